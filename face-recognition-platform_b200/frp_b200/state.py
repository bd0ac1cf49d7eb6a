"""Process-global gallery + service singletons, mirroring the reference's
``app.state.ENCODINGS`` (backend/app/state.py:78) and
``app.services.face_service.face_service`` (face_service.py:769).  The device
gallery is created lazily at the first insert, so importing this module needs no GPU."""
from .face_service import FaceService
from .gallery import GalleryDict

ENCODINGS = GalleryDict()
face_service = FaceService(ENCODINGS)
