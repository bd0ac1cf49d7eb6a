"""Import the UNMODIFIED reference ``face_service`` in the authoring container.

TEST INFRASTRUCTURE ONLY (used by ``oracle/gen_golden.py`` and by CPU tests that
are skipped when ``/root/reference`` is absent, e.g. on the GPU box).  Nothing is
copied from the reference: ``/root/reference/backend`` is put on ``sys.path`` and
its third-party imports that are missing from this image are stubbed:

* ``socketio``  (``backend/app/state.py:47`` needs ``AsyncManager``)
* ``pymongo`` / ``bson`` (``backend/app/utils/db.py:22-23``)
* ``face_recognition`` — the two functions of ``face-recognition==1.3.0`` the hot
  path calls (``face_service.py:357,410,465,576,599``), per its published source:
  ``face_distance = np.linalg.norm(encs - q, axis=1)`` and
  ``compare_faces = list(face_distance(...) <= tolerance)``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FRP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "backend/app/services/face_service.py"))


class _Anything:
    """Permissive stand-in for pymongo collections/clients."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return True


def _install_stubs():
    if "socketio" not in sys.modules:
        m = types.ModuleType("socketio")

        class AsyncManager:  # state.py:47
            def __init__(self, *a, **k):
                pass

        class AsyncServer:
            def __init__(self, *a, **k):
                pass

        m.AsyncManager = AsyncManager
        m.AsyncServer = AsyncServer
        m.ASGIApp = _Anything
        sys.modules["socketio"] = m
    if "pymongo" not in sys.modules:
        m = types.ModuleType("pymongo")
        m.MongoClient = _Anything
        m.ASCENDING, m.DESCENDING = 1, -1
        errs = types.ModuleType("pymongo.errors")
        for name in ("AutoReconnect", "PyMongoError", "ConnectionFailure",
                     "ServerSelectionTimeoutError", "DuplicateKeyError", "OperationFailure"):
            setattr(errs, name, type(name, (Exception,), {}))
        m.errors = errs
        sys.modules["pymongo"] = m
        sys.modules["pymongo.errors"] = errs
    if "bson" not in sys.modules:
        m = types.ModuleType("bson")
        m.ObjectId = type("ObjectId", (str,), {})
        sys.modules["bson"] = m
        sub = types.ModuleType("bson.objectid")
        sub.ObjectId = m.ObjectId
        sys.modules["bson.objectid"] = sub
    if "face_recognition" not in sys.modules:
        m = types.ModuleType("face_recognition")

        def face_distance(face_encodings, face_to_compare):
            if len(face_encodings) == 0:
                return np.empty((0))
            return np.linalg.norm(np.asarray(face_encodings) - face_to_compare, axis=1)

        def compare_faces(known, cand, tolerance=0.6):
            return list(face_distance(known, cand) <= tolerance)

        m.face_distance = face_distance
        m.compare_faces = compare_faces
        sys.modules["face_recognition"] = m


_cached = None


def load_reference():
    """Returns (face_service singleton, ENCODINGS dict, db module) of the
    reference.  The import creates ``data/`` and ``logs/`` relative to cwd
    (``face_service.py:47-48``), so cwd is a scratch dir during import."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    os.environ.setdefault("DISABLE_ENCRYPTION", "true")
    backend = os.path.join(REFERENCE_ROOT, "backend")
    if backend not in sys.path:
        sys.path.insert(0, backend)
    scratch = tempfile.mkdtemp(prefix="frp_ref_")
    old = os.getcwd()
    os.chdir(scratch)
    try:
        sys.dont_write_bytecode = True
        from app.services.face_service import face_service  # type: ignore
        from app.state import ENCODINGS  # type: ignore
        import app.utils.db as db  # type: ignore
        import app.services.face_service as fs_mod  # type: ignore
        # store_face persists through db.store_embedding (Mongo); keep it in memory
        fs_mod.store_embedding = lambda *a, **k: True

        class _Coll:  # face_service.py:528-529 reads result.deleted_count
            def delete_one(self, *a, **k):
                return types.SimpleNamespace(deleted_count=0)

        fs_mod.faces_collection = _Coll()
        fs_mod.BACKUP_DIR = __import__("pathlib").Path(scratch) / "backups"
        fs_mod.BACKUP_DIR.mkdir(parents=True, exist_ok=True)
    finally:
        os.chdir(old)
    _cached = (face_service, ENCODINGS, db)
    return _cached
