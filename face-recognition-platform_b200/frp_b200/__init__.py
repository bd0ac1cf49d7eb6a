"""frp_b200 — B200-native recognition hot path (5-landmark alignment + 1:N matching).

Host-side mirror of the reference's ``face_service`` / ``ENCODINGS`` seam over the
C ABI of ``libfrp_b200.so`` (``include/frp_b200.h``).  Importing this package loads
the shared library and fails loudly if it has not been built; there is no CPU
fallback.
"""
from . import _capi
from ._capi import (FrpError, METRIC_COS, METRIC_L2, MATCH_FORCE_EXACT, MATCH_FORCE_GEMM,
                    MATCH_FORCE_SCAN, MATCH_NO_THRESHOLD, MATCH_STRICT, STATUS_FALLBACK, STATUS_OK,
                    STATUS_UNCERTIFIED, device_count, kernel_launches, last_match_path)
from .gallery import Gallery, GalleryDict
from .face_service import FaceService
from . import align
from .camera_scan import BatchingMatcher, CooldownFilter, DeviceCooldown, scan_cameras, scan_cameras_recorded
from .shard import CertifiedPipeline, ShardedGalleryDict, ShardedMatcher, shard_bounds, shard_streams

__all__ = [
    "FrpError", "METRIC_COS", "METRIC_L2", "MATCH_FORCE_EXACT", "MATCH_FORCE_GEMM", "MATCH_FORCE_SCAN",
    "MATCH_NO_THRESHOLD", "MATCH_STRICT", "STATUS_OK", "STATUS_FALLBACK", "STATUS_UNCERTIFIED",
    "device_count", "kernel_launches", "last_match_path", "Gallery", "GalleryDict", "FaceService",
    "align", "ShardedMatcher", "ShardedGalleryDict", "CertifiedPipeline", "shard_bounds", "shard_streams",
    "scan_cameras", "scan_cameras_recorded", "BatchingMatcher", "CooldownFilter", "DeviceCooldown",
]
