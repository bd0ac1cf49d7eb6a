"""ctypes binding of libfrp_b200.so (include/frp_b200.h).

There is no CPU fallback: if the shared library is missing the import fails
loudly; if no sm_100 device is visible every compute call raises FrpError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libfrp_b200.so")

FRP_OK = 0
FRP_E_INVALID, FRP_E_CUDA, FRP_E_NODEVICE, FRP_E_NOMEM, FRP_E_IO, FRP_E_UNSUPPORTED, FRP_E_RANGE = \
    -1, -2, -3, -4, -5, -6, -7
METRIC_L2, METRIC_COS = 0, 1
MATCH_THRESHOLD, MATCH_NO_THRESHOLD, MATCH_FORCE_SCAN, MATCH_FORCE_GEMM, MATCH_FORCE_EXACT, MATCH_STRICT = \
    0x0, 0x1, 0x2, 0x4, 0x8, 0x10
MATCH_WIDE = 0x20
MATCH_NO_SAMPLE = 0x40
STATUS_OK, STATUS_FALLBACK, STATUS_UNCERTIFIED = 0, 1, 2


class FrpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("frp_b200 error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.isfile(_LIB_PATH):
        raise ImportError(
            "libfrp_b200.so not found at %s — build it with "
            "`python face-recognition-platform_b200/build.py` (there is no CPU fallback)" % _LIB_PATH)
    return C.CDLL(_LIB_PATH)


lib = _load()

_vp, _i, _u64, _d, _sz = C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_size_t
_pi, _pu64 = C.POINTER(C.c_int), C.POINTER(C.c_uint64)

# name -> (restype, argtypes); also the list the "exports every symbol" test checks
SIGNATURES = {
    "frp_abi_version": (_i, []),
    "frp_last_error": (C.c_char_p, []),
    "frp_device_count": (_i, [_pi]),
    "frp_device_info": (_i, [_i, C.c_char_p, _sz, _pi, _pu64, _pi, _pi]),
    "frp_gallery_create": (_i, [_i, _i, _i, _u64, C.POINTER(_vp)]),
    "frp_gallery_destroy": (_i, [_vp]),
    "frp_gallery_append_f32": (_i, [_vp, _vp, _u64, _pu64]),
    "frp_gallery_append_f64": (_i, [_vp, _vp, _u64, _pu64]),
    "frp_gallery_append_device": (_i, [_vp, _vp, _u64, _pu64, _vp]),
    "frp_gallery_update_f32": (_i, [_vp, _u64, _vp]),
    "frp_gallery_update_f64": (_i, [_vp, _u64, _vp]),
    "frp_gallery_remove": (_i, [_vp, _u64]),
    "frp_gallery_count": (_i, [_vp, _pu64, _pu64]),
    "frp_gallery_info": (_i, [_vp, _pi, _pi, _pi, _pi, _pu64]),
    "frp_gallery_gather": (_i, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "frp_gallery_get_row": (_i, [_vp, _u64, _vp]),
    "frp_gallery_save": (_i, [_vp, C.c_char_p]),
    "frp_gallery_load": (_i, [_i, C.c_char_p, C.POINTER(_vp)]),
    "frp_gallery_set_base": (_i, [_vp, _u64]),
    "frp_gallery_set_seq": (_i, [_vp, _u64, _u64, _vp]),
    "frp_match": (_i, [_vp, _vp, _i, _d, _i, _i, _vp, _vp, _vp, _vp]),
    "frp_match_device": (_i, [_vp, _vp, _i, _d, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "frp_scores": (_i, [_vp, _vp, _i, _vp, _u64, _pu64]),
    "frp_scores_device": (_i, [_vp, _vp, _i, _vp, _u64, _pu64, _vp]),
    "frp_shard_record_bytes": (_sz, [_i, _i]),
    "frp_match_record_device": (_i, [_vp, _vp, _i, _d, _i, _i, _vp, _vp]),
    "frp_merge_packed_device": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "frp_set_reserved_sms": (_i, [_i]),
    "frp_cluster": (_i, [_vp, _d, _i, _vp, _u64, _pi, _vp]),
    "frp_cooldown_create": (_i, [_i, _u64, C.POINTER(_vp)]),
    "frp_cooldown_destroy": (_i, [_vp]),
    "frp_cooldown_clear": (_i, [_vp, _vp]),
    "frp_cooldown_filter_device": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _vp, _vp, _vp, _vp]),
    "frp_match_hits": (_i, [_vp, _vp, _vp, _i, _d, _i, _i, _vp, _vp, _d, _d, _vp, _pi]),
    "frp_exchange_buffer_bytes": (_sz, [_i, _i, _i]),
    "frp_push_merge_device": (_i, [_i, _i, _i, _i, _i, _vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "frp_estimate_norm": (_i, [_vp, _i, _vp]),
    "frp_estimate_norm_f32": (_i, [_vp, _i, _vp]),
    "frp_align": (_i, [_i, _vp, _i, _i, _i, _sz, _vp, _vp, _vp, _i, _vp, _vp]),
    "frp_align_device": (_i, [_i, _vp, _i, _i, _i, _sz, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "frp_profile_enable": (_i, [_i]),
    "frp_profile_last": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "frp_profile_align_enable": (_i, [_i]),
    "frp_profile_align_last": (_i, [C.POINTER(C.c_float)]),
    "frp_kernel_launches": (_u64, []),
    "frp_last_match_path": (C.c_char_p, []),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


# frp_hit_t (include/frp_b200.h): one recorded detection
HIT_DTYPE = np.dtype([("query", np.int32), ("rank", np.int32), ("row", np.int64), ("score", np.float32),
                      ("camera", np.int32)], align=True)
assert HIT_DTYPE.itemsize == 24


def last_error() -> str:
    return (lib.frp_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != FRP_OK:
        raise FrpError(rc, last_error())


_c_char = C.c_char


def ptr(a) -> int:
    """Raw address of a C-contiguous numpy array, a torch tensor, an int, or None.

    numpy: ``ndarray.ctypes.data`` builds a helper object per access (~1.7 us); the buffer-protocol
    route is three times cheaper, which matters for a 30 us call that passes five pointers.  It
    needs a writable, non-empty buffer; anything else takes the slow route."""
    if a is None:
        return None
    if type(a) is int:
        return a
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        try:
            return C.addressof(_c_char.from_buffer(a))
        except (TypeError, ValueError):
            return a.ctypes.data
    if isinstance(a, int):
        return int(a)
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


def device_count() -> int:
    n = C.c_int(0)
    lib.frp_device_count(C.byref(n))
    return n.value


def kernel_launches() -> int:
    return int(lib.frp_kernel_launches())


def last_match_path() -> str:
    return (lib.frp_last_match_path() or b"").decode()
