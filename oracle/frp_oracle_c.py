"""ctypes loader of the plain-C oracle (oracle/frp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfrp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "frp_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/libfrp_oracle.so"], check=True,
                       capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_topk.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def l2_distances(G, q):
    G = np.ascontiguousarray(G, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    out = np.empty(len(G), np.float64)
    lib().orc_l2_distances(_p(G), C.c_int64(len(G)), C.c_int(G.shape[1]), _p(q), _p(out))
    return out


def cos_scores(G, q):
    G = np.ascontiguousarray(G, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    out = np.empty(len(G), np.float64)
    lib().orc_cos_scores(_p(G), C.c_int64(len(G)), C.c_int(G.shape[1]), _p(q), _p(out))
    return out


def topk(v, larger_is_better, thresholded, tau, k):
    v = np.ascontiguousarray(v, np.float64)
    idx = np.full(k, -1, np.int64)
    val = np.full(k, np.nan, np.float64)
    n = lib().orc_topk(_p(v), C.c_int64(len(v)), C.c_int(int(larger_is_better)), C.c_int(int(thresholded)),
                       C.c_double(tau), C.c_int(k), _p(idx), _p(val))
    idx[n:] = -1
    val[n:] = np.nan
    return idx, val, n


def warp_affine(img, M, size=112):
    img = np.ascontiguousarray(img, np.uint8)
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    out = np.empty((size, size, 3), np.uint8)
    lib().orc_warp_affine(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int64(img.strides[0]),
                          _p(M), C.c_int(size), _p(out))
    return out


def blob(aligned):
    aligned = np.ascontiguousarray(aligned, np.uint8)
    size = aligned.shape[0]
    out = np.empty((3, size, size), np.float32)
    lib().orc_blob(_p(aligned), C.c_int(size), _p(out))
    return out
