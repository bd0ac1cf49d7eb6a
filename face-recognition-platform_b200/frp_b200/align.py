"""Batched 5-landmark alignment (host mirror of insightface ``face_align``).

``norm_crop_batch`` = for every detection ``face_align.norm_crop(img, kps, 112)``
followed by ``cv2.dnn.blobFromImages(1/127.5, mean 127.5, swapRB=True)`` and the
fp16 cast — the preprocessing insightface's ``ArcFaceONNX.get`` runs in front of
the backbone (third-party; reference call site
backend/app/utils/deepfake_utils.py:138).  One CUDA kernel does all of it.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from ._capi import check, lib, ptr

OUT = 112


def estimate_norm(lmk, precision: str = "float64") -> np.ndarray:
    """face_align.estimate_norm for a batch: [n,5,2] fp32 -> [n,2,3] float64.
    precision="float64": skimage's Umeyama in float64 (what ``norm_crop_batch``
    uses when given landmarks).  precision="float32": the mixed precision insightface really
    runs (float32 landmarks + float32 template keep skimage's means / covariance / variance in
    float32); pass the result as ``M`` to reproduce insightface's crops within 1 LSB."""
    a = np.ascontiguousarray(np.asarray(lmk, dtype=np.float32).reshape(-1, 5, 2))
    M = np.empty((a.shape[0], 6), np.float64)
    if precision == "float64":
        check(lib.frp_estimate_norm(ptr(a), a.shape[0], ptr(M)))
    elif precision == "float32":
        check(lib.frp_estimate_norm_f32(ptr(a), a.shape[0], ptr(M)))
    else:
        raise ValueError("precision must be 'float64' or 'float32'")
    return M.reshape(-1, 2, 3)


def norm_crop_batch(frames, frame_idx, lmk=None, M=None, device: int = 0,
                    return_aligned: bool = False) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Host-buffer path.  frames [F,H,W,3] uint8 BGR; frame_idx [n] int32;
    lmk [n,5,2] fp32 (or M [n,2,3] float64).  Returns (blob fp16 [n,3,112,112],
    aligned uint8 [n,112,112,3] or None)."""
    fr = np.asarray(frames)
    if fr.ndim == 3:
        fr = fr[None]
    if fr.dtype != np.uint8 or fr.shape[-1] != 3:
        raise ValueError("frames must be uint8 [F,H,W,3]")
    fr = np.ascontiguousarray(fr)
    F, H, W, _ = fr.shape
    fi = np.ascontiguousarray(np.asarray(frame_idx, dtype=np.int32).reshape(-1))
    n = fi.shape[0]
    if n and (fi.min() < 0 or fi.max() >= F):
        raise ValueError("frame_idx out of range")
    lm = Mm = None
    if M is not None:
        Mm = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(n, 6))
    elif lmk is not None:
        lm = np.ascontiguousarray(np.asarray(lmk, dtype=np.float32).reshape(n, 5, 2))
    else:
        raise ValueError("need landmarks or matrices")
    blob = np.empty((n, 3, OUT, OUT), np.float16)
    aligned = np.empty((n, OUT, OUT, 3), np.uint8) if return_aligned else None
    check(lib.frp_align(int(device), ptr(fr), F, H, W, 3 * W, ptr(fi), ptr(lm), ptr(Mm), n, ptr(blob),
                        ptr(aligned)))
    return blob, aligned


def _only_112(image_size: int, mode: str) -> None:
    if int(image_size) != OUT or mode != "arcface":
        raise ValueError("only insightface's arcface template at image_size=112 is built (got %r, %r)"
                         % (image_size, mode))


def norm_crop(img, landmark, image_size: int = OUT, mode: str = "arcface") -> np.ndarray:
    """insightface ``face_align.norm_crop(img, landmark, image_size=112, mode='arcface')``:
    one BGR frame + its 5x2 landmarks -> the aligned uint8 [112,112,3] crop."""
    _only_112(image_size, mode)
    return norm_crop_batch(np.asarray(img)[None], [0], lmk=np.asarray(landmark)[None], return_aligned=True)[1][0]


def norm_crop2(img, landmark, image_size: int = OUT, mode: str = "arcface"):
    """insightface ``face_align.norm_crop2``: (aligned crop, the 2x3 matrix that produced it)."""
    _only_112(image_size, mode)
    M = estimate_norm(np.asarray(landmark)[None])[0]
    return norm_crop_batch(np.asarray(img)[None], [0], M=M[None], return_aligned=True)[1][0], M


def norm_crop_batch_device(frames_dev, F: int, H: int, W: int, stride: int, frame_idx_dev, lmk_dev, n: int,
                           out_blob_dev, out_aligned_dev=None, M_dev=None, device: Optional[int] = None,
                           stream: int = 0) -> None:
    """Device-buffer path (torch tensors or raw pointers); asynchronous on ``stream``.
    ``device`` defaults to the device of ``frames_dev`` when that is a CUDA tensor, else 0."""
    if device is None:
        dev = getattr(frames_dev, "device", None)
        device = dev.index if (dev is not None and getattr(dev, "index", None) is not None) else 0
    check(lib.frp_align_device(int(device), ptr(frames_dev), int(F), int(H), int(W), int(stride),
                               ptr(frame_idx_dev), ptr(lmk_dev), ptr(M_dev), int(n), ptr(out_blob_dev),
                               ptr(out_aligned_dev), stream or None))
