"""Batched camera-scan front end.

The reference matches one face at a time on a 4-thread pool:
``process_camera_sync`` (backend/app/routes/camera.py:171-272) loops
``for face_encoding in faces_enc: face_service.compare_faces(...)`` (:243-259) and keeps
every gallery row with ``match and distance <= confidence_threshold`` (:250);
``/camera/alerts`` (:284-391) fans that out over the cameras.  Here all faces of all
cameras of one scan go through ONE ``frp_match`` call; the per-camera result lists are
identical to the reference's (same dictionaries, same order: faces in detection order,
matches by ascending distance with ties in gallery order).

Detection and encoding (dlib / ONNX) are out of scope: the caller passes the
encodings each camera produced for the current frame.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .face_service import MAX_NATIVE_K, FaceService

DEFAULT_CONFIDENCE_THRESHOLD = 0.6   # routes/camera.py:180
DEFAULT_MAX_FACES = 10               # routes/camera.py:182


class CooldownFilter:
    """Hand-off filter between the matcher and the tracking service:
    the reference drops a (person, camera) detection that arrives within the cooldown window
    of the previous recorded one (backend/app/services/tracking_service.py:122-134, window
    set at :41,57, timestamp stored at :179).  Applying it right after the batched match
    keeps repeated hits out of the downstream Python services; state is host-side, like the
    reference's ``last_detection`` dict."""

    def __init__(self, cooldown_seconds: float = 10.0):
        self.cooldown = float(cooldown_seconds)
        self.last_detection: Dict[tuple, float] = {}

    def filter(self, results: Dict[Any, List[Dict[str, Any]]], now: float) -> Dict[Any, List[Dict[str, Any]]]:
        """``results`` = output of ``scan_cameras``; ``now`` = scan timestamp in seconds.
        Returns the detections the reference's ``record_detection`` would record."""
        out: Dict[Any, List[Dict[str, Any]]] = {}
        for cam_id, dets in results.items():
            kept = []
            for d in dets:
                key = (d["target"], cam_id)
                last = self.last_detection.get(key)
                if last is not None and (now - last) < self.cooldown:
                    continue                      # duplicate within the cooldown: not recorded
                self.last_detection[key] = now
                kept.append(d)
            out[cam_id] = kept
        return out


def scan_cameras(service: FaceService, faces_by_camera: Dict[Any, Sequence[np.ndarray]],
                 configs: Optional[Dict[Any, Dict[str, Any]]] = None) -> Dict[Any, List[Dict[str, Any]]]:
    """One scan over all cameras.  ``faces_by_camera[cam_id]`` = the face encodings
    detected in that camera's current frame (detection order).  ``configs[cam_id]`` may
    carry ``confidence_threshold`` and ``max_faces`` like the reference's per-camera
    config.  Returns ``{cam_id: [{"camera_id","target","distance","confidence"}, ...]}``."""
    configs = configs or {}
    order: List[Any] = []
    spans: Dict[Any, tuple] = {}
    rows: List[np.ndarray] = []
    for cam_id, faces in faces_by_camera.items():
        cfg = configs.get(cam_id) or {}
        max_faces = cfg.get("max_faces", DEFAULT_MAX_FACES)
        faces = list(faces)[:max_faces]                      # routes/camera.py:233-235
        spans[cam_id] = (len(rows), len(rows) + len(faces))
        order.append(cam_id)
        rows.extend(np.asarray(f, dtype=np.float32).reshape(-1) for f in faces)
    results: Dict[Any, List[Dict[str, Any]]] = {cam_id: [] for cam_id in order}
    enc = service.encodings
    if not rows or len(enc) == 0:
        return results
    Q = np.stack(rows)
    g = enc.gallery
    idx, dist, cnt, _ = g.match(Q, service.tolerance, MAX_NATIVE_K)
    dense = None
    for cam_id in order:
        lo, hi = spans[cam_id]
        thr = (configs.get(cam_id) or {}).get("confidence_threshold", DEFAULT_CONFIDENCE_THRESHOLD)
        out = results[cam_id]
        for b in range(lo, hi):
            if cnt[b] >= MAX_NATIVE_K:      # more matches than the native k cap: dense scores
                if dense is None:
                    dense = g.scores(Q)
                live = enc.live_rows()
                d = dense[b][live]
                keep = np.nonzero(d <= service.tolerance)[0]
                keep = keep[np.argsort(d[keep], kind="stable")]
                pairs = [(int(live[i]), float(d[i])) for i in keep]
            else:
                pairs = [(int(idx[b, j]), float(dist[b, j])) for j in range(int(cnt[b]))]
            for row, dd in pairs:
                if dd <= thr:
                    out.append({"camera_id": cam_id, "target": enc.name_of(row), "distance": dd,
                                "confidence": service._get_confidence_level(dd)})
    return results
