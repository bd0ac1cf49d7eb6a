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

import ctypes as C
import threading
import time
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from ._capi import HIT_DTYPE, STATUS_UNCERTIFIED, check, lib, ptr
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
    with enc.reading():
        _scan_locked(service, enc, Q, order, spans, configs, results)
    return results


def _scan_locked(service, enc, Q, order, spans, configs, results):
    g = enc.gallery
    idx, dist, cnt, st = g.match(Q, service.tolerance, MAX_NATIVE_K)
    dense = None
    for cam_id in order:
        lo, hi = spans[cam_id]
        thr = (configs.get(cam_id) or {}).get("confidence_threshold", DEFAULT_CONFIDENCE_THRESHOLD)
        out = results[cam_id]
        for b in range(lo, hi):
            # more matches than the native k cap, or an uncertifiable tie block: dense scores
            if cnt[b] >= MAX_NATIVE_K or st[b] == STATUS_UNCERTIFIED:
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


class DeviceCooldown:
    """The cooldown table of ``CooldownFilter`` held in HBM (frp_cooldown_*): keyed by (global
    gallery row, camera id), applied to the match results before they leave the device, so a
    detection the tracking service would drop as a duplicate (tracking_service.py:118-134) or
    that fails the camera's ``confidence_threshold`` (routes/camera.py:250) is never copied
    back or turned into a Python dict.  Camera ids must be integers in [0, 2^24)."""

    def __init__(self, device: int = 0, capacity: int = 1 << 20, cooldown_seconds: float = 10.0):
        self.cooldown = float(cooldown_seconds)
        self.device = int(device)
        h = C.c_void_p()
        check(lib.frp_cooldown_create(self.device, int(capacity), C.byref(h)))
        self._h = h.value

    def close(self):
        if getattr(self, "_h", None):
            lib.frp_cooldown_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        if not self._h:
            raise RuntimeError("cooldown table is closed")
        return self._h

    def clear(self):
        check(lib.frp_cooldown_clear(self.handle, None))

    def match_hits(self, gallery, queries, tau: float, k: int, camera, limit=None, now: Optional[float] = None,
                   flags: int = 0) -> np.ndarray:
        """frp_match_hits: match + confidence limit + cooldown on the device; returns the
        RECORDED hits as a structured array (query, rank, row, score, camera)."""
        q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
        B = q.shape[0]
        cam = np.ascontiguousarray(np.asarray(camera, dtype=np.int32).reshape(-1))
        if cam.shape[0] != B:
            raise ValueError("one camera id per query")
        lim = None if limit is None else np.ascontiguousarray(np.asarray(limit, dtype=np.float32).reshape(-1))
        hits = np.empty(B * k, HIT_DTYPE)
        n = C.c_int(0)
        check(lib.frp_match_hits(gallery.handle, self.handle, ptr(q), B, float(tau), int(k), int(flags), ptr(cam),
                                 ptr(lim), float(time.time() if now is None else now), self.cooldown, ptr(hits),
                                 C.byref(n)))
        return hits[:n.value]


def scan_cameras_recorded(service: FaceService, cooldown: DeviceCooldown, faces_by_camera: Dict[Any, Sequence[np.ndarray]],
                          configs: Optional[Dict[Any, Dict[str, Any]]] = None, now: Optional[float] = None,
                          top_k: int = MAX_NATIVE_K) -> Dict[Any, List[Dict[str, Any]]]:
    """``scan_cameras`` followed by ``CooldownFilter.filter`` in ONE native call: only the
    detections ``tracking_service.record_detection`` would record come back from the device.
    Camera ids must be ints.  Queries with ``top_k`` or more matches are truncated to the best
    ``top_k`` (the host path handles larger result sets)."""
    configs = configs or {}
    order, cams, limits, rows = [], [], [], []
    for cam_id, faces in faces_by_camera.items():
        cfg = configs.get(cam_id) or {}
        faces = list(faces)[:cfg.get("max_faces", DEFAULT_MAX_FACES)]
        order.append(cam_id)
        for f in faces:
            rows.append(np.asarray(f, dtype=np.float32).reshape(-1))
            cams.append(int(cam_id))
            limits.append(cfg.get("confidence_threshold", DEFAULT_CONFIDENCE_THRESHOLD))
    results: Dict[Any, List[Dict[str, Any]]] = {cam_id: [] for cam_id in order}
    enc = service.encodings
    if not rows or len(enc) == 0:
        return results
    with enc.reading():
        hits = cooldown.match_hits(enc.gallery, np.stack(rows), service.tolerance, top_k, cams, limits, now)
        for h in hits:
            d = float(h["score"])
            results[int(h["camera"])].append({"camera_id": int(h["camera"]), "target": enc.name_of(int(h["row"])),
                                              "distance": d, "confidence": service._get_confidence_level(d)})
    return results


class BatchingMatcher:
    """Micro-batching front end for concurrent callers.  The reference runs
    ``process_camera_sync`` on a 4-thread pool (routes/camera.py:29, :277-279) and every thread
    calls ``compare_faces`` per face (:243-259): four independent 1-query launches of the whole
    pipeline.  Here callers enqueue their query; the first one to find the queue empty becomes
    the leader, waits ``window_s`` for the others, runs ONE ``frp_match`` for everything queued
    with the same parameters and hands each caller its own rows.  Results are identical to
    serial calls (queries are independent inside a batch); kernel launches per query fall by
    the batch size."""

    def __init__(self, service: FaceService, window_s: float = 200e-6, max_batch: int = 256):
        self.service = service
        self.window_s = float(window_s)
        self.max_batch = int(max_batch)
        self._lock = threading.Lock()
        self._queue: List[dict] = []
        self._leader_active = False
        self.batches = 0
        self.queries = 0

    def match_topk(self, query, threshold: Optional[float] = None, top_k: int = 5, thresholded: bool = True):
        """One query -> [(target, value), ...] like ``FaceService.match_topk(query[None])[0]``."""
        return self._submit(("topk", threshold, int(top_k), bool(thresholded)), query)

    def camera_matches(self, face_encoding, camera_id=None, confidence_threshold: float = DEFAULT_CONFIDENCE_THRESHOLD):
        """The per-face body of the reference's camera loop (routes/camera.py:243-259): every
        gallery row with ``match and distance <= confidence_threshold``, ascending distance."""
        pairs = self._submit(("topk", None, MAX_NATIVE_K, True), face_encoding)
        svc = self.service
        if len(pairs) >= MAX_NATIVE_K:      # possibly more matches than the native cap: exact host path
            return [{"camera_id": camera_id, "target": r["target"], "distance": r["distance"],
                     "confidence": r["confidence"]}
                    for r in svc.compare_faces(face_encoding) if r["match"] and r["distance"] <= confidence_threshold]
        return [{"camera_id": camera_id, "target": t, "distance": d, "confidence": svc._get_confidence_level(d)}
                for t, d in pairs if d <= confidence_threshold]

    def _submit(self, key, query):
        req = {"key": key, "q": np.asarray(query, dtype=np.float32).reshape(-1), "ev": threading.Event(),
               "out": None, "err": None}
        with self._lock:
            self._queue.append(req)
            lead = not self._leader_active
            if lead:
                self._leader_active = True
        if lead:
            try:
                if self.window_s > 0:
                    time.sleep(self.window_s)       # releases the GIL: the other workers enqueue meanwhile
                while True:
                    with self._lock:
                        batch = self._queue[:self.max_batch]
                        self._queue = self._queue[self.max_batch:]
                        if not batch:
                            self._leader_active = False
                            break
                    self._run(batch)
            except BaseException:
                with self._lock:
                    self._leader_active = False
                raise
        req["ev"].wait()
        if req["err"] is not None:
            raise req["err"]
        return req["out"]

    def _run(self, batch):
        groups: Dict[tuple, List[dict]] = {}
        for r in batch:
            groups.setdefault(r["key"], []).append(r)
        for key, reqs in groups.items():
            try:
                _, threshold, top_k, thresholded = key
                outs = self.service.match_topk(np.stack([r["q"] for r in reqs]), threshold, top_k, thresholded)
                for r, o in zip(reqs, outs):
                    r["out"] = o
                self.batches += 1
                self.queries += len(reqs)
            except Exception as e:      # every caller of the failed batch sees the error
                for r in reqs:
                    r["err"] = e
            finally:
                for r in reqs:
                    r["ev"].set()
