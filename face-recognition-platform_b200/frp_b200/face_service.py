"""Drop-in ``FaceService`` for the recognition hot path.

Mirrors the public surface of the reference's
``backend/app/services/face_service.py`` for the 1:N path — same method names,
argument meaning, result dictionaries and never-raise error convention — with
the arithmetic done by libfrp_b200.so on the HBM-resident ``GalleryDict``:

=====================  =======================================  ==================
method                 reference                                 native entry point
=====================  =======================================  ==================
compare_faces          face_service.py:395-443                   frp_scores
batch_compare_faces    face_service.py:448-481                   frp_match
find_k_nearest         face_service.py:590-612                   frp_match (no thr.)
store_face (dup scan)  face_service.py:344-390 (:352-364)        frp_match (strict)
delete_face            face_service.py:517-547                   frp_gallery_remove
cluster_faces          face_service.py:552-585                   frp_cluster
get_all_targets        face_service.py:511-512                   —
match_topk (new)       routes/face.py:688-693 threshold + top-k  frp_match
=====================  =======================================  ==================

Encoding (dlib / ONNX backbones), Mongo persistence and backups are out of scope
(SURVEY.md section 8); ``store_face`` / ``delete_face`` take optional persistence
callbacks instead.
"""
from __future__ import annotations

import logging
import math
import os
import threading
import time
from collections import deque
from datetime import datetime
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from . import _capi
from ._capi import (MATCH_NO_THRESHOLD, MATCH_STRICT, METRIC_COS, METRIC_L2, STATUS_UNCERTIFIED, FrpError)
from .gallery import GalleryDict

logger = logging.getLogger("frp_b200.face_service")

DEFAULT_TOLERANCE = float(os.getenv("FACE_TOLERANCE", "0.6"))   # face_service.py:43
MAX_NATIVE_K = 64


class FaceService:
    """Service for 1:N face matching against the HBM gallery."""

    def __init__(self, encodings: Optional[GalleryDict] = None, tolerance: float = DEFAULT_TOLERANCE,
                 raise_errors: Optional[bool] = None,
                 persist: Optional[Callable[[str, list], bool]] = None,
                 unpersist: Optional[Callable[[str], bool]] = None):
        self.encodings = encodings if encodings is not None else GalleryDict(METRIC_L2)
        self.tolerance = tolerance
        self.raise_errors = (os.getenv("FRP_RAISE", "0") == "1") if raise_errors is None else raise_errors
        self._persist = persist
        self._unpersist = unpersist
        self._metrics_lock = threading.RLock()
        self._metrics = {"total_comparisons": 0, "cumulative_comparison_time": 0.0}
        self._comparison_history = deque(maxlen=5000)

    # ------------------------------------------------------------------ helpers
    @property
    def _cos(self) -> bool:
        return self.encodings.metric == METRIC_COS

    def _fail(self, what: str, exc: Exception, default):
        logger.exception("Error in %s: %s", what, exc)
        if self.raise_errors:
            raise exc
        return default

    def _get_confidence_level(self, distance: float) -> str:
        # face_service.py:486-492
        if distance < 0.4:
            return "high"
        if distance < 0.6:
            return "medium"
        return "low"

    def _calibrate_confidence(self, distance: float) -> float:
        # face_service.py:497-506: logistic in (1 - d) clipped to [0, 1], slope 12, 2 decimals
        x = min(1.0, max(0.0, 1.0 - distance))
        return round(float(100.0 / (1.0 + np.exp(-12.0 * (x - 0.5)))), 2)

    def _value_key(self) -> str:
        return "similarity" if self._cos else "distance"

    def _is_match(self, v: np.ndarray, tol: float) -> np.ndarray:
        return (v >= tol) if self._cos else (v <= tol)

    def _subset(self, target_names):
        """(names, rows) to check, in the order the reference builds them (face_service.py:403)."""
        enc = self.encodings
        if target_names is None:
            names = enc.keys()
        else:
            names = [t for t in target_names if t in enc]
        rows = np.array([enc.row_of(t) for t in names], dtype=np.int64)
        return names, rows

    # ---------------------------------------------------------------- compare
    def compare_faces(self, test_encoding: np.ndarray, target_names: Optional[List[str]] = None,
                      return_distances: bool = True) -> List[Dict[str, Any]]:
        start_time = time.time()
        try:
            # one consistent view of (device gallery, row numbering): a concurrent compaction
            # or clear waits until the scores are back (and vice versa)
            with self.encodings.reading():
                names, rows = self._subset(target_names)
                if not names:
                    logger.warning("No targets available for comparison")
                    return []
                q = np.asarray(test_encoding, dtype=np.float32).reshape(1, -1)
                all_v = self.encodings.gallery.scores(q)[0]
                v = all_v[rows]
            matches = self._is_match(v, self.tolerance)
            key = self._value_key()
            now = datetime.now().isoformat()
            vals, ms = v.tolist(), matches.tolist()
            self._comparison_history.extend({key: vals[i], "match": ms[i], "timestamp": now}
                                            for i in range(len(names)))
            if not return_distances:
                results = [{"target": names[i], "match": ms[i]} for i in range(len(names))]
            else:
                # stable argsort == the reference's stable list.sort: ties keep dict
                # insertion order (face_service.py:431)
                order = np.argsort(-v if self._cos else v, kind="stable").tolist()
                if self._cos:
                    results = [{"target": names[i], "match": ms[i], key: vals[i]} for i in order]
                else:
                    # vectorised _get_confidence_level / _calibrate_confidence (same formulas)
                    lv = np.where(v < 0.4, 0, np.where(v < 0.6, 1, 2)).tolist()
                    x = np.clip(1.0 - v, 0.0, 1.0)
                    cal = (100.0 / (1.0 + np.exp(-12.0 * (x - 0.5)))).tolist()
                    lvn = ("high", "medium", "low")
                    results = [{"target": names[i], "match": ms[i], key: vals[i], "confidence": lvn[lv[i]],
                                "confidence_score": round(cal[i], 2)} for i in order]
            dt = time.time() - start_time
            with self._metrics_lock:
                self._metrics["total_comparisons"] += len(names)
                self._metrics["cumulative_comparison_time"] += dt
            return results
        except Exception as e:  # reference convention: never raise (face_service.py:441-443)
            return self._fail("compare_faces", e, [])

    def batch_compare_faces(self, test_encodings: List[np.ndarray],
                            target_names: Optional[List[str]] = None) -> List[List[Dict[str, Any]]]:
        try:
            with self.encodings.reading():
                return self._batch_compare_locked(test_encodings, target_names)
        except Exception as e:
            return self._fail("batch_compare_faces", e, [[] for _ in test_encodings])

    def _native_matches(self, Q: np.ndarray, tol: float, flags: int = 0):
        """Threshold-then-top-64 on the device for every query (tcgen05 kernel for more than 4
        queries).  Returns (idx, score, count, dense_needed[B]): a query needs the dense
        all-scores path when its list is full (there may be more than 64 matches) or when the
        native ladder could not certify it (status 2: more than 64 exact ties at the cut)."""
        idx, sc, cnt, st = self.encodings.gallery.match(Q, tol, MAX_NATIVE_K, flags)
        return idx, sc, cnt, (cnt >= MAX_NATIVE_K) | (st == STATUS_UNCERTIFIED)

    def _batch_compare_locked(self, test_encodings, target_names):
        names, rows = self._subset(target_names)
        if not names:
            return [[] for _ in test_encodings]
        if len(test_encodings) == 0:
            return []
        Q = np.stack([np.asarray(t, dtype=np.float32).reshape(-1) for t in test_encodings])
        key = self._value_key()
        out: List[List[Dict[str, Any]]] = []
        if target_names is None:
            g = self.encodings.gallery
            idx, sc, cnt, need_dense = self._native_matches(Q, self.tolerance)
            dense = None
            for b in range(len(Q)):
                if need_dense[b]:
                    if dense is None:
                        dense = g.scores(Q)
                    out.append(self._matches_from_dense(dense[b], names, rows))
                    continue
                res = []
                for j in range(int(cnt[b])):
                    val = float(sc[b, j])
                    item = {"target": self.encodings.name_of(int(idx[b, j])), "match": True, key: val}
                    if not self._cos:
                        item["confidence"] = self._get_confidence_level(val)
                    res.append(item)
                out.append(res)
        else:
            dense = self.encodings.gallery.scores(Q)
            for b in range(len(Q)):
                out.append(self._matches_from_dense(dense[b], names, rows))
        return out

    def _matches_from_dense(self, all_v, names, rows):
        key = self._value_key()
        v = all_v[rows]
        m = self._is_match(v, self.tolerance)
        res = []
        for i in np.nonzero(m)[0]:
            val = float(v[i])
            item = {"target": names[int(i)], "match": True, key: val}
            if not self._cos:
                item["confidence"] = self._get_confidence_level(val)
            res.append(item)
        res.sort(key=(lambda x: -x[key]) if self._cos else (lambda x: x[key]))
        return res

    def find_k_nearest(self, test_encoding: np.ndarray, k: int = 5) -> List[Dict[str, Any]]:
        try:
            with self.encodings.reading():
                return self._find_k_nearest_locked(test_encoding, k)
        except Exception as e:
            return self._fail("find_k_nearest", e, [])

    def _find_k_nearest_locked(self, test_encoding, k):
        if len(self.encodings) == 0:
            return []
        k = min(int(k), len(self.encodings))
        q = np.asarray(test_encoding, dtype=np.float32).reshape(1, -1)
        key = self._value_key()
        g = self.encodings.gallery
        pairs = None
        if k <= MAX_NATIVE_K:
            idx, sc, cnt, st = g.match(q, 0.0, k, MATCH_NO_THRESHOLD)
            if st[0] != STATUS_UNCERTIFIED:
                pairs = [(int(idx[0, j]), float(sc[0, j])) for j in range(int(cnt[0]))]
        if pairs is None:
            v = g.scores(q)[0]
            rows = self.encodings.live_rows()
            order = np.argsort(-v[rows] if self._cos else v[rows], kind="stable")[:k]
            pairs = [(int(rows[i]), float(v[rows[i]])) for i in order]
        results = []
        for row, val in pairs:
            item = {"target": self.encodings.name_of(row), key: val}
            if not self._cos:
                item["confidence"] = self._get_confidence_level(val)
                item["confidence_score"] = self._calibrate_confidence(val)
            results.append(item)
        return results

    def match_topk(self, queries, threshold: Optional[float] = None, top_k: int = 5,
                   thresholded: bool = True):
        """Threshold-then-top-k for a batch of queries (the `/face/compare` and
        `/face/similar` post-processing, routes/face.py:685-693, :800-809) without
        materialising N dicts.  Returns one list of (target, value) per query."""
        tol = self.tolerance if threshold is None else threshold
        Q = np.atleast_2d(np.asarray(queries, dtype=np.float32))
        kk = min(int(top_k), MAX_NATIVE_K)
        flags = 0 if thresholded else MATCH_NO_THRESHOLD
        with self.encodings.reading():
            enc = self.encodings
            if len(enc) == 0:
                return [[] for _ in Q]
            g = enc.gallery
            idx, sc, cnt, st = g.match(Q, tol, kk, flags)
            out = []
            dense = None
            for b in range(len(Q)):
                if st[b] == STATUS_UNCERTIFIED:     # > 64 exact ties at the cut: dense scores decide
                    if dense is None:
                        dense = g.scores(Q)
                    rows = enc.live_rows()
                    v = dense[b][rows]
                    keep = np.nonzero(self._is_match(v, tol))[0] if thresholded else np.arange(len(rows))
                    order = keep[np.argsort(-v[keep] if self._cos else v[keep], kind="stable")][:kk]
                    out.append([(enc.name_of(int(rows[i])), float(v[i])) for i in order])
                else:
                    out.append([(enc.name_of(int(idx[b, j])), float(sc[b, j])) for j in range(int(cnt[b]))])
            return out

    # ------------------------------------------------------------------ mutate
    def get_all_targets(self) -> List[str]:
        return self.encodings.keys()

    def store_face(self, target_name: str, encoding: np.ndarray) -> Dict[str, Any]:
        try:
            encoding_list = encoding.tolist() if isinstance(encoding, np.ndarray) else encoding
            is_duplicate, similar_target = False, None
            enc = self.encodings
            if len(enc) and not self._cos:
                # first existing face (dict order) closer than 0.3, self excluded (face_service.py:352-364)
                q = np.asarray(encoding_list, dtype=np.float32).reshape(1, -1)
                with enc.reading():
                    idx, sc, cnt, need_dense = self._native_matches(q, 0.3, MATCH_STRICT)
                    if need_dense[0]:
                        v = enc.gallery.scores(q)[0]
                        cand = [r for r in enc.live_rows() if v[r] < 0.3]
                    else:
                        cand = sorted(int(i) for i in idx[0, :cnt[0]])
                    for r in cand:
                        name = enc.name_of(int(r))
                        if name is not None and name != target_name:
                            is_duplicate, similar_target = True, name
                            logger.warning("Potential duplicate: %s ~ %s", target_name, name)
                            break
            already_exists = target_name in enc
            if self._persist is not None and not self._persist(target_name, encoding_list):
                return {"success": False, "message": "Failed to store in database", "is_duplicate": is_duplicate}
            enc[target_name] = encoding_list
            message = "Face %s successfully for '%s'" % ("updated" if already_exists else "stored", target_name)
            if is_duplicate:
                message += " (Warning: Similar to '%s')" % similar_target
            return {"success": True, "message": message, "is_duplicate": is_duplicate,
                    "similar_to": similar_target if is_duplicate else None, "was_update": already_exists}
        except Exception as e:
            if self.raise_errors:
                raise
            logger.exception("Error storing face %s: %s", target_name, e)
            return {"success": False, "message": "Error storing face: %s" % e, "is_duplicate": False}

    def delete_face(self, target_name: str) -> Dict[str, Any]:
        try:
            removed_from_memory = False
            if target_name in self.encodings:
                del self.encodings[target_name]
                removed_from_memory = True
            removed_from_db = bool(self._unpersist(target_name)) if self._unpersist else False
            if removed_from_memory or removed_from_db:
                return {"success": True, "message": "Face '%s' deleted successfully" % target_name,
                        "removed_from_memory": removed_from_memory, "removed_from_db": removed_from_db}
            return {"success": False, "message": "Face '%s' not found in database or memory" % target_name}
        except Exception as e:
            if self.raise_errors:
                raise
            logger.exception("Error deleting face %s: %s", target_name, e)
            return {"success": False, "message": "Error deleting face: %s" % e}

    def cluster_faces(self, distance_threshold: float = 0.6) -> Dict[str, List[str]]:
        """Greedy seed clustering of face_service.py:552-585: every not-yet-assigned
        face (dict order) seeds a cluster and absorbs all unassigned faces within
        the threshold (inclusive).  The gallery is matched against itself on the device
        (frp_cluster: 256-row query slabs through the tcgen05 matcher, thresholded neighbour
        lists only); a seed always precedes its members in dict order, so grouping the labels
        in row order reproduces the reference's member order."""
        try:
            with self.encodings.reading():
                enc = self.encodings
                if len(enc) < 2:
                    return {"cluster_0": enc.keys()}
                labels, n, _ = enc.gallery.cluster(distance_threshold)
                clusters: Dict[str, List[str]] = {"cluster_%d" % c: [] for c in range(n)}
                for name, row in enc._rows.items():          # dict order == row order
                    clusters["cluster_%d" % int(labels[row])].append(name)
                return clusters
        except Exception as e:
            return self._fail("cluster_faces", e, {})

    # ----------------------------------------------------------------- metrics
    def get_performance_metrics(self) -> Dict[str, Any]:
        with self._metrics_lock:
            m = dict(self._metrics)
        total = m["total_comparisons"]
        avg = (m["cumulative_comparison_time"] / total) if total > 0 else 0.0
        return {**m, "average_comparison_time": round(avg, 4), "total_faces_stored": len(self.encodings),
                "comparison_history_size": len(self._comparison_history),
                "kernel_launches": _capi.kernel_launches()}
