"""HBM-resident gallery: ``Gallery`` (thin wrapper of ``frp_gallery_t``) and
``GalleryDict`` — the dict-compatible replacement of the reference's
``app.state.ENCODINGS`` (backend/app/state.py:78).

Mutations mirror the dict semantics the reference relies on:
* ``ENCODINGS[t] = enc.tolist()`` (face_service.py:374): new key -> appended at
  the end; existing key -> value replaced, position kept.
* ``del ENCODINGS[t]`` (face_service.py:521-523): row tombstoned.
* iteration / ``keys()`` order = insertion order = gallery row order = the tie
  order of the reference's stable sort (face_service.py:431).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Iterable, Iterator, List, Optional

import numpy as np

from . import _capi
from ._capi import FRP_E_RANGE, METRIC_COS, METRIC_L2, FrpError, check, lib, ptr


class RWLock:
    """Readers-writer lock for the gallery bookkeeping: matchers (readers) hold it across
    subset -> device call -> row->name translation, so a compaction or ``clear`` (writers) can
    never swap the device gallery or renumber rows under them.  Writer-preferring; re-entrant
    for the thread that already holds it (a writer may read, a reader may read again)."""

    def __init__(self):
        self._cond = threading.Condition(threading.Lock())
        self._readers = 0
        self._writer = None
        self._wdepth = 0
        self._waiting_writers = 0
        self._local = threading.local()

    def acquire_read(self):
        me = threading.get_ident()
        depth = getattr(self._local, "rdepth", 0)
        with self._cond:
            if self._writer == me or depth:
                self._local.rdepth = depth + 1
                return
            while self._writer is not None or self._waiting_writers:
                self._cond.wait()
            self._readers += 1
            self._local.rdepth = 1

    def release_read(self):
        me = threading.get_ident()
        with self._cond:
            self._local.rdepth -= 1
            if self._local.rdepth or self._writer == me:
                return
            self._readers -= 1
            if not self._readers:
                self._cond.notify_all()

    def acquire_write(self):
        me = threading.get_ident()
        with self._cond:
            if self._writer == me:
                self._wdepth += 1
                return
            if getattr(self._local, "rdepth", 0):
                raise RuntimeError("cannot upgrade a gallery read lock to a write lock")
            self._waiting_writers += 1
            while self._writer is not None or self._readers:
                self._cond.wait()
            self._waiting_writers -= 1
            self._writer = me
            self._wdepth = 1

    def release_write(self):
        with self._cond:
            self._wdepth -= 1
            if not self._wdepth:
                self._writer = None
                self._cond.notify_all()

    class _Ctx:
        def __init__(self, acq, rel):
            self._acq, self._rel = acq, rel

        def __enter__(self):
            self._acq()

        def __exit__(self, *exc):
            self._rel()
            return False

    def read(self):
        return RWLock._Ctx(self.acquire_read, self.release_read)

    def write(self):
        return RWLock._Ctx(self.acquire_write, self.release_write)


class Gallery:
    """Owns one ``frp_gallery_t`` on one device."""

    def __init__(self, dim: int, metric: int = METRIC_L2, device: int = 0, capacity: int = 1024,
                 _handle: Optional[int] = None):
        self.dim = int(dim)
        self.metric = int(metric)
        self.device = int(device)
        if _handle is None:
            h = C.c_void_p()
            check(lib.frp_gallery_create(self.device, self.dim, self.metric, int(capacity), C.byref(h)))
            _handle = h.value
        self._h = _handle

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.frp_gallery_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        if not self._h:
            raise RuntimeError("gallery is closed")
        return self._h

    # -- mutation -----------------------------------------------------------
    def append(self, rows) -> int:
        """Append host rows [n, dim] (float32 or float64); returns first row index."""
        a = np.asarray(rows)
        if a.ndim == 1:
            a = a[None, :]
        if a.shape[1] != self.dim:
            raise ValueError("expected rows of length %d, got %d" % (self.dim, a.shape[1]))
        first = C.c_uint64(0)
        if a.dtype == np.float32:
            a = np.ascontiguousarray(a)
            check(lib.frp_gallery_append_f32(self.handle, ptr(a), a.shape[0], C.byref(first)))
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            check(lib.frp_gallery_append_f64(self.handle, ptr(a), a.shape[0], C.byref(first)))
        return first.value

    def append_device(self, rows_dev, n: Optional[int] = None, stream: int = 0) -> int:
        """Append fp32 rows already in device memory (torch tensor or raw pointer)."""
        if hasattr(rows_dev, "shape"):
            if tuple(rows_dev.shape)[-1] != self.dim:
                raise ValueError("expected rows of length %d" % self.dim)
            n = int(rows_dev.shape[0]) if n is None else n
        first = C.c_uint64(0)
        check(lib.frp_gallery_append_device(self.handle, ptr(rows_dev), int(n), C.byref(first), stream or None))
        return first.value

    def update(self, row: int, vec) -> None:
        a = np.ascontiguousarray(np.asarray(vec, dtype=np.float64).reshape(-1))
        if a.shape[0] != self.dim:
            raise ValueError("expected a vector of length %d" % self.dim)
        check(lib.frp_gallery_update_f64(self.handle, int(row), ptr(a)))

    def remove(self, row: int) -> None:
        check(lib.frp_gallery_remove(self.handle, int(row)))

    def set_base(self, base_row: int) -> None:
        check(lib.frp_gallery_set_base(self.handle, int(base_row)))

    def set_seq(self, first_row: int, seqs) -> None:
        """Explicit global sequence numbers for rows [first_row, first_row + len(seqs))
        (mutable row shards)."""
        a = np.ascontiguousarray(np.asarray(seqs, dtype=np.uint64).reshape(-1))
        check(lib.frp_gallery_set_seq(self.handle, int(first_row), len(a), ptr(a) if len(a) else None))

    # -- queries --------------------------------------------------------------
    def counts(self):
        rows, live = C.c_uint64(0), C.c_uint64(0)
        check(lib.frp_gallery_count(self.handle, C.byref(rows), C.byref(live)))
        return rows.value, live.value

    @property
    def rows(self) -> int:
        return self.counts()[0]

    def get_row(self, row: int) -> np.ndarray:
        out = np.empty(self.dim, np.float32)
        check(lib.frp_gallery_get_row(self.handle, int(row), ptr(out)))
        return out

    def gather(self, idx: Iterable[int]) -> "Gallery":
        a = np.ascontiguousarray(np.asarray(list(idx), dtype=np.uint64))
        h = C.c_void_p()
        check(lib.frp_gallery_gather(self.handle, ptr(a) if len(a) else None, len(a), C.byref(h)))
        return Gallery(self.dim, self.metric, self.device, _handle=h.value)

    def save(self, path: str) -> None:
        check(lib.frp_gallery_save(self.handle, path.encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Gallery":
        h = C.c_void_p()
        check(lib.frp_gallery_load(int(device), path.encode(), C.byref(h)))
        dev, dim, dimp, metric = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        cap = C.c_uint64()
        check(lib.frp_gallery_info(h.value, C.byref(dev), C.byref(dim), C.byref(dimp), C.byref(metric), C.byref(cap)))
        return cls(dim.value, metric.value, dev.value, _handle=h.value)

    def match(self, queries, tau: float, k: int, flags: int = 0):
        """Host-buffer 1:N match.  Returns (idx[B,k] int64, score[B,k] f32, count[B], status[B])."""
        q = queries
        if not (type(q) is np.ndarray and q.dtype == np.float32 and q.ndim == 2 and q.flags.c_contiguous):
            q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
        if q.shape[1] != self.dim:
            raise ValueError("expected queries of length %d, got %d" % (self.dim, q.shape[1]))
        B = q.shape[0]
        k = int(k)
        # the four results are views of ONE buffer (idx | score | count | status): one address
        # look-up instead of four (the single-query call is ~30 us, a look-up costs 0.6-1.7 us)
        o_s = B * k * 8
        o_c = o_s + B * k * 4
        o_t = o_c + B * 4
        buf = np.empty(o_t + B * 4 + 8, np.uint8)
        idx = np.ndarray((B, k), np.int64, buf, 0)
        score = np.ndarray((B, k), np.float32, buf, o_s)
        count = np.ndarray((B,), np.int32, buf, o_c)
        status = np.ndarray((B,), np.int32, buf, o_t)
        base = ptr(buf)
        check(lib.frp_match(self.handle, ptr(q), B, float(tau), k, int(flags), base, base + o_s, base + o_c,
                            base + o_t))
        return idx, score, count, status

    def match_device(self, q_dev, B: int, tau: float, k: int, flags: int, out_idx, out_score, out_count,
                     out_status, stream: int = 0) -> None:
        check(lib.frp_match_device(self.handle, ptr(q_dev), int(B), float(tau), int(k), int(flags),
                                   ptr(out_idx), ptr(out_score), ptr(out_count), ptr(out_status),
                                   stream or None))

    def scores(self, queries) -> np.ndarray:
        """All scores [B, rows] in float64 from the fp32 master (NaN = deleted row).  The row
        count is the one the native call observed under its lock (``rows_out``), never an
        earlier read: a concurrent append cannot overrun the buffer (FRP_E_RANGE -> retry)."""
        q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
        if q.shape[1] != self.dim:
            raise ValueError("expected queries of length %d, got %d" % (self.dim, q.shape[1]))
        for _ in range(8):
            stride = self.rows + 64
            out = np.empty((q.shape[0], stride), np.float64)
            got = C.c_uint64(0)
            rc = lib.frp_scores(self.handle, ptr(q), q.shape[0], ptr(out), stride, C.byref(got))
            if rc == FRP_E_RANGE:
                continue                      # the gallery grew past the slack between the two calls
            check(rc)
            return out[:, :got.value]
        raise FrpError(FRP_E_RANGE, "gallery kept growing while frp_scores was sizing its output")

    def cluster(self, tau: float, flags: int = 0):
        """Greedy seed clustering on the device (frp_cluster).  Returns (labels[rows] int32 with
        -1 for deleted rows, n_clusters, stats dict)."""
        rows = self.rows
        labels = np.empty(rows + 64, np.int32)
        n = C.c_int(0)
        stats = np.zeros(4, np.uint64)
        check(lib.frp_cluster(self.handle, float(tau), int(flags), ptr(labels), len(labels), C.byref(n), ptr(stats)))
        rows_now = min(self.rows, len(labels))
        return labels[:rows_now], int(n.value), {"slabs": int(stats[0]), "rows_matched": int(stats[1]),
                                                  "dense_fallbacks": int(stats[2]), "kernel_launches": int(stats[3])}


class GalleryDict:
    """Drop-in for ``app.state.ENCODINGS``: ``Dict[str, list]`` whose values also
    live in HBM.  The device gallery is created lazily at the first insert (the
    embedding length is not known before)."""

    def __init__(self, metric: int = METRIC_L2, device: int = 0, keep_host: bool = True,
                 compact_ratio: float = 0.5, gallery_factory=None):
        # gallery_factory(dim, metric, device) -> Gallery-like; the tests inject a host
        # stand-in to exercise the dict bookkeeping without a GPU
        self._factory = gallery_factory or Gallery
        self.metric = metric
        self.device = device
        self.keep_host = keep_host
        self.compact_ratio = compact_ratio
        self._lock = RWLock()                    # readers: matchers; writers: every mutation
        self._rows: Dict[str, int] = {}          # name -> row; insertion order == row order
        self._names: List[Optional[str]] = []    # row -> name (None = tombstone)
        self._host: Dict[str, list] = {}
        self._gallery: Optional[Gallery] = None

    # -- device side ------------------------------------------------------
    @property
    def gallery(self) -> Optional[Gallery]:
        return self._gallery

    def reading(self):
        """Context manager: a consistent view of (device gallery, row numbering, names) for the
        duration of one match (subset -> native call -> ``name_of``).  Compaction, ``clear`` and
        every other mutation wait for the readers and vice versa."""
        return self._lock.read()

    def row_of(self, name: str) -> int:
        return self._rows[name]

    def name_of(self, row: int) -> Optional[str]:
        return self._names[row] if 0 <= row < len(self._names) else None

    def live_rows(self) -> np.ndarray:
        return np.fromiter(self._rows.values(), dtype=np.int64, count=len(self._rows))

    def _ensure(self, dim: int) -> Gallery:
        if self._gallery is None:
            self._gallery = self._factory(dim, self.metric, self.device)
        elif self._gallery.dim != dim:
            raise ValueError("embedding length %d does not match the gallery's %d" % (dim, self._gallery.dim))
        return self._gallery

    def _maybe_compact(self):
        dead = len(self._names) - len(self._rows)
        if dead > 1024 and dead > self.compact_ratio * len(self._names):
            self.compact()

    def compact(self):
        """Drop tombstones; live rows keep their relative order."""
        with self._lock.write():
            if self._gallery is None or len(self._names) == len(self._rows):
                return
            live = self.live_rows()
            new = self._gallery.gather(live)
            old = self._gallery
            self._gallery = new
            old.close()          # no reader holds it: the write lock excludes them
            self._names = list(self._rows.keys())
            self._rows = {n: i for i, n in enumerate(self._names)}

    def bulk_load(self, names: List[str], rows) -> None:
        """Append many (name, row) pairs at once (start-up load; the reference's
        retrieve_all_embeddings, utils/db.py:484-490, has no caller)."""
        a = np.asarray(rows)
        with self._lock.write():
            if len(set(names)) != len(names) or any(n in self._rows for n in names):
                raise ValueError("bulk_load needs new, distinct names")
            g = self._ensure(a.shape[1])
            first = g.append(a)
            assert first == len(self._names)
            for i, n in enumerate(names):
                self._rows[n] = first + i
                self._names.append(n)
                if self.keep_host:
                    self._host[n] = a[i].tolist()

    # -- dict protocol ------------------------------------------------------
    def __setitem__(self, name: str, value) -> None:
        vec = np.asarray(value, dtype=np.float64).reshape(-1)
        with self._lock.write():
            g = self._ensure(vec.shape[0])
            if name in self._rows:
                g.update(self._rows[name], vec)
            else:
                row = g.append(vec[None, :])
                assert row == len(self._names)
                self._rows[name] = row
                self._names.append(name)
            if self.keep_host:
                self._host[name] = value if isinstance(value, list) else vec.tolist()

    def __getitem__(self, name: str) -> list:
        with self._lock.read():
            row = self._rows[name]
            if self.keep_host:
                return self._host[name]
            return self._gallery.get_row(row).astype(np.float64).tolist()

    def __delitem__(self, name: str) -> None:
        with self._lock.write():
            row = self._rows.pop(name)
            self._names[row] = None
            self._host.pop(name, None)
            self._gallery.remove(row)
            self._maybe_compact()

    def __contains__(self, name) -> bool:
        return name in self._rows

    def __len__(self) -> int:
        return len(self._rows)

    def __iter__(self) -> Iterator[str]:
        return iter(list(self._rows.keys()))

    def __bool__(self) -> bool:
        return bool(self._rows)

    def keys(self):
        return list(self._rows.keys())

    def values(self):
        return [self[n] for n in self.keys()]

    def items(self):
        return [(n, self[n]) for n in self.keys()]

    def get(self, name, default=None):
        return self[name] if name in self._rows else default

    def pop(self, name, *default):
        with self._lock.write():
            if name not in self._rows:
                if default:
                    return default[0]
                raise KeyError(name)
            v = self[name]
            del self[name]
            return v

    def clear(self) -> None:
        with self._lock.write():
            if self._gallery is not None:
                self._gallery.close()
            self._gallery = None
            self._rows.clear()
            self._names.clear()
            self._host.clear()

    def update(self, other=(), **kw) -> None:
        items = other.items() if hasattr(other, "items") else other
        for k, v in items:
            self[k] = v
        for k, v in kw.items():
            self[k] = v

    def setdefault(self, name, default=None):
        if name not in self._rows:
            self[name] = default
        return self[name]

    def __repr__(self):
        return "GalleryDict(%d faces, metric=%s)" % (len(self), "cos" if self.metric == METRIC_COS else "l2")
