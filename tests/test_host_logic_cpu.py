"""Host-side logic added in round 2, exercised without a GPU: the readers-writer lock of
GalleryDict, the micro-batching front end (leader / follower hand-off, grouping by parameters,
error propagation) and the routing tables of the mutable sharded dict."""
import threading
import time

import numpy as np
import pytest

from frp_b200.camera_scan import BatchingMatcher
from frp_b200.gallery import RWLock
from frp_b200.shard import ShardedGalleryDict


def test_rwlock_excludes_writers_and_is_reentrant():
    lk = RWLock()
    state = {"readers": 0, "max_readers": 0, "writer_inside": False, "violations": 0}
    guard = threading.Lock()
    stop = time.time() + 0.6

    def reader():
        while time.time() < stop:
            with lk.read():
                with lk.read():                      # a reader may read again (FaceService -> GalleryDict)
                    with guard:
                        state["readers"] += 1
                        state["max_readers"] = max(state["max_readers"], state["readers"])
                        if state["writer_inside"]:
                            state["violations"] += 1
                    time.sleep(0.0005)
                    with guard:
                        state["readers"] -= 1

    def writer():
        while time.time() < stop:
            with lk.write():
                with lk.write():                     # compact() inside __delitem__
                    with lk.read():                  # a writer may read (pop -> __getitem__)
                        with guard:
                            if state["readers"]:
                                state["violations"] += 1
                            state["writer_inside"] = True
                        time.sleep(0.0005)
                        with guard:
                            state["writer_inside"] = False
            time.sleep(0.001)

    th = [threading.Thread(target=reader) for _ in range(4)] + [threading.Thread(target=writer) for _ in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert state["violations"] == 0
    assert state["max_readers"] >= 2                 # readers really share the lock
    with lk.read():
        with pytest.raises(RuntimeError):
            lk.acquire_write()                       # no read -> write upgrade (would deadlock)


class _FakeService:
    """Stands in for FaceService.match_topk: records every batch it is given."""

    def __init__(self, fail_on=None):
        self.batches = []
        self.fail_on = fail_on
        self.lock = threading.Lock()

    def match_topk(self, queries, threshold=None, top_k=5, thresholded=True):
        Q = np.atleast_2d(np.asarray(queries))
        with self.lock:
            self.batches.append((len(Q), threshold, top_k, thresholded))
        if self.fail_on is not None and any(float(q[0]) == self.fail_on for q in Q):
            raise ValueError("boom")
        time.sleep(0.002)
        return [[("t%d" % int(q[0]), float(q[0]) + (threshold or 0.0))][:top_k] for q in Q]

    def _get_confidence_level(self, d):
        return "high"


def test_batching_matcher_coalesces_and_routes_results():
    svc = _FakeService()
    bm = BatchingMatcher(svc, window_s=5e-3, max_batch=64)
    T, per = 6, 20
    out = {}
    start = threading.Barrier(T)

    def worker(t):
        start.wait()
        for i in range(per):
            q = np.array([t * 1000 + i, 0.0], np.float32)
            thr = 0.25 if t % 2 else None              # two parameter groups -> never mixed in one call
            out[(t, i)] = bm.match_topk(q, threshold=thr, top_k=3)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for t in range(T):
        for i in range(per):
            want = float(t * 1000 + i) + (0.25 if t % 2 else 0.0)
            assert out[(t, i)] == [("t%d" % (t * 1000 + i), want)]
    assert bm.queries == T * per
    assert sum(b[0] for b in svc.batches) == T * per
    assert len(svc.batches) < T * per                  # something was coalesced
    assert all(b[1] in (None, 0.25) and b[2] == 3 for b in svc.batches)


def test_batching_matcher_propagates_errors_to_the_failed_batch_only():
    svc = _FakeService(fail_on=7.0)
    bm = BatchingMatcher(svc, window_s=0.0)
    assert bm.match_topk(np.array([1.0, 0.0], np.float32)) == [("t1", 1.0)]
    with pytest.raises(ValueError):
        bm.match_topk(np.array([7.0, 0.0], np.float32))
    assert bm.match_topk(np.array([2.0, 0.0], np.float32)) == [("t2", 2.0)]      # the matcher is still usable


class _HostShard:
    def __init__(self, dim, metric, device):
        self.rows, self.alive, self.seq = [], [], {}

    def append(self, rows):
        first = len(self.rows)
        for r in np.atleast_2d(rows):
            self.rows.append(np.array(r, np.float64))
            self.alive.append(True)
        return first

    def set_seq(self, first_row, seqs):
        for i, s in enumerate(seqs):
            self.seq[first_row + i] = int(s)

    def update(self, row, vec):
        assert self.alive[row]
        self.rows[row] = np.array(vec, np.float64)

    def remove(self, row):
        assert self.alive[row]
        self.alive[row] = False

    def get_row(self, row):
        return self.rows[row].astype(np.float32)

    def close(self):
        pass


def test_sharded_dict_routing_tables():
    """New names go to the least-full shard with the next global sequence, updates keep their
    place, deletes free capacity, rename re-appends; every shard's sequences increase with the
    local row (the property the kernels' tie order relies on)."""
    rng = np.random.default_rng(3)
    d = ShardedGalleryDict(4, 8, 0, gallery_factory=_HostShard)
    ref = {}
    names = ["n%02d" % i for i in range(60)]
    for step in range(1500):
        n = names[rng.integers(len(names))]
        op = rng.random()
        if op < 0.6:
            v = rng.normal(size=8)
            d[n] = v
            ref[n] = v
        elif op < 0.85:
            if n in ref:
                del d[n]
                del ref[n]
        else:
            if n in ref:
                new = "%s_%d" % (n, step)
                v = ref.pop(n)
                d.rename(n, new, v)
                ref[new] = v
                names[names.index(n)] = new
        assert d.keys() == list(ref.keys())
        live = d.live_per_shard()
        assert sum(live) == len(ref)
    for r, g in d.galleries.items():
        seqs = [g.seq[i] for i in range(len(g.rows))]
        assert seqs == sorted(seqs) and len(set(seqs)) == len(seqs)
        assert sum(g.alive) == d.live_per_shard()[r]
    for n, v in ref.items():
        r, row, seq = d._where[n]
        assert d.galleries[r].alive[row] and np.allclose(d.galleries[r].rows[row], v) and d.name_of(seq) == n
    # an SPMD rank owns one shard only and keeps the same tables
    a = ShardedGalleryDict(3, 8, 0, rank=1, gallery_factory=_HostShard)
    b = ShardedGalleryDict(3, 8, 0, gallery_factory=_HostShard)
    for i in range(10):
        v = rng.normal(size=8)
        a["k%d" % i] = v
        b["k%d" % i] = v
    assert a._where == b._where and list(a.galleries) == [1]
    assert len(a.galleries[1].rows) == len(b.galleries[1].rows)
