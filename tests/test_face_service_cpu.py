"""Host-side mirror (`FaceService` + `GalleryDict`) against the UNMODIFIED reference service,
on CPU: the device gallery is replaced by a numpy test double that follows the C ABI's
documented contract (float64 scores of float32 inputs, inclusive threshold, strict flag,
threshold-then-top-k with stable row order), so what is checked here is the Python logic around
it — result dictionaries, ordering, confidence formulas, subset handling, duplicate detection,
deletion, clustering.  Skipped where /root/reference is absent (the GPU box); the same calls run
against the real kernels in tests/test_match_gpu.py."""
import numpy as np
import pytest

import frp_b200
from frp_b200 import FaceService, GalleryDict, METRIC_L2
from frp_b200._capi import MATCH_NO_THRESHOLD, MATCH_STRICT
import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


class _NumpyGallery:
    """Test double for frp_b200.Gallery (L2 metric only)."""

    def __init__(self, dim, metric, device):
        assert metric == METRIC_L2
        self.dim, self.metric, self.device = dim, metric, device
        self.data, self.alive = [], []

    rows = property(lambda self: len(self.data))

    def append(self, rows):
        first = len(self.data)
        for r in np.atleast_2d(rows):
            self.data.append(np.asarray(r, np.float32).astype(np.float64))
            self.alive.append(True)
        return first

    def update(self, row, vec):
        self.data[row] = np.asarray(vec, np.float32).astype(np.float64)

    def remove(self, row):
        self.alive[row] = False

    def gather(self, idx):
        g = _NumpyGallery(self.dim, self.metric, self.device)
        g.append(np.stack([self.data[int(i)] for i in idx]))
        return g

    def get_row(self, row):
        return self.data[row].astype(np.float32)

    def close(self):
        pass

    def scores(self, queries):
        Q = np.atleast_2d(np.asarray(queries, np.float32)).astype(np.float64)
        G = np.stack(self.data)
        out = np.stack([np.linalg.norm(G - q, axis=1) for q in Q])
        out[:, ~np.array(self.alive)] = np.nan
        return out

    def cluster(self, tau, flags=0):
        """Contract of frp_cluster: greedy seed labels in row order, -1 for deleted rows."""
        G = np.stack(self.data)
        alive = np.array(self.alive)
        labels = np.full(len(G), -1, np.int32)
        cid = 0
        for i in range(len(G)):
            if not alive[i] or labels[i] >= 0:
                continue
            d = np.linalg.norm(G - G[i], axis=1)
            labels[alive & (labels < 0) & (d <= tau)] = cid
            labels[i] = cid
            cid += 1
        return labels, cid, {}

    def match(self, queries, tau, k, flags=0):
        S = self.scores(queries)
        B = S.shape[0]
        idx = np.full((B, k), -1, np.int64)
        sc = np.full((B, k), np.nan, np.float32)
        cnt = np.zeros(B, np.int32)
        for b in range(B):
            ok = ~np.isnan(S[b])
            if not flags & MATCH_NO_THRESHOLD:
                ok &= (S[b] < tau) if flags & MATCH_STRICT else (S[b] <= tau)
            cand = np.nonzero(ok)[0]
            cand = cand[np.argsort(S[b][cand], kind="stable")][:k]
            idx[b, :len(cand)], sc[b, :len(cand)], cnt[b] = cand, S[b][cand], len(cand)
        return idx, sc, cnt, np.zeros(B, np.int32)


def _same(a, b, tol=1e-6):
    """Deep comparison; floats within tol: the C ABI returns top-k scores as float32 (gate in
    BASELINE.json: distances within 1e-3), the dense path is float64 up to summation order."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), (a, b)
        for k in a:
            _same(a[k], b[k], tol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (len(a), len(b))
        for x, y in zip(a, b):
            _same(x, y, tol)
    elif isinstance(a, float):
        assert abs(a - b) <= tol, (a, b)
    else:
        assert a == b, (a, b)


@pytest.fixture()
def pair():
    ref, ENC, _ = ref_loader.load_reference()
    ENC.clear()
    ref.tolerance = 0.6
    ours = FaceService(GalleryDict(gallery_factory=_NumpyGallery), tolerance=0.6, raise_errors=True)
    yield ref, ours, ENC
    ENC.clear()


def test_scripted_session_matches_the_reference(pair):
    ref, ours, ENC = pair
    rng = np.random.default_rng(2024)
    # float32-representable encodings: both sides then compute on identical numbers
    people = rng.normal(0, 0.05, (40, 128)).astype(np.float32).astype(np.float64)
    for i, g in enumerate(people):
        _same(ref.store_face("person_%02d" % i, g), ours.store_face("person_%02d" % i, g))
    # a near duplicate of an existing face (strict < 0.3 scan, first hit in dict order wins)
    dup = (people[7] + rng.normal(0, 0.002, 128)).astype(np.float32).astype(np.float64)
    _same(ref.store_face("lookalike", dup), ours.store_face("lookalike", dup))
    # overwriting an existing name keeps its place in the order
    upd = rng.normal(0, 0.05, 128).astype(np.float32).astype(np.float64)
    _same(ref.store_face("person_03", upd), ours.store_face("person_03", upd))
    assert ref.get_all_targets() == ours.get_all_targets()

    queries = [(people[j] + rng.normal(0, 0.02, 128)).astype(np.float32).astype(np.float64) for j in (0, 7, 11, 39)]
    queries += [rng.normal(0, 0.05, 128).astype(np.float32).astype(np.float64) for _ in range(3)]
    for q in queries:
        _same(ref.compare_faces(q), ours.compare_faces(q))
        _same(ref.compare_faces(q, return_distances=False), ours.compare_faces(q, return_distances=False))
        _same(ref.find_k_nearest(q, 5), ours.find_k_nearest(q, 5))
        sub = ["person_11", "nobody", "lookalike", "person_07"]
        _same(ref.compare_faces(q, target_names=sub), ours.compare_faces(q, target_names=sub))
    _same(ref.batch_compare_faces(queries), ours.batch_compare_faces(queries))
    _same(ref.batch_compare_faces(queries, target_names=["person_00", "person_39"]),
          ours.batch_compare_faces(queries, target_names=["person_00", "person_39"]))
    _same(ref.cluster_faces(0.6), ours.cluster_faces(0.6))

    # deletions: existing, repeated, unknown
    for name in ("person_07", "person_07", "ghost"):
        r, o_ = ref.delete_face(name), ours.delete_face(name)
        assert r["success"] == o_["success"] and r.get("removed_from_memory") == o_.get("removed_from_memory"), (r, o_)
    assert ref.get_all_targets() == ours.get_all_targets()
    for q in queries[:3]:
        _same(ref.compare_faces(q), ours.compare_faces(q))
        _same(ref.find_k_nearest(q, 50), ours.find_k_nearest(q, 50))      # k larger than the gallery
    _same(ref.cluster_faces(0.75), ours.cluster_faces(0.75))


def test_empty_gallery_and_bad_input_follow_the_never_raise_convention(pair):
    ref, ours, ENC = pair
    ours.raise_errors = False
    q = np.zeros(128)
    assert ref.compare_faces(q) == ours.compare_faces(q) == []
    assert ref.batch_compare_faces([q, q]) == ours.batch_compare_faces([q, q])
    assert ref.find_k_nearest(q, 3) == ours.find_k_nearest(q, 3)
    _same(ref.cluster_faces(0.6), ours.cluster_faces(0.6))
    _same(ref.store_face("a", np.ones(128) * 0.01), ours.store_face("a", np.ones(128) * 0.01))
    # wrong query length: the reference logs and returns [], so does the mirror
    assert ref.compare_faces(np.zeros(64)) == ours.compare_faces(np.zeros(64)) == []


def test_batched_camera_scan_equals_the_reference_per_face_loop(pair):
    """scan_cameras (one batched match for all faces of all cameras) against the loop of
    backend/app/routes/camera.py:243-259 driven by the live reference service: every gallery row
    with ``match and distance <= confidence_threshold``, faces in detection order, matches by
    ascending distance; ``max_faces`` truncation as in :233-235."""
    from frp_b200 import scan_cameras
    ref, ours, ENC = pair
    rng = np.random.default_rng(7)
    people = rng.normal(0, 0.05, (60, 128)).astype(np.float32).astype(np.float64)
    for i, g in enumerate(people):
        ref.store_face("w%02d" % i, g)
        ours.store_face("w%02d" % i, g)
    # a second enrolment of some people, so one face can match several rows
    for i in (4, 9):
        twin = (people[i] + rng.normal(0, 0.01, 128)).astype(np.float32).astype(np.float64)
        ref.store_face("w%02d_b" % i, twin)
        ours.store_face("w%02d_b" % i, twin)

    def face(j=None):
        base = people[j] if j is not None else rng.normal(0, 0.05, 128)
        return (base + rng.normal(0, 0.02, 128)).astype(np.float32).astype(np.float64)

    cams = {"gate": [face(4), face(), face(9)], "lobby": [], 7: [face(30)] + [face() for _ in range(4)],
            "dock": [face(i) for i in (1, 2, 3, 4)]}
    cfgs = {"gate": {"confidence_threshold": 0.6}, 7: {"confidence_threshold": 0.35, "max_faces": 3},
            "dock": {"max_faces": 2}}
    got = scan_cameras(ours, cams, cfgs)
    want = {}
    for cam, faces in cams.items():
        cfg = cfgs.get(cam, {})
        thr, mf = cfg.get("confidence_threshold", 0.6), cfg.get("max_faces", 10)
        out = []
        for enc in list(faces)[:mf]:
            for m in ref.compare_faces(enc, target_names=None, return_distances=True):
                if m.get("match") and m.get("distance", 0.0) <= thr:
                    out.append({"camera_id": cam, "target": m["target"], "distance": m["distance"],
                                "confidence": m["confidence"]})
        want[cam] = out
    assert list(got.keys()) == list(want.keys())
    _same(want, got)
    assert sum(len(v) for v in got.values()) >= 6
