"""Parity of the CUDA 1:N matcher (through the C ABI) against the oracle / the
reference's golden vectors.  Tolerances are BASELINE.json's: decisions and top-k
indices identical except where candidates tie within 1e-3; distances within 1e-3
absolute after the fp32 re-rank (the asserts below use far tighter bounds where the
arithmetic allows)."""
import numpy as np
import pytest

import frp_b200
from frp_b200 import (FaceService, Gallery, GalleryDict, METRIC_COS, METRIC_L2, MATCH_FORCE_EXACT,
                      MATCH_FORCE_GEMM, MATCH_FORCE_SCAN, MATCH_NO_THRESHOLD)
import frp_oracle_np as o

pytestmark = pytest.mark.gpu

TIE = 1e-3        # north_star: candidates closer than this may swap
DIST_TOL = 1e-3   # north_star: distances within 1e-3 absolute
TIGHT = 2e-6      # what fp32 master + fp64 re-rank actually delivers


def check_topk(idx, sc, cnt, widx, wsc, wcnt, all_scores, tau, larger_is_better, thresholded=True):
    """Compare a (idx, score, count) result with the oracle's, allowing swaps only
    between candidates whose oracle scores tie within TIE and threshold flips only
    within TIE of tau."""
    B, k = idx.shape
    for b in range(B):
        n, wn = int(cnt[b]), int(wcnt[b])
        s_all = all_scores[b]
        got = idx[b, :n]
        assert len(set(got.tolist())) == n
        # scores reported for the returned rows are right
        assert np.allclose(sc[b, :n], s_all[got], rtol=0, atol=DIST_TOL)
        assert np.allclose(sc[b, :n], s_all[got], rtol=0, atol=TIGHT), np.abs(sc[b, :n] - s_all[got]).max()
        if n != wn:
            # only rows within TIE of the threshold may flip
            assert thresholded
            near = np.sum(np.abs(s_all - tau) < TIE)
            assert abs(n - wn) <= near
        m = min(n, wn)
        for j in range(m):
            if got[j] != widx[b, j]:
                assert abs(s_all[got[j]] - wsc[b, j]) < TIE, (b, j, got[j], widx[b, j])
        # ordering is monotone
        d = np.diff(sc[b, :n])
        assert np.all(d <= 1e-7) if larger_is_better else np.all(d >= -1e-7)
        assert np.all(idx[b, n:] == -1) and np.all(np.isnan(sc[b, n:]))


def exact_equal(idx, cnt, widx, wcnt):
    return np.array_equal(cnt, wcnt.astype(cnt.dtype)) and np.array_equal(idx, widx)


# ----------------------------------------------------------------------------
# golden vectors produced by the UNMODIFIED reference (config C1)
# ----------------------------------------------------------------------------
def _service_from(G, metric=METRIC_L2):
    enc = GalleryDict(metric)
    for i, g in enumerate(G):
        enc["id%07d" % i] = g.tolist()
    return FaceService(enc, tolerance=0.6, raise_errors=True), enc


def test_compare_faces_matches_reference_golden(golden_l2):
    g = golden_l2
    fs, enc = _service_from(g["G"])
    names = enc.keys()
    lvl = {"high": 0, "medium": 1, "low": 2}
    for b in range(len(g["Q"])):
        res = fs.compare_faces(g["Q"][b], None, True)
        assert len(res) == len(g["G"])
        rows = np.array([names.index(r["target"]) for r in res])
        dist = np.array([r["distance"] for r in res])
        match = np.array([r["match"] for r in res])
        assert np.abs(dist - g["dist"][b]).max() < TIGHT
        # order identical except inside groups tying within TIE
        diff = rows != g["order"][b]
        assert np.all(np.abs(dist[diff] - g["dist"][b][diff]) < TIGHT)
        for j in np.nonzero(diff)[0]:
            assert np.abs(g["dist"][b][j] - g["dist"][b][g["order"][b].tolist().index(rows[j])]) < TIE
        # decisions identical except within TIE of the tolerance
        want_match = np.zeros(len(rows), bool)
        want_match[g["order"][b]] = g["match"][b]
        flip = match != want_match[rows]
        assert np.all(np.abs(dist[flip] - 0.6) < TIE)
        # with the fp32 master the only flips are the planted |d - 0.6| <= 1e-6 rows
        assert np.all(np.abs(dist[flip] - 0.6) < 2e-6)
        ok = ~flip
        assert [lvl[r["confidence"]] for r in res] == [lvl[fs._get_confidence_level(r["distance"])] for r in res]
        sc = np.array([r["confidence_score"] for r in res])
        assert np.abs(sc[ok] - np.array([o.calibrate_confidence(d) for d in dist[ok]])).max() <= 0.011


def test_threshold_topk_matches_reference_golden(golden_l2):
    g = golden_l2
    gal = Gallery(128, METRIC_L2)
    gal.append(g["G"])
    Q = g["Q"].astype(np.float32)
    all_d = np.stack([np.linalg.norm(g["G"] - q, axis=1) for q in g["Q"]])
    for k in (1, 5, 10, 50):
        widx, wd, wcnt = o.l2_match_topk(g["G"], g["Q"], 0.6, k, True)
        # the golden itself: threshold-then-top-k of the reference's sorted output
        for b in range(len(Q)):
            want = g["order"][b][g["match"][b]][:k]
            assert widx[b, :wcnt[b]].tolist() == want.tolist()
        for flags in (0, MATCH_FORCE_SCAN, MATCH_FORCE_EXACT):
            idx, sc, cnt, st = gal.match(Q, 0.6, k, flags)
            check_topk(idx, sc, cnt, widx, wd, wcnt, all_d, 0.6, False)
            assert np.all(st <= 1)
    # find_k_nearest (no threshold)
    widx, wd, wcnt = o.l2_match_topk(g["G"], g["Q"], 0.0, 5, False)
    idx, sc, cnt, st = gal.match(Q, 0.0, 5, MATCH_NO_THRESHOLD)
    check_topk(idx, sc, cnt, widx, wd, wcnt, all_d, 0.0, False, thresholded=False)
    assert np.abs(sc - g["knn_d"]).max() < TIGHT


def test_service_methods_match_reference_golden(golden_l2):
    g = golden_l2
    fs, enc = _service_from(g["G"])
    names = enc.keys()
    bat = fs.batch_compare_faces(list(g["Q"]), None)
    for b, r in enumerate(bat):
        rows = [names.index(x["target"]) for x in r]
        want = g["bat_rows"][b, :g["bat_cnt"][b]].tolist()
        if rows != want:   # only the planted 0.6 +- 1e-6 rows may differ
            sym = set(rows) ^ set(want)
            d = np.linalg.norm(g["G"][sorted(sym)] - g["Q"][b], axis=1)
            assert np.all(np.abs(d - 0.6) < 2e-6)
        assert all(x["match"] is True and x["confidence"] in ("high", "medium", "low") for x in r)
    for b in range(0, len(g["Q"]), 5):
        res = fs.find_k_nearest(g["Q"][b], 5)
        assert np.abs(np.array([r["distance"] for r in res]) - g["knn_d"][b]).max() < TIGHT
        assert set(res[0].keys()) == {"target", "distance", "confidence", "confidence_score"}
    # subset + unknown names (face_service.py:403)
    sub = [names[r] if r >= 0 else "nope" for r in g["sub_query_rows"]]
    res = fs.compare_faces(g["Q"][0], sub, True)
    assert [names.index(r["target"]) for r in res] == g["sub_rows"].tolist()
    assert np.abs(np.array([r["distance"] for r in res]) - g["sub_dist"]).max() < TIGHT
    assert fs.compare_faces(g["Q"][0], ["nope"]) == []
    # threshold + top-k like /face/compare (routes/face.py:685-693)
    tk = fs.match_topk(g["Q"], threshold=0.6, top_k=5)
    for b in range(len(tk)):
        want = g["order"][b][g["match"][b]][:5].tolist()
        got = [names.index(t) for t, _ in tk[b]]
        if got != want:
            # the only admissible difference: rows planted within 1e-6 of the tolerance, where the
            # fp32 master and the reference's float64 lists may decide differently (gate: 1e-3)
            sym = sorted(set(got) ^ set(want))
            d = np.linalg.norm(g["G"][sym] - g["Q"][b], axis=1)
            assert len(sym) and np.all(np.abs(d - 0.6) < 2e-6), (b, got, want)
            common = [r for r in got if r in want]
            assert common == [r for r in want if r in got]          # same order on the rows both hold
        dd = np.array([v for _, v in tk[b]])
        assert np.all(np.diff(dd) >= 0) and np.all(dd <= 0.6 + 2e-6)
        ref_d = np.linalg.norm(g["G"][got] - g["Q"][b], axis=1)
        assert np.abs(dd - ref_d).max(initial=0) < TIGHT


def test_store_delete_mirror_dict_semantics():
    rng = np.random.default_rng(5)
    G = rng.normal(0, 0.05, (200, 128))
    fs, enc = _service_from(G[:100])
    ref = {"id%07d" % i: G[i].tolist() for i in range(100)}
    # duplicate warning: first existing face in dict order closer than 0.3 (strict)
    r = fs.store_face("dup", G[40] + 0.001)
    assert r["success"] and r["is_duplicate"] and r["similar_to"] == "id0000040" and not r["was_update"]
    ref["dup"] = (G[40] + 0.001).tolist()
    r = fs.store_face("id0000003", G[150])            # update keeps position
    assert r["was_update"] and not r["is_duplicate"]
    ref["id0000003"] = G[150].tolist()
    assert fs.delete_face("id0000010")["removed_from_memory"]
    del ref["id0000010"]
    assert not fs.delete_face("missing")["success"]
    r = fs.store_face("id0000010", G[10])              # re-add moves to the end
    ref["id0000010"] = G[10].tolist()
    assert fs.get_all_targets() == list(ref.keys())
    for q in (G[40], G[150], G[10], G[77] + 0.01):
        res = fs.compare_faces(q)
        want = o.compare_faces(ref, q, 0.6)
        assert [x["target"] for x in res] == [x["target"] for x in want]
        assert np.abs(np.array([x["distance"] for x in res]) - np.array([x["distance"] for x in want])).max() < TIGHT
        assert [x["match"] for x in res] == [x["match"] for x in want]
        knn = fs.find_k_nearest(q, 3)
        assert [x["target"] for x in knn] == [x["target"] for x in want[:3]]
    cl = fs.cluster_faces(0.6)
    members = sorted(sum(cl.values(), []))
    assert members == sorted(ref.keys())
    assert any("dup" in v and "id0000040" in v for v in cl.values())


def test_camera_scan_batched_front_end():
    """scan_cameras == the reference's per-face loop of routes/camera.py:243-259 restated
    on the oracle: every gallery row with match and distance <= confidence_threshold."""
    rng = np.random.default_rng(21)
    G = rng.normal(0, 0.05, (600, 128))
    G[300:380] = G[5] + rng.normal(0, 0.004, (80, 128))       # > 64 matches for one face -> dense path
    fs, enc = _service_from(G)
    ref = {"id%07d" % i: g.tolist() for i, g in enumerate(G)}
    faces = {
        "cam0": [G[10] + rng.normal(0, 0.02, 128), rng.normal(0, 0.05, 128)],
        7: [G[5] + rng.normal(0, 0.003, 128)],
        "cam2": [],
        "cam3": [G[i] + rng.normal(0, 0.02, 128) for i in range(20, 34)],      # 14 faces, max_faces 10
    }
    cfg = {"cam0": {"confidence_threshold": 0.5}, "cam3": {"confidence_threshold": 0.25}}
    got = frp_b200.scan_cameras(fs, faces, cfg)
    assert list(got.keys()) == list(faces.keys())
    for cam, fl in faces.items():
        thr = cfg.get(cam, {}).get("confidence_threshold", 0.6)
        want = []
        for f in list(fl)[:10]:
            for m in o.compare_faces(ref, f, 0.6):
                if m["match"] and m["distance"] <= thr:
                    want.append({"camera_id": cam, "target": m["target"], "distance": m["distance"],
                                 "confidence": m["confidence"]})
        assert [w["target"] for w in want] == [r["target"] for r in got[cam]], cam
        assert [w["confidence"] for w in want] == [r["confidence"] for r in got[cam]]
        assert all(r["camera_id"] == cam for r in got[cam])
        if want:
            assert max(abs(w["distance"] - r["distance"]) for w, r in zip(want, got[cam])) < TIGHT
    assert len(got[7]) > 64 and got["cam2"] == []
    assert frp_b200.scan_cameras(FaceService(GalleryDict(), raise_errors=True), faces) == {c: [] for c in faces}


# ----------------------------------------------------------------------------
# cosine 512-d (insightface semantics; oracle = numpy restatement)
# ----------------------------------------------------------------------------
def test_cosine_golden(golden_cos):
    g = golden_cos
    gal = Gallery(512, METRIC_COS)
    gal.append(g["G"])
    tau, k = float(g["tau"]), int(g["k"])
    for flags in (0, MATCH_FORCE_SCAN, MATCH_FORCE_EXACT):
        idx, sc, cnt, st = gal.match(g["Q"], tau, k, flags)
        check_topk(idx, sc, cnt, g["idx"], g["score"], g["count"], g["all_scores"], tau, True)
        idx, sc, cnt, st = gal.match(g["Q"], tau, k, flags | MATCH_NO_THRESHOLD)
        check_topk(idx, sc, cnt, g["idx_nt"], g["score_nt"], np.full(len(g["Q"]), k), g["all_scores"], tau,
                   True, thresholded=False)
    dense = gal.scores(g["Q"])
    assert np.abs(dense - g["all_scores"]).max() < TIGHT
    # duplicates of row 11: query 23 == row 11 -> four exact ties, insertion order
    idx, sc, cnt, st = gal.match(g["Q"][23:24], 0.99, 10)
    dup = np.nonzero(np.all(g["G"] == g["G"][11], axis=1))[0]
    assert idx[0, :cnt[0]].tolist() == dup.tolist()


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 9, 33])
def test_cosine_random_vs_oracle(B):
    N, D, k, tau = 60_000, 512, 10, 0.40
    G = o.synth_cos_gallery(N, D, seed=100 + B)
    Q, planted = o.synth_cos_queries(G, B, seed=200 + B)
    gal = Gallery(D, METRIC_COS)
    gal.append(G)
    S = o.cosine_scores(G, Q)
    widx, wsc, wcnt = o.cos_match_topk(G, Q, tau, k, True)
    idx, sc, cnt, st = gal.match(Q, tau, k)
    check_topk(idx, sc, cnt, widx, wsc, wcnt, S, tau, True)
    assert exact_equal(idx, cnt, widx, wcnt)          # no ties in this data: bit-identical decisions
    assert np.all(st == 0), "fp16 prefilter should certify random data"
    widx, wsc, wcnt = o.cos_match_topk(G, Q, tau, k, False)
    idx, sc, cnt, st = gal.match(Q, tau, k, MATCH_NO_THRESHOLD)
    assert exact_equal(idx, cnt, widx, wcnt)
    for b in range(B):
        if planted[b] >= 0:
            assert idx[b, 0] == planted[b]


@pytest.mark.parametrize("B", [1, 7, 128, 129, 256, 300])
def test_tcgen05_path_vs_oracle(B):
    """The tensor-core prefilter (forced for every batch size) against the oracle, and
    against the streaming scan path: identical indices and scores."""
    N, D, k, tau = 50_021, 512, 10, 0.35
    G = o.synth_cos_gallery(N, D, seed=300 + B)
    Q, planted = o.synth_cos_queries(G, B, seed=400 + B)
    gal = Gallery(D, METRIC_COS)
    gal.append(G)
    for thresholded in (True, False):
        fl = 0 if thresholded else MATCH_NO_THRESHOLD
        widx, wsc, wcnt = o.cos_match_topk(G, Q, tau, k, thresholded)
        idx, sc, cnt, st = gal.match(Q, tau, k, fl | MATCH_FORCE_GEMM)
        assert frp_b200.last_match_path() == "gemm_tcgen05"
        assert exact_equal(idx, cnt, widx, wcnt)
        assert np.all(st == 0)
        i2, s2, c2, _ = gal.match(Q, tau, k, fl | MATCH_FORCE_SCAN)
        assert np.array_equal(i2, idx) and np.array_equal(s2, sc, equal_nan=True)


@pytest.mark.parametrize("dim,metric", [(128, METRIC_L2), (64, METRIC_COS), (256, METRIC_L2), (100, METRIC_COS)])
def test_tcgen05_path_other_dims(dim, metric):
    rng = np.random.default_rng(dim)
    N, B, k = 9000, 40, 12
    G = rng.normal(0, 0.06, (N, dim)).astype(np.float32)
    Q = G[rng.integers(0, N, B)] + rng.normal(0, 0.03, (B, dim)).astype(np.float32)
    gal = Gallery(dim, metric)
    gal.append(G)
    gal.remove(int(rng.integers(0, N)))
    if metric == METRIC_L2:
        S = np.stack([np.linalg.norm(G.astype(np.float64) - q.astype(np.float64), axis=1) for q in Q])
        tau = float(np.sort(S[0])[25])
    else:
        S = o.cosine_scores(G, Q)
        tau = float(np.sort(S[0])[-25])
    a = gal.match(Q, tau, k, MATCH_FORCE_GEMM)
    assert frp_b200.last_match_path() == "gemm_tcgen05"
    b = gal.match(Q, tau, k, MATCH_FORCE_SCAN)
    c = gal.match(Q, tau, k, MATCH_FORCE_EXACT)
    for r in (b, c):
        assert np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1], equal_nan=True) and np.array_equal(a[2], r[2])
    for bq in range(B):
        got = a[0][bq, :a[2][bq]]
        assert np.abs(a[1][bq, :a[2][bq]] - S[bq][got]).max(initial=0) < TIGHT


@pytest.mark.parametrize("B", [200, 256, 512])
@pytest.mark.parametrize("dim,metric", [(256, METRIC_L2), (512, METRIC_L2), (384, METRIC_COS), (256, METRIC_COS)])
def test_tcgen05_pair_kernel_dims_metrics(B, dim, metric):
    """Even numbers of 128-query groups on rows >= 256-d run the cta_group::2 CTA-pair kernel
    (both metrics: the L2 form also exercises its per-CTA norm ring); a ragged last tile, a
    partly filled second group (B = 200) and tombstones included."""
    rng = np.random.default_rng(1000 * dim + B)
    N, k = 20_011, 9
    G = rng.normal(0, 0.06, (N, dim)).astype(np.float32)
    Q = G[rng.integers(0, N, B)] + rng.normal(0, 0.03, (B, dim)).astype(np.float32)
    gal = Gallery(dim, metric)
    gal.append(G)
    dead = rng.integers(0, N, 7)
    for r in dead:
        gal.remove(int(r))
    if metric == METRIC_L2:
        S = np.stack([np.linalg.norm(G.astype(np.float64) - q.astype(np.float64), axis=1) for q in Q[:8]])
        tau = float(np.sort(S[0])[30])
    else:
        S = o.cosine_scores(G, Q[:8])
        tau = float(np.sort(S[0])[-30])
    a = gal.match(Q, tau, k, MATCH_FORCE_GEMM)
    assert frp_b200.last_match_path() == "gemm_tcgen05"
    assert np.all(a[3] == 0)
    b = gal.match(Q, tau, k, MATCH_FORCE_SCAN)
    c = gal.match(Q, tau, k, MATCH_FORCE_EXACT)
    for r in (b, c):
        assert np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1], equal_nan=True) and np.array_equal(a[2], r[2])
    assert not np.isin(a[0][a[0] >= 0], dead).any()
    for bq in range(8):
        got = a[0][bq, :a[2][bq]]
        assert np.abs(a[1][bq, :a[2][bq]] - S[bq][got]).max(initial=0) < TIGHT
        live = np.ones(N, bool)
        live[dead] = False
        order = np.argsort(S[bq] if metric == METRIC_L2 else -S[bq], kind="stable")
        order = order[live[order]]
        ok = (S[bq][order] <= tau) if metric == METRIC_L2 else (S[bq][order] >= tau)
        want = order[ok][:k]
        assert np.array_equal(got, want)


@pytest.mark.parametrize("N", [1, 100, 257, 700])
def test_tcgen05_pair_kernel_tiny_galleries(N):
    """Fewer gallery tiles than CTA pairs: one (partial) tile per pair, the rest of the grid is
    capped away; batches of 256 and 512 queries."""
    rng = np.random.default_rng(N)
    G = rng.standard_normal((N, 512)).astype(np.float32)
    gal = Gallery(512, METRIC_COS)
    gal.append(G)
    for B in (256, 512):
        Q = rng.standard_normal((B, 512)).astype(np.float32)
        Q[::3] = G[rng.integers(0, N, len(Q[::3]))] + 0.05 * Q[::3]
        a = gal.match(Q, 0.3, 5, MATCH_FORCE_GEMM)
        assert frp_b200.last_match_path() == "gemm_tcgen05" and np.all(a[3] == 0)
        b = gal.match(Q, 0.3, 5, MATCH_FORCE_EXACT)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True) and np.array_equal(a[2], b[2])
        widx, wsc, wcnt = o.cos_match_topk(G, Q[:16], 0.3, 5, True)
        assert exact_equal(a[0][:16], a[2][:16], widx, wcnt)


@pytest.mark.parametrize("dim", [64, 100, 128, 256, 384, 512])
@pytest.mark.parametrize("metric", [METRIC_L2, METRIC_COS])
def test_dims_and_metrics(dim, metric):
    rng = np.random.default_rng(dim + metric)
    N, B, k = 5000, 3, 7
    G = rng.normal(0, 0.06, (N, dim)).astype(np.float32)
    Q = G[rng.integers(0, N, B)] + rng.normal(0, 0.02, (B, dim)).astype(np.float32)
    gal = Gallery(dim, metric)
    gal.append(G[:3000])
    gal.append(G[3000:].astype(np.float64))           # f64 ingest path
    if metric == METRIC_L2:
        S = np.stack([np.linalg.norm(G.astype(np.float64) - q.astype(np.float64), axis=1) for q in Q])
        tau = float(np.sort(S[0])[20])
        widx, wsc, wcnt = o.l2_match_topk(G, Q, tau, k, True)
        lib = False
    else:
        S = o.cosine_scores(G, Q)
        tau = float(np.sort(S[0])[-20])
        widx, wsc, wcnt = o.cos_match_topk(G, Q, tau, k, True)
        lib = True
    for flags in (0, MATCH_FORCE_EXACT):
        idx, sc, cnt, st = gal.match(Q, tau, k, flags)
        check_topk(idx, sc, cnt, widx, wsc, wcnt, S, tau, lib)
    assert np.abs(gal.scores(Q) - S).max() < TIGHT


@pytest.mark.parametrize("seed", range(12))
def test_exact_ties_and_inclusive_threshold(seed):
    """Lattice-valued encodings make distances tie EXACTLY (in fp16, fp32 and fp64 alike) and
    put gallery rows exactly AT the tolerance: the inclusive `<=` of face_service.py:411 and
    the stable tie order (dict insertion order, :431) must then be reproduced bit for bit,
    on every path (scan / tcgen05 / fp32 master)."""
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.choice([64, 128, 512]))
    N = int(rng.integers(1, 400))
    B = int(rng.integers(1, 10))
    k = int(rng.integers(1, 13))
    G = rng.integers(-2, 3, (N, dim)).astype(np.float64) / 16.0
    G[rng.integers(0, N, N // 3)] = G[rng.integers(0, N, N // 3)]          # exact duplicates
    Q = rng.integers(-2, 3, (B, dim)).astype(np.float64) / 16.0
    D = np.stack([np.linalg.norm(G - q, axis=1) for q in Q])               # exact sums, one sqrt
    tau = float(np.sort(D[0])[min(N - 1, int(rng.integers(0, 2 * k + 1)))])  # a row sits exactly at tau
    gal = Gallery(dim, METRIC_L2)
    gal.append(G)
    assert np.array_equal(gal.scores(Q), D)                                # dense dump bit-exact
    for thresholded in (True, False):
        widx, wd, wcnt = o.l2_match_topk(G, Q, tau, k, thresholded)
        fl = 0 if thresholded else MATCH_NO_THRESHOLD
        for path in (MATCH_FORCE_SCAN, MATCH_FORCE_GEMM, MATCH_FORCE_EXACT):
            idx, sc, cnt, st = gal.match(Q, tau, k, fl | path)
            assert np.array_equal(cnt, wcnt.astype(np.int32)), (path, thresholded)
            assert np.array_equal(idx, widx), (path, thresholded)
            assert np.array_equal(sc, wd.astype(np.float32), equal_nan=True)
            assert np.all(st <= 1)


@pytest.mark.parametrize("scale", [1e-6, 3e-3, 40.0, 2e5])
def test_extreme_scales_stay_exact(scale):
    """The prefilter bound is computed from the MEASURED fp16 rounding error of the gallery
    and the query (Cauchy-Schwarz), so it stays rigorous for subnormal-range and
    overflowing (fp16 inf) encodings; those queries are answered by the fp32 fallback."""
    rng = np.random.default_rng(int(scale * 1e7) % 1000)
    N, dim, B, k = 3000, 128, 6, 5
    G = (rng.normal(0, 1.0, (N, dim)) * scale).astype(np.float32)
    Q = (G[rng.integers(0, N, B)] + rng.normal(0, 0.3, (B, dim)) * scale).astype(np.float32)
    D = np.stack([np.linalg.norm(G.astype(np.float64) - q.astype(np.float64), axis=1) for q in Q])
    tau = float(np.sort(D[0])[12])
    widx, wd, wcnt = o.l2_match_topk(G, Q, tau, k, True)
    gal = Gallery(dim, METRIC_L2)
    gal.append(G)
    for path in (MATCH_FORCE_SCAN, MATCH_FORCE_GEMM, 0):
        idx, sc, cnt, st = gal.match(Q, tau, k, path)
        assert np.array_equal(cnt, wcnt.astype(np.int32)) and np.array_equal(idx, widx), (scale, path)
        for b in range(B):
            assert np.allclose(sc[b, :cnt[b]], wd[b, :cnt[b]], rtol=2e-6, atol=0)
        assert np.all(st <= 1)
    # cosine is scale-free after normalisation
    gc = Gallery(dim, METRIC_COS)
    gc.append(G)
    S = o.cosine_scores(G, Q)
    widx, wsc, wcnt = o.cos_match_topk(G, Q, 0.5, k, True)
    idx, sc, cnt, st = gc.match(Q, 0.5, k)
    assert np.array_equal(idx, widx) and np.array_equal(cnt, wcnt.astype(np.int32)) and np.all(st == 0)


# ----------------------------------------------------------------------------
# edge cases
# ----------------------------------------------------------------------------
def test_empty_and_tiny_galleries():
    gal = Gallery(128, METRIC_L2)
    q = np.zeros((2, 128), np.float32)
    idx, sc, cnt, st = gal.match(q, 0.6, 5)
    assert np.all(cnt == 0) and np.all(idx == -1) and np.all(np.isnan(sc))
    assert gal.scores(q).shape == (2, 0)
    fs = FaceService(GalleryDict(), raise_errors=True)
    assert fs.compare_faces(q[0]) == [] and fs.find_k_nearest(q[0]) == [] and fs.batch_compare_faces([q[0]]) == [[]]
    gal.append(np.ones((3, 128)) * 0.01)
    idx, sc, cnt, st = gal.match(q, 0.6, 5)            # k > N
    assert np.all(cnt == 3) and idx[0, :3].tolist() == [0, 1, 2]      # exact ties: insertion order
    idx, sc, cnt, st = gal.match(q, 0.6, 2, MATCH_NO_THRESHOLD)
    assert idx[0].tolist() == [0, 1]


def test_many_exact_duplicates_take_the_certified_fallback():
    rng = np.random.default_rng(9)
    G = rng.normal(0, 0.05, (4000, 128)).astype(np.float32)
    G[1000:1100] = G[7]                               # 101 identical rows >> per-list capacity
    gal = Gallery(128, METRIC_L2)
    gal.append(G)
    idx, sc, cnt, st = gal.match(G[7:8], 0.1, 10)
    assert idx[0].tolist() == [7] + list(range(1000, 1009))
    assert np.all(sc[0] == 0.0)
    idx, sc, cnt, st = gal.match(G[7:8], 0.1, 64)
    assert idx[0].tolist() == [7] + list(range(1000, 1063)) and cnt[0] == 64


def test_tombstones_updates_and_gather():
    rng = np.random.default_rng(10)
    G = rng.normal(0, 0.05, (3000, 128))
    gal = Gallery(128, METRIC_L2)
    gal.append(G)
    q = (G[123] + 0.001).astype(np.float32)[None]
    idx, sc, cnt, st = gal.match(q, 0.6, 3, MATCH_NO_THRESHOLD)
    assert idx[0, 0] == 123
    gal.remove(123)
    with pytest.raises(frp_b200.FrpError):
        gal.remove(123)
    idx2, sc2, cnt2, _ = gal.match(q, 0.6, 3, MATCH_NO_THRESHOLD)
    assert 123 not in idx2[0].tolist() and idx2[0, 0] == idx[0, 1]
    assert np.isnan(gal.scores(q)[0, 123])
    gal.update(5, G[123])
    idx3, _, _, _ = gal.match(q, 0.6, 3, MATCH_NO_THRESHOLD)
    assert idx3[0, 0] == 5
    assert gal.counts() == (3000, 2999)
    sub = gal.gather([5, 7, 5, 2999])
    i4, s4, c4, _ = sub.match(q, 0.6, 4, MATCH_NO_THRESHOLD)
    assert i4[0, :2].tolist() == [0, 2] and sub.counts() == (4, 4)
    assert np.allclose(sub.get_row(1), G[7].astype(np.float32))


def test_snapshot_roundtrip(tmp_path):
    G = o.synth_cos_gallery(2500, 512, seed=3)
    Q, _ = o.synth_cos_queries(G, 4, seed=4)
    gal = Gallery(512, METRIC_COS)
    gal.append(G)
    gal.remove(17)
    p = str(tmp_path / "gal.frp")
    gal.save(p)
    g2 = Gallery.load(p)
    assert g2.counts() == gal.counts() and g2.metric == METRIC_COS and g2.dim == 512
    a = gal.match(Q, 0.3, 10)
    b = g2.match(Q, 0.3, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1], b[1], equal_nan=True)
    assert a[2].sum() >= 2


def test_documents_and_named_snapshot(tmp_path):
    """Row N3: start-up load from db.py-style documents (Fernet), then a named snapshot."""
    from cryptography.fernet import Fernet
    from frp_b200.persistence import encode_embedding, load_faces, load_snapshot, save_snapshot
    G = o.synth_l2_gallery(400, 128, seed=31)
    key = Fernet.generate_key()
    docs = [{"target": "p%03d" % i, "embedding": encode_embedding(g.tolist(), key)} for i, g in enumerate(G)]
    enc = GalleryDict()
    assert load_faces(enc, docs, key) == 400
    fs = FaceService(enc, raise_errors=True)
    ref = {"p%03d" % i: g.tolist() for i, g in enumerate(G)}
    q = G[7] + 0.002
    want = o.compare_faces(ref, q, 0.6)
    got = fs.compare_faces(q)
    assert [r["target"] for r in got] == [r["target"] for r in want] and got[0]["target"] == "p007"
    del enc["p003"]
    p = str(tmp_path / "gallery.frp")
    save_snapshot(enc, p)
    enc2 = load_snapshot(p)
    assert enc2.keys() == enc.keys() and "p003" not in enc2
    got2 = FaceService(enc2, raise_errors=True).compare_faces(q)
    assert [(r["target"], r["distance"]) for r in got2] == [(r["target"], r["distance"]) for r in fs.compare_faces(q)]


def test_concurrent_callers():
    # routes/camera.py:30 calls compare_faces from a 4-thread pool
    from concurrent.futures import ThreadPoolExecutor
    G = o.synth_cos_gallery(20_000, 512, seed=8)
    Q, _ = o.synth_cos_queries(G, 32, seed=9)
    gal = Gallery(512, METRIC_COS)
    gal.append(G)
    want = gal.match(Q, 0.4, 5)

    def run(b):
        return gal.match(Q[b:b + 1], 0.4, 5)

    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(run, range(32)))
    for b, r in enumerate(res):
        assert np.array_equal(r[0][0], want[0][b]) and np.array_equal(r[2][0], want[2][b])


# ----------------------------------------------------------------------------
# BASELINE.json's full size (10M x 512, batch 256, top-10): size-independent properties
# ----------------------------------------------------------------------------
def test_full_size_10m_properties():
    import torch
    N, D, B, k, tau = 10_000_000, 512, 256, 10, 0.40
    gen = torch.Generator(device="cuda").manual_seed(11)
    full = Gallery(D, METRIC_COS, capacity=N)
    halves = [Gallery(D, METRIC_COS, capacity=N // 2) for _ in range(2)]
    halves[1].set_base(N // 2)
    chunk = 1_000_000
    keep = []
    for lo in range(0, N, chunk):
        x = torch.randn((chunk, D), generator=gen, device="cuda", dtype=torch.float32)
        full.append_device(x)
        halves[lo // (N // 2)].append_device(x)
        keep.append(x[:16].cpu().numpy())
        del x
    torch.cuda.synchronize()
    rng = np.random.default_rng(12)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    planted = {}
    for j in range(0, B, 2):                       # every other query is planted
        c = int(rng.integers(0, N // chunk))
        r = int(rng.integers(0, 16))
        base = keep[c][r] / np.linalg.norm(keep[c][r])
        a = rng.uniform(0.6, 0.95)
        Q[j] = a * base + np.sqrt(1 - a * a) * Q[j] / np.linalg.norm(Q[j])
        want = float(Q[j].astype(np.float64) @ base.astype(np.float64) / np.linalg.norm(Q[j].astype(np.float64)))
        planted[j] = (c * chunk + r, want)
    idx, sc, cnt, st = full.match(Q, tau, k)
    assert frp_b200.last_match_path() == "gemm_tcgen05" and np.all(st == 0)
    for j in range(B):
        if j in planted:                           # planted recall, score ~ alpha
            assert idx[j, 0] == planted[j][0] and abs(sc[j, 0] - planted[j][1]) < 1e-5, j
        else:                                      # impostors stay below 0.40 at 10M rows
            assert cnt[j] == 0
    # pure top-k: prefix / monotone / batch-1 scan path agrees with the tensor-core path
    i2, s2, c2, st2 = full.match(Q, tau, k, MATCH_NO_THRESHOLD)
    assert np.all(st2 == 0) and np.all(c2 == k) and np.all(np.diff(s2, axis=1) <= 0)
    for j in range(B):
        assert np.array_equal(i2[j, :cnt[j]], idx[j, :cnt[j]])
    for j in (0, 1, 101, 254):
        i1, s1, c1, _ = full.match(Q[j:j + 1], tau, k, MATCH_NO_THRESHOLD | MATCH_FORCE_SCAN)
        assert frp_b200.last_match_path() == "scan_f16"
        assert np.array_equal(i1[0], i2[j]) and np.array_equal(s1[0], s2[j])
        # a gallery this large routes single queries through the tcgen05 kernel by default
        i0, s0, c0, _ = full.match(Q[j:j + 1], tau, k, MATCH_NO_THRESHOLD)
        assert frp_b200.last_match_path() == "gemm_tcgen05"
        assert np.array_equal(i0[0], i2[j]) and np.array_equal(s0[0], s2[j])
    # shard invariance at full size: two 5M shards merged == one 10M gallery
    parts = [h.match(Q, tau, k, MATCH_NO_THRESHOLD) for h in halves]
    for j in range(B):
        ent = [(-float(p[1][j, t]), int(p[0][j, t])) for p in parts for t in range(int(p[2][j]))]
        ent.sort()
        assert [e[1] for e in ent[:k]] == i2[j].tolist()
        assert np.array_equal(np.array([-e[0] for e in ent[:k]], np.float32), s2[j])


# ----------------------------------------------------------------------------
# size-independent properties at a large size (1M x 512; oracle too slow here)
# ----------------------------------------------------------------------------
def test_properties_large():
    import torch
    N, D, B, k, tau = 1_000_000, 512, 8, 10, 0.40
    gen = torch.Generator(device="cuda").manual_seed(1)
    gal = Gallery(D, METRIC_COS, capacity=N)
    chunk = 250_000
    keep = {}
    for lo in range(0, N, chunk):
        x = torch.randn((chunk, D), generator=gen, device="cuda", dtype=torch.float32)
        gal.append_device(x)
        keep[lo] = x[:4].cpu().numpy()
    torch.cuda.synchronize()
    rng = np.random.default_rng(2)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    planted_rows = [0, 250_001, 500_002, 750_003]
    for j, r in enumerate(planted_rows):
        base = keep[(r // chunk) * chunk][r % chunk]
        base = base / np.linalg.norm(base)
        n = Q[j] / np.linalg.norm(Q[j])
        Q[j] = 0.8 * base + 0.6 * n
    idx, sc, cnt, st = gal.match(Q, tau, k)
    assert np.all(st == 0)
    for j, r in enumerate(planted_rows):                 # planted recall
        assert idx[j, 0] == r and cnt[j] >= 1 and abs(sc[j, 0] - 0.8) < 0.05
    assert np.all(cnt[len(planted_rows):] == 0)          # impostors stay below 0.40
    # threshold monotonicity + top-k prefix property + path equivalence
    i2, s2, c2, _ = gal.match(Q, 0.0, k, MATCH_NO_THRESHOLD)
    for b in range(B):
        assert np.array_equal(i2[b, :cnt[b]], idx[b, :cnt[b]])
        assert np.all(np.diff(s2[b]) <= 0)
    i3, s3, c3, _ = gal.match(Q, 0.0, 5, MATCH_NO_THRESHOLD)
    assert np.array_equal(i3, i2[:, :5])
    i4, s4, c4, _ = gal.match(Q, 0.0, k, MATCH_NO_THRESHOLD | MATCH_FORCE_EXACT)
    assert np.array_equal(i4, i2) and np.array_equal(s4, s2)
    # the dense dump agrees with the top-k on the winners
    dense = gal.scores(Q[:2])
    for b in range(2):
        top = np.argsort(-dense[b], kind="stable")[:k]
        assert np.array_equal(top, i2[b])
        assert np.abs(dense[b][top] - s2[b]).max() < TIGHT
