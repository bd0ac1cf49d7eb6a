"""The oracle against the golden vectors (reference outputs / installed OpenCV)."""
import hashlib
import os

import numpy as np
import pytest

import frp_oracle_c as oc
import frp_oracle_np as o
import ref_loader

LVL = {"high": 0, "medium": 1, "low": 2}


def _enc(G):
    return {"id%07d" % i: g.tolist() for i, g in enumerate(G)}


def test_np_compare_faces_matches_reference_golden(golden_l2):
    g = golden_l2
    enc = _enc(g["G"])
    names = list(enc.keys())
    for b in range(0, len(g["Q"]), 3):
        res = o.compare_faces(enc, g["Q"][b], 0.6)
        assert [names.index(r["target"]) for r in res] == g["order"][b].tolist()
        assert np.array_equal(np.array([r["distance"] for r in res]), g["dist"][b])      # bit-exact
        assert np.array_equal(np.array([r["match"] for r in res]), g["match"][b])
        assert np.array_equal(np.array([r["confidence_score"] for r in res]), g["score"][b])
        assert [LVL[r["confidence"]] for r in res] == g["conf"][b].tolist()


def test_np_topk_and_batch_match_reference_golden(golden_l2):
    g = golden_l2
    idx, dist, cnt = o.l2_match_topk(g["G"], g["Q"], 0.6, 5, True)
    for b in range(len(g["Q"])):
        want = g["order"][b][g["match"][b]][:5]
        assert idx[b, :cnt[b]].tolist() == want.tolist()
    enc = _enc(g["G"])
    names = list(enc.keys())
    bat = o.batch_compare_faces(enc, list(g["Q"]), 0.6)
    for b, r in enumerate(bat):
        assert len(r) == g["bat_cnt"][b]
        assert [names.index(x["target"]) for x in r] == g["bat_rows"][b, :len(r)].tolist()
    # find_k_nearest: pure top-k; argpartition leaves tie order unspecified, so compare distances
    idx, dist, cnt = o.l2_match_topk(g["G"], g["Q"], 0.0, 5, False)
    assert np.allclose(dist, g["knn_d"], rtol=0, atol=0)


def test_np_subset_semantics(golden_l2):
    g = golden_l2
    enc = _enc(g["G"])
    names = list(enc.keys())
    sub = [names[r] if r >= 0 else "nope" for r in g["sub_query_rows"]]
    res = o.compare_faces(enc, g["Q"][0], 0.6, sub)
    assert [names.index(r["target"]) for r in res] == g["sub_rows"].tolist()
    assert np.array_equal(np.array([r["distance"] for r in res]), g["sub_dist"])


def test_c_oracle_l2_matches_reference_golden(golden_l2):
    g = golden_l2
    for b in range(len(g["Q"])):
        d = oc.l2_distances(g["G"], g["Q"][b])
        # sqrt(sum of squares) in a different summation order than numpy's pairwise sum
        assert np.allclose(np.sort(d), g["dist"][b], rtol=0, atol=1e-14)
        idx, val, n = oc.topk(d, 0, 1, 0.6, 5)
        want = g["order"][b][g["match"][b]][:5]
        # exact ties / threshold hits at the 1e-16 level may legitimately differ; they do not
        # occur away from the planted 0.6 +- 0 rows
        near = np.abs(g["dist"][b][g["match"][b]][:5] - 0.6) < 1e-12
        if not near.any():
            assert idx[:n].tolist() == want.tolist()


def test_c_oracle_cos_matches_np(golden_cos):
    g = golden_cos
    for b in range(len(g["Q"])):
        s = oc.cos_scores(g["G"], g["Q"][b])
        assert np.allclose(s, g["all_scores"][b], rtol=0, atol=1e-12)
    idx, sc, cnt = o.cos_match_topk(g["G"], g["Q"], float(g["tau"]), int(g["k"]), True)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(cnt, g["count"])


def test_warp_restatements_match_cv2_golden(golden_align):
    g = golden_align
    for i in range(len(g["frame_idx"])):
        img = g["frames"][g["frame_idx"][i]]
        a = o.warp_affine_fixed(img, g["M"][i])
        c = oc.warp_affine(img, g["M"][i])
        assert hashlib.sha256(a.tobytes()).hexdigest() == str(g["sha_aligned"][i])
        assert hashlib.sha256(c.tobytes()).hexdigest() == str(g["sha_aligned"][i])
        b = o.blob_from_aligned(a)
        assert hashlib.sha256(b.tobytes()).hexdigest() == str(g["sha_blob"][i])
        assert np.array_equal(oc.blob(c).astype(np.float16), b)
    assert np.allclose(np.stack([o.estimate_norm(l) for l in g["lmk"]]), g["M"], rtol=0, atol=1e-9)


def test_warp_restatement_matches_installed_cv2_live():
    cv2 = pytest.importorskip("cv2")
    frames = o.synth_frames(1, 360, 640, seed=3)
    fi, lm = o.synth_detections(1, 40, 360, 640, seed=11, sigma_range=(0.3, 1.5), oob_frac=0.2)
    for i in range(len(fi)):
        M = o.estimate_norm(lm[i])
        want = cv2.warpAffine(frames[0], M, (112, 112), borderValue=0.0)
        assert np.array_equal(o.warp_affine_fixed(frames[0], M), want)
        assert np.array_equal(oc.warp_affine(frames[0], M), want)
        blob = cv2.dnn.blobFromImages([want], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
        assert np.array_equal(o.blob_from_aligned(want), blob.astype(np.float16))


def test_norm_lut_matches_cv2_golden(golden_align):
    v = np.arange(256, dtype=np.float32)
    lut = (v - np.float32(127.5)) * np.float32(1.0 / 127.5)
    assert np.array_equal(lut, golden_align["lut_f32"])
    assert np.array_equal(lut.astype(np.float16), golden_align["lut_f16"])


def test_norm_fma_form_matches_cv2(golden_align):
    """The kernel normalises with ONE fused multiply-add, fp16(fma(v, f32(1/127.5), -1));
    exhaustively equal to cv2.dnn.blobFromImages' (v - 127.5) * f32(1/127.5) after the
    fp16 cast."""
    v = np.arange(256, dtype=np.float64)
    c = np.float64(np.float32(1.0 / 127.5))
    fma = (v * c - 1.0).astype(np.float32).astype(np.float16)     # exact product, one rounding
    assert np.array_equal(fma, golden_align["lut_f16"])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_live_reference_agrees_with_np_oracle():
    fs, ENC, _ = ref_loader.load_reference()
    G = o.synth_l2_gallery(300, 128, seed=5)
    Q, _ = o.synth_l2_queries(G, 6, seed=6)
    ENC.clear()
    for i, g in enumerate(G):
        ENC["id%07d" % i] = g.tolist()
    try:
        fs.tolerance = 0.6
        for q in Q:
            assert fs.compare_faces(q) == o.compare_faces(ENC, q, 0.6)
        assert fs.batch_compare_faces(list(Q)) == o.batch_compare_faces(ENC, list(Q), 0.6)
    finally:
        ENC.clear()


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_live_reference_clustering_agrees_with_np_oracle():
    """cluster_faces (face_service.py:552-585) of the UNMODIFIED reference on a clustered gallery
    with chains (A~B, B~C, A!~C) equals the vectorised restatement the GPU tests use at 20 k rows."""
    fs, ENC, _ = ref_loader.load_reference()
    G = o.synth_identity_gallery(25, 6, seed=3, chain_every=4)
    ENC.clear()
    for i, g in enumerate(G):
        ENC["id%07d" % i] = g.tolist()
    try:
        for thr in (0.6, 0.45):
            want = fs.cluster_faces(thr)
            got = o.cluster_faces(ENC, thr)
            assert got == want
            assert 20 <= len(want) < len(G)
        # a chain really is split by the greedy order somewhere
        sizes = sorted(len(v) for v in fs.cluster_faces(0.6).values())
        assert sizes[0] >= 1 and sizes[-1] >= 6
    finally:
        ENC.clear()


def test_umeyama_cross_checked_against_scipy_and_opencv():
    """scikit-image (whose ``_umeyama`` insightface calls) is not in this image, so the
    restatement cannot be pinned to it; it is cross-checked against two independent
    implementations of the same least-squares similarity instead: SciPy's orthogonal
    Procrustes solution and OpenCV's ``estimateAffinePartial2D`` (all five points inliers,
    LM-refined), plus the first-order optimality of the residual."""
    cv2 = pytest.importorskip("cv2")
    from scipy.linalg import orthogonal_procrustes
    rng = np.random.default_rng(77)
    dst = o.ARCFACE_DST.astype(np.float64)
    for _ in range(40):
        th, sg = rng.uniform(-0.9, 0.9), rng.uniform(0.3, 4.0)
        R0 = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        src = ((dst - 56.0) @ R0.T) * sg + rng.uniform(100, 900, 2) + rng.normal(0, 1.5, (5, 2))
        src32 = src.astype(np.float32)
        M = o.estimate_norm(src32)                       # 2x3, maps src -> dst
        s = src32.astype(np.float64)
        sc, dc = s - s.mean(0), dst - dst.mean(0)
        Rp, _ = orthogonal_procrustes(sc, dc)            # min ||sc Rp - dc||_F, Rp orthogonal
        assert np.linalg.det(Rp) > 0
        scale = np.trace(Rp.T @ sc.T @ dc) / np.sum(sc * sc)
        A = scale * Rp.T
        Mp = np.hstack([A, (dst.mean(0) - A @ s.mean(0))[:, None]])
        assert np.abs(M - Mp).max() < 1e-9
        Mc, inl = cv2.estimateAffinePartial2D(s, dst, method=cv2.LMEDS, refineIters=50)
        if inl is not None and int(inl.sum()) == 5:
            assert np.abs(M - Mc).max() < 5e-3 * max(1.0, np.abs(M).max())
        # least-squares optimum: no first-order decrease along any of the 4 similarity parameters
        def resid(Mx):
            return float(np.sum((s @ Mx[:, :2].T + Mx[:, 2] - dst) ** 2))
        r0 = resid(M)
        a, b, tx, ty = M[0, 0], M[1, 0], M[0, 2], M[1, 2]
        for da, db, dx, dy in [(1e-5, 0, 0, 0), (0, 1e-5, 0, 0), (0, 0, 1e-3, 0), (0, 0, 0, 1e-3)]:
            for sgn in (1, -1):
                a2, b2 = a + sgn * da, b + sgn * db
                M2 = np.array([[a2, -b2, tx + sgn * dx], [b2, a2, ty + sgn * dy]])
                assert resid(M2) >= r0 - 1e-9


def test_fast_topk_oracles_equal_the_plain_ones():
    """The shortlist forms used at BASELINE sizes (1 M x 512, 2 M x 128) give exactly the plain
    restatements' answers, including exact duplicates and threshold ties."""
    G = o.synth_cos_gallery(30_000, 64, seed=1)
    G[20_000] = G[7]
    G[29_999] = G[7]
    Q, _ = o.synth_cos_queries(G, 9, seed=2)
    Q[0] = G[7]
    for thr in (True, False):
        a = o.cos_match_topk(G, Q, 0.35, 10, thr, chunk=8192)
        b = o.cos_match_topk_fast(G, Q, 0.35, 10, thr, chunk=8192)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        assert np.array_equal(a[1], b[1], equal_nan=True)
    G2 = o.synth_l2_gallery(20_000, 128, seed=3)
    G2[15_000] = G2[3]
    Q2, _ = o.synth_l2_queries(G2, 7, seed=4)
    Q2[0] = G2[3]
    for thr in (True, False):
        a = o.l2_match_topk(G2, Q2, 0.6, 10, thr)
        b = o.l2_match_topk_fast(G2, Q2, 0.6, 10, thr, chunk=4096)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        assert np.array_equal(a[1], b[1], equal_nan=True)


def test_float32_umeyama_restatement():
    """umeyama_f32 (what insightface's float32 inputs make skimage compute) stays within float32
    round-off of the float64 estimate and is a proper similarity."""
    fi, lm = o.synth_detections(2, 40, seed=9)
    for l in lm:
        a, b = o.estimate_norm(l), o.estimate_norm_f32(l)
        assert np.abs(a - b).max() / np.abs(a).max() < 2e-6
        assert abs(b[0, 0] - b[1, 1]) < 1e-9 * abs(b[0, 0]) + 1e-12 and abs(b[0, 1] + b[1, 0]) < 1e-9 * abs(b[0, 0]) + 1e-9
