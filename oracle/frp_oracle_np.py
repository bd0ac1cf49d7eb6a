"""CPU oracle (numpy) for the recognition hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(``face-recognition-platform_b200/``) never imports it and has no CPU fallback.

What is restated, and from where
--------------------------------
* 128-d L2 matcher: the reference's own code,
  ``backend/app/services/face_service.py:395-443`` (compare_faces),
  ``:448-481`` (batch_compare_faces), ``:486-506`` (confidence helpers),
  ``:590-612`` (find_k_nearest).  The arithmetic lives in the un-vendored
  dependency ``face-recognition==1.3.0`` (``backend/requirements.txt:37``):
  ``api.face_distance(encs, q) = np.linalg.norm(encs - q, axis=1)`` and is also
  restated in-repo at ``face_service.py:597``.
* 512-d cosine matcher + 5-point alignment: NOT in the reference source; they
  run inside ``insightface==0.7.3`` (``backend/requirements.txt:41``; only call
  site ``backend/app/utils/deepfake_utils.py:39-48,138``).  Published algorithm
  restated here: ``face_align.estimate_norm`` (skimage ``_umeyama``),
  ``face_align.norm_crop`` (``cv2.warpAffine(img, M, (112,112), borderValue=0)``),
  ``ArcFaceONNX.get_feat`` (``cv2.dnn.blobFromImages(1/127.5, mean 127.5,
  swapRB=True)``), ``ArcFaceONNX.compute_sim`` / ``Face.normed_embedding``.

Pinning status
--------------
The reference has no tests, golden vectors or fixtures, so:
* L2 path: pinned against the UNMODIFIED reference ``face_service.py`` imported
  in the authoring container (``oracle/ref_loader.py``), vectors committed under
  ``tests/golden/`` by ``oracle/gen_golden.py``.
* warp: pinned against the installed ``cv2.warpAffine`` (OpenCV 4.13.0) — the
  fixed-point restatement below is compared with cv2 itself in the CPU tests.
* Umeyama / cosine: "parity unpinned" — insightface / scikit-image are not
  installed anywhere in this image; the restatement follows their published
  algorithm only.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# 1. 128-d L2 matcher (face_recognition / face_service.py)
# --------------------------------------------------------------------------

def face_distance(face_encodings, face_to_compare):
    """face_recognition.api.face_distance (1.3.0); in-repo restatement at
    face_service.py:597."""
    face_encodings = np.asarray(face_encodings)
    if len(face_encodings) == 0:
        return np.empty((0))
    return np.linalg.norm(face_encodings - face_to_compare, axis=1)


def confidence_level(distance: float) -> str:
    """face_service.py:486-492."""
    if distance < 0.4:
        return "high"
    elif distance < 0.6:
        return "medium"
    return "low"


def calibrate_confidence(distance: float) -> float:
    """face_service.py:497-506."""
    x = max(0.0, min(1.0, 1.0 - distance))
    calibrated = 100.0 / (1.0 + np.exp(-12.0 * (x - 0.5)))
    return round(float(calibrated), 2)


def compare_faces(encodings: dict, test_encoding, tolerance=0.6, target_names=None,
                  return_distances=True):
    """face_service.py:395-443: one dict per checked target, inclusive <=,
    stable ascending sort (ties keep dict insertion order)."""
    targets = list(encodings.keys()) if target_names is None else \
        [t for t in target_names if t in encodings]
    if not targets:
        return []
    stored = np.array([encodings[t] for t in targets])
    distances = face_distance(stored, np.asarray(test_encoding))
    matches = distances <= tolerance
    results = []
    for i, target in enumerate(targets):
        d = float(distances[i])
        item = {"target": target, "match": bool(matches[i])}
        if return_distances:
            item["distance"] = d
            item["confidence"] = confidence_level(d)
            item["confidence_score"] = calibrate_confidence(d)
        results.append(item)
    if return_distances:
        results.sort(key=lambda x: x.get("distance", 1.0))
    return results


def batch_compare_faces(encodings: dict, test_encodings, tolerance=0.6, target_names=None):
    """face_service.py:448-481: matches only, sorted ascending."""
    targets = list(encodings.keys()) if target_names is None else \
        [t for t in target_names if t in encodings]
    if not targets:
        return [[] for _ in test_encodings]
    stored = np.array([encodings[t] for t in targets])
    out = []
    for q in test_encodings:
        distances = face_distance(stored, np.asarray(q))
        matches = distances <= tolerance
        res = [{"target": t, "match": True, "distance": float(distances[i]),
                "confidence": confidence_level(distances[i])}
               for i, t in enumerate(targets) if matches[i]]
        res.sort(key=lambda x: x["distance"])
        out.append(res)
    return out


def l2_match_topk(G, Q, tau, k, thresholded=True):
    """Array form of routes/face.py:692-693 (threshold THEN top-k) over
    face_service.py:410-432.  Returns (idx[B,k] int64 (-1 pad), dist[B,k] f64
    (nan pad), count[B]).  Order: distance ascending, ties by row index
    (= dict insertion order; the reference's sort is stable)."""
    G = np.asarray(G, dtype=np.float64)
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    B = Q.shape[0]
    idx = np.full((B, k), -1, np.int64)
    dist = np.full((B, k), np.nan, np.float64)
    cnt = np.zeros(B, np.int64)
    for b in range(B):
        d = np.linalg.norm(G - Q[b], axis=1) if len(G) else np.empty(0)
        order = np.argsort(d, kind="stable")
        if thresholded:
            order = order[d[order] <= tau]
        order = order[:k]
        cnt[b] = len(order)
        idx[b, :len(order)] = order
        dist[b, :len(order)] = d[order]
    return idx, dist, cnt


def cluster_faces(encodings: dict, distance_threshold=0.6):
    """face_service.py:552-585 `cluster_faces`: faces in dict order; an unassigned face seeds
    `cluster_<n>` and absorbs every still-unassigned face j with
    face_distance([enc_i], enc_j) <= threshold (inclusive), members in dict order after the
    seed.  Fewer than two faces -> one cluster with everything.  The reference evaluates one
    pair per Python call; this restatement evaluates one seed row at a time (same float64
    arithmetic: np.linalg.norm of the row difference)."""
    if len(encodings) < 2:
        return {"cluster_0": list(encodings.keys())}
    targets = list(encodings.keys())
    E = np.array([encodings[t] for t in targets], dtype=np.float64)
    assigned = np.zeros(len(targets), bool)
    clusters = {}
    cid = 0
    for i, t in enumerate(targets):
        if assigned[i]:
            continue
        assigned[i] = True
        d = np.linalg.norm(E - E[i], axis=1)
        take = (~assigned) & (d <= distance_threshold)
        clusters["cluster_%d" % cid] = [t] + [targets[j] for j in np.nonzero(take)[0]]
        assigned |= take
        cid += 1
    return clusters


def synth_identity_gallery(n_ids, per_id, dim=128, seed=7, spread=0.02, chain_every=0):
    """Clustered 128-d encodings: `n_ids` identity centres ~ N(0, 0.05), `per_id` samples each
    at N(centre, spread) (intra-identity distance ~ spread * sqrt(2 dim) = 0.32 for 0.02,
    inter-identity ~ 0.8), shuffled so members are spread over the whole row range.  With
    chain_every = m > 0 every m-th identity gets a CHAIN A ~ B ~ C: B at 0.5 from A, C at 0.5
    from B on the same ray (so A-C = 1.0 > 0.6): greedy seed order decides where B lands."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0, 0.05, (n_ids, dim))
    rows = [centres[i] + rng.normal(0, spread, (per_id, dim)) for i in range(n_ids)]
    extra = []
    if chain_every:
        for i in range(0, n_ids, chain_every):
            u = rng.standard_normal(dim)
            u /= np.linalg.norm(u)
            a = rows[i][0]
            extra.append(a + 0.5 * u)
            extra.append(a + 1.0 * u)
    G = np.concatenate(rows + ([np.array(extra)] if extra else []), axis=0)
    return G[rng.permutation(len(G))]


# --------------------------------------------------------------------------
# 2. 512-d cosine matcher (insightface ArcFaceONNX.compute_sim)
# --------------------------------------------------------------------------

def normed(x):
    """insightface app/common.py Face.normed_embedding = e / ||e||_2."""
    x = np.asarray(x, dtype=np.float64)
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def cosine_scores(G, Q):
    """ArcFaceONNX.compute_sim for every (query, gallery row), float64."""
    return normed(np.atleast_2d(Q)) @ normed(G).T


def cos_match_topk(G, Q, tau, k, thresholded=True, chunk=262144):
    """Cosine 1:N: match iff sim >= tau; top-k by similarity descending, ties
    by row index.  Same return convention as l2_match_topk (score instead of
    distance).  Chunked over gallery rows to bound memory."""
    Qn = normed(np.atleast_2d(Q))
    B = Qn.shape[0]
    N = len(G)
    best_s = np.full((B, 0), 0.0)
    best_i = np.zeros((B, 0), np.int64)
    for lo in range(0, N, chunk):
        Gn = normed(G[lo:lo + chunk])
        s = Qn @ Gn.T
        ii = np.broadcast_to(np.arange(lo, lo + s.shape[1]), s.shape)
        s = np.concatenate([best_s, s], axis=1)
        ii = np.concatenate([best_i, ii], axis=1)
        # stable: sort by (-score, idx); idx already ascending inside each part
        order = np.lexsort((ii, -s), axis=1)[:, :max(k, 1) * 4 + 64]
        best_s = np.take_along_axis(s, order, axis=1)
        best_i = np.take_along_axis(ii, order, axis=1)
    idx = np.full((B, k), -1, np.int64)
    sc = np.full((B, k), np.nan, np.float64)
    cnt = np.zeros(B, np.int64)
    for b in range(B):
        s, ii = best_s[b], best_i[b]
        if thresholded:
            keep = s >= tau
            s, ii = s[keep], ii[keep]
        s, ii = s[:k], ii[:k]
        cnt[b] = len(s)
        idx[b, :len(s)] = ii
        sc[b, :len(s)] = s
    return idx, sc, cnt


def _shortlist(score_desc, m):
    """Per row: indices of every entry >= the m-th largest value (ties included)."""
    n = score_desc.shape[1]
    if n <= m:
        return [np.arange(n)] * score_desc.shape[0]
    kth = np.partition(score_desc, n - m, axis=1)[:, n - m]
    return [np.nonzero(score_desc[b] >= kth[b])[0] for b in range(score_desc.shape[0])]


def cos_match_topk_fast(G, Q, tau, k, thresholded=True, chunk=262144):
    """Same result as cos_match_topk (same float64 chunk products, same order) for galleries of
    millions of rows: each chunk is shortlisted with a partition (all ties of the cut kept)
    instead of a full lexsort.  Checked equal to cos_match_topk in tests/test_oracle_cpu.py."""
    Qn = normed(np.atleast_2d(Q))
    B = Qn.shape[0]
    cs = [[] for _ in range(B)]
    ci = [[] for _ in range(B)]
    for lo in range(0, len(G), chunk):
        s = Qn @ normed(G[lo:lo + chunk]).T
        for b, sel in enumerate(_shortlist(s, k + 8)):
            cs[b].append(s[b, sel])
            ci[b].append(sel + lo)
    idx = np.full((B, k), -1, np.int64)
    sc = np.full((B, k), np.nan, np.float64)
    cnt = np.zeros(B, np.int64)
    for b in range(B):
        s_, i_ = np.concatenate(cs[b]), np.concatenate(ci[b])
        order = np.lexsort((i_, -s_))
        s_, i_ = s_[order], i_[order]
        if thresholded:
            keep = s_ >= tau
            s_, i_ = s_[keep], i_[keep]
        s_, i_ = s_[:k], i_[:k]
        cnt[b] = len(s_)
        idx[b, :len(s_)] = i_
        sc[b, :len(s_)] = s_
    return idx, sc, cnt


def l2_match_topk_fast(G, Q, tau, k, thresholded=True, chunk=262144):
    """Same result as l2_match_topk for galleries of millions of rows: a float64
    |g|^2 - 2 g.q + |q|^2 pass shortlists each chunk (everything within 1e-6 of the k-th best
    squared distance), then the shortlisted rows are re-evaluated with the reference's own
    expression np.linalg.norm(g - q) (face_service.py:597) before threshold and ordering."""
    G = np.asarray(G)
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    B = Q.shape[0]
    cand = [[] for _ in range(B)]
    qn = (Q * Q).sum(1)
    for lo in range(0, len(G), chunk):
        Gc = np.asarray(G[lo:lo + chunk], dtype=np.float64)
        d2 = (Gc * Gc).sum(1)[None, :] - 2.0 * (Q @ Gc.T) + qn[:, None]
        for b, sel in enumerate(_shortlist(-d2, k + 8)):
            kth = d2[b, sel].max()
            cand[b].append(np.nonzero(d2[b] <= kth + 1e-6)[0] + lo)
    idx = np.full((B, k), -1, np.int64)
    dist = np.full((B, k), np.nan, np.float64)
    cnt = np.zeros(B, np.int64)
    for b in range(B):
        rows = np.unique(np.concatenate(cand[b]))
        d = np.linalg.norm(np.asarray(G[rows], dtype=np.float64) - Q[b], axis=1)
        order = np.lexsort((rows, d))
        rows, d = rows[order], d[order]
        if thresholded:
            keep = d <= tau
            rows, d = rows[keep], d[keep]
        rows, d = rows[:k], d[:k]
        cnt[b] = len(rows)
        idx[b, :len(rows)] = rows
        dist[b, :len(rows)] = d
    return idx, dist, cnt


# --------------------------------------------------------------------------
# 3. Alignment (insightface face_align + OpenCV)
# --------------------------------------------------------------------------

# insightface/utils/face_align.py: arcface_dst (float32), image_size 112 -> ratio 1, diff_x 0
ARCFACE_DST = np.array(
    [[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
     [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float32)


def umeyama(src, dst):
    """skimage.transform._geometric._umeyama(src, dst, estimate_scale=True)
    restated in float64 (SVD form, reflection guard, rank handling)."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    num, dim = src.shape
    src_mean = src.mean(axis=0)
    dst_mean = dst.mean(axis=0)
    src_demean = src - src_mean
    dst_demean = dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d)
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


def umeyama_f32(src, dst):
    """skimage _umeyama as insightface 0.7.3 really runs it: face_align.estimate_norm passes
    float32 landmarks and the float32 arcface_dst, numpy keeps float32 through the means, the
    deviations, A and the variance, LAPACK's float32 SVD (sgesdd) factors A, and only d, T and
    the composition are float64 (skimage allocates them as float64).  TEST INFRASTRUCTURE."""
    src = np.asarray(src, dtype=np.float32)
    dst = np.asarray(dst, dtype=np.float32)
    num, dim = src.shape
    src_mean = src.mean(axis=0)
    dst_mean = dst.mean(axis=0)
    src_demean = src - src_mean
    dst_demean = dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d)
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


def estimate_norm_f32(lmk):
    """insightface face_align.estimate_norm(lmk, 112) with the dtypes insightface uses."""
    return umeyama_f32(np.asarray(lmk, dtype=np.float32).reshape(5, 2), ARCFACE_DST)[0:2, :]


def estimate_norm(lmk):
    """face_align.estimate_norm(lmk, image_size=112): 2x3 forward affine."""
    lmk = np.asarray(lmk)
    assert lmk.shape == (5, 2)
    return umeyama(lmk, ARCFACE_DST)[0:2, :]


def _rint(x):
    # cvRound / saturate_cast<int>(double): round half to even (SSE2 cvtsd2si)
    return np.rint(x).astype(np.int64)


def invert_affine(M):
    """cv::warpAffine's in-place inversion of the forward 2x3 matrix (double),
    modules/imgproc/src/imgwarp.cpp (operation order kept)."""
    M = np.array(M, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[4] * D
    A22 = M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2] = b1
    M[5] = b2
    return M


def warp_affine_fixed(img, M, size=112):
    """cv2.warpAffine(img, M, (size,size), flags=INTER_LINEAR, BORDER_CONSTANT 0)
    for 8UC3 restated with OpenCV's fixed-point scheme: AB_BITS=10,
    INTER_BITS=5, INTER_REMAP_COEF_BITS=15."""
    img = np.asarray(img)
    H, W = img.shape[:2]
    iM = invert_affine(M)
    xs = np.arange(size, dtype=np.float64)
    adelta = _rint(iM[0] * xs * 1024.0)
    bdelta = _rint(iM[3] * xs * 1024.0)
    X0 = _rint((iM[1] * xs + iM[2]) * 1024.0) + 16          # indexed by y
    Y0 = _rint((iM[4] * xs + iM[5]) * 1024.0) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5                 # [y, x]
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx = X & 31
    fy = Y & 31
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] +
           tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None])
    return ((acc + 16384) >> 15).astype(np.uint8)


def norm_crop_cv2(img, lmk, size=112):
    """face_align.norm_crop with the INSTALLED OpenCV as the pixel oracle."""
    import cv2
    M = estimate_norm(lmk)
    return cv2.warpAffine(img, M, (size, size), borderValue=0.0)


def blob_from_aligned(aimg):
    """cv2.dnn.blobFromImages([aimg], 1/127.5, (112,112), (127.5,)*3, swapRB=True)
    -> CHW float32 (OpenCV computes (v-mean)*float(scale) in float32), then the
    north_star's fp16 cast (round to nearest even)."""
    x = aimg.astype(np.float32)
    x = (x - np.float32(127.5)) * np.float32(1.0 / 127.5)
    x = x[..., ::-1]                       # BGR -> RGB
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float16)


def align_batch(frames, frame_idx, lmks, size=112, use_cv2=True):
    """Whole alignment stage for a batch of detections.
    frames [F,H,W,3] u8 BGR; frame_idx [n]; lmks [n,5,2] f32.
    Returns (aligned_u8 [n,size,size,3], blob_f16 [n,3,size,size])."""
    n = len(frame_idx)
    al = np.empty((n, size, size, 3), np.uint8)
    bl = np.empty((n, 3, size, size), np.float16)
    for i in range(n):
        img = frames[int(frame_idx[i])]
        if use_cv2:
            al[i] = norm_crop_cv2(img, lmks[i], size)
        else:
            al[i] = warp_affine_fixed(img, estimate_norm(lmks[i]), size)
        bl[i] = blob_from_aligned(al[i])
    return al, bl


# --------------------------------------------------------------------------
# 4. Seeded synthetic inputs
# --------------------------------------------------------------------------

def synth_l2_gallery(n=1000, dim=128, seed=1234):
    rng = np.random.default_rng(seed)
    return rng.normal(0, 0.05, (n, dim))


def synth_l2_queries(G, nq, seed=4321, planted_frac=0.5):
    rng = np.random.default_rng(seed)
    n, dim = G.shape
    Q = rng.normal(0, 0.05, (nq, dim))
    planted = rng.random(nq) < planted_frac
    j = rng.integers(0, n, nq)
    Q[planted] = G[j[planted]] + rng.normal(0, 0.02, (int(planted.sum()), dim))
    return Q, np.where(planted, j, -1)


def synth_cos_gallery(n, dim=512, seed=1234):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, dim), dtype=np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    return G


def synth_cos_queries(G, nq, seed=4321, planted_frac=0.5):
    rng = np.random.default_rng(seed)
    n, dim = G.shape
    Q = rng.standard_normal((nq, dim), dtype=np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    planted = rng.random(nq) < planted_frac
    j = rng.integers(0, n, nq)
    alpha = rng.uniform(0.45, 0.95, nq).astype(np.float32)
    for b in np.nonzero(planted)[0]:
        q = alpha[b] * G[j[b]] + np.sqrt(1 - alpha[b] ** 2) * Q[b]
        Q[b] = q / np.linalg.norm(q)
    return Q, np.where(planted, j, -1)


def synth_frames(nframes=2, H=1080, W=1920, seed=99):
    import cv2
    rng = np.random.default_rng(seed)
    out = np.empty((nframes, H, W, 3), np.uint8)
    for f in range(nframes):
        raw = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        out[f] = cv2.GaussianBlur(raw, (0, 0), 3)
    return out


def synth_detections(nframes, per_frame, H=1080, W=1920, seed=7, sigma_range=(0.4, 3.0),
                     oob_frac=0.02):
    """Landmarks = ((arcface_dst - 56) @ R(theta)^T) * sigma + centre + N(0,1px)."""
    rng = np.random.default_rng(seed)
    n = nframes * per_frame
    frame_idx = np.repeat(np.arange(nframes, dtype=np.int32), per_frame)
    sigma = rng.uniform(*sigma_range, n)
    theta = rng.uniform(-0.5, 0.5, n)
    cx = rng.uniform(150, W - 150, n)
    cy = rng.uniform(150, H - 150, n)
    oob = rng.random(n) < oob_frac
    cx[oob] = rng.choice([10.0, W - 10.0], int(oob.sum()))
    cy[oob] = rng.choice([10.0, H - 10.0], int(oob.sum()))
    base = ARCFACE_DST.astype(np.float64) - 56.0
    lm = np.empty((n, 5, 2), np.float32)
    for i in range(n):
        c, s = np.cos(theta[i]), np.sin(theta[i])
        R = np.array([[c, -s], [s, c]])
        p = (base @ R.T) * sigma[i] + np.array([cx[i], cy[i]]) + rng.normal(0, 1.0, (5, 2))
        lm[i] = p.astype(np.float32)
    return frame_idx, lm
