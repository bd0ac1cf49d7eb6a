"""Generate the committed golden vectors under tests/golden/.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs
/root/reference for the L2 vectors and the installed OpenCV for the warp):

    python oracle/gen_golden.py

* l2_c1.npz      — BASELINE.json configs[0]: 1 000 x 128 synthetic encodings,
                   tolerance 0.6, outputs of the UNMODIFIED reference
                   face_service.compare_faces / batch_compare_faces /
                   find_k_nearest (face_service.py:395,448,590) via ref_loader.
* cos_small.npz  — 512-d cosine 1:N outputs of the numpy restatement (parity
                   unpinned: insightface is not installed anywhere).
* align_small.npz— 5-landmark alignment; aligned pixels from the installed
                   cv2.warpAffine, blob from cv2.dnn.blobFromImages.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import frp_oracle_np as o  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def unit(rng, dim):
    u = rng.standard_normal(dim)
    return u / np.linalg.norm(u)


def gen_l2():
    fs, ENC, _ = ref_loader.load_reference()
    rng = np.random.default_rng(2024)
    G = o.synth_l2_gallery(1000, 128, seed=1234)
    Q, planted = o.synth_l2_queries(G, 32, seed=4321)
    # threshold-stress rows: rows at exact distance 0.6 +- delta from queries 0..15
    deltas = [0.0, 1e-6, -1e-6, 1e-4, -1e-4, 1e-3, -1e-3, 1e-2, -1e-2]
    extra = []
    for b in range(16):
        for dl in deltas:
            extra.append(Q[b] + (0.6 + dl) * unit(rng, 128))
    # exact duplicates (tie order = insertion order): 4 copies of row 17 and of a planted row
    for _ in range(4):
        extra.append(G[17].copy())
    G = np.concatenate([G, np.array(extra)], axis=0)
    names = ["id%07d" % i for i in range(len(G))]
    ENC.clear()
    for n_, g in zip(names, G):
        ENC[n_] = g.tolist()
    # query 31 = exactly row 17 (distance 0 to five rows)
    Q[31] = G[17]
    name_to_row = {n_: i for i, n_ in enumerate(names)}
    order = np.zeros((len(Q), len(G)), np.int32)
    dist = np.zeros((len(Q), len(G)), np.float64)
    match = np.zeros((len(Q), len(G)), np.bool_)
    score = np.zeros((len(Q), len(G)), np.float64)
    conf = np.zeros((len(Q), len(G)), np.int8)
    lvl = {"high": 0, "medium": 1, "low": 2}
    fs.tolerance = 0.6
    for b in range(len(Q)):
        res = fs.compare_faces(Q[b], None, True)
        assert len(res) == len(G)
        for j, r in enumerate(res):
            order[b, j] = name_to_row[r["target"]]
            dist[b, j] = r["distance"]
            match[b, j] = r["match"]
            score[b, j] = r["confidence_score"]
            conf[b, j] = lvl[r["confidence"]]
    knn = np.zeros((len(Q), 5), np.int32)
    knn_d = np.zeros((len(Q), 5), np.float64)
    for b in range(len(Q)):
        res = fs.find_k_nearest(Q[b], 5)
        knn[b] = [name_to_row[r["target"]] for r in res]
        knn_d[b] = [r["distance"] for r in res]
    bat = fs.batch_compare_faces(list(Q), None)
    bat_cnt = np.array([len(r) for r in bat], np.int32)
    bat_rows = np.full((len(Q), max(1, bat_cnt.max())), -1, np.int32)
    for b, r in enumerate(bat):
        bat_rows[b, :len(r)] = [name_to_row[x["target"]] for x in r]
    # subset / unknown-name behaviour (face_service.py:403)
    sub_names = [names[5], "nope", names[3], names[17], names[1000 + 3]]
    sub = fs.compare_faces(Q[0], sub_names, True)
    sub_rows = np.array([name_to_row[r["target"]] for r in sub], np.int32)
    sub_dist = np.array([r["distance"] for r in sub])
    ENC.clear()
    np.savez_compressed(os.path.join(OUT, "l2_c1.npz"), G=G, Q=Q, order=order, dist=dist,
                        match=match, score=score, conf=conf, knn=knn, knn_d=knn_d,
                        bat_cnt=bat_cnt, bat_rows=bat_rows,
                        sub_query_rows=np.array([5, -1, 3, 17, 1003], np.int32),
                        sub_rows=sub_rows, sub_dist=sub_dist,
                        numpy_version=np.__version__)
    print("l2_c1:", G.shape, Q.shape, "matches/query", match.sum(1)[:8])


def gen_cos():
    rng = np.random.default_rng(77)
    N, D, B, k, tau = 768, 512, 24, 10, 0.40
    G = o.synth_cos_gallery(N, D, seed=1234).astype(np.float32)
    Q, planted = o.synth_cos_queries(G, B, seed=4321)
    Q = Q.astype(np.float32)
    extra = []
    for b in range(8):
        q = o.normed(Q[b].astype(np.float64))
        for dl in (0.0, 1e-6, -1e-6, 1e-4, -1e-4, 1e-3, -1e-3):
            c = tau + dl
            u = unit(rng, D)
            u = u - (u @ q) * q
            u /= np.linalg.norm(u)
            extra.append((c * q + np.sqrt(1 - c * c) * u).astype(np.float32))
    for _ in range(3):
        extra.append(G[11].copy())
    G = np.concatenate([G, np.array(extra, dtype=np.float32)], 0)
    # un-normalised rows exercise the store's normalisation
    G[5] *= 3.0
    G[6] *= 0.25
    Q[3] *= 7.0
    Q[23] = G[11]
    idx, sc, cnt = o.cos_match_topk(G, Q, tau, k, True)
    idx_nt, sc_nt, cnt_nt = o.cos_match_topk(G, Q, tau, k, False)
    S = o.cosine_scores(G, Q)
    np.savez_compressed(os.path.join(OUT, "cos_small.npz"), G=G, Q=Q, tau=tau, k=k,
                        idx=idx, score=sc, count=cnt, idx_nt=idx_nt, score_nt=sc_nt,
                        all_scores=S.astype(np.float64))
    print("cos_small:", G.shape, Q.shape, "count", cnt)


def gen_align():
    import cv2
    H, W = 270, 480
    frames = o.synth_frames(2, H, W, seed=5)
    rng = np.random.default_rng(6)
    n = 48
    frame_idx = rng.integers(0, 2, n).astype(np.int32)
    sigma = rng.uniform(0.3, 2.2, n)
    theta = rng.uniform(-0.6, 0.6, n)
    cx = rng.uniform(-10, W + 10, n)
    cy = rng.uniform(-10, H + 10, n)
    base = o.ARCFACE_DST.astype(np.float64) - 56.0
    lm = np.empty((n, 5, 2), np.float32)
    for i in range(n):
        c, s = np.cos(theta[i]), np.sin(theta[i])
        R = np.array([[c, -s], [s, c]])
        lm[i] = ((base @ R.T) * sigma[i] + [cx[i], cy[i]] + rng.normal(0, 1, (5, 2))).astype(np.float32)
    M = np.stack([o.estimate_norm(lm[i]) for i in range(n)])
    al = np.empty((n, 112, 112, 3), np.uint8)
    bl = np.empty((n, 3, 112, 112), np.float16)
    for i in range(n):
        al[i] = cv2.warpAffine(frames[frame_idx[i]], M[i], (112, 112), borderValue=0.0)
        blob = cv2.dnn.blobFromImages([al[i]], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)
        bl[i] = blob[0].astype(np.float16)
        assert np.array_equal(al[i], o.warp_affine_fixed(frames[frame_idx[i]], M[i]))
        assert np.array_equal(bl[i], o.blob_from_aligned(al[i]))
    sha_al = np.array([hashlib.sha256(a.tobytes()).hexdigest() for a in al])
    sha_bl = np.array([hashlib.sha256(b.tobytes()).hexdigest() for b in bl])
    # the 256-entry normalisation table as OpenCV computes it
    v = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    lut = cv2.dnn.blobFromImages([v], 1.0 / 127.5, (16, 16), (127.5, 127.5, 127.5))[0, 0].reshape(256)
    np.savez_compressed(os.path.join(OUT, "align_small.npz"), frames=frames, frame_idx=frame_idx,
                        lmk=lm, M=M, sha_aligned=sha_al, sha_blob=sha_bl,
                        aligned_first4=al[:4], blob_first4=bl[:4],
                        lut_f32=lut.astype(np.float32), lut_f16=lut.astype(np.float16),
                        cv2_version=cv2.__version__)
    print("align_small:", frames.shape, n, "crops; cv2", cv2.__version__)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_l2()
    gen_cos()
    gen_align()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
