"""Parity of the CUDA alignment kernel (through the C ABI) against the installed
OpenCV (golden vectors + live) and the C oracle.  Gate (BASELINE.json): aligned
pixels within 1 LSB of the uint8 cv2 output; the fixed-point formulation is expected
to be bit-exact, which is what is asserted."""
import hashlib

import numpy as np
import pytest

import frp_b200
from frp_b200 import align
import frp_oracle_c as oc
import frp_oracle_np as o

pytestmark = pytest.mark.gpu


def _oracle_batch(frames, fi, M):
    al = np.stack([oc.warp_affine(frames[fi[i]], M[i]) for i in range(len(fi))])
    bl = np.stack([oc.blob(a).astype(np.float16) for a in al])
    return al, bl


def test_golden_cv2_crops(golden_align):
    g = golden_align
    blob, al = align.norm_crop_batch(g["frames"], g["frame_idx"], lmk=g["lmk"], return_aligned=True)
    assert blob.shape == (len(g["lmk"]), 3, 112, 112) and blob.dtype == np.float16
    M = align.estimate_norm(g["lmk"])
    assert np.abs(M.reshape(-1, 6) - g["M"].reshape(-1, 6)).max() < 1e-9
    assert np.array_equal(al[:4], g["aligned_first4"])
    assert np.array_equal(blob[:4], g["blob_first4"])
    for i in range(len(al)):
        assert hashlib.sha256(al[i].tobytes()).hexdigest() == str(g["sha_aligned"][i]), i
        assert hashlib.sha256(blob[i].tobytes()).hexdigest() == str(g["sha_blob"][i]), i
    # matrices supplied by the caller give the same pixels
    blob2, al2 = align.norm_crop_batch(g["frames"], g["frame_idx"], M=g["M"], return_aligned=True)
    assert np.array_equal(al2, al) and np.array_equal(blob2, blob)


def test_1080p_crops_vs_cv2_live():
    cv2 = pytest.importorskip("cv2")
    frames = o.synth_frames(2, 1080, 1920, seed=21)
    fi, lm = o.synth_detections(2, 96, seed=22)
    blob, al = align.norm_crop_batch(frames, fi, lmk=lm, return_aligned=True)
    worst = 0
    for i in range(len(fi)):
        M = o.estimate_norm(lm[i])
        want = cv2.warpAffine(frames[fi[i]], M, (112, 112), borderValue=0.0)
        worst = max(worst, int(np.abs(want.astype(int) - al[i].astype(int)).max()))
        wb = cv2.dnn.blobFromImages([want], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
        if np.array_equal(want, al[i]):
            assert np.array_equal(wb.astype(np.float16), blob[i])
    assert worst <= 1          # the gate
    assert worst == 0          # what the fixed-point scheme delivers


def test_float32_estimate_reproduces_insightface_precision():
    """insightface hands skimage float32 landmarks, so its matrices come from a float32 Umeyama
    (oracle: umeyama_f32, LAPACK sgesdd).  frp_estimate_norm_f32 evaluates the closed form on the
    same float32 intermediates: matrices agree to ~1e-6 relative, and crops warped with them stay
    within north_star's 1 LSB of cv2.warpAffine on the oracle's float32 matrices."""
    cv2 = pytest.importorskip("cv2")
    frames = o.synth_frames(2, 1080, 1920, seed=61)
    fi, lm = o.synth_detections(2, 64, seed=62)
    M32 = align.estimate_norm(lm, precision="float32")
    M64 = align.estimate_norm(lm)
    want = np.stack([o.estimate_norm_f32(l) for l in lm])
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert (np.abs(M32 - want) / scale).max() < 2e-6
    assert (np.abs(M32 - M64) / scale).max() < 2e-6 and not np.array_equal(M32, M64)
    blob, al = align.norm_crop_batch(frames, fi, M=M32, return_aligned=True)
    worst, differing = 0, 0
    for i in range(len(fi)):
        ref = cv2.warpAffine(frames[fi[i]], want[i], (112, 112), borderValue=0.0)
        d = np.abs(ref.astype(int) - al[i].astype(int))
        worst = max(worst, int(d.max()))
        differing += int((d > 0).sum())
    assert worst <= 1                                   # the gate
    assert differing < 0.002 * al.size                  # and almost every pixel is identical


def test_large_faces_take_the_direct_path_and_odd_strides():
    # sigma up to 8 source px per template px exceeds the per-warp staging buffer
    H, W = 700, 1001                      # 3*W = 3003 is not a multiple of 16 -> repitched on upload
    frames = o.synth_frames(1, H, W, seed=31)
    fi, lm = o.synth_detections(1, 24, H, W, seed=32, sigma_range=(3.0, 8.0), oob_frac=0.3)
    M = np.stack([o.estimate_norm(l) for l in lm])
    blob, al = align.norm_crop_batch(frames, fi, lmk=lm, return_aligned=True)
    wal, wbl = _oracle_batch(frames, fi, M)
    assert np.array_equal(al, wal) and np.array_equal(blob, wbl)


def test_fully_outside_and_degenerate():
    frames = o.synth_frames(1, 128, 160, seed=41)
    lm = np.zeros((3, 5, 2), np.float32)
    lm[0] = o.ARCFACE_DST + 5000.0        # far outside the frame: all zeros -> -1.0
    lm[1] = o.ARCFACE_DST                 # identity transform
    lm[2] = 7.0                           # degenerate (all landmarks equal)
    blob, al = align.norm_crop_batch(frames, [0, 0, 0], lmk=lm, return_aligned=True)
    assert np.all(al[0] == 0) and np.all(blob[0] == np.float16(-1.0))
    want = oc.warp_affine(frames[0], o.estimate_norm(lm[1]))
    assert np.array_equal(al[1], want)
    assert np.all(al[2] == 0)
    assert align.norm_crop_batch(frames, [], lmk=np.zeros((0, 5, 2), np.float32))[0].shape == (0, 3, 112, 112)


def test_insightface_single_image_signatures():
    frames = o.synth_frames(1, 360, 480, seed=71)
    fi, lm = o.synth_detections(1, 3, 360, 480, seed=72, sigma_range=(0.6, 1.5))
    for l in lm:
        M = o.estimate_norm(l)
        want = oc.warp_affine(frames[0], M)
        assert np.array_equal(align.norm_crop(frames[0], l), want)
        got, M2 = align.norm_crop2(frames[0], l)
        assert np.array_equal(got, want) and np.abs(M2 - M).max() < 1e-9
    with pytest.raises(ValueError):
        align.norm_crop(frames[0], lm[0], image_size=128)


def test_caller_matrices_all_sign_patterns():
    """Caller-supplied 2x3 maps: every rotation quadrant, reflections, shear, scales 0.2-4 and
    translations that push part (or all, up to the int16 saturation of cv::remap) of the
    footprint outside the frame.  The per-crop staging plan picks the box corner from the
    signs of the inverse map; all four sign patterns of each row are exercised."""
    H, W = 480, 640
    frames = o.synth_frames(2, H, W, seed=61)
    rng = np.random.default_rng(62)
    Ms, fi = [], []
    for k in range(160):
        th = (k % 16) * (2 * np.pi / 16) + rng.uniform(-0.15, 0.15)
        sc = float(rng.choice([0.2, 0.5, 0.8, 1.0, 1.3, 2.0, 3.1, 4.0]))     # output px per source px
        A = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) * sc
        if k % 5 == 1:
            A = A @ np.array([[1.0, 0.0], [0.0, -1.0]])      # reflection
        if k % 5 == 2:
            A = A @ np.array([[1.0, 0.35], [0.0, 1.0]])      # shear
        if k % 5 == 3:
            A = A @ np.diag([1.0, 0.6])                      # anisotropic
        c = np.array([rng.uniform(-40, W + 40), rng.uniform(-40, H + 40)])    # source point mapped to the crop centre
        if k % 40 == 7:
            c = np.array([31000.0, -31000.0])                 # beyond the int16 range of the xy map
        t = np.array([56.0, 56.0]) - A @ c
        Ms.append(np.hstack([A, t[:, None]]))
        fi.append(k % 2)
    M = np.stack(Ms)
    blob, al = align.norm_crop_batch(frames, fi, M=M, return_aligned=True)
    wal, wbl = _oracle_batch(frames, fi, M)
    bad = [i for i in range(len(fi)) if not np.array_equal(al[i], wal[i])]
    assert not bad, bad
    assert np.array_equal(blob, wbl)
    assert sum(int(a.any()) for a in al) > 100      # most crops do show the frame


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_frames_strides_and_maps(seed):
    """Random frame sizes (pitches that are and are not 16-byte multiples, frames smaller than a
    TMA box, single-row-ish frames), random similarity / affine maps and landmark-driven crops."""
    rng = np.random.default_rng(900 + seed)
    for _ in range(5):
        H, W = int(rng.integers(8, 420)), int(rng.integers(8, 520))
        F = int(rng.integers(1, 4))
        frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
        n = 10
        fi = rng.integers(0, F, n).astype(np.int32)
        Ms = []
        for i in range(n):
            th = rng.uniform(0, 2 * np.pi)
            sc = float(np.exp(rng.uniform(np.log(0.15), np.log(6.0))))
            A = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) * sc
            if i % 3 == 0:
                A = A @ np.array([[1.0, rng.uniform(-0.5, 0.5)], [rng.uniform(-0.3, 0.3), rng.choice([-1.0, 1.0])]])
            c = np.array([rng.uniform(-0.2 * W, 1.2 * W), rng.uniform(-0.2 * H, 1.2 * H)])
            Ms.append(np.hstack([A, (np.array([56.0, 56.0]) - A @ c)[:, None]]))
        M = np.stack(Ms)
        blob, al = align.norm_crop_batch(frames, fi, M=M, return_aligned=True)
        wal, wbl = _oracle_batch(frames, fi, M)
        assert np.array_equal(al, wal), (seed, H, W)
        assert np.array_equal(blob, wbl), (seed, H, W)
        # landmark-driven crops on the same frames
        lm = np.empty((6, 5, 2), np.float32)
        for i in range(6):
            th, sg = rng.uniform(-0.6, 0.6), rng.uniform(0.2, 2.5)
            R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            ctr = np.array([rng.uniform(-0.1 * W, 1.1 * W), rng.uniform(-0.1 * H, 1.1 * H)])
            lm[i] = ((o.ARCFACE_DST.astype(np.float64) - 56.0) @ R.T) * sg + ctr + rng.normal(0, 0.7, (5, 2))
        fi2 = rng.integers(0, F, len(lm)).astype(np.int32)
        blob2, al2 = align.norm_crop_batch(frames, fi2, lmk=lm, return_aligned=True)
        M2 = np.stack([o.estimate_norm(l) for l in lm])
        wal2, wbl2 = _oracle_batch(frames, fi2, M2)
        assert np.array_equal(al2, wal2) and np.array_equal(blob2, wbl2), (seed, H, W)


def test_device_path_many_crops():
    """BASELINE.json configs[3] at full size: 4 096 detections from 16 1080p frames."""
    import torch
    F, H, W, per = 16, 1080, 1920, 256
    frames = np.random.default_rng(51).integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    frames[:2] = o.synth_frames(2, H, W, seed=51)       # two smooth frames, the rest white noise
    fi, lm = o.synth_detections(F, per, seed=52)
    n = len(fi)
    d_fr = torch.from_numpy(frames).cuda()
    d_fi = torch.from_numpy(fi).cuda()
    d_lm = torch.from_numpy(lm).cuda()
    out = torch.empty((n, 3, 112, 112), dtype=torch.float16, device="cuda")
    out_al = torch.empty((n, 112, 112, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    before = frp_b200.kernel_launches()
    align.norm_crop_batch_device(d_fr, F, H, W, 3 * W, d_fi, d_lm, n, out, out_al, stream=st)
    torch.cuda.synchronize()
    assert frp_b200.kernel_launches() == before + 3      # align_plan_kernel + align_kernel + align_direct_kernel
    al = out_al.cpu().numpy()
    blob = out.cpu().numpy()
    sel = np.random.default_rng(0).choice(n, 64, replace=False)
    M = np.stack([o.estimate_norm(lm[i]) for i in sel])
    wal, wbl = _oracle_batch(frames, fi[sel], M)
    assert np.array_equal(al[sel], wal) and np.array_equal(blob[sel], wbl)
    # checksum-of-checksums style property on the rest: blob is a pure function of aligned
    lut = ((np.arange(256, dtype=np.float32) - np.float32(127.5)) * np.float32(1 / 127.5)).astype(np.float16)
    assert np.array_equal(blob, lut[al][..., ::-1].transpose(0, 3, 1, 2))
