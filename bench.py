#!/usr/bin/env python
"""Benchmark of the recognition hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (N=1): 512-d cosine 1:N, batch 256 queries per step against a 10M x 512
gallery resident in HBM, threshold 0.40 then top-10.  For N>1 the same 10M rows are
row-sharded across the ranks (strong scaling), per-shard top-k merged after one NCCL
all-gather.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

if "reference" in sys.argv:
    # the CPU arm uses every host core at every N: torchrun exports OMP_NUM_THREADS=1, which
    # OpenBLAS would otherwise pick up when numpy loads
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "face-recognition-platform_b200"))

METRIC = "1:N match queries/s @10M x512 gallery"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--tau", type=float, default=0.40)
    ap.add_argument("--no-extras", action="store_true", help="skip the align / batch-1 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default=os.environ.get("FRP_EXCHANGE", "nccl"), choices=["nccl", "p2p"],
                    help="multi-GPU merge: NCCL all-gather + merge kernel, or the fused peer-memory kernel")
    ap.add_argument("--also-batches", default="", help="comma list of extra batch sizes timed on the same gallery")
    ap.add_argument("--pipeline", action="store_true", help="force the two-slot pipeline (default when more than one rank runs)")
    ap.add_argument("--no-pipeline", action="store_true", help="plain launches on one stream at every rank count")
    ap.add_argument("--graph", action="store_true", help="replay one CUDA graph per step (opt-in)")
    ap.add_argument("--c3", choices=["auto", "on", "off"], default="auto",
                    help="100 M-row leg (12.5 M rows per rank): auto = when 8 ranks run")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            burst = float(d["bf16_tflops"])
            return {"hbm": float(d["hbm_gbs"]), "tensor_burst": burst,
                    "tensor": float(d.get("bf16_tflops_sustained", burst)), "src": "measured"}
        except Exception:      # unreadable or incomplete file: documented fallback below
            pass
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


def config(args, n_gpus):
    return {"workload": "cosine 1:N, batch %d x gallery %d x 512 fp16 (fp32 master re-rank), tau %.2f, top-%d"
                        % (args.batch, args.rows, args.tau, args.k),
            "batch": args.batch, "gallery_rows": args.rows, "dim": 512, "top_k": args.k, "tau": args.tau,
            "parallelism": "rowshard%d" % n_gpus,
            "l2_between_steps": "inputs larger than L2 (gallery %.1f GB per pass)" % (args.rows * 1024 / 1e9 / n_gpus)}


# --------------------------------------------------------------------------
# Synthetic workloads.  bench.py generates its own inputs; `oracle/` is imported
# only by the cpu_baseline legs below (and by `--impl reference`).
# --------------------------------------------------------------------------
_ARCFACE_DST = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                         [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float32)


def synth_detections(nframes, per_frame, H=1080, W=1920, seed=7, sigma_range=(0.4, 3.0), oob_frac=0.02):
    """Landmarks = ((arcface_dst - 56) @ R(theta)^T) * sigma + centre + N(0, 1 px): magnification
    sigma source px per template px, roll theta in [-0.5, 0.5] rad, a few detections at the frame
    border."""
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
    base = _ARCFACE_DST.astype(np.float64) - 56.0
    lm = np.empty((n, 5, 2), np.float32)
    for i in range(n):
        c, s = np.cos(theta[i]), np.sin(theta[i])
        R = np.array([[c, -s], [s, c]])
        lm[i] = ((base @ R.T) * sigma[i] + np.array([cx[i], cy[i]]) + rng.normal(0, 1.0, (5, 2))).astype(np.float32)
    return frame_idx, lm


def synth_l2_gallery(n=1000, dim=128, seed=1234):
    """128-d encodings ~ N(0, 0.05): impostor distance ~0.8, no impostor match at tolerance 0.6."""
    return np.random.default_rng(seed).normal(0, 0.05, (n, dim))


def synth_l2_queries(G, nq, seed=4321, planted_frac=0.5):
    rng = np.random.default_rng(seed)
    n, dim = G.shape
    Q = rng.normal(0, 0.05, (nq, dim))
    planted = rng.random(nq) < planted_frac
    j = rng.integers(0, n, nq)
    Q[planted] = G[j[planted]] + rng.normal(0, 0.02, (int(planted.sum()), dim))
    return Q, np.where(planted, j, -1)


# --------------------------------------------------------------------------
# CPU baseline: the reference's numpy path (oracle port) on the host cores
# --------------------------------------------------------------------------
def _blas_threads():
    """Pin the BLAS pool to every host core (explicitly: the environment may say 1) and return
    the thread count actually in use."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        return max([t.get("num_threads", 1) for t in threadpool_info()] + [1])
    except Exception:
        return n


def _cpu_topk_pass(Q, G, rows_total, k, tau, chunk=65536):
    """numpy / OpenBLAS cosine 1:N with threshold + top-k over `rows_total` gallery rows: chunked
    sgemm + argpartition + merge.  G is a block of rows that is cycled when rows_total > len(G)
    (identical arithmetic per row; a block of >= 1 M x 512 fp32 = 2 GB never fits a CPU cache)."""
    B = Q.shape[0]
    best_s = np.full((B, k), -np.inf, np.float32)
    best_i = np.full((B, k), -1, np.int64)
    nblk = len(G)
    done = 0
    while done < rows_total:
        lo = done % nblk
        n = min(chunk, nblk - lo, rows_total - done)
        S = Q @ G[lo:lo + n].T
        kk = min(k, n)
        part = np.argpartition(-S, kk - 1, axis=1)[:, :kk]
        ps = np.take_along_axis(S, part, axis=1)
        cs = np.concatenate([best_s, ps], 1)
        ci = np.concatenate([best_i, part + done], 1)
        o_ = np.lexsort((ci, -cs), axis=1)[:, :k]
        best_s = np.take_along_axis(cs, o_, 1)
        best_i = np.take_along_axis(ci, o_, 1)
        done += n
    return best_i, best_s, (best_s >= tau).sum(1)


def _cpu_workload(args, block_rows):
    rng = np.random.default_rng(0)
    G = rng.standard_normal((block_rows, 512), dtype=np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    Q = rng.standard_normal((args.batch, 512), dtype=np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    return G, Q


def cpu_match_sample(args, rows_sample=1_000_000, reps=3):
    """cpu_baseline leg of the GPU arm (bounded: ~10-30 s): `batch` queries vs a 1 M-row sample
    of the gallery, median of `reps`, scaled to the full row count.  `--impl reference` runs the
    full row count instead."""
    cores = _blas_threads()
    rows_sample = min(rows_sample, args.rows)
    G, Q = _cpu_workload(args, rows_sample)
    _cpu_topk_pass(Q, G, min(rows_sample, 131072), args.k, args.tau)        # warm the BLAS pool
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        _cpu_topk_pass(Q, G, rows_sample, args.k, args.tau)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    qps = args.batch / (t * (args.rows / rows_sample))
    return {"value": qps, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": "numpy sgemm + argpartition (the oracle's arithmetic; the reference has no 512-d path), %d "
                      "queries x %d of %d rows (fp32), median of %d, scaled x%g"
                      % (args.batch, rows_sample, args.rows, reps, args.rows / rows_sample),
            "sample_seconds": t}


def run_reference(args, budget_s=150.0):
    """`--impl reference`: the CPU implementation of the path on ALL host threads, on the GPU
    arm's exact config: every step scores `batch` queries against all `rows` gallery rows.  The
    number of timed steps is capped by a wall-clock budget (>= 3) and reported as `steps`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _blas_threads()
    block = min(args.rows, 1_000_000)
    G, Q = _cpu_workload(args, block)
    # warm-up: W passes over one block (BLAS pool and page faults; a full-size pass would add
    # nothing), plus one timed block pass to size the step count against the budget
    for _ in range(args.warmup):
        _cpu_topk_pass(Q, G, block, args.k, args.tau)
    t0 = time.perf_counter()
    _cpu_topk_pass(Q, G, block, args.k, args.tau)
    est = (time.perf_counter() - t0) * args.rows / block
    steps = int(max(min(3, args.steps), min(args.steps, budget_s // max(est, 1e-3))))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        _cpu_topk_pass(Q, G, args.rows, args.k, args.tau)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    qps = args.batch / t
    sample = ("numpy sgemm + argpartition + merge, %d queries x ALL %d rows per step (a %d-row fp32 block cycled "
              "%gx), median of %d steps (of %d requested; capped by a %.0f s budget); warm-up = %d block passes"
              % (args.batch, args.rows, block, args.rows / block, steps, args.steps, budget_s, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config(args, args.gpus),
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": int(cores), "kind": "port", "sample": sample,
                             "step_seconds": [round(x, 3) for x in ts]},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _lines(self):
        try:
            return [l for l in open(self.f.name) if l.strip()]
        except Exception:
            return []

    def wait_first(self, timeout=4.0):
        """nvidia-smi needs ~1 s to start; block until it has produced a sample."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout and not self._lines():
            time.sleep(0.05)

    def mark(self):
        self.skip = len(self._lines())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in self._lines()[getattr(self, "skip", 0):]]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        # under load = samples in the upper half of observed power/clock activity
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_gallery(frp, torch, rows_local, seed, device, keep_rows=4096, dim=512, metric=None, scale=1.0):
    """Random gallery generated on the device in 1 M-row chunks.  Returns (gallery, keep, keep_idx):
    `keep` = `keep_rows` rows drawn from the WHOLE local row range (the planted queries are
    mixed from them, so matches land in every tile of every CTA, not only in the first ones)."""
    gal = frp.Gallery(dim, frp.METRIC_COS if metric is None else metric, device=device.index or 0, capacity=rows_local)
    gen = torch.Generator(device=device).manual_seed(seed)
    chunk = 1_000_000
    nchunks = max(1, (rows_local + chunk - 1) // chunk)
    per = max(1, -(-keep_rows // nchunks))          # ceil: the concatenation is cut to exactly keep_rows
    keep, keep_idx = [], []
    st = torch.cuda.current_stream().cuda_stream
    for lo in range(0, rows_local, chunk):
        n = min(chunk, rows_local - lo)
        x = torch.randn((n, dim), generator=gen, device=device, dtype=torch.float32)
        if scale != 1.0:
            x *= scale
        gal.append_device(x, stream=st)
        sel = torch.randint(0, n, (min(per, n),), generator=gen, device=device)
        keep.append(x[sel].clone())
        keep_idx.append(sel + lo)
        del x
    torch.cuda.synchronize()
    return gal, torch.cat(keep)[:keep_rows].contiguous(), torch.cat(keep_idx)[:keep_rows]


def make_queries(torch, nsteps, B, keep, seed, device, planted_frac=0.5):
    """[nsteps, B, 512] fp32: impostors + alpha-mixed copies of gallery rows."""
    gen = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn((nsteps, B, 512), generator=gen, device=device)
    q = q / q.norm(dim=-1, keepdim=True)
    if keep is not None:
        kn = keep / keep.norm(dim=-1, keepdim=True)
        sel = torch.randint(0, kn.shape[0], (nsteps, B), generator=gen, device=device)
        alpha = torch.rand((nsteps, B, 1), generator=gen, device=device) * 0.5 + 0.45
        planted = torch.rand((nsteps, B, 1), generator=gen, device=device) < planted_frac
        mix = alpha * kn[sel] + (1 - alpha * alpha).sqrt() * q
        q = torch.where(planted, mix / mix.norm(dim=-1, keepdim=True), q)
    return q.contiguous()


def run_ours(args):
    import torch
    import frp_b200 as frp
    from frp_b200 import _capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if frp.device_count() < 1:
        raise SystemExit("bench.py needs a B200 (no CPU fallback): " + _capi.last_error())
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: while the communicator comes up (NCCL prints its
        # version banner there when NCCL_DEBUG is set, as it is in this image) fd 1 points at stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    pk = peaks()
    B, k, tau, K, W = args.batch, args.k, args.tau, args.steps, args.warmup

    lo, hi = frp.shard_bounds(args.rows, world, rank)
    # planted queries are mixed from rows of EVERY rank's shard, spread over its whole row range
    per_rank_keep = 4096 // world
    gal, keep, _ = make_gallery(frp, torch, hi - lo, 1234 + rank, device, keep_rows=per_rank_keep)
    if world > 1:
        allk = torch.empty((world * per_rank_keep, 512), device=device)
        dist.all_gather_into_tensor(allk, keep.contiguous())
        keep = allk
    matcher = frp.ShardedMatcher(gal, args.rows, rank, world, exchange=args.exchange)
    exchange_check = None
    if args.exchange == "p2p" and world > 1:
        # the fused peer-memory exchange must give exactly what all-gather + merge gives
        ref_m = frp.ShardedMatcher(gal, args.rows, rank, world, exchange="nccl")
        qchk = make_queries(torch, 1, B, keep, 4242, device)[0]
        a_ = [x.clone() for x in matcher.match(qchk, tau, k)]
        b_ = [x.clone() for x in ref_m.match(qchk, tau, k)]
        torch.cuda.synchronize()
        exchange_check = bool(torch.equal(a_[0], b_[0]) and torch.equal(a_[2], b_[2]) and
                              torch.equal(torch.nan_to_num(a_[1]), torch.nan_to_num(b_[1])) and
                              int(a_[3].item()) == int(b_[3].item()))
    nq = K + W
    queries = make_queries(torch, nq, B, keep, 99, device)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        return matcher.match(queries[i], tau, k)

    # stepping: more than one rank -> two-slot pipeline over two streams (all-gather + merge of step
    # i behind the prefilter of step i+1: +2 % at 8 ranks, neutral at 2); one rank -> plain
    # launches (nothing to overlap; the extra stream only adds event traffic).  --graph replays
    # one CUDA graph per step instead (measured equal within noise at 2 ranks: sustained steps are
    # power-capped, not launch-gap-bound).
    pipelined = (args.pipeline or (world > 1 and not args.no_pipeline)) and args.exchange == "nccl" and not args.graph
    graph = None
    launches_per_graph_step = 0
    if args.graph and args.exchange == "nccl":
        try:
            q_static = queries[0].clone()
            l0 = frp.kernel_launches()
            graph, g_out = matcher.capture(q_static, tau, k)
            # capture() issues the step three times (two warm-ups + the captured one); a replay
            # launches the same kernels without passing through the library's launch counter
            launches_per_graph_step = (frp.kernel_launches() - l0) // 3
        except Exception as e:      # pragma: no cover - a capture failure must not lose the run
            sys.stderr.write("CUDA graph capture failed (%r): plain launches\n" % (e,))
            graph = None

    def run_steps(a, b):
        """Steps a..b-1.  Pipelined (default): the all-gather + merge of step i run on the
        exchange stream behind the prefilter of step i+1; every step's merged result is
        produced (and waited for on the timing stream) inside the timed region."""
        if graph is not None:
            for i in range(a, b):
                q_static.copy_(queries[i], non_blocking=True)
                graph.replay()
            return g_out
        if not pipelined:
            out_ = None
            for i in range(a, b):
                out_ = step(i)
            return out_
        prev, out_ = None, None
        for i in range(a, b):
            t = matcher.submit(queries[i], tau, k)
            if prev is not None:
                out_ = matcher.result(prev)
            prev = t
        return matcher.result(prev)

    # ---- device-resident throughput -------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    out = run_steps(0, W)
    if sampler:
        sampler.wait_first()
    barrier()
    if sampler:
        sampler.mark()
    launches0 = frp.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run_steps(W, W + K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = frp.kernel_launches() - launches0
    if graph is not None:
        launches = K * launches_per_graph_step
    fallback_steps = int(out[3].item() != 0)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # the timed region can be shorter than nvidia-smi's sampling period: keep the identical
    # load running (untimed) for ~1.2 s so the clock record reflects the steps under load
    reps = max(1, int(1200.0 / max(ms, 1e-3)))
    for _ in range(min(reps, 2000)):
        run_steps(W, W + K)
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "timed region + %d untimed repeats of the same %d steps" % (min(reps, 2000), K)
    path = frp.last_match_path()
    value = B * K / (ms / 1e3)

    # ---- per-kernel timing for the roofline (separate pass, events inside the library) ----
    _capi.lib.frp_profile_enable(1)
    pre, mid, post = [], [], []
    for i in range(W, W + min(K, 10)):
        step(i)
        a, b_, c = C.c_float(), C.c_float(), C.c_float()
        _capi.check(_capi.lib.frp_profile_last(C.byref(a), C.byref(b_), C.byref(c)))
        pre.append(a.value); mid.append(b_.value); post.append(c.value)
    _capi.lib.frp_profile_enable(0)
    barrier()
    t_pref = float(np.mean(mid)) / 1e3
    rows_local = hi - lo
    flops = 2.0 * B * rows_local * 512
    gbytes = rows_local * 1024 / 1e9
    # the tcgen05 kernel is tensor-bound from ~200 queries per pass (ridge of the measured peaks:
    # 1 636.9 TFLOP/s / 6.56 TB/s = 250 flop/B = batch 250); below that it streams at HBM speed
    if path == "gemm_tcgen05" and B >= 192:
        ach = flops / t_pref / 1e12
        roof = {"bound": "tensor", "kernel": "match_gemm_tcgen05", "achieved": ach, "peak": pk["tensor_burst"],
                "unit": "TFLOP/s", "frac": ach / pk["tensor_burst"], "traffic": None,
                "peak_source": pk["src"] + " bf16 burst (kernel timed alone by events)",
                "frac_of_sustained_peak": ach / pk["tensor"], "sustained_peak": pk["tensor"],
                "hbm_achieved_gbs": gbytes / t_pref, "hbm_frac": gbytes / t_pref / pk["hbm"]}
    else:
        passes = 1 if path == "gemm_tcgen05" else (B + 3) // 4
        ach = gbytes * passes / t_pref
        roof = {"bound": "hbm", "kernel": "match_gemm_tcgen05" if path == "gemm_tcgen05" else "match_scan_f16",
                "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["src"] + " copy bandwidth",
                "note": "%d pass(es) over the gallery per step, path %s" % (passes, path)}
    # DRAM traffic per launch: ratio (dram bytes / algorithmic bytes) measured once with
    # `ncu --set full` on the same kernel (profiles/r1_traffic.json) x this launch's bytes
    try:
        tf = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if not os.path.isfile(tf):
            tf = os.path.join(ROOT, "profiles", "r1_traffic.json")
        tr = json.load(open(tf))[roof["kernel"]]
        roof["traffic"] = tr["dram_bytes_per_algorithmic_byte"] * rows_local * 1024 * \
            ((B + 3) // 4 if roof["kernel"] == "match_scan_f16" else 1)
        roof["traffic_source"] = tr["source"] + " (dram bytes / algorithmic byte of one `ncu --set full` capture x this launch's bytes; not an in-run counter)"
    except Exception:
        pass
    roof["algorithmic_bytes"] = rows_local * 1024
    roof["algorithmic_flops"] = flops
    roof["kernel_ms"] = {"prep": float(np.mean(pre)), "prefilter": float(np.mean(mid)), "rerank": float(np.mean(post))}
    roof["kernel_share_of_step"] = float(np.mean(mid)) / (ms / K)

    # ---- end to end through the public host API (pinned host queries, results read back) ----
    e2e = None
    if world == 1:
        hq = queries.cpu().pin_memory().numpy()
        for i in range(W):
            gal.match(hq[i], tau, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nfb = 0
        for i in range(W, W + K):
            r = gal.match(hq[i], tau, k)
            nfb += int((r[3] != 0).sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": B * K / dt, "unit": UNIT, "h2d_bytes_per_step": B * 512 * 4,
               "d2h_bytes_per_step": B * k * 12 + B * 8, "api": "frp_match (Gallery.match), host buffers",
               "uncertified_queries": nfb, "ms_per_step": dt / K * 1e3}
    else:
        # multi-rank: pinned host queries -> H2D on every rank, pipelined sharded match, merged
        # result D2H into pinned memory every step; the certificate flag is checked one step late
        hq = queries.cpu().pin_memory()
        pipe = frp.CertifiedPipeline(matcher, B, 512, k, device)
        levels = []

        def e2e_run(a, b):
            for i in range(a, b):
                r = pipe.push(hq[i], tau)
                if r is not None:
                    levels.append(r[3])
            levels.append(pipe.flush()[3])

        e2e_run(0, W)
        barrier()
        t0 = time.perf_counter()
        e2e_run(W, W + K)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": B * K / dt, "unit": UNIT, "h2d_bytes_per_step": B * 512 * 4,
               "d2h_bytes_per_step": B * k * 12 + B * 4 + 4,
               "api": "CertifiedPipeline over ShardedMatcher (pinned host queries in, pinned host results out, "
                      "certificate checked one step late)",
               "fallback_steps": int(sum(1 for x in levels if x)), "ms_per_step": dt / K * 1e3}

    extras = {}
    if world == 1 and not args.no_extras and rank == 0:
        extras = side_measurements(args, frp, torch, gal, queries, pk, device, keep)
    if world > 1 and not args.no_extras:
        # alignment shards by stream with no exchange: every rank aligns the 4 096 detections
        # of its own 16 frames (weak scaling); time = max over ranks
        extras["align_sharded"] = align_sharded(frp, torch, dist, device, rank, world)
    if not args.no_extras:
        try:
            extras["pipeline_c5"] = pipeline_c5(args, frp, torch, dist, matcher, keep, device, rank, world)
        except Exception as e:  # pragma: no cover
            extras["pipeline_c5"] = {"error": repr(e)}
    if (args.c3 == "on" or (args.c3 == "auto" and world == 8)) and not args.no_extras:
        try:
            extras["c3_100M"] = c3_100m(args, frp, torch, dist, device, rank, world, pk)
        except Exception as e:  # pragma: no cover
            extras["c3_100M"] = {"error": repr(e)}
    # other batch sizes against the same resident gallery (config C3: batch 1 and 1024)
    if args.also_batches:
        other = {}
        for Bx in [int(x) for x in args.also_batches.split(",") if x]:
            nsx = max(4, min(K, 20))
            qx = make_queries(torch, nsx + 2, Bx, keep, 7 + Bx, device)
            for i in range(2):
                matcher.match(qx[i], tau, k)
            barrier()
            ex0, ex1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ex0.record()
            for i in range(2, 2 + nsx):
                ox = matcher.match(qx[i], tau, k)
            ex1.record()
            barrier()
            msx = ex0.elapsed_time(ex1)
            if dist is not None:
                tx = torch.tensor([msx], device=device)
                dist.all_reduce(tx, op=dist.ReduceOp.MAX)
                msx = float(tx.item())
            _capi.lib.frp_profile_enable(1)
            matcher.match(qx[0], tau, k)
            pb = C.c_float()
            _capi.check(_capi.lib.frp_profile_last(None, C.byref(pb), None))
            _capi.lib.frp_profile_enable(0)
            barrier()
            other[str(Bx)] = {"queries_per_s": Bx * nsx / (msx / 1e3), "ms_per_step": msx / nsx,
                              "prefilter_kernel_ms": pb.value, "match_path": frp.last_match_path(),
                              "uncertified": int(ox[3].item() != 0),
                              "local_hbm_gbs": rows_local * 1024 / 1e9 / (pb.value / 1e3),
                              "local_tflops": 2.0 * Bx * rows_local * 512 / (pb.value / 1e3) / 1e12}
        extras["batches"] = other
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_match_sample(args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 (fp32 accumulate; fp64 re-rank of fp32 master)", "data": "synthetic",
                "config": config(args, world), "match_path": path, "gpu_launches": int(launches),
                "exchange": args.exchange if world > 1 else "none", "exchange_equals_nccl": exchange_check,
                "stepping": "cuda_graph" if graph is not None else ("two_slot_pipeline" if pipelined else "plain_launches"),
                "uncertified_last_step": fallback_steps, "clocks": clocks, "e2e": e2e, "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extras:
            line["extra"] = extras
        print(json.dumps(line))
    if dist is not None:
        graph = None            # a captured all-gather must be gone before the communicator is torn down
        g_out = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def _time_steps(torch, dist, device, fn, n, warm=2):
    """n calls of fn(i) bracketed by barriers, device events, max over ranks.  Returns ms per call."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / n


def _pipelined(matcher, qs, tau, k, flags=0):
    """Closure running qs[i % len] through the two-slot pipeline; call .drain() at the end."""
    state = {"prev": None}

    def fn(i):
        t = matcher.submit(qs[i % qs.shape[0]], tau, k, flags)
        if state["prev"] is not None:
            matcher.result(state["prev"])
        state["prev"] = t

    def drain():
        if state["prev"] is not None:
            matcher.result(state["prev"])
            state["prev"] = None

    fn.drain = drain
    return fn


def c3_100m(args, frp, torch, dist, device, rank, world, pk):
    """BASELINE configs[2]: 100 M x 512 fp16 rows row-sharded over the ranks (12.5 M per rank at 8
    GPUs, generated per shard on the device with seed + rank), batch 1 / 256 / 1024, NCCL top-k
    merge; device-resident queries, pipelined steps, max over ranks."""
    from frp_b200 import _capi
    total = 100_000_000
    lo, hi = frp.shard_bounds(total, world, rank)
    gal, keep, _ = make_gallery(frp, torch, hi - lo, 4321 + rank, device, keep_rows=max(8, 512 // world))
    if world > 1:
        allk = torch.empty((world * keep.shape[0], 512), device=device)
        dist.all_gather_into_tensor(allk, keep.contiguous())
        keep = allk
    m = frp.ShardedMatcher(gal, total, rank, world)
    out = {"workload": "100 M x 512 fp16 rows over %d ranks (%d rows per rank), tau %.2f, top-%d, queries resident, "
                       "two-slot pipeline" % (world, hi - lo, args.tau, args.k), "batches": {}}
    rows_local = hi - lo
    for Bx in (1, 256, 1024):
        qx = make_queries(torch, 8, Bx, keep, 11 + Bx, device)
        fn = _pipelined(m, qx, args.tau, args.k)
        nst = 12 if Bx < 1024 else 6
        ms = _time_steps(torch, dist, device, fn, nst)
        fn.drain()
        torch.cuda.synchronize()
        _capi.lib.frp_profile_enable(1)
        o_ = m.match(qx[0], args.tau, args.k)
        pb = C.c_float()
        _capi.check(_capi.lib.frp_profile_last(None, C.byref(pb), None))
        _capi.lib.frp_profile_enable(0)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        out["batches"][str(Bx)] = {
            "queries_per_s": Bx / (ms / 1e3), "ms_per_step": ms, "prefilter_kernel_ms": pb.value,
            "kernel_share_of_step": pb.value / ms, "match_path": frp.last_match_path(),
            "uncertified": int(o_[3].item() != 0),
            "local_hbm_gbs": rows_local * 1024 / 1e9 / (pb.value / 1e3),
            "local_hbm_frac": rows_local * 1024 / 1e9 / (pb.value / 1e3) / pk["hbm"],
            "local_tflops": 2.0 * Bx * rows_local * 512 / (pb.value / 1e3) / 1e12}
    gal.close()
    return out


def pipeline_c5(args, frp, torch, dist, matcher, keep, device, rank, world):
    """BASELINE configs[4]: 32 synthetic 1080p streams per tick, 8 faces each.  Streams are dealt
    to the ranks (shard_streams); per tick every rank copies ITS streams' frames from pinned host
    memory, aligns their detections, contributes its faces' embeddings (backbone out of scope:
    seeded synthetic 512-d rows stand in for its output) to one all-gather, and all ranks run
    the row-sharded match of the 256 faces against the 10 M watchlist; rank 0 reads the result."""
    S_, faces, H, W_ = 32, 8, 1080, 1920
    mine = frp.shard_streams(range(S_), world, rank)
    ns = len(mine)
    rng = np.random.default_rng(900 + rank)
    h_frames = torch.from_numpy(rng.integers(0, 256, (ns, H, W_, 3), dtype=np.uint8)).pin_memory()
    d_frames = torch.empty((ns, H, W_, 3), dtype=torch.uint8, device=device)
    fi, lm = synth_detections(ns, faces, H, W_, seed=11 + rank)
    d_fi, d_lm = torch.from_numpy(fi).to(device), torch.from_numpy(lm).to(device)
    n_mine = ns * faces
    n_all = S_ * faces
    blob = torch.empty((n_mine, 3, 112, 112), dtype=torch.float16, device=device)
    emb_all = make_queries(torch, 4, n_all, keep, 77, device)            # [4, 256, 512] stand-in embeddings
    my_lo = sum(len(frp.shard_streams(range(S_), world, r)) for r in range(rank)) * faces
    gathered = torch.empty((n_all, 512), device=device)
    h_idx = torch.empty((n_all, args.k), dtype=torch.int64).pin_memory()
    h_cnt = torch.empty((n_all,), dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    equal_split = all(len(frp.shard_streams(range(S_), world, r)) == ns for r in range(world))

    def tick(i):
        d_frames.copy_(h_frames, non_blocking=True)
        frp.align.norm_crop_batch_device(d_frames, ns, H, W_, 3 * W_, d_fi, d_lm, n_mine, blob,
                                         device=device.index, stream=st)
        e = emb_all[i % 4]
        if world > 1 and equal_split:
            dist.all_gather_into_tensor(gathered, e[my_lo:my_lo + n_mine].contiguous())
            q = gathered
        else:
            q = e
        oi, _, oc, _ = matcher.match(q, args.tau, args.k)
        if rank == 0:
            h_idx.copy_(oi, non_blocking=True)
            h_cnt.copy_(oc, non_blocking=True)

    ms = _time_steps(torch, dist, device, tick, 10, warm=2)
    return {"ticks_per_s": 1e3 / ms, "faces_per_s": n_all * 1e3 / ms, "frames_per_s": S_ * 1e3 / ms, "ms_per_tick": ms,
            "h2d_bytes_per_tick_per_rank": int(h_frames.numel()), "streams_per_rank": ns,
            "workload": "32 x 1080p streams over %d rank(s): frames H2D from pinned memory + align (%d crops per "
                        "rank) + embedding all-gather + row-sharded cosine match of %d faces vs %d rows, top-%d"
                        % (world, n_mine, n_all, args.rows, args.k)}


def align_sharded(frp, torch, dist, device, rank, world):
    F, H, W_, per = 16, 1080, 1920, 256
    rng = np.random.default_rng(50 + rank)
    frames = torch.from_numpy(rng.integers(0, 256, (F, H, W_, 3), dtype=np.uint8)).to(device)
    fi, lm = synth_detections(F, per, H, W_, seed=70 + rank)
    d_fi, d_lm = torch.from_numpy(fi).to(device), torch.from_numpy(lm).to(device)
    ncrop = F * per
    blob = torch.empty((ncrop, 3, 112, 112), dtype=torch.float16, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(8):
        flush.fill_(i)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frp.align.norm_crop_batch_device(frames, F, H, W_, 3 * W_, d_fi, d_lm, ncrop, blob,
                                         device=device.index, stream=st)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= 3:
            ts.append(float(t.item()) / 1e3)
    tm = float(np.mean(ts))
    return {"crops_per_s": world * ncrop / tm, "call_ms_max_over_ranks": tm * 1e3, "crops_per_rank": ncrop,
            "scaling": "weak", "workload": "per rank: 4096 detections from 16 1080p frames, no exchange"}


def side_measurements(args, frp, torch, gal, queries, pk, device, keep):
    """Batch-1 matching (HBM roofline) and the alignment stage (C4), each timed alone."""
    from frp_b200 import _capi
    out = {}
    B, k, tau = 1, args.k, args.tau
    dev_out = (torch.empty((1, k), dtype=torch.int64, device=device), torch.empty((1, k), device=device),
               torch.empty((1,), dtype=torch.int32, device=device), torch.empty((1,), dtype=torch.int32, device=device))
    st = torch.cuda.current_stream().cuda_stream
    q1 = queries[:, 0:1, :].contiguous()
    for i in range(3):
        gal.match_device(q1[i], 1, tau, k, 0, *dev_out, st)
    torch.cuda.synchronize()
    n = min(20, q1.shape[0])
    _capi.lib.frp_profile_enable(1)
    mids = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        gal.match_device(q1[i], 1, tau, k, 0, *dev_out, st)
        b_ = C.c_float()
        _capi.check(_capi.lib.frp_profile_last(None, C.byref(b_), None))
        mids.append(b_.value)
    e1.record()
    torch.cuda.synchronize()
    _capi.lib.frp_profile_enable(0)
    rows = gal.rows
    t = float(np.mean(mids)) / 1e3
    gb = rows * 1024 / 1e9
    b1_path = frp.last_match_path()
    out["match_batch1"] = {"queries_per_s": n / (e0.elapsed_time(e1) / 1e3),
                           "kernel": "match_gemm_tcgen05" if b1_path == "gemm_tcgen05" else "match_scan_f16",
                           "kernel_ms": t * 1e3,
                           "roofline": {"bound": "hbm", "achieved": gb / t, "peak": pk["hbm"], "unit": "GB/s",
                                        "frac": gb / t / pk["hbm"]}}
    # pure top-k (find_k_nearest semantics: FRP_MATCH_NO_THRESHOLD) at the headline size: every
    # epilogue thread keeps a live top-16 for the whole pass
    try:
        Bn = queries.shape[1]
        bo = (torch.empty((Bn, k), dtype=torch.int64, device=device), torch.empty((Bn, k), device=device),
              torch.empty((Bn,), dtype=torch.int32, device=device), torch.empty((Bn,), dtype=torch.int32, device=device))
        res_nt = {}
        for name, fl in (("sampled_cut", frp.MATCH_NO_THRESHOLD), ("no_sampling_pass", frp.MATCH_NO_THRESHOLD | _capi.MATCH_NO_SAMPLE)):
            _capi.lib.frp_profile_enable(1)
            ts_ = []
            for i in range(6):
                gal.match_device(queries[i], Bn, tau, k, fl, *bo, st)
                pb = C.c_float()
                _capi.check(_capi.lib.frp_profile_last(None, C.byref(pb), None))
                if i >= 2:
                    ts_.append(pb.value / 1e3)
            _capi.lib.frp_profile_enable(0)
            res_nt[name] = (float(np.mean(ts_)), int((bo[3] != 0).sum().item()))
        tn = res_nt["sampled_cut"][0]
        fl = 2.0 * Bn * rows * 512
        out["match_topk_nothreshold"] = {
            "workload": "pure top-%d (no threshold), batch %d vs %d x 512" % (k, Bn, rows),
            "kernel_ms": tn * 1e3, "queries_per_s_kernel": Bn / tn, "path": frp.last_match_path(),
            "uncertified": res_nt["sampled_cut"][1],
            "tflops": fl / tn / 1e12, "tensor_frac_of_burst": fl / tn / 1e12 / pk["tensor_burst"],
            "hbm_frac": gb / tn / pk["hbm"],
            "kernel_ms_without_sampling_pass": res_nt["no_sampling_pass"][0] * 1e3,
            "note": "prefilter time = sampling pass over the first ~1.5 % of the rows + cut lift + main pass"}
    except Exception as e:  # pragma: no cover
        out["match_topk_nothreshold"] = {"error": repr(e)}

    # config C2 (BASELINE configs[1]): batch 256 vs a 1 M x 512 gallery, top-10
    try:
        g1, keep1, _ = make_gallery(frp, torch, 1_000_000, 555, device, keep_rows=2048)
        q1m = make_queries(torch, 24, 256, keep1, 556, device)
        bo = (torch.empty((256, k), dtype=torch.int64, device=device), torch.empty((256, k), device=device),
              torch.empty((256,), dtype=torch.int32, device=device), torch.empty((256,), dtype=torch.int32, device=device))
        flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=device)

        def c2_step(i):
            g1.match_device(q1m[i % 24], 256, tau, k, 0, *bo, st)

        ms_ = _time_steps(torch, None, device, c2_step, 20)
        _capi.lib.frp_profile_enable(1)
        cold = []
        for i in range(5):
            flush.fill_(i)          # the 1 GB gallery is several times the L2, but start cold anyway
            c2_step(i)
            pb = C.c_float()
            _capi.check(_capi.lib.frp_profile_last(None, C.byref(pb), None))
            cold.append(pb.value)
        _capi.lib.frp_profile_enable(0)
        hq1 = q1m.cpu().pin_memory().numpy()
        for i in range(3):
            g1.match(hq1[i], tau, k)
        t0 = time.perf_counter()
        for i in range(20):
            g1.match(hq1[i % 24], tau, k)
        dt1 = (time.perf_counter() - t0) / 20
        kms = float(np.mean(cold[1:]))
        out["c2_1m"] = {"workload": "cosine 1:N, batch 256 x gallery 1 000 000 x 512, tau %.2f, top-%d (BASELINE configs[1])" % (tau, k),
                        "queries_per_s": 256 / (ms_ / 1e3), "ms_per_step": ms_, "prefilter_kernel_ms": kms,
                        "tflops": 2.0 * 256 * 1e6 * 512 / (kms / 1e3) / 1e12,
                        "tensor_frac_of_burst": 2.0 * 256 * 1e6 * 512 / (kms / 1e3) / 1e12 / pk["tensor_burst"],
                        "hbm_frac": 1.024 / (kms / 1e3) / pk["hbm"],
                        "e2e_queries_per_s": 256 / dt1, "e2e_ms_per_step": dt1 * 1e3, "path": frp.last_match_path()}
        del flush
        g1.close()
    except Exception as e:  # pragma: no cover
        out["c2_1m"] = {"error": repr(e)}

    # alignment: 4096 detections from 16 1080p frames -> 112x112 fp16 NCHW
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import frp_oracle_np as o      # cpu_baseline legs only (align and C1 below)
    F, H, W_, per = 16, 1080, 1920, 256
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(rng.integers(0, 256, (F, H, W_, 3), dtype=np.uint8)).to(device)
    fi, lm = synth_detections(F, per, H, W_, seed=7)
    d_fi = torch.from_numpy(fi).to(device)
    d_lm = torch.from_numpy(lm).to(device)
    ncrop = F * per
    blob = torch.empty((ncrop, 3, 112, 112), dtype=torch.float16, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    times, call_times = [], []
    _capi.lib.frp_profile_align_enable(1)
    for i in range(8):
        flush.fill_(i)               # evict the frames / outputs from L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frp.align.norm_crop_batch_device(frames, F, H, W_, 3 * W_, d_fi, d_lm, ncrop, blob, stream=st)
        e1.record()
        torch.cuda.synchronize()
        km = C.c_float()
        _capi.check(_capi.lib.frp_profile_align_last(C.byref(km)))
        if i >= 3:
            times.append(km.value / 1e3)               # the kernel alone (events inside the library)
            call_times.append(e0.elapsed_time(e1) / 1e3)   # whole call incl. launch overhead
    _capi.lib.frp_profile_align_enable(0)
    t = float(np.mean(times))
    bytes_alg = F * H * W_ * 3 + ncrop * (75264 + 40)
    out["align"] = {"crops_per_s": ncrop / float(np.mean(call_times)), "kernel_ms": t * 1e3,
                    "call_ms": float(np.mean(call_times)) * 1e3, "crops": ncrop,
                    "workload": "4096 detections from 16 1080p frames -> 112x112 fp16 NCHW, L2 flushed between runs",
                    "roofline": {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": bytes_alg / t / 1e9 / pk["hbm"],
                                 "algorithmic_bytes": bytes_alg,
                                 "definition": "unique source bytes (16 frames) + landmarks + fp16 output"}}
    # CPU alignment baseline on a bounded sample (cv2 = the reference's own library)
    try:
        import cv2
        fr_h = frames[:1].cpu().numpy()
        t0 = time.perf_counter()
        m = 128
        for i in range(m):
            M = o.estimate_norm(lm[i])
            a = cv2.warpAffine(fr_h[0], M, (112, 112), borderValue=0.0)
            o.blob_from_aligned(a)
        dt = time.perf_counter() - t0
        out["align"]["cpu_baseline"] = {"value": m / dt, "unit": "crops/s", "cores": 1, "kind": "port",
                                        "sample": "%d crops: numpy Umeyama + cv2.warpAffine + normalise, 1 thread loop" % m}
    except Exception as e:  # pragma: no cover
        out["align"]["cpu_baseline"] = {"error": str(e)}

    # 128-d L2 (face_recognition encodings) at scale: 4M x 128, tolerance 0.6, batch 1 and 256.
    # Arithmetic intensity of the batched contraction is B flop per gallery byte whatever the row
    # length, so batch 256 sits on the same tensor/HBM ridge as the 512-d headline: both
    # fractions are reported.
    try:
        N2_ = 4_000_000
        g2, keep2, _ = make_gallery(frp, torch, N2_, 77, device, keep_rows=4096, dim=128, metric=frp.METRIC_L2, scale=0.05)
        gen = torch.Generator(device=device).manual_seed(78)
        res = {}
        for Bx in (1, 256):
            qx = torch.randn((6, Bx, 128), generator=gen, device=device) * 0.05
            # half the queries are noisy copies of gallery rows from the whole row range (d ~ 0.23)
            sel = torch.randint(0, keep2.shape[0], (6, Bx), generator=gen, device=device)
            planted = torch.rand((6, Bx, 1), generator=gen, device=device) < 0.5
            qx = torch.where(planted, keep2[sel] + torch.randn((6, Bx, 128), generator=gen, device=device) * 0.02, qx)
            bo = (torch.empty((Bx, k), dtype=torch.int64, device=device), torch.empty((Bx, k), device=device),
                  torch.empty((Bx,), dtype=torch.int32, device=device), torch.empty((Bx,), dtype=torch.int32, device=device))
            _capi.lib.frp_profile_enable(1)
            ts_ = []
            for i in range(6):
                g2.match_device(qx[i], Bx, 0.6, k, 0, *bo, st)
                pb = C.c_float()
                _capi.check(_capi.lib.frp_profile_last(None, C.byref(pb), None))
                if i >= 2:
                    ts_.append(pb.value / 1e3)
            _capi.lib.frp_profile_enable(0)
            tm_ = float(np.mean(ts_))
            gb_ = N2_ * 256 / 1e9
            fl_ = 2.0 * Bx * N2_ * 128
            res[str(Bx)] = {"kernel_ms": tm_ * 1e3, "path": frp.last_match_path(), "hbm_gbs": gb_ / tm_,
                            "hbm_frac": gb_ / tm_ / pk["hbm"], "queries_per_s_kernel": Bx / tm_,
                            "tflops": fl_ / tm_ / 1e12, "tensor_frac_of_burst": fl_ / tm_ / 1e12 / pk["tensor_burst"],
                            "tensor_frac_of_sustained": fl_ / tm_ / 1e12 / pk["tensor"],
                            "mean_matches_per_query": float(bo[2].float().mean().item())}
            # which roofline binds: B flop per gallery byte against the ridge of the measured peaks
            ridge_ = pk["tensor"] * 1e12 / (pk["hbm"] * 1e9)
            res[str(Bx)]["bound"] = "tensor" if Bx > ridge_ else "hbm"
            res[str(Bx)]["ridge_flop_per_byte"] = ridge_
        out["match_l2_128d"] = {"workload": "L2, tolerance 0.6, 4M x 128 gallery (256 B fp16 rows), half the queries "
                                            "planted on rows spread over the whole gallery", "batches": res}
        g2.close()
    except Exception as e:  # pragma: no cover
        out["match_l2_128d"] = {"error": repr(e)}

    # config C1: face_recognition compare_faces, 1 query vs 1k synthetic 128-d encodings,
    # tolerance 0.6, through the drop-in FaceService (all 1 000 result dicts, like the reference)
    try:
        G1 = synth_l2_gallery(1000, 128, seed=1234)
        Q1, _ = synth_l2_queries(G1, 64, seed=4321)
        enc = frp.GalleryDict()
        enc.bulk_load(["id%07d" % i for i in range(len(G1))], G1)
        fs = frp.FaceService(enc, tolerance=0.6, raise_errors=True)
        ref = {"id%07d" % i: g.tolist() for i, g in enumerate(G1)}
        for q in Q1[:4]:
            fs.compare_faces(q)
        t0 = time.perf_counter()
        for q in Q1:
            fs.compare_faces(q)
        t_gpu = (time.perf_counter() - t0) / len(Q1)
        t0 = time.perf_counter()
        for q in Q1:
            fs.match_topk(q[None], 0.6, 5)
        t_topk = (time.perf_counter() - t0) / len(Q1)
        t0 = time.perf_counter()
        for q in Q1[:16]:
            o.compare_faces(ref, q, 0.6)
        t_cpu = (time.perf_counter() - t0) / 16
        # the UNMODIFIED reference through oracle/ref_loader.py where /root/reference exists (the
        # authoring container); the GPU box has no reference tree, so the port is the baseline there
        t_live, live_kind = None, "port (no /root/reference on this box)"
        try:
            import ref_loader
            if ref_loader.reference_available():
                rfs, RENC, _ = ref_loader.load_reference()
                RENC.clear()
                for n_, g_ in ref.items():
                    RENC[n_] = g_
                rfs.tolerance = 0.6
                rfs.compare_faces(Q1[0])
                t0 = time.perf_counter()
                for q in Q1[:16]:
                    rfs.compare_faces(q)
                t_live = (time.perf_counter() - t0) / 16
                live_kind = "reference (unmodified face_service.compare_faces via oracle/ref_loader.py)"
                RENC.clear()
        except Exception as e:  # pragma: no cover
            live_kind = "port (reference import failed: %r)" % (e,)
        # hot-call latencies: store_face = duplicate scan + in-place row update
        t0 = time.perf_counter()
        for i in range(64):
            fs.store_face("id%07d" % (i * 7), G1[i * 7])
        t_store = (time.perf_counter() - t0) / 64
        gq = np.ascontiguousarray(Q1[:1], dtype=np.float32)
        gg = enc.gallery
        for _ in range(8):
            gg.match(gq, 0.6, 5)
        t0 = time.perf_counter()
        for _ in range(256):
            gg.match(gq, 0.6, 5)
        t_native = (time.perf_counter() - t0) / 256
        out["c1_compare_faces"] = {
            "workload": "1 query vs 1 000 x 128-d, tolerance 0.6 (BASELINE configs[0])",
            "drop_in_compare_faces_ms": t_gpu * 1e3, "drop_in_queries_per_s": 1.0 / t_gpu,
            "match_topk_ms": t_topk * 1e3, "match_topk_queries_per_s": 1.0 / t_topk,
            "frp_match_call_us": t_native * 1e6, "store_face_us": t_store * 1e6,
            "cpu_port_compare_faces_ms": t_cpu * 1e3, "cpu_port_queries_per_s": 1.0 / t_cpu,
            "cpu_reference_compare_faces_ms": None if t_live is None else t_live * 1e3,
            "cpu_baseline_kind": live_kind,
            "note": "compare_faces returns all 1 000 dicts (O(N) Python by contract, face_service.py:414-429); "
                    "match_topk is threshold-then-top-5 on the device; frp_match_call_us = one Gallery.match "
                    "(ctypes + frp_match, 1 query, top-5)"}
    except Exception as e:  # pragma: no cover
        out["c1_compare_faces"] = {"error": repr(e)}

    # N4: cluster_faces on the device (face_service.py:552-585): 100 k x 128 faces, 20 per identity
    try:
        ids, per_id = 5000, 20
        gen = torch.Generator(device=device).manual_seed(91)
        cen = torch.randn((ids, 128), generator=gen, device=device) * 0.05
        pts = (cen[:, None, :] + torch.randn((ids, per_id, 128), generator=gen, device=device) * 0.02).reshape(-1, 128)
        pts = pts[torch.randperm(ids * per_id, generator=gen, device=device)].contiguous()
        gc = frp.Gallery(128, frp.METRIC_L2, device=device.index or 0, capacity=ids * per_id)
        gc.append_device(pts, stream=st)
        torch.cuda.synchronize()
        gc.cluster(0.6)
        t0 = time.perf_counter()
        labels, ncl, stats = gc.cluster(0.6)
        t_cl = time.perf_counter() - t0
        out["cluster_faces"] = {"workload": "greedy seed clustering of %d x 128-d faces (%d identities x %d), threshold 0.6"
                                            % (ids * per_id, ids, per_id),
                                "seconds": t_cl, "faces_per_s": ids * per_id / t_cl, "clusters": int(ncl), **stats,
                                "note": "the reference's O(N^2) Python loop needs N^2/2 = 5e9 face_distance calls here"}
        gc.close()
    except Exception as e:  # pragma: no cover
        out["cluster_faces"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
