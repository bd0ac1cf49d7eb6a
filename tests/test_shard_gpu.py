"""Row sharding on one GPU: R "virtual shards" (R galleries with global bases) are
matched one after another, their packed records concatenated exactly as
ncclAllGather would, and merged by frp_merge_packed_device.  The result must not
depend on the number of shards.  The real NCCL path is exercised by
`bench.py --gpus N`."""
import numpy as np
import pytest
import torch

import frp_b200
from frp_b200 import Gallery, METRIC_COS, METRIC_L2, MATCH_NO_THRESHOLD
from frp_b200.shard import ShardedMatcher, record_bytes, shard_bounds
import frp_oracle_np as o

pytestmark = pytest.mark.gpu


def _sharded(G, Q, metric, tau, k, R, flags=0):
    N = len(G)
    B = len(Q)
    dev = torch.device("cuda")
    q_dev = torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32)).to(dev)
    nb = record_bytes(B, k)
    recv = torch.zeros(R * nb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    keep = []
    for r in range(R):
        lo, hi = shard_bounds(N, R, r)
        g = Gallery(G.shape[1], metric)
        g.append(G[lo:hi])
        m = ShardedMatcher(g, N, r, R)
        assert (m.lo, m.hi) == (lo, hi)
        send = recv[r * nb:(r + 1) * nb]
        idx, score, count, status = ShardedMatcher._views(send, B, k, torch)
        m._local_match(q_dev, B, tau, k, flags, idx, score, count, status, st)     # frp_match_record_device
        keep.append(g)
    out_idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    out_score = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_count = torch.empty((B,), dtype=torch.int32, device=dev)
    anyst = torch.zeros((1,), dtype=torch.int32, device=dev)
    m._cuda_merge(metric, R, B, k, recv, out_idx, out_score, out_count, anyst, st)
    torch.cuda.synchronize()
    return out_idx.cpu().numpy(), out_score.cpu().numpy(), out_count.cpu().numpy(), int(anyst.item())


@pytest.mark.parametrize("metric", [METRIC_COS, METRIC_L2])
def test_result_is_independent_of_shard_count(metric):
    N, D, B, k = 20_011, 128 if metric == METRIC_L2 else 512, 6, 10
    if metric == METRIC_COS:
        G = o.synth_cos_gallery(N, D, seed=1)
        Q, _ = o.synth_cos_queries(G, B, seed=2)
        tau = 0.3
        G[15_000] = G[100]
        G[19_999] = G[100]          # ties across shards resolve by global row
        Q[0] = G[100]
        # near-ties across shards: rows whose float64 scores differ by ~1e-9 (below float32
        # resolution) must be ordered by SCORE, not by row — the shard record carries float64
        q1 = o.normed(Q[1].astype(np.float64))
        for j, (row, c) in enumerate(((18_000, 0.61), (2_000, 0.61 - 2e-9), (9_000, 0.61 - 4e-9))):
            u = np.random.default_rng(100 + j).standard_normal(D)
            u -= (u @ q1) * q1
            u /= np.linalg.norm(u)
            G[row] = (c * q1 + np.sqrt(1 - c * c) * u).astype(np.float32)
        widx, wsc, wcnt = o.cos_match_topk(G, Q, tau, k, True)
    else:
        G = o.synth_l2_gallery(N, D, seed=3).astype(np.float32)
        Q, _ = o.synth_l2_queries(G.astype(np.float64), B, seed=4)
        Q = Q.astype(np.float32)
        tau = 0.6
        widx, wsc, wcnt = o.l2_match_topk(G, Q, tau, k, True)
    base = None
    for R in (1, 2, 3, 8):
        idx, sc, cnt, anyst = _sharded(G, Q, metric, tau, k, R)
        assert anyst == 0
        assert np.array_equal(cnt, wcnt.astype(np.int32))
        for b in range(B):
            n = int(cnt[b])
            if not np.array_equal(idx[b], widx[b]):
                # only the planted near-ties (scores ~1e-8 apart: the fp32 master of a
                # re-normalised row vs the oracle's float64 normalisation) may swap
                j = np.nonzero(idx[b, :n] != widx[b, :n])[0]
                assert metric == METRIC_COS and b == 1 and sorted(idx[b, :n]) == sorted(widx[b, :n])
                assert np.ptp(wsc[b, j.min():j.max() + 1]) < 1e-7
            assert np.abs(sc[b, :n] - wsc[b, :n]).max(initial=0) < 2e-6
        if base is None:
            base = (idx, sc, cnt)
            single = Gallery(G.shape[1], metric)
            single.append(G)
            sidx, ssc, scnt, _ = single.match(Q, tau, k)
            # one gallery holding all rows == the merged shards, bit for bit
            assert np.array_equal(sidx, idx) and np.array_equal(ssc, sc, equal_nan=True) and np.array_equal(scnt, cnt)
            single.close()
        else:
            # the float64 shard record makes the merged order independent of the shard count,
            # including the near-ties whose float32 scores are equal
            assert np.array_equal(idx, base[0]) and np.array_equal(sc, base[1], equal_nan=True)
    if metric == METRIC_COS:
        assert {18_000, 2_000, 9_000} <= set(idx[1, :cnt[1]].tolist())


def test_match_certified_takes_the_fallback_ladder():
    G = o.synth_cos_gallery(6000, 512, seed=7)
    G[2000:2100] = G[3]                       # 101 identical rows: more ties than a candidate list holds
    g = Gallery(512, METRIC_COS)
    g.append(G)
    m = ShardedMatcher(g, 6000, 0, 1)
    q = torch.from_numpy(G[3:4].copy()).cuda()
    idx, sc, cnt, level = m.match_certified(q, 0.9, 10)
    torch.cuda.synchronize()
    assert level in (1, 2)
    assert idx.cpu().numpy()[0].tolist() == [3] + list(range(2000, 2009))
    idx, sc, cnt, level = m.match_certified(torch.from_numpy(G[10:12].copy()).cuda(), 0.9, 10)
    assert level == 0 and cnt.cpu().numpy().tolist() == [1, 1]


def test_fused_push_merge_loopback_equals_allgather_merge():
    """frp_push_merge_device with one rank (its own exchange buffer as the only peer): the
    push / flag / wait / merge kernel must give what the all-gather + merge path gives,
    over several epochs (double-buffered slots)."""
    G = o.synth_cos_gallery(9000, 512, seed=9)
    g = Gallery(512, METRIC_COS)
    g.append(G)
    a = ShardedMatcher(g, 9000, 0, 1, exchange="p2p")
    b = ShardedMatcher(g, 9000, 0, 1, exchange="nccl")
    for step in range(5):
        Q, _ = o.synth_cos_queries(G, 37, seed=50 + step)
        q = torch.from_numpy(Q).cuda()
        ra = [x.clone() for x in a.match(q, 0.3, 10)]
        rb = [x.clone() for x in b.match(q, 0.3, 10)]
        torch.cuda.synchronize()
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[2], rb[2])
        assert torch.equal(torch.nan_to_num(ra[1]), torch.nan_to_num(rb[1]))
        assert int(ra[3].item()) == 0
        widx, wsc, wcnt = o.cos_match_topk(G, Q, 0.3, 10, True)
        assert np.array_equal(ra[0].cpu().numpy(), widx)


def test_single_rank_sharded_matcher_runs_end_to_end():
    G = o.synth_cos_gallery(5000, 512, seed=5)
    Q, _ = o.synth_cos_queries(G, 4, seed=6)
    g = Gallery(512, METRIC_COS)
    g.append(G)
    m = ShardedMatcher(g, 5000, 0, 1)
    idx, sc, cnt, anyst = m.match(torch.from_numpy(Q).cuda(), 0.3, 5, MATCH_NO_THRESHOLD)
    torch.cuda.synchronize()
    widx, wsc, wcnt = o.cos_match_topk(G, Q, 0.3, 5, False)
    assert np.array_equal(idx.cpu().numpy(), widx) and int(anyst.item()) == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_buffers_on_another_gpu_are_refused_not_faulted():
    """Device entry points name a device (or a gallery that lives on one); buffers of a different
    GPU are rejected with FRP_E_INVALID instead of faulting inside a kernel."""
    from frp_b200 import align
    G = o.synth_cos_gallery(2000, 512, seed=5)
    gal = Gallery(512, METRIC_COS, device=0)
    gal.append(G)
    q = torch.zeros((2, 512), device="cuda:1")
    outs = (torch.empty((2, 3), dtype=torch.int64, device="cuda:1"), torch.empty((2, 3), device="cuda:1"),
            torch.empty((2,), dtype=torch.int32, device="cuda:1"), torch.empty((2,), dtype=torch.int32, device="cuda:1"))
    with pytest.raises(frp_b200.FrpError):
        gal.match_device(q, 2, 0.3, 3, 0, *outs)
    frames = torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device="cuda:1")
    fi = torch.zeros((1,), dtype=torch.int32, device="cuda:1")
    lm = torch.from_numpy(o.ARCFACE_DST[None].astype(np.float32)).to("cuda:1")
    blob = torch.empty((1, 3, 112, 112), dtype=torch.float16, device="cuda:1")
    with pytest.raises(frp_b200.FrpError):
        align.norm_crop_batch_device(frames, 1, 64, 64, 192, fi, lm, 1, blob, device=0)
    # the device is taken from the frames tensor when it is not named
    align.norm_crop_batch_device(frames, 1, 64, 64, 192, fi, lm, 1, blob)
    torch.cuda.synchronize(1)
    assert torch.all(blob == -1.0)


# ----------------------------------------------------------------------------------------------
# the real thing: two ranks, two GPUs, NCCL (and the fused peer-memory exchange)
# ----------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out_q):
    import os
    import sys
    import traceback
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (os.path.join(root, "face-recognition-platform_b200"), os.path.join(root, "oracle")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import numpy as np
        import torch
        import torch.distributed as dist
        import frp_oracle_np as o
        from frp_b200 import Gallery, METRIC_COS, MATCH_NO_THRESHOLD
        from frp_b200.shard import CertifiedPipeline, ShardedMatcher, shard_bounds
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        N, D, B, k, tau = 300_007, 512, 64, 10, 0.35
        G = o.synth_cos_gallery(N, D, seed=1)
        G[250_000] = G[5]
        G[N - 1] = G[5]                       # exact ties on different ranks
        steps = 6
        Qs = [o.synth_cos_queries(G, B, seed=10 + s)[0] for s in range(steps)]
        Qs[0][0] = G[5]
        lo, hi = shard_bounds(N, world, rank)
        g = Gallery(D, METRIC_COS, device=rank)
        g.append(G[lo:hi])
        full = Gallery(D, METRIC_COS, device=rank)
        full.append(G)
        ok, notes = True, []
        for exchange in ("nccl", "p2p"):
            m = ShardedMatcher(g, N, rank, world, exchange=exchange)
            for s in range(steps):
                for flags in (0, MATCH_NO_THRESHOLD):
                    q = torch.from_numpy(Qs[s]).to(dev)
                    oi, os_, oc, fl = [x.clone() for x in m.match(q, tau, k, flags)]
                    torch.cuda.synchronize()
                    wi, ws, wc, wst = full.match(Qs[s], tau, k, flags)
                    same = (np.array_equal(oi.cpu().numpy(), wi) and np.array_equal(oc.cpu().numpy(), wc) and
                            np.array_equal(os_.cpu().numpy(), ws, equal_nan=True) and int(fl.item()) == 0)
                    if not same:
                        ok = False
                        notes.append("%s step %d flags %d differs" % (exchange, s, flags))
        # pipelined submit / result and the host-buffer pipeline give the same answers
        m = ShardedMatcher(g, N, rank, world)
        tickets, outs = [], []
        for s in range(steps):
            q = torch.from_numpy(Qs[s]).to(dev)
            tickets.append((m.submit(q, tau, k), q))
            if len(tickets) == 2:
                t, _ = tickets.pop(0)
                outs.append([x.clone() for x in m.result(t)])
        while tickets:
            t, _ = tickets.pop(0)
            outs.append([x.clone() for x in m.result(t)])
        torch.cuda.synchronize()
        pipe = CertifiedPipeline(m, B, D, k, dev)
        host = []
        for s in range(steps):
            r = pipe.push(torch.from_numpy(Qs[s]).pin_memory(), tau)
            if r is not None:
                host.append([x.clone() if hasattr(x, "clone") else x for x in r])
        r = pipe.flush()
        host.append([x.clone() if hasattr(x, "clone") else x for x in r])
        for s in range(steps):
            wi, ws, wc, _ = full.match(Qs[s], tau, k)
            if not (np.array_equal(outs[s][0].cpu().numpy(), wi) and np.array_equal(outs[s][1].cpu().numpy(), ws, equal_nan=True)):
                ok = False
                notes.append("pipelined step %d differs" % s)
            if not (np.array_equal(host[s][0].numpy(), wi) and np.array_equal(host[s][2].numpy(), wc) and host[s][3] == 0):
                ok = False
                notes.append("host pipeline step %d differs" % s)
        assert outs[0][0][0, :3].tolist() == [5, 250_000, N - 1]
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, ok, "; ".join(notes)))
    except Exception:      # pragma: no cover
        out_q.put((rank, False, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(600)
def test_two_rank_nccl_and_p2p_equal_single_gallery():
    """ShardedMatcher over two real ranks (NCCL all-gather + merge kernel, the fused peer-memory
    exchange, the two-slot pipeline and the host-buffer pipeline) == one gallery holding all
    rows, bit for bit.  Run by the builder on a 2-GPU box; its log is kept under profiles/."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, ok, msg in res:
        assert ok, "rank %d: %s" % (rank, msg)
