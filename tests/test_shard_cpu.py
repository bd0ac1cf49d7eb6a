"""N>1 path on CPU: world_size-2 gloo run of ShardedMatcher's partition / pack /
all-gather / merge plumbing, plain steps and the two-slot pipeline (submit / result).  The two CUDA pieces (local match, merge) are injected
with oracle-based stand-ins here; the real kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        for p in (os.path.join(ROOT, "face-recognition-platform_b200"), os.path.join(ROOT, "oracle")):
            sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        import frp_oracle_np as o
        from frp_b200.shard import ShardedMatcher, shard_bounds, unpack_records
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        N, D, B, k, tau = 1003, 64, 9, 6, 0.25
        G = o.synth_cos_gallery(N, D, seed=1)
        G[500] = G[10]            # duplicate rows on different shards: tie order = global row
        G[900] = G[10]
        Q, _ = o.synth_cos_queries(G, B, seed=2)
        Q[0] = G[10]
        lo, hi = shard_bounds(N, world, rank)

        def local_match(q_dev, B_, tau_, k_, flags, idx, score, count, status, stream):
            i, s, c = o.cos_match_topk(G[lo:hi], q_dev.numpy(), tau_, k_, True)
            i = np.where(i >= 0, i + lo, -1)
            idx.copy_(torch.from_numpy(i))
            score.copy_(torch.from_numpy(np.nan_to_num(s, nan=0.0).astype(np.float32)))
            count.copy_(torch.from_numpy(c.astype(np.int32)))
            status.zero_()

        def merge(metric, R, B_, k_, records, out_idx, out_score, out_count, any_status, stream):
            idx, sc, cnt, st = unpack_records(records.numpy(), R, B_, k_)
            for b in range(B_):
                ent = [(-float(sc[r, b, j]), int(idx[r, b, j])) for r in range(R) for j in range(cnt[r, b])]
                ent.sort()
                ent = ent[:k_]
                out_count[b] = len(ent)
                for j, (ns, ii) in enumerate(ent):
                    out_idx[b, j] = ii
                    out_score[b, j] = -ns
                for j in range(len(ent), k_):
                    out_idx[b, j] = -1
            any_status.fill_(int(st.max()))

        m = ShardedMatcher(None, N, rank, world, local_match=local_match, merge=merge)
        assert (m.lo, m.hi) == (lo, hi)
        oi, os_, oc, anyst = m.match(torch.from_numpy(Q), tau, k, metric=1)
        wi, ws, wc = o.cos_match_topk(G, Q, tau, k, True)
        ok = np.array_equal(oi.numpy(), wi) and np.array_equal(oc.numpy(), wc.astype(np.int32))
        ok = ok and int(anyst[0]) == 0
        for b in range(B):
            ok = ok and np.allclose(os_.numpy()[b, :wc[b]], ws[b, :wc[b]], atol=1e-6)
        # two-slot pipeline (submit / result): two steps in flight give the answers of two plain
        # steps, a third submit before a result is refused, tickets are checked
        Q2 = np.ascontiguousarray(Q[::-1])
        t0 = m.submit(torch.from_numpy(Q), tau, k, metric=1)
        t1 = m.submit(torch.from_numpy(Q2), tau, k, metric=1)
        try:
            m.submit(torch.from_numpy(Q), tau, k, metric=1)
            ok = False
        except RuntimeError:
            pass
        r0 = [x.clone() for x in m.result(t0)]
        r1 = [x.clone() for x in m.result(t1)]
        try:
            m.result(t0)
            ok = False
        except RuntimeError:
            pass
        w2 = o.cos_match_topk(G, Q2, tau, k, True)
        ok = ok and np.array_equal(r0[0].numpy(), wi) and np.array_equal(r0[2].numpy(), wc.astype(np.int32))
        ok = ok and np.array_equal(r1[0].numpy(), w2[0]) and np.array_equal(r1[2].numpy(), w2[2].astype(np.int32))
        t2 = m.submit(torch.from_numpy(Q), tau, k, metric=1)       # slot 0 again
        ok = ok and t2 == 2 and np.array_equal(m.wait_host(t2)[0].numpy(), wi)
        # every rank holds the same global answer
        gathered = [torch.zeros_like(oi) for _ in range(world)]
        dist.all_gather(gathered, oi)
        ok = ok and all(torch.equal(g, oi) for g in gathered)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.timeout(180)
def test_sharded_match_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    for rank, ok, msg in res:
        assert ok, "rank %d failed: %s" % (rank, msg)


def test_match_certified_walks_the_fallback_ladder():
    """Host logic of ShardedMatcher.match_certified on CPU tensors (single rank, injected
    stand-ins): an uncertified step is repeated with FRP_MATCH_FORCE_EXACT, then with
    FORCE_EXACT | WIDE, forced-kernel flags are dropped on the way, and the level is reported."""
    for p in (os.path.join(ROOT, "face-recognition-platform_b200"),):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    from frp_b200._capi import MATCH_FORCE_EXACT, MATCH_FORCE_GEMM, MATCH_NO_THRESHOLD, MATCH_WIDE
    from frp_b200.shard import ShardedMatcher, unpack_records

    def run(certify_at):
        seen = []

        def local_match(q_dev, B, tau, k, flags, idx, score, count, status, stream):
            seen.append(flags)
            idx.fill_(3)
            score.fill_(0.5)
            count.fill_(1)
            level = 2 if (flags & MATCH_WIDE) else (1 if (flags & MATCH_FORCE_EXACT) else 0)
            status.fill_(0 if level >= certify_at else 1)

        def merge(metric, R, B, k, records, out_idx, out_score, out_count, any_status, stream):
            idx, sc, cnt, st = unpack_records(records.numpy(), R, B, k)
            out_idx.copy_(torch.from_numpy(idx[0]))
            out_score.copy_(torch.from_numpy(sc[0]))
            out_count.copy_(torch.from_numpy(cnt[0]))
            any_status.fill_(int(st.max()))

        m = ShardedMatcher(None, 100, 0, 1, local_match=local_match, merge=merge)
        q = torch.zeros((4, 8))
        oi, os_, oc, level = m.match_certified(q, 0.3, 2, flags=MATCH_FORCE_GEMM | MATCH_NO_THRESHOLD, metric=1)
        assert oi.shape == (4, 2) and int(oc[0]) == 1
        return level, seen

    level, seen = run(0)
    assert level == 0 and seen == [MATCH_FORCE_GEMM | MATCH_NO_THRESHOLD]
    level, seen = run(1)
    assert level == 1 and seen[1] == MATCH_NO_THRESHOLD | MATCH_FORCE_EXACT
    level, seen = run(2)
    assert level == 2 and seen[2] == MATCH_NO_THRESHOLD | MATCH_FORCE_EXACT | MATCH_WIDE and len(seen) == 3
    level, seen = run(3)
    assert level == 3 and len(seen) == 3


def test_certified_pipeline_defers_the_flag_check_and_repairs_flagged_steps():
    """CertifiedPipeline on CPU tensors (single rank, injected stand-ins): push returns the
    PREVIOUS step's result, the host copies carry the merged answer, a step whose certificate
    flag is set is repeated through the ladder before it is handed out (level > 0), and flush
    drains the last step."""
    for p in (os.path.join(ROOT, "face-recognition-platform_b200"),):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    from frp_b200._capi import MATCH_FORCE_EXACT
    from frp_b200.shard import CertifiedPipeline, ShardedMatcher, unpack_records
    calls = []

    def local_match(q_dev, B, tau, k, flags, idx, score, count, status, stream):
        # "answer" = the first query component as row id; step 2 (marker 20) is uncertified
        # until it is repeated against the exact master
        marker = int(q_dev[0, 0])
        calls.append((marker, flags))
        idx.fill_(marker)
        score.fill_(0.25)
        count.fill_(1)
        status.fill_(1 if (marker == 20 and not (flags & MATCH_FORCE_EXACT)) else 0)

    def merge(metric, R, B, k, records, out_idx, out_score, out_count, any_status, stream):
        idx, sc, cnt, st = unpack_records(records.numpy(), R, B, k)
        out_idx.copy_(torch.from_numpy(idx[0]))
        out_score.copy_(torch.from_numpy(sc[0]))
        out_count.copy_(torch.from_numpy(cnt[0]))
        any_status.fill_(int(st.max()))

    import types
    m = ShardedMatcher(types.SimpleNamespace(metric=1), 100, 0, 1, local_match=local_match, merge=merge,
                       set_base=False)
    pipe = CertifiedPipeline(m, 4, 8, 2, torch.device("cpu"))
    outs = []
    for marker in (10, 20, 30):
        r = pipe.push(torch.full((4, 8), float(marker)), 0.3)
        outs.append(None if r is None else (int(r[0][0, 0]), int(r[2][0]), r[3]))
    last = pipe.flush()
    outs.append((int(last[0][0, 0]), int(last[2][0]), last[3]))
    assert pipe.flush() is None
    assert outs == [None, (10, 1, 0), (20, 1, 1), (30, 1, 0)]
    # the flagged step ran twice: once in the pipeline, once with FORCE_EXACT
    assert [c for c in calls if c[0] == 20] == [(20, 0), (20, MATCH_FORCE_EXACT)]
