"""Where a single-GPU bench step spends its time on the device (not a benchmark).

Runs the same call sequence as ShardedMatcher.match at world size 1 with CUDA events between
the pieces, and the plain Gallery.match_device call for comparison.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-recognition-platform_b200"))
import frp_b200 as frp  # noqa: E402


def main():
    rows, B, k, tau = int(os.environ.get("ROWS", 10_000_000)), 256, 10, 0.4
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    gal = frp.Gallery(512, frp.METRIC_COS, device=0, capacity=rows)
    for r0 in range(0, rows, 1 << 20):
        n = min(1 << 20, rows - r0)
        gal.append_device(torch.randn(n, 512, device=dev, generator=g).data_ptr(), n)
    q = torch.randn(B, 512, device=dev, generator=g)
    m = frp.ShardedMatcher(gal, rows, 0, 1)
    st = torch.cuda.current_stream().cuda_stream
    nbytes, send, recv, out_idx, out_score, out_count, any_status = m._buffers(B, k, dev)
    idx, score, count, status = m._views(send, B, k, torch)
    for _ in range(5):
        m.match(q, tau, k)
    torch.cuda.synchronize()
    N = 30
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(N)]
    for i in range(N):
        ev[i][0].record()
        gal.match_device(q, B, tau, k, 0, idx, score, count, status, st)
        ev[i][1].record()
        recv.copy_(send)
        ev[i][2].record()
        m._merge(gal.metric, 1, B, k, recv, out_idx, out_score, out_count, any_status, st)
        ev[i][3].record()
    torch.cuda.synchronize()
    d = np.array([[ev[i][j].elapsed_time(ev[i][j + 1]) for j in range(3)] for i in range(N)])
    print("match_device / copy / merge ms:", d.mean(0).round(4).tolist())
    print("step (event to event) ms:", float(np.mean([ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(N - 1)])))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        m.match(q, tau, k)
    e1.record()
    torch.cuda.synchronize()
    print("ShardedMatcher.match ms/step:", e0.elapsed_time(e1) / N)
    e0.record()
    for i in range(N):
        gal.match_device(q, B, tau, k, 0, idx, score, count, status, st)
    e1.record()
    torch.cuda.synchronize()
    print("Gallery.match_device ms/step:", e0.elapsed_time(e1) / N)
    from frp_b200 import _capi
    import ctypes as C
    _capi.lib.frp_profile_enable(1)
    gal.match_device(q, B, tau, k, 0, idx, score, count, status, st)
    t = [C.c_float(), C.c_float(), C.c_float()]
    _capi.check(_capi.lib.frp_profile_last(*[C.byref(x) for x in t]))
    print("inside match_device: prep/prefilter/rerank ms", [round(x.value, 4) for x in t])


if __name__ == "__main__":
    main()
