"""Small driver for ncu captures of one matcher configuration (not a benchmark).

    python tools/prof_match.py --rows 4000000 --dim 128 --metric l2 --batch 256 --k 10 --iters 3

Builds a synthetic gallery on cuda:0, runs `iters` matches through the device API and prints
the in-library kernel time of the last one.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-recognition-platform_b200"))
import frp_b200 as frp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--metric", default="l2")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--tau", type=float, default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--flags", type=lambda v: int(v, 0), default=0, help="FRP_MATCH_* flags, e.g. 0x4 = force tcgen05")
    a = ap.parse_args()
    metric = frp.METRIC_L2 if a.metric == "l2" else frp.METRIC_COS
    tau = a.tau if a.tau is not None else (0.6 if a.metric == "l2" else 0.4)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    gal = frp.Gallery(a.dim, metric, device=0, capacity=a.rows)
    step = 1 << 20
    scale = 0.09 if a.metric == "l2" else 1.0
    for r0 in range(0, a.rows, step):
        n = min(step, a.rows - r0)
        x = torch.randn(n, a.dim, device=dev, generator=g, dtype=torch.float32) * scale
        gal.append_device(x.data_ptr(), n)
    q = (torch.randn(a.batch, a.dim, device=dev, generator=g, dtype=torch.float32) * scale).contiguous()
    idx = torch.empty(a.batch, a.k, dtype=torch.int64, device=dev)
    sc = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
    cnt = torch.empty(a.batch, dtype=torch.int32, device=dev)
    st = torch.empty(a.batch, dtype=torch.int32, device=dev)
    import ctypes as C
    from frp_b200 import _capi
    _capi.lib.frp_profile_enable(1)
    for _ in range(a.iters):
        gal.match_device(q.data_ptr(), a.batch, tau, a.k, a.flags, idx.data_ptr(), sc.data_ptr(), cnt.data_ptr(),
                         st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    t = [C.c_float(), C.c_float(), C.c_float()]
    _capi.check(_capi.lib.frp_profile_last(*[C.byref(x) for x in t]))
    print("path", frp.last_match_path(), "prep/filter/rerank ms", [round(x.value, 4) for x in t],
          "status max", int(st.max()), "count mean", float(cnt.float().mean()))


if __name__ == "__main__":
    main()
