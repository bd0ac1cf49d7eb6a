"""Back-to-back steps of the device matcher (prep -> prefilter -> re-rank), average step time.

    FRP_PDL=0|1 python tools/prof_step.py --rows 1000000 --dim 512 --batch 256 --steps 400

Prints one JSON line.  Used for the A/B of programmatic dependent launches, where the interesting
quantity is the time between the kernels, not the kernels.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-recognition-platform_b200"))
import frp_b200 as frp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    gal = frp.Gallery(a.dim, frp.METRIC_COS, device=0, capacity=a.rows)
    step = 1 << 20
    for r0 in range(0, a.rows, step):
        n = min(step, a.rows - r0)
        gal.append_device(torch.randn(n, a.dim, device=dev, generator=g).data_ptr(), n)
    q = torch.randn(8, a.batch, a.dim, device=dev, generator=g).contiguous()
    k = 10
    idx = torch.empty(a.batch, k, dtype=torch.int64, device=dev)
    sc = torch.empty(a.batch, k, dtype=torch.float32, device=dev)
    cnt = torch.empty(a.batch, dtype=torch.int32, device=dev)
    st = torch.empty(a.batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run(n):
        for i in range(n):
            gal.match_device(q[i & 7].data_ptr(), a.batch, 0.4, k, 0, idx.data_ptr(), sc.data_ptr(), cnt.data_ptr(),
                             st.data_ptr(), stream=stream)

    run(20)
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(a.steps)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / a.steps)
    print(json.dumps({"pdl": os.environ.get("FRP_PDL", "1"), "rows": a.rows, "dim": a.dim, "batch": a.batch,
                      "ms_per_step": [round(t, 5) for t in times], "path": frp.last_match_path(),
                      "status_max": int(st.max())}))


if __name__ == "__main__":
    main()
