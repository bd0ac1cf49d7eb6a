"""Alignment stage alone on the C4 workload (4 096 detections from 16 x 1080p frames): kernel time
from the library's own CUDA events, L2 flushed between runs.  `python tools/prof_align.py [tag]`
prints one JSON line; `--once` runs a single call (for ncu)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "face-recognition-platform_b200"))
sys.path.insert(0, ROOT)
import frp_b200 as frp
from frp_b200 import _capi
from bench import synth_detections


def main():
    tag = next((a for a in sys.argv[1:] if not a.startswith("--")), "")
    once = "--once" in sys.argv
    dev = torch.device("cuda", 0)
    F, H, W, per = 16, 1080, 1920, 256
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)).to(dev)
    fi, lm = synth_detections(F, per, H, W, seed=7)
    d_fi, d_lm = torch.from_numpy(fi).to(dev), torch.from_numpy(lm).to(dev)
    n = F * per
    blob = torch.empty((n, 3, 112, 112), dtype=torch.float16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _capi.lib.frp_profile_align_enable(1)
    ts = []
    for i in range(1 if once else 10):
        flush.fill_(i)
        frp.align.norm_crop_batch_device(frames, F, H, W, 3 * W, d_fi, d_lm, n, blob, stream=st)
        torch.cuda.synchronize()
        km = C.c_float()
        _capi.check(_capi.lib.frp_profile_align_last(C.byref(km)))
        if i >= 3 or once:
            ts.append(km.value)
    bytes_alg = F * H * W * 3 + n * (75264 + 40)
    t = float(np.mean(ts)) / 1e3
    print(json.dumps({"tag": tag, "kernel_ms": t * 1e3, "min_ms": float(np.min(ts)), "gbs": bytes_alg / t / 1e9,
                      "frac_of_6556.8": bytes_alg / t / 1e9 / 6556.8, "checksum": int(blob.view(torch.int16).long().sum().item())}))


if __name__ == "__main__":
    main()
