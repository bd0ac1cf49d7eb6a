"""Latency of the small-gallery path (BASELINE configs[0]: 1 query vs 1 000 x 128-d, L2, top-5).

    FRP_SMALL_CLUSTER=0|1 FRP_SMALL_POLL=0|1 python tools/prof_c1.py

Prints one JSON line: microseconds per Gallery.match call (ctypes + frp_match) for a few
(rows, batch) shapes and a checksum of the results (must not depend on the switches).
"""
import json
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "face-recognition-platform_b200"))
import frp_b200 as frp  # noqa: E402


def sweep():
    """4 calls per (rows, k) shape, for a launch list under ncu (kernel time vs rows and rounds)."""
    rng = np.random.default_rng(6)
    for rows in (64, 1000, 4096):
        G = (rng.standard_normal((rows, 128)) * 0.09).astype(np.float32)
        Q = (G[:1] + rng.standard_normal((1, 128)) * 0.02).astype(np.float32)
        gal = frp.Gallery(128, frp.METRIC_L2, device=0, capacity=rows)
        gal.append(G)
        for k, flags in ((1, 0), (10, frp.MATCH_NO_THRESHOLD)):
            for _ in range(4):
                gal.match(Q, 0.6, k, flags)
            print("shape rows=%d k=%d%s" % (rows, k, " nothreshold" if flags else ""))
        gal.close()


def main():
    if "--sweep" in sys.argv:
        return sweep()
    rng = np.random.default_rng(5)
    out = {"cluster": os.environ.get("FRP_SMALL_CLUSTER", "1"), "poll": os.environ.get("FRP_SMALL_POLL", "1")}
    crc = 0
    for rows, B, k in ((1000, 1, 5), (1000, 16, 5), (4096, 1, 10), (4096, 64, 10)):
        G = (rng.standard_normal((rows, 128)) * 0.09).astype(np.float32)
        Q = (G[rng.integers(0, rows, B)] + rng.standard_normal((B, 128)) * 0.02).astype(np.float32)
        gal = frp.Gallery(128, frp.METRIC_L2, device=0, capacity=rows)
        gal.append(G)
        for _ in range(16):
            r = gal.match(Q, 0.6, k)
        n = 512
        t0 = time.perf_counter()
        for _ in range(n):
            r = gal.match(Q, 0.6, k)
        dt = (time.perf_counter() - t0) / n
        assert frp.last_match_path() == "small_exact"
        for a in r:
            crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
        out["%dx128_B%d_k%d_us" % (rows, B, k)] = round(dt * 1e6, 2)
        gal.close()
    out["crc"] = crc
    print(json.dumps(out))


if __name__ == "__main__":
    main()
