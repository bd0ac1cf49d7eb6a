"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with
the keys the driver reads, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                           "--steps", "1", "--warmup", "0", "--rows", "1000000"],
                          capture_output=True, text=True, timeout=240, env=env)


def test_reference_arm_prints_the_contract_line():
    # torchrun exports OMP_NUM_THREADS=1: the arm must still use every host core
    r = _run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("1:N match queries/s")
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("cosine 1:N") and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "ALL 1000000 rows" in cb["sample"]
    assert cb["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_generates_the_surveyed_workloads_itself():
    """bench.py builds its own synthetic inputs (oracle/ is reserved for the checker and the
    cpu_baseline legs); they are the SURVEY 8d workloads, identical to the oracle's generators."""
    import importlib.util
    import frp_oracle_np as o
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a, b = bench.synth_detections(4, 32, 1080, 1920, seed=7), o.synth_detections(4, 32, 1080, 1920, seed=7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    G = bench.synth_l2_gallery(1000, 128, seed=1234)
    assert np.array_equal(G, o.synth_l2_gallery(1000, 128, seed=1234))
    q1, q2 = bench.synth_l2_queries(G, 64, seed=4321), o.synth_l2_queries(G, 64, seed=4321)
    assert np.array_equal(q1[0], q2[0]) and np.array_equal(q1[1], q2[1])
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src.split("def align_sharded", 1)[1].split("\ndef ", 1)[0]
    assert "oracle" not in body
