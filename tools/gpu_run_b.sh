#!/bin/bash
# 2-GPU validation: remaining GPU tests (incl. the 2-rank NCCL test) + pipelined bench at N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_round2_gpu.py tests/test_shard_gpu.py -m gpu -q --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r2b_bench_n2.json
tail -5 gpurun_out/r2b_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-pipeline --no-extras > gpurun_out/r2b_bench_n2_nopipe.json 2> gpurun_out/r2b_bench_n2_nopipe.err
echo "bench n2 nopipe rc=$?"
head -c 600 gpurun_out/r2b_bench_n2_nopipe.json
