#!/bin/bash
# round-2 first validation: GPU parity tests, then the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
free -g > gpurun_out/r2a_free.txt; nproc >> gpurun_out/r2a_free.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
