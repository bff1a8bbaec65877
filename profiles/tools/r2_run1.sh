#!/bin/bash
# Round-2 GPU pass 1 (one B200): the whole GPU test suite, then the bench line with configs + verification.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt; free -g | head -2 >> gpurun_out/r2_gpu.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests.txt
tail -5 gpurun_out/r2_tests.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench exit $?"
tail -c 1500 gpurun_out/r2_bench_n1.err
head -c 3000 gpurun_out/r2_bench_n1.json
