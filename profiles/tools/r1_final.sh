#!/bin/bash
# Round-1 evidence run on ONE B200 (under gpurun): GPU test suite, bench.py (own arm + reference arm), launch list, one
# ncu --set full capture of the dominant kernel, microbenchmarks, and the other configs.  Outputs -> gpurun_out/.
OUT=gpurun_out; mkdir -p $OUT
# microbenchmarks (not part of the library): build on first use
mkdir -p profiles/tools/_bin
for b in mma_issue_bench tmem_ld_bench; do
  [ -x profiles/tools/_bin/$b ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/_bin/$b profiles/tools/$b.cu 2>/dev/null
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/final_gpu.txt 2>&1
echo "#### pytest -m gpu" > $OUT/final_tests.txt
timeout 900 python -m pytest tests -m gpu -x -q >> $OUT/final_tests.txt 2>&1
tail -2 $OUT/final_tests.txt
echo "#### smoke" >> $OUT/final_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/final_tests.txt 2>&1
tail -1 $OUT/final_tests.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/final_bench_n1.json 2> $OUT/final_bench_n1.err
cat $OUT/final_bench_n1.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/final_bench_reference.json 2>> $OUT/final_bench_n1.err
cat $OUT/final_bench_reference.json | cut -c1-300
# launch list of the same bench command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/final_launches.csv python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 > $OUT/final_launches.log 2>&1
python profiles/tools/launch_summary.py $OUT/final_launches.csv > $OUT/final_launch_summary.txt 2>&1; head -12 $OUT/final_launch_summary.txt
# the dominant kernel: one main-pass launch, full set
timeout 800 ncu --set full --import-source on --clock-control none -k regex:topk_tc_kernel -s 3 -c 1 -f -o $OUT/r1_tc_main_final python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > $OUT/final_ncu.log 2>&1
tail -1 $OUT/final_ncu.log | cut -c1-120
timeout 120 profiles/tools/_bin/mma_issue_bench > $OUT/final_mma_issue_bench.txt 2>&1
timeout 120 profiles/tools/_bin/tmem_ld_bench > $OUT/final_tmem_ld_bench.txt 2>&1
tail -4 $OUT/final_mma_issue_bench.txt
# the other configs (BASELINE.json configs[0..2]): parity-test cases, measured once per round
{ timeout 300 python bench_configs.py c1; timeout 300 python bench_configs.py deep; timeout 600 python bench_configs.py kmeans; } > $OUT/final_bench_configs.jsonl 2> $OUT/final_bench_configs.err
cat $OUT/final_bench_configs.jsonl | cut -c1-330
