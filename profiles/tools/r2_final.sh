#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): GPU test suite + smoke, bench.py (own arm with verified/configs, reference
# arm), the launch list of the same bench command, one ncu --set full capture of the dominant kernel's main pass at 100M, and
# of the streaming kernel.  Outputs -> gpurun_out/r2_final_*.
cd "$(dirname "$0")/../.."
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/r2_final_gpu.txt 2>&1
echo "#### pytest -m gpu" > $OUT/r2_final_tests.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 >> $OUT/r2_final_tests.txt 2>&1
tail -2 $OUT/r2_final_tests.txt
echo "#### smoke" >> $OUT/r2_final_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/r2_final_tests.txt 2>&1
tail -1 $OUT/r2_final_tests.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r2_final_bench_n1.json 2> $OUT/r2_final_bench_n1.err
cut -c1-300 $OUT/r2_final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2_final_bench_reference.json 2>> $OUT/r2_final_bench_n1.err
cut -c1-200 $OUT/r2_final_bench_reference.json
# launch list of the same bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_final_launches.csv \
  python bench.py --no-e2e --no-cpu --no-verify --no-configs --steps 2 --warmup 3 > $OUT/r2_final_launches.log 2>&1
python profiles/tools/launch_summary.py $OUT/r2_final_launches.csv > $OUT/r2_final_launch_summary.txt 2>&1; head -14 $OUT/r2_final_launch_summary.txt
# the dominant kernel: one main-pass launch (4 topk_tc launches per step: 3 sample passes + main), full set
timeout 900 ncu --set full --import-source on --clock-control none -k regex:topk_tc_kernel -s 15 -c 1 -f -o $OUT/r2_tc_main \
  python bench.py --no-e2e --no-cpu --no-verify --no-configs --steps 1 --warmup 3 > $OUT/r2_final_ncu.log 2>&1
tail -1 $OUT/r2_final_ncu.log | cut -c1-120
ncu -i $OUT/r2_tc_main.ncu-rep --page raw --csv > $OUT/r2_tc_main_raw.csv 2>/dev/null
