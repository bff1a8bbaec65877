cd /root/repo
GPANN_TC_STATS=2 GPANN_BENCH_N=12500000 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-verify 2>&1 | grep "gpann stats" | tail -5
