cd /root/repo
for lib in libgpann_cuda.so libgpann_cuda_nopipe.so; do
echo "== $lib"
GPANN_LIB=$PWD/gp-ann_b200/$lib GPANN_TC_STATS=1 GPANN_BENCH_N=12500000 timeout 300 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify 2>&1 | grep -E "gpann stats|ms_per_step" | tail -4 | cut -c1-700
done
