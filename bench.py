#!/usr/bin/env python
"""bench.py -- exact top-10 k-NN over the BigANN-shape workload (uint8 100M x 128, 10k queries), the configuration
BASELINE.json's metric is quoted on (configs[3]); it fits one B200 (12.8 GB), so N=1 runs it whole and N>1 shards
the SAME 100M rows across ranks (strong scaling), one process per GPU.

  python bench.py [--gpus N --steps K --warmup W]                 one JSON line (rank 0)
  torchrun ... bench.py --gpus N --steps K --warmup W             N ranks over NCCL
  python bench.py --impl reference ...                            the reference's own CPU path (oracle/_ref)

A step = one pass of the hot path over the whole batch: every rank scans its shard for all queries
(gpann_topk_device: tcgen05 tile kernel + per-query selection), one all-gather of the [nq,k] key blocks over
NCCL, and the merge (gpann_merge_keys_device).  `value` = distance evaluations / s with the point shards resident
in HBM; `e2e` = the same through the host-buffer API (H2D of the shard's bytes + layout build + queries H2D +
scan + merge + D2H of the result) every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gp-ann_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_TOTAL = int(os.environ.get("GPANN_BENCH_N", 100_000_000))
D = int(os.environ.get("GPANN_BENCH_D", 128))  # probes only; the workload is d = 128
NQ = int(os.environ.get("GPANN_BENCH_NQ", 10_000))
K = 10
METRIC = "exact_top10_knn_distance_evals_per_s"
UNIT = "evals/s"
WORKLOAD = f"BigANN-shape uint8 {N_TOTAL}x{D}, {NQ} queries, exact top-{K} (BASELINE.json configs[3])"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# MEASURED_PEAKS.json has no int8 figure.  The denominator of the tensor roofline is this repo's own measurement of the
# tensor pipe on this pool's B200s: profiles/tools/mma_issue_bench.cu issues back-to-back tcgen05.mma.kind::i8
# (M128 N256 K32, random operands, no loads, no epilogue) on all 148 SMs for ~150 ms and sustains 4.18-4.21 POP/s under
# the 1000 W cap (4.56 POP/s for a 5 ms burst = the nominal 8192 MAC/clk/SM); profiles/r1_mma_issue_bench.txt.
I8_TOPS_SUSTAINED = 4200.0


def measured_traffic(n_rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE main-pass launch from the committed ncu capture
    (profiles/r1_traffic.json), if it was taken at this shard size."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        t = json.load(open(path))
        if int(t["rows"]) == int(n_rows):
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(seconds_target=15.0, n_points=1_000_000, nq0=256):
    """Time the UNMODIFIED reference's ComputeGroundTruth (oracle/_ref/libgpann_ref.so; parlay, all host cores) on a
    bounded sample of the same workload shape: n_points u8-valued float rows x 128, k=10.  Returns a dict."""
    import numpy as np
    from oracle import build as obuild
    from oracle.oracle import Oracle, Ref, have_ref
    obuild.build_ref()
    rng = np.random.default_rng(6)
    P = rng.integers(0, 256, (n_points, D), dtype=np.uint8).astype(np.float32)
    cores = os.cpu_count()
    if have_ref():
        ref = Ref()
        kind, cores = "reference", ref.num_workers()
        hp = ref.lib.ref_pointset_create(P, n_points, D)

        def run(nq):
            Q = rng.integers(0, 256, (nq, D), dtype=np.uint8).astype(np.float32)
            hq = ref.lib.ref_pointset_create(Q, nq, D)
            t = time.perf_counter()
            ref.lib.ref_ground_truth_handles(hp, hq, K, None)
            dt = time.perf_counter() - t
            ref.lib.ref_pointset_free(hq)
            return dt
    else:
        obuild.build_oracle()
        o = Oracle()
        kind = "port"

        def run(nq):
            Q = rng.integers(0, 256, (nq, D), dtype=np.uint8).astype(np.float32)
            t = time.perf_counter()
            o.ground_truth(P, Q, K)
            return time.perf_counter() - t
    dt = run(nq0)  # calibration pass (also warms the page cache / threads)
    rate = n_points * nq0 / dt
    nq = int(max(nq0, min(65536, seconds_target * rate / n_points)))
    return {"run": run, "rate0": rate, "nq": nq, "n_points": n_points, "kind": kind, "cores": cores}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cal = cpu_reference_rate(seconds_target=max(4.0, 40.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cal["run"](cal["nq"])
    t = time.perf_counter()
    for _ in range(args.steps):
        cal["run"](cal["nq"])
    dt = time.perf_counter() - t
    evals = cal["n_points"] * cal["nq"] * args.steps
    value = evals / dt
    sample = f"{cal['n_points']} u8-valued float points x {D}, {cal['nq']} queries per step, k={K}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cal["cores"], "kind": cal["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "queries_per_s": value / N_TOTAL, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import gpann
    from gpann.dist import CudaShard, ShardedTopK, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)

    # ---- synthetic workload: this rank's shard, generated on the device --------------------------------------
    lo, hi = shard_bounds(N_TOTAL, rank, world)
    n_local = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(6_000 + rank)
    rows = torch.empty((n_local, D), dtype=torch.uint8, device=dev)
    chunk = 4_000_000
    for s in range(0, n_local, chunk):
        e = min(n_local, s + chunk)
        rows[s:e] = torch.randint(0, 256, (e - s, D), dtype=torch.uint8, device=dev, generator=gen)
    qgen = torch.Generator(device=dev)
    qgen.manual_seed(7)
    queries = torch.randint(0, 256, (NQ, D), device=dev, generator=qgen).to(torch.float32)

    ctx = gpann.Context(local_rank)
    stream = torch.cuda.current_stream(dev).cuda_stream
    torch.cuda.synchronize()
    pts = ctx.from_device(rows.data_ptr(), n_local, D, gpann.STORAGE_U8, id_base=lo, stream=stream, keepalive=rows)
    shard = CudaShard(ctx, pts, dev)
    step_fn = ShardedTopK(shard.local_topk, shard.merge)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks ---------------------------
    for _ in range(warmup):
        out = step_fn(queries, K)
    assert ctx.last_algo == "tensor_u8", f"expected the tcgen05 kernel, got {ctx.last_algo}"
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ctx.launch_count
    ctx.set_timing(True)
    kernel_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step_fn(queries, K)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    # per-launch duration of the dominant kernel: separate passes so reading the bracket never stalls the timed loop
    for _ in range(min(3, args.steps)):
        step_fn(queries, K)
        kernel_ms.append(ctx.last_kernel_ms())
    ctx.set_timing(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    evals = float(N_TOTAL) * NQ * args.steps
    value = evals / (ms * 1e-3)

    # ---- e2e: host buffers in, host result out, every step -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_rows = torch.empty((n_local, D), dtype=torch.uint8, pin_memory=True)
        host_rows.copy_(rows)
        host_q = torch.empty((NQ, D), dtype=torch.float32, pin_memory=True)
        host_q.copy_(queries)
        torch.cuda.synchronize()
        pts.free()
        del rows, shard, step_fn
        torch.cuda.empty_cache()
        host_np = host_rows.numpy()

        host_ptr = host_rows.data_ptr()
        merge_only = CudaShard(ctx, None, dev)

        def local_stream(q, k):
            keys = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
            ctx.topk_stream_bytes_device(host_ptr, n_local, D, 0, q.data_ptr(), q.shape[0], k, keys.data_ptr(), id_base=lo,
                                         stream=torch.cuda.current_stream(dev).cuda_stream)
            return keys

        def e2e_step():
            q = host_q.to(dev, non_blocking=True)                              # H2D of the queries
            d_out, i_out = ShardedTopK(local_stream, merge_only.merge)(q, K)   # streamed H2D + scan, all-gather, merge
            return d_out.cpu(), i_out.cpu()                                    # D2H of the result

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            res = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(N_TOTAL) * NQ * e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * D + NQ * D * 4), "d2h_bytes_per_step": int(NQ * K * 8),
               "steps": e2e_steps, "ms_per_step": 1e3 * float(dt.item()) / e2e_steps,
               "note": "per rank: pinned host u8 shard -> gpann_topk_stream_bytes_device (chunked H2D overlapped with layout "
                       "build + tcgen05 scan) -> all-gather -> gpann_merge_keys_device -> D2H"}

    if rank == 0:
        pk, pk_src = peaks()
        kms = statistics.mean(kernel_ms)
        ops_per_launch = 2.0 * D * n_local * NQ  # 2*d ops per distance evaluation
        achieved = ops_per_launch / (kms * 1e-3) / 1e12
        peak = I8_TOPS_SUSTAINED
        peak_bf16x2 = 2.0 * pk["bf16_tflops_sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": n_local, "parallelism": f"point-shard x{world}",
                       "l2": "inputs (12.8 GB of points per pass) exceed the 126 MB L2; no flush needed"},
            "queries_per_s": value / N_TOTAL,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                         "traffic": measured_traffic(n_local), "kernel": "topk_tc_kernel<false,false>", "kernel_ms": kms,
                         "peak_source": "own measurement: sustained tcgen05.mma.kind::i8 issue-only rate, all SMs, ~150 ms under the "
                                        "power cap (profiles/tools/mma_issue_bench.cu); MEASURED_PEAKS.json has no int8 figure",
                         "frac_of_2x_bf16_sustained": achieved / peak_bf16x2,
                         "peak_2x_bf16_sustained": peak_bf16x2, "bf16_peak_source": pk_src,
                         "algorithmic_ops_per_launch": ops_per_launch,
                         "algorithmic_bytes_per_launch": n_local * (D + 8),
                         "hbm_gbs_achieved": n_local * (D + 8) / (kms * 1e-3) / 1e9},
            "clocks": clocks, "gpu_launches": int(launches * world),
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            cal = cpu_reference_rate(seconds_target=12.0)
            dt = cal["run"](cal["nq"])
            line["cpu_baseline"] = {"value": cal["n_points"] * cal["nq"] / dt, "unit": UNIT, "cores": cal["cores"],
                                    "kind": cal["kind"],
                                    "sample": f"{cal['n_points']} u8-valued float points x {D}, {cal['nq']} queries, k={K}, "
                                              f"ComputeGroundTruth in {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
