#!/usr/bin/env python
"""bench.py -- exact top-10 k-NN over the BigANN-shape workload (uint8 100M x 128, 10k queries), the configuration
BASELINE.json's metric is quoted on (configs[3]); it fits one B200 (12.8 GB), so N=1 runs it whole and N>1 shards
the SAME 100M rows across ranks (strong scaling), one process per GPU.

  python bench.py [--gpus N --steps K --warmup W]                 one JSON line (rank 0)
  torchrun ... bench.py --gpus N --steps K --warmup W             N ranks, one NCCL communicator INSIDE libgpann_cuda
  python bench.py --impl reference ...                            the reference's own CPU path (oracle/_ref)

A step = ONE call of the C ABI per rank, gpann_topk_sharded_device: the rank scans its shard for all queries (tcgen05 tile
kernel + per-query selection), the per-query bounds are all-reduced (min) over NCCL after the sample passes, the [nq,k] key
blocks are all-gathered and merged.  `value` = distance evaluations / s with the point shards resident in HBM; `e2e` = the
same through the host-buffer API (H2D of the shard's bytes + layout build + queries H2D + scan + merge + D2H of the result)
every step.  The 100M rows are generated from ONE seed by global row block, so every world size scans the same set; after
the timed region sampled queries of the result are checked bit-exactly against the unmodified reference (oracle/_ref) run
over all 100M rows on the host ("verified").  On one GPU the line also carries `configs`: BASELINE.json's other
single-GPU configurations (C1, C2, the C3 k-means round) and the streaming kernel, each with its own roofline.
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
GEN_BLOCK = 500_000   # rows per generator block: block b of the global row space is seeded with POINT_SEED + b
POINT_SEED = 6_000_000
QUERY_SEED = 7


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
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE main-pass launch from the committed ncu capture, if one was
    taken at this shard size (profiles/r2_traffic.json, else round 1's)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
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


def bind_to_gpu_numa(gpu_index):
    """Pin this process (and so its first-touch pinned host buffers) to the NUMA node of its GPU: torchrun does not, and a
    rank whose staging buffer sits on the other socket pulls every H2D byte across the inter-socket link (e2e only)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True,
                             text=True, timeout=20).stdout.strip().lower()
        dom, rest = bus.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    try:  # fall-back: the "CPU Affinity" column of `nvidia-smi topo -m`
        for line in subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines():
            f = line.split()
            if f and f[0] == f"GPU{gpu_index}":
                aff = next(x for x in f[1:] if x[0].isdigit() and ("-" in x or "," in x))
                cpus = set()
                for part in aff.split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    return {"cpu_affinity": aff, "cpus": len(cpus)}
    except Exception:
        pass
    return None


# ---- synthetic data: one seed for the whole set, generated by global row block ------------------------------------------
def gen_u8_rows(torch, lo, hi, d, dev):
    """Rows [lo,hi) of the global uint8 point set.  Block b = rows [b*GEN_BLOCK, (b+1)*GEN_BLOCK) comes from its own
    generator seeded POINT_SEED + b, so a rank generates exactly the rows any other world size would hold there."""
    rows = torch.empty((hi - lo, d), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    for b in range(lo // GEN_BLOCK, (max(hi, lo + 1) - 1) // GEN_BLOCK + 1):
        b0, b1 = b * GEN_BLOCK, (b + 1) * GEN_BLOCK
        g.manual_seed(POINT_SEED + b)
        blk = torch.randint(0, 256, (GEN_BLOCK, d), dtype=torch.uint8, device=dev, generator=g)
        s, e = max(lo, b0), min(hi, b1)
        if e > s:
            rows[s - lo:e - lo] = blk[s - b0:e - b0]
    return rows


def gen_u8_queries(torch, nq, d, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(QUERY_SEED)
    return torch.randint(0, 256, (nq, d), device=dev, generator=g).to(torch.float32)


def gen_normal(torch, n, d, seed, dev, chunk=4_000_000, centres=None, sigma=1.0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, d), dtype=torch.float32, device=dev, generator=g) * sigma
        if centres is not None:
            which = torch.randint(0, centres.shape[0], (e - s,), device=dev, generator=g)
            x += centres[which]
        out[s:e] = x
    return out


# ---- the reference on the host (oracle/_ref): baseline legs and the checker ----------------------------------------------
def load_ref():
    from oracle import build as obuild
    from oracle.oracle import Oracle, Ref, have_ref
    obuild.build_ref()
    if have_ref():
        return Ref(), "reference"
    obuild.build_oracle()
    return Oracle(), "port"


def cpu_reference_rate(seconds_target=15.0, n_points=1_000_000, nq0=256):
    """Time the UNMODIFIED reference's ComputeGroundTruth (oracle/_ref/libgpann_ref.so; parlay, all host cores) on a
    bounded sample of the same workload shape: n_points u8-valued float rows x 128, k=10.  Returns a dict."""
    import numpy as np
    ref, kind = load_ref()
    rng = np.random.default_rng(6)
    P = rng.integers(0, 256, (n_points, D), dtype=np.uint8).astype(np.float32)
    cores = os.cpu_count()
    if kind == "reference":
        cores = ref.num_workers()
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
        def run(nq):
            Q = rng.integers(0, 256, (nq, D), dtype=np.uint8).astype(np.float32)
            t = time.perf_counter()
            ref.ground_truth(P, Q, K)
            return time.perf_counter() - t
    dt = run(nq0)  # calibration pass (also warms the page cache / threads)
    rate = n_points * nq0 / dt
    nq = int(max(nq0, min(65536, seconds_target * rate / n_points)))
    return {"run": run, "rate0": rate, "nq": nq, "n_points": n_points, "kind": kind, "cores": cores}


def cpu_kmeans_rate(n_points=400_000, d=96, kc=1024):
    """The reference's KMeans (src/kmeans.cpp:228-244: 20 Lloyd rounds) on a bounded sample of config 3's shape."""
    import numpy as np
    ref, kind = load_ref()
    rng = np.random.default_rng(5)
    centres = (rng.standard_normal((kc, d)) * 2.0).astype(np.float32)
    P = (rng.standard_normal((n_points, d)).astype(np.float32) + centres[rng.integers(0, kc, n_points)])
    C0 = ref.random_sample(P, kc, 555) if kind == "reference" else ref.random_sample(P, kc, 555)[1]
    cores = ref.num_workers() if kind == "reference" else os.cpu_count()
    t = time.perf_counter()
    ref.kmeans(P, C0)
    dt = time.perf_counter() - t
    return {"value": n_points * kc * 20 / dt, "unit": "point*centroid evals/s", "cores": cores, "kind": kind,
            "sample": f"{n_points} x {d} f32 mixture of {kc} Gaussians -> {kc} centroids, KMeans (20 rounds) in {dt:.1f} s"}


def verify_against_reference(torch, dist, rows_u8, lo, queries, gpu_dist, gpu_ids, rank, world, n_check=16, chunk=2_500_000, k=K, exact=True,
                             n_total=None):
    """Bit-exact check of `n_check` sampled queries of the GPU result against the unmodified reference's ComputeGroundTruth
    over ALL rows: every rank runs the reference over its own shard in host chunks (float rows, as the reference's reader
    would hold them), the per-chunk top-k lists are merged lexicographically (what one scan over the union returns), rank 0
    gathers the shards' lists and compares (distance bits, ids) with the GPU's answer."""
    import numpy as np
    ref, kind = load_ref()
    nq = queries.shape[0]
    pick = np.linspace(0, nq - 1, n_check).astype(np.int64)
    Q = np.ascontiguousarray(queries[torch.as_tensor(pick, device=queries.device)].cpu().numpy(), dtype=np.float32)
    n_local = rows_u8.shape[0]
    t0 = time.perf_counter()
    cand_d, cand_i = [], []
    stage = torch.empty((min(chunk, max(n_local, 1)), rows_u8.shape[1]), dtype=torch.float32, pin_memory=True)
    for s in range(0, n_local, chunk):
        e = min(n_local, s + chunk)
        stage[:e - s].copy_(rows_u8[s:e].to(torch.float32))
        P = stage[:e - s].numpy()
        dd, ii, cc = ref.ground_truth(P, Q, k)
        cand_d.append(dd)
        cand_i.append(ii.astype(np.int64) + (lo + s))
    if cand_d:
        cd, ci = np.concatenate(cand_d, axis=1), np.concatenate(cand_i, axis=1)
    else:
        cd, ci = np.full((n_check, k), np.inf, np.float32), np.full((n_check, k), 2**32 - 1, np.int64)
    order = np.lexsort((ci, cd), axis=1)[:, :k]
    cd, ci = np.take_along_axis(cd, order, 1), np.take_along_axis(ci, order, 1)
    if world > 1:
        td = torch.from_numpy(cd.astype(np.float32)).to(gpu_dist.device)
        ti = torch.from_numpy(ci).to(gpu_dist.device)
        gd = [torch.empty_like(td) for _ in range(world)]
        gi = [torch.empty_like(ti) for _ in range(world)]
        dist.all_gather(gd, td)
        dist.all_gather(gi, ti)
        cd = torch.cat(gd, dim=1).cpu().numpy()
        ci = torch.cat(gi, dim=1).cpu().numpy()
        order = np.lexsort((ci, cd), axis=1)[:, :k]
        cd, ci = np.take_along_axis(cd, order, 1), np.take_along_axis(ci, order, 1)
    if rank != 0:
        return None
    g_d = gpu_dist[torch.as_tensor(pick, device=gpu_dist.device)].cpu().numpy()
    g_i = gpu_ids[torch.as_tensor(pick, device=gpu_ids.device)].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    if exact:
        ok = bool(np.array_equal(g_i, ci) and np.array_equal(g_d.view(np.uint32), cd.astype(np.float32).view(np.uint32)))
        compare = "bit-exact (float32 distance bits, uint32 ids)"
    else:  # float data: north_star's tolerance -- distances within 1e-5 relative, id sets equal except at ties inside it
        rel = np.abs(g_d.astype(np.float64) - cd) <= 1e-5 * np.abs(cd) + 1e-6
        same = g_i == ci
        tie = np.zeros_like(same)
        for q, r in np.argwhere(~same):
            tie[q, r] = np.isin(g_i[q, r], ci[q]) or abs(float(cd[q, r]) - float(cd[q, -1])) <= 2e-5 * abs(float(cd[q, r])) + 1e-6
        ok = bool(rel.all() and (same | tie).all())
        compare = "distances within 1e-5 relative; ids equal except swaps / boundary ties inside that tolerance"
    return {"ok": ok, "queries_checked": int(n_check), "query_ids": [int(x) for x in pick], "rows_scanned_per_query": int(n_total or N_TOTAL),
            "checker": f"oracle/_ref ComputeGroundTruth ({kind}) over all rows in {chunk}-row host chunks + lexicographic merge",
            "compare": compare, "seconds": round(time.perf_counter() - t0, 1), "mismatching_entries": int((g_i != ci).sum())}


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


# ---- the other BASELINE.json configurations, one GPU ----------------------------------------------------------------------
def time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def config_float_topk(torch, gpann, ctx, dev, name, n, d, nq, k, pk, pk_src, steps=3, warmup=3):
    stream = torch.cuda.current_stream(dev).cuda_stream
    rows = gen_normal(torch, n, d, 3 if name == "C2" else 1, dev)
    q = gen_normal(torch, nq, d, 4 if name == "C2" else 2, dev)
    pts = ctx.from_device(rows.data_ptr(), n, d, gpann.STORAGE_F32, stream=stream, keepalive=rows)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    launches0 = ctx.launch_count
    ms = time_steps(torch, lambda: ctx.topk_device(pts, q.data_ptr(), nq, k, keys.data_ptr(), stream=stream), steps, warmup)
    launches = (ctx.launch_count - launches0) // (steps + warmup)
    algo = ctx.last_algo
    ctx.set_timing(True)
    ctx.topk_device(pts, q.data_ptr(), nq, k, keys.data_ptr(), stream=stream)
    kms = ctx.last_kernel_ms()
    ctx.set_timing(False)
    pts.free()
    flops = 2.0 * d * n * nq
    ach = flops / (kms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    return {"workload": f"{name}: f32 N(0,1) {n}x{d}, {nq} queries, exact top-{k}", "ms_per_step": ms, "evals_per_s": n * nq / (ms * 1e-3),
            "algo": algo, "gpu_launches_per_step": int(launches), "dtype": "bf16 filter (tcgen05 kind::f16) + exact f32 re-rank",
            "l2": f"inputs ({n * d * 2 / 1e9:.2f} GB bf16 layout per pass) exceed the 126 MB L2" if n * d * 2 > 126e6 else "layout fits L2",
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "kernel": "topk_tc_kernel<true,false>", "kernel_ms": kms, "peak_source": f"bf16_tflops_sustained ({pk_src})",
                         "algorithmic_flops_per_launch": flops}}


def config_kmeans(torch, gpann, ctx, dev, n, d, kc, pk, pk_src, rounds=3, cpu=True):
    stream = torch.cuda.current_stream(dev).cuda_stream
    centres = gen_normal(torch, kc, d, 55, dev, sigma=2.0)  # centres N(0,4)
    rows = gen_normal(torch, n, d, 5, dev, centres=centres)
    pts = ctx.from_device(rows.data_ptr(), n, d, gpann.STORAGE_F32, stream=stream, keepalive=rows)
    g = torch.Generator(device=dev)
    g.manual_seed(555)
    cent0 = rows[torch.randperm(n, device=dev, generator=g)[:kc]].contiguous()
    cent = cent0.clone()
    alive = torch.ones(kc, dtype=torch.uint8, device=dev)
    counts = torch.zeros(kc, dtype=torch.int64, device=dev)
    launches0 = ctx.launch_count
    # a step = ONE Lloyd round (assignment + sums + update), centroids device-resident across rounds
    ms = time_steps(torch, lambda: ctx.kmeans_run_device(pts, cent.data_ptr(), kc, 1, alive.data_ptr(), counts.data_ptr(), stream=stream),
                    rounds, 2)
    launches = (ctx.launch_count - launches0) // (rounds + 2)
    algo = ctx.last_algo
    ctx.set_timing(True)
    ctx.kmeans_run_device(pts, cent.data_ptr(), kc, 1, alive.data_ptr(), counts.data_ptr(), stream=stream)
    kms = ctx.last_kernel_ms()
    ctx.set_timing(False)
    amb = ctx.last_ambiguous
    pts.free()
    flops = 2.0 * d * n * kc
    ach_k = flops / (kms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    out = {"workload": f"C3: f32 mixture of {kc} Gaussians {n}x{d} -> {kc} centroids, one Lloyd round (assignment + sums + update)",
           "ms_per_round": ms, "point_centroid_evals_per_s": n * kc / (ms * 1e-3), "algo": algo, "gpu_launches_per_round": int(launches),
           "dtype": "bf16x3 split precision (tcgen05 kind::f16) + exact f32 re-rank of ambiguous points",
           "exact_rerank_points": int(amb), "live_centroids": int(alive.sum().item()),
           "roofline": {"bound": "tensor", "achieved": ach_k, "peak": peak, "unit": "TFLOP/s", "frac": ach_k / peak, "traffic": None,
                        "kernel": "kmeans_tc_kernel", "kernel_ms": kms, "peak_source": f"bf16_tflops_sustained ({pk_src})",
                        "algorithmic_flops_per_launch": flops,
                        "note": "2*d flops per point-centroid pair (SURVEY 8d); the three split-precision MMA chains are not counted",
                        "whole_round_frac": flops / (ms * 1e-3) / 1e12 / peak,
                        "hbm_gbs_whole_round": (n * d * 4 * 2 + n * d * 4) / (ms * 1e-3) / 1e9}}
    if cpu:
        out["cpu_baseline"] = cpu_kmeans_rate()
    return out


def config_streaming(torch, gpann, ctx, dev, pk, pk_src):
    """The HBM-streaming scan (small query batches): achieved GB/s of point bytes against the measured copy bandwidth."""
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = []
    cases = [("u8", 100_000_000 if N_TOTAL >= 100_000_000 else N_TOTAL, 128), ("f32", 10_000_000, 96)]
    for kind, n, d in cases:
        if kind == "u8":
            rows = gen_u8_rows(torch, 0, n, d, dev)
            storage, esz = gpann.STORAGE_U8, 1
        else:
            rows = gen_normal(torch, n, d, 3, dev)
            storage, esz = gpann.STORAGE_F32, 4
        # stream algo forced: the layout build of byte storage is not what is measured here
        ctx.set_algo(gpann.ALGO_STREAM)
        pts = ctx.from_device(rows.data_ptr(), n, d, storage, stream=stream, keepalive=rows)
        for nq in (1, 4, 16):
            q = (torch.randint(0, 256, (nq, d), device=dev).to(torch.float32) if kind == "u8" else torch.randn((nq, d), device=dev))
            keys = torch.empty((nq, K), dtype=torch.int64, device=dev)
            ms = time_steps(torch, lambda: ctx.topk_device(pts, q.data_ptr(), nq, K, keys.data_ptr(), stream=stream), 3, 2)
            gbs = n * d * esz / (ms * 1e-3) / 1e9
            out.append({"storage": kind, "rows": n, "d": d, "nq": nq, "ms": ms, "hbm_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
                        "evals_per_s": n * nq / (ms * 1e-3)})
        pts.free()
        del rows
        ctx.set_algo(gpann.ALGO_AUTO)
        torch.cuda.empty_cache()
    return {"kernel": "topk_stream_kernel", "bound": "hbm", "peak_gbs": pk["hbm_gbs"], "peak_source": pk_src,
            "algorithmic_bytes": "rows * d * sizeof(storage) per pass (SURVEY 8d)", "cases": out}


def gen_f32_rows(torch, lo, hi, d, dev, seed0):
    """Rows [lo,hi) of a global float32 N(0,1) set, by generator block (see gen_u8_rows)."""
    rows = torch.empty((hi - lo, d), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for b in range(lo // GEN_BLOCK, (max(hi, lo + 1) - 1) // GEN_BLOCK + 1):
        b0, b1 = b * GEN_BLOCK, (b + 1) * GEN_BLOCK
        g.manual_seed(seed0 + b)
        blk = torch.randn((GEN_BLOCK, d), dtype=torch.float32, device=dev, generator=g)
        s, e = max(lo, b0), min(hi, b1)
        if e > s:
            rows[s - lo:e - lo] = blk[s - b0:e - b0]
    return rows


def run_c5(args, rank, world, local_rank):
    """BASELINE.json configs[4]: MS-Turing-shape float32 1B x 100, 10k queries, exact top-10 over the GPU shards (f32 rows for the
    exact re-rank + the bf16 filter layout: 83 GB per GPU at 8 GPUs).  Evidence run, not the driver's contract."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import gpann
    from gpann.dist import shard_bounds
    n_total, d, nq, k = int(os.environ.get("GPANN_C5_N", 1_000_000_000)), 100, 10_000, 10
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(n_total, rank, world)
    rows = gen_f32_rows(torch, lo, hi, d, dev, 8_000_000)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    queries = torch.randn((nq, d), dtype=torch.float32, device=dev, generator=g)
    ctx = gpann.Context(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(gpann.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.current_stream(dev).cuda_stream
    torch.cuda.synchronize()
    pts = ctx.from_device(rows.data_ptr(), hi - lo, d, gpann.STORAGE_F32, id_base=lo, stream=stream, keepalive=rows)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
    f = lambda: ctx.topk_sharded_device(pts, queries.data_ptr(), nq, k, od.data_ptr(), oi.data_ptr(), stream=stream)
    t_build = time.perf_counter()
    f()  # first call builds the bf16 layout
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    for _ in range(2):
        f()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ctx.set_timing(True)
    f()
    kms = ctx.last_kernel_ms()
    ctx.set_timing(False)
    verified = None
    if not args.no_verify:
        verified = verify_against_reference(torch, dist, rows, lo, queries, od, oi, rank, world, n_check=4, chunk=2_500_000, k=k, exact=False,
                                            n_total=n_total)
    if rank == 0:
        pk, pk_src = peaks()
        step_ms = float(ms.item())
        flops = 2.0 * d * (hi - lo) * nq
        ach = flops / (kms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": n_total * nq / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": step_ms,
                "dtype": "bf16 filter (tcgen05 kind::f16) + exact f32 re-rank", "data": "synthetic",
                "config": {"workload": f"C5: MS-Turing-shape f32 N(0,1) {n_total}x{d}, {nq} queries, exact top-{k} (BASELINE.json configs[4])",
                           "rows_per_gpu": hi - lo, "hbm_per_gpu_gb": round((hi - lo) * (d * 4 + 128 * 2 + 8) / 1e9, 1),
                           "storage": "f32 rows (exact re-rank) + norm-sorted bf16 layout [n,128] + norms + ids", "layout_build_first_call_s": round(t_build, 2),
                           "step": "gpann_topk_sharded_device"},
                "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                             "kernel": "topk_tc_kernel<true,false>", "kernel_ms": kms, "peak_source": f"bf16_tflops_sustained ({pk_src})",
                             "algorithmic_flops_per_launch": flops, "whole_step_frac": 2.0 * d * n_total * nq / world / (step_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]},
                "verified": verified}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--kmeans-multi", action="store_true", help="N>1: also time the sharded C3 Lloyd round (NCCL all-reduce)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"], help="c4 = the headline (default); c5 = 1B x 100 f32 evidence run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "c5":
        run_c5(args, rank, world, local_rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import gpann
    from gpann.dist import shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)

    # ---- synthetic workload: this rank's rows of the ONE global set, generated on the device ---------------------------
    lo, hi = shard_bounds(N_TOTAL, rank, world)
    n_local = hi - lo
    rows = gen_u8_rows(torch, lo, hi, D, dev)
    queries = gen_u8_queries(torch, NQ, D, dev)

    ctx = gpann.Context(local_rank)
    if world > 1:  # the library's own NCCL communicator; torch.distributed only ships the 128-byte id
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(gpann.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.current_stream(dev).cuda_stream
    torch.cuda.synchronize()
    pts = ctx.from_device(rows.data_ptr(), n_local, D, gpann.STORAGE_U8, id_base=lo, stream=stream, keepalive=rows)
    out_dist = torch.empty((NQ, K), dtype=torch.float32, device=dev)
    out_ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)

    def step_fn():  # ONE C-ABI call: local scan, bounds all-reduce(min), key all-gather, merge
        ctx.topk_sharded_device(pts, queries.data_ptr(), NQ, K, out_dist.data_ptr(), out_ids.data_ptr(), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks ---------------------------
    for _ in range(warmup):
        step_fn()
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
        step_fn()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    # per-launch duration of the dominant kernel: separate passes so reading the bracket never stalls the timed loop
    for _ in range(min(3, args.steps)):
        step_fn()
        kernel_ms.append(ctx.last_kernel_ms())
    ctx.set_timing(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    evals = float(N_TOTAL) * NQ * args.steps
    value = evals / (ms * 1e-3)
    res_fingerprint = int(out_ids.to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF  # same at every world size (one seed)

    verified = None
    if not args.no_verify:
        verified = verify_against_reference(torch, dist, rows, lo, queries, out_dist, out_ids, rank, world)
        if rank == 0:
            verified["result_fingerprint"] = res_fingerprint

    # ---- optional: the sharded k-means round (config 3 at N GPUs), NCCL all-reduce of sums + counts per round ------------
    kmeans_multi = None
    if args.kmeans_multi:
        pts.free()
        del rows
        torch.cuda.empty_cache()
        kn, kd, kc = int(os.environ.get("GPANN_BENCH_KM_N", 100_000_000)), 96, 1024
        klo, khi = shard_bounds(kn, rank, world)
        centres = gen_normal(torch, kc, kd, 55, dev, sigma=2.0)
        krows = gen_normal(torch, khi - klo, kd, 5000 + rank, dev, centres=centres)
        kpts = ctx.from_device(krows.data_ptr(), khi - klo, kd, gpann.STORAGE_F32, id_base=klo, stream=stream, keepalive=krows)
        cent = centres.clone() + 0.5  # identical start on every rank
        alive = torch.ones(kc, dtype=torch.uint8, device=dev)
        counts = torch.zeros(kc, dtype=torch.int64, device=dev)
        f = lambda: ctx.kmeans_run_device(kpts, cent.data_ptr(), kc, 1, alive.data_ptr(), counts.data_ptr(), stream=stream)
        for _ in range(2):
            f()
        barrier()
        e0.record()
        for _ in range(3):
            f()
        e1.record()
        barrier()
        kms_t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(kms_t, op=dist.ReduceOp.MAX)
        csum = float(cent.double().sum().item())
        kmeans_multi = {"workload": f"C3: f32 {kn}x{kd} -> {kc} centroids sharded over {world} GPUs, one Lloyd round", "ms_per_round": float(kms_t.item()),
                        "point_centroid_evals_per_s": kn * kc / (float(kms_t.item()) * 1e-3), "centroid_checksum": csum,
                        "total_count": int(counts.sum().item()), "exchange": "ncclAllReduce(sum) of [kc,d] f32 sums + [kc] u64 counts per round"}
        kpts.free()
        del krows
        torch.cuda.empty_cache()
        rows = gen_u8_rows(torch, lo, hi, D, dev)
        pts = ctx.from_device(rows.data_ptr(), n_local, D, gpann.STORAGE_U8, id_base=lo, stream=stream, keepalive=rows)

    # ---- e2e: host buffers in, host result out, every step -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_rows = torch.empty((n_local, D), dtype=torch.uint8, pin_memory=True)
        host_rows.copy_(rows)
        host_q = torch.empty((NQ, D), dtype=torch.float32, pin_memory=True)
        host_q.copy_(queries)
        torch.cuda.synchronize()
        pts.free()
        del rows
        torch.cuda.empty_cache()
        host_ptr = host_rows.data_ptr()
        gathered = torch.empty((world, NQ, K), dtype=torch.int64, device=dev)
        local_keys = torch.empty((NQ, K), dtype=torch.int64, device=dev)
        e_dist = torch.empty((NQ, K), dtype=torch.float32, device=dev)
        e_ids = torch.empty((NQ, K), dtype=torch.int32, device=dev)

        def e2e_step():
            q = host_q.to(dev, non_blocking=True)                              # H2D of the queries
            keys = local_keys if world > 1 else gathered[0]
            ctx.topk_stream_bytes_device(host_ptr, n_local, D, 0, q.data_ptr(), NQ, K, keys.data_ptr(), id_base=lo, stream=stream)
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(-1), keys.view(-1))
            ctx.merge_keys_device(gathered.data_ptr(), world, NQ, K, e_dist.data_ptr(), e_ids.data_ptr(), stream=stream)
            return e_dist.cpu(), e_ids.cpu()                                    # D2H of the result

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
        same = bool(torch.equal(res[1].to(dev), out_ids))
        e2e = {"value": float(N_TOTAL) * NQ * e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * D + NQ * D * 4), "d2h_bytes_per_step": int(NQ * K * 8),
               "steps": e2e_steps, "ms_per_step": 1e3 * float(dt.item()) / e2e_steps, "same_ids_as_resident_path": same, "numa_binding": numa,
               "note": "per rank: pinned host u8 shard -> gpann_topk_stream_bytes_device (chunked H2D overlapped with layout "
                       "build + tcgen05 scan) -> all-gather -> gpann_merge_keys_device -> D2H"}
        del host_rows
    else:
        pts.free()
        del rows
    torch.cuda.empty_cache()

    if rank == 0:
        pk, pk_src = peaks()
        kms = statistics.mean(kernel_ms)
        ops_per_launch = 2.0 * D * n_local * NQ  # 2*d ops per distance evaluation
        achieved = ops_per_launch / (kms * 1e-3) / 1e12
        peak = I8_TOPS_SUSTAINED
        peak_bf16x2 = 2.0 * pk["bf16_tflops_sustained"]
        whole_step = 2.0 * D * N_TOTAL * NQ / world / (ms / args.steps * 1e-3) / 1e12  # per GPU, over the whole step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": n_local, "parallelism": f"point-shard x{world}",
                       "seed": f"one seed for the whole set: block b of {GEN_BLOCK} rows <- torch generator {POINT_SEED}+b; queries <- {QUERY_SEED}",
                       "step": "gpann_topk_sharded_device: scan + ncclAllReduce(min) of bounds + ncclAllGather of keys + merge",
                       "l2": "inputs (12.8 GB of points per pass) exceed the 126 MB L2; no flush needed"},
            "queries_per_s": value / N_TOTAL,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                         "traffic": measured_traffic(n_local), "kernel": "topk_tc_kernel<false,false>", "kernel_ms": kms,
                         "peak_source": "own measurement: sustained tcgen05.mma.kind::i8 issue-only rate, all SMs, ~150 ms under the "
                                        "power cap (profiles/tools/mma_issue_bench.cu); MEASURED_PEAKS.json has no int8 figure",
                         "frac_of_2x_bf16_sustained": achieved / peak_bf16x2,
                         "peak_2x_bf16_sustained": peak_bf16x2, "bf16_peak_source": pk_src,
                         "whole_step_achieved_per_gpu": whole_step, "whole_step_frac": whole_step / peak,
                         "algorithmic_ops_per_launch": ops_per_launch,
                         "algorithmic_bytes_per_launch": n_local * (D + 8),
                         "hbm_gbs_achieved": n_local * (D + 8) / (kms * 1e-3) / 1e9},
            "clocks": clocks, "gpu_launches": int(launches * world),
        }
        if verified is not None:
            line["verified"] = verified
        if e2e:
            line["e2e"] = e2e
        if kmeans_multi:
            line["kmeans_sharded"] = kmeans_multi
        if world == 1 and not args.no_configs:
            cfg = {}
            try:
                cfg["C1"] = config_float_topk(torch, gpann, ctx, dev, "C1", 1_000_000, 128, 10_000, 10, pk, pk_src)
                torch.cuda.empty_cache()
                cfg["C2"] = config_float_topk(torch, gpann, ctx, dev, "C2", 10_000_000, 96, 10_000, 100, pk, pk_src)
                torch.cuda.empty_cache()
                cfg["C3"] = config_kmeans(torch, gpann, ctx, dev, int(os.environ.get("GPANN_BENCH_KM_N", 100_000_000)), 96, 1024, pk, pk_src,
                                          cpu=not args.no_cpu)
                torch.cuda.empty_cache()
                cfg["streaming"] = config_streaming(torch, gpann, ctx, dev, pk, pk_src)
            except Exception as ex:  # a failing side configuration must not lose the headline line
                cfg["error"] = repr(ex)
            line["configs"] = cfg
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
