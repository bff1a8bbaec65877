#!/usr/bin/env python
"""bench_configs.py -- the OTHER BASELINE.json configurations (bench.py reports the headline one, configs[3]).

  python bench_configs.py deep    [--n 10000000 --nq 10000 --k 100]   configs[1]: DEEP-shape f32 10M x 96, exact top-100
  python bench_configs.py kmeans  [--n 100000000 --kc 1024]           configs[2]: k-means round, 100M x 96 f32 -> 1024
  python bench_configs.py c1      [--n 1000000]                        configs[0]: f32 1M x 128, top-10

One JSON line per run (device-resident inputs, CUDA-event timing of >= 3 steps after >= 2 warm-ups).  Not the
driver's contract -- evidence for DESIGN.md section 6 / profiles/.  Inputs are generated on the device.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gp-ann_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["deep", "kmeans", "c1"])
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--d", type=int, default=None)
    ap.add_argument("--kc", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--algo", default="auto", choices=["auto", "stream", "tensor"])
    ap.add_argument("--mixture", type=int, default=1, help="kmeans: 1 = mixture of kc Gaussians (SURVEY C3), 0 = N(0,1)")
    args = ap.parse_args()
    import torch
    import gpann
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = gpann.Context(0)
    ctx.set_algo({"auto": gpann.ALGO_AUTO, "stream": gpann.ALGO_STREAM, "tensor": gpann.ALGO_TENSOR}[args.algo])
    stream = torch.cuda.current_stream(dev).cuda_stream
    pk, pk_src = peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    if args.what in ("deep", "c1"):
        n = args.n or (10_000_000 if args.what == "deep" else 1_000_000)
        d = args.d or (96 if args.what == "deep" else 128)
        k = args.k or (100 if args.what == "deep" else 10)
        rows = gen_normal(torch, n, d, 3, dev)
        q = gen_normal(torch, args.nq, d, 4, dev)
        pts = ctx.from_device(rows.data_ptr(), n, d, gpann.STORAGE_F32, stream=stream, keepalive=rows)
        keys = torch.empty((args.nq, k), dtype=torch.int64, device=dev)
        ctx.set_timing(True)
        kms = []
        for _ in range(args.warmup):
            ctx.topk_device(pts, q.data_ptr(), args.nq, k, keys.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ctx.topk_device(pts, q.data_ptr(), args.nq, k, keys.data_ptr(), stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        ctx.topk_device(pts, q.data_ptr(), args.nq, k, keys.data_ptr(), stream=stream)
        kms = ctx.last_kernel_ms()
        flop = 2.0 * d * n * args.nq
        peak = pk["bf16_tflops_sustained"]
        line = {"config": f"f32 {n}x{d}, {args.nq} queries, exact top-{k}", "algo": ctx.last_algo, "ms_per_step": ms,
                "evals_per_s": n * args.nq / (ms * 1e-3), "queries_per_s": args.nq / (ms * 1e-3),
                "roofline": {"bound": "tensor", "achieved": flop / (kms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                             "frac": flop / (kms * 1e-3) / 1e12 / peak, "kernel_ms": kms,
                             "peak_source": f"bf16_tflops_sustained ({pk_src})"}}
        print(json.dumps(line), flush=True)
        return

    # ---- k-means: one Lloyd round = assignment + per-cluster sums (gpann_kmeans_partial_device) -----------------
    n = args.n or 100_000_000
    d = args.d or 96
    kc = args.kc
    g = torch.Generator(device=dev)
    g.manual_seed(55)
    centres = torch.randn((kc, d), dtype=torch.float32, device=dev, generator=g) * 4.0 if args.mixture else None
    rows = gen_normal(torch, n, d, 5, dev, centres=centres)
    pts = ctx.from_device(rows.data_ptr(), n, d, gpann.STORAGE_F32, stream=stream, keepalive=rows)
    pick = torch.randperm(n, device=dev, generator=g)[:kc]
    cent = rows[pick].clone()
    sums = torch.empty((kc, d), dtype=torch.float32, device=dev)
    counts = torch.empty((kc,), dtype=torch.int64, device=dev)
    assign = torch.empty((n,), dtype=torch.int32, device=dev)

    def round_():
        ctx.kmeans_partial_device(pts, cent.data_ptr(), kc, sums.data_ptr(), counts.data_ptr(), dev_assign=assign.data_ptr(),
                                  stream=stream)
    for _ in range(args.warmup):
        round_()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        round_()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flop = 2.0 * d * n * kc
    peak = pk["bf16_tflops_sustained"]
    line = {"config": f"k-means round f32 {n}x{d} -> {kc} centroids ({'mixture' if args.mixture else 'N(0,1)'})",
            "algo": ctx.last_algo, "ms_per_round": ms, "point_centroid_evals_per_s": n * kc / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flop / (ms * 1e-3) / 1e12 / peak, "peak_source": f"bf16_tflops_sustained ({pk_src})",
                         "note": "whole round (assignment + sums), not the assignment kernel alone"},
            "hbm_gbs": n * d * 4 * 2 / (ms * 1e-3) / 1e9, "ambiguous_points": ctx.last_ambiguous}
    ctx.set_timing(True)
    round_()
    line["assign_kernel_ms"] = ctx.last_kernel_ms()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
