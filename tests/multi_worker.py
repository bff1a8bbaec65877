"""Worker of tests/test_gpu_multi.py (launched with torchrun, one rank per GPU): the multi-process form of the sharded hot
path -- every rank owns a row shard behind its own gpann ctx, the ranks share ONE NCCL communicator inside libgpann_cuda
(gpann_comm_init_rank) -- must return on every rank exactly what one GPU returns for the whole set."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-ann_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import gpann
    from gpann.dist import shard_bounds
    from helpers import assert_topk_exact, assert_topk_close
    from oracle import build as obuild
    from oracle.oracle import Oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = gpann.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(gpann.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, uid.cpu().numpy().tobytes())
    assert ctx.comm_size == world
    obuild.build_oracle()
    oracle = Oracle()
    st = torch.cuda.current_stream(dev).cuda_stream

    # ---- exact top-k, byte data: bounds all-reduce(min) + key all-gather + merge inside gpann_topk_sharded_device ----
    rng = np.random.default_rng(31)
    n, nq, k = 1_200_007, 1500, 10
    P = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    Q = rng.integers(0, 256, (nq, 128)).astype(np.float32)
    lo, hi = shard_bounds(n, rank, world)
    pts = ctx.upload_bytes(P[lo:hi], id_base=lo)
    q = torch.from_numpy(Q).to(dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
    for _ in range(2):
        ctx.topk_sharded_device(pts, q.data_ptr(), nq, k, od.data_ptr(), oi.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert ctx.last_algo == "tensor_u8", ctx.last_algo
    pick = np.linspace(0, nq - 1, 40).astype(int)
    d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Q[pick], k)
    assert_topk_exact(od.cpu().numpy()[pick], oi.cpu().numpy().view(np.uint32)[pick], d0, i0)
    # every rank holds the same merged result
    mine = oi.to(torch.int64).sum()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert int(mine.item()) == int(ref.item())
    pts.free()

    # ---- float data, k = 100, one rank with an EMPTY shard (rows only on rank 0): collectives must still line up ----
    Pf = rng.standard_normal((30_000, 96)).astype(np.float32)
    Qf = rng.standard_normal((200, 96)).astype(np.float32)
    flo, fhi = (0, len(Pf)) if rank == 0 else (len(Pf), len(Pf))
    fpts = ctx.upload(Pf[flo:fhi], storage=gpann.STORAGE_F32, id_base=flo)
    qf = torch.from_numpy(Qf).to(dev)
    fd = torch.empty((200, 100), dtype=torch.float32, device=dev)
    fi = torch.empty((200, 100), dtype=torch.int32, device=dev)
    ctx.topk_sharded_device(fpts, qf.data_ptr(), 200, 100, fd.data_ptr(), fi.data_ptr(), stream=st)
    torch.cuda.synchronize()
    d0, i0, _ = oracle.ground_truth(Pf, Qf, 100)
    assert_topk_close(fd.cpu().numpy(), fi.cpu().numpy().view(np.uint32), d0, i0)
    fpts.free()

    # ---- k-means: sums + counts all-reduced every round; identical centroids on every rank, same as one GPU ----
    kn, kd, kc = 120_000, 96, 64
    centres = rng.standard_normal((kc, kd)).astype(np.float32) * 4
    X = (centres[rng.integers(0, kc, kn)] + rng.standard_normal((kn, kd))).astype(np.float32)
    C0 = np.concatenate([X[:kc - 4], 300 + rng.standard_normal((4, kd)).astype(np.float32)]).astype(np.float32)  # 4 clusters die
    klo, khi = shard_bounds(kn, rank, world)
    kpts = ctx.upload(X[klo:khi], storage=gpann.STORAGE_F32)
    cent = torch.from_numpy(C0.copy()).to(dev)
    alive = torch.ones(kc, dtype=torch.uint8, device=dev)
    counts = torch.zeros(kc, dtype=torch.int64, device=dev)
    assign = torch.empty(khi - klo, dtype=torch.int32, device=dev)
    ctx.kmeans_run_device(kpts, cent.data_ptr(), kc, 5, alive.data_ptr(), counts.data_ptr(), dev_assign=assign.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert int(counts.sum().item()) == kn
    with gpann.Context(local) as solo:  # the whole set on one GPU
        a1, c1, s1 = solo.kmeans_run(solo.upload(X, storage=gpann.STORAGE_F32), C0, rounds=5)
    live = alive.cpu().numpy().astype(bool)
    assert int(live.sum()) == len(c1) <= kc - 4  # the four far-away seeds (at least) lose every member
    np.testing.assert_allclose(cent.cpu().numpy()[live], c1, rtol=1e-5, atol=1e-5)
    assert np.array_equal(counts.cpu().numpy()[live].astype(np.uint64), s1)
    remap = np.cumsum(live) - 1
    assert np.array_equal(remap[assign.cpu().numpy()], a1[klo:khi])
    kpts.free()
    dist.barrier()
    if rank == 0:
        print("MULTI OK", world, "ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
