"""Multi-GPU form of the hot path: one process per GPU (torch.distributed), the POINT SET sharded by rows.

Top-k (reference ComputeGroundTruth, src/recall.h:25-39): rank r owns rows [lo_r, hi_r) and reports global ids
(id_base = lo_r), computes its local k best keys per query, the [nq,k] uint64 key blocks are exchanged with ONE
all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) and every rank merges world*k keys per query.  Keys
order like std::pair<float,uint32_t>, so the merged result -- ties included -- equals the single-scan result.

k-means (reference KMeans, src/kmeans.cpp:228-244): centroids replicated, every rank assigns its rows and sums
them per cluster, ONE all-reduce(sum) of [kc,d] float sums + [kc] counts per Lloyd round, identical division and
empty-cluster removal on every rank.

The exchange is 8*nq*k bytes per rank (0.8 MB at nq=10k, k=10) respectively 4*kc*d+8*kc bytes (0.4 MB), i.e.
latency-bound and <1% of a step, so it stays a library collective rather than a fused peer-memory kernel
("multi-GPU").
"""

def shard_bounds(n, rank, world):
    """Contiguous row range of `rank`: [rank*ceil(n/world), min(n, (rank+1)*ceil(n/world)))  (SURVEY.md 8e)."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class ShardedTopK:
    """Per-rank driver.  `local_topk(queries, k) -> uint64 keys [nq,k]` and `merge(all_keys [world,nq,k], k) ->
    (dist, ids)` are injected so the collective plumbing is identical for the CUDA path (torch cuda tensors +
    gpann_topk_device / gpann_merge_keys_device) and for the gloo CPU tests."""

    def __init__(self, local_topk, merge, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local_topk = local_topk
        self.merge = merge

    def __call__(self, queries, k):
        import torch
        keys = self.local_topk(queries, k)  # torch int64 tensor viewing uint64 keys, [nq,k]
        if self.world == 1:
            gathered = keys.unsqueeze(0)
        else:
            gathered = torch.empty((self.world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
            # list-of-views form: accepted by both NCCL and gloo; writes land in `gathered`
            self.dist.all_gather(list(gathered.unbind(0)), keys.contiguous(), group=self.group)
        return self.merge(gathered, k)


def allreduce_kmeans_partials(sums, counts, group=None):
    """sum-all-reduce of the per-shard cluster sums [kc,d] (float32) and counts [kc] (int64 view of uint64)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return sums, counts


class CudaShard:
    """The CUDA realisation of one rank: a gpann Context + this rank's rows, torch tensors for the buffers."""

    def __init__(self, ctx, pts, device):
        import torch
        self.torch = torch
        self.ctx, self.pts, self.device = ctx, pts, device

    def local_topk(self, queries, k):
        torch = self.torch
        q = queries.contiguous()
        keys = torch.empty((q.shape[0], k), dtype=torch.int64, device=self.device)
        self.ctx.topk_device(self.pts, q.data_ptr(), q.shape[0], k, keys.data_ptr(),
                             stream=torch.cuda.current_stream(self.device).cuda_stream)
        return keys

    def merge(self, gathered, k):
        torch = self.torch
        world, nq = gathered.shape[0], gathered.shape[1]
        dist_out = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        id_out = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        self.ctx.merge_keys_device(gathered.data_ptr(), world, nq, k, dist_out.data_ptr(), id_out.data_ptr(),
                                   stream=torch.cuda.current_stream(self.device).cuda_stream)
        return dist_out, id_out

    def kmeans_round(self, centroids, group=None):
        """One sharded Lloyd round on device tensors; returns (new centroids incl. zero rows, counts)."""
        torch = self.torch
        kc, d = centroids.shape
        sums = torch.empty((kc, d), dtype=torch.float32, device=self.device)
        counts = torch.empty((kc,), dtype=torch.int64, device=self.device)
        self.ctx.kmeans_partial_device(self.pts, centroids.data_ptr(), kc, sums.data_ptr(), counts.data_ptr(),
                                       stream=torch.cuda.current_stream(self.device).cuda_stream)
        allreduce_kmeans_partials(sums, counts, group)
        nz = counts > 0
        sums[nz] = sums[nz] / counts[nz].to(torch.float32).unsqueeze(1)
        return sums, counts
