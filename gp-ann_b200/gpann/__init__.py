"""gpann -- ctypes binding of libgpann_cuda.so (include/gpann.h) plus a host-side mirror of the reference's
entry points for this path (ComputeGroundTruth, ComputeDistanceToKthNeighbor, KMeans, ObjectiveValue,
ClosestLeaders: reference src/recall.h:10-39, src/kmeans.cpp:228-249, src/defs.cpp:87-96).

The C++ drop-in headers (gp-ann_b200/cxx/) are what a gp-ann build links; this module is what the tests, the
bench and the torch.distributed multi-GPU driver use.  There is no CPU fallback: if the CUDA library or a
CUDA device is missing, `Context()` raises.
"""
from .lib import (ALGO_AUTO, ALGO_STREAM, ALGO_TENSOR, EMPTY_KEY, INVALID_ID, L2, MIPS, STORAGE_AUTO, STORAGE_F32,
                  STORAGE_S8, STORAGE_U8, Context, GpannError, Points, comm_unique_id, lib_path, load_library, pack_key,
                  unpack_keys)
from .api import (closest_leaders, compute_distance_to_kth_neighbor, compute_ground_truth, kmeans, kmeans_by_steps, objective_value,
                  remove_empty_clusters)

__all__ = [n for n in dir() if not n.startswith("_")]
