// defs.cpp (standalone builds only) -- ClosestLeaders of gp-ann's src/defs.cpp:87-96.  Inside a gp-ann checkout the
// reference's own defs.cpp provides it.  One query against <= ~1500 leaders is a handful of microseconds of host work and
// it is called per point from inside worker threads (knn_graph.h:75-81), so the single-query form stays lock-free host
// code; the batched form (all points against the leaders at once) is gpann_topk on the leader set.
#include "defs.h"
#include "dist.h"

TopN ClosestLeaders(PointSet& points, PointSet& leader_points, uint32_t my_id, int k) {
    TopN best(k);
    float* me = points.GetPoint(my_id);
    for (uint32_t j = 0; j < leader_points.n; ++j) best.Add({ distance(leader_points.GetPoint(j), me, points.d), j });
    return best;
}
