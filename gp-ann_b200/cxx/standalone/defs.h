// defs.h -- the part of gp-ann's src/defs.h the hot path touches (PointSet :10-21, NNVec :46, ClosestLeaders :50,
// partition helpers used by recall.h), for STANDALONE builds of the drop-in layer.  Inside a gp-ann checkout the
// reference's own defs.h is used instead; the types are layout-identical.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <utility>
#include <vector>

#include "topn.h"

struct PointSet {
    std::vector<float> coordinates;  // row-major n x d, may be empty
    size_t d = 0, n = 0;
    float* GetPoint(size_t i) { return &coordinates[i * d]; }
    void Drop() { std::vector<float>().swap(coordinates); }
    void Alloc() { coordinates.resize(n * d, 0.f); }
    void Resize(size_t new_n) {
        n = new_n;
        coordinates.resize(new_n * d);
    }
    bool empty() const { return n == 0; }
};

using Partition = std::vector<int>;
using NNVec = std::vector<std::pair<float, uint32_t>>;

inline int NumPartsInPartition(const Partition& partition) {
    return partition.empty() ? 0 : *std::max_element(partition.begin(), partition.end()) + 1;
}

TopN ClosestLeaders(PointSet& points, PointSet& leader_points, uint32_t my_id, int k);
