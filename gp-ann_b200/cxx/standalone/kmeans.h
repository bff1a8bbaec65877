// kmeans.h -- drop-in for gp-ann's src/kmeans.h:5-8 (same signatures).  RandomSample, KMeans and ObjectiveValue are
// implemented in kmeans_gpu.cpp over libgpann_cuda; BalancedKMeans (kmeans.cpp:253-524, SURVEY.md 8f "next") is
// declared for source compatibility and is still provided by the reference's own kmeans.cpp.
#pragma once

#include <cstddef>
#include <vector>

#include "defs.h"

PointSet RandomSample(PointSet& points, size_t num_samples, int seed);
std::vector<int> KMeans(PointSet& P, PointSet& centroids);
double ObjectiveValue(PointSet& points, PointSet& centroids, const std::vector<int>& closest_center);
std::vector<int> BalancedKMeans(PointSet& points, PointSet& centroids, size_t max_cluster_size);
