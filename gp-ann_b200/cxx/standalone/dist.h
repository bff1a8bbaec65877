// dist.h -- drop-in for gp-ann's src/dist.h:3-15 (same names, same signatures).
// These are the SCALAR pair functions: the reference calls them billions of times from inside parlay worker
// threads (recall.h:104, partitioning.cpp:128, kmeans_tree_router.cpp:117 ...), so they stay lock-free host code
// (dist.cpp).  The batched loops that dominate the run time go through recall.h / kmeans.h to the GPU instead.
#pragma once

float sqr_l2_dist(const float* a, const float* b, unsigned size);
float inner_product(float* p, float* q, unsigned d);
float mips_distance(float* p, float* q, unsigned d);
float vec_norm(float* p, unsigned d);
bool L2Normalize(float* p, unsigned d);
float distance(float* p, float* q, unsigned d);
float pos_distance(float* p, float* q, unsigned d);
