// dist.cpp -- host scalar distances behind dist.h (drop-in for gp-ann's src/dist.cpp).
// Same arithmetic as the reference's release build: squared L2 accumulates 8 interleaved fp32 lanes with FMA over
// the leading multiple of 16 elements (what hnswlib's 256-bit kernel does, src/dist.cpp:11-16), adds the lanes in
// index order and finishes the remainder sequentially.  Metric selection stays a compile-time switch
// (-DMIPS_DISTANCE), as in the reference (src/dist.cpp:75-81).
#include "dist.h"

#include <cmath>
#include <iostream>

float sqr_l2_dist(const float* a, const float* b, unsigned size) {
    float acc[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    const unsigned body = size & ~15u;
    for (unsigned i = 0; i < body; i += 8)
        for (unsigned l = 0; l < 8; ++l) {
            const float diff = a[i + l] - b[i + l];
            acc[l] = std::fmaf(diff, diff, acc[l]);
        }
    float total = acc[0];
    for (unsigned l = 1; l < 8; ++l) total += acc[l];
    float rest = 0.f;
    for (unsigned i = body; i < size; ++i) {
        const float diff = a[i] - b[i];
        rest = std::fmaf(diff, diff, rest);
    }
    return total + rest;
}

float inner_product(float* p, float* q, unsigned d) {
    float dot = 0.f;
    for (unsigned i = 0; i < d; ++i) dot = std::fmaf(q[i], p[i], dot);
    return dot;
}

float mips_distance(float* p, float* q, unsigned d) { return 1.0f - inner_product(p, q, d); }

float vec_norm(float* p, unsigned d) {
    float s = 0.f;
    for (unsigned i = 0; i < d; ++i) s += p[i] * p[i];
    return s;
}

bool L2Normalize(float* p, unsigned d) {
    const double norm = vec_norm(p, d);
    if (norm < 1e-10) {
        std::cerr << "Vector is fully zero " << std::endl;
        for (unsigned i = 0; i < d; ++i) std::cerr << p[i] << " ";
        std::cerr << std::endl;
        return false;
    }
    const double root = std::sqrt(norm);
    for (unsigned i = 0; i < d; ++i) p[i] /= root;
    return true;
}

float distance(float* p, float* q, unsigned d) {
#ifdef MIPS_DISTANCE
    return mips_distance(p, q, d);
#else
    return sqr_l2_dist(p, q, d);
#endif
}

float pos_distance(float* p, float* q, unsigned d) {
#ifdef MIPS_DISTANCE
    return distance(p, q, d) + 1.0;
#else
    return distance(p, q, d);
#endif
}
