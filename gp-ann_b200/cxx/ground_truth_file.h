// ground_truth_file.h -- ComputeGroundTruth straight from a point FILE: the reference's reader widens
// .u8bin / .i8bin to float (src/points_io.cpp:63-111: 4x the host memory, 4x the PCIe bytes, 51 GB for BigANN-100M) before
// ComputeGroundTruth (src/recall.h:25-39) scans it.  Here the file is read in row chunks AS STORED -- bytes stay bytes --
// and every chunk is streamed through the GPU(s) by gpann_topk_stream_bytes / gpann_topk_stream with global ids; the
// per-chunk lists are merged lexicographically by (distance, id), which is what one scan over the whole file returns.
// Host memory: one chunk.  The result type and the ground-truth FILE format (points_io.cpp:136-179: u32 nq, u32 k, ids, then
// distances) are the reference's, so WriteGroundTruth / ReadGroundTruth of either side interoperate.
// Like recall.h this header DEFINES its function: include it from one translation unit.
#pragma once

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "defs.h"
#include "gpann_dropin.h"

std::vector<NNVec> ComputeGroundTruthFromFile(const std::string& points_path, PointSet& queries, int k, size_t chunk_rows = 0) {
    auto ends_with = [&](const char* suffix) {
        const std::string t(suffix);
        return points_path.size() >= t.size() && points_path.compare(points_path.size() - t.size(), t.size(), t) == 0;
    };
    const bool is_u8 = ends_with(".u8bin"), is_i8 = ends_with(".i8bin"), is_f32 = ends_with(".fbin");
    if (!is_u8 && !is_i8 && !is_f32)
        throw std::runtime_error("Invalid file ending for the pointset path. Valid options are [.fbin, .u8bin, .i8bin]");
    std::ifstream in(points_path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + points_path);
    uint32_t n = 0, d = 0;
    in.read(reinterpret_cast<char*>(&n), sizeof(n));
    in.read(reinterpret_cast<char*>(&d), sizeof(d));
    if (d != queries.d) throw std::runtime_error("ComputeGroundTruthFromFile: dimensions of points and queries differ");
    const size_t esz = is_f32 ? 4 : 1;
    if (chunk_rows == 0) chunk_rows = std::max<size_t>(1, (size_t(1) << 30) / (size_t(d) * esz));  // ~1 GB of file per chunk
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    const size_t nq = queries.n;
    std::vector<char> chunk(std::min<size_t>(chunk_rows, std::max<uint32_t>(n, 1)) * d * esz);
    std::vector<float> dist(nq * k), best_d(nq * k);
    std::vector<uint32_t> ids(nq * k), count(nq), best_i(nq * k), best_c(nq, 0);
    std::vector<std::pair<float, uint32_t>> merged(2 * size_t(k));
    for (size_t row0 = 0; row0 < n; row0 += chunk_rows) {
        const size_t rows = std::min<size_t>(chunk_rows, n - row0);
        in.read(chunk.data(), rows * d * esz);
        if (!in) throw std::runtime_error("short read from " + points_path);
        int rc;
        if (is_f32)
            rc = gpann_topk_stream(s.ctx, reinterpret_cast<const float*>(chunk.data()), rows, d, GPANN_STORAGE_AUTO, (uint32_t) row0,
                                   queries.coordinates.data(), nq, k, 0, dist.data(), ids.data(), count.data(), nullptr);
        else
            rc = gpann_topk_stream_bytes(s.ctx, chunk.data(), rows, d, is_i8 ? 1 : 0, (uint32_t) row0, queries.coordinates.data(), nq, k, 0,
                                         dist.data(), ids.data(), count.data());
        s.check(rc);
        for (size_t q = 0; q < nq; ++q) {  // lexicographic merge of two ascending lists: O(k) per query and chunk
            size_t m = 0;
            for (uint32_t r = 0; r < best_c[q]; ++r) merged[m++] = { best_d[q * k + r], best_i[q * k + r] };
            for (uint32_t r = 0; r < count[q]; ++r) merged[m++] = { dist[q * k + r], ids[q * k + r] };
            std::inplace_merge(merged.begin(), merged.begin() + best_c[q], merged.begin() + m);
            best_c[q] = (uint32_t) std::min<size_t>(m, k);
            for (uint32_t r = 0; r < best_c[q]; ++r) {
                best_d[q * k + r] = merged[r].first;
                best_i[q * k + r] = merged[r].second;
            }
        }
    }
    std::vector<NNVec> res(nq);
    for (size_t q = 0; q < nq; ++q) {
        res[q].resize(best_c[q]);
        for (uint32_t r = 0; r < best_c[q]; ++r) res[q][r] = { best_d[q * k + r], best_i[q * k + r] };
    }
    return res;
}
