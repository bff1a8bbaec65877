// points_io.cpp (standalone builds only) -- the file formats of gp-ann's src/points_io.cpp, host code:
//   .fbin / .u8bin / .i8bin : u32 n, u32 d, then n*d coordinates (float / uint8 / int8); bytes are widened to float (:63-111)
//   ground-truth file       : u32 num_queries, u32 num_neighbors, ids[nq*k] (u32), then distances[nq*k] (f32)  (:136-179)
#include "points_io.h"

#include <fstream>
#include <stdexcept>

namespace {
bool ends_with(const std::string& s, const char* suffix) {
    const std::string t(suffix);
    return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}

template <typename T>
PointSet ReadAs(const std::string& path, int64_t size) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    uint32_t n = 0, d = 0;
    in.read(reinterpret_cast<char*>(&n), sizeof(n));
    in.read(reinterpret_cast<char*>(&d), sizeof(d));
    if (size != -1) n = static_cast<uint32_t>(size);
    PointSet points;
    points.n = n;
    points.d = d;
    points.coordinates.resize(points.n * points.d);
    std::vector<T> buffer(1 << 20);
    size_t done = 0;
    while (done < points.coordinates.size()) {
        const size_t take = std::min(buffer.size(), points.coordinates.size() - done);
        in.read(reinterpret_cast<char*>(buffer.data()), take * sizeof(T));
        for (size_t i = 0; i < take; ++i) points.coordinates[done + i] = static_cast<float>(buffer[i]);
        done += take;
    }
    return points;
}
}  // namespace

PointSet ReadPoints(const std::string& path, int64_t size) {
    if (ends_with(path, ".fbin")) return ReadAs<float>(path, size);
    if (ends_with(path, ".u8bin")) return ReadAs<uint8_t>(path, -1);
    if (ends_with(path, ".i8bin")) return ReadAs<int8_t>(path, -1);
    throw std::runtime_error("Invalid file ending for the pointset path. Valid options are [.fbin, .u8bin, .i8bin]");
}

void WritePoints(PointSet& points, const std::string& path) {
    std::ofstream out(path, std::ios::binary);
    const uint32_t n = static_cast<uint32_t>(points.n), d = static_cast<uint32_t>(points.d);
    out.write(reinterpret_cast<const char*>(&n), sizeof(n));
    out.write(reinterpret_cast<const char*>(&d), sizeof(d));
    out.write(reinterpret_cast<const char*>(points.coordinates.data()), points.coordinates.size() * sizeof(float));
}

std::vector<NNVec> ReadGroundTruth(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    uint32_t nq = 0, k = 0;
    in.read(reinterpret_cast<char*>(&nq), sizeof(nq));
    in.read(reinterpret_cast<char*>(&k), sizeof(k));
    std::vector<uint32_t> ids(static_cast<size_t>(nq) * k);
    std::vector<float> dist(ids.size());
    in.read(reinterpret_cast<char*>(ids.data()), ids.size() * sizeof(uint32_t));
    in.read(reinterpret_cast<char*>(dist.data()), dist.size() * sizeof(float));
    std::vector<NNVec> gt(nq, NNVec(k));
    for (size_t q = 0; q < nq; ++q)
        for (size_t j = 0; j < k; ++j) gt[q][j] = { dist[q * k + j], ids[q * k + j] };
    return gt;
}

void WriteGroundTruth(const std::string& path, const std::vector<NNVec>& ground_truth) {
    const uint32_t nq = static_cast<uint32_t>(ground_truth.size());
    const uint32_t k = nq ? static_cast<uint32_t>(ground_truth[0].size()) : 0;
    std::vector<uint32_t> ids(static_cast<size_t>(nq) * k);
    std::vector<float> dist(ids.size());
    for (size_t q = 0; q < nq; ++q)
        for (size_t j = 0; j < k; ++j) {
            ids[q * k + j] = ground_truth[q][j].second;
            dist[q * k + j] = ground_truth[q][j].first;
        }
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(&nq), sizeof(nq));
    out.write(reinterpret_cast<const char*>(&k), sizeof(k));
    out.write(reinterpret_cast<const char*>(ids.data()), ids.size() * sizeof(uint32_t));
    out.write(reinterpret_cast<const char*>(dist.data()), dist.size() * sizeof(float));
}
