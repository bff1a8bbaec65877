// gpann_dropin.h -- glue between the reference-shaped C++ entry points (recall.h, kmeans.h, ClosestLeaders) and
// the C ABI of libgpann_cuda.so (include/gpann.h).
//
// One process-wide gpann_ctx, created on first use: a group over the devices listed in $GPANN_DEVICES ("0,1,2,3":
// gpann_init_multi -- row-sharded point sets, NCCL all-gather / all-reduce inside the library, so ComputeGroundTruth and
// KMeans keep their single-call signatures at 8 GPUs, src/recall.h:25-39, src/kmeans.cpp:228-244), else the single
// device $GPANN_DEVICE (default 0).  The reference calls
// KMeans concurrently from parlay workers (kmeans_tree_router.cpp:22-37, partitioning.cpp:468-476) while a
// gpann_ctx is externally synchronised, so every batched call holds the context mutex.  Errors become
// std::runtime_error, the reference's own convention (kmeans.cpp:229-231); there is no CPU fallback.
#pragma once

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gpann.h"

namespace gpann_dropin {

#ifdef MIPS_DISTANCE
constexpr int kMetric = GPANN_METRIC_MIPS;
#else
constexpr int kMetric = GPANN_METRIC_L2;
#endif

struct Session {
    gpann_ctx* ctx = nullptr;
    gpann_ctx* single = nullptr;  // entry points that take a single-device ctx (BalancedKMeans, segmented scans): ctx itself, or
                                  // a ctx on the group's first device created on first use
    int first_device = 0;
    std::mutex mutex;
    Session() {
        std::vector<int> devices;
        if (const char* list = std::getenv("GPANN_DEVICES")) {
            for (const char* p = list; *p;) {
                char* end = nullptr;
                const long v = std::strtol(p, &end, 10);
                if (end == p) break;
                devices.push_back((int) v);
                p = (*end == ',') ? end + 1 : end;
            }
        }
        const char* dev = std::getenv("GPANN_DEVICE");
        first_device = !devices.empty() ? devices[0] : (dev ? std::atoi(dev) : 0);
        const int rc = devices.size() > 1 ? gpann_init_multi(&ctx, devices.data(), (int) devices.size()) : gpann_init(&ctx, first_device);
        if (rc != 0)
            throw std::runtime_error("gpann: no usable CUDA device (gpann_init returned " + std::to_string(rc) +
                                     "); the GPU hot path has no CPU fallback");
    }
    ~Session() {
        if (single && single != ctx) gpann_destroy(single);
        gpann_destroy(ctx);
    }
    gpann_ctx* single_device() {
        if (!single) {
            if (gpann_num_devices(ctx) == 1) single = ctx;
            else if (gpann_init(&single, first_device) != 0) throw std::runtime_error("gpann: cannot create a single-device context");
        }
        return single;
    }
    void check_single(int rc) const {
        if (rc != 0) throw std::runtime_error(std::string("gpann: ") + gpann_last_error(single));
    }
    void check(int rc) const {
        if (rc != 0) throw std::runtime_error(std::string("gpann: ") + gpann_last_error(ctx));
    }
};

inline Session& session() {
    static Session s;
    return s;
}

// RAII upload of a PointSet-like (coordinates/n/d) object.
template <class PS>
struct DevicePoints {
    gpann_points* h = nullptr;
    DevicePoints(Session& s, PS& ps, int storage = GPANN_STORAGE_AUTO) {
        s.check(gpann_points_upload(s.ctx, ps.coordinates.data(), ps.n, ps.d, storage, 0, &h));
    }
    ~DevicePoints() { gpann_points_free(h); }
    DevicePoints(const DevicePoints&) = delete;
    DevicePoints& operator=(const DevicePoints&) = delete;
};

}  // namespace gpann_dropin
