// topn.h -- drop-in for gp-ann's src/topn.h:6-32: bounded accumulator of the n lexicographically smallest
// (distance, id) pairs.  Same public surface (n, value_type, Add, Take, Top); implemented as a max-heap over a
// std::vector.  Stays host code: it is used per query inside worker threads for small candidate sets
// (router votes, leaders); the big scans use the GPU selection in libgpann_cuda instead.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

struct TopN {
    using value_type = std::pair<float, uint32_t>;

    explicit TopN(size_t n) : n(n) { heap.reserve(n + 1); }

    size_t n;
    std::vector<value_type> heap;  // max-heap: heap.front() is the current n-th best

    void Add(value_type e) {
        if (heap.size() < n) {
            heap.push_back(e);
            std::push_heap(heap.begin(), heap.end());
        } else if (n > 0 && e < heap.front()) {  // ties on distance keep the smaller id
            std::pop_heap(heap.begin(), heap.end());
            heap.back() = e;
            std::push_heap(heap.begin(), heap.end());
        }
    }

    // ascending; leaves the accumulator empty
    std::vector<value_type> Take() {
        std::sort_heap(heap.begin(), heap.end());
        std::vector<value_type> out;
        out.swap(heap);
        return out;
    }

    const value_type& Top() const { return heap.front(); }
    size_t size() const { return heap.size(); }
};
