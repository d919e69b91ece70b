// extract_sim.cpp -- host instantiation of lutgen_rs_b200/csrc/extract_core.h for the CPU unit tests.
//
// TEST INFRASTRUCTURE ONLY. The palette-extraction kernel bodies are plain functions of a global index; the
// CUDA kernels (csrc/kernels_extract.cuh) call them with one thread per index, this file calls them in loops --
// in a scrambled order, since nothing may depend on the order in which threads run or in which the compaction
// appends. The two Oklab conversions between the stages are done by the caller with the oracle's functions (on
// the GPU: the remappers' conversion kernels, tested bit exact against the same). Nothing in the product path
// links or calls this.
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../lutgen_rs_b200/csrc/extract_core.h"

namespace {
// a fixed pseudo-random permutation of 0..n-1
std::vector<uint32_t> scrambled(uint32_t n, uint64_t seed) {
    std::vector<uint32_t> p(n);
    std::iota(p.begin(), p.end(), 0u);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    for (uint32_t i = n; i > 1; --i) {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        std::swap(p[i - 1], p[x % i]);
    }
    return p;
}
}  // namespace

extern "C" {

// Stage 1: histogram and compaction. keys / weights / rgb_list must hold min(npix, 2^24) entries. Returns the number of points.
uint32_t lext_sim_points(const uint8_t* rgb, size_t npix, uint64_t seed, uint32_t* keys, uint32_t* weights, uint8_t* rgb_list) {
    using namespace lext;
    std::vector<uint32_t> counts(NKEYS, 0u);
    for (size_t i = npix; i-- > 0;) hist_body(i, rgb, counts.data());
    // visit the keys in a scrambled order: blocks of 4096 keys permuted, reversed inside
    std::vector<uint32_t> blocks = scrambled(NKEYS / 4096, seed);
    uint32_t n = 0;
    for (uint32_t b : blocks)
        for (uint32_t k = 4096; k-- > 0;) compact_body(b * 4096 + k, counts.data(), &n, keys, weights, rgb_list);
    return n;
}

// Stage 2: everything between the two conversions. lab: n x 3 floats (Oklab of rgb_list); out_lab: 256 x 3 floats. Returns the number of centroids.
uint32_t lext_sim_kmeans(uint32_t n, const uint32_t* keys, const uint32_t* weights, const float* lab, uint32_t color_count, uint32_t iterations,
                         uint64_t seed, float* out_lab) {
    using namespace lext;
    std::vector<int32_t> q(3 * (size_t)n);
    std::vector<unsigned long long> dmin(n);
    memset(q.data(), 0xA5, q.size() * sizeof(int32_t));
    memset(dmin.data(), 0xA5, dmin.size() * sizeof(unsigned long long));
    std::vector<uint32_t> order = scrambled(n, seed + 1);
    for (uint32_t i : order) quantise_body(i, lab, q.data(), dmin.data());
    int32_t centroids[3 * MAX_COLORS];
    long long sums[4 * MAX_COLORS];
    memset(centroids, 0, sizeof centroids);
    memset(sums, 0, sizeof sums);
    uint32_t n_centroids = 0;
    const uint32_t k = color_count < n ? color_count : n;
    for (uint32_t j = 0; j < k; ++j) {
        Seed s;
        memset(&s, 0, sizeof s);
        order = scrambled(n, seed + 2 + j);
        for (uint32_t i : order) seed_score_body(i, j, q.data(), weights, centroids, &n_centroids, dmin.data(), &s);
        for (uint32_t i : order) seed_pick_body(i, j, keys, weights, &n_centroids, dmin.data(), &s);
        for (uint32_t i : order) seed_commit_body(i, j, keys, q.data(), weights, dmin.data(), &s, centroids, &n_centroids);
    }
    for (uint32_t it = 0; it < iterations; ++it) {
        order = scrambled(n, seed + 1000 + it);
        for (uint32_t i : order) assign_body(i, q.data(), weights, centroids, &n_centroids, sums);
        for (uint32_t j = MAX_COLORS; j-- > 0;) update_body(j, &n_centroids, sums, centroids);
    }
    for (uint32_t j = 0; j < MAX_COLORS; ++j) output_body(j, centroids, out_lab);
    return n_centroids;
}
}
