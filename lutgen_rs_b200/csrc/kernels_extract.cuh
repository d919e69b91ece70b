// kernels_extract.cuh -- palette extraction (k-means in Oklab): the kernel bodies of extract_core.h, one
// thread per pixel / colour key / point / centroid. See extract_core.h for the specification and for what
// this is and is not with respect to the reference (`lutgen extract`, crates/cli/src/main.rs:772-815).
#pragma once
#include <cstdint>

#include "extract_core.h"

namespace lutgen {

constexpr int EXT_THREADS = 256;

__global__ void __launch_bounds__(EXT_THREADS) k_ext_hist(const uint8_t* __restrict__ rgb, size_t npix, uint32_t* counts) {
    const size_t i = (size_t)blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < npix) lext::hist_body(i, rgb, counts);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_compact(const uint32_t* __restrict__ counts, uint32_t* n_points, uint32_t* keys, uint32_t* weights,
                                                            uint8_t* rgb_list) {
    const uint32_t key = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (key < lext::NKEYS) lext::compact_body(key, counts, n_points, keys, weights, rgb_list);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_quantise(const float* __restrict__ lab, uint32_t n, int32_t* q, unsigned long long* dmin) {
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < n) lext::quantise_body(i, lab, q, dmin);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_seed_score(uint32_t n, uint32_t j, const int32_t* __restrict__ q, const uint32_t* __restrict__ weights,
                                                               const int32_t* centroids, const uint32_t* n_centroids, unsigned long long* dmin,
                                                               lext::Seed* seed) {
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < n) lext::seed_score_body(i, j, q, weights, centroids, n_centroids, dmin, seed);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_seed_pick(uint32_t n, uint32_t j, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ weights,
                                                              const uint32_t* n_centroids, const unsigned long long* __restrict__ dmin, lext::Seed* seed) {
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < n) lext::seed_pick_body(i, j, keys, weights, n_centroids, dmin, seed);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_seed_commit(uint32_t n, uint32_t j, const uint32_t* __restrict__ keys, const int32_t* __restrict__ q,
                                                                const uint32_t* __restrict__ weights, const unsigned long long* __restrict__ dmin,
                                                                const lext::Seed* seed, int32_t* centroids, uint32_t* n_centroids) {
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < n) lext::seed_commit_body(i, j, keys, q, weights, dmin, seed, centroids, n_centroids);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_assign(uint32_t n, const int32_t* __restrict__ q, const uint32_t* __restrict__ weights,
                                                           const int32_t* centroids, const uint32_t* n_centroids, long long* sums) {
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (i < n) lext::assign_body(i, q, weights, centroids, n_centroids, sums);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_update(const uint32_t* n_centroids, long long* sums, int32_t* centroids) {
    lext::update_body(threadIdx.x, n_centroids, sums, centroids);
}

__global__ void __launch_bounds__(EXT_THREADS) k_ext_output(const int32_t* centroids, float* lab) { lext::output_body(threadIdx.x, centroids, lab); }

}  // namespace lutgen
