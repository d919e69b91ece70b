// extract_core.h -- palette extraction by k-means in Oklab (`lutgen extract`,
// crates/cli/src/main.rs:772-815, which runs the `quantette` crate's PalettePipeline with
// ColorSpace::Oklab, QuantizeMethod::kmeans() and palette_size(n) over the image).
//
// PARITY UNPINNED. `quantette` is not vendored under the reference and nothing in the reference pins
// its output (no test, no fixture); its k-means is an online, sampled variant seeded by Wu's
// quantizer, driven by its own RNG -- it cannot be restated from the call site. What is implemented
// here is a DETERMINISTIC k-means on the same data in the same colour space, specified below and
// restated independently in oracle/lutgen_oracle.c (lutgen_oracle_extract_palette); the CUDA path is
// bit exact with that restatement, and its palettes are compared with scikit-learn's k-means for
// quality in the tests. A caller gets "a k-means palette in Oklab", not quantette's palette.
//
// Specification.
//   points     the distinct colours of the image; weight w = number of pixels; coordinates q = Oklab (the
//              conversion of the remappers, bit exact) in 16.16 fixed point, round to nearest even
//   distance   d(p, c) = sum over the three coordinates of (q_p - c)^2, exact in 64-bit integers
//   seeding    centroid 0 = the point with the largest weight; centroid j = the point maximising
//              min(w, 65535) * min_{i<j} d(p, centroid i); ties go to the smallest colour key
//              (r | g << 8 | b << 16); seeding stops early when every point sits on a centroid
//   iteration  (Lloyd) every point joins its nearest centroid (ties: lowest index); a centroid with members
//              moves to round(sum w q / sum w) (half away from zero), one without stays; `iterations` times
//   output     centroids, in index order, back through Oklab -> sRGB8
// Every step is a sum, a minimum or a maximum over points, so the order in which threads (or the
// compaction) visit them cannot change the result: integer atomics only, no floating-point sums.
//
// The kernel bodies below are plain C++ over a global index (`gid`) that compiles for the host as well:
// csrc/kernels_extract.cuh wraps them in kernels, tests/extract_sim.cpp in loops (in scrambled order),
// so the CPU tests run the code the GPU runs. The host build is test infrastructure only.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LEXT_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define LEXT_HD inline
#endif

namespace lext {

constexpr int FIX_BITS = 16;
constexpr uint32_t MAX_COLORS = 256;
constexpr uint32_t WEIGHT_CAP = 65535u;   // seeding weighs a point by min(w, WEIGHT_CAP): the product with d stays below 2^64
constexpr uint32_t NKEYS = 1u << 24;

struct Seed {                      // per seeding step, zeroed before it
    unsigned long long best_score; // max over points of the seeding score
    uint32_t best_inv_key;         // max over the points that reach it of ~key (i.e. the smallest key)
    uint32_t pad;
};

LEXT_HD void atomic_add_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
LEXT_HD uint32_t atomic_inc_u32(uint32_t* p) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}
LEXT_HD void atomic_max_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LEXT_HD void atomic_max_u64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LEXT_HD void atomic_add_i64(long long* p, long long v) {
#ifdef __CUDA_ARCH__
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);  // two's complement
#else
    *p += v;
#endif
}

LEXT_HD int32_t to_fixed(float v) {
#ifdef __CUDA_ARCH__
    return __float2int_rn(v * 65536.0f);
#else
    return (int32_t)lrintf(v * 65536.0f);  // default rounding mode: to nearest even, like cvt.rni
#endif
}

LEXT_HD unsigned long long dist2(const int32_t* a, const int32_t* b) {
    unsigned long long d = 0;
    for (int c = 0; c < 3; ++c) {
        long long t = (long long)a[c] - (long long)b[c];
        d += (unsigned long long)(t * t);
    }
    return d;
}

LEXT_HD long long div_round(long long s, long long w) {  // w > 0; round half away from zero
    return s >= 0 ? (2 * s + w) / (2 * w) : -((-2 * s + w) / (2 * w));
}

// counts[key] += 1 for pixel gid
LEXT_HD void hist_body(size_t gid, const uint8_t* rgb, uint32_t* counts) {
    const uint32_t key = (uint32_t)rgb[3 * gid] | ((uint32_t)rgb[3 * gid + 1] << 8) | ((uint32_t)rgb[3 * gid + 2] << 16);
    atomic_add_u32(&counts[key], 1u);
}

// colour `key` (gid over all 2^24) joins the point list if it occurs: its key, weight and sRGB bytes
LEXT_HD void compact_body(uint32_t key, const uint32_t* counts, uint32_t* n_points, uint32_t* keys, uint32_t* weights, uint8_t* rgb_list) {
    const uint32_t w = counts[key];
    if (!w) return;
    const uint32_t i = atomic_inc_u32(n_points);
    keys[i] = key;
    weights[i] = w;
    rgb_list[3 * i] = (uint8_t)key;
    rgb_list[3 * i + 1] = (uint8_t)(key >> 8);
    rgb_list[3 * i + 2] = (uint8_t)(key >> 16);
}

// Oklab (float, from the remappers' conversion kernel) -> fixed point; the distance to the nearest seed starts at "infinite"
LEXT_HD void quantise_body(uint32_t gid, const float* lab, int32_t* q, unsigned long long* dmin) {
    for (int c = 0; c < 3; ++c) q[3 * gid + c] = to_fixed(lab[3 * gid + c]);
    dmin[gid] = ~0ull;
}

// The seeding score of point gid at step j (after folding centroid j-1 into its distance).
LEXT_HD unsigned long long seed_score(uint32_t gid, uint32_t j, const uint32_t* weights, const unsigned long long* dmin) {
    if (j == 0) return weights[gid];
    const unsigned long long w = weights[gid] < WEIGHT_CAP ? weights[gid] : WEIGHT_CAP;
    return w * dmin[gid];
}

// step j, pass 1: fold centroid j-1 into dmin (if it exists), raise best_score
LEXT_HD void seed_score_body(uint32_t gid, uint32_t j, const int32_t* q, const uint32_t* weights, const int32_t* centroids, const uint32_t* n_centroids,
                             unsigned long long* dmin, Seed* seed) {
    if (j > 0) {
        if (*n_centroids < j) return;  // seeding has stopped: nothing new to fold in, nothing to pick
        const unsigned long long d = dist2(q + 3 * gid, centroids + 3 * (j - 1));
        if (d < dmin[gid]) dmin[gid] = d;
    }
    atomic_max_u64(&seed->best_score, seed_score(gid, j, weights, dmin));
}

// pass 2: among the points that reach best_score, the smallest key
LEXT_HD void seed_pick_body(uint32_t gid, uint32_t j, const uint32_t* keys, const uint32_t* weights, const uint32_t* n_centroids,
                            const unsigned long long* dmin, Seed* seed) {
    if (j > 0 && *n_centroids < j) return;
    const unsigned long long best = seed->best_score;
    if (best == 0 && j > 0) return;  // every point sits on a centroid already
    if (seed_score(gid, j, weights, dmin) == best) atomic_max_u32(&seed->best_inv_key, ~keys[gid]);
}

// pass 3: that point becomes centroid j
LEXT_HD void seed_commit_body(uint32_t gid, uint32_t j, const uint32_t* keys, const int32_t* q, const uint32_t* weights, const unsigned long long* dmin,
                              const Seed* seed, int32_t* centroids, uint32_t* n_centroids) {
    if (j > 0 && (*n_centroids < j || seed->best_score == 0)) return;
    if (~keys[gid] != seed->best_inv_key || seed_score(gid, j, weights, dmin) != seed->best_score) return;
    for (int c = 0; c < 3; ++c) centroids[3 * j + c] = q[3 * gid + c];
    *n_centroids = j + 1;
}

// Lloyd, pass 1: point gid adds itself to its nearest centroid's sums (sums[4 j + 0..2] = sum w q, [4 j + 3] = sum w)
LEXT_HD void assign_body(uint32_t gid, const int32_t* q, const uint32_t* weights, const int32_t* centroids, const uint32_t* n_centroids, long long* sums) {
    const uint32_t k = *n_centroids;
    uint32_t best = 0;
    unsigned long long bd = dist2(q + 3 * gid, centroids);
    for (uint32_t j = 1; j < k; ++j) {
        const unsigned long long d = dist2(q + 3 * gid, centroids + 3 * j);
        if (d < bd) {
            bd = d;
            best = j;
        }
    }
    const long long w = weights[gid];
    for (int c = 0; c < 3; ++c) atomic_add_i64(&sums[4 * best + c], w * (long long)q[3 * gid + c]);
    atomic_add_i64(&sums[4 * best + 3], w);
}

// Lloyd, pass 2: centroid j moves to the weighted mean of its members; its sums are cleared for the next pass
LEXT_HD void update_body(uint32_t j, const uint32_t* n_centroids, long long* sums, int32_t* centroids) {
    if (j >= *n_centroids) return;
    const long long w = sums[4 * j + 3];
    if (w > 0)
        for (int c = 0; c < 3; ++c) centroids[3 * j + c] = (int32_t)div_round(sums[4 * j + c], w);
    for (int c = 0; c < 4; ++c) sums[4 * j + c] = 0;
}

// centroid j as float Oklab (exact: |q| < 2^18 fits a float's 24-bit significand, and the scale is a power of two)
LEXT_HD void output_body(uint32_t j, const int32_t* centroids, float* lab) {
    for (int c = 0; c < 3; ++c) lab[3 * j + c] = (float)centroids[3 * j + c] * (1.0f / 65536.0f);
}

}  // namespace lext
