// kernels_fast.cuh -- the throughput remap kernel: tile-bounded nearest-k in place of the
// reference's kd-tree (kiddo nearest_n / nearest_one; rbf.rs:85-97, nearest_neighbor.rs:48).
//
// One CTA (256 threads) owns a TILE of colours that are close together in sRGB, hence in
// Oklab: a tr x tg x tb brick of Hald-identity entries (LUT / TABLE front-ends; power-of-two
// sides so indices are shifts) or 1024 consecutive pixels (IMAGE). Per tile:
//   1. every thread converts its colours to Oklab (f32, MUFU-seeded cube root) and the CTA
//      reduces their bounding box (REDUX on order-preserving integer keys);
//   2. for every palette colour p the CTA computes LB(p)/UB(p), the smallest/largest squared
//      distance from p to any point of the (slightly inflated) box;
//   3. U_k = k-th smallest UB bounds every pixel's k-th neighbour distance, so colours with
//      LB > U_k are OUT for the whole tile; a colour with at most k-1 rivals whose LB <= its UB
//      is IN for the whole tile; the rest are UNCERTAIN;
//   4. colours whose weight cannot reach 2^-24 of the denominator anywhere in the tile
//      (w(LB) <= 2^-24 w(min UB)) are dropped from both lists: below f32 resolution of the sums;
//   5. each thread accumulates the IN colours straight from a compacted shared-memory list
//      (broadcast LDS, no selection) and picks its k' = k - |IN| nearest among the UNCERTAIN
//      ones with a register-resident bitonic network.
// The bounds are conservative, so the set of neighbours equals the brute-force set except at
// near-ties in step 5 (gap below the f32 error bound); those entries are appended to a work
// list and recomputed by the exact f64 kernel (kernels_exact.cuh), which makes the result
// independent of the fast path's approximations wherever they could matter.
//
// Weights are evaluated in f32 relative to the tile's smallest possible distance d0 = min LB:
//   Gaussian exp(-s d)      -> ex2(c1*d + c0),        c1 = -s*log2(e), c0 = -c1*d0
//   Shepard  1/sqrt(d)^p    -> ex2(c1*lg2(d) + c0),   c1 = -p/2,       c0 = -c1*lg2(d0)
//   Linear   d
// (a common factor cancels in num/den; the f64 reference underflows far later than f32 would
// without the shift). A denominator that still leaves [1e-30, 1e30] flags the entry.
#pragma once
#include "kernels_exact.cuh"

namespace lutgen {

constexpr int FAST_THREADS = 256;
constexpr int FAST_MAX_EPT = 4;                    // entries per thread
constexpr int FAST_TILE_MAX = FAST_THREADS * FAST_MAX_EPT;
constexpr int FAST_MAX_HITS = 16;
constexpr uint32_t FAST_BUCKET_MAX = 16;           // uncertain sets up to this size use the register buckets (measured: 16 beats 8)
constexpr int KIND_NN = 3;
// Step 4 threshold: a colour whose weight cannot reach 2^-16 of the tile's largest anywhere in the tile is dropped.
// The f32 weights themselves carry relative errors of about 2^-14 (the distance error of the fast conversion, 3.7e-7
// at d = 0.1, times the Gaussian's slope of 185 per unit of distance at shape 128), so what is dropped is below what
// the arithmetic resolves. Measured on the level-16 headline CLUT against 2^-24 (round 1's value): 6 % faster,
// 2480 of 50.3 M bytes change, each by one (2^-12: 14 % faster, 77 935 bytes; 2^-20: 1 %, 101 bytes) -- the +-1 LSB
// contract holds with the golden CLUTs' byte counts inside the tests' limits (tests/test_gpu_parity.py).
#ifndef FAST_NEGLIGIBLE_LOG2_V
#define FAST_NEGLIGIBLE_LOG2_V -16.0f
#endif
constexpr float FAST_NEGLIGIBLE_LOG2 = FAST_NEGLIGIBLE_LOG2_V;
constexpr float FAST_TIE_TOL = 5e-10f;             // (gap)^2 <= TOL * d  ->  recompute exactly

struct FastArgs {
    Front f;
    const float4* pal;        // n x (l*lum, a, b, 0)  f32
    const uint32_t* pal_rgb;  // n x packed r | g<<8 | b<<16 (from pal_rgb4)
    const float* pal_ab;      // n x 2 (NearestNeighbor preserve)
    uint32_t n, k;            // k = 0: every colour
    int preserve;
    float lum, inv_lum;
    float c1;                 // weight slope (see header)
    uint32_t sh_r, sh_g;      // log2 of the brick's r and g sides (LUT / TABLE)
    uint32_t tr, tg, tb;      // brick sides
    uint32_t tiles_r, tiles_g;
    uint32_t b_lo;            // first blue plane covered by the grid (multiple of tb)
    uint32_t ept;             // entries per thread
    uint32_t ntiles;          // tiles in this launch (CTAs stride over them)
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t* tile_counter;   // zeroed before the launch
    uint32_t* stats;          // optional [128]: histogram of |UNCERTAIN| and |IN| per tile
    const ColorTables* tables;
};

__device__ __forceinline__ uint32_t f2key(float x) {
    uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

template <int KIND>
__device__ __forceinline__ float fast_weight(float d, float c1, float c0) {
    if (KIND == 0) return ex2_approx(fmaf(c1, d, c0));
    if (KIND == 1) return ex2_approx(fmaf(c1, lg2_approx(d), c0));
    return d;
}

__device__ __forceinline__ float dist3(float L, float A, float B, float4 p) {
    float dx = L - p.x, dy = A - p.y, dz = B - p.z;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

__device__ __forceinline__ void flag_entry(const FastArgs& a, unsigned long long slot) {
    uint32_t pos = atomicAdd(a.flag_count, 1u);
    a.flag_list[pos] = (uint32_t)slot;
    if (a.stats) atomicAdd(&a.stats[127], 1u);
}

// Ascending bitonic network over M registers (all indices compile-time after unrolling).
template <int M>
__device__ __forceinline__ void bitonic_sort(float (&v)[M]) {
#pragma unroll
    for (int size = 2; size <= M; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                int j = i ^ stride;
                if (j > i) {
                    bool up = ((i & size) == 0);
                    float lo = fminf(v[i], v[j]), hi = fmaxf(v[i], v[j]);
                    v[i] = up ? lo : hi;
                    v[j] = up ? hi : lo;
                }
            }
        }
    }
}

// Pick the kp nearest of mU (kp < mU <= M) uncertain colours and accumulate them. Returns true when
// the choice is within the f32 error bound of flipping (the entry must be recomputed exactly).
// The sorted copy lives in registers; distances are recomputed for the accumulation pass (6 cheap
// instructions each) rather than kept in a second register array.
// kp is uniform across the CTA: a chain of uniform branches reads s[kp-1], s[kp] from registers
// without turning the array into local memory.
template <int M, int I = 1>
__device__ __forceinline__ void pick_threshold(const float (&s)[M], int kp, float& tau, float& nxt) {
    if (kp == I) {
        tau = s[I - 1];
        nxt = s[I];
    } else if constexpr (I + 1 < M) {
        pick_threshold<M, I + 1>(s, kp, tau, nxt);
    }
}

template <int M, int KIND>
__device__ __forceinline__ bool select_accumulate(const float4* __restrict__ unc, int mU, int kp, float L, float A, float B, float c1,
                                                  float c0, float& n0, float& n1, float& n2, float& den) {
    constexpr bool KEEP = M <= 8;  // small sets keep their distances; M = 16 recomputes them (registers)
    float s[M], du[KEEP ? M : 1];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        s[i] = i < mU ? dist3(L, A, B, unc[i]) : INFINITY;
        if (KEEP) du[i] = s[i];
    }
    bitonic_sort<M>(s);
    float tau = s[0], nxt = s[M - 1];
    pick_threshold<M>(s, kp, tau, nxt);
    float gap = nxt - tau;
    bool risky = (gap * gap <= FAST_TIE_TOL * nxt) && (nxt < INFINITY);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (i < mU) {
            float4 p = unc[i];
            float d = KEEP ? du[KEEP ? i : 0] : dist3(L, A, B, p);
            float w = fast_weight<KIND>(d, c1, c0);
            w = d <= tau ? w : 0.0f;
            n0 = fmaf(p.x, w, n0);
            n1 = fmaf(p.y, w, n1);
            n2 = fmaf(p.z, w, n2);
            den += w;
        }
    }
    return risky;
}

// Same contract for any number of uncertain colours, without register arrays: kp + 1 passes, each
// extracting the next (distance, index) in ascending order. Only tiles with more than 16
// uncertain colours come here.
struct LoopAcc {
    float n0, n1, n2, den;
    int risky;
};
template <int KIND>
__device__ __noinline__ LoopAcc select_accumulate_loop(const float4* __restrict__ unc, int mU, int kp, float L, float A, float B, float c1,
                                                       float c0) {
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
    float last = -1.0f, tau = 0.0f;
    int last_j = -1;
    for (int pass = 0; pass <= kp; ++pass) {
        float best = INFINITY;
        int best_j = -1;
        for (int j = 0; j < mU; ++j) {
            float d = dist3(L, A, B, unc[j]);
            bool after = d > last || (d == last && j > last_j);
            bool take = after && (d < best || best_j < 0);
            best = take ? d : best;
            best_j = take ? j : best_j;
        }
        if (best_j < 0) return LoopAcc{n0, n1, n2, den, 1};  // NaNs: let the exact kernel decide
        if (pass == kp) {
            float gap = best - tau;
            return LoopAcc{n0, n1, n2, den, gap * gap <= FAST_TIE_TOL * best ? 1 : 0};
        }
        float4 p = unc[best_j];
        float w = fast_weight<KIND>(best, c1, c0);
        n0 = fmaf(p.x, w, n0);
        n1 = fmaf(p.y, w, n1);
        n2 = fmaf(p.z, w, n2);
        den += w;
        last = best; last_j = best_j; tau = best;
    }
    return LoopAcc{n0, n1, n2, den, 0};
}

// Many uncertain colours, moderate kp (<= 32): one pass keeping the kp + 1 best (distance, index)
// pairs sorted in a small per-thread array (local memory, L1 resident); about 10 instructions per
// colour plus the occasional insertion, instead of kp + 1 full passes.
template <int KIND>
__device__ __noinline__ LoopAcc select_accumulate_insert(const float4* __restrict__ unc, int mU, int kp, float L, float A, float B, float c1,
                                                         float c0) {
    constexpr int CAP = 33;
    float td[CAP];
    unsigned short tj[CAP];
    const int keep = kp + 1;  // the extra one decides whether the k-boundary is a near-tie
    int cnt = 0;
    for (int j = 0; j < mU; ++j) {
        float4 pj = unc[j];
        // the list is sorted by LB (in .w): nothing from here on can beat the current (kp+1)-th best
        if (cnt == keep && pj.w >= td[keep - 1]) break;
        float d = dist3(L, A, B, pj);
        if (cnt < keep || d < td[cnt - 1]) {
            int i = cnt < keep ? cnt : keep - 1;
            while (i > 0 && td[i - 1] > d) {
                td[i] = td[i - 1];
                tj[i] = tj[i - 1];
                --i;
            }
            td[i] = d;
            tj[i] = (unsigned short)j;
            if (cnt < keep) ++cnt;
        }
    }
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
    if (cnt < keep) return LoopAcc{n0, n1, n2, den, 1};  // NaNs swallowed candidates: exact kernel decides
    for (int i = 0; i < kp; ++i) {
        float4 p = unc[tj[i]];
        float w = fast_weight<KIND>(td[i], c1, c0);
        n0 = fmaf(p.x, w, n0);
        n1 = fmaf(p.y, w, n1);
        n2 = fmaf(p.z, w, n2);
        den += w;
    }
    float gap = td[kp] - td[kp - 1];
    return LoopAcc{n0, n1, n2, den, gap * gap <= FAST_TIE_TOL * td[kp] ? 1 : 0};
}

// Dynamic shared memory (bytes): pal[n] float4 | cand[n] float4 | lb[n] f32 | ub[n] f32 |
// cand_idx[n] u32 | lab[3][tile] f32 | rgb[tile] u32
__host__ __device__ inline size_t fast_smem_bytes(uint32_t n, uint32_t tile) {
    return (size_t)n * (16 + 16 + 4 + 4 + 4) + (size_t)tile * 16 + 64;
}

// Negligibility test of step 4 for a colour whose distance is at least `lb`, against the tile's
// guaranteed-nearest distance `ubmin` (log2 of it in lg_ubmin for Shepard).
template <int KIND>
__device__ __forceinline__ bool negligible_at(float lb, float ubmin, float lg_ubmin, float c1) {
    if (KIND == 0) return c1 < 0.f && c1 * (lb - ubmin) < FAST_NEGLIGIBLE_LOG2;
    if (KIND == 1) return c1 < 0.f && lb > 0.f && ubmin > 0.f && c1 * (lg2_approx(lb) - lg_ubmin) < FAST_NEGLIGIBLE_LOG2 - 1e-3f;
    return false;
}

// Outcome of steps 2-4 for one tile, published through shared memory.
struct TileClass {
    uint32_t nin, nunc;          // list lengths (negligible colours dropped)
    uint32_t nin_all, nunc_all;  // class sizes
    uint32_t nhit, take_all;
    float lbmin, ubmin;
};

__device__ __forceinline__ void bounds_of(float4 p, const float (&lo)[3], const float (&hi)[3], float& lbv, float& ubv) {
    float pv[3] = {p.x, p.y, p.z};
    lbv = 0.f;
    ubv = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float below = lo[c] - pv[c], above = pv[c] - hi[c];
        float ex = fmaxf(0.f, fmaxf(below, above));
        float fx = fmaxf(fabsf(pv[c] - lo[c]), fabsf(pv[c] - hi[c]));
        lbv = fmaf(ex, ex, lbv);
        ubv = fmaf(fx, fx, ubv);
    }
    lbv *= (1.0f - 2e-6f);
    ubv *= (1.0f + 2e-6f);
}

// Steps 2-4 for palettes of at most 32 colours (the common case: the built-in palettes average 22):
// one warp, one colour per lane, ranks by shuffles, lists by ballots -- no shared-memory atomics
// and a single CTA barrier afterwards, while the other seven warps wait.
template <int KIND>
__device__ __forceinline__ void classify_one_warp(uint32_t lane, uint32_t n, uint32_t k, float c1, const float (&lo)[3], const float (&hi)[3],
                                                  const uint32_t* s_box, const float4* s_pal, const uint32_t* pal_rgb, float4* s_cand,
                                                  uint32_t* s_cidx, uint32_t* s_hit, TileClass* out) {
    const uint32_t FULL = 0xffffffffu;
    const bool active = lane < n;
    float4 p = active ? s_pal[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    float lbv, ubv;
    bounds_of(p, lo, hi, lbv, ubv);
    if (!active) { lbv = INFINITY; ubv = INFINITY; }
    const float lbmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(lbv)));
    const float ubmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ubv)));
    uint32_t nhit = 0;
    if (KIND != KIND_NN) {
        uint32_t q = active ? pal_rgb[lane] : 0u, qr = q & 255u, qg = (q >> 8) & 255u, qb = (q >> 16) & 255u;
        bool inside = active && qr >= s_box[6] && qr <= s_box[9] && qg >= s_box[7] && qg <= s_box[10] && qb >= s_box[8] && qb <= s_box[11];
        uint32_t bh = __ballot_sync(FULL, inside);
        nhit = __popc(bh);
        uint32_t pos = __popc(bh & ((1u << lane) - 1u));
        if (inside && pos < FAST_MAX_HITS) s_hit[pos] = q;
    }
    uint32_t rank_u = 0, rank_l = 0, rivals = 0;
    if (k != 0) {
        for (uint32_t q = 0; q < n; ++q) {
            float uq = __shfl_sync(FULL, ubv, q), lq = __shfl_sync(FULL, lbv, q);
            rank_u += (uq < ubv || (uq == ubv && q < lane)) ? 1u : 0u;
            rank_l += (lq < lbv || (lq == lbv && q < lane)) ? 1u : 0u;
            rivals += (q != lane && lq <= ubv) ? 1u : 0u;
        }
    }
    float uk = INFINITY, lk = INFINITY;
    if (k != 0) {
        uint32_t bu = __ballot_sync(FULL, active && rank_u == k - 1), bl = __ballot_sync(FULL, active && rank_l == k - 1);
        if (bu) uk = __shfl_sync(FULL, ubv, __ffs(bu) - 1);
        if (bl) lk = __shfl_sync(FULL, lbv, __ffs(bl) - 1);
    }
    const float lg_ubmin = KIND == 1 ? lg2_approx(ubmin) : 0.f;
    const bool moot = (k != 0) && (KIND != KIND_NN) && negligible_at<KIND>(lk, ubmin, lg_ubmin, c1);
    const bool take_all = (k == 0) || moot;
    bool is_in = false, is_unc = false, negl = false;
    if (active) {
        if (take_all) is_in = true;
        else if (lbv <= uk) { is_in = rivals <= k - 1; is_unc = !is_in; }
        if (KIND != KIND_NN) negl = negligible_at<KIND>(lbv, ubmin, lg_ubmin, c1);
    }
    const uint32_t nin_all = __popc(__ballot_sync(FULL, is_in)), nunc_all = __popc(__ballot_sync(FULL, is_unc));
    const bool list_in = is_in && !negl, list_unc = is_unc && !negl;
    const uint32_t bin = __ballot_sync(FULL, list_in), bun = __ballot_sync(FULL, list_unc);
    const uint32_t nin = __popc(bin), nunc = __popc(bun), lt = (1u << lane) - 1u;
    // IN colours first, UNCERTAIN ones after them, palette order in both; s_cidx mirrors the kernel's
    // general layout (IN from the front, UNCERTAIN from the back)
    if (list_in) { uint32_t pos = __popc(bin & lt); s_cand[pos] = p; s_cidx[pos] = lane; }
    if (KIND != KIND_NN && nunc > FAST_BUCKET_MAX) {
        // large uncertain sets are listed by ascending LB (kept in .w) so the insertion selection can
        // stop at the first colour whose LB already exceeds the thread's current k-th distance
        uint32_t rank = 0;
        for (uint32_t q = 0; q < 32; ++q) {
            float lq = __shfl_sync(FULL, lbv, q);
            rank += (((bun >> q) & 1u) && (lq < lbv || (lq == lbv && q < lane))) ? 1u : 0u;
        }
        if (list_unc) { p.w = lbv; s_cand[nin + rank] = p; s_cidx[n - 1 - rank] = lane; }
    } else if (list_unc) {
        uint32_t pos = __popc(bun & lt);
        p.w = lbv;
        s_cand[nin + pos] = p;
        s_cidx[n - 1 - pos] = lane;
    }
    if (lane == 0) {
        out->nin = nin; out->nunc = nunc; out->nin_all = nin_all; out->nunc_all = nunc_all;
        out->nhit = nhit; out->take_all = take_all ? 1u : 0u; out->lbmin = lbmin; out->ubmin = ubmin;
    }
}

// Persistent CTAs: the grid is a few CTAs per SM; each stages the colour tables and the palette
// once and then walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int KIND, int MODE>
__global__ void __launch_bounds__(FAST_THREADS, 4) k_remap_fast(const FastArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmemTables st;
    __shared__ uint32_t s_box[12];        // Oklab min keys [0..3), max keys [3..6), sRGB min [6..9), max [9..12)
    __shared__ uint32_t s_lbmin, s_ubmin, s_uk, s_lk;
    __shared__ uint32_t s_wcnt[2][FAST_THREADS / 32];
    __shared__ uint32_t s_nhit, s_next, s_geo[3];
    __shared__ TileClass s_cls;
    __shared__ uint32_t s_hit[FAST_MAX_HITS];
    __shared__ uint32_t s_chan[3][16];    // identity channel value of each brick coordinate

    const uint32_t n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = a.ept * FAST_THREADS;
    float4* s_pal = reinterpret_cast<float4*>(smem_raw);
    float4* s_cand = s_pal + n;
    float* s_lb = reinterpret_cast<float*>(s_cand + n);
    float* s_ub = s_lb + n;
    uint32_t* s_cidx = reinterpret_cast<uint32_t*>(s_ub + n);
    float* s_lab = reinterpret_cast<float*>(s_cidx + n);
    uint32_t* s_rgb = reinterpret_cast<uint32_t*>(s_lab + 3 * tile);
    // scratch for the sorted bound arrays of the general classification: the candidate buffer,
    // which is only filled afterwards (2 * n_pad floats <= n float4 because n_pad < 2n)
    uint32_t n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    float* srt_u = reinterpret_cast<float*>(s_cand);
    float* srt_l = srt_u + n_pad;

    const uint32_t cs = a.f.cs, csm1 = a.f.csm1;
    const uint32_t begin = (uint32_t)a.f.begin, end = (uint32_t)(a.f.begin + a.f.count);  // < 2^32 by construction
    const uint32_t lr = tid & (a.tr - 1), lg = (tid >> a.sh_r) & (a.tg - 1);
    const uint32_t lb_step = FAST_THREADS >> (a.sh_r + a.sh_g), lb0 = tid >> (a.sh_r + a.sh_g);
    const uint32_t k = a.k;
    const float c1 = a.c1;

    for (uint32_t i = tid; i < n; i += FAST_THREADS) s_pal[i] = a.pal[i];
    stage_tables(st, a.tables);  // ends with __syncthreads()

    // Tiles differ a lot in cost (0 to 30 uncertain colours), so they are handed out dynamically:
    // the first wave is blockIdx.x, later ones come from a device-wide counter.
    for (uint32_t bid = blockIdx.x; bid < a.ntiles;) {
        __syncthreads();  // previous tile's shared state (and s_next) is no longer read
        if (tid == 0) s_next = gridDim.x + atomicAdd(a.tile_counter, 1u);
        // ---- tile geometry (the divisions run in the first two warps only) -----------------
        uint32_t r0 = 0, g0 = 0, b0 = 0, img_base = 0;
        if (MODE == FRONT_IMAGE) {
            img_base = bid * tile;
        } else if (tid < 64) {
            uint32_t ir = bid % a.tiles_r, ig = (bid / a.tiles_r) % a.tiles_g, ib = bid / (a.tiles_r * a.tiles_g);
            r0 = ir << a.sh_r; g0 = ig << a.sh_g; b0 = a.b_lo + ib * a.tb;
            if (tid == 0) { s_geo[0] = r0; s_geo[1] = g0; s_geo[2] = b0; }
            if (tid < 48) {
                uint32_t c = tid >> 4, i = tid & 15;
                uint32_t v = min((c == 0 ? r0 : c == 1 ? g0 : b0) + i, csm1);
                s_chan[c][i] = MODE == FRONT_TABLE ? v : hald_channel(v, csm1);
            }
        }
        if (tid >= 64 && tid < 67) { s_box[tid - 64] = 0xffffffffu; s_box[6 + tid - 64] = 0xffffffffu; }
        else if (tid >= 67 && tid < 70) { s_box[tid - 64] = 0u; s_box[6 + tid - 64] = 0u; }
        if (tid == 96) { s_lbmin = 0x7f800000u; s_ubmin = 0x7f800000u; s_uk = 0x7f800000u; s_lk = 0x7f800000u; s_nhit = 0; }
        __syncthreads();
        if (MODE != FRONT_IMAGE) {
            r0 = s_geo[0]; g0 = s_geo[1]; b0 = s_geo[2];
            // whole tile outside the requested entry range: nothing to do (uniform)
            uint32_t first = (b0 * cs + g0) * cs + r0;
            uint32_t last = (min(b0 + a.tb - 1, csm1) * cs + min(g0 + a.tg - 1, csm1)) * cs + min(r0 + a.tr - 1, csm1);
            if (b0 > csm1 || last < begin || first >= end) { bid = s_next; continue; }
        }

        // ---- 1. colours of this tile -> Oklab, bounding box ----------------------------
        uint32_t kmin[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, kmax[3] = {0u, 0u, 0u};
        uint32_t cmin[3] = {255u, 255u, 255u}, cmax[3] = {0u, 0u, 0u};
        uint32_t valid_mask = 0;
        for (uint32_t e = 0; e < a.ept; ++e) {
            uint32_t li = e * FAST_THREADS + tid;
            uint32_t R, G, B;
            // `valid`: this thread must produce the colour; `in_tile`: the colour belongs to the
            // tile's geometry. The bounding box is taken over in_tile so that the classification --
            // hence every byte of the result -- does not depend on how the rows were sharded.
            bool valid, in_tile;
            if (MODE == FRONT_IMAGE) {
                uint32_t pix = img_base + li;
                valid = in_tile = pix < end;
                uchar4 v = valid ? reinterpret_cast<const uchar4*>(a.f.out)[pix] : make_uchar4(0, 0, 0, 0);
                R = v.x; G = v.y; B = v.z;
            } else {
                uint32_t lb = lb0 + e * lb_step;
                uint32_t r = r0 + lr, g = g0 + lg, b = b0 + lb;
                uint32_t flat = (b * cs + g) * cs + r;
                in_tile = r < cs && g < cs && b < cs && lb < a.tb;
                valid = in_tile && flat >= begin && flat < end;
                R = s_chan[0][lr]; G = s_chan[1][lg]; B = s_chan[2][lb & 15];
            }
            Lab c = linear_to_oklab_fast(st.decode[R], st.decode[G], st.decode[B]);
            float Lp = c.l * a.lum;
            s_lab[li] = Lp;
            s_lab[tile + li] = c.a;
            s_lab[2 * tile + li] = c.b;
            s_rgb[li] = R | (G << 8) | (B << 16);
            if (valid) valid_mask |= 1u << e;
            if (in_tile) {
                uint32_t k0 = f2key(Lp), k1 = f2key(c.a), k2 = f2key(c.b);
                kmin[0] = min(kmin[0], k0); kmax[0] = max(kmax[0], k0);
                kmin[1] = min(kmin[1], k1); kmax[1] = max(kmax[1], k1);
                kmin[2] = min(kmin[2], k2); kmax[2] = max(kmax[2], k2);
                cmin[0] = min(cmin[0], R); cmax[0] = max(cmax[0], R);
                cmin[1] = min(cmin[1], G); cmax[1] = max(cmax[1], G);
                cmin[2] = min(cmin[2], B); cmax[2] = max(cmax[2], B);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t lo = __reduce_min_sync(0xffffffffu, kmin[c]), hi = __reduce_max_sync(0xffffffffu, kmax[c]);
            uint32_t clo = __reduce_min_sync(0xffffffffu, cmin[c]), chi = __reduce_max_sync(0xffffffffu, cmax[c]);
            if (lane == 0) {
                atomicMin(&s_box[c], lo); atomicMax(&s_box[3 + c], hi);
                atomicMin(&s_box[6 + c], clo); atomicMax(&s_box[9 + c], chi);
            }
        }
        __syncthreads();
        const uint32_t next_bid = s_next;
        if (s_box[0] == 0xffffffffu) { bid = next_bid; continue; }  // no colour in this tile (uniform)

        // ---- 2. LB / UB of every palette colour against the inflated box ----------------
        const float PAD = 2e-5f;
        float lo[3], hi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = key2f(s_box[c]) - PAD; hi[c] = key2f(s_box[3 + c]) + PAD; }
        uint32_t nin = 0, nunc = 0;          // list lengths (negligible colours dropped)
        uint32_t nin_all = 0, nunc_all = 0;  // class sizes
        uint32_t nhit_raw = 0;
        bool take_all = false;
        float lbmin = 0.f, ubmin = 0.f;
        if (n <= 32) {
            if (warp == 0) classify_one_warp<KIND>(lane, n, k, c1, lo, hi, s_box, s_pal, a.pal_rgb, s_cand, s_cidx, s_hit, &s_cls);
            __syncthreads();
            nin = s_cls.nin; nunc = s_cls.nunc; nin_all = s_cls.nin_all; nunc_all = s_cls.nunc_all;
            nhit_raw = s_cls.nhit; take_all = s_cls.take_all != 0; lbmin = s_cls.lbmin; ubmin = s_cls.ubmin;
        } else {
        for (uint32_t t = tid; t < n; t += FAST_THREADS) {
            float lbv, ubv;
            bounds_of(s_pal[t], lo, hi, lbv, ubv);
            s_lb[t] = lbv;
            s_ub[t] = ubv;
            atomicMin(&s_lbmin, __float_as_uint(lbv));
            atomicMin(&s_ubmin, __float_as_uint(ubv));
            // palette colours whose sRGB lies inside the tile's sRGB box can be hit exactly (rbf.rs:66)
            if (KIND != KIND_NN) {
                uint32_t q = a.pal_rgb[t], qr = q & 255u, qg = (q >> 8) & 255u, qb = (q >> 16) & 255u;
                if (qr >= s_box[6] && qr <= s_box[9] && qg >= s_box[7] && qg <= s_box[10] && qb >= s_box[8] && qb <= s_box[11]) {
                    uint32_t h = atomicAdd(&s_nhit, 1u);
                    if (h < FAST_MAX_HITS) s_hit[h] = q;
                }
            }
        }
        __syncthreads();

        // ---- 3./4. classification --------------------------------------------------------
        ubmin = __uint_as_float(s_ubmin);
        lbmin = __uint_as_float(s_lbmin);
        const float lg_ubmin = KIND == 1 ? lg2_approx(ubmin) : 0.f;
        if (k != 0) {
            // U_k: k-th smallest UB (every pixel's k-th neighbour is at most this far);
            // L_k: k-th smallest LB (... and at least this far)
            // Sorted copies of both bound arrays (CTA-wide bitonic network in the not-yet-used
            // candidate buffer): O(n log^2 n / 256) per thread instead of O(n) rank loops.
            for (uint32_t i = tid; i < n_pad; i += FAST_THREADS) {
                srt_u[i] = i < n ? s_ub[i] : INFINITY;
                srt_l[i] = i < n ? s_lb[i] : INFINITY;
            }
            __syncthreads();
            for (uint32_t size = 2; size <= n_pad; size <<= 1) {
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t i = tid; i < (n_pad >> 1); i += FAST_THREADS) {
                        uint32_t lo_i = ((i & ~(stride - 1)) << 1) | (i & (stride - 1)), hi_i = lo_i + stride;
                        bool up = (lo_i & size) == 0;
                        float a0 = srt_u[lo_i], a1 = srt_u[hi_i], b0v = srt_l[lo_i], b1v = srt_l[hi_i];
                        float amin = fminf(a0, a1), amax = fmaxf(a0, a1), bmin = fminf(b0v, b1v), bmax = fmaxf(b0v, b1v);
                        srt_u[lo_i] = up ? amin : amax; srt_u[hi_i] = up ? amax : amin;
                        srt_l[lo_i] = up ? bmin : bmax; srt_l[hi_i] = up ? bmax : bmin;
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) { s_uk = __float_as_uint(srt_u[k - 1]); s_lk = __float_as_uint(srt_l[k - 1]); }
            __syncthreads();
        }
        const float uk = __uint_as_float(s_uk);
        // If even the k-th neighbour's weight is negligible everywhere in the tile, "nearest k" no
        // longer constrains anything: every colour that can matter is taken, nothing is chosen.
        const bool moot = (k != 0) && (KIND != KIND_NN) && negligible_at<KIND>(__uint_as_float(s_lk), ubmin, lg_ubmin, c1);
        take_all = (k == 0) || moot;
        for (uint32_t t0 = 0; t0 < n; t0 += FAST_THREADS) {
            uint32_t t = t0 + tid;
            bool is_in = false, is_unc = false, negl = false;
            if (t < n) {
                float l = s_lb[t], u = s_ub[t];
                if (take_all) {
                    is_in = true;
                } else if (l <= uk) {
                    // rivals = #{q != t : LB(q) <= UB(t)} = upper_bound(sorted LB, UB(t)) - 1 (t counts itself)
                    uint32_t lo_b = 0, hi_b = n;
                    while (lo_b < hi_b) {
                        uint32_t mid = (lo_b + hi_b) >> 1;
                        if (srt_l[mid] <= u) lo_b = mid + 1; else hi_b = mid;
                    }
                    uint32_t rivals = lo_b - 1;
                    is_in = rivals <= k - 1;
                    is_unc = !is_in;
                }
                if (KIND != KIND_NN) negl = negligible_at<KIND>(l, ubmin, lg_ubmin, c1);
            }
            nin_all += __syncthreads_count(is_in);
            nunc_all += __syncthreads_count(is_unc);
            bool list_in = is_in && !negl, list_unc = is_unc && !negl;
            uint32_t bin = __ballot_sync(0xffffffffu, list_in), bun = __ballot_sync(0xffffffffu, list_unc);
            if (lane == 0) { s_wcnt[0][warp] = __popc(bin); s_wcnt[1][warp] = __popc(bun); }
            __syncthreads();
            uint32_t pre_in = nin, pre_unc = nunc, tot_in = 0, tot_unc = 0;
#pragma unroll
            for (int w = 0; w < FAST_THREADS / 32; ++w) {
                uint32_t ci = s_wcnt[0][w], cu = s_wcnt[1][w];
                if ((uint32_t)w < warp) { pre_in += ci; pre_unc += cu; }
                tot_in += ci; tot_unc += cu;
            }
            uint32_t lt = (1u << lane) - 1u;
            // IN colours fill s_cidx from the front, UNCERTAIN ones from the back (palette order both)
            if (list_in) s_cidx[pre_in + __popc(bin & lt)] = t;
            if (list_unc) s_cidx[n - 1 - (pre_unc + __popc(bun & lt))] = t;
            nin += tot_in; nunc += tot_unc;
            __syncthreads();
        }
        for (uint32_t j = tid; j < nin; j += FAST_THREADS) s_cand[j] = s_pal[s_cidx[j]];
        for (uint32_t j = tid; j < nunc; j += FAST_THREADS) {
            uint32_t t = s_cidx[n - 1 - j];
            float4 p = s_pal[t];
            p.w = s_lb[t];
            uint32_t pos = j;
            if (KIND != KIND_NN && nunc > FAST_BUCKET_MAX) {  // ascending LB, see classify_one_warp
                pos = 0;
                for (uint32_t i = 0; i < nunc; ++i) {
                    float li = s_lb[s_cidx[n - 1 - i]];
                    pos += (li < p.w || (li == p.w && i < j)) ? 1u : 0u;
                }
            }
            s_cand[nin + pos] = p;
        }
        __syncthreads();
        nhit_raw = s_nhit;
        }  // general (n > 32) classification
        if (a.stats && tid == 0) {
            atomicAdd(&a.stats[min(nunc, 63u)], 1u);
            atomicAdd(&a.stats[64 + min(nin, 63u)], 1u);
        }
        // neighbours still to choose per pixel among the uncertain colours
        const int kp = (int)k - (int)nin_all;
        const bool choose = !take_all && kp > 0 && (int)nunc_all > kp && (int)nunc > kp;
        const uint32_t direct = (take_all || kp <= 0 || choose) ? nin : nin + nunc;  // listed colours taken unconditionally
        const uint32_t nhit = min(nhit_raw, (uint32_t)FAST_MAX_HITS);
        const bool hits_overflow = nhit_raw > FAST_MAX_HITS;

        // weight shift
        float c0 = 0.f;
        if (KIND == 0) c0 = -c1 * lbmin;
        if (KIND == 1) c0 = -c1 * lg2_approx(lbmin > 0.f ? lbmin : ubmin);

        // ---- 5. per colour ----------------------------------------------------------------
        for (uint32_t e = 0; e < a.ept; ++e) {
            if (!((valid_mask >> e) & 1u)) continue;
            uint32_t li = e * FAST_THREADS + tid;
            float L = s_lab[li], A = s_lab[tile + li], B = s_lab[2 * tile + li];
            uint32_t packed = s_rgb[li];
            PixelIO px;
            px.r = packed & 255u; px.g = (packed >> 8) & 255u; px.b = packed >> 16;
            if (MODE == FRONT_IMAGE) {
                px.slot = img_base + li; px.mode = FRONT_IMAGE;
            } else {
                uint32_t lb = lb0 + e * lb_step;
                px.slot = ((b0 + lb) * cs + (g0 + lg)) * cs + (r0 + lr); px.mode = MODE;
            }
            if (hits_overflow) { flag_entry(a, px.slot); continue; }

            if (KIND == KIND_NN) {
                // nearest_one: IN holds at most one colour; otherwise argmin over the uncertain ones
                uint32_t best_j = 0;
                bool risky = false;
                if (nin == 0) {
                    float best = INFINITY, second = INFINITY;
                    for (uint32_t j = 0; j < nunc; ++j) {
                        float d = dist3(L, A, B, s_cand[j]);
                        bool better = d < best;
                        second = better ? best : fminf(second, d);
                        best_j = better ? j : best_j;
                        best = better ? d : best;
                    }
                    float gap = second - best;
                    risky = (gap * gap <= FAST_TIE_TOL * second) && (second < INFINITY);
                }
                if (risky) { flag_entry(a, px.slot); continue; }
                uint32_t idx = (nin != 0) ? s_cidx[0] : s_cidx[n - 1 - best_j];
                if (!a.preserve) {
                    uint32_t q = a.pal_rgb[idx];
                    front_store<MODE>(a.f, px, q & 255u, (q >> 8) & 255u, (q >> 16) & 255u);
                } else {
                    // nearest_neighbor.rs:57-59: pixel's own (unscaled, exact) l with the colour's a, b
                    Lab ex = linear_to_oklab_exact(st.decode[px.r], st.decode[px.g], st.decode[px.b]);
                    Lab o = {ex.l, a.pal_ab[2 * idx], a.pal_ab[2 * idx + 1]};
                    float r, g, b;
                    oklab_to_linear_exact(o, r, g, b);
                    front_store<MODE>(a.f, px, f32_to_srgb8(r, st.encode2), f32_to_srgb8(g, st.encode2), f32_to_srgb8(b, st.encode2));
                }
                continue;
            }

            // rbf.rs:66-68 palette.contains(&color): identical sRGB <=> identical Oklab triple
            bool hit = false;
            for (uint32_t h = 0; h < nhit; ++h) hit |= (s_hit[h] == packed);
            if (hit) {
                if (MODE != FRONT_IMAGE) front_store<MODE>(a.f, px, px.r, px.g, px.b);
                continue;
            }
            float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
#pragma unroll 4
            for (uint32_t j = 0; j < direct; ++j) {
                float4 p = s_cand[j];
                float w = fast_weight<KIND>(dist3(L, A, B, p), c1, c0);
                n0 = fmaf(p.x, w, n0);
                n1 = fmaf(p.y, w, n1);
                n2 = fmaf(p.z, w, n2);
                den += w;
            }
            bool risky = false;
            if (choose) {
                const float4* unc = s_cand + nin;
                if (nunc <= 4) risky = select_accumulate<4, KIND>(unc, (int)nunc, kp, L, A, B, c1, c0, n0, n1, n2, den);
                else if (nunc <= 8) risky = select_accumulate<8, KIND>(unc, (int)nunc, kp, L, A, B, c1, c0, n0, n1, n2, den);
                else if (FAST_BUCKET_MAX >= 16 && nunc <= 16) risky = select_accumulate<16, KIND>(unc, (int)nunc, kp, L, A, B, c1, c0, n0, n1, n2, den);
                else {
                    LoopAcc la = kp <= 32 ? select_accumulate_insert<KIND>(unc, (int)nunc, kp, L, A, B, c1, c0)
                                          : select_accumulate_loop<KIND>(unc, (int)nunc, kp, L, A, B, c1, c0);
                    n0 += la.n0; n1 += la.n1; n2 += la.n2; den += la.den;
                    risky = la.risky != 0;
                }
            }
            // The reference's f64 sums have their own failure modes (every exp() underflowing ->
            // 0/0 -> black; 1/sqrt(d)^p overflowing -> inf/inf -> black). log2 of the true denominator
            // is lg2(den) - c0; near either end of the f64 range the exact kernel reproduces them.
            if (KIND != 2 && fabsf(lg2_approx(den) - c0) > 1000.0f) risky = true;
            if (risky || !(den > 1e-30f) || !(den < 1e30f)) { flag_entry(a, px.slot); continue; }
            float inv = __frcp_rn(den);
            Lab o;
            o.l = a.preserve ? L * a.inv_lum : n0 * inv * a.inv_lum;
            o.a = n1 * inv;
            o.b = n2 * inv;
            float r, g, b;
            oklab_to_linear_fast(o, r, g, b);
            front_store<MODE>(a.f, px, f32_to_srgb8(r, st.encode2), f32_to_srgb8(g, st.encode2), f32_to_srgb8(b, st.encode2));
        }
        bid = next_bid;
    }
}

// Diagnostics: the fast f32 Oklab conversion for arrays (tests bound its distance to the exact one).
// Colours 2i and 2i + 1 go through the packed conversion of the warp-tile and blur kernels, in the form
// those kernels use it (partial sums over r and g, then b); an odd tail through the scalar one.
__global__ void k_srgb_to_oklab_fast(const uint8_t* __restrict__ rgb, size_t n, float* __restrict__ out, const ColorTables* __restrict__ tables) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    if (i + 1 >= n) {
        Lab o = linear_to_oklab_fast(tables->decode[rgb[3 * i]], tables->decode[rgb[3 * i + 1]], tables->decode[rgb[3 * i + 2]]);
        out[3 * i] = o.l; out[3 * i + 1] = o.a; out[3 * i + 2] = o.b;
        return;
    }
    const f2 r = make_float2(tables->decode[rgb[3 * i]], tables->decode[rgb[3 * i + 3]]);
    const f2 g = make_float2(tables->decode[rgb[3 * i + 1]], tables->decode[rgb[3 * i + 4]]);
    const f2 b = make_float2(tables->decode[rgb[3 * i + 2]], tables->decode[rgb[3 * i + 5]]);
    const f2 pl = fma2(splat(0.5363325363f), g, fma2(splat(0.4122214708f), r, splat(LMS_FLOOR)));
    const f2 pm = fma2(splat(0.6806995451f), g, fma2(splat(0.2119034982f), r, splat(LMS_FLOOR)));
    const f2 ps = fma2(splat(0.2817188376f), g, fma2(splat(0.0883024619f), r, splat(LMS_FLOOR)));
    f2 L, A, B;
    lms_to_lab2(fma2(splat(0.0514459929f), b, pl), fma2(splat(0.1073969566f), b, pm), fma2(splat(0.6299787005f), b, ps), 1.0f, L, A, B);
    out[3 * i] = L.x; out[3 * i + 1] = A.x; out[3 * i + 2] = B.x;
    out[3 * i + 3] = L.y; out[3 * i + 4] = A.y; out[3 * i + 5] = B.y;
}

}  // namespace lutgen
