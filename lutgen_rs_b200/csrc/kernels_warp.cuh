// kernels_warp.cuh -- the throughput remap kernel for palettes of at most 32 colours (the
// built-in palettes average 22): the tile-bounded nearest-k of kernels_fast.cuh with one WARP per
// tile instead of one CTA, and the arithmetic in packed f32x2 (FFMA2 / FMUL2 / FADD2, sm_100a).
//
// Why a warp per tile: the smaller the tile, the tighter its Oklab bounding box and the fewer
// palette colours are left UNCERTAIN at the k-boundary (tests/tools/sim_tiles.py: 6.9 per 8x8x16 CTA tile,
// 2.7 per 8x4x4 warp tile for the headline configuration); per-entry selection among the uncertain
// colours was the most expensive part of the CTA kernel. A warp classifies a 26-colour palette in
// ~200 instructions (one colour per lane, ranks from one broadcast loop), needs no CTA barrier,
// keeps its entries' Oklab values in registers, and takes tiles from a device-wide counter.
//
// Why f32x2: the kernel is bound by instruction ISSUE, not by the FMA pipe. A lane owns EPT entries
// of its tile; two of them share every instruction as a float2 (the palette colour is the scalar
// broadcast operand FFMA2 accepts), which halves the issue slots of all the per-entry arithmetic.
//
//   tile        : 8 (r) x 4 (g) x EPT (b) Hald-identity entries (LUT / TABLE front-ends; lane =
//                 (r, g), entry e = b plane) or 32*EPT consecutive pixels (IMAGE).
//   classify    : as kernels_fast.cuh steps 2-4 (LB/UB against the inflated box, U_k, IN / OUT /
//                 UNCERTAIN, negligible colours dropped, "moot" tiles take every colour).
//   accumulate  : the candidate list is walked once for the lane's EPT entries together (one
//                 broadcast LDS.128 serves all of them).
//   Gaussian    : the list holds q = (-2C p, C(|p|^2 - lbmin)), C = shape*log2(e), so that
//                 s = C(d - lbmin) = q.xyz . c + q.w + C|c|^2 costs 3 FFMA2 + 1 FADD2 per pair,
//                 w = ex2(-s), and the sums run over q.xyz (-2C is divided out once per entry).
//   NN          : the same expansion ranks by d - |c|^2.
//   near-ties   : as in kernels_fast.cuh every decision closer than the f32 error bound (here a
//                 per-tile constant) sends the entry to the exact f64 kernel's work list.
//
// Code size matters: warps run free of each other, so the SM's instruction cache sees the whole
// kernel body at once. Everything rare or repeated per entry (selection networks, flagging) is out
// of line, and loops that would unroll into kilobytes are kept rolled.
#pragma once
#include "kernels_fast.cuh"

namespace lutgen {

#ifndef WT_MINB
#define WT_MINB 3
#endif
constexpr int WT_THREADS = 256;
constexpr int WT_WARPS = WT_THREADS / 32;
constexpr int WT_BATCH = 4;          // warp tiles per hand-out
constexpr int WT_MAX_N = 32;         // one palette colour per lane
constexpr int WT_MAX_CS = 1089;      // level 33
constexpr float WT_TIE_SLOPE = 2.24e-5f;   // sqrt(FAST_TIE_TOL): conversion error, scales with sqrt(d)
constexpr float WT_TIE_CONST = 1.5e-6f;    // rounding of the expanded distance (|p|^2 - 2 p.c + |c|^2)

struct WarpArgs {
    Front f;
    const float4* pal;        // n x (l*lum, a, b, 0)
    const uint32_t* pal_rgb;  // n x packed r | g<<8 | b<<16
    const float* pal_ab;      // n x 2 (NearestNeighbor preserve)
    uint32_t n, k;            // k = 0: every colour
    int preserve;
    float lum, inv_lum;
    float c1;                 // weight slope (kernels_fast.cuh header)
    uint32_t tiles_r, tiles_g;
    float inv_tiles_r, inv_tiles_g;
    uint32_t b_lo;            // first blue plane covered (multiple of EPT)
    uint32_t ntiles;          // tiles this launch walks (local numbering)
    uint32_t ntiles_global;   // tiles of the whole range
    uint32_t shard_index, shard_count;  // this launch takes hand-outs index, index + count, ... (1 hand-out = `batch` warp tiles / 1 CTA tile)
    uint32_t batch;           // warp tiles per hand-out: WT_BATCH, or 1 when the range has too few tiles to fill the GPU otherwise
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t* tile_counter;   // zeroed before the launch
    uint32_t* stats;
    const ColorTables* tables;
    int words_ok;             // FRONT_LUT, cs % 4 == 0: 1 = destinations 4-byte aligned (whole tiles leave as words),
                              // 2 = 16-byte aligned (whole hand-outs leave as 16-byte pieces), 0 = byte stores
};

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void oklab_to_linear2(f2 L, f2 A, f2 B, f2& r, f2& g, f2& b) {
    f2 l_ = fma2(splat(0.3963377774f), A, fma2(splat(0.2158037573f), B, L));
    f2 m_ = fma2(splat(-0.1055613458f), A, fma2(splat(-0.0638541728f), B, L));
    f2 s_ = fma2(splat(-0.0894841775f), A, fma2(splat(-1.2914855480f), B, L));
    f2 l = mul2(mul2(l_, l_), l_), m = mul2(mul2(m_, m_), m_), s = mul2(mul2(s_, s_), s_);
    r = fma2(splat(4.0767416621f), l, fma2(splat(-3.3077115913f), m, mul2(splat(0.2309699292f), s)));
    g = fma2(splat(-1.2684380046f), l, fma2(splat(2.6097574011f), m, mul2(splat(-0.3413193965f), s)));
    b = fma2(splat(-0.0041960863f), l, fma2(splat(-0.7034186147f), m, mul2(splat(1.7076147010f), s)));
}

// Score of candidate q for two entries (smaller = nearer):
//   Gaussian   C (d - lbmin)  = q.xyz . c + (q.w + C |c|^2)            (cc = C |c|^2)
//   NN         d - |c|^2      = q.xyz . c + q.w
//   Shepard / Linear  d       (differences: small distances keep their relative precision)
template <int KIND>
__device__ __forceinline__ f2 wt_score2(float4 q, f2 L, f2 A, f2 B, f2 cc) {
    if (KIND == 0) return fma2(splat(q.x), L, fma2(splat(q.y), A, fma2(splat(q.z), B, add2(splat(q.w), cc))));
    if (KIND == KIND_NN) return fma2(splat(q.x), L, fma2(splat(q.y), A, fma2(splat(q.z), B, splat(q.w))));
    f2 dx = fma2(splat(q.x), splat(-1.0f), L), dy = fma2(splat(q.y), splat(-1.0f), A), dz = fma2(splat(q.z), splat(-1.0f), B);
    return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}
template <int KIND>
__device__ __forceinline__ f2 wt_weight2(f2 s, float c1, float c0) {
    if (KIND == 0) return make_float2(ex2_approx(-s.x), ex2_approx(-s.y));
    if (KIND == 1) {
        f2 e = fma2(splat(c1), make_float2(lg2_approx(s.x), lg2_approx(s.y)), splat(c0));
        return make_float2(ex2_approx(e.x), ex2_approx(e.y));
    }
    return s;
}

// Optimal sorting networks for 4 and 8 keys (5 and 19 compare-exchanges), bitonic beyond.
__device__ __forceinline__ void cswap(float& a, float& b) {
    float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}
template <int M>
__device__ __forceinline__ void wt_sort(float (&v)[M]) {
    if constexpr (M == 4) {
        cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[1], v[2]);
    } else if constexpr (M == 8) {
        cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
        cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[6]); cswap(v[5], v[7]);
        cswap(v[1], v[2]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[3], v[7]);
        cswap(v[1], v[5]); cswap(v[2], v[6]);
        cswap(v[1], v[4]); cswap(v[3], v[6]);
        cswap(v[2], v[4]); cswap(v[3], v[5]);
        cswap(v[3], v[4]);
    } else {
        bitonic_sort<M>(v);
    }
}

struct PairAcc {
    f2 n0, n1, n2, den;
    uint32_t risky;  // bit 0: first entry of the pair, bit 1: second
};

// kp-th and (kp+1)-th smallest of the first mU of `sc` (component C of each pair).
template <int M, int C>
__device__ __forceinline__ void wt_threshold(const f2 (&sc)[M], int mU, int kp, float& tau, float& nxt) {
    float s[M];
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = i < mU ? (C == 0 ? sc[i].x : sc[i].y) : INFINITY;
    wt_sort<M>(s);
    tau = s[0];
    nxt = s[M - 1];
    pick_threshold<M>(s, kp, tau, nxt);
}

// Choose the kp nearest of mU (kp < mU <= M) uncertain colours for the two entries of a pair and
// accumulate them.
template <int M, int KIND>
__device__ __forceinline__ PairAcc wt_select(const float4* __restrict__ unc, int mU, int kp, f2 L, f2 A, f2 B, f2 cc, float c1, float c0, float tol) {
    f2 sc[M];
#pragma unroll
    for (int i = 0; i < M; ++i) sc[i] = i < mU ? wt_score2<KIND>(unc[i], L, A, B, cc) : splat(INFINITY);
    float tau0, nxt0, tau1, nxt1;
    wt_threshold<M, 0>(sc, mU, kp, tau0, nxt0);
    wt_threshold<M, 1>(sc, mU, kp, tau1, nxt1);
    PairAcc r;
    r.n0 = r.n1 = r.n2 = r.den = splat(0.f);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (i < mU) {
            float4 q = unc[i];
            f2 w = wt_weight2<KIND>(sc[i], c1, c0);
            w.x = sc[i].x <= tau0 ? w.x : 0.0f;
            w.y = sc[i].y <= tau1 ? w.y : 0.0f;
            r.n0 = fma2(splat(q.x), w, r.n0);
            r.n1 = fma2(splat(q.y), w, r.n1);
            r.n2 = fma2(splat(q.z), w, r.n2);
            r.den = add2(r.den, w);
        }
    }
    r.risky = (((nxt0 - tau0 <= tol) && (nxt0 < INFINITY)) ? 1u : 0u) | (((nxt1 - tau1 <= tol) && (nxt1 < INFINITY)) ? 2u : 0u);
    return r;
}

// More than 16 uncertain colours (rare with warp tiles): per entry, one pass keeping the kp + 1
// best scores sorted in a small local array.
template <int KIND>
__device__ __forceinline__ PairAcc wt_select_insert(const float4* __restrict__ unc, int mU, int kp, f2 L, f2 A, f2 B, f2 cc, float c1, float c0,
                                                    float tol) {
    constexpr int CAP = 33;
    PairAcc r;
    r.n0 = r.n1 = r.n2 = r.den = splat(0.f);
    r.risky = 0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float td[CAP];
        unsigned char tj[CAP];
        const int keep = kp + 1;
        int cnt = 0;
        for (int j = 0; j < mU; ++j) {
            f2 s2 = wt_score2<KIND>(unc[j], L, A, B, cc);
            float d = half ? s2.y : s2.x;
            if (cnt < keep || d < td[cnt - 1]) {
                int i = cnt < keep ? cnt : keep - 1;
                while (i > 0 && td[i - 1] > d) {
                    td[i] = td[i - 1];
                    tj[i] = tj[i - 1];
                    --i;
                }
                td[i] = d;
                tj[i] = (unsigned char)j;
                if (cnt < keep) ++cnt;
            }
        }
        if (cnt < keep) { r.risky |= 1u << half; continue; }
        float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
        for (int i = 0; i < kp; ++i) {
            float4 q = unc[tj[i]];
            f2 w2 = wt_weight2<KIND>(splat(td[i]), c1, c0);
            n0 = fmaf(q.x, w2.x, n0);
            n1 = fmaf(q.y, w2.x, n1);
            n2 = fmaf(q.z, w2.x, n2);
            den += w2.x;
        }
        if (td[kp] - td[kp - 1] <= tol) r.risky |= 1u << half;
        if (half) { r.n0.y = n0; r.n1.y = n1; r.n2.y = n2; r.den.y = den; }
        else { r.n0.x = n0; r.n1.x = n1; r.n2.x = n2; r.den.x = den; }
    }
    return r;
}

// One call site per pair of entries; the bucket (4 / 8 / 16 / more) is chosen inside. Out of line:
// a copy of every network per pair made the kernel 100 KB of SASS, and warps that run free of each
// other then spend their time waiting for instruction fetch.
template <int KIND>
__device__ __noinline__ PairAcc wt_select_any(const float4* __restrict__ unc, int mU, int kp, f2 L, f2 A, f2 B, f2 cc, float c1, float c0, float tol) {
    if (mU <= 4) return wt_select<4, KIND>(unc, mU, kp, L, A, B, cc, c1, c0, tol);
    if (mU <= 8) return wt_select<8, KIND>(unc, mU, kp, L, A, B, cc, c1, c0, tol);
    if (mU <= 16) return wt_select<16, KIND>(unc, mU, kp, L, A, B, cc, c1, c0, tol);
    return wt_select_insert<KIND>(unc, mU, kp, L, A, B, cc, c1, c0, tol);
}

// Entries (bits of `mask`) this lane hands to the exact kernel. Out of line and rare.
__device__ __noinline__ void wt_flag_mask(uint32_t* flag_count, uint32_t* flag_list, uint32_t* stats, uint32_t mask, uint32_t slot0,
                                          uint32_t slot_step) {
    while (mask) {
        const uint32_t e = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t pos = atomicAdd(flag_count, 1u);
        flag_list[pos] = slot0 + e * slot_step;
        if (stats) atomicAdd(&stats[127], 1u);
    }
}

// n / d for n < 2^24 with inv = 1.0f / d (host): float estimate, one correction either way.
__device__ __forceinline__ void wt_divmod(uint32_t n, uint32_t d, float inv, uint32_t& q, uint32_t& r) {
    q = (uint32_t)(__uint2float_rz(n) * inv);
    int rem = (int)(n - q * d);
    if (rem < 0) { --q; rem += (int)d; }
    if (rem >= (int)d) { ++q; rem -= (int)d; }
    r = (uint32_t)rem;
}

// ---------------------------------------------------------------------------------------------
// Pieces shared by the two kernels below.

// A warp tile after step 1: the lane's EPT colours as Oklab register pairs, where they are stored,
// the tile's sRGB box and its (inflated) Oklab box.
template <int EPT>
struct WtTile {
    f2 Lp[EPT / 2], Ap[EPT / 2], Bp[EPT / 2];
    uint32_t R, G, Bq[EPT];      // BRICK: channel values (R, G per lane, B per entry); IMAGE: Bq = packed rgb
    uint32_t slot0, slot_step;   // entry e's slot = slot0 + e * slot_step
    uint32_t valid_mask;
    uint32_t blo[3], bhi[3];
    float lo[3], hi[3];
};

struct WtShared {
    const SmemTables* st;
    const float* lin;     // decode[channel value of grid coordinate v]   (BRICK)
    const uint8_t* chv;   // channel value of grid coordinate v            (BRICK)
};

// Step 1 for a brick tile with origin (r0, g0, b0): lane = (r, g), entry e = b plane.
template <int EPT>
__device__ __forceinline__ void wt_convert_brick(const WarpArgs& a, const WtShared& sh, uint32_t lane, uint32_t r0, uint32_t g0, uint32_t b0,
                                                 WtTile<EPT>& T) {
    constexpr int NP = EPT / 2;
    const uint32_t cs = a.f.cs, csm1 = a.f.csm1;
    const uint32_t begin = (uint32_t)a.f.begin, end = (uint32_t)(a.f.begin + a.f.count);
    const uint32_t r = r0 + (lane & 7u), g = g0 + (lane >> 3), rc = min(r, csm1), gc = min(g, csm1);
    const bool rg_in = r < cs && g < cs;
    T.R = sh.chv[rc]; T.G = sh.chv[gc];
    const float linr = sh.lin[rc], ling = sh.lin[gc];
    const float pl = fmaf(0.5363325363f, ling, 0.4122214708f * linr);
    const float pm = fmaf(0.6806995451f, ling, 0.2119034982f * linr);
    const float ps = fmaf(0.2817188376f, ling, 0.0883024619f * linr);
    T.slot0 = (b0 * cs + g) * cs + r;
    T.slot_step = cs * cs;
    T.valid_mask = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const uint32_t bA = b0 + 2 * i, bB = bA + 1, bcA = min(bA, csm1), bcB = min(bB, csm1);
        const f2 linb = make_float2(sh.lin[bcA], sh.lin[bcB]);
        T.Bq[2 * i] = sh.chv[bcA];
        T.Bq[2 * i + 1] = sh.chv[bcB];
        lms_to_lab2(fma2(splat(0.0514459929f), linb, splat(pl)), fma2(splat(0.1073969566f), linb, splat(pm)),
                    fma2(splat(0.6299787005f), linb, splat(ps)), a.lum, T.Lp[i], T.Ap[i], T.Bp[i]);
        const uint32_t flatA = T.slot0 + (2 * i) * T.slot_step, flatB = flatA + T.slot_step;
        if (rg_in && bA < cs && flatA >= begin && flatA < end) T.valid_mask |= 1u << (2 * i);
        if (rg_in && bB < cs && flatB >= begin && flatB < end) T.valid_mask |= 2u << (2 * i);
    }
    T.blo[0] = sh.chv[min(r0, csm1)]; T.bhi[0] = sh.chv[min(r0 + 7u, csm1)];
    T.blo[1] = sh.chv[min(g0, csm1)]; T.bhi[1] = sh.chv[min(g0 + 3u, csm1)];
    T.blo[2] = sh.chv[min(b0, csm1)]; T.bhi[2] = sh.chv[min(b0 + EPT - 1, csm1)];
}

// Step 1 for 32*EPT consecutive pixels starting at `base` (base < end).
template <int EPT>
__device__ __forceinline__ void wt_convert_image(const WarpArgs& a, const WtShared& sh, uint32_t lane, uint32_t base, WtTile<EPT>& T) {
    constexpr int NP = EPT / 2;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t end = (uint32_t)(a.f.begin + a.f.count);
    T.R = T.G = 0;
    T.slot0 = base + lane;
    T.slot_step = 32u;
    T.valid_mask = 0;
    uint32_t cmin[3] = {255u, 255u, 255u}, cmax[3] = {0u, 0u, 0u};
    f2 lin[3][NP];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const uint32_t pix = T.slot0 + e * 32u;
        const bool valid = pix < end;
        // past the end: repeat the tile's first pixel (always valid) so the box is unaffected
        uchar4 v = reinterpret_cast<const uchar4*>(a.f.out)[valid ? pix : base];
        T.Bq[e] = v.x | ((uint32_t)v.y << 8) | ((uint32_t)v.z << 16);
        if (e & 1) { lin[0][e / 2].y = sh.st->decode[v.x]; lin[1][e / 2].y = sh.st->decode[v.y]; lin[2][e / 2].y = sh.st->decode[v.z]; }
        else { lin[0][e / 2].x = sh.st->decode[v.x]; lin[1][e / 2].x = sh.st->decode[v.y]; lin[2][e / 2].x = sh.st->decode[v.z]; }
        if (valid) T.valid_mask |= 1u << e;
        cmin[0] = min(cmin[0], (uint32_t)v.x); cmax[0] = max(cmax[0], (uint32_t)v.x);
        cmin[1] = min(cmin[1], (uint32_t)v.y); cmax[1] = max(cmax[1], (uint32_t)v.y);
        cmin[2] = min(cmin[2], (uint32_t)v.z); cmax[2] = max(cmax[2], (uint32_t)v.z);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        f2 l = fma2(splat(0.4122214708f), lin[0][i], fma2(splat(0.5363325363f), lin[1][i], mul2(splat(0.0514459929f), lin[2][i])));
        f2 m = fma2(splat(0.2119034982f), lin[0][i], fma2(splat(0.6806995451f), lin[1][i], mul2(splat(0.1073969566f), lin[2][i])));
        f2 s = fma2(splat(0.0883024619f), lin[0][i], fma2(splat(0.2817188376f), lin[1][i], mul2(splat(0.6299787005f), lin[2][i])));
        lms_to_lab2(l, m, s, a.lum, T.Lp[i], T.Ap[i], T.Bp[i]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T.blo[c] = __reduce_min_sync(FULL, cmin[c]);
        T.bhi[c] = __reduce_max_sync(FULL, cmax[c]);
    }
}

// Per-lane extent of the tile's colours, then the tile's Oklab bounding box (warp reduction), inflated by PAD.
template <int EPT>
__device__ __forceinline__ void wt_minmax(const WtTile<EPT>& T, float (&mn)[3], float (&mx)[3], bool first) {
    constexpr int NP = EPT / 2;
    if (first) {
        mn[0] = mx[0] = T.Lp[0].x; mn[1] = mx[1] = T.Ap[0].x; mn[2] = mx[2] = T.Bp[0].x;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        mn[0] = fminf(mn[0], fminf(T.Lp[i].x, T.Lp[i].y)); mx[0] = fmaxf(mx[0], fmaxf(T.Lp[i].x, T.Lp[i].y));
        mn[1] = fminf(mn[1], fminf(T.Ap[i].x, T.Ap[i].y)); mx[1] = fmaxf(mx[1], fmaxf(T.Ap[i].x, T.Ap[i].y));
        mn[2] = fminf(mn[2], fminf(T.Bp[i].x, T.Bp[i].y)); mx[2] = fmaxf(mx[2], fmaxf(T.Bp[i].x, T.Bp[i].y));
    }
}

__device__ __forceinline__ void wt_box_reduce(const float (&mn)[3], const float (&mx)[3], float (&lo)[3], float (&hi)[3]) {
    constexpr uint32_t FULL = 0xffffffffu;
    // warp reduction on integer keys: a, b + 2 are positive floats (|a|, |b| < 0.5 inside the sRGB
    // gamut; the 1.2e-7 rounding disappears in PAD); L*lum can be large and goes through f2key
    const float PAD = 2e-5f;
    lo[0] = key2f(__reduce_min_sync(FULL, f2key(mn[0]))) - PAD;
    hi[0] = key2f(__reduce_max_sync(FULL, f2key(mx[0]))) + PAD;
#pragma unroll
    for (int c = 1; c < 3; ++c) {
        lo[c] = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(mn[c] + 2.0f))) - 2.0f - PAD;
        hi[c] = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(mx[c] + 2.0f))) - 2.0f + PAD;
    }
}

template <int EPT>
__device__ __forceinline__ void wt_box(WtTile<EPT>& T) {
    float mn[3], mx[3];
    wt_minmax<EPT>(T, mn, mx, true);
    wt_box_reduce(mn, mx, T.lo, T.hi);
}

// What the list holds for palette colour p (header).
template <int KIND>
__device__ __forceinline__ float4 wt_list_entry(float4 p, float psq, float c1, float lbmin) {
    if (KIND == 0) {
        const float C = -c1;
        return make_float4(-2.f * C * p.x, -2.f * C * p.y, -2.f * C * p.z, C * (psq - lbmin));
    }
    if (KIND == KIND_NN) return make_float4(-2.f * p.x, -2.f * p.y, -2.f * p.z, psq);
    return p;
}

// Outcome of the classification of one tile (warp uniform).
struct WtClass {
    uint32_t nin, nunc;          // list lengths: cand[0 .. nin) IN, cand[nin .. nin + nunc) UNCERTAIN
    uint32_t nin_all, nunc_all;  // class sizes (negligible colours included)
    uint32_t nhit;               // exact-hit candidates (may exceed FAST_MAX_HITS: then every entry is flagged)
    bool take_all;
    bool force_exact;            // the tile cannot be handled here
    float lbmin, ubmin, ubmax_unc;
};

// Where a tile's staged bytes go: a per-warp shared-memory image of `pitch` bytes per (b, g) row.
// slot = the tile's place in a hand-out of WT_BATCH tiles that are consecutive in r (its 24 bytes of
// every row start at byte 24 * slot); defer = the caller writes the hand-out's rows out in one go
// (wt_flush_batch) instead of tile by tile.
struct WtStage {
    uint8_t* const* extras;   // shared-memory copy of Front::extra (a dynamic index into the kernel parameters would move them to local memory)
    uint32_t* buf;
    uint32_t pitch;   // bytes per row: 24 (one tile) or 24 * WT_BATCH
    uint32_t slot;
    bool defer;
    uint32_t staged;  // out: bit `slot` set when the tile was staged and its store deferred
};

// One tile's rows (24 bytes each) from the staging image to the CLUT and every extra destination,
// as aligned 32-bit words.
template <int EPT>
__device__ __noinline__ void wt_flush_tile(uint8_t* out, uint8_t* const* extras, int nextra, uint32_t cs, const uint32_t* stage, uint32_t pitch,
                                           uint32_t slot, uint32_t flat0, uint32_t slot_step) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int i = 0; i < (EPT * 4 * 6 + 31) / 32; ++i) {
        const uint32_t wi = lane + 32u * i;          // word of the tile: row (e, g) = wi / 6
        if (wi < (uint32_t)EPT * 24u) {
            const uint32_t row = (wi * 43u) >> 8, e = row >> 2, gi = row & 3u, q = wi - row * 6u;   // wi / 6 for wi < 256
            const size_t off = 3ull * (flat0 + e * slot_step + gi * cs) + 4u * q;
            const uint32_t v = stage[(row * pitch + slot * 24u) / 4u + q];
            *reinterpret_cast<uint32_t*>(out + off) = v;
#pragma unroll 1
            for (int x = 0; x < nextra; ++x) *reinterpret_cast<uint32_t*>(extras[x] + off) = v;
        }
    }
}

// A whole hand-out (WT_BATCH tiles consecutive in r, all staged): every (b, g) row is 96 contiguous
// bytes = six 16-byte stores per destination -- the unit in which the rows travel to the peer GPUs.
template <int EPT>
__device__ __noinline__ void wt_flush_batch(uint8_t* out, uint8_t* const* extras, int nextra, uint32_t cs, const uint32_t* stage, uint32_t flat0,
                                            uint32_t slot_step) {
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t ROW16 = 24u * WT_BATCH / 16u;   // 16-byte pieces per row
#pragma unroll
    for (int i = 0; i < (EPT * 4 * (int)ROW16 + 31) / 32; ++i) {
        const uint32_t wi = lane + 32u * i;
        if (wi < (uint32_t)EPT * 4u * ROW16) {
            const uint32_t row = wi / ROW16, e = row >> 2, gi = row & 3u, q = wi - row * ROW16;
            const size_t off = 3ull * (flat0 + e * slot_step + gi * cs) + 16u * q;
            const uint4 v = reinterpret_cast<const uint4*>(stage)[wi];
            *reinterpret_cast<uint4*>(out + off) = v;
#pragma unroll 1
            for (int x = 0; x < nextra; ++x) *reinterpret_cast<uint4*>(extras[x] + off) = v;
        }
    }
}

// Entry-by-entry stores of one lane's four results (tiles cut by the launch's range or by the cube's
// edge, images, the table front end). Out of line: with up to eight destinations the byte stores
// are a few hundred instructions that whole tiles never execute.
template <int MODE>
__device__ __noinline__ void wt_store_entries(uint8_t* out, uint8_t* const* extras, int nextra, uint32_t slot0, uint32_t slot_step, uint32_t smask,
                                              uint4 res) {
    const uint32_t v[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (!((smask >> e) & 1u)) continue;
        const uint32_t slot = slot0 + e * slot_step;
        if (MODE == FRONT_TABLE) {
            reinterpret_cast<uint32_t*>(out)[slot] = v[e];
        } else if (MODE == FRONT_IMAGE) {
            uchar4* q = reinterpret_cast<uchar4*>(out) + slot;   // alpha stays (mod.rs:30-50 writes bytes 0..3)
            uchar4 px = *q;
            px.x = (unsigned char)v[e]; px.y = (unsigned char)(v[e] >> 8); px.z = (unsigned char)(v[e] >> 16);
            *q = px;
        } else {
            uint8_t* q = out + 3ull * slot;
            q[0] = (uint8_t)v[e]; q[1] = (uint8_t)(v[e] >> 8); q[2] = (uint8_t)(v[e] >> 16);
#pragma unroll 1
            for (int x = 0; x < nextra; ++x) {
                uint8_t* qx = extras[x] + 3ull * slot;
                qx[0] = (uint8_t)v[e]; qx[1] = (uint8_t)(v[e] >> 8); qx[2] = (uint8_t)(v[e] >> 16);
            }
        }
    }
}

// Stores the lane's EPT results (packed r | g<<8 | b<<16; bit e of `smask`: entry e has one).
// FRONT_LUT tiles that lie wholly inside the launch go through shared memory and leave as aligned
// words -- each (b, g) row of a tile is 8 entries = 24 contiguous bytes -- to the CLUT and to every
// extra destination (peer GPUs over NVLink: runs of 24 or 96 bytes instead of single bytes).
// Entries without a result (flagged for the exact kernel, which runs afterwards on the same stream
// and writes all destinations too) are sent as zeros in that case.
template <int MODE, int EPT, bool MULTI>
__device__ __forceinline__ void wt_store(const WarpArgs& a, const WtTile<EPT>& T, const uint32_t (&res)[EPT], uint32_t smask, WtStage& sg) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    if (!MULTI) {
        // one destination: plain stores, entry by entry (the CLUT stays in L2, which merges the bytes)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            if (!((smask >> e) & 1u)) continue;
            const uint32_t slot = T.slot0 + e * T.slot_step, v = res[e];
            if (MODE == FRONT_TABLE) {
                reinterpret_cast<uint32_t*>(a.f.out)[slot] = v;
            } else if (MODE == FRONT_IMAGE) {
                uchar4* q = reinterpret_cast<uchar4*>(a.f.out) + slot;   // alpha stays (mod.rs:30-50 writes bytes 0..3)
                uchar4 px = *q;
                px.x = (unsigned char)v; px.y = (unsigned char)(v >> 8); px.z = (unsigned char)(v >> 16);
                *q = px;
            } else {
                uint8_t* q = a.f.out + 3ull * slot;
                q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
            }
        }
        return;
    }
    if (MODE == FRONT_LUT) {
        const bool whole = a.words_ok && __all_sync(FULL, T.valid_mask == (1u << EPT) - 1u);
        if (whole) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(sg.buf);
            if (!sg.defer) __syncwarp();   // the previous tile's words have been read
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const uint32_t v = ((smask >> e) & 1u) ? res[e] : 0u;
                uint8_t* q = sb + (e * 4u + (lane >> 3)) * sg.pitch + sg.slot * 24u + (lane & 7u) * 3u;
                q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
            }
            if (sg.defer) { sg.staged |= 1u << sg.slot; return; }
            __syncwarp();
            wt_flush_tile<EPT>(a.f.mc ? a.f.mc : a.f.out, sg.extras, a.f.mc ? 0 : a.f.nextra, a.f.cs, sg.buf, sg.pitch, sg.slot,
                               __shfl_sync(FULL, T.slot0, 0), T.slot_step);
            return;
        }
    }
    if (EPT == 4) wt_store_entries<MODE>(a.f.out, sg.extras, a.f.nextra, T.slot0, T.slot_step, smask, make_uint4(res[0], res[1], res[2], res[3]));
}

// Step 5: everything per entry, given the candidate list.
//   cand  : per-warp list (wt_list_entry of every listed colour), cidx: palette index of each
//   hits  : packed sRGB of the palette colours that may equal a colour of the tile exactly
template <int KIND, int MODE, int EPT, bool MULTI>
__device__ __forceinline__ void wt_entries(const WarpArgs& a, const SmemTables& st, const WtTile<EPT>& T, const WtClass& cl,
                                           const float4* __restrict__ cand, const uint32_t* __restrict__ cidx,
                                           const uint32_t* __restrict__ hits, WtStage& stage) {
    constexpr int NP = EPT / 2;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    const float c1 = a.c1;
    const uint32_t k = a.k, nin = cl.nin, nunc = cl.nunc;
    const int kp = (int)k - (int)cl.nin_all;
    const bool choose = !cl.take_all && kp > 0 && (int)cl.nunc_all > kp && (int)nunc > kp;
    const uint32_t direct = (cl.take_all || kp <= 0 || choose) ? nin : nin + nunc;
    // per-tile conditions that send every entry to the exact kernel: more exact-hit candidates than
    // the list holds; weights so far from 1 that the reference's f64 sums leave their range; more
    // neighbours to choose than the selection handles
    float c0 = 0.f;
    bool all_exact = cl.force_exact || cl.nhit > (uint32_t)FAST_MAX_HITS || (choose && kp > 32);
    if (KIND == 0) all_exact |= !(-c1 * cl.lbmin < 890.0f);
    if (KIND == 1) {
        c0 = -c1 * lg2_approx(cl.lbmin > 0.f ? cl.lbmin : cl.ubmin);
        all_exact |= !(fabsf(c0) < 890.0f);
    }
    const uint32_t nhit = min(cl.nhit, (uint32_t)FAST_MAX_HITS);
    float tol = WT_TIE_SLOPE * sqrt_approx(cl.ubmax_unc);
    if (KIND == 0 || KIND == KIND_NN) tol += WT_TIE_CONST;
    if (KIND == 0) tol *= -c1;

    if (all_exact) {
        if (T.valid_mask) wt_flag_mask(a.flag_count, a.flag_list, a.stats, T.valid_mask, T.slot0, T.slot_step);
        return;
    }
    uint32_t flag_mask = 0;
    if constexpr (KIND == KIND_NN) {
        uint32_t best_j[EPT];
        uint32_t risky_mask = 0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) best_j[e] = 0;
        if (nin == 0) {
            float best[EPT], second[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) { best[e] = INFINITY; second[e] = INFINITY; }
#pragma unroll 1
            for (uint32_t j = 0; j < nunc; ++j) {
                const float4 q = cand[j];
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    f2 d2 = wt_score2<KIND_NN>(q, T.Lp[i], T.Ap[i], T.Bp[i], splat(0.f));
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = 2 * i + h;
                        const float d = h ? d2.y : d2.x;
                        const bool better = d < best[e];
                        second[e] = better ? best[e] : fminf(second[e], d);
                        best_j[e] = better ? j : best_j[e];
                        best[e] = better ? d : best[e];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if ((second[e] - best[e] <= tol) && (second[e] < INFINITY)) risky_mask |= 1u << e;
        }
        flag_mask = risky_mask & T.valid_mask;
        uint32_t res[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            res[e] = 0;
            if (!(((T.valid_mask & ~risky_mask) >> e) & 1u)) continue;
            const uint32_t idx = cidx[best_j[e]];
            if (!a.preserve) {
                res[e] = a.pal_rgb[idx] & 0xffffffu;
            } else {
                const uint32_t pr = BRICK ? T.R : (T.Bq[e] & 255u), pg = BRICK ? T.G : ((T.Bq[e] >> 8) & 255u), pb = BRICK ? T.Bq[e] : (T.Bq[e] >> 16);
                Lab ex = linear_to_oklab_exact(st.decode[pr], st.decode[pg], st.decode[pb]);
                Lab o = {ex.l, a.pal_ab[2 * idx], a.pal_ab[2 * idx + 1]};
                float r, g, b;
                oklab_to_linear_exact(o, r, g, b);
                res[e] = f32_to_srgb8(r, st.encode) | (f32_to_srgb8(g, st.encode) << 8) | (f32_to_srgb8(b, st.encode) << 16);
            }
        }
        wt_store<MODE, EPT, MULTI>(a, T, res, T.valid_mask & ~risky_mask, stage);
        if (flag_mask) wt_flag_mask(a.flag_count, a.flag_list, a.stats, flag_mask, T.slot0, T.slot_step);
        return;
    } else {
    f2 n0[NP], n1[NP], n2[NP], den[NP], cc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        n0[i] = n1[i] = n2[i] = den[i] = splat(0.f);
        cc[i] = KIND == 0 ? mul2(splat(-c1), fma2(T.Lp[i], T.Lp[i], fma2(T.Ap[i], T.Ap[i], mul2(T.Bp[i], T.Bp[i])))) : splat(0.f);
    }
#pragma unroll 2
    for (uint32_t j = 0; j < direct; ++j) {
        const float4 q = cand[j];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            f2 w = wt_weight2<KIND>(wt_score2<KIND>(q, T.Lp[i], T.Ap[i], T.Bp[i], cc[i]), c1, c0);
            n0[i] = fma2(splat(q.x), w, n0[i]);
            n1[i] = fma2(splat(q.y), w, n1[i]);
            n2[i] = fma2(splat(q.z), w, n2[i]);
            den[i] = add2(den[i], w);
        }
    }
    uint32_t risky_mask = 0;
    if (choose) {
        const float4* unc = cand + nin;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            PairAcc pa = wt_select_any<KIND>(unc, (int)nunc, kp, T.Lp[i], T.Ap[i], T.Bp[i], cc[i], c1, c0, tol);
            n0[i] = add2(n0[i], pa.n0); n1[i] = add2(n1[i], pa.n1); n2[i] = add2(n2[i], pa.n2); den[i] = add2(den[i], pa.den);
            risky_mask |= pa.risky << (2 * i);
        }
    }
    const float fs = KIND == 0 ? 0.5f / c1 : 1.0f;  // the Gaussian sums ran over -2C p = 2 c1 p
    uint32_t res[EPT], smask = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const f2 d = den[i];
        if (!(d.x > 1e-30f) || !(d.x < 1e30f)) risky_mask |= 1u << (2 * i);
        if (!(d.y > 1e-30f) || !(d.y < 1e30f)) risky_mask |= 2u << (2 * i);
        const f2 inv = mul2(make_float2(rcp_approx(d.x), rcp_approx(d.y)), splat(fs));
        f2 oL = a.preserve ? T.Lp[i] : mul2(n0[i], inv);
        oL = mul2(oL, splat(a.inv_lum));
        f2 r2, g2, b2;
        oklab_to_linear2(oL, mul2(n1[i], inv), mul2(n2[i], inv), r2, g2, b2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = 2 * i + h;
            res[e] = 0;
            if (!((T.valid_mask >> e) & 1u)) continue;
            const uint32_t pr = BRICK ? T.R : (T.Bq[e] & 255u), pg = BRICK ? T.G : ((T.Bq[e] >> 8) & 255u), pb = BRICK ? T.Bq[e] : (T.Bq[e] >> 16);
            // rbf.rs:66-68 palette.contains(&color): identical sRGB <=> identical Oklab triple
            if (nhit) {
                const uint32_t packed = pr | (pg << 8) | (pb << 16);
                bool hit = false;
#pragma unroll 1
                for (uint32_t hh = 0; hh < nhit; ++hh) hit |= (hits[hh] == packed);
                if (hit) {
                    if (MODE != FRONT_IMAGE) { res[e] = packed; smask |= 1u << e; }  // images: the pixel stays as it is
                    continue;
                }
            }
            if ((risky_mask >> e) & 1u) { flag_mask |= 1u << e; continue; }
            res[e] = f32_to_srgb8(h ? r2.y : r2.x, st.encode) | (f32_to_srgb8(h ? g2.y : g2.x, st.encode) << 8) |
                     (f32_to_srgb8(h ? b2.y : b2.x, st.encode) << 16);
            smask |= 1u << e;
        }
    }
    wt_store<MODE, EPT, MULTI>(a, T, res, smask, stage);
    if (flag_mask) wt_flag_mask(a.flag_count, a.flag_list, a.stats, flag_mask, T.slot0, T.slot_step);
    }  // RBF kinds
}

// ---------------------------------------------------------------------------------------------
// Palettes of at most 32 colours: every warp on its own, one palette colour per lane.
// GROUPS = 2 (brick front ends, single destination): the tile is 8 blue planes deep, two groups of four
// entries per lane that share ONE classification -- the second group's Oklab values wait in shared memory
// while the first is processed. Halves the classification's share of the work for a slightly larger box.
template <int KIND, int MODE, int EPT, bool MULTI, int GROUPS>
__global__ void __launch_bounds__(WT_THREADS, WT_MINB) k_remap_warp(const WarpArgs a) {
    static_assert(EPT == 4, "entries are processed in two pairs (wt_store_entries takes four results)");
    static_assert(GROUPS == 1 || (GROUPS == 2 && MODE != FRONT_IMAGE && !MULTI), "two groups: brick tiles of the lean kernel only");
    constexpr uint32_t TD = (uint32_t)EPT * GROUPS;   // blue planes per tile
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    __shared__ SmemTables st;
    __shared__ float s_lin[BRICK ? WT_MAX_CS : 1];
    __shared__ uint8_t s_chv[BRICK ? WT_MAX_CS + 3 : 4];
    __shared__ __align__(16) float4 s_cand[WT_WARPS][WT_MAX_N];
    __shared__ __align__(8) float2 s_bnd[WT_WARPS][WT_MAX_N];
    __shared__ uint32_t s_cidx[WT_WARPS][WT_MAX_N];
    __shared__ uint32_t s_hit[WT_WARPS][FAST_MAX_HITS];
    __shared__ f2 s_stash[GROUPS == 2 ? WT_WARPS : 1][6][GROUPS == 2 ? 32 : 1];   // the second group's Lp, Ap, Bp pairs
    __shared__ __align__(16) uint32_t s_stage[MULTI ? WT_WARPS : 1][MULTI ? EPT * 24 * WT_BATCH : 1];
    __shared__ uint8_t* s_extra[LUTGEN_MAX_PEERS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = a.n, k = a.k;
    const uint32_t cs = a.f.cs, csm1 = a.f.csm1;
    const uint32_t begin = (uint32_t)a.f.begin, end = (uint32_t)(a.f.begin + a.f.count);
    const float c1 = a.c1;

    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WT_THREADS) s_chv[v] = (uint8_t)(MODE == FRONT_TABLE ? v : hald_channel(v, csm1));
    }
    if (tid == 0) {
#pragma unroll
        for (int x = 0; x < LUTGEN_MAX_PEERS; ++x) s_extra[x] = a.f.extra[x];
    }
    stage_tables(st, a.tables);  // ends with __syncthreads()
    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WT_THREADS) s_lin[v] = st.decode[s_chv[v]];
        __syncthreads();
    }
    WtShared sh;
    sh.st = &st; sh.lin = s_lin; sh.chv = s_chv;
    // FRONT_LUT: a hand-out of WT_BATCH tiles is one run of 8 * WT_BATCH entries per (b, g) row when the
    // tile rows hold a multiple of WT_BATCH tiles; with cs % 16 == 0 those runs are 16-byte aligned
    const bool batch_rows = MULTI && a.words_ok == 2 && (cs & 15u) == 0 && (a.tiles_r % (uint32_t)WT_BATCH) == 0;

    // this lane's palette colour
    const bool active = lane < n;
    const float4 p = active ? a.pal[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t prgb = active ? (a.pal_rgb[lane] & 0xffffffu) : 0u;
    const float psq = fmaf(p.x, p.x, fmaf(p.y, p.y, p.z * p.z));
    float4* cand = s_cand[warp];
    float2* bnd = s_bnd[warp];
    uint32_t* cidx = s_cidx[warp];
    uint32_t* hits = s_hit[warp];

    // the hand-out after the current one is requested before the current one is processed: the atomic's
    // round trip (1.6 % of the stall samples when taken at the point of use) hides behind a hand-out's work
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.tile_counter, a.batch);
    for (;;) {
        uint32_t t0 = __shfl_sync(FULL, ticket, 0);
        if (t0 >= a.ntiles) break;
        if (lane == 0) ticket = atomicAdd(a.tile_counter, a.batch);
        if (a.shard_count > 1) t0 = ((t0 / a.batch) * a.shard_count + a.shard_index) * a.batch;
        if (t0 >= a.ntiles_global) break;
        uint32_t ir = 0, ig = 0, ib = 0;
        if (BRICK) {
            uint32_t rest;
            wt_divmod(t0, a.tiles_r, a.inv_tiles_r, rest, ir);
            wt_divmod(rest, a.tiles_g, a.inv_tiles_g, ib, ig);
        }
        const uint32_t t1 = min(t0 + a.batch, a.ntiles_global);
        WtStage sg;
        sg.extras = s_extra;
        sg.buf = s_stage[MULTI ? warp : 0];
        sg.defer = batch_rows && t1 - t0 == (uint32_t)WT_BATCH;
        sg.pitch = sg.defer ? 24u * WT_BATCH : 24u;
        sg.staged = 0;
        uint32_t batch_flat0 = 0, batch_step = 0;
        __syncwarp();  // the previous hand-out's staged rows have been read
#pragma unroll 1
        for (uint32_t t = t0; t < t1; ++t) {
            // ---- 1. colours of this tile -> Oklab (register pairs), bounding box ----------------
            WtTile<EPT> T;
            sg.slot = sg.defer ? t - t0 : 0u;
            uint32_t valid1 = 0, tile_b0 = 0;
            if (BRICK) {
                const uint32_t r0 = ir * 8u, g0 = ig * 4u, b0 = a.b_lo + ib * TD;
                tile_b0 = b0;
                if (++ir == a.tiles_r) { ir = 0; if (++ig == a.tiles_g) { ig = 0; ++ib; } }
                const uint32_t first = (b0 * cs + g0) * cs + r0;
                const uint32_t last = (min(b0 + TD - 1, csm1) * cs + min(g0 + 3u, csm1)) * cs + min(r0 + 7u, csm1);
                if (b0 > csm1 || last < begin || first >= end) continue;  // uniform
                float mn[3], mx[3];
                if (GROUPS == 2) {
                    // second group first: its colours go to the stash, its extent into the box
                    wt_convert_brick<EPT>(a, sh, lane, r0, g0, b0 + (uint32_t)EPT, T);
                    wt_minmax<EPT>(T, mn, mx, true);
                    valid1 = T.valid_mask;
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < EPT / 2; ++i) {
                        s_stash[warp][i][lane] = T.Lp[i]; s_stash[warp][2 + i][lane] = T.Ap[i]; s_stash[warp][4 + i][lane] = T.Bp[i];
                    }
                }
                wt_convert_brick<EPT>(a, sh, lane, r0, g0, b0, T);
                wt_minmax<EPT>(T, mn, mx, GROUPS == 1);
                if (GROUPS == 2) T.bhi[2] = s_chv[min(b0 + TD - 1, csm1)];
                if (!__any_sync(FULL, (T.valid_mask | valid1) != 0)) continue;
                wt_box_reduce(mn, mx, T.lo, T.hi);
            } else {
                wt_convert_image<EPT>(a, sh, lane, t * (32u * EPT), T);
                if (!__any_sync(FULL, T.valid_mask != 0)) continue;
                wt_box<EPT>(T);
            }

            // ---- 2.-4. classification, one palette colour per lane --------------------------------
            WtClass cl;
            cl.force_exact = false;
            float lbv, ubv;
            bounds_of(p, T.lo, T.hi, lbv, ubv);
            if (!active) { lbv = INFINITY; ubv = INFINITY; }
            cl.lbmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(lbv)));
            cl.ubmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ubv)));
            cl.nhit = 0;
            __syncwarp();  // the previous tile's reads of hits / cand / cidx / bnd are over
            if (KIND != KIND_NN) {
                const uint32_t qr = prgb & 255u, qg = (prgb >> 8) & 255u, qb = prgb >> 16;
                const bool inside = active && qr >= T.blo[0] && qr <= T.bhi[0] && qg >= T.blo[1] && qg <= T.bhi[1] && qb >= T.blo[2] && qb <= T.bhi[2];
                const uint32_t bh = __ballot_sync(FULL, inside);
                cl.nhit = __popc(bh);
                const uint32_t pos = __popc(bh & ((1u << lane) - 1u));
                if (inside && pos < FAST_MAX_HITS) hits[pos] = prgb;
            }
            const float lg_ubmin = KIND == 1 ? lg2_approx(cl.ubmin) : 0.f;
            bool negl = false;
            if (KIND != KIND_NN) negl = negligible_at<KIND>(lbv, cl.ubmin, lg_ubmin, c1);
            cl.take_all = (k == 0);
            if (k != 0 && KIND != KIND_NN) {
                // the k-th smallest LB is negligible <=> fewer than k colours have a non-negligible LB
                cl.take_all = (uint32_t)__popc(__ballot_sync(FULL, active && !negl)) < k;
            }
            bool is_in = active, is_unc = false;
            if (!cl.take_all) {
                // U_k = k-th smallest UB. Keys made unique by the lane number in the low mantissa bits
                // (UB rounded UP to a multiple of 32 ulps first, so the bound stays conservative).
                const uint32_t ukey = ((__float_as_uint(ubv) + 31u) & ~31u) | lane;
                bnd[lane] = make_float2(lbv, __uint_as_float(ukey));
                __syncwarp();
                uint32_t rank_u = 0, rivals = 0;
#pragma unroll 4
                for (uint32_t q = 0; q < n; ++q) {
                    const float2 bq = bnd[q];
                    rank_u += (__float_as_uint(bq.y) < ukey) ? 1u : 0u;
                    rivals += (bq.x <= ubv) ? 1u : 0u;  // counts the lane itself once (LB <= UB)
                }
                const uint32_t bu = __ballot_sync(FULL, active && rank_u == k - 1);
                const float uk = __uint_as_float(__shfl_sync(FULL, ukey, __ffs(bu) - 1) | 31u);
                const bool near = active && lbv <= uk;
                is_in = near && rivals <= k;  // rivals - 1 (others) <= k - 1
                is_unc = near && !is_in;
            }
            cl.nin_all = __popc(__ballot_sync(FULL, is_in));
            cl.nunc_all = __popc(__ballot_sync(FULL, is_unc));
            const bool list_in = is_in && !negl, list_unc = is_unc && !negl;
            const uint32_t bin = __ballot_sync(FULL, list_in), bun = __ballot_sync(FULL, list_unc), lt = (1u << lane) - 1u;
            cl.nin = __popc(bin);
            cl.nunc = __popc(bun);
            // largest distance an uncertain colour can have: scales the near-tie tolerance
            cl.ubmax_unc = __uint_as_float(__reduce_max_sync(FULL, list_unc ? __float_as_uint(ubv) : 0u));
            {
                const float4 q = wt_list_entry<KIND>(p, psq, c1, cl.lbmin);
                if (list_in) { const uint32_t pos = __popc(bin & lt); cand[pos] = q; if (KIND == KIND_NN) cidx[pos] = lane; }
                if (list_unc) { const uint32_t pos = cl.nin + __popc(bun & lt); cand[pos] = q; if (KIND == KIND_NN) cidx[pos] = lane; }
                __syncwarp();
            }
            if (a.stats && lane == 0) {
                atomicAdd(&a.stats[min(cl.nunc, 63u)], 1u);
                atomicAdd(&a.stats[64 + min(cl.nin, 63u)], 1u);
            }
            // ---- 5. per entry ------------------------------------------------------------------
            if (MULTI && t == t0) { batch_flat0 = __shfl_sync(FULL, T.slot0, 0); batch_step = T.slot_step; }
#pragma unroll 1
            for (int grp = 0; grp < GROUPS; ++grp) {
                if (GROUPS == 2 && grp == 1) {
                    // the second group: same classification, its own colours
#pragma unroll
                    for (int i = 0; i < EPT / 2; ++i) {
                        T.Lp[i] = s_stash[warp][i][lane]; T.Ap[i] = s_stash[warp][2 + i][lane]; T.Bp[i] = s_stash[warp][4 + i][lane];
                    }
#pragma unroll
                    for (int e = 0; e < EPT; ++e) T.Bq[e] = s_chv[min(tile_b0 + (uint32_t)EPT + e, csm1)];
                    T.slot0 += (uint32_t)EPT * T.slot_step;
                    T.valid_mask = valid1;
                }
                wt_entries<KIND, MODE, EPT, MULTI>(a, st, T, cl, cand, cidx, hits, sg);
            }
        }
        if (MULTI && sg.staged) {
            __syncwarp();
            if (sg.staged == (1u << WT_BATCH) - 1u) {
                wt_flush_batch<EPT>(a.f.mc ? a.f.mc : a.f.out, sg.extras, a.f.mc ? 0 : a.f.nextra, cs, sg.buf, batch_flat0, batch_step);
            } else {
                // some tiles of the hand-out went out entry by entry (or were skipped): the staged ones go tile by tile.
                // batch_flat0 is the first tile's origin only if that tile reached step 5; recompute from the tile number
                uint32_t rest, ir0, ig0, ib0;
                wt_divmod(t0, a.tiles_r, a.inv_tiles_r, rest, ir0);
                wt_divmod(rest, a.tiles_g, a.inv_tiles_g, ib0, ig0);
                const uint32_t flat_first = ((a.b_lo + ib0 * (uint32_t)EPT) * cs + ig0 * 4u) * cs + ir0 * 8u;
#pragma unroll 1
                for (uint32_t sl = 0; sl < (uint32_t)WT_BATCH; ++sl)
                    if ((sg.staged >> sl) & 1u)
                        wt_flush_tile<EPT>(a.f.mc ? a.f.mc : a.f.out, sg.extras, a.f.mc ? 0 : a.f.nextra, cs, sg.buf, sg.pitch, sl, flat_first + 8u * sl, cs * cs);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Palettes of 33 .. 256 colours: two levels. The CTA's eight warp tiles form one CTA tile
// (16 x 8 x 2 EPT entries, or 256 EPT pixels); the CTA, one palette colour per thread, bounds every
// colour against the CTA tile's box and drops those that are OUT for the whole CTA tile
// (LB > U, U >= the k-th smallest UB, found by bisection on block-wide counts). Each warp then
// classifies the survivors (a few dozen) against its own, smaller box exactly as the kernel above
// does with a whole small palette -- lanes take survivors lane, lane + 32, ...
//
// Dropping is safe for the warp level: an OUT colour is farther than the k-th neighbour of every
// colour of the CTA tile, so it is neither one of the k nearest nor closer than any of them, and
// ranks / rival counts taken over the survivors alone equal those over the whole palette. The
// "moot" rule needs the number of colours with a non-negligible LB; OUT colours are counted with
// their CTA-level LB (an over-estimate, so the rule only fires less often).
constexpr int WB_MAX_N = 256;    // one palette colour per thread
constexpr int WB_MAX_M = 128;    // survivors a warp can classify (4 per lane); beyond: exact kernel
constexpr int WB_BISECT = 12;

template <int KIND, int MODE, int EPT, bool MULTI>
__global__ void __launch_bounds__(WT_THREADS, WT_MINB) k_remap_warp_big(const WarpArgs a) {
    static_assert(EPT % 2 == 0, "entries are processed in pairs");
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    __shared__ SmemTables st;
    __shared__ float s_lin[BRICK ? WT_MAX_CS : 1];
    __shared__ uint8_t s_chv[BRICK ? WT_MAX_CS + 3 : 4];
    __shared__ __align__(16) float4 s_cand[WT_WARPS][WB_MAX_M];
    __shared__ __align__(8) float2 s_bnd[WT_WARPS][WB_MAX_M];
    __shared__ uint32_t s_cidx[WT_WARPS][WB_MAX_M];
    __shared__ uint32_t s_riv[WT_WARPS][WB_MAX_M];
    __shared__ __align__(16) float4 s_surv[WB_MAX_N];   // survivors of the CTA level (palette order)
    __shared__ uint32_t s_sidx[WB_MAX_N];               // their palette indices
    __shared__ uint32_t s_hit[FAST_MAX_HITS];
    __shared__ uint32_t s_stage[MULTI ? WT_WARPS : 1][MULTI ? EPT * 24 : 1];
    __shared__ uint8_t* s_extra[LUTGEN_MAX_PEERS];
    __shared__ float s_wbox[WT_WARPS][6];
    __shared__ uint32_t s_wsrgb[WT_WARPS][6];
    __shared__ float s_cbox[6];
    __shared__ uint32_t s_csrgb[6];
    __shared__ uint32_t s_red[4];      // ubmin, ubmax, lbmin (bit patterns), hit count
    __shared__ uint32_t s_wcnt[WT_WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = a.n, k = a.k;
    const uint32_t cs = a.f.cs, csm1 = a.f.csm1;
    const uint32_t begin = (uint32_t)a.f.begin, end = (uint32_t)(a.f.begin + a.f.count);
    const float c1 = a.c1;

    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WT_THREADS) s_chv[v] = (uint8_t)(MODE == FRONT_TABLE ? v : hald_channel(v, csm1));
    }
    if (tid == 0) {
#pragma unroll
        for (int x = 0; x < LUTGEN_MAX_PEERS; ++x) s_extra[x] = a.f.extra[x];
    }
    stage_tables(st, a.tables);  // ends with __syncthreads()
    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WT_THREADS) s_lin[v] = st.decode[s_chv[v]];
        __syncthreads();
    }
    WtShared sh;
    sh.st = &st; sh.lin = s_lin; sh.chv = s_chv;

    // this thread's palette colour (CTA level)
    const bool owner = tid < n;
    const float4 p = owner ? a.pal[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t prgb = owner ? (a.pal_rgb[tid] & 0xffffffu) : 0u;
    float4* cand = s_cand[warp];
    float2* bnd = s_bnd[warp];
    uint32_t* cidx = s_cidx[warp];
    uint32_t* riv = s_riv[warp];

    for (;;) {
        __syncthreads();  // the previous CTA tile's shared state is no longer read
        if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        if (tid < 4) s_red[tid] = tid == 1 ? 0u : (tid == 3 ? 0u : 0x7f800000u);
        __syncthreads();
        uint32_t ct = s_tile;
        if (ct >= a.ntiles) break;
        ct = ct * a.shard_count + a.shard_index;
        if (ct >= a.ntiles_global) break;

        // ---- 1. every warp converts its tile; boxes of the eight tiles -> CTA box ------------------
        WtTile<EPT> T;
        if (BRICK) {
            uint32_t ir, ig, ib, rest;
            wt_divmod(ct, a.tiles_r, a.inv_tiles_r, rest, ir);
            wt_divmod(rest, a.tiles_g, a.inv_tiles_g, ib, ig);
            const uint32_t R0 = ir * 16u, G0 = ig * 8u, B0 = a.b_lo + ib * (2u * EPT);
            const uint32_t first = (B0 * cs + G0) * cs + R0;
            const uint32_t last = (min(B0 + 2u * EPT - 1, csm1) * cs + min(G0 + 7u, csm1)) * cs + min(R0 + 15u, csm1);
            if (B0 > csm1 || last < begin || first >= end) continue;  // uniform over the CTA
            wt_convert_brick<EPT>(a, sh, lane, R0 + 8u * (warp & 1u), G0 + 4u * ((warp >> 1) & 1u), B0 + (uint32_t)EPT * (warp >> 2), T);
        } else {
            // a warp tile past the end of the image repeats the CTA tile's first pixels
            uint32_t base = ct * (256u * EPT) + warp * (32u * EPT);
            if (base >= end) base = ct * (256u * EPT);
            wt_convert_image<EPT>(a, sh, lane, base, T);
            if (ct * (256u * EPT) + warp * (32u * EPT) >= end) T.valid_mask = 0;
        }
        wt_box<EPT>(T);
        if (lane < 6) {
            s_wbox[warp][lane] = lane < 3 ? T.lo[lane] : T.hi[lane - 3];
            s_wsrgb[warp][lane] = lane < 3 ? T.blo[lane] : T.bhi[lane - 3];
        }
        const bool cta_any = __syncthreads_or(T.valid_mask != 0);
        if (!cta_any) continue;
        if (tid < 6) {
            float v = s_wbox[0][tid];
            uint32_t u = s_wsrgb[0][tid];
#pragma unroll
            for (int w = 1; w < WT_WARPS; ++w) {
                v = tid < 3 ? fminf(v, s_wbox[w][tid]) : fmaxf(v, s_wbox[w][tid]);
                u = tid < 3 ? min(u, s_wsrgb[w][tid]) : max(u, s_wsrgb[w][tid]);
            }
            s_cbox[tid] = v;
            s_csrgb[tid] = u;
        }
        __syncthreads();

        // ---- 2. CTA level: bounds of every colour against the CTA box, U, survivors ----------------
        float clo[3] = {s_cbox[0], s_cbox[1], s_cbox[2]}, chi[3] = {s_cbox[3], s_cbox[4], s_cbox[5]};
        float lbc, ubc;
        bounds_of(p, clo, chi, lbc, ubc);
        if (!owner) { lbc = INFINITY; ubc = INFINITY; }
        {
            const uint32_t wmin = __reduce_min_sync(FULL, __float_as_uint(ubc));
            const uint32_t wmax = __reduce_max_sync(FULL, owner ? __float_as_uint(ubc) : 0u);
            const uint32_t lmin = __reduce_min_sync(FULL, __float_as_uint(lbc));
            if (lane == 0) { atomicMin(&s_red[0], wmin); atomicMax(&s_red[1], wmax); atomicMin(&s_red[2], lmin); }
        }
        if (KIND != KIND_NN) {
            const uint32_t qr = prgb & 255u, qg = (prgb >> 8) & 255u, qb = prgb >> 16;
            if (owner && qr >= s_csrgb[0] && qr <= s_csrgb[3] && qg >= s_csrgb[1] && qg <= s_csrgb[4] && qb >= s_csrgb[2] && qb <= s_csrgb[5]) {
                const uint32_t h = atomicAdd(&s_red[3], 1u);
                if (h < FAST_MAX_HITS) s_hit[h] = prgb;
            }
        }
        __syncthreads();
        const float ubmin_c = __uint_as_float(s_red[0]), ubmax_c = __uint_as_float(s_red[1]);
        const uint32_t nhit_c = s_red[3];
        float U = ubmax_c;
        if (k != 0) {
            // smallest of WB_BISECT probes v with #{UB <= v} >= k: an upper bound of the k-th smallest UB
            float lo_v = ubmin_c, hi_v = ubmax_c;
#pragma unroll 1
            for (int it = 0; it < WB_BISECT; ++it) {
                const float mid = 0.5f * (lo_v + hi_v);
                const uint32_t c = (uint32_t)__syncthreads_count(owner && ubc <= mid);
                if (c >= k) hi_v = mid; else lo_v = mid;
            }
            U = hi_v;
        }
        const bool surv = owner && lbc <= U;
        uint32_t out_nonnegl = 0;
        if (KIND != KIND_NN && k != 0) {
            const float lg_c = KIND == 1 ? lg2_approx(ubmin_c) : 0.f;
            out_nonnegl = (uint32_t)__syncthreads_count(owner && !surv && !negligible_at<KIND>(lbc, ubmin_c, lg_c, c1));
        }
        {
            const uint32_t bs = __ballot_sync(FULL, surv);
            if (lane == 0) s_wcnt[warp] = __popc(bs);
            __syncthreads();
            uint32_t pre = 0;
#pragma unroll
            for (int w = 0; w < WT_WARPS; ++w) pre += ((uint32_t)w < warp) ? s_wcnt[w] : 0u;
            if (surv) {
                const uint32_t pos = pre + __popc(bs & ((1u << lane) - 1u));
                s_surv[pos] = p;
                s_sidx[pos] = tid;
            }
        }
        __syncthreads();
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < WT_WARPS; ++w) m += s_wcnt[w];

        // ---- 3. warp level: classify the m survivors against the warp's own box -------------------
        if (!__any_sync(FULL, T.valid_mask != 0)) continue;  // nothing to produce (the next barrier is at the loop top)
        WtClass cl;
        cl.force_exact = m > (uint32_t)WB_MAX_M;
        cl.nhit = nhit_c;
        cl.take_all = (k == 0);
        cl.nin = cl.nunc = cl.nin_all = cl.nunc_all = 0;
        cl.lbmin = cl.ubmin = INFINITY;
        cl.ubmax_unc = 0.f;
        if (!cl.force_exact) {
            const uint32_t nslots = (m + 31u) >> 5;
            // bounds against the warp box; U_k keys carry the survivor number in 7 low mantissa bits
            float lbm = INFINITY, ubm = INFINITY;
#pragma unroll 1
            for (uint32_t s = 0; s < nslots; ++s) {
                const uint32_t j = lane + 32u * s;
                if (j < m) {
                    float lbv, ubv;
                    bounds_of(s_surv[j], T.lo, T.hi, lbv, ubv);
                    const uint32_t ukey = ((__float_as_uint(ubv) + 127u) & ~127u) | j;
                    bnd[j] = make_float2(lbv, __uint_as_float(ukey));
                    lbm = fminf(lbm, lbv);
                    ubm = fminf(ubm, ubv);
                }
            }
            cl.lbmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(lbm)));
            cl.ubmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ubm)));
            __syncwarp();
            const float lg_ubmin = KIND == 1 ? lg2_approx(cl.ubmin) : 0.f;
            if (k != 0 && KIND != KIND_NN) {
                uint32_t nn_cnt = 0;
#pragma unroll 1
                for (uint32_t s = 0; s < nslots; ++s) {
                    const uint32_t j = lane + 32u * s;
                    const bool nonnegl = j < m && !negligible_at<KIND>(bnd[j].x, cl.ubmin, lg_ubmin, c1);
                    nn_cnt += __popc(__ballot_sync(FULL, nonnegl));
                }
                cl.take_all = nn_cnt + out_nonnegl < k;
            }
            float uk = INFINITY;
            if (!cl.take_all) {
#pragma unroll 1
                for (uint32_t s = 0; s < nslots; ++s) {
                    const uint32_t j = lane + 32u * s;
                    const float2 me = j < m ? bnd[j] : make_float2(INFINITY, INFINITY);
                    const uint32_t ukey = __float_as_uint(me.y);
                    const float ubr = __uint_as_float(ukey | 127u);
                    uint32_t rank_u = 0, rivals = 0;
#pragma unroll 4
                    for (uint32_t q = 0; q < m; ++q) {
                        const float2 bq = bnd[q];
                        rank_u += (__float_as_uint(bq.y) < ukey) ? 1u : 0u;
                        rivals += (bq.x <= ubr) ? 1u : 0u;
                    }
                    if (j < m) riv[j] = rivals;
                    const uint32_t bu = __ballot_sync(FULL, j < m && rank_u == k - 1);
                    if (bu) uk = __uint_as_float(__shfl_sync(FULL, ukey, __ffs(bu) - 1) | 127u);
                }
            }
            __syncwarp();
            uint32_t ubmax_bits = 0;
#pragma unroll 1
            for (uint32_t s = 0; s < nslots; ++s) {
                const uint32_t j = lane + 32u * s;
                bool is_in = j < m, is_unc = false, negl = false;
                float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                float lbv = INFINITY, ubr = INFINITY;
                if (j < m) {
                    pj = s_surv[j];
                    const float2 me = bnd[j];
                    lbv = me.x;
                    ubr = __uint_as_float(__float_as_uint(me.y) | 127u);
                    if (KIND != KIND_NN) negl = negligible_at<KIND>(lbv, cl.ubmin, lg_ubmin, c1);
                    if (!cl.take_all) {
                        const bool near = lbv <= uk;
                        is_in = near && riv[j] <= k;
                        is_unc = near && !is_in;
                    }
                }
                cl.nin_all += __popc(__ballot_sync(FULL, is_in));
                cl.nunc_all += __popc(__ballot_sync(FULL, is_unc));
                const bool list_in = is_in && !negl, list_unc = is_unc && !negl;
                const uint32_t bin = __ballot_sync(FULL, list_in), bun = __ballot_sync(FULL, list_unc), lt = (1u << lane) - 1u;
                // IN colours fill the list from the front, UNCERTAIN ones from the back (reversed below)
                const float4 q = wt_list_entry<KIND>(pj, fmaf(pj.x, pj.x, fmaf(pj.y, pj.y, pj.z * pj.z)), c1, cl.lbmin);
                if (list_in) { const uint32_t pos = cl.nin + __popc(bin & lt); cand[pos] = q; cidx[pos] = s_sidx[j]; }
                if (list_unc) { const uint32_t pos = WB_MAX_M - 1u - (cl.nunc + __popc(bun & lt)); cand[pos] = q; cidx[pos] = s_sidx[j]; }
                if (list_unc) ubmax_bits = max(ubmax_bits, __float_as_uint(ubr));
                cl.nin += __popc(bin);
                cl.nunc += __popc(bun);
            }
            cl.ubmax_unc = __uint_as_float(__reduce_max_sync(FULL, ubmax_bits));
            __syncwarp();
            // move the UNCERTAIN block right behind the IN block (nin + nunc <= m <= WB_MAX_M: the
            // source [M - nunc, M) never lies below the destination; copy low to high in chunks)
            {
                const uint32_t src0 = WB_MAX_M - cl.nunc, dst0 = cl.nin;
                if (src0 != dst0) {
#pragma unroll 1
                    for (uint32_t i0 = 0; i0 < cl.nunc; i0 += 32u) {
                        const uint32_t i = i0 + lane;
                        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                        uint32_t ci = 0;
                        if (i < cl.nunc) { q = cand[src0 + i]; ci = cidx[src0 + i]; }
                        __syncwarp();
                        if (i < cl.nunc) { cand[dst0 + i] = q; cidx[dst0 + i] = ci; }
                        __syncwarp();
                    }
                }
            }
        }
        if (a.stats && lane == 0) {
            atomicAdd(&a.stats[min(cl.nunc, 63u)], 1u);
            atomicAdd(&a.stats[64 + min(cl.nin, 63u)], 1u);
        }
        // ---- 4. per entry ----------------------------------------------------------------------
        WtStage sg;
        sg.extras = s_extra;
        sg.buf = s_stage[MULTI ? warp : 0];
        sg.pitch = 24u;
        sg.slot = 0;
        sg.defer = false;
        sg.staged = 0;
        wt_entries<KIND, MODE, EPT, MULTI>(a, st, T, cl, cand, cidx, s_hit, sg);
    }
}

}  // namespace lutgen
