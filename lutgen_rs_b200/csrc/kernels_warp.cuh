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
//   centred     : everything per entry is relative to the centre m of the tile's Oklab box: an entry is
//                 delta = c - m, a listed colour is p - m (small numbers: the expansion below loses no
//                 precision to |c|^2, and needs no per-entry constant).
//   Gaussian    : the list holds q = (-2C (p - m), C(|p - m|^2 - lbmin)), C = shape*log2(e), so that
//                 s = C(d - lbmin - |delta|^2) = q.xyz . delta + q.w costs 3 FFMA2 per pair (|delta|^2 is
//                 common to all colours of an entry: it cancels in num / den and in every comparison),
//                 w = ex2(-s), and the sums run over q.xyz (-2C is divided out and m added once per entry).
//   NN          : the same expansion ranks by d - |delta|^2.
//   near-ties   : as in kernels_fast.cuh every decision closer than the f32 error bound (here a
//                 per-tile constant) sends the entry to the exact f64 kernel's work list.
//
// Code size matters: warps run free of each other, so the SM's instruction cache sees the whole
// kernel body at once. Everything rare or repeated per entry (selection networks, flagging) is out
// of line, and loops that would unroll into kilobytes are kept rolled.
#pragma once
#include <cuda_fp16.h>

#include "kernels_fast.cuh"

namespace lutgen {

// Resident warps. The lean kernel is bound by instruction delivery and issue, not by latency that more resident warps
// would hide: 2 x 256 threads (128 registers), 2 x 320 (96), 3 x 192 (96) and 4 x 128 (128) per SM all land within 2 % of
// each other on the headline CLUT (tools/time_variants.py, round 2); 2 x 384 (80 registers) is 7 % slower. Three CTAs of
// 192 threads are the default: level with the best on the RBF kinds, 4 % ahead on NearestNeighbor.
#ifndef WT_MINB
#define WT_MINB 3
#endif
constexpr int WT_THREADS = 256;
constexpr int WT_WARPS = WT_THREADS / 32;
// CTA size of k_remap_warp (warps run free of each other: any multiple of 32 works; k_remap_warp_big's CTA tile needs WT_THREADS)
#ifndef WL_THREADS_N
#define WL_THREADS_N 192
#endif
constexpr int WL_THREADS = WL_THREADS_N;
constexpr int WL_WARPS = WL_THREADS / 32;
#ifndef WT_DIRECT_UNROLL
#define WT_DIRECT_UNROLL 4
#endif
constexpr int WT_DIRECT_UNROLL_N = WT_DIRECT_UNROLL;   // colours per trip of the accumulation loop
constexpr int WT_BATCH = 4;          // warp tiles per hand-out
constexpr int WT_MAX_N = 32;         // one palette colour per lane
constexpr int WT_MAX_CS = 1089;      // level 33
// Near-tie tolerance of the warp-tile kernels (squared-distance units): gap <= WT_TIE_SLOPE sqrt(d) + WT_TIE_CONST sends the
// decision to the reference's own arithmetic. Two scores are compared, and each is off by at most
//   2 sqrt(d) E   conversion: E = |fast Oklab - the reference's f32 Oklab| as a vector. The domain is the 2^24 sRGB colours
//                 and the kernels are deterministic, so E is MEASURED, not estimated: 5.87e-7 over the whole cube
//                 (tools/measure_fast_oklab.py; tests/test_gpu_parity.py::test_fast_oklab_within_tile_padding holds it
//                 below WT_OKLAB_ERR); the image front end sums the LMS terms in another order (+1 ulp of an LMS value,
//                 <= 1e-7 after the cube root and the matrix), hence WT_OKLAB_ERR = 8e-7.
//   2.5e-7 sqrt(d)  rounding of the list entry: p - m and |p - m|^2 are rounded once per component (2^-24 relative), and
//                 |p - m| <= sqrt(d) + the tile's half diagonal
// => 4 E + 5e-7 per sqrt(d), and 1e-7 for the fma chain of the score itself. (Round 1 assumed E = 3e-6 per component:
// nine times the measured value, and seven times as many entries sent to wt_resolve_tie as necessary -- the resolver
// was 6 % of the kernel's issue slots.) L is scaled by lum_factor before it enters a distance, and so is its error:
// the host passes both constants multiplied by max(1, lum) (WarpArgs::tie_slope, tie_const, pad).
constexpr float WT_OKLAB_ERR = 8e-7f;
constexpr float WT_TIE_SLOPE = 4.0f * WT_OKLAB_ERR + 5e-7f;
constexpr float WT_TIE_CONST = 1e-7f;
constexpr float WT_BOX_PAD = 2e-5f;        // the tile's Oklab box is inflated by this much (times max(1, lum)): >> E and the rounding of the +2 trick

struct WarpArgs {
    Front f;
    const float4* pal;        // n x (l*lum, a, b, 0)
    const uint32_t* pal_rgb;  // n x packed r | g<<8 | b<<16
    const float* pal_ab;      // n x 2 (NearestNeighbor preserve)
    uint32_t n, k;            // k = 0: every colour
    int preserve;
    float lum, inv_lum;
    float c1;                 // weight slope (kernels_fast.cuh header)
    float rank_scale;         // squared distances times this fit fp16 with room to spare (k_remap_warp's rank counting)
    float tie_slope, tie_const, pad;   // WT_TIE_SLOPE, WT_TIE_CONST, WT_BOX_PAD times max(1, lum)
    int whole_tiles;          // brick front ends: no tile of the launch is cut by the cube's edge or by the range
    uint32_t tiles_r, tiles_g;
    float inv_tiles_r, inv_tiles_g;
    uint32_t b_lo;            // first blue plane covered (multiple of EPT)
    uint32_t ntiles;          // tiles this launch walks (local numbering)
    uint32_t ntiles_global;   // tiles of the whole range
    uint32_t shard_index, shard_count;  // this launch takes hand-outs index, index + count, ... (1 hand-out = `batch` warp tiles / 1 CTA tile)
    uint32_t batch;           // warp tiles per hand-out: WT_BATCH, or 1 when the range has too few tiles to fill the GPU otherwise
    uint32_t tail_cut, ticket_end;   // k_remap_warp: single-tile hand-outs from ticket tail_cut on; tickets end at ticket_end
    uint32_t deal;            // shards: runs of `deal` consecutive tiles go to the shards in turn (the hand-out, or a whole strip)
    uint32_t* strip_counter;  // MULTI, strip exchange: tiles finished per local strip (zero at launch, left zero), else nullptr
    uint32_t* flag_list;
    uint32_t* flag_count;
    uint32_t* tile_counter;   // zero at launch; k_remap_warp's last CTA zeroes it (and done_counter) again
    uint32_t* done_counter;   // CTAs that have finished (k_remap_warp)
    uint32_t* stats;
    const ColorTables* tables;
    // the reference's f64 arithmetic for entries the f32 path cannot decide (wt_exact_entry)
    const double* pal64;      // n x 3: (l*lum, a, b), the f32 Oklab triple widened (rbf.rs:29-35)
    double lum64, param64;
    int words_ok;             // FRONT_LUT, cs % 4 == 0: 1 = destinations 4-byte aligned (whole tiles leave as words),
                              // 2 = 16-byte aligned (whole hand-outs leave as 16-byte pieces), 0 = byte stores
};

// Nanosecond clock shared by all SMs (low word): k_remap_warp's timeline under LUTGEN_CUDA_STATS (stats[128..132)).
__device__ __forceinline__ uint32_t wt_clock_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (uint32_t)t;
}

__device__ __forceinline__ __half2 u32_as_half2(uint32_t v) {
    return *reinterpret_cast<const __half2*>(&v);
}
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

// Score of candidate q for two entries (smaller = nearer); L, A, B are the entries' offsets from the tile centre:
//   Gaussian   C (d - lbmin - |delta|^2)  = q.xyz . delta + q.w
//   NN         d - |delta|^2              = q.xyz . delta + q.w
//   Shepard / Linear  d       (differences: small distances keep their relative precision)
template <int KIND>
__device__ __forceinline__ f2 wt_score2(float4 q, f2 L, f2 A, f2 B) {
    if (KIND == 0 || KIND == KIND_NN) return fma2(splat(q.x), L, fma2(splat(q.y), A, fma2(splat(q.z), B, splat(q.w))));
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

struct PairAcc {
    f2 n0, n1, n2, den;
    f2 tau, nxt;     // kp-th and (kp+1)-th smallest score of each entry (meaningful where a tie bit is set)
    uint32_t risky;  // bits 0, 1: near-tie at the k-boundary (first / second entry of the pair; wt_resolve_tie settles it);
                     // bits 2, 3: the selection itself gave up (the entry needs the reference's arithmetic throughout)
};

// ---- choosing among the uncertain colours: kp of mU (0 < kp < mU) ------------------------------------
// With a palette of a few dozen colours the choice is nearly always lopsided (measured on the headline CLUT:
// one colour to take or to leave out in 61 % of the choosing tiles, two in 29 %, three in 8 %):
//   from below : the c = kp smallest scores are the chosen ones; their weights are added afterwards
//   from above : the c = mU - kp largest are the ones left out; the caller's direct loop has added every
//                uncertain colour, theirs are subtracted afterwards (Gaussian and Shepard weights fall with the
//                distance, so what is subtracted is the smallest part of the sum; Linear, whose weights grow,
//                never comes here)
// One rolled pass over the list keeps the c + 1 smallest scores (or negated scores) per entry -- an insertion
// of 2c + 1 FMNMX -- each carrying its list position in four mantissa bits, and m[c] - m[c-1] is the gap the
// near-tie test looks at. Three small loops where the sorting networks this replaces were 900 instructions:
// warps run free of each other, so the kernel's working set of code has to fit the SM's 32 KB instruction
// cache -- beyond it 'no instruction' was 47 % of the stall samples.
constexpr int WT_TRACK = 6;        // c <= WT_TRACK - 1: the usual deep choice
constexpr int WT_TRACK_WIDE = 9;   // c <= 8 = every choice among up to 16 colours (c = min(kp, mU - kp))

template <int D>
__device__ __forceinline__ void wt_track(float (&m)[D], float v) {   // keep the D smallest, ascending
    float t[D];
#pragma unroll
    for (int i = 0; i + 1 < D; ++i) t[i] = fmaxf(m[i], v);
    m[0] = fminf(m[0], v);
#pragma unroll
    for (int i = 1; i < D; ++i) m[i] = fminf(m[i], t[i - 1]);
}

// Weight of the tagged score's colour into one entry's sums (sign = -1: taken out again). The weight is taken
// from the tagged score itself: the tag is at most 16 ulp of the score.
template <int KIND>
__device__ __forceinline__ void wt_add_one(const float4* __restrict__ unc, float tagged, float sign, float c1, float c0, uint32_t tm, float& n0, float& n1,
                                           float& n2, float& den) {
    const float4 q = unc[__float_as_uint(tagged) & tm];
    const float w = sign * wt_weight2<KIND>(splat(sign * tagged), c1, c0).x;
    n0 = fmaf(q.x, w, n0); n1 = fmaf(q.y, w, n1); n2 = fmaf(q.z, w, n2); den += w;
}

// mU <= 16, c = C <= WT_TRACK - 1; top: C counts the colours left out (see above).
template <int KIND, int C>
__device__ __forceinline__ PairAcc wt_select_c(const float4* __restrict__ unc, int mU, bool top, f2 L, f2 A, f2 B, float c1, float c0, float tol, uint32_t tm,
                                               float tagtol) {
    PairAcc r;
    r.n0 = r.n1 = r.n2 = r.den = splat(0.f);
    float mx[C + 1], my[C + 1];
#pragma unroll
    for (int t = 0; t <= C; ++t) mx[t] = my[t] = INFINITY;
    const float sgn = top ? -1.0f : 1.0f;
#pragma unroll 1
    for (int i = 0; i < mU; ++i) {
        const f2 sg = mul2(wt_score2<KIND>(unc[i], L, A, B), splat(sgn));
        wt_track<C + 1>(mx, __uint_as_float((__float_as_uint(sg.x) & ~tm) | (uint32_t)i));
        wt_track<C + 1>(my, __uint_as_float((__float_as_uint(sg.y) & ~tm) | (uint32_t)i));
    }
    // the boundary lies between the C-th and the (C+1)-th tracked score; the tags cost up to 16 ulp of each
    r.risky = (((mx[C] - mx[C - 1] <= fmaf(fabsf(mx[C]), tagtol, tol)) && (mx[C] < INFINITY)) ? 1u : 0u) |
              (((my[C] - my[C - 1] <= fmaf(fabsf(my[C]), tagtol, tol)) && (my[C] < INFINITY)) ? 2u : 0u);
    r.tau = top ? make_float2(-mx[C], -my[C]) : make_float2(mx[C - 1], my[C - 1]);
    r.nxt = top ? make_float2(-mx[C - 1], -my[C - 1]) : make_float2(mx[C], my[C]);
    static_assert(C <= 2, "deeper choices: wt_select_deep");
#pragma unroll
    for (int t = 0; t < C; ++t) {
        wt_add_one<KIND>(unc, mx[t], sgn, c1, c0, tm, r.n0.x, r.n1.x, r.n2.x, r.den.x);
        wt_add_one<KIND>(unc, my[t], sgn, c1, c0, tm, r.n0.y, r.n1.y, r.n2.y, r.den.y);
    }
    return r;
}

// c = 3 .. D - 1, ONE copy of the code for all of them (11 % of the choosing tiles of a small palette; D = WT_TRACK
// covers them; D = WT_TRACK_WIDE covers every choice among up to 16 colours, which the 33..256-colour kernel meets
// daily): the D smallest are tracked whatever c is, the boundary pair is picked by uniform selects, and a second rolled
// pass adds (or takes out again) every colour whose tagged score is at most the c-th.
template <int KIND, int D>
__device__ __noinline__ PairAcc wt_select_deep(const float4* __restrict__ unc, int mU, int c, bool top, f2 L, f2 A, f2 B, float c1, float c0, float tol,
                                               uint32_t tm, float tagtol) {
    PairAcc r;
    r.n0 = r.n1 = r.n2 = r.den = splat(0.f);
    float mx[D], my[D];
#pragma unroll
    for (int t = 0; t < D; ++t) mx[t] = my[t] = INFINITY;
    const float sgn = top ? -1.0f : 1.0f;
#pragma unroll 1
    for (int i = 0; i < mU; ++i) {
        const f2 sg = mul2(wt_score2<KIND>(unc[i], L, A, B), splat(sgn));
        wt_track<D>(mx, __uint_as_float((__float_as_uint(sg.x) & ~tm) | (uint32_t)i));
        wt_track<D>(my, __uint_as_float((__float_as_uint(sg.y) & ~tm) | (uint32_t)i));
    }
    float lox = mx[2], hix = mx[3], loy = my[2], hiy = my[3];   // c = 3
#pragma unroll
    for (int t = 4; t < D; ++t)
        if (c >= t) { lox = mx[t - 1]; hix = mx[t]; loy = my[t - 1]; hiy = my[t]; }   // uniform
    r.risky = (((hix - lox <= fmaf(fabsf(hix), tagtol, tol)) && (hix < INFINITY)) ? 1u : 0u) |
              (((hiy - loy <= fmaf(fabsf(hiy), tagtol, tol)) && (hiy < INFINITY)) ? 2u : 0u);
    r.tau = top ? make_float2(-hix, -hiy) : make_float2(lox, loy);
    r.nxt = top ? make_float2(-lox, -loy) : make_float2(hix, hiy);
#pragma unroll 1
    for (int i = 0; i < mU; ++i) {
        const float4 q = unc[i];
        const f2 sc = wt_score2<KIND>(q, L, A, B);
        const f2 sg = mul2(sc, splat(sgn));
        f2 w = mul2(wt_weight2<KIND>(sc, c1, c0), splat(sgn));
        w.x = __uint_as_float((__float_as_uint(sg.x) & ~tm) | (uint32_t)i) <= lox ? w.x : 0.f;
        w.y = __uint_as_float((__float_as_uint(sg.y) & ~tm) | (uint32_t)i) <= loy ? w.y : 0.f;
        r.n0 = fma2(splat(q.x), w, r.n0);
        r.n1 = fma2(splat(q.y), w, r.n1);
        r.n2 = fma2(splat(q.z), w, r.n2);
        r.den = add2(r.den, w);
    }
    return r;
}

template <int KIND>
__device__ __forceinline__ PairAcc wt_select_small(const float4* __restrict__ unc, int mU, int c, bool top, f2 L, f2 A, f2 B, float c1, float c0,
                                                   float tol, uint32_t tm, float tagtol) {
    if (c == 1) return wt_select_c<KIND, 1>(unc, mU, top, L, A, B, c1, c0, tol, tm, tagtol);
    if (c == 2) return wt_select_c<KIND, 2>(unc, mU, top, L, A, B, c1, c0, tol, tm, tagtol);
    if (c < WT_TRACK) return wt_select_deep<KIND, WT_TRACK>(unc, mU, c, top, L, A, B, c1, c0, tol, tm, tagtol);
    return wt_select_deep<KIND, WT_TRACK_WIDE>(unc, mU, c, top, L, A, B, c1, c0, tol, tm, tagtol);
}

// Anything else (more than 16 uncertain colours, or neither side of the choice is small): per entry, one pass
// keeping the kp + 1 best scores sorted in a small local array. Rare with warp tiles; out of line.
template <int KIND>
__device__ __noinline__ PairAcc wt_select_insert(const float4* __restrict__ unc, int mU, int kp, f2 L, f2 A, f2 B, float c1, float c0, float tol) {
    constexpr int CAP = 33;
    PairAcc r;
    r.n0 = r.n1 = r.n2 = r.den = r.tau = r.nxt = splat(0.f);
    r.risky = 0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float td[CAP];
        unsigned char tj[CAP];
        const int keep = kp + 1;
        int cnt = 0;
        for (int j = 0; j < mU; ++j) {
            f2 s2 = wt_score2<KIND>(unc[j], L, A, B);
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
        if (cnt < keep) { r.risky |= 4u << half; continue; }
        float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
        for (int i = 0; i < kp; ++i) {
            float4 q = unc[tj[i]];
            f2 w2 = wt_weight2<KIND>(splat(td[i]), c1, c0);
            n0 = fmaf(q.x, w2.x, n0);
            n1 = fmaf(q.y, w2.x, n1);
            n2 = fmaf(q.z, w2.x, n2);
            den += w2.x;
        }
        if (td[kp] - td[kp - 1] <= tol) {
            // a near-tie: wt_resolve_tie settles it when the list positions fit its band keys, else the whole entry is redone
            if (mU <= 256) { r.risky |= 1u << half; if (half) { r.tau.y = td[kp - 1]; r.nxt.y = td[kp]; } else { r.tau.x = td[kp - 1]; r.nxt.x = td[kp]; } }
            else r.risky |= 4u << half;
        }
        if (half) { r.n0.y = n0; r.n1.y = n1; r.n2.y = n2; r.den.y = den; }
        else { r.n0.x = n0; r.n1.x = n1; r.n2.x = n2; r.den.x = den; }
    }
    return r;
}

// ---- near-ties at the k-boundary, settled by the lane that met them ---------------------------------
// The f32 scores of two uncertain colours are closer than their error bound: which of them the reference
// takes depends on its own arithmetic -- f32 Oklab with a correctly rounded cube root, widened, squared
// distances in f64 in kiddo's fold order (rbf.rs:59-63, 85-97; nearest_neighbor.rs:47-48). Only the ORDER
// is in doubt, so only the distances of the colours inside the doubtful band are recomputed that way; the
// weights of whatever is chosen stay f32 like every other entry's. About 300 instructions for one lane,
// where the f64 weights, sums and divisions of wt_exact_entry are 1600 for the whole warp -- executed for
// 0.13 % of the entries, those evicted the hot loop from the instruction cache (0.30 -> 0.44 ms).
//   certain : s_i <  tau - tol          (tau = kp-th smallest score, nxt = the next one)
//   band    : tau - tol <= s_i <= nxt + tol   -> the kp - |certain| nearest by (f64 distance, palette index)
// A score is within tol / 2 of its true value, so a colour below the band is nearer than at least
// mU - kp others and one above it farther than at least kp + 1 -- whatever the order inside the band.
// False: the band holds more than WT_BAND colours (the caller falls back to the full f64 evaluation).
// (A band of 8 was measured: the larger resolver makes both warp-tile kernels 2.5 % slower and settles nothing more. What
// did reach the f64 list pass with 256 colours -- six entries of a level-16 CLUT, 70 us of ONE thread's latency for them --
// were uncertain sets of more than 32 colours: the chosen set is a 64-bit mask now.)
#ifndef WT_BAND_N
#define WT_BAND_N 4
#endif
constexpr int WT_BAND = WT_BAND_N;   // (4 until round 2's end: with 256 colours a handful of entries per CLUT overflowed it, and each of them costs the f64 list pass 70 us of one thread's latency)

// linear_to_oklab_exact with ONE out-of-line copy of the correctly rounded cube root (the resolvers run rarely, but often
// enough that their code competes with the hot loop for the instruction cache: three inlined copies were 140 instructions).
__device__ __noinline__ float wt_cbrt_cr(float x) { return cbrtf_cr(x); }
__device__ __forceinline__ void wt_exact_colour(const float* __restrict__ decode, uint32_t rgb, double lum64, double (&e)[3]) {
    const float r = decode[rgb & 255u], g = decode[(rgb >> 8) & 255u], b = decode[rgb >> 16];
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(0.4122214708f, r), __fmul_rn(0.5363325363f, g)), __fmul_rn(0.0514459929f, b));
    const float m = __fadd_rn(__fadd_rn(__fmul_rn(0.2119034982f, r), __fmul_rn(0.6806995451f, g)), __fmul_rn(0.1073969566f, b));
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(0.0883024619f, r), __fmul_rn(0.2817188376f, g)), __fmul_rn(0.6299787005f, b));
    const float l_ = wt_cbrt_cr(l), m_ = wt_cbrt_cr(m), s_ = wt_cbrt_cr(s);
    const float ol = __fsub_rn(__fadd_rn(__fmul_rn(0.2104542553f, l_), __fmul_rn(0.7936177850f, m_)), __fmul_rn(0.0040720468f, s_));
    const float oa = __fadd_rn(__fsub_rn(__fmul_rn(1.9779984951f, l_), __fmul_rn(2.4285922050f, m_)), __fmul_rn(0.4505937099f, s_));
    const float ob = __fsub_rn(__fadd_rn(__fmul_rn(0.0259040371f, l_), __fmul_rn(0.7827717662f, m_)), __fmul_rn(0.8086757660f, s_));
    e[0] = __dmul_rn((double)ol, lum64);
    e[1] = (double)oa;
    e[2] = (double)ob;
}

// out: the uncertain colours' part of the entry's sums in the convention of the selection that met the tie --
// the chosen ones added (from below), or the ones not chosen subtracted (from above: the direct loop has added all).
template <int KIND>
__device__ __noinline__ bool wt_resolve_tie(const double* __restrict__ pal64, double lum64, const float* __restrict__ decode,
                                            const float4* __restrict__ unc, const uint32_t* __restrict__ uidx, int mU, int kp, bool top, uint32_t rgb,
                                            float L, float A, float B, float c1, float c0, float tau, float nxt, float tol, float4& out) {
    if (mU > 64) return false;
    double e[3];
    wt_exact_colour(decode, rgb, lum64, e);
    const float lo = tau - tol, hi = nxt + tol;
    // the band, ascending by (f64 distance, palette index), in registers: every index below is static
    double bd[WT_BAND];
    uint32_t bk[WT_BAND];   // palette index << 8 | list position
#pragma unroll
    for (int t = 0; t < WT_BAND; ++t) { bd[t] = INFINITY; bk[t] = 0xffffffffu; }
    int nb = 0, certain = 0;
#pragma unroll 1
    for (int i = 0; i < mU; ++i) {
        const float sc = wt_score2<KIND>(unc[i], splat(L), splat(A), splat(B)).x;
        if (sc < lo) ++certain;
        else if (sc <= hi) {
            if (nb == WT_BAND) return false;
            ++nb;
            const uint32_t j = uidx[i];
            double d = sq_dist_exact(e[0], e[1], e[2], pal64 + 3 * j);
            uint32_t key = (j << 8) | (uint32_t)i;
#pragma unroll
            for (int t = 0; t < WT_BAND; ++t) {
                const bool before = d < bd[t] || (d == bd[t] && key < bk[t]);
                const double td = bd[t];
                const uint32_t tk = bk[t];
                bd[t] = before ? d : td;
                bk[t] = before ? key : tk;
                d = before ? td : d;
                key = before ? tk : key;
            }
        }
    }
    const int need = kp - certain;
    unsigned long long chosen = 0;   // by list position (mU <= 64)
#pragma unroll
    for (int t = 0; t < WT_BAND; ++t)
        if (t < need && t < nb) chosen |= 1ull << (bk[t] & 255u);
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, den = 0.f;
#pragma unroll 1
    for (int i = 0; i < mU; ++i) {
        const float4 q = unc[i];
        const float sc = wt_score2<KIND>(q, splat(L), splat(A), splat(B)).x;
        const bool in_set = sc < lo || ((chosen >> i) & 1ull);
        if (in_set != top) {
            float w = wt_weight2<KIND>(splat(sc), c1, c0).x;
            w = top ? -w : w;
            n0 = fmaf(q.x, w, n0); n1 = fmaf(q.y, w, n1); n2 = fmaf(q.z, w, n2); den += w;
        }
    }
    out = make_float4(n0, n1, n2, den);
    return true;
}

// NearestNeighbor: the colours whose score is within tol of the best one, by (f64 distance, palette index);
// returns the list position of the winner (nearest_neighbor.rs:48: kiddo's nearest_one, first minimum).
__device__ __noinline__ uint32_t wt_resolve_tie_nn(const double* __restrict__ pal64, double lum64, const float* __restrict__ decode,
                                                   const float4* __restrict__ cand, const uint32_t* __restrict__ cidx, uint32_t ncand, uint32_t rgb,
                                                   float L, float A, float B, float best, float tol) {
    double e[3];
    wt_exact_colour(decode, rgb, lum64, e);
    double bd = INFINITY;
    uint32_t bj = 0xffffffffu, bpos = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < ncand; ++i) {
        const float sc = wt_score2<KIND_NN>(cand[i], splat(L), splat(A), splat(B)).x;
        if (sc <= best + tol) {
            const uint32_t j = cidx[i];
            const double d = sq_dist_exact(e[0], e[1], e[2], pal64 + 3 * j);
            if (d < bd || (d == bd && j < bj)) { bd = d; bj = j; bpos = i; }
        }
    }
    return bpos;
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

// ---------------------------------------------------------------------------------------------
// One colour in the reference's own arithmetic, by the whole warp (palettes of at most 32 colours:
// lane j owns palette colour j). This is what k_rbf_exact / k_nn_exact do with one thread per colour
// in a second launch over a work list; here the warp that met the undecidable entry resolves it on the
// spot, so the result leaves with the rest of its tile (one kernel, no list, no second pass whose
// few thousand serial threads were 6 % of the step). r, g, b are warp uniform.
//   rbf.rs:57-111: f32 Oklab (no contraction, correctly rounded cube root), widened; palette.contains;
//   squared distances in kiddo's fold order; the k nearest by (distance, index); weights and sums in
//   f64, ascending distance; division, casts and the conversion back as the Rust does them.
//   nearest_neighbor.rs:46-61: first minimum in palette order.
// Returns r | g<<8 | b<<16; bit 31 set: the colour is a palette colour (rbf.rs:66: left as it is).
// `scratch`: 32 x 4 doubles of this warp.
struct WtExactArgs {   // by value: a reference to the kernel's parameter block would have it copied to the stack
    const double* pal64;
    const uint32_t* pal_rgb;
    const float* pal_ab;
    double lum64, param64;
    uint32_t n, k;
    int preserve;
};
__device__ __forceinline__ WtExactArgs wt_exact_args(const WarpArgs& a) {
    WtExactArgs x;
    x.pal64 = a.pal64; x.pal_rgb = a.pal_rgb; x.pal_ab = a.pal_ab; x.lum64 = a.lum64; x.param64 = a.param64;
    x.n = a.n; x.k = a.k; x.preserve = a.preserve;
    return x;
}

template <int KIND>
__device__ __noinline__ uint32_t wt_exact_entry(const WtExactArgs a, const SmemTables& st, double* __restrict__ scratch, uint32_t r, uint32_t g,
                                                uint32_t b) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, n = a.n;
    const Lab c = linear_to_oklab_exact(st.decode[r], st.decode[g], st.decode[b]);
    const double c0 = __dmul_rn((double)c.l, a.lum64), c1 = (double)c.a, c2 = (double)c.b;
    const bool act = lane < n;
    double p[3] = {0.0, 0.0, 0.0}, d = INFINITY;
    if (act) {
        p[0] = a.pal64[3 * lane]; p[1] = a.pal64[3 * lane + 1]; p[2] = a.pal64[3 * lane + 2];
        d = sq_dist_exact(c0, c1, c2, p);
    }
    Lab o;
    if (KIND == KIND_NN) {
        // distances are >= 0: their bit patterns order like the values
        const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
        const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
        const uint32_t hmin = __reduce_min_sync(FULL, act ? hi : 0xffffffffu);
        const bool top = act && hi == hmin;
        const uint32_t lmin = __reduce_min_sync(FULL, top ? lo : 0xffffffffu);
        const uint32_t best = __ffs(__ballot_sync(FULL, top && lo == lmin)) - 1u;
        if (!a.preserve) return a.pal_rgb[best] & 0xffffffu;
        o.l = c.l; o.a = a.pal_ab[2 * best]; o.b = a.pal_ab[2 * best + 1];
    } else {
        if (__any_sync(FULL, act && p[0] == c0 && p[1] == c1 && p[2] == c2)) return r | (g << 8) | (b << 16) | 0x80000000u;
        const uint32_t k = a.k, cnt = k ? k : n;
        uint32_t rank = lane;
        if (k != 0) {
            __syncwarp();
            scratch[lane] = d;
            __syncwarp();
            rank = 0;
#pragma unroll 2
            for (uint32_t q = 0; q < n; ++q) {
                const double dq = scratch[q];
                rank += (dq < d || (dq == d && q < lane)) ? 1u : 0u;
            }
        }
        __syncwarp();
        if (act && rank < cnt) {
            const double w = radial_basis_exact<(KIND < 3 ? KIND : 0)>(d, a.param64);
            scratch[4 * rank] = __dmul_rn(p[0], w);
            scratch[4 * rank + 1] = __dmul_rn(p[1], w);
            scratch[4 * rank + 2] = __dmul_rn(p[2], w);
            scratch[4 * rank + 3] = w;
        }
        __syncwarp();
        // lanes 0..3 sum one component each, ascending distance as the reference's loop does
        double acc = 0.0;
        if (lane < 4) {
#pragma unroll 1
            for (uint32_t i = 0; i < cnt; ++i) acc = __dadd_rn(acc, scratch[4 * i + lane]);
        }
        const double n0 = __shfl_sync(FULL, acc, 0), n1 = __shfl_sync(FULL, acc, 1), n2 = __shfl_sync(FULL, acc, 2), den = __shfl_sync(FULL, acc, 3);
        o.l = a.preserve ? __double2float_rn(__ddiv_rn(c0, a.lum64)) : __double2float_rn(__ddiv_rn(__ddiv_rn(n0, den), a.lum64));
        o.a = __double2float_rn(__ddiv_rn(n1, den));
        o.b = __double2float_rn(__ddiv_rn(n2, den));
    }
    float fr, fg, fb;
    oklab_to_linear_exact(o, fr, fg, fb);
    return f32_to_srgb8(fr, st.encode2) | (f32_to_srgb8(fg, st.encode2) << 8) | (f32_to_srgb8(fb, st.encode2) << 16);
}

// The entries in `flag_mask` of every lane, one after the other through wt_exact_entry; the owning lane
// takes the result into res / smask. Warp uniform.
template <int KIND, int MODE, int EPT>
__device__ __forceinline__ void wt_resolve_exact(const WarpArgs& a, const SmemTables& st, double* scratch, uint32_t R, uint32_t G,
                                                 const uint32_t (&Bq)[EPT], uint32_t flag_mask, uint32_t (&res)[EPT], uint32_t& smask) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t pend = __ballot_sync(FULL, flag_mask != 0);
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        const uint32_t m = __shfl_sync(FULL, flag_mask, src);
        const uint32_t sr = __shfl_sync(FULL, R, src), sg = __shfl_sync(FULL, G, src);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const uint32_t q = __shfl_sync(FULL, Bq[e], src);
            if (!((m >> e) & 1u)) continue;  // uniform
            const uint32_t cr = BRICK ? sr : (q & 255u), cg = BRICK ? sg : ((q >> 8) & 255u), cb = BRICK ? q : (q >> 16);
            const uint32_t v = wt_exact_entry<KIND>(wt_exact_args(a), st, scratch, cr, cg, cb);
            if (a.stats && lane == 0) atomicAdd(&a.stats[127], 1u);
            if ((int)lane == src && !((v >> 31) && MODE == FRONT_IMAGE)) {   // images: a palette-coloured pixel stays as it is
                res[e] = v & 0xffffffu;
                smask |= 1u << e;
            }
        }
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
    float mid[3];                // centre of the box; after wt_centre, Lp / Ap / Bp are offsets from it
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
    const float pl = fmaf(0.5363325363f, ling, 0.4122214708f * linr) + LMS_FLOOR;
    const float pm = fmaf(0.6806995451f, ling, 0.2119034982f * linr) + LMS_FLOOR;
    const float ps = fmaf(0.2817188376f, ling, 0.0883024619f * linr) + LMS_FLOOR;
    T.slot0 = (b0 * cs + g) * cs + r;
    T.slot_step = cs * cs;
    T.valid_mask = 0;
    static_assert(NP == 2, "the rolled loop below stores its pair with a uniform branch");
#ifdef WT_CONV_ROLLED
#pragma unroll 1
#else
#pragma unroll   // both pairs inline: measured 3 % faster than one rolled copy (unlike the per-entry part below)
#endif
    for (int i = 0; i < NP; ++i) {
        const uint32_t bA = b0 + 2 * i, bB = bA + 1, bcA = min(bA, csm1), bcB = min(bB, csm1);
        const f2 linb = make_float2(sh.lin[bcA], sh.lin[bcB]);
        const uint32_t qA = sh.chv[bcA], qB = sh.chv[bcB];
        f2 L, A, B;
        lms_to_lab2(fma2(splat(0.0514459929f), linb, splat(pl)), fma2(splat(0.1073969566f), linb, splat(pm)),
                    fma2(splat(0.6299787005f), linb, splat(ps)), a.lum, L, A, B);
        if (i) { T.Lp[1] = L; T.Ap[1] = A; T.Bp[1] = B; T.Bq[2] = qA; T.Bq[3] = qB; }
        else { T.Lp[0] = L; T.Ap[0] = A; T.Bp[0] = B; T.Bq[0] = qA; T.Bq[1] = qB; }
        if (!a.whole_tiles) {
            const uint32_t flatA = T.slot0 + (2 * i) * T.slot_step, flatB = flatA + T.slot_step;
            if (rg_in && bA < cs && flatA >= begin && flatA < end) T.valid_mask |= 1u << (2 * i);
            if (rg_in && bB < cs && flatB >= begin && flatB < end) T.valid_mask |= 2u << (2 * i);
        }
    }
    if (a.whole_tiles) T.valid_mask = (1u << EPT) - 1u;   // every tile of the launch lies wholly inside its range (the usual case)
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
        f2 l = fma2(splat(0.4122214708f), lin[0][i], fma2(splat(0.5363325363f), lin[1][i], fma2(splat(0.0514459929f), lin[2][i], splat(LMS_FLOOR))));
        f2 m = fma2(splat(0.2119034982f), lin[0][i], fma2(splat(0.6806995451f), lin[1][i], fma2(splat(0.1073969566f), lin[2][i], splat(LMS_FLOOR))));
        f2 s = fma2(splat(0.0883024619f), lin[0][i], fma2(splat(0.2817188376f), lin[1][i], fma2(splat(0.6299787005f), lin[2][i], splat(LMS_FLOOR))));
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

__device__ __forceinline__ void wt_box_reduce(const float (&mn)[3], const float (&mx)[3], float (&lo)[3], float (&hi)[3], const float PAD) {
    constexpr uint32_t FULL = 0xffffffffu;
    // warp reduction on integer keys: a, b + 2 are positive floats (|a|, |b| < 0.5 inside the sRGB
    // gamut; the 1.2e-7 rounding disappears in PAD); L*lum can be large and goes through f2key
    lo[0] = key2f(__reduce_min_sync(FULL, f2key(mn[0]))) - PAD;
    hi[0] = key2f(__reduce_max_sync(FULL, f2key(mx[0]))) + PAD;
#pragma unroll
    for (int c = 1; c < 3; ++c) {
        lo[c] = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(mn[c] + 2.0f))) - 2.0f - PAD;
        hi[c] = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(mx[c] + 2.0f))) - 2.0f + PAD;
    }
}

template <int EPT>
__device__ __forceinline__ void wt_box(WtTile<EPT>& T, float pad) {
    float mn[3], mx[3];
    wt_minmax<EPT>(T, mn, mx, true);
    wt_box_reduce(mn, mx, T.lo, T.hi, pad);
}

// What the list holds for palette colour p, relative to the tile centre m (header).
template <int KIND>
__device__ __forceinline__ float4 wt_list_entry(float4 p, const float (&m)[3], float c1, float lbmin) {
    const float x = p.x - m[0], y = p.y - m[1], z = p.z - m[2];
    const float sq = fmaf(x, x, fmaf(y, y, z * z));
    if (KIND == 0) {
        const float C = -c1;
        return make_float4(-2.f * C * x, -2.f * C * y, -2.f * C * z, C * (sq - lbmin));
    }
    if (KIND == KIND_NN) return make_float4(-2.f * x, -2.f * y, -2.f * z, sq);
    return make_float4(x, y, z, 0.f);
}

// The tile's entries as offsets from the centre of its Oklab box (what the scores are written in).
template <int EPT>
__device__ __forceinline__ void wt_centre(WtTile<EPT>& T) {
#pragma unroll
    for (int c = 0; c < 3; ++c) T.mid[c] = 0.5f * (T.lo[c] + T.hi[c]);
#pragma unroll
    for (int i = 0; i < EPT / 2; ++i) {
        T.Lp[i] = add2(T.Lp[i], splat(-T.mid[0]));
        T.Ap[i] = add2(T.Ap[i], splat(-T.mid[1]));
        T.Bp[i] = add2(T.Bp[i], splat(-T.mid[2]));
    }
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

// Stores to an NVSwitch multicast address (one store, replicated by the switch into every GPU's copy of the
// buffer). PTX defines multicast addresses for multimem.* operations only: a plain st.global to one is
// undefined behaviour, even though both become STG in SASS.
__device__ __forceinline__ void multimem_st_u32(uint8_t* mc_addr, uint32_t v) {
    asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_v4(uint8_t* mc_addr, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

// One tile's rows (24 bytes each) from the staging image to the CLUT and every extra destination,
// as aligned 32-bit words. mc: `out` is a multicast address (and there are no extras).
template <int EPT>
__device__ __noinline__ void wt_flush_tile(uint8_t* out, bool mc, uint8_t* const* extras, int nextra, uint32_t cs, const uint32_t* stage, uint32_t pitch,
                                           uint32_t slot, uint32_t flat0, uint32_t slot_step) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int i = 0; i < (EPT * 4 * 6 + 31) / 32; ++i) {
        const uint32_t wi = lane + 32u * i;          // word of the tile: row (e, g) = wi / 6
        if (wi < (uint32_t)EPT * 24u) {
            const uint32_t row = (wi * 43u) >> 8, e = row >> 2, gi = row & 3u, q = wi - row * 6u;   // wi / 6 for wi < 256
            const size_t off = 3ull * (flat0 + e * slot_step + gi * cs) + 4u * q;
            const uint32_t v = stage[(row * pitch + slot * 24u) / 4u + q];
            if (mc) multimem_st_u32(out + off, v);
            else *reinterpret_cast<uint32_t*>(out + off) = v;
#pragma unroll 1
            for (int x = 0; x < nextra; ++x) *reinterpret_cast<uint32_t*>(extras[x] + off) = v;
        }
    }
}

// A whole hand-out (WT_BATCH tiles consecutive in r, all staged): every (b, g) row is 96 contiguous
// bytes = six 16-byte stores per destination -- the unit in which the rows travel to the peer GPUs.
template <int EPT>
__device__ __noinline__ void wt_flush_batch(uint8_t* out, bool mc, uint8_t* const* extras, int nextra, uint32_t cs, const uint32_t* stage, uint32_t flat0,
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
            if (mc) multimem_st_v4(out + off, v);
            else *reinterpret_cast<uint4*>(out + off) = v;
#pragma unroll 1
            for (int x = 0; x < nextra; ++x) *reinterpret_cast<uint4*>(extras[x] + off) = v;
        }
    }
}

// The barrier that ends an exchange, run by ONE thread once every store of this GPU to the peers has been issued and
// fenced: it publishes this rank's step number in every rank's flag array (release, system scope) and waits until all
// ranks have published theirs in its own array. The step number lives on the device (word LUTGEN_BAR_EPOCH of the
// rank's own array), so the host passes nothing per launch.
__device__ __forceinline__ void wt_peer_barrier(unsigned* const* flags, int nranks, int rank) {
    unsigned* mine = flags[rank];
    const unsigned epoch = mine[LUTGEN_BAR_EPOCH] + 1u;
    mine[LUTGEN_BAR_EPOCH] = epoch;
    // ONE system-scope fence, then plain flag stores: a st.release.sys per peer is a fence each, one after the other
    // (8 ranks: 20 us between "stores landed" and "barrier passed" in tools/time_exchange_timeline.py, on every rank)
    __threadfence_system();
    for (int t = 0; t < nranks; ++t) {
        unsigned* remote = flags[t] + rank;
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    }
    for (int t = 0; t < nranks; ++t) {
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + t) : "memory");
        } while ((int)(v - epoch) < 0);
    }
}

// Strip exchange. A strip is the tiles_r warp tiles that share (g, b): EPT planes x 4 rows of cs entries, and the 4 rows
// of a plane are ONE run of 12 cs bytes in the CLUT. The strips are dealt to the GPUs in turn; a GPU's warps write their
// tiles into the GPU's own CLUT exactly as the one-GPU kernel does, and the warp that completes a strip reads it back
// (L2) and sends its EPT runs to the peers in 16-byte pieces, 512 contiguous bytes per warp instruction -- whole 128-byte
// lines on the NVLink side whatever the hand-out size, where staged hand-outs travel as 96-byte rows.
// Needs cs % 16 == 0 (16-byte aligned runs) and 16-byte aligned destinations.
template <int EPT>
__device__ __noinline__ void wt_push_strip(const uint8_t* own, uint8_t* dst, uint8_t* const* extras, int nextra, uint32_t cs, uint32_t b0, uint32_t g0) {
    // (scalars, not the kernel's argument struct: a reference to it would make the caller copy it to local memory)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t pieces = (cs * 3u) >> 2;   // 16-byte pieces in 4 rows
#pragma unroll 1
    for (uint32_t e = 0; e < (uint32_t)EPT; ++e) {
        const size_t off = 3ull * (((size_t)(b0 + e) * cs + g0) * cs);
        const uint4* src = reinterpret_cast<const uint4*>(own + off);
#pragma unroll 1
        for (uint32_t i0 = 0; i0 < pieces; i0 += 192u) {
            uint4 v[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const uint32_t i = i0 + 32u * j + lane;
                if (i < pieces) v[j] = __ldcg(src + i);
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const uint32_t i = i0 + 32u * j + lane;
                if (i < pieces) {
                    if (dst) multimem_st_v4(dst + off + 16ull * i, v[j]);
#pragma unroll 1
                    for (int x = 0; x < nextra; ++x) *reinterpret_cast<uint4*>(extras[x] + off + 16ull * i) = v[j];
                }
            }
        }
    }
}

// What wt_strip_done needs, in shared memory (one copy per CTA) so that the call passes a pointer, not a dozen values.
struct WtStripCfg {
    uint32_t* counter;
    uint8_t* own;
    uint8_t* mc;
    int nextra;
    uint32_t cs, b_lo, tiles_g, deal, shard_count, shard_index;
};

// `n` more tiles of local strip `strip` are in this GPU's CLUT. The warp whose count completes the strip sends it.
template <int EPT>
__device__ __noinline__ void wt_strip_done(const WtStripCfg* c, uint8_t* const* extras, uint32_t strip, uint32_t n) {
    uint32_t full = 0;
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) {
        // this warp's stores before its count: a release on the atomic (MEMBAR.ALL.GPU), not __threadfence(), whose
        // MEMBAR.SC.GPU + invalidation of the SM's L1 cost 1.4 us per tile when measured here
        uint32_t prev;
        asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(c->counter + strip), "r"(n) : "memory");
        full = prev + n == c->deal;
        if (full) c->counter[strip] = 0u;   // as the next launch expects to find it
    }
    if (__shfl_sync(0xffffffffu, full, 0)) {
        __threadfence();   // the other warps' stores (counted before ours) before our loads (once per strip)
        const uint32_t gs = strip * c->shard_count + c->shard_index, sb = gs / c->tiles_g;
        wt_push_strip<EPT>(c->own, c->mc, extras, c->nextra, c->cs, c->b_lo + sb * (uint32_t)EPT, (gs - sb * c->tiles_g) * 4u);
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
        // one destination: plain stores, entry by entry (the CLUT stays in L2, which merges the bytes). Strip exchange:
        // the same, into this GPU's own CLUT -- whole strips leave for the peers later (wt_push_strip)
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
            wt_flush_tile<EPT>(a.f.mc ? a.f.mc : a.f.out, a.f.mc != nullptr, sg.extras, a.f.mc ? 0 : a.f.nextra, a.f.cs, sg.buf, sg.pitch, sg.slot,
                               __shfl_sync(FULL, T.slot0, 0), T.slot_step);
            return;
        }
    }
    if (EPT == 4) wt_store_entries<MODE>(a.f.out, sg.extras, a.f.nextra, T.slot0, T.slot_step, smask, make_uint4(res[0], res[1], res[2], res[3]));
}

// Step 5: everything per entry, given the candidate list.
//   cand  : per-warp list (wt_list_entry of every listed colour), cidx: palette index of each
//   hits  : packed sRGB of the palette colours that may equal a colour of the tile exactly
//   INLINE_EXACT : entries the f32 arithmetic cannot decide are resolved here by wt_exact_entry (`scratch`:
//           32 x 4 doubles per warp) and stored with their tile; otherwise they go to the work list of the
//           exact kernels' second pass
template <int KIND, int MODE, int EPT, bool MULTI, bool INLINE_EXACT>
__device__ __forceinline__ void wt_entries(const WarpArgs& a, const SmemTables& st, const WtTile<EPT>& T, const WtClass& cl,
                                           const float4* __restrict__ cand, const uint32_t* __restrict__ cidx,
                                           const uint32_t* __restrict__ hits, WtStage& stage, double* __restrict__ scratch) {
    constexpr int NP = EPT / 2;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    const float c1 = a.c1;
    const uint32_t k = a.k, nin = cl.nin, nunc = cl.nunc;
    const int kp = (int)k - (int)cl.nin_all;
    const bool choose = !cl.take_all && kp > 0 && (int)cl.nunc_all > kp && (int)nunc > kp;
    // choosing from above (wt_select_c): every uncertain colour is added by the direct loop, the few left out are taken out again
    const int left_out = (int)nunc - kp;
    const bool top = choose && KIND != 2 && KIND != KIND_NN && left_out < kp && left_out < WT_TRACK_WIDE && nunc <= 32u;
    const uint32_t direct = (cl.take_all || kp <= 0 || (choose && !top)) ? nin : nin + nunc;
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
    float tol = fmaf(a.tie_slope, sqrt_approx(cl.ubmax_unc), a.tie_const);
    if (KIND == 0) tol *= -c1;

    if (all_exact) {
        if (INLINE_EXACT) {
            uint32_t res[EPT], smask = 0;
#pragma unroll
            for (int e = 0; e < EPT; ++e) res[e] = 0;
            wt_resolve_exact<KIND, MODE, EPT>(a, st, scratch, T.R, T.G, T.Bq, T.valid_mask, res, smask);
            wt_store<MODE, EPT, MULTI>(a, T, res, smask, stage);
        } else if (T.valid_mask) {
            wt_flag_mask(a.flag_count, a.flag_list, a.stats, T.valid_mask, T.slot0, T.slot_step);
        }
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
                    f2 d2 = wt_score2<KIND_NN>(q, T.Lp[i], T.Ap[i], T.Bp[i]);
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
            for (int e = 0; e < EPT; ++e) {
                if (!((second[e] - best[e] <= tol) && (second[e] < INFINITY)) || !((T.valid_mask >> e) & 1u)) continue;
                // near-tie: the reference's own distances decide (this lane alone; rare)
                const uint32_t rgb = BRICK ? (T.R | (T.G << 8) | (T.Bq[e] << 16)) : T.Bq[e];
                best_j[e] = wt_resolve_tie_nn(a.pal64, a.lum64, st.decode, cand, cidx, nunc, rgb, (e & 1) ? T.Lp[e / 2].y : T.Lp[e / 2].x,
                                              (e & 1) ? T.Ap[e / 2].y : T.Ap[e / 2].x, (e & 1) ? T.Bp[e / 2].y : T.Bp[e / 2].x, best[e], tol);
                if (a.stats) atomicAdd(&a.stats[126], 1u);
            }
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
                res[e] = f32_to_srgb8(r, st.encode2) | (f32_to_srgb8(g, st.encode2) << 8) | (f32_to_srgb8(b, st.encode2) << 16);
            }
        }
        uint32_t smask = T.valid_mask & ~risky_mask;
        if (INLINE_EXACT) wt_resolve_exact<KIND, MODE, EPT>(a, st, scratch, T.R, T.G, T.Bq, flag_mask, res, smask);
        wt_store<MODE, EPT, MULTI>(a, T, res, smask, stage);
        if (!INLINE_EXACT && flag_mask) wt_flag_mask(a.flag_count, a.flag_list, a.stats, flag_mask, T.slot0, T.slot_step);
        return;
    } else {
    static_assert(NP == 2, "the loop below picks its pair with selects");
    // One pair of entries at a time, rolled: ONE copy of the accumulate / select / convert-back code for both pairs
    // (half the code -- what decides whether the SM's instruction cache holds the loop -- and half the accumulators
    // live at a time), at the price of walking the candidate list twice.
    const float fs = KIND == 0 ? 0.5f / c1 : 1.0f;  // the Gaussian sums run over -2C (p - m) = 2 c1 (p - m)
    const float4* unc = cand + nin;
    const int mU = (int)nunc;
    const int csel = top ? left_out : kp;
    // the tracked scores carry their list position in the low mantissa bits: four bits (16 ulp of the score) up to 16 uncertain
    // colours, five (32 ulp) up to 32 -- the 33..256-colour kernel meets 17..32 in a sixth of its tiles, and the insertion path
    // they used to take was 9 % of its instructions (profiles/r2_shepard256_k_remap_warp_big_ncu.txt)
    const bool small = csel < WT_TRACK_WIDE && mU <= 32;
    const uint32_t tm = mU > 16 ? 31u : 15u;
    const float tagtol = mU > 16 ? 8e-6f : 4e-6f;
    uint32_t res[EPT], smask = 0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) res[e] = 0;
#ifdef WT_PAIR_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int pr = 0; pr < NP; ++pr) {
        const f2 L = pr ? T.Lp[1] : T.Lp[0], A = pr ? T.Ap[1] : T.Ap[0], B = pr ? T.Bp[1] : T.Bp[0];
        const uint32_t bq0 = pr ? T.Bq[2] : T.Bq[0], bq1 = pr ? T.Bq[3] : T.Bq[1];
        const uint32_t vm = (T.valid_mask >> (2 * pr)) & 3u;
        f2 n0 = splat(0.f), n1 = n0, n2 = n0, den = n0;
#pragma unroll WT_DIRECT_UNROLL_N
        for (uint32_t j = 0; j < direct; ++j) {
            const float4 q = cand[j];
            const f2 w = wt_weight2<KIND>(wt_score2<KIND>(q, L, A, B), c1, c0);
            n0 = fma2(splat(q.x), w, n0);
            n1 = fma2(splat(q.y), w, n1);
            n2 = fma2(splat(q.z), w, n2);
            den = add2(den, w);
        }
        uint32_t risky = 0;   // bit h: entry h of the pair needs the reference's arithmetic throughout
        if (choose) {
            PairAcc pa = small ? wt_select_small<KIND>(unc, mU, csel, top, L, A, B, c1, c0, tol, tm, tagtol) : wt_select_insert<KIND>(unc, mU, kp, L, A, B, c1, c0, tol);
            if (pa.risky & 3u) {
                // near-ties: the reference's own distances decide the order inside the doubtful band (this lane alone; rare)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!((pa.risky >> h) & 1u)) continue;
                    pa.risky &= ~(1u << h);
                    if (!((vm >> h) & 1u)) continue;
                    const uint32_t bq = h ? bq1 : bq0;
                    const uint32_t rgb = BRICK ? (T.R | (T.G << 8) | (bq << 16)) : bq;
                    const float nx = h ? pa.nxt.y : pa.nxt.x;
                    float4 o;
                    if (wt_resolve_tie<KIND>(a.pal64, a.lum64, st.decode, unc, cidx + nin, mU, kp, top, rgb, h ? L.y : L.x, h ? A.y : A.x, h ? B.y : B.x, c1, c0,
                                             h ? pa.tau.y : pa.tau.x, nx, fmaf(fabsf(nx), tagtol, tol), o)) {
                        if (h) { pa.n0.y = o.x; pa.n1.y = o.y; pa.n2.y = o.z; pa.den.y = o.w; }
                        else { pa.n0.x = o.x; pa.n1.x = o.y; pa.n2.x = o.z; pa.den.x = o.w; }
                        if (a.stats) atomicAdd(&a.stats[126], 1u);
                    } else {
                        pa.risky |= 4u << h;
                    }
                }
            }
            n0 = add2(n0, pa.n0); n1 = add2(n1, pa.n1); n2 = add2(n2, pa.n2); den = add2(den, pa.den);
            risky = pa.risky >> 2;
        }
        // num / den, the centre added back, to linear sRGB
        if (!(den.x > 1e-30f) || !(den.x < 1e30f)) risky |= 1u;
        if (!(den.y > 1e-30f) || !(den.y < 1e30f)) risky |= 2u;
        const f2 inv = mul2(make_float2(rcp_approx(den.x), rcp_approx(den.y)), splat(fs));
        f2 oL = a.preserve ? add2(L, splat(T.mid[0])) : fma2(n0, inv, splat(T.mid[0]));
        oL = mul2(oL, splat(a.inv_lum));
        f2 r2, g2, b2;
        oklab_to_linear2(oL, fma2(n1, inv, splat(T.mid[1])), fma2(n2, inv, splat(T.mid[2])), r2, g2, b2);
        uint32_t out0 = srgb8_pack3(f32_to_srgb8_wide(r2.x, st.encode2), f32_to_srgb8_wide(g2.x, st.encode2), f32_to_srgb8_wide(b2.x, st.encode2));
        uint32_t out1 = srgb8_pack3(f32_to_srgb8_wide(r2.y, st.encode2), f32_to_srgb8_wide(g2.y, st.encode2), f32_to_srgb8_wide(b2.y, st.encode2));
        uint32_t ok = vm & ~risky;   // entries with a result
        if (nhit) {
            // rbf.rs:66-68 palette.contains(&color): identical sRGB <=> identical Oklab triple; the colour is left as it is
            const uint32_t p0 = BRICK ? (T.R | (T.G << 8) | (bq0 << 16)) : bq0, p1 = BRICK ? (T.R | (T.G << 8) | (bq1 << 16)) : bq1;
            uint32_t hit = 0;
#pragma unroll 1
            for (uint32_t hh = 0; hh < nhit; ++hh) hit |= (hits[hh] == p0 ? 1u : 0u) | (hits[hh] == p1 ? 2u : 0u);
            hit &= vm;
            if (hit & 1u) out0 = p0;
            if (hit & 2u) out1 = p1;
            risky &= ~hit;
            ok = MODE == FRONT_IMAGE ? (ok & ~hit) : (ok | hit);   // images: the pixel stays as it is
        }
        flag_mask |= (risky & vm) << (2 * pr);
        smask |= ok << (2 * pr);
        if (pr) { res[2] = out0; res[3] = out1; }
        else { res[0] = out0; res[1] = out1; }
    }
    if (INLINE_EXACT) wt_resolve_exact<KIND, MODE, EPT>(a, st, scratch, T.R, T.G, T.Bq, flag_mask, res, smask);
    wt_store<MODE, EPT, MULTI>(a, T, res, smask, stage);
    if (!INLINE_EXACT && flag_mask) wt_flag_mask(a.flag_count, a.flag_list, a.stats, flag_mask, T.slot0, T.slot_step);
    }  // RBF kinds
}

// ---------------------------------------------------------------------------------------------
// Palettes of at most 32 colours: every warp on its own, one palette colour per lane.
// GROUPS = 2 (brick front ends, single destination): the tile is 8 blue planes deep, two groups of four
// entries per lane that share ONE classification -- the second group's Oklab values wait in shared memory
// while the first is processed. Halves the classification's share of the work for a slightly larger box.
// XCH (several destinations, FRONT_LUT): 0 = one destination; 1 = tiles staged in shared memory and stored to every
// destination by the warp that computed them; 2 = strip exchange (wt_push_strip): the one-destination stores, plus a
// count per strip -- its hot loop is the one-destination kernel's (the staged variant's larger code misses the
// instruction cache: 83-87 % hit rate against 99.5 %, measured).
template <int KIND, int MODE, int EPT, int XCH, int GROUPS>
__global__ void __launch_bounds__(WL_THREADS, WT_MINB) k_remap_warp(const WarpArgs a) {
    constexpr bool MULTI = XCH == 1, STRIPS = XCH == 2;
    static_assert(EPT == 4, "entries are processed in two pairs (wt_store_entries takes four results)");
    static_assert(GROUPS == 1 || (GROUPS == 2 && MODE != FRONT_IMAGE && !MULTI), "two groups: brick tiles of the lean kernel only");
    constexpr uint32_t TD = (uint32_t)EPT * GROUPS;   // blue planes per tile
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool BRICK = MODE != FRONT_IMAGE;
    __shared__ SmemTables st;
    __shared__ float s_lin[BRICK ? WT_MAX_CS : 1];
    __shared__ uint8_t s_chv[BRICK ? WT_MAX_CS + 3 : 4];
    __shared__ __align__(16) float4 s_cand[WL_WARPS][WT_MAX_N];
    __shared__ __align__(16) uint32_t s_rank[WL_WARPS][WT_MAX_N];   // (UB, LB) of the lane's colour as an fp16 pair
    __shared__ uint32_t s_cidx[WL_WARPS][WT_MAX_N];
    __shared__ uint32_t s_hit[WL_WARPS][FAST_MAX_HITS];
    __shared__ f2 s_stash[GROUPS == 2 ? WL_WARPS : 1][6][GROUPS == 2 ? 32 : 1];   // the second group's Lp, Ap, Bp pairs
    __shared__ __align__(16) uint32_t s_stage[MULTI ? WL_WARPS : 1][MULTI ? EPT * 24 * WT_BATCH : 1];
    __shared__ uint8_t* s_extra[LUTGEN_MAX_PEERS];
    __shared__ WtStripCfg s_strip[STRIPS ? 1 : 1];
    __shared__ __align__(16) double s_exact[WL_WARPS][128];   // wt_exact_entry's scratch

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = a.n, k = a.k;
    const uint32_t cs = a.f.cs, csm1 = a.f.csm1;
    const uint32_t begin = (uint32_t)a.f.begin, end = (uint32_t)(a.f.begin + a.f.count);
    const float c1 = a.c1;

    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WL_THREADS) s_chv[v] = (uint8_t)(MODE == FRONT_TABLE ? v : hald_channel(v, csm1));
    }
    if (tid == 0) {
#pragma unroll
        for (int x = 0; x < LUTGEN_MAX_PEERS; ++x) s_extra[x] = a.f.extra[x];
        if (STRIPS) {
            s_strip[0] = {a.strip_counter, a.f.out, a.f.mc, a.f.mc ? 0 : a.f.nextra, a.f.cs, a.b_lo, a.tiles_g, a.deal, a.shard_count, a.shard_index};
        }
    }
    stage_tables(st, a.tables);  // ends with __syncthreads()
    if (BRICK) {
#pragma unroll 1
        for (uint32_t v = tid; v < cs; v += WL_THREADS) s_lin[v] = st.decode[s_chv[v]];
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
    const __half k16 = __ushort2half_rn((unsigned short)(k < 1024u ? k : 1024u));
    const uint32_t nq4 = (n + 3u) >> 2;
    float4* cand = s_cand[warp];
    uint32_t* rnk = s_rank[warp];
    uint32_t* cidx = s_cidx[warp];
    uint32_t* hits = s_hit[warp];

    // the hand-out after the current one is requested before the current one is processed: the atomic's
    // round trip (1.6 % of the stall samples when taken at the point of use) hides behind a hand-out's work
    // Tickets advance by a.batch. Up to a.tail_cut a ticket IS the first tile of a hand-out of a.batch tiles; beyond it
    // every ticket stands for a single tile (tile tail_cut + (ticket - tail_cut) / batch): the last couple of tiles per warp
    // are handed out one by one, so the warps finish within a tile of each other instead of within a hand-out (a warp
    // has only ~9 hand-outs of a level-16 CLUT: the uneven end was 5 % of the kernel).
    if (a.stats && tid == 0) atomicMin(&a.stats[128], wt_clock_ns());
    uint32_t ticket = 0;
    uint32_t pend_strip = 0, pend_n = 0;   // strip exchange: tiles whose stores are issued but not yet counted
    if (lane == 0) ticket = atomicAdd(a.tile_counter, a.batch);
    for (;;) {
        const uint32_t tk = __shfl_sync(FULL, ticket, 0);
        if (tk >= a.ticket_end) break;
        if (lane == 0) ticket = atomicAdd(a.tile_counter, a.batch);
        uint32_t t0 = tk, nb = a.batch;
        if (tk >= a.tail_cut) { t0 = a.tail_cut + (tk - a.tail_cut) / a.batch; nb = 1u; }
        const uint32_t t_local = t0;
        if (a.shard_count > 1) { const uint32_t u = t0 / a.deal; t0 = (u * a.shard_count + a.shard_index) * a.deal + (t0 - u * a.deal); }
        if (t0 >= a.ntiles_global) break;
        uint32_t ir = 0, ig = 0, ib = 0;
        if (BRICK) {
            uint32_t rest;
            wt_divmod(t0, a.tiles_r, a.inv_tiles_r, rest, ir);
            wt_divmod(rest, a.tiles_g, a.inv_tiles_g, ib, ig);
        }
        const uint32_t t1 = min(t0 + nb, a.ntiles_global);
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
                wt_box_reduce(mn, mx, T.lo, T.hi, a.pad);
            } else {
                wt_convert_image<EPT>(a, sh, lane, t * (32u * EPT), T);
                if (!__any_sync(FULL, T.valid_mask != 0)) continue;
                wt_box<EPT>(T, a.pad);
            }
            wt_centre<EPT>(T);

            // ---- 2.-4. classification, one palette colour per lane --------------------------------
            WtClass cl;
            cl.force_exact = false;
            float lbv, ubv;
            bounds_of(p, T.lo, T.hi, lbv, ubv);
            if (!active) { lbv = INFINITY; ubv = INFINITY; }
            cl.lbmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(lbv)));
            cl.ubmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ubv)));
            cl.nhit = 0;
            __syncwarp();  // the previous tile's reads of hits / cand / cidx / rnk are over
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
                // U_k = the k-th smallest UB bounds every entry's k-th neighbour distance: a colour with LB > U_k is OUT;
                // one with at most k colours (itself included) whose LB <= its UB is IN. Both counts come from ONE packed
                // comparison per palette colour: the bounds as fp16 pairs (UB rounded up, LB rounded down -- conservative),
                // (UB_q, LB_q) <= (UB_me, UB_me) as HSET2, summed with HADD2, four colours per 16-byte load. (The scalar
                // loop this replaces -- LDS, two compares, two adds per colour -- was 9 % of the kernel's instructions.)
                const __half ub16 = __float2half_ru(ubv * a.rank_scale), lb16 = __float2half_rd(lbv * a.rank_scale);
                rnk[lane] = (uint32_t)__half_as_ushort(ub16) | ((uint32_t)__half_as_ushort(lb16) << 16);
                __syncwarp();
                const __half2 me = __half2half2(ub16);
                __half2 acc0 = __float2half2_rn(0.f), acc1 = acc0;
#pragma unroll 1
                for (uint32_t q4 = 0; q4 < nq4; ++q4) {   // rolled: code size (see wt_select_c); nq4 = ceil(n / 4)
                    const uint4 v = reinterpret_cast<const uint4*>(rnk)[q4];   // inactive lanes hold (inf, inf): never <= a finite bound
                    acc0 = __hadd2(acc0, __hle2(u32_as_half2(v.x), me));
                    acc1 = __hadd2(acc1, __hle2(u32_as_half2(v.y), me));
                    acc0 = __hadd2(acc0, __hle2(u32_as_half2(v.z), me));
                    acc1 = __hadd2(acc1, __hle2(u32_as_half2(v.w), me));
                }
                const __half2 cnt = __hadd2(acc0, acc1);   // low: #{UB_q <= UB_me}, high: #{LB_q <= UB_me} (counts up to 32 are exact)
                const bool kth = active && __hge(__low2half(cnt), k16);
                const uint32_t uk = __reduce_min_sync(FULL, kth ? (uint32_t)__half_as_ushort(ub16) : 0xffffu);   // bounds are >= 0: bit patterns order like values
                const bool near = active && (uint32_t)__half_as_ushort(lb16) <= uk;
                is_in = near && __hle(__high2half(cnt), k16);
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
                const float4 q = wt_list_entry<KIND>(p, T.mid, c1, cl.lbmin);
                if (list_in) { const uint32_t pos = __popc(bin & lt); cand[pos] = q; cidx[pos] = lane; }
                if (list_unc) { const uint32_t pos = cl.nin + __popc(bun & lt); cand[pos] = q; cidx[pos] = lane; }
                __syncwarp();
            }
            if (a.stats && lane == 0) {
                atomicAdd(&a.stats[min(cl.nunc, 63u)], 1u);
                atomicAdd(&a.stats[64 + min(cl.nin, 63u)], 1u);
            }
            // ---- 5. per entry ------------------------------------------------------------------
            if (STRIPS && pend_n) {
                // strip exchange: the previous hand-out's tiles are counted here, a classification after their stores were
                // issued -- the fence the count needs finds nothing left to wait for
                wt_strip_done<EPT>(s_strip, s_extra, pend_strip, pend_n);
                pend_n = 0;
            }
            if (MULTI && t == t0) { batch_flat0 = __shfl_sync(FULL, T.slot0, 0); batch_step = T.slot_step; }
#pragma unroll 1
            for (int grp = 0; grp < GROUPS; ++grp) {
                if (GROUPS == 2 && grp == 1) {
                    // the second group: same classification, its own colours
#pragma unroll
                    for (int i = 0; i < EPT / 2; ++i) {
                        T.Lp[i] = add2(s_stash[warp][i][lane], splat(-T.mid[0]));
                        T.Ap[i] = add2(s_stash[warp][2 + i][lane], splat(-T.mid[1]));
                        T.Bp[i] = add2(s_stash[warp][4 + i][lane], splat(-T.mid[2]));
                    }
#pragma unroll
                    for (int e = 0; e < EPT; ++e) T.Bq[e] = s_chv[min(tile_b0 + (uint32_t)EPT + e, csm1)];
                    T.slot0 += (uint32_t)EPT * T.slot_step;
                    T.valid_mask = valid1;
                }
                wt_entries<KIND, MODE, EPT, MULTI, true>(a, st, T, cl, cand, cidx, hits, sg, s_exact[warp]);
            }
        }
        if (STRIPS) {
            // strip exchange: these tiles are counted as finished later, when their stores have long landed (below)
            pend_strip = t_local / a.deal;
            pend_n = t1 - t0;
        } else if (MULTI && sg.staged) {
            __syncwarp();
            if (sg.staged == (1u << WT_BATCH) - 1u) {
                wt_flush_batch<EPT>(a.f.mc ? a.f.mc : a.f.out, a.f.mc != nullptr, sg.extras, a.f.mc ? 0 : a.f.nextra, cs, sg.buf, batch_flat0, batch_step);
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
                        wt_flush_tile<EPT>(a.f.mc ? a.f.mc : a.f.out, a.f.mc != nullptr, sg.extras, a.f.mc ? 0 : a.f.nextra, cs, sg.buf, sg.pitch, sl, flat_first + 8u * sl, cs * cs);
            }
        }
    }
    if (STRIPS && pend_n) wt_strip_done<EPT>(s_strip, s_extra, pend_strip, pend_n);
    if (a.stats && lane == 0) atomicMax(&a.stats[129], wt_clock_ns());   // out of tiles
    // The last WARP to finish leaves the two counters as every launch expects to find them (zero): no memset
    // between launches, and a captured graph can be replayed as it is. (Warps, not CTAs: no CTA barrier at the end.)
    //
    // MULTI with a barrier: that warp is also the one that knows every store of this GPU to the peers' CLUTs has been
    // issued. It publishes this rank's step number in every rank's flag array (release, system scope) and waits until
    // all ranks have published theirs in its own array: when the kernel ends, the CLUT is complete on this GPU -- the
    // exchange and its barrier cost no launch of their own.
    __syncwarp();
    if (lane == 0) {
        if ((MULTI || STRIPS) && a.f.bar_nranks > 1) __threadfence_system();   // this warp's stores (ordered before by the barrier above) before its count
        else __threadfence();
        if (a.stats) atomicMax(&a.stats[130], wt_clock_ns());   // this warp's stores have landed
        const uint32_t prev = atomicAdd(a.done_counter, 1u);
        if (prev == gridDim.x * (uint32_t)WL_WARPS - 1u) {
            __threadfence();
            *a.tile_counter = 0u;
            *a.done_counter = 0u;
            if ((MULTI || STRIPS) && a.f.bar_nranks > 1) wt_peer_barrier(a.f.bar_flags, a.f.bar_nranks, a.f.bar_rank);
            if (a.stats) a.stats[131] = wt_clock_ns();   // every rank's stores have landed
        }
    }
}

// The exchange without the arithmetic (tools/time_exchange.py): every warp sends strips (pattern 0) or hand-outs of
// WT_BATCH tiles as 96-byte rows (pattern 1) of this GPU's share of a CLUT that is already in `own`, then the barrier.
// What it takes is the floor of a fused step on N GPUs; the difference between the patterns is what the size of the
// contiguous runs costs on the NVLink side.
__global__ void __launch_bounds__(WT_THREADS) k_exchange_probe(Front f, uint32_t tiles_r, uint32_t tiles_g, int pattern, uint32_t* done_counter) {
    __shared__ __align__(16) uint32_t s_stage[WT_WARPS][4 * 24 * WT_BATCH];
    __shared__ uint8_t* s_extra[LUTGEN_MAX_PEERS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, cs = f.cs;
    if (threadIdx.x == 0)
        for (int x = 0; x < LUTGEN_MAX_PEERS; ++x) s_extra[x] = f.extra[x];
    __syncthreads();
    const uint32_t gw = blockIdx.x * WT_WARPS + warp, nw = gridDim.x * WT_WARPS;
    const uint32_t nstrips = tiles_g * (cs / 4u);
    if (pattern == 0) {
        for (uint32_t ls = gw; ls * f.shard_count + f.shard_index < nstrips; ls += nw) {
            const uint32_t gs = ls * f.shard_count + f.shard_index, sb = gs / tiles_g;
            wt_push_strip<4>(f.out, f.mc, s_extra, f.mc ? 0 : f.nextra, cs, sb * 4u, (gs - sb * tiles_g) * 4u);
        }
    } else {
        const uint32_t per_row = tiles_r / WT_BATCH, nh = nstrips * per_row;
        for (uint32_t lh = gw; lh * f.shard_count + f.shard_index < nh; lh += nw) {
            const uint32_t gh = lh * f.shard_count + f.shard_index, gs = gh / per_row, ir = (gh - gs * per_row) * WT_BATCH, sb = gs / tiles_g;
            const uint32_t flat0 = ((sb * 4u) * cs + (gs - sb * tiles_g) * 4u) * cs + ir * 8u;
            __syncwarp();
            // 16 rows of 96 bytes into the staging image (6 pieces per row)
            for (uint32_t wi = lane; wi < 96u; wi += 32u) {
                const uint32_t row = wi / 6u, q = wi - row * 6u;
                reinterpret_cast<uint4*>(s_stage[warp])[wi] =
                    __ldcg(reinterpret_cast<const uint4*>(f.out + 3ull * (flat0 + (row >> 2) * cs * cs + (row & 3u) * cs) + 16u * q));
            }
            __syncwarp();
            wt_flush_batch<4>(f.mc ? f.mc : f.out, f.mc != nullptr, s_extra, f.mc ? 0 : f.nextra, cs, s_stage[warp], flat0, cs * cs);
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_system();
        if (atomicAdd(done_counter, 1u) == nw - 1u) {
            __threadfence();
            *done_counter = 0u;
            if (f.bar_nranks > 1) wt_peer_barrier(f.bar_flags, f.bar_nranks, f.bar_rank);
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

// Three CTAs per SM here: this kernel's CTAs meet at barriers (31 % of the stall samples with two), which more
// resident warps do hide.
#ifndef WB_MINB
#define WB_MINB 3
#endif
template <int KIND, int MODE, int EPT, bool MULTI>
__global__ void __launch_bounds__(WT_THREADS, WB_MINB) k_remap_warp_big(const WarpArgs a) {
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
        wt_box<EPT>(T, a.pad);
        wt_centre<EPT>(T);
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
                const float4 q = wt_list_entry<KIND>(pj, T.mid, c1, cl.lbmin);
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
        wt_entries<KIND, MODE, EPT, MULTI, false>(a, st, T, cl, cand, cidx, s_hit, sg, nullptr);
    }
}

}  // namespace lutgen
