// kernels_blur.cuh -- GaussianBlurRemapper (crates/lib/src/interpolation/gaussian_blur.rs), the
// remapper the CLI and Studio use by default: a nearest-neighbour volume of Oklab colours over the
// [r][g][b] cell grid, three separable clamped-edge Gaussian blurs (B, R, G axis order), then back
// to sRGB in Hald order. Everything here is f32 and evaluated exactly as the Rust does (no FMA
// contraction, taps summed in order), so the CLUT is bit-identical to the CPU algorithm.
//
// The reference rotates the volume between passes to keep the blurred axis contiguous for the
// CPU cache; the result of a pass does not depend on the layout, so here each pass reads strided
// lines through a shared-memory tile instead and the volume never moves.
#pragma once
#include "blur_launch.h"
#include "color.cuh"

namespace lutgen {

// Margin of k_blur_nn_volume's fast decision (squared-distance units, per unit of max(1, lum)): 4 x 8e-7 per sqrt(d) for the
// conversion (see the kernel), 5e-7 per sqrt(d) + 1.5e-6 for the rounding of the expanded distance.
constexpr float BLUR_TIE_SLOPE = 3.7e-6f, BLUR_TIE_CONST = 1.5e-6f;

// Order-preserving integer keys of floats (warp min / max by REDUX).
__device__ __forceinline__ uint32_t blur_f2key(float x) {
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float blur_key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// sq_dist, gaussian_blur.rs:504-510 (f32, left to right).
__device__ __forceinline__ float blur_sq_dist(float c0, float c1, float c2, const float* __restrict__ p) {
    float dl = __fsub_rn(c0, p[0]), da = __fsub_rn(c1, p[1]), db = __fsub_rn(c2, p[2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db));
}

// GaussianBlurRemapper::new, gaussian_blur.rs:34-53: palette -> [l * lum, a, b] (f32), plus what
// generate_lut_inner stores per cell, [l*lum / lum, a, b] (gaussian_blur.rs:232-236).
__global__ void k_blur_palette(const uint8_t* __restrict__ pal_rgb4, uint32_t n, float lum, const float* __restrict__ decode,
                               float* __restrict__ pal3, float* __restrict__ out3) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lab o = linear_to_oklab_exact(decode[pal_rgb4[4 * i]], decode[pal_rgb4[4 * i + 1]], decode[pal_rgb4[4 * i + 2]]);
    float l = __fmul_rn(o.l, lum);
    pal3[3 * i] = l; pal3[3 * i + 1] = o.a; pal3[3 * i + 2] = o.b;
    out3[3 * i] = __fdiv_rn(l, lum); out3[3 * i + 1] = o.a; out3[3 * i + 2] = o.b;
}

// Nearest-neighbour volume with the reference's hint rule (find_nearest_with_hint,
// gaussian_blur.rs:57-80; row loop of par_generate_lut_inner, l.320-358: the hint restarts at 0 for
// every (r, g) row and then walks b). With hint h the reference returns
//     h                         if d(h) < threshold            (early exit, l.61-63)
//     argmin, h winning ties    otherwise                      (strict `<` scan from best = h)
// Let M be the smallest distance of the cell, I the first palette index reaching it, and
// S = {t : d(t) < threshold or d(t) == M}. Then the result is h if h is in S, else I -- and S = {I}
// unless two colours tie at M or two colours lie within the (tiny) threshold. So the hint matters
// only in AMBIGUOUS cells (second-smallest distance == M or < threshold); everywhere else the
// result is I whatever came before. One warp per row, 64 cells at a time (two per lane, the palette
// scan in packed f32x2): all cells are resolved in parallel and only ambiguous ones -- rare --
// replay the chain.
// `chan[i]` = (i as f32 * scale).round() as u8 (l.206-210), filled on the host.

__device__ __forceinline__ float2 blur_sub2(float2 a, float b) { return __ffma2_rn(make_float2(b, b), make_float2(-1.0f, -1.0f), a); }  // a - b, exact product
// a + b with ONE rounding, as a packed instruction ptxas cannot fuse with the multiply that produced
// b: `one` is 1.0f passed at run time (a * 1.0 is exact). ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 -- even under --fmad=false, unlike the scalar .rn forms -- and the reference rounds the
// product before it adds.
__device__ __forceinline__ float2 blur_add2(float2 a, float2 b, float one) { return __ffma2_rn(a, make_float2(one, one), b); }

// The exact per-cell scan: the reference's f32 arithmetic operation for operation (exact Oklab, sq_dist
// rounded as written). M = smallest distance, I = first index reaching it, S2 = second smallest.
__device__ __forceinline__ void blur_scan_exact(float lr, float lg, float lb, float lum, const float* __restrict__ pal, uint32_t n, float& c0, float& c1,
                                                float& c2, float& M, float& S2, uint32_t& I) {
    const Lab o = linear_to_oklab_exact(lr, lg, lb);
    c0 = __fmul_rn(o.l, lum); c1 = o.a; c2 = o.b;
    float best = INFINITY, second = INFINITY;
    uint32_t bi = 0;
    for (uint32_t t = 0; t < n; ++t) {
        const float d = blur_sq_dist(c0, c1, c2, pal + 3 * t);
        const bool better = d < best;
        second = better ? best : fminf(second, d);
        best = better ? d : best;
        bi = better ? t : bi;
    }
    M = best; S2 = second; I = bi;
}

// Per cell the kernel first decides with the FAST arithmetic of the +-1 LSB kernels (Oklab within
// 3e-6 of the exact one, expanded distance |p|^2 - 2 p.c, packed f32x2 for the lane's two cells):
// if the two smallest distances are further apart than that arithmetic can be wrong, and no second
// colour comes near the early-exit threshold, the nearest colour is I whatever the hint and whatever
// the last bits of the reference's own rounding. Every other cell (a fraction of a percent) is
// rescanned with the exact arithmetic above before anything is decided.
__global__ void __launch_bounds__(BLUR_ROW_WARPS * 32) k_blur_nn_volume(float* __restrict__ colors, uint32_t size, uint32_t ch, const uint8_t* __restrict__ chan,
                                                                         const float* __restrict__ pal3, const float* __restrict__ out3, uint32_t n,
                                                                         float lum, float threshold_sq, const ColorTables* __restrict__ tables,
                                                                         int stage_palette, float one, const uint32_t* __restrict__ end_prev,
                                                                         uint32_t* __restrict__ end_cur, uint32_t* __restrict__ changed) {
    // end_prev / end_cur / changed (all NULL for par_generate_lut_inner, where every row starts from
    // hint 0): generate_lut_inner carries the hint from the last cell of a row into the next row
    // (gaussian_blur.rs:213). Rows are computed in parallel here, so the sequential chain is reached by
    // iteration: a row starts from end_prev[row - 1], the hint the previous pass left at the end of the
    // row before it, writes its own end to end_cur[row] and counts a change; the host repeats the pass
    // until nothing changes (the starting hint matters only in ambiguous cells: two passes in practice).
    extern __shared__ __align__(16) float s_blur_nn[];   // decode[256] | when it fits: palette 3n | q 4n (-2p, |p|^2)
    float* s_decode = s_blur_nn;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_decode[i] = tables->decode[i];
    float4* s_q = reinterpret_cast<float4*>(s_blur_nn + 256 + ((3 * n + 3) & ~3u));
    if (stage_palette) {
        for (uint32_t i = threadIdx.x; i < 3 * n; i += blockDim.x) s_blur_nn[256 + i] = pal3[i];
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
            const float x = pal3[3 * t], y = pal3[3 * t + 1], z = pal3[3 * t + 2];
            s_q[t] = make_float4(-2.f * x, -2.f * y, -2.f * z, fmaf(x, x, fmaf(y, y, z * z)));
        }
    }
    const float* __restrict__ s_pal = stage_palette ? s_blur_nn + 256 : pal3;
    __syncthreads();
    const uint32_t FULL = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows = size * size;
    const float es = fmaxf(1.0f, lum);   // L, and the error of the fast conversion in it, are scaled by lum
    (void)one;
    for (uint32_t row = blockIdx.x * BLUR_ROW_WARPS + warp; row < rows; row += gridDim.x * BLUR_ROW_WARPS) {
        const uint32_t r = row / size, g = row % size;
        const float lr = s_decode[chan[r]], lg = s_decode[chan[g]];
        uint32_t hint = (end_prev && row > 0) ? end_prev[row - 1] : 0u;  // result of the previous cell
        for (uint32_t b0 = 0; b0 < size; b0 += 64) {
            // cells b0 + lane (x) and b0 + 32 + lane (y); past the end they repeat the last cell
            const uint32_t bx = min(b0 + lane, size - 1), by = min(b0 + 32 + lane, size - 1);
            const float lbx = s_decode[chan[bx]], lby = s_decode[chan[by]];
            float2 best = make_float2(INFINITY, INFINITY), second = best;
            uint32_t ix = 0, iy = 0;
            bool riskx = true, risky = true;
            if (stage_palette) {
                // fast decision
                const f2 linb = make_float2(lbx, lby);
                const float pl = fmaf(0.5363325363f, lg, 0.4122214708f * lr) + LMS_FLOOR, pm = fmaf(0.6806995451f, lg, 0.2119034982f * lr) + LMS_FLOOR,
                            ps = fmaf(0.2817188376f, lg, 0.0883024619f * lr) + LMS_FLOOR;
                f2 L, A, B;
                lms_to_lab2(fma2(splat(0.0514459929f), linb, splat(pl)), fma2(splat(0.1073969566f), linb, splat(pm)),
                            fma2(splat(0.6299787005f), linb, splat(ps)), lum, L, A, B);
                // Palettes of at most 32 colours: only the colours that can be the nearest, or within the doubtful margin of
                // it, of ANY of the warp's 64 cells are scanned (ascending palette index, as the full scan goes). One colour
                // per lane is bounded against the cells' Oklab box: U = the smallest of the largest distances bounds every
                // cell's nearest distance, and a colour whose smallest distance exceeds U + slack is farther than
                // best + slack from every cell -- it is neither the nearest nor a second that could make a cell doubtful
                // (slack = the early-exit threshold plus the largest margin the tests below use). Doubtful cells are
                // rescanned over the whole palette as before. Typically 2-5 of 26 colours survive: the scan was half of the
                // kernel's instructions (profiles/r2_blur_k_blur_nn_volume_ncu.txt).
                uint32_t todo = 0xffffffffu;
                if (n <= 32u) {
                    float lo[3], hi[3];
                    {
                        const float v[3][2] = {{L.x, L.y}, {A.x, A.y}, {B.x, B.y}};
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            lo[c] = blur_key2f(__reduce_min_sync(FULL, blur_f2key(fminf(v[c][0], v[c][1])))) - 1e-5f;
                            hi[c] = blur_key2f(__reduce_max_sync(FULL, blur_f2key(fmaxf(v[c][0], v[c][1])))) + 1e-5f;
                        }
                    }
                    float lbv = INFINITY, ubv = INFINITY;
                    if (lane < n) {
                        const float pv[3] = {s_pal[3 * lane], s_pal[3 * lane + 1], s_pal[3 * lane + 2]};
                        lbv = 0.f; ubv = 0.f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float ex = fmaxf(0.f, fmaxf(lo[c] - pv[c], pv[c] - hi[c]));
                            const float fx = fmaxf(fabsf(pv[c] - lo[c]), fabsf(pv[c] - hi[c]));
                            lbv = fmaf(ex, ex, lbv);
                            ubv = fmaf(fx, fx, ubv);
                        }
                        lbv *= (1.0f - 4e-6f);
                        ubv *= (1.0f + 4e-6f);
                    }
                    const float U = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ubv)));                      // >= 0: bits order like values
                    const float Umax = __uint_as_float(__reduce_max_sync(FULL, lane < n ? __float_as_uint(ubv) : 0u));
                    const float slack = threshold_sq + 2.0f * es * (BLUR_TIE_SLOPE * sqrtf(Umax) + BLUR_TIE_CONST);
                    todo = __ballot_sync(FULL, lane < n && lbv <= U + slack);
                } else {
                    todo = 0u;
                }
                if (n <= 32u) {
#pragma unroll 1
                    while (todo) {
                        const uint32_t t = __ffs(todo) - 1u;
                        todo &= todo - 1u;
                        const float4 q = s_q[t];
                        const f2 d = fma2(splat(q.x), L, fma2(splat(q.y), A, fma2(splat(q.z), B, splat(q.w))));   // d - |c|^2
                        const bool bx_ = d.x < best.x, by_ = d.y < best.y;
                        second.x = bx_ ? best.x : fminf(second.x, d.x);
                        second.y = by_ ? best.y : fminf(second.y, d.y);
                        best.x = bx_ ? d.x : best.x;
                        best.y = by_ ? d.y : best.y;
                        ix = bx_ ? t : ix;
                        iy = by_ ? t : iy;
                    }
                } else {
#pragma unroll 2
                    for (uint32_t t = 0; t < n; ++t) {
                        const float4 q = s_q[t];
                        const f2 d = fma2(splat(q.x), L, fma2(splat(q.y), A, fma2(splat(q.z), B, splat(q.w))));   // d - |c|^2
                        const bool bx_ = d.x < best.x, by_ = d.y < best.y;
                        second.x = bx_ ? best.x : fminf(second.x, d.x);
                        second.y = by_ ? best.y : fminf(second.y, d.y);
                        best.x = bx_ ? d.x : best.x;
                        best.y = by_ ? d.y : best.y;
                        ix = bx_ ? t : ix;
                        iy = by_ ? t : iy;
                    }
                }
                // |c|^2 back in; what the fast arithmetic can be off by: the conversion -- MEASURED over all 2^24 colours,
                // 5.87e-7 as a vector (tools/measure_fast_oklab.py; bound 8e-7, scaled with lum like L itself), 2 sqrt(d)
                // of it in each of the two distances compared -- and the rounding of the expansion |p|^2 - 2 p.c
                // (three roundings of terms up to ~2 per distance). Round 1 assumed 3e-6 per coordinate: seven times the
                // margin, and every doubtful cell costs its warp a serial scan of the whole palette by one lane.
                const f2 cc = fma2(L, L, fma2(A, A, mul2(B, B)));
                const f2 d2 = add2(second, cc), gap = make_float2(second.x - best.x, second.y - best.y);
                const float mx = es * (BLUR_TIE_SLOPE * sqrtf(fmaxf(d2.x, 0.f)) + BLUR_TIE_CONST), my = es * (BLUR_TIE_SLOPE * sqrtf(fmaxf(d2.y, 0.f)) + BLUR_TIE_CONST);
                riskx = !(gap.x > mx) || !(d2.x > threshold_sq + mx);
                risky = !(gap.y > my) || !(d2.y > threshold_sq + my);
            }
            float c0x = 0.f, c1x = 0.f, c2x = 0.f, c0y = 0.f, c1y = 0.f, c2y = 0.f;
            bool ambx = false, amby = false;
            if (riskx) {
                float M, S2;
                blur_scan_exact(lr, lg, lbx, lum, s_pal, n, c0x, c1x, c2x, M, S2, ix);
                best.x = M;
                ambx = S2 == M || S2 < threshold_sq;
            }
            if (risky) {
                float M, S2;
                blur_scan_exact(lr, lg, lby, lum, s_pal, n, c0y, c1y, c2y, M, S2, iy);
                best.y = M;
                amby = S2 == M || S2 < threshold_sq;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t cell = b0 + 32u * half + lane;
                if (b0 + 32u * half >= size) break;  // uniform
                const float M = half ? best.y : best.x;   // exact where the cell is ambiguous
                const uint32_t I = half ? iy : ix;
                const bool in_row = cell < size;
                const bool amb = in_row && (half ? amby : ambx);
                uint32_t res = I;
                uint32_t am = __ballot_sync(FULL, amb);
                while (am) {
                    const uint32_t j = __ffs(am) - 1;
                    am &= am - 1;
                    // the hint that reaches cell j: the previous group's last result, or cell j-1's
                    const uint32_t prev = __shfl_sync(FULL, res, j == 0 ? 0 : j - 1);
                    const uint32_t h = j == 0 ? hint : prev;
                    if (lane == j && h != I) {
                        const float dh = blur_sq_dist(half ? c0y : c0x, half ? c1y : c1x, half ? c2y : c2x, s_pal + 3 * h);
                        if (dh < threshold_sq || dh == M) res = h;
                    }
                }
                const uint32_t valid = min(size - (b0 + 32u * half), 32u);
                hint = __shfl_sync(FULL, res, valid - 1);
                if (in_row) {
                    const float* t = out3 + 3 * res;
                    float* dst = colors + ((size_t)row * size + cell) * ch;
                    if (ch == 2) { dst[0] = t[1]; dst[1] = t[2]; }
                    else { dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2]; }
                }
            }
        }
        if (end_cur && lane == 0) {
            if (end_prev[row] != hint) atomicAdd(changed, 1u);
            end_cur[row] = hint;
        }
    }
}

// blur_inner / par_blur_inner, gaussian_blur.rs:100-162: out[p] = sum_k kw[k] * in[clamp(p+k-half)],
// accumulated in tap order with a rounded multiply and a rounded add per tap (no FMA: blur_add2).
//
// One thread produces 4 consecutive positions p0 .. p0+3 of TWO neighbouring columns, as float2
// accumulators (FMUL2 + FFMA2 via blur_add2: the two columns share the weight, which is the scalar
// operand). Output q's tap k reads position clamp(p0 + q + k - half); stepping j = q + k, every output
// needs the SAME input row clamp(p0 + j - half) at step j (with its own weight kw[j - q]), and each
// output still meets its taps in increasing k. So one 8-byte shared-memory load serves 8 outputs, and
// the sums are bit-identical to the one-output-at-a-time loop.
//
// Two things keep the loop uniform. (1) The tile carries `half` replicated rows before position 0 and
// half + 3 after position size-1, so the clamp is done once, at staging, and the input address just
// advances by `pitch`. (2) The weights are zero-padded (three before, up to six after): steps at which
// an output is outside its window add +0.0 * v, which leaves a sum unchanged bit for bit -- v is
// finite (a volume is either finite or NaN throughout: its cells are palette colours or blurs of
// them), s never is -0.0 (it starts at +0.0), and a NaN sum stays NaN.
constexpr int BLUR_R = 4;
constexpr uint32_t BLUR_PITCH = 20;   // the usual tile pitch (16 or 18 columns): a compile-time constant for the tap loop's addresses

// PITCH > 0: the tile pitch is known at compile time, so the four input rows of a trip are immediate offsets from ONE
// running address (the IMADs that computed them from a run-time pitch shared the FMA pipe with the packed arithmetic,
// where an IMAD between FFMA2s costs as much as an FFMA2: tools/microbench_fp32.cu). The weights arrive as one 16-byte
// load per trip; with the trip unrolled twice their four live registers rotate by renaming instead of by moves.
template <uint32_t PITCH>
__device__ __forceinline__ void blur_taps4x2(const float* __restrict__ row0, uint32_t pitch_rt, const float* __restrict__ kwz, int steps, float one,
                                             float2 (&out)[BLUR_R]) {
    // row0: the input row of step 0 (position p0 - half); kwz[j] = kw[j] for 0 <= j < klen, else 0 (j < steps + 4)
    float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
    // weights live in four registers named by j mod 4 (ra = kw[j] for j = 0 mod 4, ...): no shifting
    float ra, rb = 0.0f, rc = 0.0f, rd = 0.0f;
    // One step = four rounded products, then four rounded adds: written products first so that an add does not wait on
    // the multiply issued just before it (ptxas otherwise pairs them back to back through one temporary to save registers,
    // and the dependent FFMA2 sits out the FMUL2's latency: 'wait' was 21 % of the stall samples).
#define BLUR_STEP(V, W0, W1, W2, W3)                                                                            \
    {                                                                                                           \
        const float2 p0 = __fmul2_rn(make_float2(W0, W0), V), p1 = __fmul2_rn(make_float2(W1, W1), V);         \
        const float2 p2 = __fmul2_rn(make_float2(W2, W2), V), p3 = __fmul2_rn(make_float2(W3, W3), V);         \
        s0 = blur_add2(s0, p0, one); s1 = blur_add2(s1, p1, one); s2 = blur_add2(s2, p2, one); s3 = blur_add2(s3, p3, one); \
    }
    const uint32_t pitch = PITCH ? PITCH : pitch_rt;
    const float* q = row0;
    const float4* kq = reinterpret_cast<const float4*>(kwz);   // 16-byte aligned: the tiles in front of it are multiples of 16 bytes
    const uint32_t p4 = 4u * pitch;
#pragma unroll 2
    for (int j = 0; j < steps; j += 4, q += p4, ++kq) {
        const float2 v0 = *reinterpret_cast<const float2*>(q), v1 = *reinterpret_cast<const float2*>(q + pitch);
        const float2 v2 = *reinterpret_cast<const float2*>(q + 2u * pitch), v3 = *reinterpret_cast<const float2*>(q + 3u * pitch);
        const float4 w = *kq;
        ra = w.x; BLUR_STEP(v0, ra, rd, rc, rb);
        rb = w.y; BLUR_STEP(v1, rb, ra, rd, rc);
        rc = w.z; BLUR_STEP(v2, rc, rb, ra, rd);
        rd = w.w; BLUR_STEP(v3, rd, rc, rb, ra);
    }
#undef BLUR_STEP
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3;
}

// One blur pass along one axis of the [r][g][b][ch] volume, whatever the axis. Position p of the
// axis is `pstride` floats from position p-1. A CTA owns `cols` columns of one outer slab and
// stages all `size` positions of them (plus the replicated rows): tile[p + half][c], c fastest,
// `pitch` floats per row. Column c of a slab lives at (c / cpl) * run_stride + (c % cpl): runs of
// `cpl` contiguous floats (cpl = 0: the whole slab is one run).
//   B axis: pstride ch,         columns = (row, channel): cpl ch, run_stride size*ch, one slab
//   R axis: pstride size^2*ch,  columns = the size^2*ch floats of a (g, b, channel) plane, one slab
//   G axis: pstride size*ch,    columns = the size*ch floats of a (b, channel) line, one slab per r
// rows of a tile: the positions rounded up to whole blocks of BLUR_R (the last block's spare outputs read on), half before, half + 3 after
__host__ __device__ inline uint32_t blur_tile_rows(uint32_t size, int klen) { return (size + 3u) / 4u * 4u + 2u * (uint32_t)(klen / 2) + 3u; }
__host__ __device__ inline uint32_t blur_kwz_len(int klen) { return (uint32_t)((klen + 3 + 3) / 4 * 4 + 4); }

__device__ __forceinline__ void blur_cp_async4(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// ---- bulk-copy staging (the R and G passes: every tile row is one run of `cols` contiguous floats) ------------------
// One cp.async.bulk per tile row -- the copy engine moves the 64 bytes and signals an mbarrier -- instead of sixteen
// 4-byte cp.async with their address arithmetic (12 instructions each: 9 % of the kernel's instructions, and IMADs the
// tap loop's FMA pipe pays for). Needs 16-byte aligned rows on both sides: cols, total_cols, pstride, outer_stride
// multiples of 4 floats and a 16-byte aligned volume (BlurAxis::bulk, decided on the host).
__device__ __forceinline__ uint32_t blur_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void blur_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(blur_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void blur_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(blur_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void blur_mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = blur_smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void blur_bulk_row(float* smem_dst, const float* gmem_src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(blur_smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(blur_smem_u32(bar)) : "memory");
}

// Where tile t starts in the volume, and how many of its columns exist.
__device__ __forceinline__ size_t blur_tile_base(const BlurAxis& P, uint32_t t, uint32_t& ncols) {
    const uint32_t outer = t / P.groups, grp = t % P.groups;
    const uint32_t col0 = grp * P.cols;
    ncols = min(P.cols, P.total_cols - col0);
    const uint32_t cpl = P.cpl ? P.cpl : P.total_cols;
    return (size_t)outer * P.outer_stride + (size_t)(col0 / cpl) * P.run_stride + (col0 % cpl);
}

// Asynchronous staging of one tile's `size` positions (rows half .. half + size of `tile`).
template <bool BULK>
__device__ __forceinline__ void blur_stage_async(const float* __restrict__ src, float* __restrict__ s_body, const BlurAxis& P, uint32_t t,
                                                 unsigned long long* bar) {
    uint32_t ncols;
    const size_t base = blur_tile_base(P, t, ncols);
    if (BULK) {
        // the barrier's phase ends when its one arrival (below) has happened and every row's bytes have landed
        if (threadIdx.x == 0) blur_mbar_expect_tx(bar, P.size * ncols * (uint32_t)sizeof(float));
        // one row per thread and trip (the compiler serialises a warp's 32 copies through an elect loop: UBLKCP takes its
        // addresses from uniform registers; issuing them from one lane per warp in a uniform loop was measured slower)
        const float* q = src + base;
        const uint32_t bytes = ncols * (uint32_t)sizeof(float);
        for (uint32_t p = threadIdx.x; p < P.size; p += blockDim.x) blur_bulk_row(s_body + p * P.pitch, q + (size_t)p * P.pstride, bytes, bar);
        return;
    }
    if (P.cpl == 0) {
        // the columns are contiguous floats: thread = (row of the wave, column), waves of 256 / cols rows
        const uint32_t c = threadIdx.x % P.cols, prow = threadIdx.x / P.cols, pstep = blockDim.x / P.cols;
        if (c < ncols && prow < pstep) {
            const float* q = src + base + c;
            for (uint32_t p = prow; p < P.size; p += pstep) blur_cp_async4(s_body + p * P.pitch + c, q + (size_t)p * P.pstride);
        }
    } else {
        // runs of cpl (= channels: 2 or 3) columns, each run one contiguous row of size * cpl floats
        const uint32_t rcols = P.cpl, nruns = (ncols + rcols - 1) / rcols, per_run = P.size * rcols;
        for (uint32_t run = 0; run < nruns; ++run) {
            const float* q = src + base + (size_t)run * P.run_stride;
            float* tl = s_body + run * rcols;
            for (uint32_t x = threadIdx.x; x < per_run; x += blockDim.x) {
                const uint32_t p = rcols == 3 ? x / 3u : (rcols == 2 ? x >> 1 : x / rcols), cc = x - p * rcols;
                if (run * rcols + cc < ncols) blur_cp_async4(tl + p * P.pitch + cc, q + x);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// Persistent CTAs walk tiles blockIdx.x, blockIdx.x + gridDim.x, ...; two tile buffers: while one is
// blurred the next tile's rows arrive asynchronously (the staging loads were 43 % of the stall
// samples of the single-buffer version) -- by cp.async.bulk rows signalling an mbarrier per buffer (BULK), or by
// 4-byte cp.async groups where rows are not 16-byte aligned runs (the B pass, odd sizes).
template <uint32_t PITCH, bool BULK>
__global__ void __launch_bounds__(256, 4) k_blur_axis(const float* __restrict__ src, float* __restrict__ dst, const BlurAxis P, uint32_t ntiles,
                                                    const float* __restrict__ kernel, int klen, float one) {
    extern __shared__ __align__(16) float s_blur[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int half = klen / 2;
    const uint32_t trows = blur_tile_rows(P.size, klen);
    const size_t tile_floats = (size_t)trows * P.pitch;
    float* s_kwz = s_blur + 2 * tile_floats;                   // zero-padded weights behind the two tiles
    for (uint32_t i = threadIdx.x; i < blur_kwz_len(klen); i += blockDim.x) s_kwz[i] = (int)i < klen ? kernel[i] : 0.0f;
    if (BULK) {
        if (threadIdx.x == 0) {
            blur_mbar_init(&s_bar[0], 1);
            blur_mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const int steps = klen ? (klen + BLUR_R - 1 + 3) / 4 * 4 : 0;   // klen + 3 steps, in fours; no taps: the sums stay 0.0
    // the thread's place in a full tile (all but a slab's last group): worked out once, not per tile
    const uint32_t full_pairs = (P.cols + 1) / 2, full_pair = threadIdx.x % full_pairs, full_pb0 = threadIdx.x / full_pairs, full_pbstep = blockDim.x / full_pairs;
    const uint32_t pad_w = (P.cols + 1u) & ~1u, pad_c = threadIdx.x % pad_w, pad_k = threadIdx.x / pad_w, pad_kstep = blockDim.x / pad_w;
    uint32_t t = blockIdx.x, buf = 0, it = 0;
    if (t < ntiles) blur_stage_async<BULK>(src, s_blur + (size_t)half * P.pitch, P, t, &s_bar[0]);
    for (; t < ntiles; t += gridDim.x, buf ^= 1u, ++it) {
        float* s_tile = s_blur + buf * tile_floats;
        float* s_body = s_tile + (size_t)half * P.pitch;        // row of position 0
        const uint32_t tn = t + gridDim.x;
        if (BULK) {
            // the other buffer was last read before the barrier that ended the previous trip
            if (tn < ntiles) blur_stage_async<BULK>(src, s_blur + (buf ^ 1u) * tile_floats + (size_t)half * P.pitch, P, tn, &s_bar[buf ^ 1u]);
            blur_mbar_wait(&s_bar[buf], (it >> 1) & 1u);        // this buffer's (it / 2)-th use
        } else {
            if (tn < ntiles) {
                blur_stage_async<BULK>(src, s_blur + (buf ^ 1u) * tile_floats + (size_t)half * P.pitch, P, tn, nullptr);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
        }
        uint32_t ncols;
        const size_t base = blur_tile_base(P, t, ncols);
        {
            // the clamp of the reference, once: rows before position 0 repeat it, rows after size-1 repeat that
            const uint32_t npad = trows - P.size, wcols = (ncols + 1u) & ~1u;
            const uint32_t c = ncols == P.cols ? pad_c : threadIdx.x % wcols, k0 = ncols == P.cols ? pad_k : threadIdx.x / wcols;
            const uint32_t kstep = ncols == P.cols ? pad_kstep : blockDim.x / wcols;
            if (k0 < kstep) {
                const float first = s_body[c], last = s_body[(P.size - 1) * P.pitch + c];
                for (uint32_t k = k0; k < npad; k += kstep) {
                    const bool lo = k < (uint32_t)half;
                    s_tile[(lo ? k : P.size + k) * P.pitch + c] = lo ? first : last;   // tile rows [0, half) and [half + size, trows)
                }
            }
        }
        __syncthreads();
        const uint32_t pblocks = (P.size + BLUR_R - 1) / BLUR_R;
        uint32_t pair = full_pair, pb0 = full_pb0, pbstep = full_pbstep;
        if (ncols != P.cols) {   // the last group of a slab may be narrower
            const uint32_t npairs = (ncols + 1) / 2;
            pair = threadIdx.x % npairs; pb0 = threadIdx.x / npairs; pbstep = blockDim.x / npairs;
        }
        if (pb0 < pbstep) {  // blockDim.x is not a multiple of npairs: the remainder threads idle
            const uint32_t c = 2 * pair;
            size_t offA, offB;
            if (P.cpl == 0) {   // contiguous columns
                offA = base + c; offB = offA + 1;
            } else {
                offA = base + (size_t)(c / P.cpl) * P.run_stride + (c % P.cpl);
                offB = base + (size_t)((c + 1) / P.cpl) * P.run_stride + ((c + 1) % P.cpl);
            }
            const bool hasB = c + 1 < ncols;
            // the pair's two columns are neighbours in the volume and 8-byte aligned: one store per position
            const bool pair_store = hasB && offB == offA + 1 && ((offA | P.pstride) & 1u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0;
            for (uint32_t pb = pb0; pb < pblocks; pb += pbstep) {
                float2 o[BLUR_R];
                // step 0 reads position pb*4 - half = tile row pb*4
                blur_taps4x2<PITCH>(s_tile + (size_t)(pb * BLUR_R) * P.pitch + c, P.pitch, s_kwz, steps, one, o);
                float* d = dst + offA + (size_t)(pb * BLUR_R) * P.pstride;
#pragma unroll
                for (int q = 0; q < BLUR_R; ++q, d += P.pstride) {
                    const uint32_t pq = pb * BLUR_R + q;
                    if (pq < P.size) {
                        if (pair_store) {
                            *reinterpret_cast<float2*>(d) = o[q];
                        } else {
                            *d = o[q].x;
                            if (hasB) dst[offB + (size_t)pq * P.pstride] = o[q].y;
                        }
                    }
                }
            }
        }
        __syncthreads();  // this buffer is refilled by the next iteration's staging
    }
}

// colors_to_lut / cell_to_rgb, gaussian_blur.rs:433-500: the [r][g][b] volume to Hald order
// (b outer, g, r inner), entries [begin, begin + count). A tile is BLUR_TR r x BLUR_TB b cells of one g,
// transposed through shared memory: a warp reads one r's run of BLUR_TB * ch contiguous floats (384 bytes), a
// thread converts four cells, and a warp's stores are the 96 contiguous bytes of 32 r entries of one b -- as 24 words
// through a per-warp staging row when they are aligned, else as bytes. CTAs are persistent and walk tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...; the NEXT tile's floats are loaded into registers before the current tile is converted
// (twelve loads in flight per thread: the kernel is bound by how many bytes an SM keeps in flight, not by instructions --
// quadrupling the tile and halving the instructions per cell left round 1's 125 us where they were).
struct BlurTile { uint32_t g, r0, b0; };
__device__ __forceinline__ BlurTile blur_lut_tile(uint32_t t, uint32_t tiles_r, float inv_tr, uint32_t size, float inv_size, uint32_t b_first) {
    // t = (tb * size + g) * tiles_r + tr   (float estimate + correction: t < 2^24)
    uint32_t q1 = (uint32_t)(__uint2float_rz(t) * inv_tr);
    int rem = (int)(t - q1 * tiles_r);
    if (rem < 0) { --q1; rem += (int)tiles_r; }
    if (rem >= (int)tiles_r) { ++q1; rem -= (int)tiles_r; }
    uint32_t tb = (uint32_t)(__uint2float_rz(q1) * inv_size);
    int g_ = (int)(q1 - tb * size);
    if (g_ < 0) { --tb; g_ += (int)size; }
    if (g_ >= (int)size) { ++tb; g_ -= (int)size; }
    BlurTile T;
    T.g = (uint32_t)g_; T.r0 = (uint32_t)rem * BLUR_TR; T.b0 = b_first + tb * BLUR_TB;
    return T;
}
constexpr int BLUR_LUT_ROWS = BLUR_TR / 8, BLUR_LUT_X = (BLUR_TB * 3 + 31) / 32;   // per thread: rows (one per trip of its warp) x floats per row

__device__ __forceinline__ void blur_lut_fetch(const float* __restrict__ colors, uint32_t size, uint32_t ch, const BlurTile& T, uint32_t warp, uint32_t lane,
                                               float (&v)[BLUR_LUT_ROWS][BLUR_LUT_X]) {
    const uint32_t run = T.b0 < size ? min((uint32_t)BLUR_TB, size - T.b0) * ch : 0u;   // floats of one r's cells in this tile
#pragma unroll
    for (int i = 0; i < BLUR_LUT_ROWS; ++i) {
        const uint32_t lr = warp + 8u * i;
        const bool row = T.r0 + lr < size;
        const float* q = colors + (((size_t)(row ? T.r0 + lr : 0u) * size + T.g) * size + (run ? T.b0 : 0u)) * ch;
#pragma unroll
        for (int j = 0; j < BLUR_LUT_X; ++j) {
            const uint32_t x = lane + 32u * j;
            v[i][j] = (row && x < run) ? __ldcs(q + x) : 0.f;   // read once
        }
    }
}

// One cell: the blurred Oklab value (preserve: the cell's own L) back to packed sRGB.
template <int CH>
__device__ __forceinline__ uint32_t blur_cell_rgb(const SmemTables& st, const float* __restrict__ cell, float own_r, float own_g, float own_b) {
    Lab o;
    if (CH == 2) {
        const Lab own = linear_to_oklab_exact(own_r, own_g, own_b);
        o.l = own.l; o.a = cell[0]; o.b = cell[1];
    } else {
        o.l = cell[0]; o.a = cell[1]; o.b = cell[2];
    }
    float r, g, b;
    oklab_to_linear_exact(o, r, g, b);
    return srgb8_pack3(f32_to_srgb8_wide(r, st.encode2), f32_to_srgb8_wide(g, st.encode2), f32_to_srgb8_wide(b, st.encode2));
}

template <int CH>
__global__ void __launch_bounds__(256) k_blur_to_lut(const float* __restrict__ colors, uint32_t size, const uint8_t* __restrict__ chan,
                                                      uint8_t* __restrict__ out_rgb, unsigned long long begin, unsigned long long count,
                                                      uint32_t b_first, uint32_t tiles_r, uint32_t ntiles, const ColorTables* __restrict__ tables) {
    __shared__ SmemTables st;
    __shared__ float s_c[BLUR_TR][BLUR_TB * 3 + 1];
    __shared__ __align__(16) uint8_t s_out[8][96];
    static_assert(BLUR_TR == 32 && BLUR_TB % 8 == 0, "a warp stores the 32 r entries of one b; eight warps share the b cells");
    stage_tables(st, tables);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float inv_tr = 1.0f / (float)tiles_r, inv_size = 1.0f / (float)size;
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out_rgb) & 3u) == 0 && (size & 3u) == 0;
    const unsigned long long plane = (unsigned long long)size * size, end = begin + count;
    float v[BLUR_LUT_ROWS][BLUR_LUT_X];
    uint32_t t = blockIdx.x;
    BlurTile T = blur_lut_tile(t < ntiles ? t : 0u, tiles_r, inv_tr, size, inv_size, b_first);
    if (t < ntiles) blur_lut_fetch(colors, size, (uint32_t)CH, T, warp, lane, v);
    for (; t < ntiles; t += gridDim.x) {
        __syncthreads();  // the previous tile's s_c is no longer read
#pragma unroll
        for (int i = 0; i < BLUR_LUT_ROWS; ++i)
#pragma unroll
            for (int j = 0; j < BLUR_LUT_X; ++j) s_c[warp + 8u * i][lane + 32u * j] = v[i][j];
        __syncthreads();
        const BlurTile C = T;
        const uint32_t tn = t + gridDim.x;
        if (tn < ntiles) {
            T = blur_lut_tile(tn, tiles_r, inv_tr, size, inv_size, b_first);
            blur_lut_fetch(colors, size, (uint32_t)CH, T, warp, lane, v);   // in flight while this tile is converted
        }
        const uint32_t g = C.g, ri = C.r0 + lane;
        if (C.b0 >= size) continue;
        const bool r_ok = ri < size;
        const float own_r = (CH == 2 && r_ok) ? st.decode[chan[ri]] : 0.f, own_g = CH == 2 ? st.decode[chan[g]] : 0.f;
        const uint32_t nb = min((uint32_t)BLUR_TB, size - C.b0);
        // entry of (b0 + warp, g, r0): the warp's first row of this tile; its next one is 8 planes on
        unsigned long long px0 = ((unsigned long long)(C.b0 + warp) * size + g) * size + C.r0;
        const unsigned long long tile_first = ((unsigned long long)C.b0 * size + g) * size + C.r0;
        const unsigned long long tile_last = ((unsigned long long)(C.b0 + nb - 1u) * size + g) * size + min(C.r0 + 31u, size - 1u);
        // the usual tile: whole rows of 32 entries, all inside the launch's range, word-aligned in the CLUT -- no test per cell
        if (out_aligned && C.r0 + 31u < size && tile_first >= begin && tile_last < end) {
            uint8_t* sb = s_out[warp];
            uint8_t* dst = out_rgb + 3ull * px0;
            const size_t step = (size_t)(24ull * plane);   // 8 planes of 3-byte entries
#pragma unroll 1
            for (uint32_t bb = warp; bb < nb; bb += 8u, dst += step) {
                const float own_b = CH == 2 ? st.decode[chan[C.b0 + bb]] : 0.f;
                const uint32_t packed = blur_cell_rgb<CH>(st, &s_c[lane][bb * CH], own_r, own_g, own_b);
                __syncwarp();   // the previous row's words have been read
                sb[3u * lane] = (uint8_t)packed; sb[3u * lane + 1] = (uint8_t)(packed >> 8); sb[3u * lane + 2] = (uint8_t)(packed >> 16);
                __syncwarp();
                if (lane < 24u) reinterpret_cast<uint32_t*>(dst)[lane] = reinterpret_cast<const uint32_t*>(sb)[lane];
            }
            continue;
        }
#pragma unroll 1
        for (uint32_t bb = warp; bb < nb; bb += 8u, px0 += 8ull * plane) {
            const unsigned long long px = px0 + lane;
            if (!(r_ok && px >= begin && px < end)) continue;
            const float own_b = CH == 2 ? st.decode[chan[C.b0 + bb]] : 0.f;
            const uint32_t packed = blur_cell_rgb<CH>(st, &s_c[lane][bb * CH], own_r, own_g, own_b);
            uint8_t* q = out_rgb + 3ull * px;
            q[0] = (uint8_t)packed; q[1] = (uint8_t)(packed >> 8); q[2] = (uint8_t)(packed >> 16);
        }
    }
}

}  // namespace lutgen
