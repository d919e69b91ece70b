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
#include "kernels_exact.cuh"

namespace lutgen {

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
// every (r, g) row and then walks b). One warp per row, 256 cells at a time:
//   parallel part: each lane converts its cells and finds (M, I) = (smallest distance, first index
//                  reaching it) over the whole palette;
//   serial part  : lane 0 replays the hint chain. With hint h the reference returns h when
//                  d(h) < threshold or d(h) == M, else I -- one distance per cell whose hint differs
//                  from I, nothing otherwise.
// `chan[i]` = (i as f32 * scale).round() as u8 (l.206-210), filled on the host.
constexpr int BLUR_ROW_WARPS = 8;
constexpr int BLUR_SEG = 256;

__global__ void __launch_bounds__(BLUR_ROW_WARPS * 32) k_blur_nn_volume(float* __restrict__ colors, uint32_t size, uint32_t ch, const uint8_t* __restrict__ chan,
                                                                         const float* __restrict__ pal3, const float* __restrict__ out3, uint32_t n,
                                                                         float lum, float threshold_sq, const ColorTables* __restrict__ tables) {
    __shared__ float s_decode[256];
    __shared__ float s_col[BLUR_ROW_WARPS][3][BLUR_SEG];
    __shared__ float s_m[BLUR_ROW_WARPS][BLUR_SEG];
    __shared__ uint32_t s_i[BLUR_ROW_WARPS][BLUR_SEG];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_decode[i] = tables->decode[i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows = size * size;
    for (uint32_t row = blockIdx.x * BLUR_ROW_WARPS + warp; row < rows; row += gridDim.x * BLUR_ROW_WARPS) {
        const uint32_t r = row / size, g = row % size;
        const float lr = s_decode[chan[r]], lg = s_decode[chan[g]];
        uint32_t hint = 0;
        for (uint32_t b0 = 0; b0 < size; b0 += BLUR_SEG) {
            const uint32_t seg = min((uint32_t)BLUR_SEG, size - b0);
            for (uint32_t j = lane; j < seg; j += 32) {
                Lab o = linear_to_oklab_exact(lr, lg, s_decode[chan[b0 + j]]);
                float c0 = __fmul_rn(o.l, lum), c1 = o.a, c2 = o.b;
                float best = INFINITY;
                uint32_t bi = 0;
                for (uint32_t t = 0; t < n; ++t) {
                    float d = blur_sq_dist(c0, c1, c2, pal3 + 3 * t);
                    if (d < best) { best = d; bi = t; }
                }
                s_col[warp][0][j] = c0; s_col[warp][1][j] = c1; s_col[warp][2][j] = c2;
                s_m[warp][j] = best; s_i[warp][j] = bi;
            }
            __syncwarp();
            if (lane == 0) {
                for (uint32_t j = 0; j < seg; ++j) {
                    uint32_t I = s_i[warp][j];
                    if (hint != I) {
                        float dh = blur_sq_dist(s_col[warp][0][j], s_col[warp][1][j], s_col[warp][2][j], pal3 + 3 * hint);
                        if (!(dh < threshold_sq || dh == s_m[warp][j])) hint = I;
                    }
                    s_i[warp][j] = hint;
                }
            }
            hint = __shfl_sync(0xffffffffu, hint, 0);
            __syncwarp();
            for (uint32_t j = lane; j < seg; j += 32) {
                const float* t = out3 + 3 * s_i[warp][j];
                float* dst = colors + ((size_t)row * size + b0 + j) * ch;
                if (ch == 2) { dst[0] = t[1]; dst[1] = t[2]; }
                else { dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2]; }
            }
            __syncwarp();
        }
    }
}

// blur_inner / par_blur_inner, gaussian_blur.rs:100-162: out[p] = sum_k kw[k] * in[clamp(p+k-half)],
// accumulated in tap order with a rounded multiply and a rounded add per tap.
//
// One thread produces BLUR_R consecutive outputs p0 .. p0+R-1 of a line. Output q's tap k reads
// input position clamp(p0 + q + k - half); stepping j = q + k, every output needs the SAME input
// in[clamp(p0 + j - half)] at step j (with its own weight kw[j - q]), and each output still meets its
// taps in increasing k. So one shared-memory load and one clamp serve R outputs, and the sums are
// bit-identical to the one-output-at-a-time loop.
constexpr int BLUR_R = 4;

__device__ __forceinline__ void blur_taps4(const float* __restrict__ line, uint32_t line_stride, int pos0, int maxpos,
                                           const float* __restrict__ kw, int klen, int half, float (&out)[BLUR_R]) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    // weights live in four registers named by j mod 4 (ra = kw[j] for j = 0 mod 4, ...): no shifting
    float ra = 0.0f, rb = 0.0f, rc = 0.0f, rd = 0.0f;
    const int steps = klen + BLUR_R - 1;
    // input at step j; away from the line's ends the clamp is the identity and the address is affine
    const bool interior = pos0 - half >= 0 && pos0 + steps - 1 - half <= maxpos;
    const float* ptr = line + (ptrdiff_t)(pos0 - half) * (ptrdiff_t)line_stride;
#define BLUR_IN(J) (interior ? ptr[(uint32_t)(J) * line_stride] : line[(uint32_t)min(max(pos0 + (J) - half, 0), maxpos) * line_stride])
#define BLUR_ACC(S, W, V) S = __fadd_rn(S, __fmul_rn(W, V))
    // head: steps 0..3 (outputs join one by one), generic form
    int j = 0;
    for (; j < 4 && j < steps; ++j) {
        float v = BLUR_IN(j);
        float w = j < klen ? kw[j] : 0.0f;
        if (j == 0) ra = w; else if (j == 1) rb = w; else if (j == 2) rc = w; else rd = w;
        if (j < klen) BLUR_ACC(s0, w, v);
        if (j >= 1 && j - 1 < klen) BLUR_ACC(s1, (j == 1 ? ra : j == 2 ? rb : rc), v);
        if (j >= 2 && j - 2 < klen) BLUR_ACC(s2, (j == 2 ? ra : rb), v);
        if (j >= 3 && j - 3 < klen) BLUR_ACC(s3, ra, v);
    }
    // body: four steps per iteration while every output is active (j + 3 < klen), no predicates
    for (; j + 3 < klen; j += 4) {
        float v0 = BLUR_IN(j), v1 = BLUR_IN(j + 1), v2 = BLUR_IN(j + 2), v3 = BLUR_IN(j + 3);
        ra = kw[j];     BLUR_ACC(s0, ra, v0); BLUR_ACC(s1, rd, v0); BLUR_ACC(s2, rc, v0); BLUR_ACC(s3, rb, v0);
        rb = kw[j + 1]; BLUR_ACC(s0, rb, v1); BLUR_ACC(s1, ra, v1); BLUR_ACC(s2, rd, v1); BLUR_ACC(s3, rc, v1);
        rc = kw[j + 2]; BLUR_ACC(s0, rc, v2); BLUR_ACC(s1, rb, v2); BLUR_ACC(s2, ra, v2); BLUR_ACC(s3, rd, v2);
        rd = kw[j + 3]; BLUR_ACC(s0, rd, v3); BLUR_ACC(s1, rc, v3); BLUR_ACC(s2, rb, v3); BLUR_ACC(s3, ra, v3);
    }
    // tail: remaining steps (outputs leave one by one), generic form; j is a multiple of 4 here
    for (; j < steps; ++j) {
        float v = BLUR_IN(j);
        const int m = j & 3;
        float w = j < klen ? kw[j] : 0.0f;
        if (m == 0) ra = w; else if (m == 1) rb = w; else if (m == 2) rc = w; else rd = w;
        // weight of output q at step j is kw[j - q], held in the register named (j - q) mod 4
        float w1 = m == 0 ? rd : m == 1 ? ra : m == 2 ? rb : rc;
        float w2 = m == 0 ? rc : m == 1 ? rd : m == 2 ? ra : rb;
        float w3 = m == 0 ? rb : m == 1 ? rc : m == 2 ? rd : ra;
        if (j < klen) BLUR_ACC(s0, w, v);
        if (j >= 1 && j - 1 < klen) BLUR_ACC(s1, w1, v);
        if (j >= 2 && j - 2 < klen) BLUR_ACC(s2, w2, v);
        if (j >= 3 && j - 3 < klen) BLUR_ACC(s3, w3, v);
    }
#undef BLUR_IN
#undef BLUR_ACC
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3;
}

// Blur along the innermost (b) axis: a CTA stages `lines` consecutive (r, g) rows -- each a
// contiguous run of size*ch floats -- in shared memory.
__global__ void __launch_bounds__(256) k_blur_inner(const float* __restrict__ src, float* __restrict__ dst, uint32_t size, uint32_t ch, uint32_t lines,
                                                     const float* __restrict__ kernel, int klen) {
    extern __shared__ float s_blur[];
    float* s_kw = s_blur;
    float* s_tile = s_blur + klen;
    const uint32_t row_len = size * ch, rows = size * size;
    const uint32_t row0 = blockIdx.x * lines, nrows = min(lines, rows - row0);
    for (int i = threadIdx.x; i < klen; i += blockDim.x) s_kw[i] = kernel[i];
    const float* base = src + (size_t)row0 * row_len;
    for (uint32_t i = threadIdx.x; i < nrows * row_len; i += blockDim.x) s_tile[i] = base[i];
    __syncthreads();
    const int half = klen / 2, maxpos = (int)size - 1;
    const uint32_t pblocks = (size + BLUR_R - 1) / BLUR_R, per_line = pblocks * ch;
    for (uint32_t i = threadIdx.x; i < nrows * per_line; i += blockDim.x) {
        uint32_t line = i / per_line, within = i % per_line, pb = within / ch, c = within % ch;
        float o[BLUR_R];
        blur_taps4(s_tile + line * row_len + c, ch, (int)(pb * BLUR_R), maxpos, s_kw, klen, half, o);
        float* d = dst + (size_t)(row0 + line) * row_len + c;
#pragma unroll
        for (int q = 0; q < BLUR_R; ++q)
            if (pb * BLUR_R + q < size) d[(pb * BLUR_R + q) * ch] = o[q];
    }
}

// Blur along a strided axis (g: stride size cells, r: stride size^2 cells). A CTA owns `cols`
// consecutive floats of the fastest-varying remainder (a run of b cells x channels, contiguous in
// memory) and stages all `size` positions of the axis for them: tile[pos][col].
__global__ void __launch_bounds__(256) k_blur_strided(const float* __restrict__ src, float* __restrict__ dst, uint32_t size, uint32_t axis_stride_floats,
                                                       uint32_t groups_per_outer, uint32_t outer_stride_floats, uint32_t inner_floats, uint32_t cols,
                                                       const float* __restrict__ kernel, int klen) {
    extern __shared__ float s_blur[];
    float* s_kw = s_blur;
    float* s_tile = s_blur + klen;
    const uint32_t outer = blockIdx.x / groups_per_outer, grp = blockIdx.x % groups_per_outer;
    const uint32_t col0 = grp * cols, ncols = min(cols, inner_floats - col0);
    const size_t base = (size_t)outer * outer_stride_floats + col0;
    for (int i = threadIdx.x; i < klen; i += blockDim.x) s_kw[i] = kernel[i];
    for (uint32_t i = threadIdx.x; i < size * ncols; i += blockDim.x) {
        uint32_t pos = i / ncols, col = i % ncols;
        s_tile[pos * cols + col] = src[base + (size_t)pos * axis_stride_floats + col];
    }
    __syncthreads();
    const int half = klen / 2, maxpos = (int)size - 1;
    const uint32_t pblocks = (size + BLUR_R - 1) / BLUR_R;
    for (uint32_t i = threadIdx.x; i < pblocks * ncols; i += blockDim.x) {
        uint32_t pb = i / ncols, col = i % ncols;
        float o[BLUR_R];
        blur_taps4(s_tile + col, cols, (int)(pb * BLUR_R), maxpos, s_kw, klen, half, o);
#pragma unroll
        for (int q = 0; q < BLUR_R; ++q)
            if (pb * BLUR_R + q < size) dst[base + (size_t)(pb * BLUR_R + q) * axis_stride_floats + col] = o[q];
    }
}

// colors_to_lut / cell_to_rgb, gaussian_blur.rs:433-500: Hald order (b outer, g, r inner), entries
// [begin, begin + count) only.
__global__ void __launch_bounds__(256) k_blur_to_lut(const float* __restrict__ colors, uint32_t size, uint32_t ch, const uint8_t* __restrict__ chan,
                                                      uint8_t* __restrict__ out_rgb, unsigned long long begin, unsigned long long count,
                                                      const ColorTables* __restrict__ tables) {
    __shared__ SmemTables st;
    stage_tables(st, tables);
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long px = begin + i;
    uint32_t ri = (uint32_t)(px % size), gi = (uint32_t)((px / size) % size), bi = (uint32_t)(px / ((unsigned long long)size * size));
    const float* c = colors + (((size_t)ri * size + gi) * size + bi) * ch;
    Lab o;
    if (ch == 2) {
        Lab own = linear_to_oklab_exact(st.decode[chan[ri]], st.decode[chan[gi]], st.decode[chan[bi]]);
        o.l = own.l; o.a = c[0]; o.b = c[1];
    } else {
        o.l = c[0]; o.a = c[1]; o.b = c[2];
    }
    float r, g, b;
    oklab_to_linear_exact(o, r, g, b);
    uint8_t* q = out_rgb + 3ull * px;
    q[0] = (uint8_t)f32_to_srgb8(r, st.encode);
    q[1] = (uint8_t)f32_to_srgb8(g, st.encode);
    q[2] = (uint8_t)f32_to_srgb8(b, st.encode);
}

}  // namespace lutgen
