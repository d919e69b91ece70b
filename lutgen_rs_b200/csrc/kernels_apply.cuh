// kernels_apply.cuh -- Hald CLUT application (identity::correct_pixel /
// correct_image_with_level, crates/lib/src/identity.rs:40-78) and the
// GaussianSamplingRemapper kernels (gaussian_sample.rs:49-76).
#pragma once
#include "kernels_exact.cuh"

namespace lutgen {

// RGB8 CLUT -> one 32-bit word per entry (r | g<<8 | b<<16) so the gather is a
// single aligned 4-byte load instead of three byte loads.
__global__ void k_expand_lut(const uint8_t* __restrict__ lut_rgb, size_t n, uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = lut_rgb + 3 * i;
    out[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

// correct_pixel, identity.rs:40-52: q(c) = c*(cs-1)/255 (floor); the pixel at
// x = r%cs + (g%level)*cs, y = b*level + g/level of the level^3-wide image is
// flat entry (b*cs + g)*cs + r. Alpha (top byte) is carried through.
// What bounds the kernel on pixels that share no CLUT cells (uniform random input) is not HBM and not the L2
// slices but the SM's request port into the crossbar: one 32-byte sector request per gather that misses L1, at most
// one request per clock and SM (ncu: l1tex__m_l1tex2xbar_req_cycles_active 91 % of peak at 254 Gpix/s, lts__throughput
// 68 %, DRAM 21 % -- profiles/r2_apply_k_apply_ncu.txt). The L1 cannot take the gathers off that port: it tags
// 128-byte lines, a scattered 4-byte gather fills one 32-byte sector of a line, so its ~196 KB hold ~49 KB of distinct
// CLUT sectors -- 5 % of a level-8 CLUT, which is the hit rate measured, with or without evict-last on the CLUT loads,
// no-allocate on the frame's own loads and stores and the largest L1 carve-out (tried in round 2, no change: 254 vs
// 250 Gpix/s). Pixels of real images share cells with their neighbours and run at 85-89 % of HBM bandwidth.
__device__ __forceinline__ uint32_t apply_one(uint32_t p, const uint32_t* __restrict__ lut, uint32_t cs, uint32_t csm1) {
    uint32_t r = (p & 255u) * csm1 / 255u;
    uint32_t g = ((p >> 8) & 255u) * csm1 / 255u;
    uint32_t b = ((p >> 16) & 255u) * csm1 / 255u;
    uint32_t v = __ldg(lut + ((b * cs + g) * cs + r));
    return (v & 0x00ffffffu) | (p & 0xff000000u);
}

// One thread = 4 consecutive pixels (one 16-byte load, four gathers, one
// 16-byte store). `px` is 4-byte aligned; `head` pixels precede the first
// 16-byte boundary and are handled, with the tail, by scalar lanes.
__global__ void __launch_bounds__(256) k_apply(uint32_t* __restrict__ px, size_t npix, size_t head, size_t nvec,
                                               const uint32_t* __restrict__ lut, uint32_t cs, uint32_t csm1) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4* v = reinterpret_cast<uint4*>(px + head);
    for (size_t i = tid; i < nvec; i += stride) {
        uint4 p = v[i];
        p.x = apply_one(p.x, lut, cs, csm1);
        p.y = apply_one(p.y, lut, cs, csm1);
        p.z = apply_one(p.z, lut, cs, csm1);
        p.w = apply_one(p.w, lut, cs, csm1);
        v[i] = p;
    }
    // scalar head + tail
    size_t tail_begin = head + nvec * 4;
    size_t nscalar = head + (npix - tail_begin);
    for (size_t i = tid; i < nscalar; i += stride) {
        size_t k = i < head ? i : tail_begin + (i - head);
        px[k] = apply_one(px[k], lut, cs, csm1);
    }
}

// correct_pixel for independent RGB triples (CLI `patch`, cli/main.rs:1028-1042).
__global__ void k_correct_pixels(const uint8_t* __restrict__ in, size_t n, const uint32_t* __restrict__ lut, uint32_t cs, uint32_t csm1,
                                 uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = (uint32_t)in[3 * i] | ((uint32_t)in[3 * i + 1] << 8) | ((uint32_t)in[3 * i + 2] << 16);
    uint32_t v = apply_one(p, lut, cs, csm1);
    out[3 * i] = (uint8_t)(v & 255u);
    out[3 * i + 1] = (uint8_t)((v >> 8) & 255u);
    out[3 * i + 2] = (uint8_t)((v >> 16) & 255u);
}

// ---------------------------------------------------------------------------
// Gaussian sampling.
//
// The reference re-seeds one RNG per pixel (gaussian_sample.rs:52), so every
// pixel sees the SAME `iterations x 4` normals: the remapper is the stencil
//     out(p) = round( sum_k NN(sat_u8(round(p + noise_k))) / K )
// where NN is the nearest-neighbour remap, a pure function of the 24-bit
// colour. We therefore (1) tabulate NN over the whole sRGB cube once per
// remapper (k_nn_exact in FRONT_TABLE mode, 2^24 x u32), (2) turn the noise
// table into K integer offsets (validated on the host to reproduce
// sat_u8(round(c + n)) for every c), and (3) gather-and-average.
// ---------------------------------------------------------------------------

// Philox4x32-10 (Salmon et al. 2011), the library's own counter-based stream.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// noise[k][0..4) = mean + std * N(0,1): counter (k, pair, 0, 0), key = seed;
// two Box-Muller pairs per iteration, 53-bit uniforms.
__global__ void k_philox_normals(double* __restrict__ noise, size_t iterations, double mean, double std_dev, unsigned long long seed) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= iterations * 2) return;
    size_t k = t >> 1;
    uint32_t pair = (uint32_t)(t & 1);
    uint32_t c[4] = {(uint32_t)k, (uint32_t)(k >> 32), pair, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned long long a = ((unsigned long long)c[0] << 32) | c[1], b = ((unsigned long long)c[2] << 32) | c[3];
    double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    noise[4 * k + 2 * pair] = mean + std_dev * (rad * co);
    noise[4 * k + 2 * pair + 1] = mean + std_dev * (rad * s);
}

__device__ __forceinline__ uint32_t clamp255(int v) { return (uint32_t)min(max(v, 0), 255); }

// GaussianSamplingRemapper::remap_pixel, gaussian_sample.rs:49-76.
// offsets: K x short4 (r, g, b, unused) in the reference's iteration order.
// EXACT_F64 = false: K is a power of two, so px/K and every partial sum are
//   exact in f64 and the reference's `mean.round()` equals (S + K/2) >> log2 K.
// EXACT_F64 = true : replay the reference's f64 recurrence mean += px/total in
//   iteration order (divk[c] = c / K, correctly rounded like the reference's).
template <int MODE, bool EXACT_F64>
__global__ void __launch_bounds__(256) k_sampling(Front f, const uint32_t* __restrict__ nn_table, const short4* __restrict__ offsets,
                                                  uint32_t iterations, uint32_t log2k) {
    constexpr int CHUNK = 1024;
    __shared__ short4 s_off[CHUNK];
    __shared__ double s_div[EXACT_F64 ? 256 : 1];
    if (EXACT_F64) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_div[i] = __ddiv_rn((double)i, (double)iterations);
    }
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    PixelIO px;
    bool live = front_load<MODE>(f, i, px);
    uint32_t sr = 0, sg = 0, sb = 0;
    double mr = 0.0, mg = 0.0, mb = 0.0;
    const int r0 = (int)px.r, g0 = (int)px.g, b0 = (int)px.b;
    for (uint32_t base = 0; base < iterations; base += CHUNK) {
        uint32_t m = min((uint32_t)CHUNK, iterations - base);
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) s_off[j] = offsets[base + j];
        __syncthreads();
        if (live) {
            // (r, g) and (b, 0) as signed 16-bit pairs: one VIADDMNMX.S16x2.RELU per pair is
            // sat_u8(c + o) for two channels (add, min with 255, max with 0), one PRMT gathers the three
            // bytes into the table index, three IDP.4A pick the result's channels into their sums. The
            // kernel was bound by the ALU pipe (87 %) at 23 instructions per iteration; this is 9.
            const uint32_t c_rg = (uint32_t)r0 | ((uint32_t)g0 << 16), c_b = (uint32_t)b0;
            const uint2* s_off2 = reinterpret_cast<const uint2*>(s_off);   // short4 (x, y, z, w) = {x | y << 16, z | w << 16}
#pragma unroll 4
            for (uint32_t j = 0; j < m; ++j) {
                const uint2 o = s_off2[j];
                const uint32_t x_rg = __viaddmin_s16x2_relu(c_rg, o.x, 0x00ff00ffu), x_b = __viaddmin_s16x2_relu(c_b, o.y, 0x00ff00ffu);
                const uint32_t idx = __byte_perm(x_rg, x_b, 0x7420);   // R = x_rg byte 0, G = x_rg byte 2, B = x_b byte 0, 0
                uint32_t v = __ldg(nn_table + idx);
                if (EXACT_F64) {
                    mr = __dadd_rn(mr, s_div[v & 255u]);
                    mg = __dadd_rn(mg, s_div[(v >> 8) & 255u]);
                    mb = __dadd_rn(mb, s_div[(v >> 16) & 255u]);
                } else {
                    sr = __dp4a(v, 0x00000001u, sr);
                    sg = __dp4a(v, 0x00000100u, sg);
                    sb = __dp4a(v, 0x00010000u, sb);
                }
            }
        }
    }
    if (!live) return;
    uint32_t r, g, b;
    if (EXACT_F64) {
        // f64::round (half away from zero) then saturating `as u8`
        r = (uint32_t)fmin(fmax(round(mr), 0.0), 255.0);
        g = (uint32_t)fmin(fmax(round(mg), 0.0), 255.0);
        b = (uint32_t)fmin(fmax(round(mb), 0.0), 255.0);
    } else {
        uint32_t half = log2k ? (1u << (log2k - 1)) : 0u;
        r = (sr + half) >> log2k; g = (sg + half) >> log2k; b = (sb + half) >> log2k;
    }
    front_store<MODE>(f, px, r, g, b);
}

// ---------------------------------------------------------------------------
// Diagnostics: on-box FP32 FMA and MUFU peak (the generation kernels' roofline
// denominators must be MEASURED at the clocks this GPU sustains, SURVEY.md
// hard part 10). 8 independent dependency chains per thread.
__global__ void __launch_bounds__(256) k_peak_fma(float* __restrict__ sink, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_peak_mufu(float* __restrict__ sink, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
            x0 = lg2_approx(x0); x1 = lg2_approx(x1); x2 = lg2_approx(x2); x3 = lg2_approx(x3);
        }
    }
    float s = (x0 + x1) + (x2 + x3);
    if (s == 123.456f) sink[0] = s;
}

}  // namespace lutgen
