// color.cuh -- sRGB8 <-> Oklab device arithmetic (oklab 1.1.2 + fast-srgb8 1.0.0,
// third-party crates the reference calls at crates/lib/src/interpolation/rbf.rs:32,59,100
// and nearest_neighbor.rs:23,47,58).
//
// Two flavours:
//   *_exact : bit-for-bit the f32 arithmetic the Rust crates perform (no FMA
//             contraction -- every op goes through __fmul_rn/__fadd_rn -- and a
//             CORRECTLY ROUNDED cube root). Used wherever the result feeds an
//             equality test or an argmin that must match the CPU oracle byte for
//             byte (NearestNeighbor, `palette.contains`, the exact RBF kernel).
//   *_fast  : FMA-contracted, MUFU-seeded cube root (a few ulp). Used by the
//             throughput RBF kernel, whose contract is +-1 LSB.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lutgen {

// 256-entry sRGB EOTF table + 104-entry encoder table.
struct ColorTables {
    float decode[256];
    uint32_t encode[104];
    uint2 encode2[104];   // the same table rearranged for f32_to_srgb8_wide
};

static const uint32_t k_srgb_decode_bits[256] = {
#include "srgb_decode_table.inc"
};

static const uint32_t k_srgb_encode_table[104] = {
    0x0073000d, 0x007a000d, 0x0080000d, 0x0087000d, 0x008d000d, 0x0094000d, 0x009a000d, 0x00a1000d,
    0x00a7001a, 0x00b4001a, 0x00c1001a, 0x00ce001a, 0x00da001a, 0x00e7001a, 0x00f4001a, 0x0101001a,
    0x010e0033, 0x01280033, 0x01410033, 0x015b0033, 0x01750033, 0x018f0033, 0x01a80033, 0x01c20033,
    0x01dc0067, 0x020f0067, 0x02430067, 0x02760067, 0x02aa0067, 0x02dd0067, 0x03110067, 0x03440067,
    0x037800ce, 0x03df00ce, 0x044600ce, 0x04ad00ce, 0x051400ce, 0x057b00c5, 0x05dd00bc, 0x063b00b5,
    0x06970158, 0x07420142, 0x07e30130, 0x087b0120, 0x090b0112, 0x09940106, 0x0a1700fc, 0x0a9500f2,
    0x0b0f01cb, 0x0bf401ae, 0x0ccb0195, 0x0d950180, 0x0e56016e, 0x0f0d015e, 0x0fbc0150, 0x10630143,
    0x11070264, 0x1238023e, 0x1357021d, 0x14660201, 0x156601e9, 0x165a01d3, 0x174401c0, 0x182401af,
    0x18fe0331, 0x1a9602fe, 0x1c1502d2, 0x1d7e02ad, 0x1ed4028d, 0x201a0270, 0x21520256, 0x227d0240,
    0x239f0443, 0x25c003fe, 0x27bf03c4, 0x29a10392, 0x2b6a0367, 0x2d1d0341, 0x2ebe031f, 0x304d0300,
    0x31d105b0, 0x34a80555, 0x37520507, 0x39d504c5, 0x3c37048b, 0x3e7c0458, 0x40a8042a, 0x42bd0401,
    0x44c20798, 0x488e071e, 0x4c1c06b6, 0x4f76065d, 0x52a50610, 0x55ac05cc, 0x5892058f, 0x5b590559,
    0x5e0c0a23, 0x631c0980, 0x67db08f6, 0x6c55087f, 0x70940818, 0x74a007bd, 0x787d076c, 0x7c330723};

// Device copy of the tables (global memory; kernels stage it into shared memory
// because per-lane indices diverge and __constant__ serialises divergent reads).
struct Lab {
    float l, a, b;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Correctly rounded cbrtf for x >= 0 (x is an LMS value: 0 or in [1e-5, 1]).
// A MUFU seed for x^(-1/3) gets one Newton step in double (relative error
// ~2^-38), which leaves at most one float rounding decision open: the float
// midpoint m nearest the estimate. m has 25 significant bits, so m*m is exact in
// double and fma(m*m, m, -x) carries the exact sign of m^3 - x.
__device__ __forceinline__ float cbrtf_cr(float x) {
    if (!(x > 0.0f)) return x;
    float t0 = ex2_approx(-0.33333334f * lg2_approx(x));
    double xd = (double)x, t = (double)t0;
    double r = fma(-(xd * t), t * t, 1.0);
    t = fma(t * 0.33333333333333331, r, t);
    double yd = (xd * t) * t;
    float yf = __double2float_rn(yd);
    double yfd = (double)yf;
    if (yfd != yd) {
        bool up = yd > yfd;
        float nb = __uint_as_float(__float_as_uint(yf) + (up ? 1u : 0xffffffffu));
        double m = 0.5 * (yfd + (double)nb);
        double s = fma(m * m, m, -xd);
        if (up ? (s < 0.0) : (s > 0.0)) yf = nb;
    }
    return yf;
}

// A few-ulp cube root for the +-1 LSB kernels: MUFU seed + one f32 Newton step.
__device__ __forceinline__ float cbrtf_fast(float x) {
    if (!(x > 0.0f)) return 0.0f;
    float t = ex2_approx(-0.33333334f * lg2_approx(x));   // x^(-1/3)
    float r = fmaf(-(x * t), t * t, 1.0f);
    t = fmaf(t * 0.33333334f, r, t);
    return (x * t) * t;
}

// oklab::linear_srgb_to_oklab, evaluated exactly as Rust does (left to right,
// every product and sum rounded to f32).
__device__ __forceinline__ Lab linear_to_oklab_exact(float r, float g, float b) {
    float l = __fadd_rn(__fadd_rn(__fmul_rn(0.4122214708f, r), __fmul_rn(0.5363325363f, g)), __fmul_rn(0.0514459929f, b));
    float m = __fadd_rn(__fadd_rn(__fmul_rn(0.2119034982f, r), __fmul_rn(0.6806995451f, g)), __fmul_rn(0.1073969566f, b));
    float s = __fadd_rn(__fadd_rn(__fmul_rn(0.0883024619f, r), __fmul_rn(0.2817188376f, g)), __fmul_rn(0.6299787005f, b));
    float l_ = cbrtf_cr(l), m_ = cbrtf_cr(m), s_ = cbrtf_cr(s);
    Lab o;
    o.l = __fsub_rn(__fadd_rn(__fmul_rn(0.2104542553f, l_), __fmul_rn(0.7936177850f, m_)), __fmul_rn(0.0040720468f, s_));
    o.a = __fadd_rn(__fsub_rn(__fmul_rn(1.9779984951f, l_), __fmul_rn(2.4285922050f, m_)), __fmul_rn(0.4505937099f, s_));
    o.b = __fsub_rn(__fadd_rn(__fmul_rn(0.0259040371f, l_), __fmul_rn(0.7827717662f, m_)), __fmul_rn(0.8086757660f, s_));
    return o;
}

__device__ __forceinline__ Lab linear_to_oklab_fast(float r, float g, float b) {
    float l = 0.4122214708f * r + 0.5363325363f * g + 0.0514459929f * b;
    float m = 0.2119034982f * r + 0.6806995451f * g + 0.1073969566f * b;
    float s = 0.0883024619f * r + 0.2817188376f * g + 0.6299787005f * b;
    float l_ = cbrtf_fast(l), m_ = cbrtf_fast(m), s_ = cbrtf_fast(s);
    Lab o;
    o.l = 0.2104542553f * l_ + 0.7936177850f * m_ - 0.0040720468f * s_;
    o.a = 1.9779984951f * l_ - 2.4285922050f * m_ + 0.4505937099f * s_;
    o.b = 0.0259040371f * l_ + 0.7827717662f * m_ - 0.8086757660f * s_;
    return o;
}

// Packed f32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100a): two values per instruction; a scalar
// operand is broadcast for free. Used by the +-1 LSB kernels and by decisions that are re-checked exactly.
typedef float2 f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 splat(float x) { return make_float2(x, x); }

// x^(1/3) for two LMS values > 0 (callers add LMS_FLOOR, so black is 1e-30: its cube root, 1e-10, is zero at
// Oklab's resolution). MUFU seed t = x^(-1/3) (1 + e), y0 = x t^2 = x^(1/3) (1 + e)^2, and with
// r = 1 - x t^3 = -3e + O(e^2) the first-order correction y = y0 + (2/3) y0 r leaves an error of O(e^2),
// like a Newton step on t would, in six packed instructions instead of eight.
constexpr float LMS_FLOOR = 1e-30f;
__device__ __forceinline__ f2 cbrt2(f2 x) {
    f2 lg = make_float2(lg2_approx(x.x), lg2_approx(x.y));
    lg = mul2(lg, splat(-0.33333334f));
    const f2 t = make_float2(ex2_approx(lg.x), ex2_approx(lg.y));
    const f2 y0 = mul2(x, mul2(t, t));
    const f2 r = fma2(mul2(y0, t), splat(-1.0f), splat(1.0f));   // 1 - x t^3
    return fma2(mul2(y0, splat(0.6666667f)), r, y0);
}

// Two entries' Oklab from their LMS values.
__device__ __forceinline__ void lms_to_lab2(f2 l, f2 m, f2 s, float lum, f2& L, f2& A, f2& B) {
    f2 l_ = cbrt2(l), m_ = cbrt2(m), s_ = cbrt2(s);
    L = mul2(fma2(splat(0.2104542553f), l_, fma2(splat(0.7936177850f), m_, mul2(splat(-0.0040720468f), s_))), splat(lum));
    A = fma2(splat(1.9779984951f), l_, fma2(splat(-2.4285922050f), m_, mul2(splat(0.4505937099f), s_)));
    B = fma2(splat(0.0259040371f), l_, fma2(splat(0.7827717662f), m_, mul2(splat(-0.8086757660f), s_)));
}

// oklab::oklab_to_linear_srgb, exact evaluation order.
__device__ __forceinline__ void oklab_to_linear_exact(Lab c, float& r, float& g, float& b) {
    float l_ = __fadd_rn(__fadd_rn(c.l, __fmul_rn(0.3963377774f, c.a)), __fmul_rn(0.2158037573f, c.b));
    float m_ = __fsub_rn(__fsub_rn(c.l, __fmul_rn(0.1055613458f, c.a)), __fmul_rn(0.0638541728f, c.b));
    float s_ = __fsub_rn(__fsub_rn(c.l, __fmul_rn(0.0894841775f, c.a)), __fmul_rn(1.2914855480f, c.b));
    float l = __fmul_rn(__fmul_rn(l_, l_), l_);
    float m = __fmul_rn(__fmul_rn(m_, m_), m_);
    float s = __fmul_rn(__fmul_rn(s_, s_), s_);
    r = __fadd_rn(__fsub_rn(__fmul_rn(4.0767416621f, l), __fmul_rn(3.3077115913f, m)), __fmul_rn(0.2309699292f, s));
    g = __fsub_rn(__fadd_rn(__fmul_rn(-1.2684380046f, l), __fmul_rn(2.6097574011f, m)), __fmul_rn(0.3413193965f, s));
    b = __fadd_rn(__fsub_rn(__fmul_rn(-0.0041960863f, l), __fmul_rn(0.7034186147f, m)), __fmul_rn(1.7076147010f, s));
}

__device__ __forceinline__ void oklab_to_linear_fast(Lab c, float& r, float& g, float& b) {
    float l_ = c.l + 0.3963377774f * c.a + 0.2158037573f * c.b;
    float m_ = c.l - 0.1055613458f * c.a - 0.0638541728f * c.b;
    float s_ = c.l - 0.0894841775f * c.a - 1.2914855480f * c.b;
    float l = l_ * l_ * l_, m = m_ * m_ * m_, s = s_ * s_ * s_;
    r = 4.0767416621f * l - 3.3077115913f * m + 0.2309699292f * s;
    g = -1.2684380046f * l + 2.6097574011f * m - 0.3413193965f * s;
    b = -0.0041960863f * l - 0.7034186147f * m + 1.7076147010f * s;
}

// fast_srgb8::f32_to_srgb8: clamp to [2^-13, 1 - 2^-24] (NaN -> the minimum), look up
//   e = table[(fu - 0x39000000) >> 20],  bias = (e >> 16) << 9,  scale = e & 0xffff,  t = (fu >> 12) & 0xff,
//   result = (bias + scale * t) >> 16.
// Evaluated in seven instructions per channel instead of thirteen: with i = fu >> 20 (the table index plus
// 0x390), fu >> 12 = (i << 8) | t, so
//   (bias << 9) + scale * t  =  [(bias << 9) - scale * (i << 8)] + scale * (fu >> 12)      (mod 2^32)
// and the bracket is a per-bucket constant: encode2[i - 0x390] = (bracket, scale). No mask for t, no shifts for
// bias and scale. Returns the sum: the sRGB byte is its byte 2, bits 24.. are zero (srgb8_pack3 picks the bytes).
// Checked against the literal formula for all 109,051,904 floats of the clamped range (tests/test_srgb_encode.py).
__host__ __device__ inline uint2 srgb_encode2_entry(uint32_t e, uint32_t index) {
    const uint32_t bias = (e >> 16) << 9, scale = e & 0xffffu;
    uint2 r;
    r.x = bias - scale * ((index + 0x390u) << 8);
    r.y = scale;
    return r;
}
__device__ __forceinline__ uint32_t f32_to_srgb8_wide(float f, const uint2* __restrict__ enc2) {
    const float in = fminf(fmaxf(f, __uint_as_float(0x39000000u)), __uint_as_float(0x3f7fffffu));   // NaN -> the minimum, like the original
    const uint32_t fu = __float_as_uint(in);
    const uint2 e = (enc2 - 0x390)[fu >> 20];
    return e.x + e.y * (fu >> 12);
}
__device__ __forceinline__ uint32_t srgb8_pack3(uint32_t wr, uint32_t wg, uint32_t wb) {   // r | g << 8 | b << 16 from three wide results
    return __byte_perm(__byte_perm(wr, wg, 0x7762), wb, 0x7610);
}
__device__ __forceinline__ uint32_t f32_to_srgb8(float f, const uint2* __restrict__ enc2) { return f32_to_srgb8_wide(f, enc2) >> 16; }

// Shared staging of the colour tables for kernels whose lanes index them divergently.
struct SmemTables {
    float decode[256];
    uint2 encode2[104];
};
__device__ __forceinline__ void stage_tables(SmemTables& s, const ColorTables* __restrict__ t) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s.decode[i] = t->decode[i];
    for (int i = threadIdx.x; i < 104; i += blockDim.x) s.encode2[i] = t->encode2[i];
    __syncthreads();
}

// identity::generate's channel value for grid coordinate v of a level-L CLUT:
// v * 255 / (cs - 1), integer floor (crates/lib/src/identity.rs:16-20).
__device__ __forceinline__ uint32_t hald_channel(uint32_t v, uint32_t cs_minus_1) { return v * 255u / cs_minus_1; }

}  // namespace lutgen
