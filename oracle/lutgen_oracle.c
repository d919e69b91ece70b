/*
 * lutgen_oracle.c -- CPU restatement of the lutgen-rs Hald-CLUT hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may
 * load it. The CUDA library never links, loads or calls anything in this file.
 *
 * The reference (ozwaldorf/lutgen-rs, crate `lutgen` 0.15.0) is pure Rust and no
 * Rust toolchain exists in this environment, so the reference itself cannot be
 * compiled or run here. This file restates, function by function, what the
 * reference's source says; every function cites the reference file:line it
 * follows (paths relative to the reference root).
 *
 * Third-party arithmetic the reference pulls from crates that are NOT vendored
 * under the reference tree (Cargo.lock pins: oklab 1.1.2, fast-srgb8 1.0.0,
 * kiddo 5.2.4, rand 0.10.0 / rand_distr 0.6.0) is restated from the published
 * algorithms:
 *   - oklab: Bjorn Ottosson's Oklab matrices, f32, no FMA contraction (Rust never
 *     contracts), cube root = f32::cbrt.  Rust's f32::cbrt is the platform libm
 *     cbrtf (<= 1 ulp, flavour varies by libc). This oracle uses the CORRECTLY
 *     ROUNDED cube root (what glibc >= 2.41 / CORE-MATH returns); it is <= 1 ulp
 *     from every faithful libm and is reproducible bit-for-bit on the GPU.
 *   - fast-srgb8: 256-entry decode table (exact sRGB EOTF rounded to f32) and
 *     the 104-entry table encoder (Giesen's float->sRGB8), SURVEY.md appendix A.
 *   - kiddo nearest_n / nearest_one: brute force, ascending squared distance,
 *     ties broken by palette index (kiddo's tie order is traversal dependent and
 *     pinned by nothing in the reference).
 *   - rand StdRng + rand_distr Normal: NOT reproduced (north star: the stream may
 *     differ). The noise table (iterations x 4 f64 normals, shared by every
 *     pixel because the reference re-seeds per pixel, gaussian_sample.rs:52) is
 *     an explicit input so the GPU and this oracle can share one table.
 *
 * PARITY PIN: checked by tests/test_oracle.py against the reference's own
 * committed outputs docs/assets/{gruvbox-dark,catppuccin-mocha,nord}-hald-clut.png
 * (fixtures in tests/golden/, max 1 LSB, <= 152 of 3,000,000 bytes differ) and the
 * known-answer test crates/cli/src/main.rs:1192-1226. Shepard/Linear functors,
 * lum_factor != 1, preserve = true and the sampling remapper are pinned by
 * nothing in the reference ("parity unpinned" for those rows).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int lutgen_oracle_max_threads(void);

/* ------------------------------------------------------------------------- */
/* fast-srgb8 (third party, un-vendored): decode table + table encoder        */
/* ------------------------------------------------------------------------- */

static float g_srgb_decode[256];
static int g_tables_ready = 0;

/* fast_srgb8::srgb8_to_f32: exact sRGB EOTF of i/255, rounded to f32. */
static void init_tables(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < 256; ++i) {
        double c = (double)i / 255.0;
        double v = (c <= 0.04045) ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
        g_srgb_decode[i] = (float)v;
    }
    g_tables_ready = 1;
}

static const uint32_t k_to_srgb8_table[104] = {
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

static inline uint32_t f32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float bits_f32(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* fast_srgb8::f32_to_srgb8: clamp (NaN -> minimum -> 0), 104-entry table. */
static inline uint8_t f32_to_srgb8(float f) {
    const uint32_t MINV = 0x39000000u, MAXV = 0x3f7fffffu;
    float in = f;
    if (!(in > bits_f32(MINV))) in = bits_f32(MINV);
    if (in > bits_f32(MAXV)) in = bits_f32(MAXV);
    uint32_t fu = f32_bits(in);
    uint32_t e = k_to_srgb8_table[(fu - MINV) >> 20];
    uint32_t bias = (e >> 16) << 9;
    uint32_t scale = e & 0xffffu;
    uint32_t t = (fu >> 12) & 0xffu;
    return (uint8_t)((bias + scale * t) >> 16);
}

/* ------------------------------------------------------------------------- */
/* oklab 1.1.2 (third party, un-vendored)                                    */
/* ------------------------------------------------------------------------- */

/* Correctly rounded f32 cube root (see header). The double estimate is within
 * one double ulp; the only float rounding decision that can go wrong is against
 * the float midpoint nearest the estimate, and fma(m*m, m, -x) has the exact sign
 * of m^3 - x because m (25 bits) squares exactly in double. */
static float cr_cbrtf(float x) {
    if (x == 0.0f || x != x || isinf(x)) return x;
    double xd = fabs((double)x);
    double yd = cbrt(xd);
    float yf = (float)yd;
    if ((double)yf != yd) {
        int up = yd > (double)yf;
        float nb = nextafterf(yf, up ? INFINITY : 0.0f);
        double m = 0.5 * ((double)yf + (double)nb);
        double s = fma(m * m, m, -xd); /* sign(m^3 - x), exact */
        if (up) {
            if (s < 0.0) yf = nb; /* m^3 < x: root above the midpoint */
        } else {
            if (s > 0.0) yf = nb; /* m^3 > x: root below the midpoint */
        }
    }
    return x < 0.0f ? -yf : yf;
}

typedef struct {
    float l, a, b;
} oklab_t;

/* oklab::srgb_to_oklab -> linear_srgb_to_oklab (Ottosson's M1, cbrt, M2), f32,
 * evaluated left to right without contraction. Reference call sites:
 * crates/lib/src/interpolation/rbf.rs:32,59; nearest_neighbor.rs:23,47. */
static oklab_t srgb_to_oklab(uint8_t r8, uint8_t g8, uint8_t b8) {
    float r = g_srgb_decode[r8], g = g_srgb_decode[g8], b = g_srgb_decode[b8];
    float l = 0.4122214708f * r + 0.5363325363f * g + 0.0514459929f * b;
    float m = 0.2119034982f * r + 0.6806995451f * g + 0.1073969566f * b;
    float s = 0.0883024619f * r + 0.2817188376f * g + 0.6299787005f * b;
    float l_ = cr_cbrtf(l), m_ = cr_cbrtf(m), s_ = cr_cbrtf(s);
    oklab_t o;
    o.l = 0.2104542553f * l_ + 0.7936177850f * m_ - 0.0040720468f * s_;
    o.a = 1.9779984951f * l_ - 2.4285922050f * m_ + 0.4505937099f * s_;
    o.b = 0.0259040371f * l_ + 0.7827717662f * m_ - 0.8086757660f * s_;
    return o;
}

/* oklab::oklab_to_srgb -> oklab_to_linear_srgb + fast_srgb8::f32_to_srgb8.
 * Reference call sites: rbf.rs:100; nearest_neighbor.rs:58. */
static void oklab_to_srgb(oklab_t c, uint8_t out[3]) {
    float l_ = c.l + 0.3963377774f * c.a + 0.2158037573f * c.b;
    float m_ = c.l - 0.1055613458f * c.a - 0.0638541728f * c.b;
    float s_ = c.l - 0.0894841775f * c.a - 1.2914855480f * c.b;
    float l = l_ * l_ * l_;
    float m = m_ * m_ * m_;
    float s = s_ * s_ * s_;
    float r = 4.0767416621f * l - 3.3077115913f * m + 0.2309699292f * s;
    float g = -1.2684380046f * l + 2.6097574011f * m - 0.3413193965f * s;
    float b = -0.0041960863f * l - 0.7034186147f * m + 1.7076147010f * s;
    out[0] = f32_to_srgb8(r);
    out[1] = f32_to_srgb8(g);
    out[2] = f32_to_srgb8(b);
}

ORACLE_API void lutgen_oracle_srgb_to_oklab(const uint8_t* rgb, size_t n, float* out_lab) {
    init_tables();
    for (size_t i = 0; i < n; ++i) {
        oklab_t o = srgb_to_oklab(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        out_lab[3 * i] = o.l;
        out_lab[3 * i + 1] = o.a;
        out_lab[3 * i + 2] = o.b;
    }
}

ORACLE_API void lutgen_oracle_oklab_to_srgb(const float* lab, size_t n, uint8_t* out_rgb) {
    init_tables();
    for (size_t i = 0; i < n; ++i) {
        oklab_t c = {lab[3 * i], lab[3 * i + 1], lab[3 * i + 2]};
        oklab_to_srgb(c, out_rgb + 3 * i);
    }
}

ORACLE_API void lutgen_oracle_srgb_decode_table(float* out256) {
    init_tables();
    memcpy(out256, g_srgb_decode, sizeof g_srgb_decode);
}

ORACLE_API float lutgen_oracle_cbrtf(float x) { return cr_cbrtf(x); }

/* ------------------------------------------------------------------------- */
/* identity.rs                                                                */
/* ------------------------------------------------------------------------- */

/* identity::generate, crates/lib/src/identity.rs:7-34. out_rgb: 3*level^6 bytes. */
ORACLE_API void lutgen_oracle_identity(uint8_t level8, uint8_t* out_rgb) {
    uint32_t level = level8, cs = level * level;
    size_t i = 0;
    for (uint32_t blue = 0; blue < cs; ++blue) {
        uint8_t b = (uint8_t)(blue * 255 / (cs - 1));
        for (uint32_t green = 0; green < cs; ++green) {
            uint8_t g = (uint8_t)(green * 255 / (cs - 1));
            for (uint32_t red = 0; red < cs; ++red) {
                uint8_t r = (uint8_t)(red * 255 / (cs - 1));
                out_rgb[i++] = r;
                out_rgb[i++] = g;
                out_rgb[i++] = b;
            }
        }
    }
}

/* identity::correct_pixel, identity.rs:40-52 (get_pixel(x, y) on an n x n RGB8
 * image, n = level^3, row-major). */
static inline void correct_pixel(const uint8_t in[3], const uint8_t* lut, uint32_t level, uint8_t out[3]) {
    uint32_t cs = level * level;
    uint32_t r = (uint32_t)in[0] * (cs - 1) / 255;
    uint32_t g = (uint32_t)in[1] * (cs - 1) / 255;
    uint32_t b = (uint32_t)in[2] * (cs - 1) / 255;
    uint32_t x = (r % cs) + (g % level) * cs;
    uint32_t y = (b * level) + (g / level);
    size_t n = (size_t)cs * level;
    const uint8_t* p = lut + 3 * ((size_t)y * n + x);
    out[0] = p[0];
    out[1] = p[1];
    out[2] = p[2];
}

ORACLE_API void lutgen_oracle_correct_pixel(const uint8_t* in_rgb, const uint8_t* lut, uint8_t level, uint8_t* out_rgb) {
    correct_pixel(in_rgb, lut, level, out_rgb);
}

/* identity::correct_image_with_level, identity.rs:71-78: serial loop, alpha kept.
 * threads <= 1 mirrors the reference (single thread); threads > 1 is the
 * "all cores for fairness" variant used only as a reported baseline. */
ORACLE_API void lutgen_oracle_correct_image(uint8_t* rgba, size_t npix, const uint8_t* lut, uint8_t level, int threads) {
    if (threads <= 1) {
        for (size_t i = 0; i < npix; ++i) {
            uint8_t o[3];
            correct_pixel(rgba + 4 * i, lut, level, o);
            rgba[4 * i] = o[0];
            rgba[4 * i + 1] = o[1];
            rgba[4 * i + 2] = o[2];
        }
        return;
    }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long i = 0; i < (long long)npix; ++i) {
        uint8_t o[3];
        correct_pixel(rgba + 4 * i, lut, level, o);
        rgba[4 * i] = o[0];
        rgba[4 * i + 1] = o[1];
        rgba[4 * i + 2] = o[2];
    }
}

/* identity::detect_level, identity.rs:85-99. Returns -1 where the reference
 * panics (assert_eq failures). */
ORACLE_API int lutgen_oracle_detect_level(uint32_t width, uint32_t height) {
    uint32_t level = 2;
    while (level * level * level < width) ++level;
    if (width != level * level * level) return -1;
    if (width != height) return -1;
    return (int)(level & 0xff); /* `level as u8` */
}

/* ------------------------------------------------------------------------- */
/* remappers                                                                  */
/* ------------------------------------------------------------------------- */

enum { K_GAUSSIAN = 0, K_SHEPARD = 1, K_LINEAR = 2, K_NEAREST = 3, K_SAMPLING = 4, K_BLUR = 5 };

typedef struct {
    int32_t kind;
    int32_t preserve;
    const uint8_t* palette; /* palette_len x [u8;3] */
    uint64_t palette_len;
    double param; /* Gaussian shape / Shepard power */
    uint64_t nearest;
    double lum_factor;
    double mean, std_dev;
    uint64_t iterations;
    uint64_t seed;
} lutgen_oracle_desc;

typedef struct {
    lutgen_oracle_desc d;
    uint8_t* pal_rgb;    /* n x 3 */
    double* pal_lab;     /* n x 3: [l*lum, a, b] widened from f32 */
    float* pal_ab;       /* n x 2 f32 (a, b) for NearestNeighbor preserve */
    int use_tree;        /* rbf.rs:38: nearest != 0 && nearest < palette.len() */
    double* noise;       /* iterations x 4 */
} oracle_remapper;

/* splitmix64 + Marsaglia polar: the oracle's OWN normal stream (the reference's
 * ChaCha12 + ziggurat stream is not reproduced, see header). */
static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double unit_open(uint64_t* s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static void fill_normals(double* out, size_t n, double mean, double std_dev, uint64_t seed) {
    uint64_t s = seed;
    size_t i = 0;
    while (i < n) {
        double u = 2.0 * unit_open(&s) - 1.0, v = 2.0 * unit_open(&s) - 1.0;
        double q = u * u + v * v;
        if (q >= 1.0 || q == 0.0) continue;
        double f = sqrt(-2.0 * log(q) / q);
        out[i++] = mean + std_dev * (u * f);
        if (i < n) out[i++] = mean + std_dev * (v * f);
    }
}

/* RBFRemapper::with_function (rbf.rs:22-53), NearestNeighborRemapper::new
 * (nearest_neighbor.rs:19-37), GaussianSamplingRemapper::new (gaussian_sample.rs:27-46). */
ORACLE_API void* lutgen_oracle_remapper_create(const lutgen_oracle_desc* desc) {
    init_tables();
    if (!desc || desc->kind < 0 || desc->kind > K_BLUR) return NULL;
    if (desc->kind == K_SAMPLING && !isfinite(desc->std_dev)) return NULL; /* Normal::new(..).unwrap() */
    oracle_remapper* r = (oracle_remapper*)calloc(1, sizeof *r);
    r->d = *desc;
    size_t n = desc->palette_len;
    r->pal_rgb = (uint8_t*)malloc(3 * n + 1);
    r->pal_lab = (double*)malloc(sizeof(double) * 3 * n + 8);
    r->pal_ab = (float*)malloc(sizeof(float) * 2 * n + 8);
    memcpy(r->pal_rgb, desc->palette, 3 * n);
    r->d.palette = r->pal_rgb;
    for (size_t i = 0; i < n; ++i) {
        oklab_t o = srgb_to_oklab(r->pal_rgb[3 * i], r->pal_rgb[3 * i + 1], r->pal_rgb[3 * i + 2]);
        r->pal_lab[3 * i] = (double)o.l * desc->lum_factor;
        r->pal_lab[3 * i + 1] = (double)o.a;
        r->pal_lab[3 * i + 2] = (double)o.b;
        r->pal_ab[2 * i] = o.a;
        r->pal_ab[2 * i + 1] = o.b;
    }
    r->use_tree = (desc->nearest != 0 && desc->nearest < n);
    if (desc->kind == K_SAMPLING) {
        size_t k = desc->iterations;
        r->noise = (double*)malloc(sizeof(double) * 4 * k + 8);
        fill_normals(r->noise, 4 * k, desc->mean, desc->std_dev, desc->seed);
    }
    return r;
}

ORACLE_API void lutgen_oracle_remapper_destroy(void* h) {
    oracle_remapper* r = (oracle_remapper*)h;
    if (!r) return;
    free(r->pal_rgb);
    free(r->pal_lab);
    free(r->pal_ab);
    free(r->noise);
    free(r);
}

/* Test hook: replace the iterations x 4 noise table (so the GPU and the oracle
 * can consume the same normals and be compared byte for byte). */
ORACLE_API int lutgen_oracle_sampling_set_noise(void* h, const double* noise) {
    oracle_remapper* r = (oracle_remapper*)h;
    if (!r || r->d.kind != K_SAMPLING) return -1;
    memcpy(r->noise, noise, sizeof(double) * 4 * r->d.iterations);
    return 0;
}
ORACLE_API int lutgen_oracle_sampling_get_noise(void* h, double* noise) {
    oracle_remapper* r = (oracle_remapper*)h;
    if (!r || r->d.kind != K_SAMPLING) return -1;
    memcpy(noise, r->noise, sizeof(double) * 4 * r->d.iterations);
    return 0;
}

/* kiddo SquaredEuclidean::dist: fold(0, |acc, (a-b)*(a-b)|), f64, no FMA. */
static inline double sq_dist(const double a[3], const double b[3]) {
    double acc = 0.0;
    for (int i = 0; i < 3; ++i) {
        double d = a[i] - b[i];
        acc = acc + d * d;
    }
    return acc;
}

/* RadialBasisFn impls, rbf.rs:153-180. */
static inline double radial_basis(const oracle_remapper* r, double d) {
    switch (r->d.kind) {
        case K_LINEAR: return d;                                   /* rbf.rs:158 */
        case K_SHEPARD: return 1.0 / pow(sqrt(d), r->d.param);     /* rbf.rs:167 */
        default: return exp(-r->d.param * d);                      /* rbf.rs:178 */
    }
}

/* NearestNeighborRemapper::remap_pixel, nearest_neighbor.rs:46-61. */
static void nn_remap_pixel(const oracle_remapper* r, uint8_t* px) {
    oklab_t c = srgb_to_oklab(px[0], px[1], px[2]);
    double q[3] = {(double)c.l * r->d.lum_factor, (double)c.a, (double)c.b};
    size_t n = r->d.palette_len, best = 0;
    double bd = INFINITY;
    for (size_t i = 0; i < n; ++i) {
        double d = sq_dist(q, r->pal_lab + 3 * i);
        if (d < bd) {
            bd = d;
            best = i;
        }
    }
    if (!r->d.preserve) {
        px[0] = r->pal_rgb[3 * best];
        px[1] = r->pal_rgb[3 * best + 1];
        px[2] = r->pal_rgb[3 * best + 2];
    } else {
        oklab_t o = {c.l, r->pal_ab[2 * best], r->pal_ab[2 * best + 1]};
        oklab_to_srgb(o, px);
    }
}

/* RBFRemapper::remap_pixel, rbf.rs:57-111. `scratch` holds palette_len doubles
 * + palette_len size_t for the brute-force nearest_n. */
static void rbf_remap_pixel(const oracle_remapper* r, uint8_t* px, double* dist, uint32_t* order) {
    oklab_t c = srgb_to_oklab(px[0], px[1], px[2]);
    double color[3] = {(double)c.l * r->d.lum_factor, (double)c.a, (double)c.b};
    size_t n = r->d.palette_len;
    /* rbf.rs:66-68: palette.contains(&color) -> leave the pixel untouched */
    for (size_t i = 0; i < n; ++i) {
        const double* p = r->pal_lab + 3 * i;
        if (p[0] == color[0] && p[1] == color[1] && p[2] == color[2]) return;
    }
    double num[3] = {0.0, 0.0, 0.0}, den = 0.0;
    if (!r->use_tree) {
        /* rbf.rs:74-84: every palette colour, palette order */
        for (size_t i = 0; i < n; ++i) {
            const double* p = r->pal_lab + 3 * i;
            double w = radial_basis(r, sq_dist(color, p));
            num[0] += p[0] * w;
            num[1] += p[1] * w;
            num[2] += p[2] * w;
            den += w;
        }
    } else {
        /* rbf.rs:85-97: tree.nearest_n(&color, nearest): ascending distance */
        size_t k = (size_t)r->d.nearest;
        for (size_t i = 0; i < n; ++i) {
            dist[i] = sq_dist(color, r->pal_lab + 3 * i);
            order[i] = (uint32_t)i;
        }
        for (size_t a = 0; a < k; ++a) { /* partial selection sort, stable on index */
            size_t m = a;
            for (size_t b = a + 1; b < n; ++b)
                if (dist[order[b]] < dist[order[m]] ||
                    (dist[order[b]] == dist[order[m]] && order[b] < order[m]))
                    m = b;
            uint32_t t = order[a];
            order[a] = order[m];
            order[m] = t;
            const double* p = r->pal_lab + 3 * order[a];
            double w = radial_basis(r, dist[order[a]]);
            num[0] += p[0] * w;
            num[1] += p[1] * w;
            num[2] += p[2] * w;
            den += w;
        }
    }
    oklab_t o;
    o.l = r->d.preserve ? (float)(color[0] / r->d.lum_factor) : (float)(num[0] / den / r->d.lum_factor);
    o.a = (float)(num[1] / den);
    o.b = (float)(num[2] / den);
    oklab_to_srgb(o, px);
}

/* Rust `f64 as u8`: saturating, NaN -> 0. f64::round: half away from zero. */
static inline uint8_t f64_as_u8(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

/* GaussianSamplingRemapper::remap_pixel, gaussian_sample.rs:49-76. px is RGBA. */
static void sampling_remap_pixel(const oracle_remapper* r, uint8_t* px) {
    double mean[3] = {0.0, 0.0, 0.0};
    size_t iters = r->d.iterations;
    double total = (double)iters;
    for (size_t k = 0; k < iters; ++k) {
        uint8_t p[4] = {px[0], px[1], px[2], px[3]};
        for (int c = 0; c < 4; ++c) p[c] = f64_as_u8(round((double)p[c] + r->noise[4 * k + c]));
        nn_remap_pixel(r, p);
        mean[0] += (double)p[0] / total;
        mean[1] += (double)p[1] / total;
        mean[2] += (double)p[2] / total;
    }
    px[0] = f64_as_u8(round(mean[0]));
    px[1] = f64_as_u8(round(mean[1]));
    px[2] = f64_as_u8(round(mean[2]));
}

static void remap_pixel(const oracle_remapper* r, uint8_t* px, double* dist, uint32_t* order) {
    switch (r->d.kind) {
        case K_NEAREST: nn_remap_pixel(r, px); break;
        case K_SAMPLING: sampling_remap_pixel(r, px); break;
        default: rbf_remap_pixel(r, px, dist, order); break;
    }
}

/* InterpolatedRemapper::{remap_image, par_remap_image}, interpolation/mod.rs:30-50.
 * rgba in place, alpha untouched. threads <= 1: sequential; else OpenMP dynamic
 * (rayon par_pixels_mut stand-in). */
ORACLE_API void lutgen_oracle_remap_image(const void* h, uint8_t* rgba, size_t npix, int threads) {
    const oracle_remapper* r = (const oracle_remapper*)h;
    size_t n = r->d.palette_len;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        double* dist = (double*)malloc(sizeof(double) * (n + 1));
        uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
#pragma omp for schedule(dynamic, 4096)
        for (long long i = 0; i < (long long)npix; ++i) remap_pixel(r, rgba + 4 * i, dist, order);
        free(dist);
        free(order);
    }
}

/* GenerateLut blanket impl, lib.rs:135-147: identity -> Rgb->Rgba (alpha 255)
 * -> remap_image -> Rgba->Rgb. Rows [row_begin, row_end) of the n x n LUT only
 * (the whole LUT when 0..n), written at their place in out_rgb (3*level^6). */
ORACLE_API void lutgen_oracle_generate_lut_rows(const void* h, uint8_t level8, uint8_t* out_rgb, uint32_t row_begin,
                                                uint32_t row_end, int threads) {
    const oracle_remapper* r = (const oracle_remapper*)h;
    uint32_t level = level8, cs = level * level;
    size_t side = (size_t)cs * level, n = r->d.palette_len;
    if (threads < 1) threads = 1;
    long long e0 = (long long)row_begin * (long long)side, e1 = (long long)row_end * (long long)side;
#pragma omp parallel num_threads(threads)
    {
        double* dist = (double*)malloc(sizeof(double) * (n + 1));
        uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
#pragma omp for schedule(dynamic, 4096)
        for (long long i = e0; i < e1; ++i) {
            uint32_t red = (uint32_t)(i % cs), green = (uint32_t)((i / cs) % cs), blue = (uint32_t)(i / ((long long)cs * cs));
            uint8_t px[4] = {(uint8_t)(red * 255 / (cs - 1)), (uint8_t)(green * 255 / (cs - 1)),
                             (uint8_t)(blue * 255 / (cs - 1)), 255};
            remap_pixel(r, px, dist, order);
            out_rgb[3 * i] = px[0];
            out_rgb[3 * i + 1] = px[1];
            out_rgb[3 * i + 2] = px[2];
        }
        free(dist);
        free(order);
    }
}

ORACLE_API void lutgen_oracle_generate_lut(const void* h, uint8_t level, uint8_t* out_rgb, int threads) {
    uint32_t side = (uint32_t)level * level * level;
    lutgen_oracle_generate_lut_rows(h, level, out_rgb, 0, side, threads);
}

/* ------------------------------------------------------------------------- */
/* GaussianBlurRemapper, crates/lib/src/interpolation/gaussian_blur.rs        */
/* (the remapper the CLI and Studio use by default)        */
/* ------------------------------------------------------------------------- */

/* sq_dist, gaussian_blur.rs:504-510: f32, left to right, no FMA. */
static inline float sq_dist_f32(const float a[3], const float b[3]) {
    float dl = a[0] - b[0], da = a[1] - b[1], db = a[2] - b[2];
    return dl * dl + da * da + db * db;
}

/* find_nearest_with_hint, gaussian_blur.rs:57-80. */
static size_t blur_find_nearest(const float* pal, size_t n, const float color[3], size_t hint, float threshold_sq) {
    float hint_dist = sq_dist_f32(color, pal + 3 * hint);
    if (hint_dist < threshold_sq) return hint;
    size_t best = hint;
    float best_dist = hint_dist;
    for (size_t i = 0; i < n; ++i) {
        if (i == hint) continue;
        float d = sq_dist_f32(color, pal + 3 * i);
        if (d < best_dist) {
            best_dist = d;
            best = i;
        }
    }
    return best;
}

/* Rust `f32 as u8`: saturating, NaN -> 0. */
static inline uint8_t f32_as_u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

/* blur_inner / par_blur_inner (gaussian_blur.rs:100-162) along an arbitrary axis of the
 * [r][g][b][ch] volume. The reference rotates the volume between passes so the blurred axis is
 * always innermost; the arithmetic per output (taps in order, sum += kw * src, clamped edge) does
 * not depend on the memory layout, so the rotations are not restated. */
static void blur_axis(const float* src, float* dst, size_t size, size_t ch, size_t axis_stride, const float* kernel, int klen, int threads) {
    int half = klen / 2, max = (int)size - 1;
    size_t cells = size * size * size;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long cell = 0; cell < (long long)cells; ++cell) {
        size_t pos = ((size_t)cell / axis_stride) % size; /* coordinate along the blurred axis */
        for (size_t c = 0; c < ch; ++c) {
            float sum = 0.0f;
            for (int ki = 0; ki < klen; ++ki) {
                int p = (int)pos + ki - half;
                p = p < 0 ? 0 : (p > max ? max : p);
                size_t src_cell = (size_t)cell + ((size_t)p - pos) * axis_stride; /* size_t wraparound is fine */
                sum += kernel[ki] * src[src_cell * ch + c];
            }
            dst[(size_t)cell * ch + c] = sum;
        }
    }
}

/* GaussianBlurRemapper::{new, par_generate_lut_inner / generate_lut_inner, colors_to_lut}
 * (gaussian_blur.rs:34-53, 197-431, 433-500). sequential_hints = 0 follows par_generate_lut_inner
 * (what the CLI and Studio call: the NN hint restarts at 0 for every (r, g) row, l.339);
 * sequential_hints = 1 follows generate_lut_inner (the hint carries across rows, l.213). */
ORACLE_API int lutgen_oracle_blur_generate(const void* h, uint8_t level8, uint8_t* out_rgb, int sequential_hints, int threads) {
    const oracle_remapper* r = (const oracle_remapper*)h;
    if (!r || r->d.kind != K_BLUR || r->d.palette_len == 0) return -1;
    if (threads < 1) threads = 1;
    size_t n = r->d.palette_len;
    float lum = (float)r->d.lum_factor, radius = (float)r->d.param;
    int preserve = r->d.preserve;
    float* pal = (float*)malloc(sizeof(float) * 3 * n);
    for (size_t i = 0; i < n; ++i) { /* new(): [l * lum_factor, a, b] in f32 */
        oklab_t o = srgb_to_oklab(r->pal_rgb[3 * i], r->pal_rgb[3 * i + 1], r->pal_rgb[3 * i + 2]);
        pal[3 * i] = o.l * lum;
        pal[3 * i + 1] = o.a;
        pal[3 * i + 2] = o.b;
    }
    size_t level = level8, size = level * level, cells = size * size * size;
    float scale = 255.0f / (float)(size - 1);
    size_t ch = preserve ? 2 : 3;
    float step = 1.0f / (float)size, threshold_sq = step * step * 0.5f;
    float* colors = (float*)malloc(sizeof(float) * cells * ch);
    float* next = (float*)malloc(sizeof(float) * cells * ch);
    uint8_t* chan = (uint8_t*)malloc(size);
    for (size_t i = 0; i < size; ++i) chan[i] = f32_as_u8(roundf((float)i * scale));

    /* NN volume in [r][g][b] order */
    if (sequential_hints) threads = 1;
    size_t carried = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16) if (!sequential_hints)
    for (long long row = 0; row < (long long)(size * size); ++row) {
        size_t rr = (size_t)row / size, gg = (size_t)row % size;
        size_t hint = sequential_hints ? carried : 0;
        for (size_t b = 0; b < size; ++b) {
            oklab_t o = srgb_to_oklab(chan[rr], chan[gg], chan[b]);
            float color[3] = {o.l * lum, o.a, o.b};
            size_t nearest = blur_find_nearest(pal, n, color, hint, threshold_sq);
            hint = nearest;
            const float* t = pal + 3 * nearest;
            float* dst = colors + ((size_t)row * size + b) * ch;
            if (preserve) {
                dst[0] = t[1];
                dst[1] = t[2];
            } else {
                dst[0] = t[0] / lum;
                dst[1] = t[1];
                dst[2] = t[2];
            }
        }
        if (sequential_hints) carried = hint;
    }

    /* build_kernel, gaussian_blur.rs:87-96 */
    /* `(radius * 3.0).ceil() as i32`: NaN -> 0, saturating; a negative half makes `-half..=half`
     * empty, i.e. a kernel with no taps (every blurred value becomes 0.0) */
    float h3 = ceilf(radius * 3.0f);
    int half = (h3 != h3) ? 0 : (h3 > 1e6f ? 1000000 : (h3 < -1e6f ? -1000000 : (int)h3));
    float two_sigma_sq = 2.0f * radius * radius;
    int klen = half < 0 ? 0 : 2 * half + 1;
    float* kernel = (float*)malloc(sizeof(float) * (size_t)(klen + 1));
    float ksum = 0.0f;
    for (int i = -half; i <= half && klen; ++i) {
        kernel[i + half] = expf((float)(-(i * i)) / two_sigma_sq);
        ksum += kernel[i + half];
    }
    for (int i = 0; i < klen; ++i) kernel[i] /= ksum;

    int bt = r->d.kind == K_BLUR ? lutgen_oracle_max_threads() : 1;
    if (sequential_hints) bt = 1;
    blur_axis(colors, next, size, ch, 1, kernel, klen, bt);           /* pass 1: B axis */
    blur_axis(next, colors, size, ch, size * size, kernel, klen, bt); /* pass 2: R axis */
    blur_axis(colors, next, size, ch, size, kernel, klen, bt);        /* pass 3: G axis */

    /* colors_to_lut: Hald order (b outer, g, r inner) */
#pragma omp parallel for num_threads(bt) schedule(static)
    for (long long px = 0; px < (long long)cells; ++px) {
        size_t ri = (size_t)px % size, gi = ((size_t)px / size) % size, bi = (size_t)px / (size * size);
        const float* c = next + ((ri * size + gi) * size + bi) * ch;
        oklab_t o;
        if (preserve) {
            oklab_t own = srgb_to_oklab(chan[ri], chan[gi], chan[bi]);
            o.l = own.l;
            o.a = c[0];
            o.b = c[1];
        } else {
            o.l = c[0];
            o.a = c[1];
            o.b = c[2];
        }
        oklab_to_srgb(o, out_rgb + 3 * px);
    }
    free(pal);
    free(colors);
    free(next);
    free(chan);
    free(kernel);
    return 0;
}

ORACLE_API int lutgen_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * Palette extraction (`lutgen extract`, crates/cli/src/main.rs:772-815).
 *
 * PARITY UNPINNED against the reference: its `quantette` dependency (PalettePipeline, ColorSpace::Oklab,
 * QuantizeMethod::kmeans) is not vendored, no test or fixture of the reference pins its output, and its
 * sampled online k-means cannot be restated from the call site. This is the restatement of the
 * deterministic k-means SPECIFIED in lutgen_rs_b200/csrc/extract_core.h (distinct colours weighted by
 * pixel count, 16.16 fixed-point Oklab, weighted-maximin seeding, Lloyd iterations in integers), written
 * independently of the kernels: a sorted list of distinct colours and plain loops.
 * Returns the number of colours written to out_rgb (at most color_count).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t key, w;
    int32_t q[3];
} extract_point;

static uint64_t extract_dist(const int32_t* a, const int32_t* b) {
    uint64_t d = 0;
    for (int c = 0; c < 3; ++c) {
        int64_t t = (int64_t)a[c] - (int64_t)b[c];
        d += (uint64_t)(t * t);
    }
    return d;
}

ORACLE_API uint32_t lutgen_oracle_extract_palette(const uint8_t* rgb, size_t npix, uint32_t color_count, uint32_t iterations, uint8_t* out_rgb) {
    init_tables();
    if (npix == 0 || color_count == 0) return 0;
    if (color_count > 256) color_count = 256;
    uint32_t* counts = (uint32_t*)calloc((size_t)1 << 24, sizeof(uint32_t));
    for (size_t i = 0; i < npix; ++i) counts[(uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16)]++;
    size_t n = 0;
    for (uint32_t key = 0; key < (1u << 24); ++key) n += counts[key] != 0;
    extract_point* p = (extract_point*)malloc(n * sizeof(extract_point));
    uint64_t* dmin = (uint64_t*)malloc(n * sizeof(uint64_t));
    size_t m = 0;
    for (uint32_t key = 0; key < (1u << 24); ++key) { /* ascending key: the first maximum found is the tie winner */
        if (!counts[key]) continue;
        oklab_t o = srgb_to_oklab((uint8_t)key, (uint8_t)(key >> 8), (uint8_t)(key >> 16));
        p[m].key = key;
        p[m].w = counts[key];
        p[m].q[0] = (int32_t)lrintf(o.l * 65536.0f);
        p[m].q[1] = (int32_t)lrintf(o.a * 65536.0f);
        p[m].q[2] = (int32_t)lrintf(o.b * 65536.0f);
        dmin[m] = UINT64_MAX;
        ++m;
    }
    free(counts);
    int32_t cent[256][3];
    uint32_t k = 0;
    /* seeding */
    {
        size_t best = 0;
        for (size_t i = 1; i < n; ++i)
            if (p[i].w > p[best].w) best = i;
        memcpy(cent[0], p[best].q, sizeof cent[0]);
        k = 1;
    }
    while (k < color_count) {
        uint64_t best_score = 0;
        size_t best = 0;
        for (size_t i = 0; i < n; ++i) {
            uint64_t d = extract_dist(p[i].q, cent[k - 1]);
            if (d < dmin[i]) dmin[i] = d;
            uint64_t w = p[i].w < 65535u ? p[i].w : 65535u;
            uint64_t score = w * dmin[i];
            if (score > best_score) {
                best_score = score;
                best = i;
            }
        }
        if (best_score == 0) break; /* every point sits on a centroid */
        memcpy(cent[k], p[best].q, sizeof cent[k]);
        ++k;
    }
    /* Lloyd */
    for (uint32_t it = 0; it < iterations; ++it) {
        int64_t sum[256][4];
        memset(sum, 0, sizeof sum);
        for (size_t i = 0; i < n; ++i) {
            uint32_t bj = 0;
            uint64_t bd = extract_dist(p[i].q, cent[0]);
            for (uint32_t j = 1; j < k; ++j) {
                uint64_t d = extract_dist(p[i].q, cent[j]);
                if (d < bd) {
                    bd = d;
                    bj = j;
                }
            }
            for (int c = 0; c < 3; ++c) sum[bj][c] += (int64_t)p[i].w * p[i].q[c];
            sum[bj][3] += p[i].w;
        }
        int moved = 0;
        for (uint32_t j = 0; j < k; ++j) {
            int64_t w = sum[j][3];
            if (w <= 0) continue;
            for (int c = 0; c < 3; ++c) {
                int64_t s = sum[j][c];
                int32_t v = (int32_t)(s >= 0 ? (2 * s + w) / (2 * w) : -((-2 * s + w) / (2 * w)));
                moved |= v != cent[j][c];
                cent[j][c] = v;
            }
        }
        if (!moved) break; /* a fixed point: further passes change nothing */
    }
    for (uint32_t j = 0; j < k; ++j) {
        oklab_t c = {(float)cent[j][0] * (1.0f / 65536.0f), (float)cent[j][1] * (1.0f / 65536.0f), (float)cent[j][2] * (1.0f / 65536.0f)};
        oklab_to_srgb(c, out_rgb + 3 * j);
    }
    free(p);
    free(dmin);
    return k;
}
