// kernels_exact.cuh -- identity, palette setup, and the "exact" remap kernels.
//
// The exact kernels follow the reference's arithmetic operation for operation
// (f32 Oklab without contraction, f64 distances / weights / accumulation in the
// reference's order), one thread per LUT entry or pixel. They are the parity
// anchor: NearestNeighbor results are byte-identical to the CPU oracle, and the
// throughput RBF kernel (kernels_fast.cuh) sends every entry it cannot prove
// safe (near-ties at the nearest-k boundary) back through these.
#pragma once
#include "color.cuh"

namespace lutgen {

constexpr int LUTGEN_MAX_PEERS = 7;

enum FrontMode { FRONT_LUT = 0, FRONT_IMAGE = 1, FRONT_TABLE = 2, FRONT_LIST = 3 };

// Where a thread's colour comes from and where its result goes.
//  FRONT_LUT   : entry e = begin + i of a level-L Hald identity (identity.rs:7-34);
//                result -> out[3e .. 3e+3)            (GenerateLut, lib.rs:135-147)
//  FRONT_IMAGE : pixel i of an RGBA8 buffer, in place, alpha kept (mod.rs:30-50)
//  FRONT_TABLE : colour i of the full 2^24 sRGB cube (r = i & 255, g, b);
//                result -> ((uint32_t*)out)[i] = r | g<<8 | b<<16
//  FRONT_LIST  : the colours are the entry / pixel / table ids in `list` (flagged by the
//                fast kernel); list_mode (FRONT_LUT / FRONT_IMAGE / FRONT_TABLE) selects
//                how an id maps to a colour and where the result is stored.
struct Front {
    uint8_t* out;
    unsigned long long begin;
    unsigned long long count;
    uint32_t cs;     // level^2
    uint32_t csm1;   // level^2 - 1
    const uint32_t* list;
    const uint32_t* list_count;
    int list_mode;
    // FRONT_LUT only: further copies of the CLUT that receive every stored entry -- the other ranks'
    // buffers, mapped into this process (CUDA IPC) and written over NVLink while the kernel runs
    uint8_t* extra[LUTGEN_MAX_PEERS];
    int nextra;
    // ... and optionally the same buffers as ONE multicast address (NVSwitch replicates a store to every
    // GPU, this one included): whole tiles are stored there once, as words, instead of to out + extra
    uint8_t* mc;
    // FRONT_LUT only: this launch is shard `shard_index` of `shard_count` of the range. Kernels that
    // walk tiles take every shard_count-th hand-out (balanced whatever the colours cost); the others
    // are given the shard's contiguous rows instead (api.cu: rows_of_shard).
    uint32_t shard_index, shard_count;
    // GaussianBlurRemapper only: generate_lut_inner's hint chain across rows instead of par_generate_lut_inner's
    int sequential;
    // FRONT_LUT, k_remap_warp only: the exchange's barrier, done by the kernel's last CTA once every CTA's stores to
    // the peers are out. bar_flags: DEVICE table of bar_nranks pointers, entry i = rank i's flag array
    // (LUTGEN_BAR_WORDS u32, zero at the start, mapped into this process): slot t < 8 is written by rank t with
    // rank t's step number, word LUTGEN_BAR_EPOCH is the owner's own step counter. bar_nranks = 0: no barrier.
    unsigned* const* bar_flags;
    int bar_nranks, bar_rank;
};
constexpr int LUTGEN_BAR_EPOCH = 8;
constexpr int LUTGEN_BAR_WORDS = 16;

struct RemapParams {
    const double* pal_lab;  // N x 3: [l*lum, a, b], f32 Oklab widened (rbf.rs:29-35)
    const float* pal_ab;    // N x 2 f32 (nearest_neighbor.rs:25-27)
    const uint8_t* pal_rgb; // N x 4 (rgb + pad)
    uint32_t n;
    uint32_t nearest;       // effective k (0 => every colour, rbf.rs:38)
    int preserve;
    double lum;
    double param;           // shape / power
};

struct PixelIO {
    uint32_t r, g, b;
    unsigned long long slot;  // entry index (LUT), pixel index (IMAGE) or colour (TABLE)
    int mode;                 // FRONT_LUT / FRONT_IMAGE / FRONT_TABLE
};

__device__ __forceinline__ void front_decode(const Front& f, int mode, unsigned long long id, PixelIO& p) {
    p.slot = id;
    p.mode = mode;
    if (mode == FRONT_IMAGE) {
        uchar4 v = reinterpret_cast<const uchar4*>(f.out)[id];
        p.r = v.x; p.g = v.y; p.b = v.z;
    } else if (mode == FRONT_TABLE) {
        p.r = (uint32_t)(id & 255u); p.g = (uint32_t)((id >> 8) & 255u); p.b = (uint32_t)((id >> 16) & 255u);
    } else {
        uint32_t red = (uint32_t)(id % f.cs), green = (uint32_t)((id / f.cs) % f.cs), blue = (uint32_t)(id / ((unsigned long long)f.cs * f.cs));
        p.r = hald_channel(red, f.csm1); p.g = hald_channel(green, f.csm1); p.b = hald_channel(blue, f.csm1);
    }
}

// Colour number i of this launch (false: past the end).
template <int MODE>
__device__ __forceinline__ bool front_load(const Front& f, unsigned long long i, PixelIO& p) {
    if (MODE == FRONT_LIST) {
        if (i >= (unsigned long long)*f.list_count) return false;
        front_decode(f, f.list_mode, f.list[i], p);
        return true;
    }
    if (i >= f.count) return false;
    front_decode(f, MODE, f.begin + i, p);
    return true;
}

__device__ __forceinline__ void front_store_mode(const Front& f, int mode, const PixelIO& p, uint32_t r, uint32_t g, uint32_t b) {
    if (mode == FRONT_TABLE) {
        reinterpret_cast<uint32_t*>(f.out)[p.slot] = r | (g << 8) | (b << 16);
    } else if (mode == FRONT_IMAGE) {
        uchar4* q = reinterpret_cast<uchar4*>(f.out) + p.slot;
        uchar4 v = *q;
        v.x = (unsigned char)r; v.y = (unsigned char)g; v.z = (unsigned char)b;
        *q = v;
    } else {
        uint8_t* q = f.out + 3ull * p.slot;
        q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b;
#pragma unroll
        for (int x = 0; x < LUTGEN_MAX_PEERS; ++x) {  // static indices: a dynamic one would move the whole parameter block to local memory
            if (x >= f.nextra) break;
            uint8_t* qx = f.extra[x] + 3ull * p.slot;
            qx[0] = (uint8_t)r; qx[1] = (uint8_t)g; qx[2] = (uint8_t)b;
        }
    }
}

template <int MODE>
__device__ __forceinline__ void front_store(const Front& f, const PixelIO& p, uint32_t r, uint32_t g, uint32_t b) {
    front_store_mode(f, MODE == FRONT_LIST ? p.mode : MODE, p, r, g, b);
}

// Threads of a launch walk colours i, i + stride, ...: one colour each for the direct front-ends
// (the grid covers the range), a grid-stride loop for FRONT_LIST (its length lives on the device).
template <int MODE>
__device__ __forceinline__ unsigned long long front_stride() {
    return MODE == FRONT_LIST ? (unsigned long long)gridDim.x * blockDim.x : ~0ull >> 1;
}

// ---------------------------------------------------------------------------
// identity::generate (crates/lib/src/identity.rs:7-34), entries [begin, begin+count).
// One thread writes 4 entries = 12 bytes = three aligned 32-bit words when the
// range is 4-aligned; the tail and unaligned heads fall back to byte stores.
__global__ void k_identity(uint8_t* __restrict__ out, unsigned long long begin, unsigned long long count, uint32_t cs,
                           uint32_t csm1) {
    unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 4ull;
    if (i >= count) return;
    unsigned long long e = begin + i;
    uint32_t v[12];
    int n = (count - i) < 4ull ? (int)(count - i) : 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        unsigned long long ej = e + j;
        uint32_t red = (uint32_t)(ej % cs), green = (uint32_t)((ej / cs) % cs), blue = (uint32_t)(ej / ((unsigned long long)cs * cs));
        v[3 * j] = hald_channel(red, csm1);
        v[3 * j + 1] = hald_channel(green, csm1);
        v[3 * j + 2] = hald_channel(blue, csm1);
    }
    uint8_t* q = out + 3ull * e;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(q) & 3u) == 0)) {
        uint32_t* w = reinterpret_cast<uint32_t*>(q);
        w[0] = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        w[1] = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
        w[2] = v[8] | (v[9] << 8) | (v[10] << 16) | (v[11] << 24);
    } else {
        for (int j = 0; j < 3 * n; ++j) q[j] = (uint8_t)v[j];
    }
}

// Palette -> Oklab, once per remapper, with the SAME exact conversion the pixel
// path uses so `palette.contains(&color)` (rbf.rs:66) compares like with like.
// RBFRemapper::with_function rbf.rs:29-35; NearestNeighborRemapper::new nearest_neighbor.rs:22-28.
__global__ void k_palette_setup(const uint8_t* __restrict__ pal_rgb4, uint32_t n, double lum, const float* __restrict__ decode,
                                double* __restrict__ pal_lab, float* __restrict__ pal_ab, float4* __restrict__ pal_lab32) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lab o = linear_to_oklab_exact(decode[pal_rgb4[4 * i]], decode[pal_rgb4[4 * i + 1]], decode[pal_rgb4[4 * i + 2]]);
    double l = __dmul_rn((double)o.l, lum);
    pal_lab[3 * i] = l;
    pal_lab[3 * i + 1] = (double)o.a;
    pal_lab[3 * i + 2] = (double)o.b;
    pal_ab[2 * i] = o.a;
    pal_ab[2 * i + 1] = o.b;
    pal_lab32[i] = make_float4((float)l, o.a, o.b, 0.0f);
}

// kiddo SquaredEuclidean::dist: fold(0, acc + (a-b)*(a-b)) in f64, no FMA.
__device__ __forceinline__ double sq_dist_exact(double c0, double c1, double c2, const double* __restrict__ p) {
    double d0 = __dsub_rn(c0, p[0]), d1 = __dsub_rn(c1, p[1]), d2 = __dsub_rn(c2, p[2]);
    double acc = __dadd_rn(0.0, __dmul_rn(d0, d0));
    acc = __dadd_rn(acc, __dmul_rn(d1, d1));
    acc = __dadd_rn(acc, __dmul_rn(d2, d2));
    return acc;
}

// NearestNeighborRemapper::remap_pixel, nearest_neighbor.rs:46-61.
// argmin over f64 squared distances, first minimum in palette order.
template <int MODE>
__global__ void __launch_bounds__(256) k_nn_exact(Front f, RemapParams P, const ColorTables* __restrict__ tables) {
    __shared__ SmemTables st;
    stage_tables(st, tables);
    PixelIO px;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; front_load<MODE>(f, i, px); i += front_stride<MODE>()) {
        Lab c = linear_to_oklab_exact(st.decode[px.r], st.decode[px.g], st.decode[px.b]);
        double c0 = __dmul_rn((double)c.l, P.lum), c1 = (double)c.a, c2 = (double)c.b;
        uint32_t best = 0;
        double bd = INFINITY;
        for (uint32_t j = 0; j < P.n; ++j) {
            double d = sq_dist_exact(c0, c1, c2, P.pal_lab + 3 * j);
            if (d < bd) { bd = d; best = j; }
        }
        if (!P.preserve) {
            const uint8_t* q = P.pal_rgb + 4 * best;
            front_store<MODE>(f, px, q[0], q[1], q[2]);
        } else {
            Lab o = {c.l, P.pal_ab[2 * best], P.pal_ab[2 * best + 1]};
            float r, g, b;
            oklab_to_linear_exact(o, r, g, b);
            front_store<MODE>(f, px, f32_to_srgb8(r, st.encode2), f32_to_srgb8(g, st.encode2), f32_to_srgb8(b, st.encode2));
        }
    }
}

// RadialBasisFn impls, rbf.rs:153-180.
template <int KIND>
__device__ __forceinline__ double radial_basis_exact(double d, double param) {
    if (KIND == 2) return d;                              // LinearFn      rbf.rs:158
    if (KIND == 1) return 1.0 / pow(sqrt(d), param);      // InverseDistanceFn rbf.rs:167
    return exp(__dmul_rn(-param, d));                     // GaussianFn    rbf.rs:178
}

// The k <= RK nearest colours of a flagged entry (list mode: the fast kernel only flags entries whose
// distances are finite), kept sorted in REGISTERS: the list pass is a few thousand threads, each
// one entry, and with the sorted array in local memory every insertion was a chain of dependent
// loads and stores (47 % of the pass's instructions, 0.1 IPC). Same order as the array version:
// ascending distance, a later colour at an equal distance after an earlier one, so the f64 sums
// below run in the oracle's order. The list holds RK >= k candidates; the first k are summed.
template <int KIND, int RK>
__device__ __forceinline__ void rbf_exact_topk_regs(const RemapParams& P, double c0, double c1, double c2, double& n0, double& n1, double& n2,
                                                    double& den) {
    double td[RK];
    uint32_t tj[RK];
#pragma unroll
    for (int i = 0; i < RK; ++i) { td[i] = INFINITY; tj[i] = 0; }
    for (uint32_t j = 0; j < P.n; ++j) {
        const double d = sq_dist_exact(c0, c1, c2, P.pal_lab + 3 * j);
        if (d < td[RK - 1]) {
#pragma unroll
            for (int i = RK - 1; i >= 1; --i) {
                const bool shift = td[i - 1] > d, place = !shift && td[i] > d;
                td[i] = shift ? td[i - 1] : (place ? d : td[i]);
                tj[i] = shift ? tj[i - 1] : (place ? j : tj[i]);
            }
            const bool place0 = td[0] > d;
            td[0] = place0 ? d : td[0];
            tj[0] = place0 ? j : tj[0];
        }
    }
#pragma unroll
    for (int a = 0; a < RK; ++a) {
        if ((uint32_t)a < P.nearest) {
            const double* p = P.pal_lab + 3 * tj[a];
            double w = radial_basis_exact<KIND>(td[a], P.param);
            n0 = __dadd_rn(n0, __dmul_rn(p[0], w));
            n1 = __dadd_rn(n1, __dmul_rn(p[1], w));
            n2 = __dadd_rn(n2, __dmul_rn(p[2], w));
            den = __dadd_rn(den, w);
        }
    }
}

// RBFRemapper::remap_pixel, rbf.rs:57-111, operation for operation in f64.
// nearest_n is brute force: k passes, each extracting the next (distance, index)
// in ascending lexicographic order, so the f64 accumulation order equals the
// oracle's (kiddo returns neighbours by ascending distance).
template <int KIND, int MODE>
__device__ __forceinline__ void rbf_exact_one(const Front& f, const RemapParams& P, const SmemTables& st, const PixelIO& px) {
    Lab c = linear_to_oklab_exact(st.decode[px.r], st.decode[px.g], st.decode[px.b]);
    double c0 = __dmul_rn((double)c.l, P.lum), c1 = (double)c.a, c2 = (double)c.b;
    const uint32_t n = P.n;
    // rbf.rs:66-68 palette.contains(&color): pixel left untouched
    bool hit = false;
    for (uint32_t j = 0; j < n; ++j) {
        const double* p = P.pal_lab + 3 * j;
        hit |= (p[0] == c0 && p[1] == c1 && p[2] == c2);
    }
    if (hit) {
        if ((MODE == FRONT_LIST ? px.mode : MODE) != FRONT_IMAGE) front_store<MODE>(f, px, px.r, px.g, px.b);
        return;
    }
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, den = 0.0;
    if (P.nearest == 0) {
        for (uint32_t j = 0; j < n; ++j) {
            const double* p = P.pal_lab + 3 * j;
            double w = radial_basis_exact<KIND>(sq_dist_exact(c0, c1, c2, p), P.param);
            n0 = __dadd_rn(n0, __dmul_rn(p[0], w));
            n1 = __dadd_rn(n1, __dmul_rn(p[1], w));
            n2 = __dadd_rn(n2, __dmul_rn(p[2], w));
            den = __dadd_rn(den, w);
        }
    } else if (MODE == FRONT_LIST && P.nearest <= 16) {
        rbf_exact_topk_regs<KIND, 16>(P, c0, c1, c2, n0, n1, n2, den);
    } else if (P.nearest <= 64) {
        // One pass over the palette keeping the k best (distance, index) pairs sorted ascending in
        // a per-thread array: a later colour at an equal distance sorts after an earlier one, so
        // the array ends as kiddo's nearest_n order and the f64 sums run in the oracle's order.
        double td[64];
        uint32_t tj[64];
        const uint32_t k = P.nearest;
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < n; ++j) {
            double d = sq_dist_exact(c0, c1, c2, P.pal_lab + 3 * j);
            if (cnt < k || d < td[cnt - 1]) {
                uint32_t i = cnt < k ? cnt : k - 1;
                while (i > 0 && td[i - 1] > d) {
                    td[i] = td[i - 1];
                    tj[i] = tj[i - 1];
                    --i;
                }
                td[i] = d;
                tj[i] = j;
                if (cnt < k) ++cnt;
            }
        }
        for (uint32_t a = 0; a < cnt; ++a) {
            const double* p = P.pal_lab + 3 * tj[a];
            double w = radial_basis_exact<KIND>(td[a], P.param);
            n0 = __dadd_rn(n0, __dmul_rn(p[0], w));
            n1 = __dadd_rn(n1, __dmul_rn(p[1], w));
            n2 = __dadd_rn(n2, __dmul_rn(p[2], w));
            den = __dadd_rn(den, w);
        }
    } else {
        // k > 64: k passes, each extracting the next (distance, index) in ascending order
        double last_d = -1.0;
        uint32_t last_j = 0;
        bool first = true;
        for (uint32_t a = 0; a < P.nearest; ++a) {
            double bd = INFINITY;
            uint32_t bj = n;
            for (uint32_t j = 0; j < n; ++j) {
                double d = sq_dist_exact(c0, c1, c2, P.pal_lab + 3 * j);
                bool after = first || d > last_d || (d == last_d && j > last_j);
                if (after && (d < bd || (bj == n && d == bd))) { bd = d; bj = j; }
            }
            if (bj == n) break;  // NaN distances: nothing orderable left
            first = false; last_d = bd; last_j = bj;
            const double* p = P.pal_lab + 3 * bj;
            double w = radial_basis_exact<KIND>(bd, P.param);
            n0 = __dadd_rn(n0, __dmul_rn(p[0], w));
            n1 = __dadd_rn(n1, __dmul_rn(p[1], w));
            n2 = __dadd_rn(n2, __dmul_rn(p[2], w));
            den = __dadd_rn(den, w);
        }
    }
    Lab o;
    o.l = P.preserve ? __double2float_rn(__ddiv_rn(c0, P.lum)) : __double2float_rn(__ddiv_rn(__ddiv_rn(n0, den), P.lum));
    o.a = __double2float_rn(__ddiv_rn(n1, den));
    o.b = __double2float_rn(__ddiv_rn(n2, den));
    float r, g, b;
    oklab_to_linear_exact(o, r, g, b);
    front_store<MODE>(f, px, f32_to_srgb8(r, st.encode2), f32_to_srgb8(g, st.encode2), f32_to_srgb8(b, st.encode2));
}

template <int KIND, int MODE>
__global__ void __launch_bounds__(256) k_rbf_exact(Front f, RemapParams P, const ColorTables* __restrict__ tables) {
    __shared__ SmemTables st;
    stage_tables(st, tables);
    PixelIO px;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; front_load<MODE>(f, i, px); i += front_stride<MODE>())
        rbf_exact_one<KIND, MODE>(f, P, st, px);
}

// oklab::srgb_to_oklab / oklab_to_srgb for arrays (C ABI utility + parity tests).
__global__ void k_srgb_to_oklab(const uint8_t* __restrict__ rgb, size_t n, float* __restrict__ out, const ColorTables* __restrict__ tables) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lab o = linear_to_oklab_exact(tables->decode[rgb[3 * i]], tables->decode[rgb[3 * i + 1]], tables->decode[rgb[3 * i + 2]]);
    out[3 * i] = o.l; out[3 * i + 1] = o.a; out[3 * i + 2] = o.b;
}
__global__ void k_oklab_to_srgb(const float* __restrict__ lab, size_t n, uint8_t* __restrict__ out, const ColorTables* __restrict__ tables) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lab c = {lab[3 * i], lab[3 * i + 1], lab[3 * i + 2]};
    float r, g, b;
    oklab_to_linear_exact(c, r, g, b);
    out[3 * i] = (uint8_t)f32_to_srgb8(r, tables->encode2);
    out[3 * i + 1] = (uint8_t)f32_to_srgb8(g, tables->encode2);
    out[3 * i + 2] = (uint8_t)f32_to_srgb8(b, tables->encode2);
}

}  // namespace lutgen
