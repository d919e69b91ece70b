// api.cu -- the extern "C" boundary declared in include/lutgen_cuda.h.
//
// Host-side plumbing only: argument validation (returning status codes where
// the reference panics), device buffers, streams, host<->device copies and
// kernel dispatch. All arithmetic lives in the kernels_*.cuh headers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lutgen_cuda.h"
#include "kernels_apply.cuh"
#include "blur_launch.h"
#include "png_launch.h"
#include "png_decode_launch.h"
#include "png_decode_core.h"
#include "kernels_exact.cuh"
#include "kernels_fast.cuh"
#include "kernels_warp.cuh"
#include "kernels_extract.cuh"

using namespace lutgen;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                       \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(LUTGEN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK()                                                                                     \
    do {                                                                                                   \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                                \
        cudaError_t _e = cudaGetLastError();                                                               \
        if (_e != cudaSuccess)                                                                             \
            return fail(LUTGEN_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int kBatchStreams = 4;

struct Ctx {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t batch_streams[kBatchStreams] = {};
    ColorTables* tables = nullptr;  // device
    DevBuf lut_rgb;                 // RGB8 CLUT (host-API scratch)
    DevBuf lut_rgbx;                // expanded CLUT for apply
    DevBuf frame;                   // RGBA8 image scratch
    DevBuf batch_frames[kBatchStreams];
    DevBuf misc;
    DevBuf png_in, png_out, png_scratch;  // PNG writer: pixels (host API), file bytes, per-piece scratch
    DevBuf pngd_z, pngd_scratch, pngd_out;  // PNG reader: the zlib stream (+ piece offsets, palette), scratch, pixels (host API)
    int* pngd_status = nullptr;          // page-locked, 4 ints
    DevBuf extract;                       // palette extraction: histogram, point list, centroids
    uint32_t* stats = nullptr;      // device [128], only with LUTGEN_CUDA_STATS set
    // k_remap_warp's tile / finished-CTA counter pairs: zero between launches (the kernel's last CTA resets its
    // pair), handed out round robin so that launches in flight on different streams do not share one
    uint32_t* counters = nullptr;
    std::atomic<uint32_t> counter_next{0};
    uint32_t* strip_counters = nullptr;   // kStripRings x kStripMax, zero between launches (k_remap_warp's strip exchange)
    std::atomic<uint32_t> strip_next{0};
    int index = 0;                  // position in g_ctx
};
constexpr uint32_t kCounterPairs = 1024;
constexpr uint32_t kStripMax = 65536, kStripRings = 4;   // strips of a level-32 CLUT; launches in flight per GPU

// One context per GPU this process drives (ONE process, up to 8 devices). g_ctx[0] is the primary one:
// every entry point runs on it unless it shards its work over the others itself (generate_lut, apply_hald_batch).
// `g` is the calling thread's current context; UseCtx switches it (and the CUDA device) for a scope.
constexpr int kMaxDevices = 8;
Ctx g_ctx[kMaxDevices];
int g_ndev = 0;
thread_local Ctx* tl_ctx = &g_ctx[0];
#define g (*tl_ctx)

struct UseCtx {
    Ctx* prev;
    explicit UseCtx(int i) : prev(tl_ctx) {
        tl_ctx = &g_ctx[i];
        cudaSetDevice(tl_ctx->device);
    }
    ~UseCtx() {
        tl_ctx = prev;
        if (prev->ready) cudaSetDevice(prev->device);
    }
};

// One host thread per additional GPU, bound to its context for life: the entry points that shard their work over
// the process's GPUs hand every device its share as a job; the calling thread does device 0's share itself. (One
// thread issuing for all devices would serialise on every copy to or from pageable host memory -- such copies block
// their caller -- and a caller's RgbImage is pageable.)
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    std::string err;
};
Worker* g_workers[kMaxDevices] = {};

void worker_main(int index) {
    tl_ctx = &g_ctx[index];
    cudaSetDevice(tl_ctx->device);
    Worker& w = *g_workers[index];
    std::unique_lock<std::mutex> lk(w.mu);
    for (;;) {
        w.cv.wait(lk, [&] { return w.has_job || w.quit; });
        if (w.quit) return;
        lk.unlock();
        int rc = w.job();
        lk.lock();
        w.rc = rc;
        w.err = rc ? g_err : std::string();
        w.has_job = false;
        w.done = true;
        w.cv.notify_all();
    }
}

void worker_submit(int index, std::function<int()> job) {
    Worker& w = *g_workers[index];
    std::lock_guard<std::mutex> lk(w.mu);
    w.job = std::move(job);
    w.has_job = true;
    w.done = false;
    w.cv.notify_all();
}

int worker_wait(int index) {
    Worker& w = *g_workers[index];
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.done; });
    w.done = false;
    if (w.rc) g_err = w.err;
    return w.rc;
}

void workers_start(int ndev) {
    for (int i = 1; i < ndev; ++i) {
        g_workers[i] = new Worker();
        g_workers[i]->th = std::thread(worker_main, i);
    }
}

void workers_stop() {
    for (int i = 1; i < kMaxDevices; ++i) {
        Worker* w = g_workers[i];
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
        g_workers[i] = nullptr;
    }
}

int ensure_ready() {
    if (!g.ready) return fail(LUTGEN_ERR_NO_DEVICE, "lutgen_cuda_init has not succeeded in this process (no CPU fallback exists)");
    cudaError_t e = cudaSetDevice(g.device);
    if (e != cudaSuccess) return fail(LUTGEN_ERR_CUDA, "cudaSetDevice(%d): %s", g.device, cudaGetErrorString(e));
    return LUTGEN_OK;
}

inline bool level_ok(uint8_t level) { return level >= 2 && level <= 33; }
inline size_t lut_entries(uint8_t level) {
    size_t l = level;
    return l * l * l * l * l * l;
}
inline unsigned blocks_for(unsigned long long n, unsigned per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace

// ---------------------------------------------------------------------------
// remapper object
// ---------------------------------------------------------------------------
struct lutgen_remapper {
    lutgen_remapper_desc d;
    uint32_t n = 0;
    uint32_t k_eff = 0;              // 0: every colour
    std::vector<uint8_t> pal_host;   // n x 3
    int ctx_index = 0;               // the context (GPU) this object's device state lives on
    // the same remapper on the process's other GPUs, made on first use by the entry points that shard their work over
    // all of them (generate_lut); replica[0] is unused
    mutable lutgen_remapper* replica[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* arena = nullptr;           // one stream-ordered allocation holding every palette array below (and the upal_* ones)
    uint8_t* pal_rgb4 = nullptr;     // device n x 4
    double* pal_lab = nullptr;       // device n x 3
    float* pal_ab = nullptr;         // device n x 2
    float4* pal_lab32 = nullptr;     // device n x float4 (l*lum, a, b, 0)
    // sampling
    std::vector<double> noise;       // iterations x 4 (host)
    short4* offsets = nullptr;       // device, iteration order
    mutable uint32_t* nn_table = nullptr;  // device 2^24 x u32, built lazily
    mutable bool nn_table_ready = false;
    mutable std::mutex mu;
    mutable std::mutex table_mu;     // held for the whole of a table build (nn_table / full_table): concurrent callers build it once
    mutable cudaEvent_t nn_table_ev = nullptr;
    // whole-cube table of the remapper itself (r | g<<8 | b<<16 per sRGB colour): large images
    // are remapped as "build the level-16 CLUT once, then gather" (remap_pixel is a pure function
    // of the colour; alpha is untouched)
    mutable uint32_t* full_table = nullptr;
    mutable bool full_table_ready = false;
    mutable cudaEvent_t full_table_ev = nullptr;
    // fast-path work list (flagged entries), device
    mutable uint32_t* flag_list = nullptr;
    mutable uint32_t* flag_count = nullptr;
    mutable size_t flag_cap = 0;
    bool fast_ok = false;
    // NearestNeighbor fast path: palette with repeated colours dropped (first occurrence wins,
    // exactly what a strict `<` argmin returns), so duplicates never look like ties
    uint32_t un = 0;
    uint8_t* upal_rgb4 = nullptr;
    float* upal_ab = nullptr;
    float4* upal_lab32 = nullptr;
    double* upal_lab = nullptr;
    // GaussianBlurRemapper: f32 palette ([l*lum, a, b] and what a cell stores), tap weights, and the
    // blurred volume of the last level asked for (row shards of one CLUT reuse it)
    float* blur_pal3 = nullptr;
    float* blur_out3 = nullptr;
    float* blur_kernel = nullptr;
    int blur_klen = 0;
    mutable float* blur_vol[2] = {nullptr, nullptr};
    mutable size_t blur_vol_cap = 0;
    mutable uint8_t* blur_chan = nullptr;
    mutable int blur_level = 0;       // level whose blurred volume sits in blur_vol[blur_result]
    mutable int blur_sequential = 0;  // ... computed with generate_lut_inner's (1) or par_generate_lut_inner's (0) hint rule
    mutable uint32_t* blur_ends = nullptr;  // sequential mode: 2 x size^2 row-end hints + 1 change counter
    mutable int blur_result = 0;
    mutable cudaEvent_t blur_ev = nullptr;
};

namespace {

RemapParams params_of(const lutgen_remapper* r) {
    RemapParams P;
    P.pal_lab = r->pal_lab;
    P.pal_ab = r->pal_ab;
    P.pal_rgb = r->pal_rgb4;
    P.n = r->n;
    P.nearest = r->k_eff;
    P.preserve = r->d.preserve;
    P.lum = r->d.lum_factor;
    P.param = r->d.param;
    return P;
}

inline uint8_t f64_as_u8(double v) {  // Rust `as u8`: saturating, NaN -> 0
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

// Noise table -> integer channel offsets such that, for every c in 0..=255,
// sat_u8(round(c + noise)) == clamp(c + o, 0, 255)   (gaussian_sample.rs:57).
int derive_offsets(const lutgen_remapper* r, std::vector<short>& out) {
    size_t K = (size_t)r->d.iterations;
    out.assign(4 * K, 0);
    for (size_t k = 0; k < K; ++k) {
        for (int ch = 0; ch < 3; ++ch) {
            double nz = r->noise[4 * k + ch];
            int cand;
            if (nz != nz) cand = -256;
            else {
                double t = std::floor(nz + 0.5);
                cand = t > 256.0 ? 256 : (t < -256.0 ? -256 : (int)t);
            }
            int found = 1000;
            for (int delta : {0, -1, 1}) {
                int o = cand + delta;
                bool ok = true;
                for (int c = 0; c < 256 && ok; ++c) {
                    int want = f64_as_u8(std::round((double)c + nz));
                    int got = c + o;
                    got = got < 0 ? 0 : (got > 255 ? 255 : got);
                    ok = (want == got);
                }
                if (ok) { found = o; break; }
            }
            if (found == 1000)
                return fail(LUTGEN_ERR_UNSUPPORTED, "noise value %.17g (iteration %zu) is not representable as an integer channel offset", nz, k);
            out[4 * k + ch] = (short)found;
        }
    }
    return LUTGEN_OK;
}

int upload_offsets(lutgen_remapper* r) {
    std::vector<short> off;
    int rc = derive_offsets(r, off);
    if (rc != LUTGEN_OK) return rc;
    size_t K = (size_t)r->d.iterations;
    if (K == 0) return LUTGEN_OK;
    // second copy ordered by (blue, green, red): consecutive gathers of the order-independent
    // integer-sum mode then walk the NN table almost linearly (L1 hits)
    std::vector<size_t> order(K);
    for (size_t k = 0; k < K; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        if (off[4 * a + 2] != off[4 * b + 2]) return off[4 * a + 2] < off[4 * b + 2];
        if (off[4 * a + 1] != off[4 * b + 1]) return off[4 * a + 1] < off[4 * b + 1];
        return off[4 * a] < off[4 * b];
    });
    off.resize(8 * K);
    for (size_t k = 0; k < K; ++k)
        for (int c = 0; c < 4; ++c) off[4 * (K + k) + c] = off[4 * order[k] + c];
    CU_TRY(cudaMemcpyAsync(r->offsets, off.data(), sizeof(short) * 8 * K, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int launch_apply(uint8_t* rgba_dev, size_t npix, const uint32_t* lut_rgbx, uint8_t level, cudaStream_t s);

// ---- exact kernels ------------------------------------------------------------------------
template <int MODE>
int launch_nn_exact(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    if (MODE != FRONT_LIST && f.count == 0) return LUTGEN_OK;
    RemapParams P = params_of(r);
    unsigned grid = MODE == FRONT_LIST ? (unsigned)g.sm_count * 4u : blocks_for(f.count, 256);
    k_nn_exact<MODE><<<grid, 256, 0, s>>>(f, P, g.tables);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

template <int KIND, int MODE>
int launch_rbf_exact_km(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    if (MODE != FRONT_LIST && f.count == 0) return LUTGEN_OK;
    RemapParams P = params_of(r);
    unsigned grid = MODE == FRONT_LIST ? (unsigned)g.sm_count * 4u : blocks_for(f.count, 256);
    k_rbf_exact<KIND, MODE><<<grid, 256, 0, s>>>(f, P, g.tables);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

template <int MODE>
int launch_rbf_exact(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    switch (r->d.kind) {
        case LUTGEN_GAUSSIAN: return launch_rbf_exact_km<0, MODE>(f, r, s);
        case LUTGEN_SHEPARD: return launch_rbf_exact_km<1, MODE>(f, r, s);
        default: return launch_rbf_exact_km<2, MODE>(f, r, s);
    }
}

// ---- fast (tile-bounded) kernel + exact pass over what it flagged ---------------------------
bool fast_path_supported(const lutgen_remapper* r) {
    if (getenv("LUTGEN_CUDA_EXACT_ONLY")) return false;
    double lum = r->d.lum_factor;
    // the f32 tile bounds assume Oklab-sized coordinates; anything exotic goes to the f64 kernel
    // and the error bounds of the fast conversion are absolute in (L * lum, a, b): the CTA-tile kernel's constants
    // (FAST_TIE_TOL, its box padding) hold them up to lum = 16; the warp-tile kernels scale theirs with lum
    if (!(lum >= 1e-3 && lum <= 16.0)) return false;
    bool nn = r->d.kind == LUTGEN_NEAREST || r->d.kind == LUTGEN_SAMPLING;
    uint32_t n = nn ? r->un : r->n;
    if (n == 0 || n > 2048) return false;
    if (!nn && r->d.param != r->d.param) return false;
    return true;
}

// A launch that is "shard i of n" of its range but runs on a kernel that does not interleave tiles:
// the shard's contiguous rows (ceil(rows / n) each), as device.shard_rows does on the host side.
Front rows_of_shard(const Front& f) {
    Front o = f;
    if (f.shard_count <= 1) return o;
    uint32_t level = 2;
    while (level * level < f.cs) ++level;
    const unsigned long long row_len = (unsigned long long)level * f.cs;  // level^3 entries per row
    const unsigned long long rows = f.count / row_len, per = (rows + f.shard_count - 1) / f.shard_count;
    unsigned long long r0 = per * f.shard_index, r1 = r0 + per;
    if (r0 > rows) r0 = rows;
    if (r1 > rows) r1 = rows;
    o.begin = f.begin + r0 * row_len;
    o.count = (r1 - r0) * row_len;
    o.shard_index = 0;
    o.shard_count = 1;
    return o;
}

// Warp-tile kernels (kernels_warp.cuh): 0 = not applicable, 1 = k_remap_warp (palettes of at most 32
// colours), 2 = k_remap_warp_big (33 .. 256 colours, nearest-k with k <= 32 or NearestNeighbor).
template <int KIND>
int warp_path(const lutgen_remapper* r, uint32_t n, uint32_t k) {
    if (getenv("LUTGEN_CUDA_NO_WARP")) return 0;  // read per call: tests compare the two families of kernels
    if (KIND == 0) {
        // the expanded distance is scaled by C = shape*log2(e): keep it where f32 has room
        double C = r->d.param * 1.4426950408889634;
        if (!(C >= 1e-3 && C <= 2000.0)) return 0;
    }
    if (n <= (uint32_t)WT_MAX_N) return 1;
    if (n <= (uint32_t)WB_MAX_N && k != 0 && k <= 32) return 2;
    return 0;
}

template <int KIND, int MODE, int EPT, bool BIG>
int launch_warp(const Front& f, const lutgen_remapper* r, cudaStream_t s, const FastArgs& fa) {
    // NearestNeighbor on fine grids: tiles 8 planes deep (two groups per lane, one classification), 10 % faster.
    // Measured and rejected for the RBF kinds (the larger box leaves more colours to choose from per entry:
    // level-16 Gaussian 0.306 -> 0.342 ms) and for coarse grids (level 8: 0.076 -> 0.127 ms).
    constexpr int GROUPS_LEAN = (!BIG && MODE != FRONT_IMAGE) ? 2 : 1;
    const bool multi = MODE == FRONT_LUT && f.nextra > 0;
    const bool deep = GROUPS_LEAN == 2 && !multi && f.cs >= 128 && !getenv("LUTGEN_CUDA_WARP_SHALLOW") &&
                      (KIND == KIND_NN || getenv("LUTGEN_CUDA_WARP_DEEP"));
    WarpArgs a{};
    a.f = f;
    a.pal = fa.pal; a.pal_rgb = fa.pal_rgb; a.pal_ab = fa.pal_ab;
    a.n = fa.n; a.k = fa.k; a.preserve = fa.preserve; a.lum = fa.lum; a.inv_lum = fa.inv_lum; a.c1 = fa.c1;
    a.flag_list = fa.flag_list; a.flag_count = fa.flag_count; a.tile_counter = fa.tile_counter;
    a.stats = fa.stats; a.tables = fa.tables;
    a.pal64 = KIND == KIND_NN ? r->upal_lab : r->pal_lab;
    a.lum64 = r->d.lum_factor;
    a.param64 = r->d.param;
    a.rank_scale = (float)(32768.0 / (r->d.lum_factor * r->d.lum_factor + 1.0));   // squared Oklab distances stay below lum^2 + 1
    {
        const float es = (float)std::max(1.0, r->d.lum_factor);   // L, and the error of the fast conversion in it, are scaled by lum
        a.tie_slope = WT_TIE_SLOPE * es; a.tie_const = WT_TIE_CONST * es; a.pad = WT_BOX_PAD * es;
    }
    if (!BIG) {
        const uint32_t pair = g.counter_next.fetch_add(1, std::memory_order_relaxed) % kCounterPairs;
        a.tile_counter = g.counters + 2 * pair;
        a.done_counter = a.tile_counter + 1;
    }
    a.words_ok = 0;
    if (MODE == FRONT_LUT && (f.cs & 3u) == 0) {
        uintptr_t bits = reinterpret_cast<uintptr_t>(f.out);
        for (int x = 0; x < f.nextra; ++x) bits |= reinterpret_cast<uintptr_t>(f.extra[x]);
        bits |= reinterpret_cast<uintptr_t>(f.mc);
        a.words_ok = (bits & 3u) == 0;
        if (a.words_ok && (bits & 15u) == 0) a.words_ok = 2;   // 16-byte stores of whole hand-outs possible
        if (const char* e = getenv("LUTGEN_CUDA_STORE")) a.words_ok = std::min(a.words_ok, atoi(e));  // 0 bytes, 1 tile words, 2 hand-out rows
    }
    // BIG: the CTA's eight warp tiles form one 16 x 8 x 2 EPT CTA tile (or 256 EPT pixels)
    const uint32_t tr = BIG ? 16 : 8, tg = BIG ? 8 : 4, tb = BIG ? 2 * EPT : (deep ? 2 * EPT : EPT), tpix = (BIG ? 256u : 32u) * EPT;
    if (MODE == FRONT_IMAGE) {
        a.tiles_r = a.tiles_g = 1;
        a.inv_tiles_r = a.inv_tiles_g = 1.0f;
        a.ntiles = blocks_for(f.count, tpix);
    } else {
        uint32_t cs = f.cs;
        a.tiles_r = (cs + tr - 1) / tr;
        a.tiles_g = (cs + tg - 1) / tg;
        a.inv_tiles_r = 1.0f / (float)a.tiles_r;
        a.inv_tiles_g = 1.0f / (float)a.tiles_g;
        unsigned long long plane = (unsigned long long)cs * cs;
        // tile origins are multiples of the tile in every axis whatever the shard (see launch_fast)
        uint32_t b_lo = (uint32_t)(f.begin / plane) / tb * tb, b_hi = (uint32_t)((f.begin + f.count - 1) / plane);
        a.b_lo = b_lo;
        a.ntiles = a.tiles_r * a.tiles_g * ((b_hi - b_lo) / tb + 1);
        const unsigned long long slab = plane * tb;
        a.whole_tiles = (cs % tr == 0) && (cs % tg == 0) && (cs % tb == 0) && (f.begin % slab == 0) && (f.count % slab == 0);
    }
    a.ntiles_global = a.ntiles;
    a.shard_index = 0;
    a.shard_count = 1;
    a.deal = 1;
    a.strip_counter = nullptr;
    // hand-outs of WT_BATCH warp tiles keep neighbouring tiles in one warp (and feed the 96-byte row stores of
    // the MULTI variant); small ranges take single tiles so that every SM gets work
    const char* batch_str = getenv("LUTGEN_CUDA_BATCH");   // experiments (tools/time_shares.py, tools/time_exchange.py)
    const int batch_env = batch_str ? atoi(batch_str) : 0;
    // Several GPUs, whole CLUT, rows that are multiples of 16 bytes: the strip exchange (wt_push_strip) is possible --
    // whole strips are dealt to the shards, so what a GPU sends is 12 cs contiguous bytes at a time however small its
    // hand-outs are. Measured and NOT chosen by default (tools/time_exchange.py, level 16, 8 / 4 GPUs): sent alone, the
    // CLUT takes 90 us to arrive as 3072-byte runs and 112-121 us as the staged form's 96-byte rows, but the fused step
    // takes 127 / 148 us with strips and 127 / 134 us staged: a strip leaves only when its 32 tiles are done, its sender
    // stalls on the saturated fabric like any other warp, and the count per tile (a release + an atomic, 0.7 us) makes
    // the arithmetic slower (tools/time_multi_dest.py, half a CLUT: 174 us against 162 us). LUTGEN_CUDA_STRIPS=1 selects it.
    const char* strips_env = getenv("LUTGEN_CUDA_STRIPS");   // 0 / 1: never / whenever possible
    const bool strips = !BIG && multi && f.shard_count > 1 && a.whole_tiles && a.words_ok == 2 && (f.cs & 15u) == 0 && f.begin == 0 &&
                        f.count == (unsigned long long)f.cs * f.cs * f.cs && a.ntiles / a.tiles_r <= kStripMax &&
                        strips_env && atoi(strips_env) != 0;
    {
        const uint32_t share = a.ntiles / (MODE == FRONT_LUT && f.shard_count > 1 ? f.shard_count : 1u);
        // measured (tools/time_shares.py, level 16): hand-outs of 4 win 3 % on the whole CLUT (131072 tiles, 55 per resident
        // warp) and lose 6 / 16 / 24 % on a half / quarter / eighth of it, where the last hand-out of a warp is a larger
        // part of its work. The staged exchange keeps them while a warp has three: its rows on NVLink are 24 bytes per tile.
        const uint32_t per_warp = (multi && !strips) ? 3u : 24u;
        a.batch = share >= (uint32_t)g.sm_count * (BIG ? WT_WARPS : WL_WARPS) * per_warp * WT_BATCH ? (uint32_t)WT_BATCH : 1u;
        if (batch_env == 1 || batch_env == WT_BATCH) a.batch = (uint32_t)batch_env;
    }
    if (strips) {
        const uint32_t nstrips = a.ntiles_global / a.tiles_r;
        const uint32_t mine = nstrips > f.shard_index ? (nstrips - f.shard_index + f.shard_count - 1) / f.shard_count : 0;
        a.shard_index = f.shard_index;
        a.shard_count = f.shard_count;
        a.deal = a.tiles_r;
        a.ntiles = mine * a.tiles_r;
        if (a.tiles_r % a.batch) a.batch = 1;   // a hand-out stays inside one strip
        a.strip_counter = g.strip_counters + (size_t)(g.strip_next.fetch_add(1, std::memory_order_relaxed) % kStripRings) * kStripMax;
    } else if (MODE == FRONT_LUT && f.shard_count > 1) {
        // hand-outs (a.batch warp tiles, or one CTA tile) index, index + count, ...
        const uint32_t per = BIG ? 1u : a.batch, handouts = (a.ntiles_global + per - 1) / per;
        a.shard_index = f.shard_index;
        a.shard_count = f.shard_count;
        a.deal = per;
        a.ntiles = handouts > f.shard_index ? ((handouts - f.shard_index + f.shard_count - 1) / f.shard_count) * per : 0;
    }
    if (a.shard_count > 1 && a.ntiles == 0 && !(f.bar_nranks > 1)) {   // with a barrier the (empty) kernel still has to take part in it
        if (BIG) CU_TRY(cudaMemsetAsync(r->flag_count, 0, 2 * sizeof(uint32_t), s));
        return LUTGEN_OK;
    }
    // MULTI: the variant that stages tiles and writes them to several destinations (multi-GPU); a single
    // destination takes the leaner kernel (the extra code costs instruction-cache misses)
    const int variant = strips ? 3 : (multi ? 1 : (deep ? 2 : 0));
    static int per_sm[4] = {0, 0, 0, 0};  // per instantiation
    if (per_sm[variant] == 0) {
        int v = 0;
        if (BIG) {
            if (multi) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp_big<KIND, MODE, EPT, MODE == FRONT_LUT>, WT_THREADS, 0));
            else CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp_big<KIND, MODE, EPT, false>, WT_THREADS, 0));
        } else {
            if (strips) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp<KIND, MODE, EPT, MODE == FRONT_LUT ? 2 : 0, 1>, WL_THREADS, 0));
            else if (multi) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp<KIND, MODE, EPT, MODE == FRONT_LUT ? 1 : 0, 1>, WL_THREADS, 0));
            else if (deep) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp<KIND, MODE, EPT, 0, GROUPS_LEAN>, WL_THREADS, 0));
            else CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_remap_warp<KIND, MODE, EPT, 0, 1>, WL_THREADS, 0));
        }
        per_sm[variant] = v < 1 ? 1 : v;
    }
    unsigned grid = (unsigned)per_sm[variant] * (unsigned)g.sm_count;
    unsigned need = BIG ? a.ntiles : blocks_for(a.ntiles, WL_WARPS * a.batch);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    a.tail_cut = a.ticket_end = a.ntiles;
    if (!BIG && a.batch > 1 && (a.shard_count == 1 || strips)) {
        // the last two tiles per resident warp go out one by one (k_remap_warp's ticket scheme)
        const uint32_t singles = std::min(a.ntiles, grid * (uint32_t)WL_WARPS * 2u);
        a.tail_cut = (a.ntiles - singles) / a.batch * a.batch;
        a.ticket_end = a.tail_cut + (a.ntiles - a.tail_cut) * a.batch;
    }
    if (BIG) CU_TRY(cudaMemsetAsync(r->flag_count, 0, 2 * sizeof(uint32_t), s));
    if (BIG) {
        if (multi) k_remap_warp_big<KIND, MODE, EPT, MODE == FRONT_LUT><<<grid, WT_THREADS, 0, s>>>(a);
        else k_remap_warp_big<KIND, MODE, EPT, false><<<grid, WT_THREADS, 0, s>>>(a);
    } else {
        if (strips) k_remap_warp<KIND, MODE, EPT, MODE == FRONT_LUT ? 2 : 0, 1><<<grid, WL_THREADS, 0, s>>>(a);
        else if (multi) k_remap_warp<KIND, MODE, EPT, MODE == FRONT_LUT ? 1 : 0, 1><<<grid, WL_THREADS, 0, s>>>(a);
        else if (deep) k_remap_warp<KIND, MODE, EPT, 0, GROUPS_LEAN><<<grid, WL_THREADS, 0, s>>>(a);
        else k_remap_warp<KIND, MODE, EPT, 0, 1><<<grid, WL_THREADS, 0, s>>>(a);
    }
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

template <int KIND, int MODE>
int launch_fast(const Front& f_in, const lutgen_remapper* r, cudaStream_t s) {
    // interleaved shards exist in the warp-tile kernels only
    const bool nn_ = KIND == KIND_NN;
    const Front f = (f_in.shard_count > 1 && !warp_path<KIND>(r, nn_ ? r->un : r->n, nn_ ? 1u : r->k_eff)) ? rows_of_shard(f_in) : f_in;
    if (f.count == 0) return LUTGEN_OK;
    const int wp_kind = warp_path<KIND>(r, nn_ ? r->un : r->n, nn_ ? 1u : r->k_eff);
    if (wp_kind != 1) {   // k_remap_warp resolves what it cannot decide itself: no work list, no second pass
        std::lock_guard<std::mutex> lk(r->mu);
        if (!r->flag_count) CU_TRY(cudaMalloc((void**)&r->flag_count, 64));
        if (r->flag_cap < f.count) {
            if (r->flag_list) CU_TRY(cudaFree(r->flag_list));
            r->flag_list = nullptr;
            r->flag_cap = 0;
            size_t cap = ((size_t)f.count + 1023) & ~(size_t)1023;
            CU_TRY(cudaMalloc((void**)&r->flag_list, cap * sizeof(uint32_t)));
            r->flag_cap = cap;
        }
    }
    const bool nn = KIND == KIND_NN;
    FastArgs a{};
    a.f = f;
    a.pal = nn ? r->upal_lab32 : r->pal_lab32;
    a.pal_rgb = reinterpret_cast<const uint32_t*>(nn ? r->upal_rgb4 : r->pal_rgb4);
    a.pal_ab = nn ? r->upal_ab : r->pal_ab;
    a.n = nn ? r->un : r->n;
    a.k = nn ? 1u : r->k_eff;
    a.preserve = r->d.preserve;
    a.lum = (float)r->d.lum_factor;
    a.inv_lum = (float)(1.0 / r->d.lum_factor);
    a.stats = g.stats;
    a.c1 = KIND == 0 ? (float)(-r->d.param * 1.4426950408889634) : (KIND == 1 ? (float)(-0.5 * r->d.param) : 0.0f);
    a.flag_list = r->flag_list;
    a.flag_count = r->flag_count;
    a.tile_counter = r->flag_count + 1;
    a.tables = g.tables;
    Front fl = f;
    fl.shard_index = 0;
    fl.shard_count = 1;
    fl.list = r->flag_list;
    fl.list_count = r->flag_count;
    fl.list_mode = MODE;
    if (const int wp = wp_kind) {
        if (wp == 1) return launch_warp<KIND, MODE, 4, false>(f, r, s, a);
        int rc = launch_warp<KIND, MODE, 4, true>(f, r, s, a);
        if (rc != LUTGEN_OK) return rc;
        if (nn) return launch_nn_exact<FRONT_LIST>(fl, r, s);
        return launch_rbf_exact_km<KIND < 3 ? KIND : 0, FRONT_LIST>(fl, r, s);
    }
    unsigned grid;
    if (MODE == FRONT_IMAGE) {
        a.tr = a.tg = a.tb = 1;
        a.sh_r = a.sh_g = 0;
        a.tiles_r = a.tiles_g = 1;
        a.ept = FAST_MAX_EPT;
        grid = blocks_for(f.count, FAST_TILE_MAX);
    } else {
        uint32_t cs = f.cs;
        // brick sides: powers of two, at most 8 x 8 x 16 (1024 colours, 4 per thread)
        a.sh_r = 0;
        while (a.sh_r < 3 && (2u << a.sh_r) <= cs) ++a.sh_r;
        a.sh_g = a.sh_r;
        uint32_t sh_b = 0;
        while (sh_b < 4 && (2u << sh_b) <= cs) ++sh_b;
        a.tr = 1u << a.sh_r;
        a.tg = 1u << a.sh_g;
        a.tb = 1u << sh_b;
        a.ept = (a.tr * a.tg * a.tb + FAST_THREADS - 1) / FAST_THREADS;
        a.tiles_r = (cs + a.tr - 1) / a.tr;
        a.tiles_g = (cs + a.tg - 1) / a.tg;
        unsigned long long plane = (unsigned long long)cs * cs;
        // tile origins are multiples of the brick in every axis whatever the shard, so a colour
        // always meets the same tile (and the same classification): results are shard-independent
        uint32_t b_lo = (uint32_t)(f.begin / plane) / a.tb * a.tb, b_hi = (uint32_t)((f.begin + f.count - 1) / plane);
        a.b_lo = b_lo;
        uint32_t tiles_b = (b_hi - b_lo) / a.tb + 1;
        grid = a.tiles_r * a.tiles_g * tiles_b;
    }
    size_t smem = fast_smem_bytes(a.n, a.ept * FAST_THREADS);
    static bool attr_done[kMaxDevices] = {};  // per <KIND, MODE> instantiation and device
    if (!attr_done[g.index]) {
        CU_TRY(cudaFuncSetAttribute(k_remap_fast<KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done[g.index] = true;
    }
    // persistent CTAs: as many as fit on the device at once, striding over the tiles
    a.ntiles = grid;
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_remap_fast<KIND, MODE>, FAST_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned resident = (unsigned)per_sm * (unsigned)g.sm_count;
    if (grid > resident) grid = resident;
    CU_TRY(cudaMemsetAsync(r->flag_count, 0, 2 * sizeof(uint32_t), s));
    k_remap_fast<KIND, MODE><<<grid, FAST_THREADS, smem, s>>>(a);
    LAUNCH_CHECK();
    if (nn) return launch_nn_exact<FRONT_LIST>(fl, r, s);
    return launch_rbf_exact_km<KIND < 3 ? KIND : 0, FRONT_LIST>(fl, r, s);
}

template <int MODE>
int launch_nn(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    if (r->fast_ok) return launch_fast<KIND_NN, MODE>(f, r, s);
    return launch_nn_exact<MODE>(rows_of_shard(f), r, s);
}

template <int MODE>
int launch_rbf(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    if (!r->fast_ok) return launch_rbf_exact<MODE>(rows_of_shard(f), r, s);
    switch (r->d.kind) {
        case LUTGEN_GAUSSIAN: return launch_fast<0, MODE>(f, r, s);
        case LUTGEN_SHEPARD: return launch_fast<1, MODE>(f, r, s);
        default: return launch_fast<2, MODE>(f, r, s);
    }
}

// NN over the whole sRGB cube, once per sampling remapper (see kernels_apply.cuh).
int ensure_nn_table(const lutgen_remapper* r, cudaStream_t s) {
    std::lock_guard<std::mutex> build(r->table_mu);
    std::unique_lock<std::mutex> lk(r->mu);
    if (!r->nn_table_ready) {
        if (!r->nn_table) CU_TRY(cudaMalloc((void**)&r->nn_table, sizeof(uint32_t) << 24));
        if (!r->nn_table_ev) CU_TRY(cudaEventCreateWithFlags(&r->nn_table_ev, cudaEventDisableTiming));
        Front f{};
        f.out = reinterpret_cast<uint8_t*>(r->nn_table);
        f.begin = 0;
        f.count = 1ull << 24;
        f.cs = 256;
        f.csm1 = 255;
        lk.unlock();
        int rc = launch_nn<FRONT_TABLE>(f, r, s);
        if (rc != LUTGEN_OK) return rc;
        lk.lock();
        CU_TRY(cudaEventRecord(r->nn_table_ev, s));
        r->nn_table_ready = true;
    }
    CU_TRY(cudaStreamWaitEvent(s, r->nn_table_ev, 0));
    return LUTGEN_OK;
}

template <int MODE>
int launch_sampling(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    int rc = ensure_nn_table(r, s);
    if (rc != LUTGEN_OK) return rc;
    if (f.count == 0) return LUTGEN_OK;
    uint64_t K = r->d.iterations;
    bool pow2 = K != 0 && (K & (K - 1)) == 0 && K <= (1ull << 24);
    uint32_t log2k = 0;
    while (pow2 && (1ull << log2k) < K) ++log2k;
    if (pow2)
        k_sampling<MODE, false><<<blocks_for(f.count, 256), 256, 0, s>>>(f, r->nn_table, r->offsets + K, (uint32_t)K, log2k);
    else
        k_sampling<MODE, true><<<blocks_for(f.count, 256), 256, 0, s>>>(f, r->nn_table, r->offsets, (uint32_t)K, 0);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

// ---- GaussianBlurRemapper -----------------------------------------------------------------
// generate_lut_inner / par_generate_lut_inner, gaussian_blur.rs:197-431 (par semantics: the NN
// hint restarts per (r, g) row, which is what the CLI and Studio call).
int ensure_blur_volume(const lutgen_remapper* r, uint32_t level, int sequential, cudaStream_t s) {
    std::unique_lock<std::mutex> lk(r->mu);
    if (!r->blur_ev) CU_TRY(cudaEventCreateWithFlags(&r->blur_ev, cudaEventDisableTiming));
    if (r->blur_level != (int)level || r->blur_sequential != sequential) {
        const uint32_t size = level * level, ch = r->d.preserve ? 2u : 3u;
        const size_t cells = (size_t)size * size * size, bytes = cells * ch * sizeof(float);
        if (r->blur_vol_cap < bytes) {
            for (int i = 0; i < 2; ++i) {
                if (r->blur_vol[i]) CU_TRY(cudaFree(r->blur_vol[i]));
                r->blur_vol[i] = nullptr;
            }
            r->blur_vol_cap = 0;
            CU_TRY(cudaMalloc((void**)&r->blur_vol[0], bytes));
            CU_TRY(cudaMalloc((void**)&r->blur_vol[1], bytes));
            r->blur_vol_cap = bytes;
        }
        if (!r->blur_chan) CU_TRY(cudaMalloc((void**)&r->blur_chan, 2048));
        // (i as f32 * scale).round() as u8, scale = 255.0 / (size - 1) as f32   (gaussian_blur.rs:200, 206-210)
        std::vector<uint8_t> chan(size);
        const float scale = 255.0f / (float)(size - 1);
        for (uint32_t i = 0; i < size; ++i) {
            float v = roundf((float)i * scale);
            chan[i] = !(v > 0.0f) ? 0 : (v >= 255.0f ? 255 : (uint8_t)v);
        }
        CU_TRY(cudaMemcpyAsync(r->blur_chan, chan.data(), size, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaStreamSynchronize(s));  // `chan` is a stack-lifetime pageable buffer
        const float lum = (float)r->d.lum_factor;
        const float step = 1.0f / (float)size, threshold_sq = step * step * 0.5f;
        float* a = r->blur_vol[0];
        float* b = r->blur_vol[1];
        unsigned rows = size * size;
        unsigned grid_nn = blocks_for(rows, BLUR_ROW_WARPS);
        if (grid_nn > (unsigned)g.sm_count * 8u) grid_nn = (unsigned)g.sm_count * 8u;
        static bool attr_done[kMaxDevices] = {};
        if (!attr_done[g.index]) {
            CU_TRY(blur_set_attributes());
            attr_done[g.index] = true;
        }
        const int stage_pal = (size_t)r->n * 28 <= 64 * 1024;   // palette (3 floats) + expanded form (4 floats) per colour
        const size_t smem_nn = 1024 + (stage_pal ? (size_t)r->n * 28 + 32 : 0);
        if (!sequential) {
            blur_launch_nn_volume(grid_nn, smem_nn, s, a, size, ch, r->blur_chan, r->blur_pal3, r->blur_out3, r->n, lum, threshold_sq, g.tables, stage_pal,
                                  nullptr, nullptr, nullptr);
            LAUNCH_CHECK();
        } else {
            // generate_lut_inner: the hint runs on from row to row (gaussian_blur.rs:213). Passes until no row's
            // final hint changes; the first pass starts every row from 0 like the parallel variant does.
            if (r->blur_ends) { CU_TRY(cudaFree(r->blur_ends)); r->blur_ends = nullptr; }
            CU_TRY(cudaMalloc((void**)&r->blur_ends, sizeof(uint32_t) * (2 * (size_t)rows + 1)));
            CU_TRY(cudaMemsetAsync(r->blur_ends, 0, sizeof(uint32_t) * (2 * (size_t)rows + 1), s));
            uint32_t* ends[2] = {r->blur_ends, r->blur_ends + rows};
            uint32_t* changed = r->blur_ends + 2 * (size_t)rows;
            for (uint32_t pass = 0;; ++pass) {
                if (pass > rows + 1) return fail(LUTGEN_ERR_CUDA, "blur: the row hint chain did not settle");
                CU_TRY(cudaMemsetAsync(changed, 0, sizeof(uint32_t), s));
                blur_launch_nn_volume(grid_nn, smem_nn, s, a, size, ch, r->blur_chan, r->blur_pal3, r->blur_out3, r->n, lum, threshold_sq, g.tables,
                                      stage_pal, ends[pass & 1], ends[(pass & 1) ^ 1], changed);
                LAUNCH_CHECK();
                uint32_t nchanged = 0;
                CU_TRY(cudaMemcpyAsync(&nchanged, changed, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                CU_TRY(cudaStreamSynchronize(s));
                if (nchanged == 0) break;
            }
        }
        const int klen = r->blur_klen;
        // columns per CTA: an even number that keeps tile + taps within ~48 KB (several CTAs per SM),
        // but never fewer than one cell's channels; the B pass takes whole rows (ch columns each)
        auto launch_axis = [&](const float* src, float* dst, BlurAxis P, uint32_t outers) -> int {
            P.groups = (P.total_cols + P.cols - 1) / P.cols;
            // rows 4 apart (two position blocks handled by one warp) must not share banks: pitch = 4 mod 8 floats
            P.pitch = P.cols + ((12u - (P.cols & 7u)) & 7u);
            // two tiles (double buffer): `size` positions rounded up to blocks of 4 + the replicated rows either
            // side, each; zero-padded weights behind them
            const size_t tile = (size_t)((P.size + 3u) / 4u * 4u + 2u * (uint32_t)(klen / 2) + 3u) * P.pitch;
            size_t smem = (2 * tile + (size_t)((klen + 6) / 4 * 4 + 4)) * sizeof(float);
            if (smem > (size_t)BLUR_MAX_DYN_SMEM) return fail(LUTGEN_ERR_UNSUPPORTED, "blur: %zu bytes of shared memory per tile (level %u, radius %g)", smem, level, r->d.param);
            const uint32_t ntiles = P.groups * outers;
            unsigned per_sm = (unsigned)((227 * 1024) / (smem + 1024));
            if (per_sm > 8) per_sm = 8;
            if (per_sm < 1) per_sm = 1;
            unsigned grid = per_sm * (unsigned)g.sm_count;
            if (grid > ntiles) grid = ntiles;
            blur_launch_axis(grid, smem, s, src, dst, P, ntiles, r->blur_kernel, klen);
            LAUNCH_CHECK();
            return LUTGEN_OK;
        };
        uint32_t want = 6144u / (size + 2u * (uint32_t)(klen / 2) + 6u);   // two buffers of ~24 KB: four CTAs per SM
        if (want > 16) want = 16;
        if (want < 2) want = 2;
        {   // pass 1: B axis (innermost)
            BlurAxis P{};
            uint32_t lines = want / ch;
            if (lines < 1) lines = 1;
            if ((lines * ch) & 1u) ++lines;  // an even number of columns per full tile
            P.size = size; P.pstride = ch; P.cols = lines * ch;
            P.pitch = P.cols;  // even; the staging is asynchronous (cp.async), its bank conflicts are off the critical path
            P.cpl = ch; P.run_stride = size * ch; P.total_cols = rows * ch; P.outer_stride = 0;
            int rc = launch_axis(a, b, P, 1);
            if (rc) return rc;
        }
        {   // pass 2: R axis (stride size^2 cells)
            BlurAxis P{};
            P.size = size; P.pstride = size * size * ch; P.cols = want & ~1u; P.pitch = P.cols;
            P.cpl = 0; P.run_stride = 0; P.total_cols = size * size * ch; P.outer_stride = 0;
            int rc = launch_axis(b, a, P, 1);
            if (rc) return rc;
        }
        {   // pass 3: G axis (stride size cells), one slab per r
            BlurAxis P{};
            P.size = size; P.pstride = size * ch; P.cols = want & ~1u; P.pitch = P.cols;
            P.cpl = 0; P.run_stride = 0; P.total_cols = size * ch; P.outer_stride = size * size * ch;
            int rc = launch_axis(a, b, P, size);
            if (rc) return rc;
        }
        r->blur_result = 1;
        r->blur_level = (int)level;
        r->blur_sequential = sequential;
        CU_TRY(cudaEventRecord(r->blur_ev, s));
    }
    CU_TRY(cudaStreamWaitEvent(s, r->blur_ev, 0));
    return LUTGEN_OK;
}

int launch_blur_lut(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    uint32_t level = 2;
    while (level * level < f.cs) ++level;
    int rc = ensure_blur_volume(r, level, f.sequential, s);
    if (rc != LUTGEN_OK) return rc;
    if (f.count == 0) return LUTGEN_OK;
    const uint32_t ch = r->d.preserve ? 2u : 3u;
    {
        const uint32_t size = f.cs;
        const unsigned long long plane = (unsigned long long)size * size;
        const uint32_t b_first = (uint32_t)(f.begin / plane) / BLUR_TB * BLUR_TB, b_last = (uint32_t)((f.begin + f.count - 1) / plane);
        const uint32_t tiles_r = (size + BLUR_TR - 1) / BLUR_TR, tiles_b = (b_last - b_first) / BLUR_TB + 1;
        const uint32_t ntiles = tiles_r * size * tiles_b;
        unsigned grid = (unsigned)g.sm_count * 8u;
        if (grid > ntiles) grid = ntiles;
        blur_launch_to_lut(grid, s, r->blur_vol[r->blur_result], size, ch, r->blur_chan, f.out, f.begin, f.count, b_first, tiles_r, ntiles, g.tables);
    }
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

template <int MODE>
int launch_remap(const Front& f, const lutgen_remapper* r, cudaStream_t s);

// The remapper tabulated over all 2^24 colours (= its level-16 CLUT in packed form).
int ensure_full_table(const lutgen_remapper* r, cudaStream_t s) {
    std::lock_guard<std::mutex> build(r->table_mu);
    std::unique_lock<std::mutex> lk(r->mu);
    if (!r->full_table_ready) {
        if (!r->full_table) CU_TRY(cudaMalloc((void**)&r->full_table, sizeof(uint32_t) << 24));
        if (!r->full_table_ev) CU_TRY(cudaEventCreateWithFlags(&r->full_table_ev, cudaEventDisableTiming));
        Front f{};
        f.out = reinterpret_cast<uint8_t*>(r->full_table);
        f.begin = 0;
        f.count = 1ull << 24;
        f.cs = 256;
        f.csm1 = 255;
        lk.unlock();
        int rc = launch_remap<FRONT_TABLE>(f, r, s);
        if (rc != LUTGEN_OK) return rc;
        lk.lock();
        CU_TRY(cudaEventRecord(r->full_table_ev, s));
        r->full_table_ready = true;
    }
    CU_TRY(cudaStreamWaitEvent(s, r->full_table_ev, 0));
    return LUTGEN_OK;
}

// Remap `f.count` colours with remapper r on stream s (no synchronisation).
template <int MODE>
int launch_remap(const Front& f, const lutgen_remapper* r, cudaStream_t s) {
    if (MODE == FRONT_IMAGE) {
        // Big images: one pass over the colour cube costs less than one pass over the pixels.
        unsigned long long threshold = r->d.kind == LUTGEN_SAMPLING ? (1ull << 24) : (1ull << 20);
        if (f.count >= threshold && !getenv("LUTGEN_CUDA_NO_TABLE")) {
            int rc = ensure_full_table(r, s);
            if (rc != LUTGEN_OK) return rc;
            return launch_apply(f.out, (size_t)f.count, r->full_table, 16, s);
        }
    }
    switch (r->d.kind) {
        case LUTGEN_NEAREST: return launch_nn<MODE>(f, r, s);
        case LUTGEN_SAMPLING: return launch_sampling<MODE>(f, r, s);
        case LUTGEN_BLUR:
            if (MODE == FRONT_LUT) return launch_blur_lut(f, r, s);
            return fail(LUTGEN_ERR_INVALID, "GaussianBlurRemapper implements GenerateLut only (gaussian_blur.rs:512-541), not remap_image");
        default: return launch_rbf<MODE>(f, r, s);
    }
}

int expand_lut(const uint8_t* lut_rgb_dev, uint8_t level, cudaStream_t s) {
    size_t E = lut_entries(level);
    CU_TRY(g.lut_rgbx.reserve(E * 4));
    k_expand_lut<<<blocks_for(E, 256), 256, 0, s>>>(lut_rgb_dev, E, (uint32_t*)g.lut_rgbx.p);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

int launch_apply(uint8_t* rgba_dev, size_t npix, const uint32_t* lut_rgbx, uint8_t level, cudaStream_t s) {
    if (npix == 0) return LUTGEN_OK;
    if (reinterpret_cast<uintptr_t>(rgba_dev) & 3u) return fail(LUTGEN_ERR_INVALID, "rgba_dev must be 4-byte aligned");
    uint32_t cs = (uint32_t)level * level;
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(rgba_dev) & 15u)) & 15u) / 4;
    if (head > npix) head = npix;
    size_t nvec = (npix - head) / 4;
    size_t want = (nvec + 255) / 256;
    if (want == 0) want = 1;
    // enough resident work to cover the gather latency, capped so the grid-stride loop amortises launch
    size_t cap = (size_t)g.sm_count * 64;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    k_apply<<<grid, 256, 0, s>>>((uint32_t*)rgba_dev, npix, head, nvec, lut_rgbx, cs, cs - 1);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------
extern "C" {

const char* lutgen_cuda_last_error(void) { return g_err.c_str(); }
const char* lutgen_cuda_version(void) { return "lutgen-b200 0.1.0 (sm_100a)"; }
uint64_t lutgen_cuda_launch_count(void) { return g_launches.load(); }

int lutgen_cuda_device_count(int* count) {
    if (!count) return fail(LUTGEN_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(LUTGEN_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return LUTGEN_OK;
}

namespace {
// Everything one GPU needs: streams, colour tables, counters. `g` is the context being set up.
int init_current_ctx(int device, int visible) {
    if (device >= visible) return fail(LUTGEN_ERR_NO_DEVICE, "device %d requested but only %d visible", device, visible);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(LUTGEN_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library carries sm_100a code only", device, prop.name, prop.major,
                    prop.minor);
    CU_TRY(cudaSetDevice(device));
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    {
        // stream-ordered allocations (remapper palettes) come from the default pool; never hand memory back at a sync
        cudaMemPool_t pool;
        CU_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (int i = 0; i < kBatchStreams; ++i) CU_TRY(cudaStreamCreateWithFlags(&g.batch_streams[i], cudaStreamNonBlocking));
    ColorTables host;
    for (int i = 0; i < 256; ++i) memcpy(&host.decode[i], &k_srgb_decode_bits[i], 4);
    for (int i = 0; i < 104; ++i) {
        host.encode[i] = k_srgb_encode_table[i];
        host.encode2[i] = srgb_encode2_entry(k_srgb_encode_table[i], (uint32_t)i);
    }
    CU_TRY(cudaMalloc((void**)&g.tables, sizeof(ColorTables)));
    CU_TRY(cudaMemcpy(g.tables, &host, sizeof host, cudaMemcpyHostToDevice));
    CU_TRY(cudaMalloc((void**)&g.counters, 2 * kCounterPairs * sizeof(uint32_t)));
    CU_TRY(cudaMemset(g.counters, 0, 2 * kCounterPairs * sizeof(uint32_t)));
    CU_TRY(cudaMalloc((void**)&g.strip_counters, kStripRings * kStripMax * sizeof(uint32_t)));
    CU_TRY(cudaMemset(g.strip_counters, 0, kStripRings * kStripMax * sizeof(uint32_t)));
    if (getenv("LUTGEN_CUDA_STATS")) {
        CU_TRY(cudaMalloc((void**)&g.stats, 136 * sizeof(uint32_t)));
        CU_TRY(cudaMemset(g.stats, 0, 136 * sizeof(uint32_t)));
        CU_TRY(cudaMemset(g.stats + 128, 0xff, sizeof(uint32_t)));
    }
    g.ready = true;
    return LUTGEN_OK;
}

void release_current_ctx() {
    if (!g.ready) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    g.lut_rgb.release();
    g.lut_rgbx.release();
    g.frame.release();
    g.misc.release();
    g.png_in.release();
    g.png_out.release();
    g.png_scratch.release();
    g.pngd_z.release();
    g.pngd_scratch.release();
    g.pngd_out.release();
    if (g.pngd_status) cudaFreeHost(g.pngd_status);
    g.pngd_status = nullptr;
    g.extract.release();
    for (int i = 0; i < kBatchStreams; ++i) {
        g.batch_frames[i].release();
        if (g.batch_streams[i]) cudaStreamDestroy(g.batch_streams[i]);
        g.batch_streams[i] = nullptr;
    }
    if (g.tables) cudaFree(g.tables);
    g.tables = nullptr;
    if (g.counters) cudaFree(g.counters);
    g.counters = nullptr;
    if (g.strip_counters) cudaFree(g.strip_counters);
    g.strip_counters = nullptr;
    if (g.stats) cudaFree(g.stats);
    g.stats = nullptr;
    if (g.stream) cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    g.ready = false;
    g.device = -1;
}

std::mutex g_init_mu;
}  // namespace

int lutgen_cuda_init_devices(const int* devices, int ndevices) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!devices || ndevices < 1) return fail(LUTGEN_ERR_INVALID, "no devices given");
    if (ndevices > kMaxDevices) return fail(LUTGEN_ERR_UNSUPPORTED, "%d devices (limit %d)", ndevices, kMaxDevices);
    for (int i = 0; i < ndevices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j] || devices[i] < 0) return fail(LUTGEN_ERR_INVALID, "device list must name distinct, non-negative devices");
    if (g_ndev > 0) {
        bool same = g_ndev == ndevices;
        for (int i = 0; same && i < ndevices; ++i) same = g_ctx[i].device == devices[i];
        if (same) return LUTGEN_OK;
        return fail(LUTGEN_ERR_INVALID, "already initialised on %d device(s), the first being device %d (lutgen_cuda_shutdown first)", g_ndev,
                    g_ctx[0].device);
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(LUTGEN_ERR_NO_DEVICE, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    int rc = LUTGEN_OK;
    for (int i = 0; i < ndevices && rc == LUTGEN_OK; ++i) {
        Ctx* prev = tl_ctx;
        tl_ctx = &g_ctx[i];
        g.index = i;
        rc = init_current_ctx(devices[i], n);
        tl_ctx = prev;
    }
    if (rc != LUTGEN_OK) {
        for (int i = 0; i < ndevices; ++i) {
            Ctx* prev = tl_ctx;
            tl_ctx = &g_ctx[i];
            release_current_ctx();
            tl_ctx = prev;
        }
        return rc;
    }
    g_ndev = ndevices;
    cudaSetDevice(g_ctx[0].device);
    workers_start(ndevices);
    return LUTGEN_OK;
}

int lutgen_cuda_device_count_in_use(int* count) {
    if (!count) return fail(LUTGEN_ERR_INVALID, "count is NULL");
    *count = g_ndev;
    return LUTGEN_OK;
}

int lutgen_cuda_init(int device) {
    if (device < 0) {
        // LUTGEN_CUDA_DEVICES=0,1,2,3: this process shards generate_lut / apply_hald_batch over those GPUs (no caller changes)
        if (const char* list = getenv("LUTGEN_CUDA_DEVICES")) {
            int devs[kMaxDevices], nd = 0;
            const char* q = list;
            while (*q && nd < kMaxDevices) {
                char* end = nullptr;
                long v = strtol(q, &end, 10);
                if (end == q) break;
                devs[nd++] = (int)v;
                q = (*end == ',') ? end + 1 : end;
                if (*end != ',') break;
            }
            if (nd > 0) return lutgen_cuda_init_devices(devs, nd);
        }
        const char* e = getenv("LUTGEN_CUDA_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        device = e ? atoi(e) : 0;
    }
    {
        std::lock_guard<std::mutex> lk(g_init_mu);
        if (g_ndev > 0) {
            if (g_ctx[0].device == device) return LUTGEN_OK;
            return fail(LUTGEN_ERR_INVALID, "already initialised with device %d as the primary one", g_ctx[0].device);
        }
    }
    return lutgen_cuda_init_devices(&device, 1);
}

int lutgen_cuda_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_ndev == 0) return LUTGEN_OK;
    workers_stop();
    {
        UseCtx u(0);
        png_release();
    }
    for (int i = 0; i < g_ndev; ++i) {
        Ctx* prev = tl_ctx;
        tl_ctx = &g_ctx[i];
        release_current_ctx();
        tl_ctx = prev;
    }
    g_ndev = 0;
    return LUTGEN_OK;
}

int lutgen_cuda_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(LUTGEN_ERR_INVALID, "out is NULL");
    int rc = ensure_ready();
    if (rc) return rc;
    CU_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return LUTGEN_OK;
}
int lutgen_cuda_host_free(void* p) {
    if (!p) return LUTGEN_OK;
    int rc = ensure_ready();
    if (rc) return rc;
    CU_TRY(cudaFreeHost(p));
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// identity.rs
// ---------------------------------------------------------------------------
int lutgen_cuda_detect_level(uint32_t width, uint32_t height, uint8_t* level) {
    if (!level) return fail(LUTGEN_ERR_INVALID, "level is NULL");
    uint64_t l = 2;
    while (l * l * l < width) ++l;
    if (width != l * l * l) return fail(LUTGEN_ERR_INVALID, "hald clut width %u is not a cube (assert_eq!(width, level^3), identity.rs:95)", width);
    if (width != height) return fail(LUTGEN_ERR_INVALID, "hald clut is not square: %u x %u (identity.rs:96)", width, height);
    *level = (uint8_t)l;
    return LUTGEN_OK;
}

int lutgen_cuda_identity_dev(uint8_t level, uint8_t* out_rgb_dev, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33 (the reference divides by zero / overflows)", level);
    if (!out_rgb_dev) return fail(LUTGEN_ERR_INVALID, "out_rgb_dev is NULL");
    cudaStream_t s = (cudaStream_t)stream;  // 0 is CUDA's default stream, as everywhere in CUDA
    size_t E = lut_entries(level);
    uint32_t cs = (uint32_t)level * level;
    k_identity<<<blocks_for((E + 3) / 4, 256), 256, 0, s>>>(out_rgb_dev, 0, E, cs, cs - 1);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

int lutgen_cuda_identity(uint8_t level, uint8_t* out_rgb) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33 (the reference divides by zero / overflows)", level);
    if (!out_rgb) return fail(LUTGEN_ERR_INVALID, "out_rgb is NULL");
    std::lock_guard<std::mutex> lk(g.mu);
    size_t bytes = 3 * lut_entries(level);
    CU_TRY(g.lut_rgb.reserve(bytes));
    rc = lutgen_cuda_identity_dev(level, (uint8_t*)g.lut_rgb.p, g.stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out_rgb, g.lut_rgb.p, bytes, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int lutgen_cuda_apply_hald_dev(uint8_t* rgba_dev, size_t npix, const uint8_t* lut_rgb_dev, uint8_t level, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (!lut_rgb_dev || (!rgba_dev && npix)) return fail(LUTGEN_ERR_INVALID, "NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;  // 0 is CUDA's default stream, as everywhere in CUDA
    if (npix) {
        rc = expand_lut(lut_rgb_dev, level, s);
        if (rc) return rc;
        rc = launch_apply(rgba_dev, npix, (const uint32_t*)g.lut_rgbx.p, level, s);
        if (rc) return rc;
    }
    return LUTGEN_OK;
}

namespace {
int apply_batch_share(uint8_t* const* frames, const size_t* npix, size_t nframes, const uint8_t* lut_rgb, uint8_t level, size_t first, size_t stride);
}
int lutgen_cuda_apply_hald(uint8_t* rgba, size_t npix, const uint8_t* lut_rgb, uint8_t level) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (!lut_rgb || (!rgba && npix)) return fail(LUTGEN_ERR_INVALID, "NULL buffer");
    if (npix == 0) return LUTGEN_OK;
    // one frame through the chunked upload / kernel / download pipeline of the batch entry point
    return apply_batch_share(&rgba, &npix, 1, lut_rgb, level, 0, 1);
}

namespace {
// Frames first, first + stride, ... on the current context: CLUT up once, then a pipeline over kBatchStreams streams in
// CHUNKS of 2 M or 4 M pixels, whatever the frame sizes: chunk i's upload overlaps chunk i-1's kernel and chunk i-2's
// download (separate copy engines per direction). With whole frames as the unit (round 1) a single frame's upload and
// download did not overlap at all. LUTGEN_CUDA_APPLY_CHUNK=<pixels> overrides the chunk size (0: whole frames).
int apply_batch_share(uint8_t* const* frames, const size_t* npix, size_t nframes, const uint8_t* lut_rgb, uint8_t level, size_t first, size_t stride) {
    int rc = ensure_ready();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t lut_bytes = 3 * lut_entries(level), max_px = 0;
    for (size_t i = first; i < nframes; i += stride)
        if (npix[i] > max_px) max_px = npix[i];
    if (max_px == 0) return LUTGEN_OK;
    size_t total_px = 0;
    for (size_t i = first; i < nframes; i += stride) total_px += npix[i];
    // measured (tools/time_apply_e2e.py, pinned 4K frames): one frame 1.53 ms whole, 1.12 ms in 2 M-pixel chunks; four
    // frames 3.68 ms whole, 3.63 / 3.50 ms in chunks of 2 M / 4 M pixels, 4.35 / 5.40 ms in chunks of 1 M / 0.5 M (a copy
    // costs ~10 us to set up, and the two directions share the link: 37 GB/s each way when both run)
    size_t chunk = total_px >= ((size_t)1 << 24) ? (size_t)1 << 22 : (size_t)1 << 21;
    if (const char* e = getenv("LUTGEN_CUDA_APPLY_CHUNK")) chunk = (size_t)strtoull(e, nullptr, 10);
    if (chunk == 0 || chunk > max_px) chunk = max_px;
    chunk = (chunk + 3) & ~(size_t)3;   // k_apply takes four pixels per thread: chunk starts stay 16-byte aligned in the frame
    CU_TRY(g.lut_rgb.reserve(lut_bytes));
    CU_TRY(cudaMemcpyAsync(g.lut_rgb.p, lut_rgb, lut_bytes, cudaMemcpyHostToDevice, g.stream));
    rc = expand_lut((const uint8_t*)g.lut_rgb.p, level, g.stream);
    if (rc) return rc;
    cudaEvent_t lut_ready;
    CU_TRY(cudaEventCreateWithFlags(&lut_ready, cudaEventDisableTiming));
    int result = LUTGEN_OK;
    cudaError_t e = cudaEventRecord(lut_ready, g.stream);
    for (int i = 0; i < kBatchStreams && e == cudaSuccess; ++i) {
        e = g.batch_frames[i].reserve(chunk * 4);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(g.batch_streams[i], lut_ready, 0);
    }
    size_t k = 0;
    for (size_t i = first; i < nframes && e == cudaSuccess && result == LUTGEN_OK; i += stride) {
        for (size_t p0 = 0; p0 < npix[i] && e == cudaSuccess && result == LUTGEN_OK; p0 += chunk) {
            const size_t n = std::min(chunk, npix[i] - p0);
            int b = (int)(k++ % kBatchStreams);
            cudaStream_t s = g.batch_streams[b];
            uint8_t* dev = (uint8_t*)g.batch_frames[b].p;
            uint8_t* host = frames[i] + 4 * p0;
            e = cudaMemcpyAsync(dev, host, n * 4, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) break;
            result = launch_apply(dev, n, (const uint32_t*)g.lut_rgbx.p, level, s);
            if (result != LUTGEN_OK) break;
            e = cudaMemcpyAsync(host, dev, n * 4, cudaMemcpyDeviceToHost, s);
        }
    }
    // whatever happened, nothing may still be copying from or into the caller's frames when this returns
    for (int i = 0; i < kBatchStreams; ++i) {
        cudaError_t es = cudaStreamSynchronize(g.batch_streams[i]);
        if (e == cudaSuccess) e = es;
    }
    cudaEventDestroy(lut_ready);
    if (result != LUTGEN_OK) return result;
    if (e != cudaSuccess) return fail(LUTGEN_ERR_CUDA, "apply_hald_batch: %s", cudaGetErrorString(e));
    return LUTGEN_OK;
}
}  // namespace

int lutgen_cuda_apply_hald_batch(uint8_t* const* frames, const size_t* npix, size_t nframes, const uint8_t* lut_rgb, uint8_t level) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (!lut_rgb || (nframes && (!frames || !npix))) return fail(LUTGEN_ERR_INVALID, "NULL buffer");
    if (nframes == 0) return LUTGEN_OK;
    for (size_t i = 0; i < nframes; ++i)
        if (npix[i] && !frames[i]) return fail(LUTGEN_ERR_INVALID, "frame %zu is NULL", i);
    // several GPUs in this process: frames are dealt out round robin, every GPU runs its own pipeline (no exchange)
    const int ndev = tl_ctx == &g_ctx[0] ? (int)std::min<size_t>((size_t)g_ndev, nframes) : 1;
    for (int i = 1; i < ndev; ++i) worker_submit(i, [=] { return apply_batch_share(frames, npix, nframes, lut_rgb, level, (size_t)i, (size_t)ndev); });
    int result = apply_batch_share(frames, npix, nframes, lut_rgb, level, 0, (size_t)(ndev > 1 ? ndev : 1));
    std::string first_err = result ? g_err : std::string();
    for (int i = 1; i < ndev; ++i) {
        int wrc = worker_wait(i);
        if (wrc && result == LUTGEN_OK) { result = wrc; first_err = g_err; }
    }
    if (result) g_err = first_err;
    return result;
}

int lutgen_cuda_correct_pixels(const uint8_t* in_rgb, size_t n, const uint8_t* lut_rgb, uint8_t level, uint8_t* out_rgb) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (!lut_rgb || (n && (!in_rgb || !out_rgb))) return fail(LUTGEN_ERR_INVALID, "NULL buffer");
    if (n == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t lut_bytes = 3 * lut_entries(level);
    uint32_t cs = (uint32_t)level * level;
    CU_TRY(g.lut_rgb.reserve(lut_bytes));
    CU_TRY(g.misc.reserve(6 * n));
    uint8_t* din = (uint8_t*)g.misc.p;
    uint8_t* dout = din + 3 * n;
    CU_TRY(cudaMemcpyAsync(g.lut_rgb.p, lut_rgb, lut_bytes, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(cudaMemcpyAsync(din, in_rgb, 3 * n, cudaMemcpyHostToDevice, g.stream));
    rc = expand_lut((const uint8_t*)g.lut_rgb.p, level, g.stream);
    if (rc) return rc;
    k_correct_pixels<<<blocks_for(n, 256), 256, 0, g.stream>>>(din, n, (const uint32_t*)g.lut_rgbx.p, cs, cs - 1, dout);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out_rgb, dout, 3 * n, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// remappers
// ---------------------------------------------------------------------------
int lutgen_cuda_remapper_destroy(lutgen_remapper* r) {
    if (!r) return LUTGEN_OK;
    for (int i = 1; i < kMaxDevices; ++i) {
        if (r->replica[i]) lutgen_cuda_remapper_destroy(r->replica[i]);
        r->replica[i] = nullptr;
    }
    UseCtx use(r->ctx_index);
    if (g.ready) {
        cudaDeviceSynchronize();
        if (r->arena) cudaFreeAsync(r->arena, g.stream);   // back to the pool: creating the next remapper costs no cudaMalloc
        cudaFree(r->offsets);
        cudaFree(r->nn_table);
        cudaFree(r->flag_list);
        cudaFree(r->flag_count);
        cudaFree(r->full_table);
        if (r->nn_table_ev) cudaEventDestroy(r->nn_table_ev);
        if (r->full_table_ev) cudaEventDestroy(r->full_table_ev);
        cudaFree(r->blur_pal3);
        cudaFree(r->blur_out3);
        cudaFree(r->blur_kernel);
        cudaFree(r->blur_vol[0]);
        cudaFree(r->blur_vol[1]);
        cudaFree(r->blur_chan);
        cudaFree(r->blur_ends);
        if (r->blur_ev) cudaEventDestroy(r->blur_ev);
    }
    delete r;
    return LUTGEN_OK;
}

int lutgen_cuda_remapper_create(const lutgen_remapper_desc* desc, lutgen_remapper** out) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!desc || !out) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (desc->kind < LUTGEN_GAUSSIAN || desc->kind > LUTGEN_BLUR) return fail(LUTGEN_ERR_INVALID, "unknown remapper kind %d", desc->kind);
    if (desc->kind == LUTGEN_BLUR) {
        if (desc->palette_len == 0) return fail(LUTGEN_ERR_INVALID, "empty palette (palette_oklab[hint] is out of bounds, gaussian_blur.rs:58)");
        float r3 = ceilf((float)desc->param * 3.0f);
        if (r3 > 4096.0f) return fail(LUTGEN_ERR_UNSUPPORTED, "blur radius %g needs more than 8193 taps", desc->param);
    }
    if (desc->palette_len && !desc->palette) return fail(LUTGEN_ERR_INVALID, "palette is NULL");
    if (desc->palette_len > (1u << 20)) return fail(LUTGEN_ERR_UNSUPPORTED, "palette of %llu colours (limit 2^20)", (unsigned long long)desc->palette_len);
    if (desc->kind == LUTGEN_SAMPLING) {
        // rand_distr::Normal::new(mean, std_dev).unwrap() (gaussian_sample.rs:36) rejects a non-finite std_dev
        if (!std::isfinite(desc->std_dev)) return fail(LUTGEN_ERR_INVALID, "std_dev must be finite (Normal::new(..).unwrap() panics)");
        if (desc->iterations > (1ull << 31)) return fail(LUTGEN_ERR_UNSUPPORTED, "iterations > 2^31");
    }
    if ((desc->kind == LUTGEN_NEAREST || desc->kind == LUTGEN_SAMPLING) && desc->palette_len == 0)
        return fail(LUTGEN_ERR_INVALID, "empty palette (kiddo nearest_one on an empty tree panics)");
    std::lock_guard<std::mutex> lk(g.mu);
    lutgen_remapper* r = new lutgen_remapper();
    r->ctx_index = g.index;
    r->d = *desc;
    r->n = (uint32_t)desc->palette_len;
    r->pal_host.assign(desc->palette, desc->palette + 3 * (size_t)r->n);
    r->d.palette = r->pal_host.data();
    r->k_eff = (desc->nearest != 0 && desc->nearest < desc->palette_len) ? (uint32_t)desc->nearest : 0u;  // rbf.rs:38
    auto bail = [&](int code) {
        lutgen_cuda_remapper_destroy(r);
        return code;
    };
#define CU_TRY_R(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) return bail(fail(LUTGEN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e))); \
    } while (0)
    size_t n = r->n, na = n ? n : 1;
    std::vector<uint8_t> rgb4(4 * na, 0);
    for (size_t i = 0; i < n; ++i) {
        rgb4[4 * i] = r->pal_host[3 * i];
        rgb4[4 * i + 1] = r->pal_host[3 * i + 1];
        rgb4[4 * i + 2] = r->pal_host[3 * i + 2];
    }
    // one allocation from the device's stream-ordered pool (kept warm: lutgen_cuda_init sets its release threshold),
    // carved into the palette arrays -- and a second set for the de-duplicated palette of the NearestNeighbor kinds.
    // The CLI builds a remapper per invocation (crates/cli/src/main.rs:334-348): construction is on the end-to-end path.
    {
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t per_set = up(4 * na) + up(sizeof(double) * 3 * na) + up(sizeof(float) * 2 * na) + up(sizeof(float4) * na);
        const bool two = desc->kind == LUTGEN_NEAREST || desc->kind == LUTGEN_SAMPLING;
        CU_TRY_R(cudaMallocAsync(&r->arena, per_set * (two ? 2 : 1), g.stream));
        uint8_t* q = (uint8_t*)r->arena;
        r->pal_rgb4 = q; q += up(4 * na);
        r->pal_lab = (double*)q; q += up(sizeof(double) * 3 * na);
        r->pal_ab = (float*)q; q += up(sizeof(float) * 2 * na);
        r->pal_lab32 = (float4*)q; q += up(sizeof(float4) * na);
        if (two) {
            r->upal_rgb4 = q; q += up(4 * na);
            r->upal_lab = (double*)q; q += up(sizeof(double) * 3 * na);
            r->upal_ab = (float*)q; q += up(sizeof(float) * 2 * na);
            r->upal_lab32 = (float4*)q;
        }
    }
    CU_TRY_R(cudaMemcpyAsync(r->pal_rgb4, rgb4.data(), 4 * na, cudaMemcpyHostToDevice, g.stream));
    if (n) {
        k_palette_setup<<<blocks_for(n, 128), 128, 0, g.stream>>>(r->pal_rgb4, r->n, desc->lum_factor, g.tables->decode, r->pal_lab, r->pal_ab,
                                                                  r->pal_lab32);
        g_launches.fetch_add(1);
        CU_TRY_R(cudaGetLastError());
    }
    if (desc->kind == LUTGEN_SAMPLING) {
        size_t K = (size_t)desc->iterations, Ka = K ? K : 1;
        r->noise.assign(4 * K, 0.0);
        CU_TRY_R(cudaMalloc((void**)&r->offsets, sizeof(short4) * Ka * 2));  // iteration order + sorted copy
        if (K) {
            CU_TRY_R(g.misc.reserve(sizeof(double) * 4 * K));
            k_philox_normals<<<blocks_for(2 * K, 128), 128, 0, g.stream>>>((double*)g.misc.p, K, desc->mean, desc->std_dev, desc->seed);
            g_launches.fetch_add(1);
            CU_TRY_R(cudaGetLastError());
            CU_TRY_R(cudaMemcpyAsync(r->noise.data(), g.misc.p, sizeof(double) * 4 * K, cudaMemcpyDeviceToHost, g.stream));
            CU_TRY_R(cudaStreamSynchronize(g.stream));
            int rc2 = upload_offsets(r);
            if (rc2) return bail(rc2);
        }
    }
    if (desc->kind == LUTGEN_BLUR) {
        // GaussianBlurRemapper::new + build_kernel (gaussian_blur.rs:34-53, 87-96), all f32. expf is
        // the host libm's, the same one a CPU build of the reference would call.
        const float lum = (float)desc->lum_factor, radius = (float)desc->param;
        CU_TRY_R(cudaMalloc((void**)&r->blur_pal3, sizeof(float) * 3 * na));
        CU_TRY_R(cudaMalloc((void**)&r->blur_out3, sizeof(float) * 3 * na));
        blur_launch_palette(blocks_for(n, 128), g.stream, r->pal_rgb4, r->n, lum, g.tables->decode, r->blur_pal3, r->blur_out3);
        g_launches.fetch_add(1);
        CU_TRY_R(cudaGetLastError());
        float h3 = ceilf(radius * 3.0f);
        int half = (h3 != h3) ? 0 : (h3 < -1e6f ? -1000000 : (int)h3);  // `as i32`: NaN -> 0
        const float two_sigma_sq = 2.0f * radius * radius;
        std::vector<float> kw;
        for (int i = -half; i <= half; ++i) kw.push_back(expf((float)(-(i * i)) / two_sigma_sq));
        float sum = 0.0f;
        for (float w : kw) sum += w;
        for (float& w : kw) w /= sum;
        r->blur_klen = (int)kw.size();
        CU_TRY_R(cudaMalloc((void**)&r->blur_kernel, sizeof(float) * (kw.size() + 1)));
        if (!kw.empty()) CU_TRY_R(cudaMemcpyAsync(r->blur_kernel, kw.data(), sizeof(float) * kw.size(), cudaMemcpyHostToDevice, g.stream));
        CU_TRY_R(cudaStreamSynchronize(g.stream));
    }
    if (desc->kind == LUTGEN_NEAREST || desc->kind == LUTGEN_SAMPLING) {
        // unique colours, first occurrence kept (see lutgen_remapper::un)
        std::vector<uint8_t> u4;
        std::vector<uint32_t> seen;
        for (size_t i = 0; i < n; ++i) {
            uint32_t key = rgb4[4 * i] | (rgb4[4 * i + 1] << 8) | (rgb4[4 * i + 2] << 16);
            bool dup = false;
            for (uint32_t s2 : seen) dup |= (s2 == key);
            if (dup) continue;
            seen.push_back(key);
            u4.insert(u4.end(), {rgb4[4 * i], rgb4[4 * i + 1], rgb4[4 * i + 2], 0});
        }
        r->un = (uint32_t)seen.size();
        if (r->un) {
            CU_TRY_R(cudaMemcpyAsync(r->upal_rgb4, u4.data(), 4 * (size_t)r->un, cudaMemcpyHostToDevice, g.stream));
            k_palette_setup<<<blocks_for(r->un, 128), 128, 0, g.stream>>>(r->upal_rgb4, r->un, desc->lum_factor, g.tables->decode, r->upal_lab,
                                                                          r->upal_ab, r->upal_lab32);
            g_launches.fetch_add(1);
            CU_TRY_R(cudaGetLastError());
        }
    }
    CU_TRY_R(cudaStreamSynchronize(g.stream));
    r->fast_ok = fast_path_supported(r);
#undef CU_TRY_R
    *out = r;
    return LUTGEN_OK;
}

int lutgen_cuda_sampling_get_noise(const lutgen_remapper* r, double* noise_out) {
    if (!r || !noise_out) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (r->d.kind != LUTGEN_SAMPLING) return fail(LUTGEN_ERR_INVALID, "not a sampling remapper");
    memcpy(noise_out, r->noise.data(), sizeof(double) * r->noise.size());
    return LUTGEN_OK;
}

int lutgen_cuda_sampling_set_noise(lutgen_remapper* r, const double* noise) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !noise) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (r->d.kind != LUTGEN_SAMPLING) return fail(LUTGEN_ERR_INVALID, "not a sampling remapper");
    std::lock_guard<std::mutex> lk(g.mu);
    for (int i = 1; i < kMaxDevices; ++i) {   // the other GPUs' copies are rebuilt, with this table, on next use
        if (r->replica[i]) lutgen_cuda_remapper_destroy(r->replica[i]);
        r->replica[i] = nullptr;
    }
    std::vector<double> old = r->noise;
    r->noise.assign(noise, noise + r->noise.size());
    {
        std::lock_guard<std::mutex> lk2(r->mu);
        r->full_table_ready = false;  // the tabulated remapper depends on the noise
    }
    CU_TRY(cudaDeviceSynchronize());
    rc = upload_offsets(r);
    if (rc) {
        r->noise = old;
        upload_offsets(r);
    }
    return rc;
}

namespace {
int generate_lut_rows_impl(const lutgen_remapper* r, uint8_t level, uint8_t* lut_rgb_dev, uint32_t row_begin, uint32_t row_end, void* stream,
                           int sequential);
}

int lutgen_cuda_generate_lut_rows_dev(const lutgen_remapper* r, uint8_t level, uint8_t* lut_rgb_dev, uint32_t row_begin, uint32_t row_end,
                                      void* stream) {
    return generate_lut_rows_impl(r, level, lut_rgb_dev, row_begin, row_end, stream, 0);
}

namespace {
int generate_lut_rows_impl(const lutgen_remapper* r, uint8_t level, uint8_t* lut_rgb_dev, uint32_t row_begin, uint32_t row_end, void* stream,
                           int sequential) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !lut_rgb_dev) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    uint32_t side = (uint32_t)level * level * level;
    if (row_begin > row_end || row_end > side) return fail(LUTGEN_ERR_INVALID, "rows [%u, %u) outside 0..%u", row_begin, row_end, side);
    cudaStream_t s = (cudaStream_t)stream;  // 0 is CUDA's default stream, as everywhere in CUDA
    Front f{};
    f.out = lut_rgb_dev;
    f.begin = (unsigned long long)row_begin * side;
    f.count = (unsigned long long)(row_end - row_begin) * side;
    f.cs = (uint32_t)level * level;
    f.csm1 = f.cs - 1;
    f.sequential = sequential;
    rc = launch_remap<FRONT_LUT>(f, r, s);
    if (rc) return rc;
    return LUTGEN_OK;
}

int generate_lut_impl(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag, int sequential);
}  // namespace

int lutgen_cuda_generate_lut(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag) {
    return generate_lut_impl(r, level, out_rgb, abort_flag, 0);
}

int lutgen_cuda_generate_lut_sequential(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag) {
    return generate_lut_impl(r, level, out_rgb, abort_flag, 1);
}

namespace {
int generate_lut_share(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag, int sequential, uint32_t chunks,
                       uint32_t first, uint32_t stride);
}

int lutgen_cuda_generate_lut_part(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, uint32_t part_index, uint32_t part_count,
                                  const volatile uint8_t* abort_flag) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !out_rgb) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (part_count == 0 || part_index >= part_count) return fail(LUTGEN_ERR_INVALID, "part %u of %u", part_index, part_count);
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    const uint32_t side = (uint32_t)level * level * level, ngroups = (side + 16u * level - 1) / (16u * level);
    // the same dealing-out as lutgen_cuda_generate_lut does over the GPUs of one process
    uint32_t chunks = std::min<uint32_t>(ngroups, part_count * (ngroups >= 2u * part_count ? 2u : 1u));
    if (abort_flag) chunks = std::min<uint32_t>(ngroups, std::max<uint32_t>(chunks, 8u));
    if (part_index >= chunks) return LUTGEN_OK;   // more parts than row groups: nothing for this one
    return generate_lut_share(r, level, out_rgb, abort_flag, 0, chunks, part_index, part_count);
}

int lutgen_cuda_generate_lut_rows_multi_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint32_t row_begin,
                                            uint32_t row_end, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !luts_rgb_dev || nluts < 1) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (nluts > 1 + LUTGEN_MAX_PEERS) return fail(LUTGEN_ERR_UNSUPPORTED, "%d destinations (limit %d)", nluts, 1 + LUTGEN_MAX_PEERS);
    for (int i = 0; i < nluts; ++i)
        if (!luts_rgb_dev[i]) return fail(LUTGEN_ERR_INVALID, "destination %d is NULL", i);
    if (r->d.kind == LUTGEN_SAMPLING || r->d.kind == LUTGEN_BLUR)
        return fail(LUTGEN_ERR_UNSUPPORTED, "multi-destination rows: RBF and NearestNeighbor remappers only (others: generate locally, then all-gather)");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    uint32_t side = (uint32_t)level * level * level;
    if (row_begin > row_end || row_end > side) return fail(LUTGEN_ERR_INVALID, "rows [%u, %u) outside 0..%u", row_begin, row_end, side);
    cudaStream_t s = (cudaStream_t)stream;
    Front f{};
    f.out = luts_rgb_dev[0];
    f.nextra = nluts - 1;
    for (int i = 1; i < nluts; ++i) f.extra[i - 1] = luts_rgb_dev[i];
    f.begin = (unsigned long long)row_begin * side;
    f.count = (unsigned long long)(row_end - row_begin) * side;
    f.cs = (uint32_t)level * level;
    f.csm1 = f.cs - 1;
    return launch_remap<FRONT_LUT>(f, r, s);
}

int lutgen_cuda_generate_lut_shard_multi_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint32_t shard_index,
                                             uint32_t shard_count, void* stream) {
    return lutgen_cuda_generate_lut_shard_mc_dev(r, level, luts_rgb_dev, nluts, nullptr, shard_index, shard_count, stream);
}

int lutgen_cuda_generate_lut_shard_mc_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint8_t* multicast_lut,
                                          uint32_t shard_index, uint32_t shard_count, void* stream) {
    return lutgen_cuda_generate_lut_exchange_dev(r, level, luts_rgb_dev, nluts, multicast_lut, shard_index, shard_count, nullptr, stream);
}

int lutgen_cuda_generate_lut_exchange_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint8_t* multicast_lut,
                                          uint32_t shard_index, uint32_t shard_count, unsigned* const* barrier_flags_dev_table, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !luts_rgb_dev || nluts < 1) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (nluts > 1 + LUTGEN_MAX_PEERS) return fail(LUTGEN_ERR_UNSUPPORTED, "%d destinations (limit %d)", nluts, 1 + LUTGEN_MAX_PEERS);
    for (int i = 0; i < nluts; ++i)
        if (!luts_rgb_dev[i]) return fail(LUTGEN_ERR_INVALID, "destination %d is NULL", i);
    if (shard_count == 0 || shard_index >= shard_count) return fail(LUTGEN_ERR_INVALID, "shard %u of %u", shard_index, shard_count);
    if (r->d.kind == LUTGEN_SAMPLING || r->d.kind == LUTGEN_BLUR)
        return fail(LUTGEN_ERR_UNSUPPORTED, "multi-destination shards: RBF and NearestNeighbor remappers only (others: generate locally, then all-gather)");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    uint32_t side = (uint32_t)level * level * level;
    Front f{};
    f.out = luts_rgb_dev[0];
    f.nextra = nluts - 1;
    for (int i = 1; i < nluts; ++i) f.extra[i - 1] = luts_rgb_dev[i];
    f.begin = 0;
    f.count = (unsigned long long)side * side;
    f.cs = (uint32_t)level * level;
    f.csm1 = f.cs - 1;
    f.shard_index = shard_index;
    f.shard_count = shard_count;
    f.mc = nluts > 1 ? multicast_lut : nullptr;
    if (barrier_flags_dev_table) {
        // only k_remap_warp's MULTI variant ends with the barrier: small palettes of the RBF / NearestNeighbor kinds
        const bool nn = r->d.kind == LUTGEN_NEAREST;
        const uint32_t n = nn ? r->un : r->n;
        if (nluts < 2 || shard_count != (uint32_t)nluts || !r->fast_ok || n == 0 || n > (uint32_t)WT_MAX_N || getenv("LUTGEN_CUDA_NO_WARP"))
            return fail(LUTGEN_ERR_UNSUPPORTED, "in-kernel exchange barrier: needs one shard per destination and a palette of at most %d colours on the warp-tile kernel", WT_MAX_N);
        if (r->d.kind == LUTGEN_GAUSSIAN) {
            double C = r->d.param * 1.4426950408889634;
            if (!(C >= 1e-3 && C <= 2000.0)) return fail(LUTGEN_ERR_UNSUPPORTED, "in-kernel exchange barrier: Gaussian shape outside the warp-tile kernel's range");
        }
        f.bar_flags = barrier_flags_dev_table;
        f.bar_nranks = nluts;
        f.bar_rank = (int)shard_index;
    }
    return launch_remap<FRONT_LUT>(f, r, (cudaStream_t)stream);
}

// Barrier between the per-GPU processes of one node through flags in peer memory: rank r writes
// `epoch` into slot r of every rank's flag array (release, system scope: everything this GPU wrote
// before -- its rows in the peers' CLUT buffers -- is visible to whoever sees the flag) and waits
// until all slots of its own array have reached `epoch`. A few microseconds, where an NCCL
// barrier is a 40 us all-reduce.
__global__ void k_peer_barrier(unsigned* const* flags, int nranks, int rank, unsigned epoch) {
    const int t = threadIdx.x;
    if (t >= nranks) return;
    __threadfence_system();
    unsigned* remote = flags[t] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned* mine = flags[rank] + t;
    unsigned v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while ((int)(v - epoch) < 0);
}

int lutgen_cuda_peer_barrier_dev(unsigned* const* flags_dev_table, int nranks, int rank, unsigned epoch, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!flags_dev_table || nranks < 1 || nranks > 1 + LUTGEN_MAX_PEERS || rank < 0 || rank >= nranks)
        return fail(LUTGEN_ERR_INVALID, "peer barrier: %d ranks, rank %d", nranks, rank);
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(flags_dev_table, nranks, rank, epoch);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

// CUDA IPC: one process per GPU; a rank exports its CLUT buffer and maps the other ranks'.
int lutgen_cuda_device_alloc(size_t bytes, void** out) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    CU_TRY(cudaMalloc(out, bytes ? bytes : 1));
    return LUTGEN_OK;
}

int lutgen_cuda_device_free(void* p) {
    int rc = ensure_ready();
    if (rc) return rc;
    CU_TRY(cudaFree(p));
    return LUTGEN_OK;
}

int lutgen_cuda_ipc_export(void* dev_ptr, unsigned char handle[64]) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!dev_ptr || !handle) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return LUTGEN_OK;
}

int lutgen_cuda_ipc_import(const unsigned char handle[64], void** dev_ptr_out) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!handle || !dev_ptr_out) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU_TRY(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return LUTGEN_OK;
}

int lutgen_cuda_ipc_close(void* dev_ptr) {
    int rc = ensure_ready();
    if (rc) return rc;
    CU_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return LUTGEN_OK;
}

namespace {
// The current context's share of a CLUT into the caller's host image: chunks `first`, first + `stride`, ... of `chunks`
// equal runs of row groups (the whole CLUT for first = 0, stride = 1). `r` lives on the current context.
int generate_lut_share(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag, int sequential, uint32_t chunks,
                       uint32_t first, uint32_t stride) {
    int rc = ensure_ready();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t bytes = 3 * lut_entries(level);
    uint32_t side = (uint32_t)level * level * level;
    CU_TRY(g.lut_rgb.reserve(bytes));
    // Row chunks: (1) a set abort flag is honoured between launches (lib.rs:149-170); (2) the
    // download of chunk i rides a second stream while chunk i+1 is being computed, so a large
    // CLUT costs about max(kernel, PCIe copy) instead of their sum. Chunk edges sit on multiples
    // of 16 blue planes (16*level rows), the kernels' brick depth, so no brick is visited twice.
    const uint32_t group = 16u * level;
    uint32_t ngroups = (side + group - 1) / group;
    cudaStream_t copy_stream = g.batch_streams[0];
    cudaEvent_t done[64];
    if (chunks > 64u) chunks = 64u;
    for (uint32_t c = 0; c < chunks; ++c) CU_TRY(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
    int result = LUTGEN_OK;
    for (uint32_t c = first; c < chunks && result == LUTGEN_OK; c += stride) {
        uint32_t r0 = (uint32_t)((uint64_t)ngroups * c / chunks) * group, r1 = (uint32_t)((uint64_t)ngroups * (c + 1) / chunks) * group;
        if (r1 > side) r1 = side;
        rc = generate_lut_rows_impl(r, level, (uint8_t*)g.lut_rgb.p, r0, r1, g.stream, sequential);
        if (rc) { result = rc; break; }
        cudaError_t e = cudaEventRecord(done[c], g.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_stream, done[c], 0);
        size_t off = (size_t)r0 * side * 3, len = (size_t)(r1 - r0) * side * 3;
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgb + off, (uint8_t*)g.lut_rgb.p + off, len, cudaMemcpyDeviceToHost, copy_stream);
        if (e != cudaSuccess) { result = fail(LUTGEN_ERR_CUDA, "chunk %u: %s", c, cudaGetErrorString(e)); break; }
        if (abort_flag) {
            cudaStreamSynchronize(g.stream);
            if (*abort_flag) result = LUTGEN_ABORTED;
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(g.stream), e2 = cudaStreamSynchronize(copy_stream);
    for (uint32_t c = 0; c < chunks; ++c) cudaEventDestroy(done[c]);
    if (result != LUTGEN_OK) return result;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(LUTGEN_ERR_CUDA, "generate_lut: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    return LUTGEN_OK;
}

// The same remapper on context `index` (made on first use; r->mu held by the caller).
int ensure_replica(const lutgen_remapper* r, int index) {
    if (r->replica[index]) return LUTGEN_OK;
    UseCtx use(index);
    lutgen_remapper* rep = nullptr;
    lutgen_remapper_desc d = r->d;
    d.palette = r->pal_host.data();
    int rc = lutgen_cuda_remapper_create(&d, &rep);
    if (rc) return rc;
    if (r->d.kind == LUTGEN_SAMPLING && !r->noise.empty()) {   // the same noise table, whatever the caller has set
        rc = lutgen_cuda_sampling_set_noise(rep, r->noise.data());
        if (rc) {
            lutgen_cuda_remapper_destroy(rep);
            return rc;
        }
    }
    r->replica[index] = rep;
    return LUTGEN_OK;
}

int generate_lut_impl(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, const volatile uint8_t* abort_flag, int sequential) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !out_rgb) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    const size_t bytes = 3 * lut_entries(level);
    const uint32_t side = (uint32_t)level * level * level, ngroups = (side + 16u * level - 1) / (16u * level);
    // Several GPUs in this process (lutgen_cuda_init_devices / LUTGEN_CUDA_DEVICES): the row groups are dealt out to
    // them round robin (dark and light ends of the cube cost differently), every GPU computes its groups with its
    // own copy of the remapper and copies them straight into the caller's image over its own PCIe link -- no exchange
    // between the GPUs is needed for a CLUT that is wanted on the host. GaussianBlurRemapper stays on one GPU: its
    // cost is the blurred volume, which every GPU would have to compute in full.
    const int ndev = (tl_ctx == &g_ctx[0] && r->ctx_index == 0 && r->d.kind != LUTGEN_BLUR && bytes >= (4u << 20)) ? std::min<int>(g_ndev, (int)ngroups) : 1;
    if (ndev > 1) {
        {
            std::lock_guard<std::mutex> lk(r->mu);
            for (int i = 1; i < ndev; ++i) {
                rc = ensure_replica(r, i);
                if (rc) return rc;
            }
        }
        // at least two chunks per GPU so that a chunk's download overlaps the next one's kernel
        uint32_t chunks = std::min<uint32_t>(ngroups, (uint32_t)ndev * (ngroups >= 2u * ndev ? 2u : 1u));
        if (abort_flag) chunks = std::min<uint32_t>(ngroups, std::max<uint32_t>(chunks, 8u));
        for (int i = 1; i < ndev; ++i) {
            const lutgen_remapper* rep = r->replica[i];
            worker_submit(i, [=] { return generate_lut_share(rep, level, out_rgb, abort_flag, sequential, chunks, (uint32_t)i, (uint32_t)ndev); });
        }
        int result = generate_lut_share(r, level, out_rgb, abort_flag, sequential, chunks, 0u, (uint32_t)ndev);
        std::string first_err = result ? g_err : std::string();
        for (int i = 1; i < ndev; ++i) {
            int wrc = worker_wait(i);
            if (wrc && (result == LUTGEN_OK || result == LUTGEN_ABORTED) && wrc != LUTGEN_ABORTED) { result = wrc; first_err = g_err; }
            else if (wrc == LUTGEN_ABORTED && result == LUTGEN_OK) result = wrc;
        }
        if (result && result != LUTGEN_ABORTED) g_err = first_err;
        return result;
    }
    uint32_t chunks = (bytes >= (8u << 20)) ? 4u : 1u;
    if (abort_flag && chunks < 8u) chunks = 8u;
    if (chunks > ngroups) chunks = ngroups;
    return generate_lut_share(r, level, out_rgb, abort_flag, sequential, chunks, 0u, 1u);
}
}  // namespace

// ---------------------------------------------------------------------------
// PNG writer (png.cu)
// ---------------------------------------------------------------------------
namespace {
int png_args_ok(uint32_t width, uint32_t height, int channels) {
    if (channels < 1 || channels > 4) return fail(LUTGEN_ERR_INVALID, "%d channels (1..4)", channels);
    if (width == 0 || height == 0) return fail(LUTGEN_ERR_INVALID, "empty image (%u x %u): PNG has no zero-sized images", width, height);
    if ((unsigned long long)width * channels >= (1ull << 31) || (unsigned long long)height * ((unsigned long long)width * channels + 1ull) >= (1ull << 40))
        return fail(LUTGEN_ERR_UNSUPPORTED, "image of %u x %u x %d is beyond the encoder's limits", width, height, channels);
    return LUTGEN_OK;
}

// pixels already on the device (g.mu held): encode into g.png_out, size in *total (host), synchronises `s`.
int png_encode_locked(const uint8_t* pixels_dev, uint32_t width, uint32_t height, int channels, uint8_t* out_dev, size_t* total, cudaStream_t s) {
    CU_TRY(g.png_scratch.reserve(png_scratch_bytes(width, height, (uint32_t)channels)));
    unsigned long long* total_dev = static_cast<unsigned long long*>(g.png_scratch.p);  // the scratch block starts with the file size
    unsigned launches = 0;
    cudaError_t e = png_encode_async(pixels_dev, width, height, (uint32_t)channels, out_dev, g.png_scratch.p, s, &launches);
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(LUTGEN_ERR_CUDA, "PNG encoder launch failed: %s", cudaGetErrorString(e));
    unsigned long long t = 0;
    CU_TRY(cudaMemcpyAsync(&t, total_dev, sizeof t, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    *total = (size_t)t;
    return LUTGEN_OK;
}
}  // namespace

int lutgen_cuda_png_bound(uint32_t width, uint32_t height, int channels, size_t* bound) {
    if (!bound) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    int rc = png_args_ok(width, height, channels);
    if (rc) return rc;
    *bound = (size_t)png_bound(width, height, (uint32_t)channels);
    return LUTGEN_OK;
}

int lutgen_cuda_encode_png_dev(const uint8_t* pixels_dev, uint32_t width, uint32_t height, int channels, uint8_t* out_dev, size_t capacity,
                               size_t* out_len, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!pixels_dev || !out_dev || !out_len) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    rc = png_args_ok(width, height, channels);
    if (rc) return rc;
    if (capacity < png_bound(width, height, (uint32_t)channels))
        return fail(LUTGEN_ERR_INVALID, "out_dev holds %zu bytes, lutgen_cuda_png_bound() asks for %llu", capacity,
                    png_bound(width, height, (uint32_t)channels));
    std::lock_guard<std::mutex> lk(g.mu);
    return png_encode_locked(pixels_dev, width, height, channels, out_dev, out_len, (cudaStream_t)stream);
}

int lutgen_cuda_encode_png(const uint8_t* pixels, uint32_t width, uint32_t height, int channels, uint8_t* out, size_t capacity, size_t* out_len) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!pixels || !out || !out_len) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    rc = png_args_ok(width, height, channels);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t bytes = (size_t)width * height * channels;
    CU_TRY(g.png_in.reserve(bytes));
    CU_TRY(g.png_out.reserve((size_t)png_bound(width, height, (uint32_t)channels)));
    CU_TRY(cudaMemcpyAsync(g.png_in.p, pixels, bytes, cudaMemcpyHostToDevice, g.stream));
    size_t total = 0;
    rc = png_encode_locked((const uint8_t*)g.png_in.p, width, height, channels, (uint8_t*)g.png_out.p, &total, g.stream);
    if (rc) return rc;
    *out_len = total;
    if (total > capacity) return fail(LUTGEN_ERR_INVALID, "the PNG is %zu bytes, the buffer holds %zu", total, capacity);
    CU_TRY(cudaMemcpyAsync(out, g.png_out.p, total, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int lutgen_cuda_generate_lut_png(const lutgen_remapper* r, uint8_t level, uint8_t* out, size_t capacity, size_t* out_len,
                                 const volatile uint8_t* abort_flag) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || !out || !out_len) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    std::lock_guard<std::mutex> lk(g.mu);
    const uint32_t side = (uint32_t)level * level * level;
    CU_TRY(g.lut_rgb.reserve(3 * lut_entries(level)));
    CU_TRY(g.png_out.reserve((size_t)png_bound(side, side, 3)));
    rc = generate_lut_rows_impl(r, level, (uint8_t*)g.lut_rgb.p, 0, side, g.stream, 0);
    if (rc) return rc;
    size_t total = 0;
    rc = png_encode_locked((const uint8_t*)g.lut_rgb.p, side, side, 3, (uint8_t*)g.png_out.p, &total, g.stream);
    if (rc) return rc;
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    *out_len = total;
    if (total > capacity) return fail(LUTGEN_ERR_INVALID, "the PNG is %zu bytes, the buffer holds %zu", total, capacity);
    CU_TRY(cudaMemcpyAsync(out, g.png_out.p, total, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    return LUTGEN_OK;
}

int lutgen_cuda_apply_hald_png(const uint8_t* rgba, uint32_t width, uint32_t height, const uint8_t* lut_rgb, uint8_t level, uint8_t* out,
                               size_t capacity, size_t* out_len) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!level_ok(level)) return fail(LUTGEN_ERR_INVALID, "level %u outside 2..=33", level);
    if (!rgba || !lut_rgb || !out || !out_len) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    rc = png_args_ok(width, height, 4);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g.mu);
    const size_t npix = (size_t)width * height, lut_bytes = 3 * lut_entries(level);
    CU_TRY(g.lut_rgb.reserve(lut_bytes));
    CU_TRY(g.frame.reserve(npix * 4));
    CU_TRY(g.png_out.reserve((size_t)png_bound(width, height, 4)));
    CU_TRY(cudaMemcpyAsync(g.lut_rgb.p, lut_rgb, lut_bytes, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(cudaMemcpyAsync(g.frame.p, rgba, npix * 4, cudaMemcpyHostToDevice, g.stream));
    rc = lutgen_cuda_apply_hald_dev((uint8_t*)g.frame.p, npix, (const uint8_t*)g.lut_rgb.p, level, g.stream);
    if (rc) return rc;
    size_t total = 0;
    rc = png_encode_locked((const uint8_t*)g.frame.p, width, height, 4, (uint8_t*)g.png_out.p, &total, g.stream);
    if (rc) return rc;
    *out_len = total;
    if (total > capacity) return fail(LUTGEN_ERR_INVALID, "the PNG is %zu bytes, the buffer holds %zu", total, capacity);
    CU_TRY(cudaMemcpyAsync(out, g.png_out.p, total, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// PNG reader (png_decode.cu over png_decode_core.h)
// ---------------------------------------------------------------------------
namespace {
struct IdatSpan {
    size_t off, len;
};

int png_parse_checked(const uint8_t* file, size_t len, lpngd::Info& info, std::vector<IdatSpan>& idat) {
    if (!file) return fail(LUTGEN_ERR_INVALID, "file is NULL");
    const int prc = lpngd::png_parse(file, len, info, idat);
    if (prc == lpngd::P_NOT_PNG) return fail(LUTGEN_ERR_INVALID, "not a PNG file (signature)");
    if (prc == lpngd::P_CORRUPT) return fail(LUTGEN_ERR_INVALID, "corrupt PNG (chunk structure or CRC)");
    if (prc == lpngd::P_UNSUPPORTED)
        return fail(LUTGEN_ERR_UNSUPPORTED, "PNG flavour not decoded here (16-bit samples, interlace, colour key or unknown critical chunk): keep the caller's decoder for it");
    return LUTGEN_OK;
}

// g.mu held. File bytes on the host -> pixels in out_dev (height x width x out_channels), on stream s; synchronises s.
int png_decode_locked(const uint8_t* file, const lpngd::Info& info, const std::vector<IdatSpan>& idat, int out_channels, uint8_t* out_dev, cudaStream_t s) {
    PngDecodePlan plan{};
    plan.width = info.width;
    plan.height = info.height;
    plan.color_type = info.color_type;
    plan.channels = info.channels;
    plan.out_channels = out_channels;
    plan.stream_bytes = (size_t)info.height * ((size_t)info.width * info.channels + 1);
    size_t zbytes = 0;
    for (const IdatSpan& c : idat) zbytes += c.len;
    if (zbytes < 6) return fail(LUTGEN_ERR_INVALID, "corrupt PNG (no image data)");
    plan.zbytes = zbytes;
    // the chunk pattern of this library's own writer: one IDAT per 16 KB piece of the scanline stream + a 4-byte one (Adler-32)
    const size_t npieces = (plan.stream_bytes + lpngd::PIECE - 1) / lpngd::PIECE;
    const bool own = idat.size() == npieces + 1 && idat.back().len == 4 && idat.front().len >= 2;
    plan.npieces = own ? (uint32_t)npieces : 0u;
    // device layout of pngd_z: [the file as it is][pad][the zlib stream, contiguous][pad][piece / chunk offsets][palette].
    // ONE upload of the file (a copy per IDAT chunk costs 20 us each from pageable memory: 60 ms for a level-16 CLUT);
    // the IDAT payloads are gathered into one stream on the device.
    const size_t file_len = idat.back().off + idat.back().len;   // everything behind the last IDAT payload is not needed
    const size_t off_z = (file_len + 255) & ~(size_t)255, off_tab = (off_z + zbytes + 255) & ~(size_t)255;
    const size_t ntab = 2 * idat.size() + npieces + 4, off_pal = off_tab + sizeof(unsigned long long) * ntab;
    CU_TRY(g.pngd_z.reserve(off_pal + 1024));
    CU_TRY(g.pngd_scratch.reserve(png_decode_scratch_bytes(plan)));
    if (!g.pngd_status) CU_TRY(cudaHostAlloc((void**)&g.pngd_status, 4 * sizeof(int), cudaHostAllocDefault));
    uint8_t* fdev = (uint8_t*)g.pngd_z.p;
    uint8_t* zdev = fdev + off_z;
    CU_TRY(cudaMemcpyAsync(fdev, file, file_len, cudaMemcpyHostToDevice, s));
    // table: [src offset in the file, dst offset in the stream] per chunk, then (own files) the pieces' offsets in the stream
    std::vector<unsigned long long> tab;
    size_t pos = 0;
    for (size_t i = 0; i < idat.size(); ++i) {
        tab.push_back(idat[i].off);
        tab.push_back(pos);
        pos += idat[i].len;
    }
    tab.push_back(0);
    tab.push_back(pos);
    const size_t piece_tab = tab.size();
    if (own) {
        size_t p2 = 0;
        for (size_t i = 0; i < npieces; ++i) {
            tab.push_back(p2 + (i == 0 ? 2 : 0));   // the zlib header rides in front of the first piece
            p2 += idat[i].len;
        }
        tab.push_back(p2);
    }
    unsigned long long* tab_dev = (unsigned long long*)(fdev + off_tab);
    CU_TRY(cudaMemcpyAsync(tab_dev, tab.data(), sizeof(unsigned long long) * tab.size(), cudaMemcpyHostToDevice, s));
    if (info.color_type == 3) CU_TRY(cudaMemcpyAsync(fdev + off_pal, info.palette, sizeof(info.palette), cudaMemcpyHostToDevice, s));
    png_decode_gather(fdev, tab_dev, (uint32_t)idat.size(), zdev, s);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned launches = 0;
    cudaError_t e = png_decode_async(plan, zdev, own ? tab_dev + piece_tab : nullptr, (const uint32_t*)(fdev + off_pal), out_dev, g.pngd_scratch.p,
                                     g.pngd_status, s, &launches);
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(LUTGEN_ERR_CUDA, "PNG decoder launch failed: %s", cudaGetErrorString(e));
    CU_TRY(cudaStreamSynchronize(s));
    if (g.pngd_status[0] != 0) return fail(LUTGEN_ERR_INVALID, "corrupt PNG (inflate error %d)", g.pngd_status[0]);
    if (g.pngd_status[1] != 1) return fail(LUTGEN_ERR_INVALID, "corrupt PNG (Adler-32 of the image data does not match)");
    if (g.pngd_status[3] != 0) return fail(LUTGEN_ERR_INVALID, "corrupt PNG (scanline with an unknown filter type)");
    return LUTGEN_OK;
}

int png_out_channels(const lpngd::Info& info, int asked) {
    if (asked == 0) return info.color_type == 3 ? info.out_channels : info.channels;   // native (a palette image has no native pixel form)
    return asked;
}
}  // namespace

int lutgen_cuda_png_info(const uint8_t* file, size_t len, uint32_t* width, uint32_t* height, int* channels, int* has_alpha) {
    lpngd::Info info;
    std::vector<IdatSpan> idat;
    int rc = png_parse_checked(file, len, info, idat);
    if (rc) return rc;
    if (width) *width = info.width;
    if (height) *height = info.height;
    if (channels) *channels = info.color_type == 3 ? info.out_channels : info.channels;
    if (has_alpha) *has_alpha = (info.color_type == 4 || info.color_type == 6 || info.has_trns) ? 1 : 0;
    return LUTGEN_OK;
}

int lutgen_cuda_decode_png_dev(const uint8_t* file, size_t len, int out_channels, uint8_t* out_dev, size_t capacity, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out_dev) return fail(LUTGEN_ERR_INVALID, "out_dev is NULL");
    if (out_channels < 0 || out_channels > 4) return fail(LUTGEN_ERR_INVALID, "%d output channels (0 = as in the file, 1..4)", out_channels);
    lpngd::Info info;
    std::vector<IdatSpan> idat;
    rc = png_parse_checked(file, len, info, idat);
    if (rc) return rc;
    const int oc = png_out_channels(info, out_channels);
    if ((oc == 1 || oc == 2) && !(info.color_type == 0 || info.color_type == 4)) return fail(LUTGEN_ERR_INVALID, "grey output asked of a colour image");
    const size_t need = (size_t)info.width * info.height * (size_t)oc;
    if (need > capacity) return fail(LUTGEN_ERR_INVALID, "the image is %zu bytes, the buffer holds %zu", need, capacity);
    std::lock_guard<std::mutex> lk(g.mu);
    return png_decode_locked(file, info, idat, oc, out_dev, (cudaStream_t)stream);
}

int lutgen_cuda_decode_png(const uint8_t* file, size_t len, int out_channels, uint8_t* out, size_t capacity) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out) return fail(LUTGEN_ERR_INVALID, "out is NULL");
    if (out_channels < 0 || out_channels > 4) return fail(LUTGEN_ERR_INVALID, "%d output channels (0 = as in the file, 1..4)", out_channels);
    lpngd::Info info;
    std::vector<IdatSpan> idat;
    rc = png_parse_checked(file, len, info, idat);
    if (rc) return rc;
    const int oc = png_out_channels(info, out_channels);
    if ((oc == 1 || oc == 2) && !(info.color_type == 0 || info.color_type == 4)) return fail(LUTGEN_ERR_INVALID, "grey output asked of a colour image");
    const size_t need = (size_t)info.width * info.height * (size_t)oc;
    if (need > capacity) return fail(LUTGEN_ERR_INVALID, "the image is %zu bytes, the buffer holds %zu", need, capacity);
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.pngd_out.reserve(need));
    rc = png_decode_locked(file, info, idat, oc, (uint8_t*)g.pngd_out.p, g.stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out, g.pngd_out.p, need, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int lutgen_cuda_apply_hald_png_file(const uint8_t* image_png, size_t image_len, const uint8_t* hald_png, size_t hald_len, uint8_t* out, size_t capacity,
                                    size_t* out_len) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out || !out_len) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    lpngd::Info ii, hi;
    std::vector<IdatSpan> iidat, hidat;
    rc = png_parse_checked(image_png, image_len, ii, iidat);
    if (rc) return rc;
    rc = png_parse_checked(hald_png, hald_len, hi, hidat);
    if (rc) return rc;
    uint8_t level = 0;
    rc = lutgen_cuda_detect_level(hi.width, hi.height, &level);
    if (rc) return rc;
    rc = png_args_ok(ii.width, ii.height, 4);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g.mu);
    const size_t npix = (size_t)ii.width * ii.height;
    CU_TRY(g.lut_rgb.reserve(3 * lut_entries(level)));
    CU_TRY(g.frame.reserve(npix * 4));
    CU_TRY(g.png_out.reserve((size_t)png_bound(ii.width, ii.height, 4)));
    // both files are decoded on the device: the CLUT as RGB8, the image as RGBA8 (load_image's to_rgb8 / to_rgba8)
    rc = png_decode_locked(hald_png, hi, hidat, 3, (uint8_t*)g.lut_rgb.p, g.stream);
    if (rc) return rc;
    rc = png_decode_locked(image_png, ii, iidat, 4, (uint8_t*)g.frame.p, g.stream);
    if (rc) return rc;
    rc = lutgen_cuda_apply_hald_dev((uint8_t*)g.frame.p, npix, (const uint8_t*)g.lut_rgb.p, level, g.stream);
    if (rc) return rc;
    size_t total = 0;
    rc = png_encode_locked((const uint8_t*)g.frame.p, ii.width, ii.height, 4, (uint8_t*)g.png_out.p, &total, g.stream);
    if (rc) return rc;
    *out_len = total;
    if (total > capacity) return fail(LUTGEN_ERR_INVALID, "the PNG is %zu bytes, the buffer holds %zu", total, capacity);
    CU_TRY(cudaMemcpyAsync(out, g.png_out.p, total, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// palette extraction (kernels_extract.cuh over extract_core.h)
// ---------------------------------------------------------------------------
int lutgen_cuda_extract_palette(const uint8_t* rgb, size_t npix, uint32_t color_count, uint32_t iterations, uint8_t* palette_rgb, uint32_t* count_out) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!palette_rgb || !count_out || (npix && !rgb)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (color_count == 0 || color_count > lext::MAX_COLORS) return fail(LUTGEN_ERR_INVALID, "color_count %u outside 1..=%u", color_count, lext::MAX_COLORS);
    if (npix >= (1ull << 32)) return fail(LUTGEN_ERR_UNSUPPORTED, "more than 2^32 pixels in one call");
    if (iterations > 1000) return fail(LUTGEN_ERR_INVALID, "%u Lloyd passes (at most 1000; two launches each)", iterations);
    *count_out = 0;
    if (npix == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    const size_t cap = npix < lext::NKEYS ? npix : lext::NKEYS;  // at most this many distinct colours
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += up(bytes);
        return o;
    };
    const size_t o_counts = take(sizeof(uint32_t) * lext::NKEYS), o_in = take(3 * npix), o_keys = take(4 * cap), o_weights = take(4 * cap),
                 o_list = take(3 * cap), o_lab = take(12 * cap), o_q = take(12 * cap), o_dmin = take(8 * cap),
                 o_small = take(4096 + 12 * lext::MAX_COLORS + 32 * lext::MAX_COLORS + 12 * lext::MAX_COLORS + 3 * lext::MAX_COLORS);
    CU_TRY(g.extract.reserve(off));
    uint8_t* base = (uint8_t*)g.extract.p;
    uint32_t* counts = (uint32_t*)(base + o_counts);
    uint8_t* in = base + o_in;
    uint32_t* keys = (uint32_t*)(base + o_keys);
    uint32_t* weights = (uint32_t*)(base + o_weights);
    uint8_t* list = base + o_list;
    float* lab = (float*)(base + o_lab);
    int32_t* q = (int32_t*)(base + o_q);
    unsigned long long* dmin = (unsigned long long*)(base + o_dmin);
    // small block: seed (16 B) | n_points | n_centroids | ... | centroids | sums | out_lab | out_rgb
    uint8_t* small = base + o_small;
    lext::Seed* seed = (lext::Seed*)small;
    uint32_t* n_points = (uint32_t*)(small + 64);
    uint32_t* n_centroids = (uint32_t*)(small + 128);
    int32_t* centroids = (int32_t*)(small + 4096);
    long long* sums = (long long*)(small + 4096 + 12 * lext::MAX_COLORS);
    float* out_lab = (float*)(small + 4096 + 12 * lext::MAX_COLORS + 32 * lext::MAX_COLORS);
    uint8_t* out_rgb = small + 4096 + 12 * lext::MAX_COLORS + 32 * lext::MAX_COLORS + 12 * lext::MAX_COLORS;
    cudaStream_t s = g.stream;
    CU_TRY(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * lext::NKEYS, s));
    CU_TRY(cudaMemsetAsync(small, 0, up(4096 + 12 * lext::MAX_COLORS + 32 * lext::MAX_COLORS + 12 * lext::MAX_COLORS + 3 * lext::MAX_COLORS), s));
    CU_TRY(cudaMemcpyAsync(in, rgb, 3 * npix, cudaMemcpyHostToDevice, s));
    k_ext_hist<<<blocks_for(npix, EXT_THREADS), EXT_THREADS, 0, s>>>(in, npix, counts);
    LAUNCH_CHECK();
    k_ext_compact<<<lext::NKEYS / EXT_THREADS, EXT_THREADS, 0, s>>>(counts, n_points, keys, weights, list);
    LAUNCH_CHECK();
    uint32_t n = 0;
    CU_TRY(cudaMemcpyAsync(&n, n_points, sizeof n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (n == 0 || n > cap) return fail(LUTGEN_ERR_CUDA, "palette extraction: %u distinct colours out of %zu pixels", n, npix);
    const unsigned grid = blocks_for(n, EXT_THREADS);
    k_srgb_to_oklab<<<grid, 256, 0, s>>>(list, n, lab, g.tables);
    LAUNCH_CHECK();
    k_ext_quantise<<<grid, EXT_THREADS, 0, s>>>(lab, n, q, dmin);
    LAUNCH_CHECK();
    const uint32_t k = color_count < n ? color_count : n;
    for (uint32_t j = 0; j < k; ++j) {
        CU_TRY(cudaMemsetAsync(seed, 0, sizeof(lext::Seed), s));
        k_ext_seed_score<<<grid, EXT_THREADS, 0, s>>>(n, j, q, weights, centroids, n_centroids, dmin, seed);
        LAUNCH_CHECK();
        k_ext_seed_pick<<<grid, EXT_THREADS, 0, s>>>(n, j, keys, weights, n_centroids, dmin, seed);
        LAUNCH_CHECK();
        k_ext_seed_commit<<<grid, EXT_THREADS, 0, s>>>(n, j, keys, q, weights, dmin, seed, centroids, n_centroids);
        LAUNCH_CHECK();
    }
    for (uint32_t it = 0; it < iterations; ++it) {
        k_ext_assign<<<grid, EXT_THREADS, 0, s>>>(n, q, weights, centroids, n_centroids, sums);
        LAUNCH_CHECK();
        k_ext_update<<<1, EXT_THREADS, 0, s>>>(n_centroids, sums, centroids);
        LAUNCH_CHECK();
    }
    k_ext_output<<<1, EXT_THREADS, 0, s>>>(centroids, out_lab);
    LAUNCH_CHECK();
    k_oklab_to_srgb<<<1, 256, 0, s>>>(out_lab, lext::MAX_COLORS, out_rgb, g.tables);
    LAUNCH_CHECK();
    uint32_t found = 0;
    uint8_t host_rgb[3 * lext::MAX_COLORS];
    CU_TRY(cudaMemcpyAsync(&found, n_centroids, sizeof found, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(host_rgb, out_rgb, sizeof host_rgb, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (found > k) return fail(LUTGEN_ERR_CUDA, "palette extraction: %u centroids for %u requested", found, k);
    std::memcpy(palette_rgb, host_rgb, 3 * (size_t)found);
    *count_out = found;
    return LUTGEN_OK;
}

int lutgen_cuda_remap_image_dev(const lutgen_remapper* r, uint8_t* rgba_dev, size_t npix, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || (!rgba_dev && npix)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (reinterpret_cast<uintptr_t>(rgba_dev) & 3u) return fail(LUTGEN_ERR_INVALID, "rgba_dev must be 4-byte aligned");
    if (npix >= (1ull << 32)) return fail(LUTGEN_ERR_UNSUPPORTED, "more than 2^32 pixels in one call");
    cudaStream_t s = (cudaStream_t)stream;  // 0 is CUDA's default stream, as everywhere in CUDA
    Front f{};
    f.out = rgba_dev;
    f.begin = 0;
    f.count = npix;
    rc = launch_remap<FRONT_IMAGE>(f, r, s);
    if (rc) return rc;
    return LUTGEN_OK;
}

int lutgen_cuda_remap_image(const lutgen_remapper* r, uint8_t* rgba, size_t npix, const volatile uint8_t* abort_flag) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!r || (!rgba && npix)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    if (npix == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.frame.reserve(npix * 4));
    CU_TRY(cudaMemcpyAsync(g.frame.p, rgba, npix * 4, cudaMemcpyHostToDevice, g.stream));
    rc = lutgen_cuda_remap_image_dev(r, (uint8_t*)g.frame.p, npix, g.stream);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(g.stream));
    // The reference skips pixels once the flag is set and the caller discards the
    // image (mod.rs:36-42, lib.rs:152-157): report it and leave `rgba` untouched.
    if (abort_flag && *abort_flag) return LUTGEN_ABORTED;
    CU_TRY(cudaMemcpyAsync(rgba, g.frame.p, npix * 4, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------
int lutgen_cuda_measure_peaks(double* fp32_tflops, double* mufu_tops) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!fp32_tflops || !mufu_tops) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.misc.reserve(256));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int blocks = g.sm_count * 8, iters = 4096;
    float ms = 0.f;
    double best_fma = 0.0, best_mufu = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(cudaEventRecord(e0, g.stream));
        k_peak_fma<<<blocks, 256, 0, g.stream>>>((float*)g.misc.p, iters, 0.999f, 0.001f);
        LAUNCH_CHECK();
        CU_TRY(cudaEventRecord(e1, g.stream));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double tf = (double)blocks * 256 * iters * 64 * 2 / (ms * 1e-3) / 1e12;
        if (rep && tf > best_fma) best_fma = tf;
        CU_TRY(cudaEventRecord(e0, g.stream));
        k_peak_mufu<<<blocks, 256, 0, g.stream>>>((float*)g.misc.p, iters / 4);
        LAUNCH_CHECK();
        CU_TRY(cudaEventRecord(e1, g.stream));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double tm = (double)blocks * 256 * (iters / 4) * 64 / (ms * 1e-3) / 1e12;
        if (rep && tm > best_mufu) best_mufu = tm;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *fp32_tflops = best_fma;
    *mufu_tops = best_mufu;
    return LUTGEN_OK;
}

// ---------------------------------------------------------------------------
// oklab
// ---------------------------------------------------------------------------
int lutgen_cuda_srgb_to_oklab(const uint8_t* rgb, size_t n, float* out_lab) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (n && (!rgb || !out_lab)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (n == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.misc.reserve(16 * n));
    float* dout = (float*)g.misc.p;
    uint8_t* din = (uint8_t*)g.misc.p + 12 * n;
    CU_TRY(cudaMemcpyAsync(din, rgb, 3 * n, cudaMemcpyHostToDevice, g.stream));
    k_srgb_to_oklab<<<blocks_for(n, 256), 256, 0, g.stream>>>(din, n, dout, g.tables);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out_lab, dout, 12 * n, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int lutgen_cuda_debug_exchange_probe_dev(uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint8_t* multicast_lut, uint32_t shard_index,
                                         uint32_t shard_count, unsigned* const* barrier_flags_dev_table, int pattern, void* stream) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!luts_rgb_dev || nluts < 2 || nluts > 1 + LUTGEN_MAX_PEERS || !level_ok(level) || (level & 3) || shard_count == 0 || shard_index >= shard_count ||
        pattern < 0 || pattern > 1)
        return fail(LUTGEN_ERR_INVALID, "exchange probe: level %u, %d destinations, shard %u of %u, pattern %d", level, nluts, shard_index, shard_count, pattern);
    uintptr_t bits = reinterpret_cast<uintptr_t>(multicast_lut);
    for (int i = 0; i < nluts; ++i) bits |= reinterpret_cast<uintptr_t>(luts_rgb_dev[i]);
    if (bits & 15u) return fail(LUTGEN_ERR_INVALID, "exchange probe: buffers must be 16-byte aligned");
    Front f{};
    f.out = luts_rgb_dev[0];
    f.nextra = nluts - 1;
    for (int i = 1; i < nluts; ++i) f.extra[i - 1] = luts_rgb_dev[i];
    f.cs = (uint32_t)level * level;
    f.csm1 = f.cs - 1;
    f.shard_index = shard_index;
    f.shard_count = shard_count;
    f.mc = multicast_lut;
    if (barrier_flags_dev_table) {
        f.bar_flags = barrier_flags_dev_table;
        f.bar_nranks = nluts;
        f.bar_rank = (int)shard_index;
    }
    const uint32_t pair = g.counter_next.fetch_add(1, std::memory_order_relaxed) % kCounterPairs;
    k_exchange_probe<<<2 * g.sm_count, WT_THREADS, 0, (cudaStream_t)stream>>>(f, f.cs / 8u, f.cs / 4u, pattern, g.counters + 2 * pair + 1);
    LAUNCH_CHECK();
    return LUTGEN_OK;
}

int lutgen_cuda_debug_timeline(uint32_t* out4, int reset) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out4) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!g.stats) return fail(LUTGEN_ERR_UNSUPPORTED, "set LUTGEN_CUDA_STATS=1 before lutgen_cuda_init");
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(out4, g.stats + 128, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (reset) {
        CU_TRY(cudaMemset(g.stats + 128, 0, 8 * sizeof(uint32_t)));
        CU_TRY(cudaMemset(g.stats + 128, 0xff, sizeof(uint32_t)));
    }
    return LUTGEN_OK;
}

int lutgen_cuda_debug_stats(uint32_t* out128, int reset) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (!out128) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (!g.stats) return fail(LUTGEN_ERR_UNSUPPORTED, "set LUTGEN_CUDA_STATS=1 before lutgen_cuda_init");
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(out128, g.stats, 128 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (reset) CU_TRY(cudaMemset(g.stats, 0, 128 * sizeof(uint32_t)));
    return LUTGEN_OK;
}

int lutgen_cuda_srgb_to_oklab_fast(const uint8_t* rgb, size_t n, float* out_lab) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (n && (!rgb || !out_lab)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (n == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.misc.reserve(16 * n));
    float* dout = (float*)g.misc.p;
    uint8_t* din = (uint8_t*)g.misc.p + 12 * n;
    CU_TRY(cudaMemcpyAsync(din, rgb, 3 * n, cudaMemcpyHostToDevice, g.stream));
    k_srgb_to_oklab_fast<<<blocks_for((n + 1) / 2, 256), 256, 0, g.stream>>>(din, n, dout, g.tables);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out_lab, dout, 12 * n, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

int lutgen_cuda_oklab_to_srgb(const float* lab, size_t n, uint8_t* out_rgb) {
    int rc = ensure_ready();
    if (rc) return rc;
    if (n && (!lab || !out_rgb)) return fail(LUTGEN_ERR_INVALID, "NULL argument");
    if (n == 0) return LUTGEN_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    CU_TRY(g.misc.reserve(16 * n));
    float* din = (float*)g.misc.p;
    uint8_t* dout = (uint8_t*)g.misc.p + 12 * n;
    CU_TRY(cudaMemcpyAsync(din, lab, 12 * n, cudaMemcpyHostToDevice, g.stream));
    k_oklab_to_srgb<<<blocks_for(n, 256), 256, 0, g.stream>>>(din, n, dout, g.tables);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out_rgb, dout, 3 * n, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    return LUTGEN_OK;
}

}  // extern "C"
