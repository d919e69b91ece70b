// png.cu -- PNG encoder kernels: the data-parallel pipeline of png_core.h run by CUDA threads
// (replaces the `image` crate's PNG writer behind `lut.save(path)` /
// `image.save(path)`, crates/cli/src/main.rs:767, 811, 839, 862).
//
//   k_png_choose_filters  four warps per scanline: the five filters' absolute residual sums, the smallest wins
//   k_png_deflate         one CTA per 16 KB piece of the filtered stream: filter, run detection, histogram,
//                         Huffman code, bit stream, CRC-32 and Adler-32 parts, all in shared memory
//   k_png_layout          one CTA: prefix sums of the piece sizes, Adler-32 of the stream, fixed chunks
//   k_png_place           one CTA per piece: its IDAT chunk copied to its place in the file
#include "png_launch.h"

#include <cstdlib>
#include <mutex>

#include "png_core.h"

namespace lutgen {
namespace {

struct CtaExec {
    template <class F>
    __host__ __device__ __forceinline__ void par(F f) {
#ifdef __CUDA_ARCH__
        f((int)threadIdx.x);
        __syncthreads();
#else
        (void)f;
#endif
    }
};

constexpr int FILTER_WARPS = 8;  // warps per CTA

// SPLIT warps share one scanline (one warp per 12 KB row does not keep enough loads in flight to fill the SM).
template <int SPLIT>
__global__ void __launch_bounds__(FILTER_WARPS * 32) k_png_choose_filters(lpng::Image im, uint8_t* filters) {
    constexpr int ROWS = FILTER_WARPS / SPLIT;
    __shared__ uint32_t part[FILTER_WARPS][5];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t y = blockIdx.x * ROWS + warp / SPLIT, piece = warp % SPLIT;
    uint32_t cost[5] = {0u, 0u, 0u, 0u, 0u};
    if (y < im.height) {
        const uint8_t* row = im.px + (size_t)y * im.rb;
        lpng::row_costs(row, y > 0 ? row - im.rb : nullptr, im.rb, im.bpp, piece * 32u + lane, 32u * SPLIT, cost);
    }
#pragma unroll
    for (int ft = 0; ft < 5; ++ft) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cost[ft] += __shfl_xor_sync(0xFFFFFFFFu, cost[ft], d);
        if (lane == 0) part[warp][ft] = cost[ft];
    }
    __syncthreads();
    if (piece == 0 && lane == 0 && y < im.height) {
        uint32_t total[5];
#pragma unroll
        for (int ft = 0; ft < 5; ++ft) {
            total[ft] = 0;
            for (int p = 0; p < SPLIT; ++p) total[ft] += part[warp + p][ft];
        }
        filters[y] = (uint8_t)lpng::pick_filter(total);
    }
}

template <int SPLIT>
void launch_choose_filters(const lpng::Image& im, uint8_t* filters, cudaStream_t s) {
    constexpr int ROWS = FILTER_WARPS / SPLIT;
    k_png_choose_filters<SPLIT><<<(im.height + ROWS - 1) / ROWS, FILTER_WARPS * 32, 0, s>>>(im, filters);
}

__global__ void __launch_bounds__(lpng::THREADS, 5)
    k_png_deflate(lpng::Image im, const lpng::Tables* tables, uint32_t nchunks, uint32_t* slots, lpng::ChunkMeta* meta) {
    extern __shared__ uint4 png_smem[];
    lpng::Shared& S = *reinterpret_cast<lpng::Shared*>(png_smem);
    CtaExec ex;
    const uint32_t c = blockIdx.x;
    lpng::encode_chunk(ex, S, im, *tables, c, nchunks, slots + (size_t)c * lpng::OUT_WORDS, meta + c);
}

__global__ void __launch_bounds__(lpng::THREADS) k_png_layout(const lpng::ChunkMeta* meta, uint32_t nchunks, uint32_t width, uint32_t height, uint32_t bpp,
                                                             unsigned long long* offsets, uint8_t* out, unsigned long long* total) {
    __shared__ lpng::LayoutShared L;
    CtaExec ex;
    lpng::layout(ex, L, meta, nchunks, width, height, bpp, offsets, out, total);
}

__global__ void __launch_bounds__(lpng::THREADS) k_png_place(const lpng::ChunkMeta* meta, const uint32_t* slots, const unsigned long long* offsets,
                                                            uint8_t* out) {
    CtaExec ex;
    const uint32_t c = blockIdx.x;
    lpng::place_chunk(ex, meta[c], slots + (size_t)c * lpng::OUT_WORDS, out + offsets[c]);
}

std::mutex g_png_mu;
lpng::Tables* g_png_tables = nullptr;  // device

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ScratchLayout {
    size_t filters, meta, offsets, slots, total;
    uint32_t nchunks;
};

ScratchLayout scratch_layout(uint32_t width, uint32_t height, uint32_t bpp) {
    ScratchLayout L;
    unsigned long long stream = (unsigned long long)height * ((unsigned long long)width * bpp + 1ull);
    L.nchunks = (uint32_t)((stream + lpng::CHUNK - 1) / lpng::CHUNK);
    L.filters = 256;  // the first 8 bytes of the block receive the file size
    L.meta = align_up(L.filters + height, 256);
    L.offsets = align_up(L.meta + sizeof(lpng::ChunkMeta) * (size_t)L.nchunks, 256);
    L.slots = align_up(L.offsets + sizeof(unsigned long long) * (size_t)L.nchunks, 256);
    L.total = L.slots + (size_t)L.nchunks * lpng::OUT_BYTES;
    return L;
}

}  // namespace

unsigned long long png_bound(uint32_t width, uint32_t height, uint32_t bpp) { return lpng::png_bound(width, height, bpp); }

size_t png_scratch_bytes(uint32_t width, uint32_t height, uint32_t bpp) { return scratch_layout(width, height, bpp).total; }

cudaError_t png_prepare() {
    std::lock_guard<std::mutex> lk(g_png_mu);
    if (g_png_tables) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_png_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lpng::Shared));
    if (e != cudaSuccess) return e;
    lpng::Tables host;
    lpng::make_tables(host);
    lpng::Tables* dev = nullptr;
    e = cudaMalloc(&dev, sizeof(lpng::Tables));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(dev, &host, sizeof(lpng::Tables), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return e;
    }
    g_png_tables = dev;
    return cudaSuccess;
}

void png_release() {
    std::lock_guard<std::mutex> lk(g_png_mu);
    if (g_png_tables) cudaFree(g_png_tables);
    g_png_tables = nullptr;
}

cudaError_t png_encode_async(const uint8_t* pixels_dev, uint32_t width, uint32_t height, uint32_t bpp, uint8_t* out_dev, void* scratch_dev,
                             cudaStream_t s, unsigned* launches) {
    cudaError_t e = png_prepare();
    if (e != cudaSuccess) return e;
    const ScratchLayout L = scratch_layout(width, height, bpp);
    uint8_t* base = static_cast<uint8_t*>(scratch_dev);
    unsigned long long* total_dev = reinterpret_cast<unsigned long long*>(base);
    uint8_t* filters = base + L.filters;
    lpng::ChunkMeta* meta = reinterpret_cast<lpng::ChunkMeta*>(base + L.meta);
    unsigned long long* offsets = reinterpret_cast<unsigned long long*>(base + L.offsets);
    uint32_t* slots = reinterpret_cast<uint32_t*>(base + L.slots);
    lpng::Image im{pixels_dev, filters, width, height, bpp, width * bpp, (unsigned long long)height * ((unsigned long long)width * bpp + 1ull)};
    // warps per scanline: LUTGEN_CUDA_PNG_SPLIT (1, 2, 4, 8) overrides the default, for measurements
    int split = 4;
    if (const char* e = std::getenv("LUTGEN_CUDA_PNG_SPLIT")) split = std::atoi(e);
    switch (split) {
        case 1: launch_choose_filters<1>(im, filters, s); break;
        case 2: launch_choose_filters<2>(im, filters, s); break;
        case 8: launch_choose_filters<8>(im, filters, s); break;
        default: launch_choose_filters<4>(im, filters, s); break;
    }
    k_png_deflate<<<L.nchunks, lpng::THREADS, sizeof(lpng::Shared), s>>>(im, g_png_tables, L.nchunks, slots, meta);
    k_png_layout<<<1, lpng::THREADS, 0, s>>>(meta, L.nchunks, width, height, bpp, offsets, out_dev, total_dev);
    k_png_place<<<L.nchunks, lpng::THREADS, 0, s>>>(meta, slots, offsets, out_dev);
    if (launches) *launches = 4;
    return cudaGetLastError();
}

}  // namespace lutgen
