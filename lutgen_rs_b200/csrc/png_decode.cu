// png_decode.cu -- PNG decoder kernels (png_decode_core.h): inflate, Adler-32, scanline reconstruction, channel expansion.
// Replaces the decode half of `load_image` (crates/cli/src/main.rs:602-613: image::open + to_rgba8 / to_rgb8) for the
// files `lutgen apply` reads, so that a frame can go file -> pixels -> corrected pixels -> file without its pixels ever
// crossing PCIe (only the compressed bytes do).
//
//   k_pngd_inflate_pieces  one thread per 16 KB piece of a file this library wrote (independent block sequences)
//   k_pngd_inflate_serial  one thread: any other zlib stream (a deflate stream is one serial chain of bits)
//   k_pngd_adler_parts / k_pngd_adler_check   Adler-32 of the inflated stream against the stream's trailer
//   k_pngd_unfilter        Sub / Up / Average / Paeth as an anti-diagonal wavefront: a warp per band of 32 rows
//   k_pngd_expand          to the channel count asked for (palette, grey -> RGB, alpha added or dropped)
#include <cstdlib>
#include "png_decode_launch.h"

#include "png_decode_core.h"

namespace lutgen {
namespace {

constexpr int ADLER_SPAN = 4096;

// One WARP per piece. Lane 0 decodes into the warp's 16 KB of shared memory -- a match is a chain of byte copies each
// reading what the copy before it wrote, 30 cycles apart in shared memory and an L2 round trip apart in global memory
// (the first version, one thread per piece writing global memory, took 94 ms for a level-16 CLUT; this one ~1 ms) --
// then the whole warp stores the piece, 16 bytes per lane.
constexpr int PIECE_WARPS = 4;
__global__ void __launch_bounds__(PIECE_WARPS * 32) k_pngd_inflate_pieces(const uint8_t* __restrict__ z, const unsigned long long* __restrict__ off,
                                                                          uint32_t npieces, uint8_t* __restrict__ dst, unsigned long long total,
                                                                          int* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t s_piece[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * PIECE_WARPS + warp;
    if (i >= npieces) return;
    uint8_t* buf = s_piece + (size_t)warp * lpngd::PIECE;
    const unsigned long long o0 = (unsigned long long)i * lpngd::PIECE;
    const uint32_t expect = (uint32_t)(total - o0 < (unsigned long long)lpngd::PIECE ? total - o0 : (unsigned long long)lpngd::PIECE);
    if (lane == 0) {
        lpngd::InflateScratch X;
        const int rc = lpngd::inflate_piece(z + off[i], (size_t)(off[i + 1] - off[i]), buf, expect, X);
        if (rc != lpngd::OK) atomicCAS(status, 0, rc);
    }
    __syncwarp();
    // o0 is a multiple of 16 KB and the stream buffer is 256-byte aligned: 16-byte stores
    uint4* g4 = reinterpret_cast<uint4*>(dst + o0);
    const uint4* s4 = reinterpret_cast<const uint4*>(buf);
    const uint32_t n16 = expect / 16u;
    for (uint32_t k = lane; k < n16; k += 32u) g4[k] = s4[k];
    for (uint32_t k = n16 * 16u + lane; k < expect; k += 32u) dst[o0 + k] = buf[k];
}

// Any other zlib stream: one thread (a deflate stream is one serial chain of bits), with the 32 KB match history in
// shared memory.
__global__ void k_pngd_inflate_serial(const uint8_t* __restrict__ z, size_t n, uint8_t* __restrict__ dst, size_t cap, int* __restrict__ status) {
    __shared__ uint8_t window[32768];
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    lpngd::InflateScratch X;
    size_t produced = 0;
    int rc = lpngd::inflate_zlib(z, n, dst, cap, &produced, X, window);
    if (rc == lpngd::OK && produced != cap) rc = lpngd::ERR_LENGTH;
    status[0] = rc;
}

// parts[i] = (a, b) of span i (ADLER_SPAN bytes each, the last one shorter)
__global__ void k_pngd_adler_parts(const uint8_t* __restrict__ p, unsigned long long n, uint2* __restrict__ parts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long o = i * ADLER_SPAN;
    if (o >= n) return;
    const size_t len = (size_t)(n - o < (unsigned long long)ADLER_SPAN ? n - o : (unsigned long long)ADLER_SPAN);
    uint32_t a, b;
    lpngd::adler_span(p + o, len, a, b);
    parts[i] = make_uint2(a, b);
}

// One warp folds the parts (each lane a contiguous run of them, then the 32 runs in order) and compares with the four
// bytes that follow the deflate data (big endian).
__global__ void k_pngd_adler_check(const uint2* __restrict__ parts, unsigned long long n, const uint8_t* __restrict__ trailer, int* __restrict__ status) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const unsigned long long nparts = (n + ADLER_SPAN - 1) / ADLER_SPAN, per = (nparts + 31) / 32;
    const unsigned long long p0 = threadIdx.x * per, p1 = p0 + per < nparts ? p0 + per : nparts;
    // sums of this lane's run alone (A from 0), and its length in bytes
    uint32_t a = 0, b = 0;
    unsigned long long len = 0;
    for (unsigned long long i = p0; i < p1; ++i) {
        const size_t l = (size_t)(n - i * ADLER_SPAN < (unsigned long long)ADLER_SPAN ? n - i * ADLER_SPAN : (unsigned long long)ADLER_SPAN);
        // appending a span to sums that started from zero: b += l * a + b_span, a += a_span
        b = (uint32_t)((b + (unsigned long long)(l % lpngd::ADLER_MOD) * a + parts[i].y) % lpngd::ADLER_MOD);
        a = (a + parts[i].x) % lpngd::ADLER_MOD;
        len += l;
    }
    uint32_t A = 1, B = 0;
    for (int t = 0; t < 32; ++t) {
        const uint32_t ta = __shfl_sync(0xffffffffu, a, t), tb = __shfl_sync(0xffffffffu, b, t);
        const unsigned long long tl = __shfl_sync(0xffffffffu, len, t);
        lpngd::adler_append(A, B, ta, tb, (size_t)tl);
    }
    if (threadIdx.x == 0) {
        const uint32_t want = ((uint32_t)trailer[0] << 24) | ((uint32_t)trailer[1] << 16) | ((uint32_t)trailer[2] << 8) | trailer[3];
        status[1] = (((B << 16) | A) == want) ? 1 : 0;
    }
}

__device__ __forceinline__ uint32_t load_pixel(const uint8_t* p, uint32_t bpp) {
    uint32_t v = 0;
    for (uint32_t k = 0; k < bpp; ++k) v |= (uint32_t)p[k] << (8u * k);
    return v;
}
__device__ __forceinline__ uint32_t load_pixel_cg(const uint8_t* p, uint32_t bpp) {   // written by another warp: not through L1
    uint32_t v = 0;
    for (uint32_t k = 0; k < bpp; ++k) v |= (uint32_t)__ldcg(p + k) << (8u * k);
    return v;
}
__device__ __forceinline__ void store_pixel(uint8_t* p, uint32_t v, uint32_t bpp) {
    for (uint32_t k = 0; k < bpp; ++k) p[k] = (uint8_t)(v >> (8u * k));
}

// Reconstruction. Band = 32 consecutive rows = one warp, lane i = row band*32 + i, and at step s lane i does pixel
// x = s - i: the pixel above it was done by lane i-1 one step earlier (it arrives by shuffle), the one above-left
// the step before that. Lane 0's row above belongs to the previous band: that band publishes how many pixels of its
// last row are done (progress[band], release) and this one reads them 32 at a time (one per lane, acquire), so the
// round trip through L2 is paid once per 32 steps. The filtered bytes of the next eight steps are loaded while the
// current eight are worked on (their latency was the whole step in the first version: 18.7 ms for a level-16 CLUT).
// Bands are numbered in launch order, so a band only ever waits for CTAs that were scheduled before its own.
constexpr int UNF_AHEAD = 8;
__global__ void __launch_bounds__(256) k_pngd_unfilter(const uint8_t* __restrict__ stream, uint32_t w, uint32_t h, uint32_t bpp, uint8_t* __restrict__ out,
                                                        int* progress, int* status) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, band = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nbands = (h + 31u) / 32u;
    if (band >= nbands) return;
    const size_t rb = (size_t)w * bpp;
    const uint32_t r = band * 32u + lane;
    const bool valid = r < h;
    const uint8_t* src = stream + (size_t)(valid ? r : 0) * (rb + 1);
    const uint32_t ft = valid ? src[0] : 0u;
    if (ft > 4u) atomicOr(status + 3, 1);   // not a filter type (PNG specification 9.2): the file is corrupt
    ++src;
    uint8_t* dstrow = out + (size_t)(valid ? r : 0) * rb;
    const uint8_t* above_row = band > 0 ? out + (size_t)(band * 32u - 1u) * rb : nullptr;
    const bool publish = band + 1 < nbands;   // full band with a follower: lane 31 is its last row
    uint32_t a = 0, prev_b = 0, last_out = 0, above = 0;
    const uint32_t steps = w + 31u;
    const uint32_t mA = ft == 1u ? 255u : 0u, mB = ft == 2u ? 255u : 0u, mV = ft == 3u ? 255u : 0u, mP = ft == 4u ? 255u : 0u;
    // Stores: a lane writes its own row, so a warp's store touches 32 rows -- 32 sectors per instruction, and as single
    // bytes that was 50 M partial-sector writes for a level-16 CLUT (the whole 17 ms of the first version). Pixels leave as
    // 32-bit words instead: one per RGBA pixel, three per four RGB pixels (when the rows are word aligned).
    const bool words4 = bpp == 4u && (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
    const bool words3 = bpp == 3u && (reinterpret_cast<uintptr_t>(out) & 3u) == 0 && (rb & 3u) == 0;
    uint32_t pk0 = 0, pk1 = 0, pk2 = 0;
    uint32_t blk[UNF_AHEAD], nxt[UNF_AHEAD];
#pragma unroll
    for (int j = 0; j < UNF_AHEAD; ++j) {
        const int x = j - (int)lane;
        blk[j] = (valid && x >= 0 && x < (int)w) ? load_pixel(src + (size_t)x * bpp, bpp) : 0u;
    }
    for (uint32_t s0 = 0; s0 < steps; s0 += UNF_AHEAD) {
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) {
            const int x = (int)(s0 + UNF_AHEAD) + j - (int)lane;
            nxt[j] = (valid && x >= 0 && x < (int)w) ? load_pixel(src + (size_t)x * bpp, bpp) : 0u;
        }
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) {
            const uint32_t s = s0 + j;
            // the row above the band, 32 pixels at a time (s0 is a multiple of 8: the test is uniform)
            if (above_row && s < w && (s & 31u) == 0) {
                const uint32_t need = min(s + 32u, w);
                if (lane == 0) {
                    // relaxed polls with a pause, ONE fence once the count is there: an acquire load per poll compiles to an
                    // L1 invalidation each (CCTL.IVALL), and a hundred waiting bands polling flushed the working bands' L1
                    // sixteen million times (33 % of the stall samples, 15 ms for a level-16 CLUT)
                    int v;
                    for (;;) {
                        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(progress + band - 1) : "memory");
                        if ((uint32_t)v >= need) break;
                        __nanosleep(200);
                    }
                    __threadfence();
                }
                __syncwarp();
                const uint32_t xa = s + lane;
                above = xa < w ? load_pixel_cg(above_row + (size_t)xa * bpp, bpp) : 0u;
            }
            uint32_t b = __shfl_up_sync(FULL, last_out, 1);
            const uint32_t from_above = __shfl_sync(FULL, above, s & 31u);
            if (lane == 0) b = above_row ? from_above : 0u;
            const int x = (int)s - (int)lane;
            uint32_t cur = 0;
            if (valid && x >= 0 && x < (int)w) {
                cur = bpp == 3u ? lpngd::unfilter_pixel_masked<3>(blk[j], a, b, prev_b, mA, mB, mV, mP)
                                : (bpp == 4u ? lpngd::unfilter_pixel_masked<4>(blk[j], a, b, prev_b, mA, mB, mV, mP)
                                             : (bpp == 1u ? lpngd::unfilter_pixel_masked<1>(blk[j], a, b, prev_b, mA, mB, mV, mP)
                                                          : lpngd::unfilter_pixel_masked<2>(blk[j], a, b, prev_b, mA, mB, mV, mP)));
                if (words4) {
                    reinterpret_cast<uint32_t*>(dstrow)[x] = cur;
                } else if (words3) {
                    const uint32_t q = (uint32_t)x & 3u;
                    if (q == 0) pk0 = cur;
                    else if (q == 1) { pk0 |= cur << 24; pk1 = cur >> 8; }
                    else if (q == 2) { pk1 |= cur << 16; pk2 = cur >> 16; }
                    else {
                        pk2 |= cur << 8;
                        uint32_t* q32 = reinterpret_cast<uint32_t*>(dstrow + (size_t)(x - 3) * 3u);
                        q32[0] = pk0; q32[1] = pk1; q32[2] = pk2;
                    }
                    if (x == (int)w - 1 && q != 3u) {   // the row's last one to three pixels
                        uint8_t* t = dstrow + (size_t)(x - (int)q) * 3u;
                        const uint32_t nb = (q + 1u) * 3u;
                        for (uint32_t k = 0; k < nb; ++k) t[k] = (uint8_t)((k < 4 ? pk0 >> (8u * k) : (k < 8 ? pk1 >> (8u * (k - 4)) : pk2 >> (8u * (k - 8)))));
                    }
                } else {
                    store_pixel(dstrow + (size_t)x * bpp, cur, bpp);
                }
                a = cur;
            }
            prev_b = b;
            last_out = cur;
            if (publish && lane == 31u && x >= 0 && x < (int)w && (((uint32_t)x & 31u) == 31u || x == (int)w - 1)) {
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + band), "r"(x + 1) : "memory");
            }
        }
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) blk[j] = nxt[j];
    }
}

// The same wavefront with the pixel size known at compile time and no branch that depends on the lane's position in its
// row: the generic kernel above picks its pixel routine and its store form at run time inside every step, and its packing
// of RGB pixels into words branches on x mod 4 -- which differs from lane to lane, so a warp walked all four arms at
// every step (277 instructions per step, 15 ms for a level-16 CLUT). WORDS: rows start on word boundaries (RGBA always
// when `out` is aligned; RGB when the row length is a multiple of 4 bytes too): pixels leave as 32-bit words.
template <int BPP>
__device__ __forceinline__ uint32_t load_pixel_t(const uint8_t* p) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < BPP; ++k) v |= (uint32_t)p[k] << (8 * k);
    return v;
}
template <int BPP>
__device__ __forceinline__ uint32_t load_pixel_cg_t(const uint8_t* p) {   // written by another warp: not through L1
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < BPP; ++k) v |= (uint32_t)__ldcg(p + k) << (8 * k);
    return v;
}
template <int BPP, bool WORDS>
__global__ void __launch_bounds__(256) k_pngd_unfilter_t(const uint8_t* __restrict__ stream, uint32_t w, uint32_t h, uint8_t* __restrict__ out, int* progress,
                                                          int* status) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, band = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nbands = (h + 31u) / 32u;
    if (band >= nbands) return;
    const size_t rb = (size_t)w * BPP;
    const uint32_t r = band * 32u + lane;
    const bool valid = r < h;
    const uint8_t* src = stream + (size_t)(valid ? r : 0) * (rb + 1);
    const uint32_t ft = valid ? src[0] : 0u;
    if (ft > 4u) atomicOr(status + 3, 1);   // not a filter type (PNG specification 9.2): the file is corrupt
    ++src;
    uint8_t* dstrow = out + (size_t)(valid ? r : 0) * rb;
    const uint8_t* above_row = band > 0 ? out + (size_t)(band * 32u - 1u) * rb : nullptr;
    const bool publish = band + 1 < nbands;   // full band with a follower: lane 31 is its last row
    uint32_t a = 0, prev_b = 0, last_out = 0, above = 0;
    const uint32_t steps = w + 31u;
    const uint32_t mA = ft == 1u ? 255u : 0u, mB = ft == 2u ? 255u : 0u, mV = ft == 3u ? 255u : 0u, mP = ft == 4u ? 255u : 0u;
    uint32_t pk0 = 0, pk1 = 0, pk2 = 0;
    uint32_t blk[UNF_AHEAD], nxt[UNF_AHEAD];
#pragma unroll
    for (int j = 0; j < UNF_AHEAD; ++j) {
        const int x = j - (int)lane;
        blk[j] = (valid && x >= 0 && x < (int)w) ? load_pixel_t<BPP>(src + (size_t)x * BPP) : 0u;
    }
    for (uint32_t s0 = 0; s0 < steps; s0 += UNF_AHEAD) {
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) {
            const int x = (int)(s0 + UNF_AHEAD) + j - (int)lane;
            nxt[j] = (valid && x >= 0 && x < (int)w) ? load_pixel_t<BPP>(src + (size_t)x * BPP) : 0u;
        }
        // the row above the band, 32 pixels at a time: once per four trips (s0 advances by 8), outside the unrolled steps
        if (above_row && s0 < w && (s0 & 31u) == 0) {
            const uint32_t need = min(s0 + 32u, w);
            if (lane == 0) {
                int v;
                for (;;) {   // relaxed polls with a pause, ONE fence once the count is there (see the generic kernel)
                    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(progress + band - 1) : "memory");
                    if ((uint32_t)v >= need) break;
                    __nanosleep(100);
                }
                __threadfence();
            }
            __syncwarp();
            const uint32_t xa = s0 + lane;
            above = xa < w ? load_pixel_cg_t<BPP>(above_row + (size_t)xa * BPP) : 0u;
        }
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) {
            const uint32_t s = s0 + j;
            uint32_t b = __shfl_up_sync(FULL, last_out, 1);
            const uint32_t from_above = __shfl_sync(FULL, above, s & 31u);
            if (lane == 0) b = above_row ? from_above : 0u;
            const int x = (int)s - (int)lane;
            const bool in_row = valid && x >= 0 && x < (int)w;
            uint32_t cur = lpngd::unfilter_pixel_masked<BPP>(blk[j], a, b, prev_b, mA, mB, mV, mP);
            cur = in_row ? cur : 0u;
            if (WORDS && BPP == 4) {
                if (in_row) reinterpret_cast<uint32_t*>(dstrow)[x] = cur;
            } else if (WORDS && BPP == 3) {
                // four pixels = three words, assembled without a branch on x mod 4: pixel q lands at bit 24 q
                const uint32_t q = (uint32_t)x & 3u, sh = (q * 24u) & 31u;
                const uint32_t lo = cur << sh, hi = sh ? cur >> (32u - sh) : 0u;   // cur < 2^24: hi is empty for q = 0 and q = 3
                pk0 = q == 0u ? lo : (q == 1u ? (pk0 | lo) : pk0);
                pk1 = q == 1u ? hi : (q == 2u ? (pk1 | lo) : pk1);
                pk2 = q == 2u ? hi : (q == 3u ? (pk2 | lo) : pk2);
                if (in_row && q == 3u) {
                    uint32_t* q32 = reinterpret_cast<uint32_t*>(dstrow + (size_t)(x - 3) * 3u);
                    q32[0] = pk0; q32[1] = pk1; q32[2] = pk2;
                } else if (in_row && x == (int)w - 1) {   // the row's last one to three pixels (once per row)
                    uint8_t* t = dstrow + (size_t)(x - (int)q) * 3u;
                    const uint32_t nb = (q + 1u) * 3u;
                    for (uint32_t k = 0; k < nb; ++k) t[k] = (uint8_t)((k < 4 ? pk0 >> (8u * k) : (k < 8 ? pk1 >> (8u * (k - 4)) : pk2 >> (8u * (k - 8)))));
                }
            } else if (in_row) {
#pragma unroll
                for (int k = 0; k < BPP; ++k) dstrow[(size_t)x * BPP + k] = (uint8_t)(cur >> (8 * k));
            }
            a = in_row ? cur : a;
            prev_b = b;
            last_out = cur;
            if (publish && lane == 31u && x >= 0 && x < (int)w && (((uint32_t)x & 31u) == 31u || x == (int)w - 1)) {
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + band), "r"(x + 1) : "memory");
            }
        }
#pragma unroll
        for (int j = 0; j < UNF_AHEAD; ++j) blk[j] = nxt[j];
    }
}

__global__ void k_pngd_expand(const uint8_t* __restrict__ rows, unsigned long long npix, int color_type, int channels, int out_channels,
                              const uint32_t* __restrict__ palette, uint8_t* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const uint8_t* p = rows + i * (unsigned long long)channels;
    uint32_t r, g, b, al = 255u;
    if (color_type == 3) {
        const uint32_t e = palette[p[0]];
        r = e & 255u; g = (e >> 8) & 255u; b = (e >> 16) & 255u; al = e >> 24;
    } else if (color_type == 0) {
        r = g = b = p[0];
    } else if (color_type == 4) {
        r = g = b = p[0]; al = p[1];
    } else if (color_type == 2) {
        r = p[0]; g = p[1]; b = p[2];
    } else {
        r = p[0]; g = p[1]; b = p[2]; al = p[3];
    }
    uint8_t* q = out + i * (unsigned long long)out_channels;
    if (out_channels == 1) {
        q[0] = (uint8_t)r;   // only asked for grey sources
    } else if (out_channels == 2) {
        q[0] = (uint8_t)r; q[1] = (uint8_t)al;
    } else {
        q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b;
        if (out_channels == 4) q[3] = (uint8_t)al;
    }
}

// One CTA per IDAT chunk: its payload to its place in the contiguous stream.
__global__ void k_pngd_gather(const uint8_t* __restrict__ file, const unsigned long long* __restrict__ tab, uint8_t* __restrict__ z) {
    const unsigned long long src = tab[2 * blockIdx.x], dst = tab[2 * blockIdx.x + 1], len = tab[2 * blockIdx.x + 3] - dst;
    for (unsigned long long i = threadIdx.x; i < len; i += blockDim.x) z[dst + i] = file[src + i];
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DecodeLayout {
    size_t stream, rows, parts, progress, status, total;
};
DecodeLayout decode_layout(const PngDecodePlan& p) {
    DecodeLayout L;
    const size_t rowbytes = (size_t)p.height * p.width * (size_t)p.channels;
    L.stream = 0;
    L.rows = align_up(L.stream + p.stream_bytes, 256);
    L.parts = align_up(L.rows + rowbytes, 256);
    L.progress = align_up(L.parts + sizeof(uint2) * ((p.stream_bytes + ADLER_SPAN - 1) / ADLER_SPAN + 1), 256);
    L.status = align_up(L.progress + sizeof(int) * ((p.height + 31) / 32 + 1), 256);
    L.total = L.status + 256;
    return L;
}

}  // namespace

size_t png_decode_scratch_bytes(const PngDecodePlan& plan) { return decode_layout(plan).total; }

void png_decode_gather(const uint8_t* file_dev, const unsigned long long* tab_dev, uint32_t nchunks, uint8_t* z_dev, cudaStream_t s) {
    k_pngd_gather<<<nchunks, 256, 0, s>>>(file_dev, tab_dev, z_dev);
}

cudaError_t png_decode_async(const PngDecodePlan& plan, const uint8_t* z_dev, const unsigned long long* piece_off_dev, const uint32_t* palette_dev,
                             uint8_t* out_dev, void* scratch_dev, int* status_host, cudaStream_t s, unsigned* launches) {
    {
        // per device: more than 48 KB of dynamic shared memory has to be asked for
        static bool attr_done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {
            cudaError_t ea = cudaFuncSetAttribute(k_pngd_inflate_pieces, cudaFuncAttributeMaxDynamicSharedMemorySize, PIECE_WARPS * lpngd::PIECE);
            if (ea != cudaSuccess) return ea;
            attr_done[dev] = true;
        }
    }
    const DecodeLayout L = decode_layout(plan);
    uint8_t* base = static_cast<uint8_t*>(scratch_dev);
    uint8_t* stream = base + L.stream;
    uint8_t* rows = base + L.rows;
    uint2* parts = reinterpret_cast<uint2*>(base + L.parts);
    int* progress = reinterpret_cast<int*>(base + L.progress);
    int* status = reinterpret_cast<int*>(base + L.status);
    unsigned nl = 0;
    cudaError_t e = cudaMemsetAsync(base + L.progress, 0, L.total - L.progress, s);   // progress + status
    if (e != cudaSuccess) return e;
    const unsigned long long total = plan.stream_bytes;
    const unsigned nparts = (unsigned)((total + ADLER_SPAN - 1) / ADLER_SPAN);
    const uint8_t* trailer = z_dev + plan.zbytes - 4;
    bool done = false;
    status_host[0] = status_host[1] = status_host[2] = status_host[3] = 0;
    if (plan.npieces > 0 && piece_off_dev) {
        k_pngd_inflate_pieces<<<(plan.npieces + PIECE_WARPS - 1) / PIECE_WARPS, PIECE_WARPS * 32, PIECE_WARPS * lpngd::PIECE, s>>>(z_dev, piece_off_dev, plan.npieces, stream,
                                                                                                                   total, status);
        k_pngd_adler_parts<<<(nparts + 127) / 128, 128, 0, s>>>(stream, total, parts);
        k_pngd_adler_check<<<1, 32, 0, s>>>(parts, total, trailer, status);
        nl += 3;
        e = cudaMemcpyAsync(status_host, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        done = status_host[0] == 0 && status_host[1] == 1;
        status_host[2] = 1;
        if (!done) {
            e = cudaMemsetAsync(status, 0, 2 * sizeof(int), s);
            if (e != cudaSuccess) return e;
        }
    }
    if (!done) {
        k_pngd_inflate_serial<<<1, 32, 0, s>>>(z_dev, plan.zbytes, stream, (size_t)total, status);
        k_pngd_adler_parts<<<(nparts + 127) / 128, 128, 0, s>>>(stream, total, parts);
        k_pngd_adler_check<<<1, 32, 0, s>>>(parts, total, trailer, status);
        nl += 3;
        e = cudaMemcpyAsync(status_host, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
        status_host[2] = 2;
    }
    const bool direct = plan.out_channels == plan.channels && plan.color_type != 3;
    const unsigned nbands = (plan.height + 31) / 32;
    {
        uint8_t* dst = direct ? out_dev : rows;
        const size_t rowb = (size_t)plan.width * (size_t)plan.channels;
        const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
        const unsigned grid = (nbands + 7) / 8;
        if (getenv("LUTGEN_CUDA_PNG_GENERIC_UNFILTER")) k_pngd_unfilter<<<grid, 256, 0, s>>>(stream, plan.width, plan.height, (uint32_t)plan.channels, dst, progress, status);
        else if (plan.channels == 4 && aligned) k_pngd_unfilter_t<4, true><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
        else if (plan.channels == 4) k_pngd_unfilter_t<4, false><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
        else if (plan.channels == 3 && aligned && (rowb & 3u) == 0) k_pngd_unfilter_t<3, true><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
        else if (plan.channels == 3) k_pngd_unfilter_t<3, false><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
        else if (plan.channels == 2) k_pngd_unfilter_t<2, false><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
        else k_pngd_unfilter_t<1, false><<<grid, 256, 0, s>>>(stream, plan.width, plan.height, dst, progress, status);
    }
    e = cudaMemcpyAsync(status_host + 3, status + 3, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    ++nl;
    if (!direct) {
        const unsigned long long npix = (unsigned long long)plan.width * plan.height;
        k_pngd_expand<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(rows, npix, plan.color_type, plan.channels, plan.out_channels, palette_dev, out_dev);
        ++nl;
    }
    if (launches) *launches = nl;
    return cudaGetLastError();
}

}  // namespace lutgen
