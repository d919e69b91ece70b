// png_decode_core.h -- PNG decoding (the other half: `load_image`, crates/cli/src/main.rs:602-613,
// i.e. image::open(path) -> to_rgba8() / to_rgb8() for the files `lutgen apply` reads: the image and the --hald-clut).
//
//   container : signature, IHDR, PLTE, tRNS, IDAT..., IEND parsed on the host (png_parse): a few dozen bytes of
//               headers; the IDAT payloads are handed to the device as one zlib stream. 8 bits per sample, colour
//               types 0 / 2 / 3 / 4 / 6, no interlace; anything else is reported as unsupported so that the caller
//               can keep its own decoder for it (16-bit and interlaced PNGs do not occur among CLUTs and are rare
//               among photographs).
//   inflate   : RFC 1951, all three block types, canonical Huffman codes decoded length by length (inflate_stream:
//               no lookup tables to build, ~100 bytes of state -- the decoder is one thread). A deflate stream is
//               one serial chain of bits, so a foreign PNG is inflated by ONE thread (k_png_inflate_serial). Files
//               written by THIS library are different: every 16 KB piece of the scanline stream is its own
//               byte-aligned block sequence in its own IDAT chunk (png_core.h), so one thread per piece inflates them
//               all at once (k_png_inflate_pieces) -- the `--hald-clut` a previous `lutgen generate` saved decodes in
//               parallel. Whether a file is of that kind is decided by trying: the chunk pattern must fit, every
//               piece must consume exactly its chunk and produce exactly its 16 KB, and the Adler-32 of the whole must
//               be the stream's; otherwise the serial path runs.
//   unfilter  : PNG specification section 9. Sub / Up / Average / Paeth make a pixel depend on its left, upper and
//               upper-left neighbours: an anti-diagonal wavefront. A warp takes 32 consecutive rows, lane i row i, lane i
//               one pixel behind lane i-1, so the row above arrives by shuffle; bands of 32 rows follow each other
//               through a progress counter per band (k_png_unfilter).
//   expand    : to the channel count asked for (image::DynamicImage::to_rgb8 / to_rgba8): palette lookup, grey to
//               RGB, alpha added (255, or the tRNS entry of a palette colour) or dropped.
//
// Like png_core.h this header is plain integer C++ that also compiles for the host: tests/png_dec_sim.cpp builds the
// same inflate / unfilter code with g++ so that the CPU tests exercise what the GPU runs (zlib and PIL are the
// checkers). The host build is TEST infrastructure only; the product path (lutgen_cuda_decode_png*) has no CPU route.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define LPNGD_HD __host__ __device__
#else
#define LPNGD_HD
#endif

namespace lpngd {

constexpr int PIECE = 16384;   // lpng::CHUNK: scanline-stream bytes per piece of this library's own files
constexpr uint32_t ADLER_MOD = 65521u;

enum : int {
    OK = 0,
    ERR_TRUNCATED = -1,      // ran out of input
    ERR_BAD_BLOCK = -2,      // reserved block type, LEN != ~NLEN
    ERR_BAD_CODE = -3,       // over-subscribed / incomplete code, invalid symbol, repeat without a previous length
    ERR_BAD_DISTANCE = -4,   // distance reaches before the start of the output
    ERR_OVERFLOW = -5,       // more output than the caller said there would be
    ERR_LENGTH = -6,         // piece mode: did not end where it had to
};

// ---------------------------------------------------------------------------------------------
// Bit reader, least significant bit first (RFC 1951, 3.1.1). The buffer is refilled ahead of use (a table lookup
// peeks at bits it may not consume), zeros past the end; whether more bits were CONSUMED than exist is asked
// afterwards (br_overrun).
// ---------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p;
    size_t n, pos;
    uint32_t buf;
    int cnt;
};

LPNGD_HD inline void br_init(BitReader& r, const uint8_t* p, size_t n) {
    r.p = p; r.n = n; r.pos = 0; r.buf = 0; r.cnt = 0;
}
LPNGD_HD inline void br_fill(BitReader& r) {   // at least 25 bits afterwards
    while (r.cnt <= 24) {
        const uint32_t b = r.pos < r.n ? r.p[r.pos] : 0u;
        ++r.pos;
        r.buf |= b << r.cnt;
        r.cnt += 8;
    }
}
LPNGD_HD inline uint32_t br_bits(BitReader& r, int need) {   // need <= 16
    br_fill(r);
    const uint32_t v = r.buf & ((1u << need) - 1u);
    r.buf >>= need;
    r.cnt -= need;
    return v;
}
LPNGD_HD inline bool br_overrun(const BitReader& r) { return r.pos * 8 > r.n * 8 + (size_t)r.cnt; }
// Drops the rest of the current byte and gives the buffered whole bytes back: the next read is r.p[r.pos].
LPNGD_HD inline void br_align(BitReader& r) {
    r.pos -= (size_t)(r.cnt / 8);
    r.buf = 0;
    r.cnt = 0;
}
LPNGD_HD inline bool br_at_end(const BitReader& r) {   // nothing but (at most seven) zero padding bits left
    const size_t consumed = r.pos * 8 - (size_t)r.cnt;
    if (consumed >= r.n * 8) return true;
    if (r.n * 8 - consumed >= 8) return false;
    return (r.buf & ((1u << (r.n * 8 - consumed)) - 1u)) == 0;
}

// ---------------------------------------------------------------------------------------------
// Canonical Huffman codes (RFC 1951, 3.2.2). count[l] codes of length l, their symbols in `symbol` ordered by (length,
// symbol value): the length-by-length decoder needs nothing else. In front of it a lookup table over the next
// FAST_BITS stream bits resolves every code of at most that length in one step (entry = symbol << 4 | length;
// 0 = a longer code: decode length by length). Stream bits are the code's bits first-to-last from the least
// significant bit, so the table is indexed by the bit-reversed code.
// ---------------------------------------------------------------------------------------------
constexpr int FAST_BITS = 9;

struct Huffman {
    uint16_t count[16];
    uint16_t* symbol;
    uint16_t* fast;       // 1 << fast_bits entries
    int fast_bits;
};

// Returns 0 for a complete code, > 0 for an incomplete one (allowed for a single distance code), < 0 if over-subscribed.
LPNGD_HD inline int huff_build(Huffman& h, const uint8_t* length, int n) {
    for (int l = 0; l < 16; ++l) h.count[l] = 0;
    for (int s = 0; s < n; ++s) h.count[length[s]]++;
    for (int i = 0; i < (1 << h.fast_bits); ++i) h.fast[i] = 0;
    if (h.count[0] == n) return 0;   // no codes at all: nothing can be decoded (caught on use)
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= h.count[l];
        if (left < 0) return left;
    }
    uint16_t offs[16], next_code[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
    uint32_t code = 0;
    for (int l = 1; l < 16; ++l) {
        code = (code + (l > 1 ? h.count[l - 1] : 0)) << 1;
        next_code[l] = (uint16_t)code;
    }
    for (int s = 0; s < n; ++s) {
        const int l = length[s];
        if (l == 0) continue;
        h.symbol[offs[l]++] = (uint16_t)s;
        const uint32_t c = next_code[l]++;
        if (l <= h.fast_bits) {
            uint32_t rev = 0;
            for (int i = 0; i < l; ++i) rev |= ((c >> i) & 1u) << (l - 1 - i);
            for (uint32_t k = rev; k < (1u << h.fast_bits); k += 1u << l) h.fast[k] = (uint16_t)((s << 4) | l);
        }
    }
    return left;
}

LPNGD_HD inline int huff_decode(BitReader& r, const Huffman& h) {
    br_fill(r);
    const uint32_t e = h.fast[r.buf & ((1u << h.fast_bits) - 1u)];
    if (e) {
        r.buf >>= (e & 15u);
        r.cnt -= (int)(e & 15u);
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int)(r.buf & 1u);
        r.buf >>= 1;
        --r.cnt;
        const int count = h.count[l];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return ERR_BAD_CODE;
}

// ---------------------------------------------------------------------------------------------
// Inflate. Decodes blocks from `r` into dst[*produced ..), at most `cap` bytes in all, until a final block has been
// decoded (stop_at_final) or the input is used up exactly at a block boundary (piece mode: !stop_at_final).
// Matches may reach back to dst[0].
// ---------------------------------------------------------------------------------------------
struct InflateScratch {   // ~2.7 KB: per decoding thread (local memory on the device)
    uint8_t lengths[19 + 288 + 32 + 5];
    uint16_t lsym[288];
    uint16_t dsym[32];
    uint16_t lfast[1 << FAST_BITS];
    uint16_t dfast[1 << 7];
};

// Where the inflated bytes go. DirectSink: a buffer that is also the history (matches read it back). WindowSink: the
// buffer is slow to read back (global memory: a match of distance 3 is a chain of dependent L2 round trips), so the last
// 32 KB -- as far as a deflate match can reach -- are kept in a second, fast buffer (shared memory) that the matches read.
struct DirectSink {
    uint8_t* dst;
    size_t cap, out;
    LPNGD_HD bool room(size_t n) const { return out + n <= cap; }
    LPNGD_HD void put(uint8_t v) { dst[out++] = v; }
    LPNGD_HD bool reaches(size_t dist) const { return dist <= out; }
    LPNGD_HD void copy(size_t dist, uint32_t len) {
        for (uint32_t i = 0; i < len; ++i, ++out) dst[out] = dst[out - dist];
    }
};
struct WindowSink {
    uint8_t* dst;
    uint8_t* win;   // 32768 bytes
    size_t cap, out;
    LPNGD_HD bool room(size_t n) const { return out + n <= cap; }
    LPNGD_HD void put(uint8_t v) {
        win[out & 32767u] = v;
        dst[out++] = v;
    }
    LPNGD_HD bool reaches(size_t dist) const { return dist <= out; }
    LPNGD_HD void copy(size_t dist, uint32_t len) {
        for (uint32_t i = 0; i < len; ++i) put(win[(out - dist) & 32767u]);
    }
};

template <class Sink>
LPNGD_HD inline int inflate_blocks(BitReader& r, Sink& sink, bool stop_at_final, InflateScratch& X) {
    const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    Huffman lcode, dcode;
    lcode.symbol = X.lsym; lcode.fast = X.lfast; lcode.fast_bits = FAST_BITS;
    dcode.symbol = X.dsym; dcode.fast = X.dfast; dcode.fast_bits = 7;
    for (;;) {
        if (!stop_at_final && br_at_end(r)) break;   // piece mode: the chunk is used up
        const uint32_t final_block = br_bits(r, 1);
        const uint32_t type = br_bits(r, 2);
        if (br_overrun(r)) return ERR_TRUNCATED;
        if (type == 0) {
            br_align(r);
            if (r.pos + 4 > r.n) return ERR_TRUNCATED;
            const uint32_t len = r.p[r.pos] | ((uint32_t)r.p[r.pos + 1] << 8), nlen = r.p[r.pos + 2] | ((uint32_t)r.p[r.pos + 3] << 8);
            r.pos += 4;
            if ((len ^ 0xffffu) != nlen) return ERR_BAD_BLOCK;
            if (r.pos + len > r.n) return ERR_TRUNCATED;
            if (!sink.room(len)) return ERR_OVERFLOW;
            for (uint32_t i = 0; i < len; ++i) sink.put(r.p[r.pos + i]);
            r.pos += len;
        } else if (type == 1 || type == 2) {
            if (type == 1) {
                for (int s = 0; s < 144; ++s) X.lengths[s] = 8;
                for (int s = 144; s < 256; ++s) X.lengths[s] = 9;
                for (int s = 256; s < 280; ++s) X.lengths[s] = 7;
                for (int s = 280; s < 288; ++s) X.lengths[s] = 8;
                huff_build(lcode, X.lengths, 288);
                for (int s = 0; s < 30; ++s) X.lengths[s] = 5;
                huff_build(dcode, X.lengths, 30);
            } else {
                const int nlen = (int)br_bits(r, 5) + 257, ndist = (int)br_bits(r, 5) + 1, ncode = (int)br_bits(r, 4) + 4;
                if (br_overrun(r)) return ERR_TRUNCATED;
                if (nlen > 286 || ndist > 30) return ERR_BAD_CODE;
                for (int i = 0; i < 19; ++i) X.lengths[i] = 0;
                for (int i = 0; i < ncode; ++i) X.lengths[clorder[i]] = (uint8_t)br_bits(r, 3);
                // the code length code is decoded through the distance code's tables (free at this point)
                if (huff_build(dcode, X.lengths, 19) != 0) return ERR_BAD_CODE;   // it must be complete
                int idx = 0;
                while (idx < nlen + ndist) {
                    int sym = huff_decode(r, dcode);
                    if (sym < 0) return sym;
                    if (br_overrun(r)) return ERR_TRUNCATED;
                    if (sym < 16) {
                        X.lengths[19 + idx++] = (uint8_t)sym;
                    } else {
                        int prev = 0, rep;
                        if (sym == 16) {
                            if (idx == 0) return ERR_BAD_CODE;
                            prev = X.lengths[19 + idx - 1];
                            rep = 3 + (int)br_bits(r, 2);
                        } else if (sym == 17) {
                            rep = 3 + (int)br_bits(r, 3);
                        } else {
                            rep = 11 + (int)br_bits(r, 7);
                        }
                        if (idx + rep > nlen + ndist) return ERR_BAD_CODE;
                        while (rep--) X.lengths[19 + idx++] = (uint8_t)prev;
                    }
                }
                if (br_overrun(r)) return ERR_TRUNCATED;
                const uint8_t* ll = X.lengths + 19;
                if (ll[256] == 0) return ERR_BAD_CODE;   // no end-of-block code
                int e = huff_build(lcode, ll, nlen);
                if (e < 0 || (e > 0 && nlen - lcode.count[0] != 1)) return ERR_BAD_CODE;
                e = huff_build(dcode, ll + nlen, ndist);
                if (e < 0 || (e > 0 && ndist - dcode.count[0] != 1)) return ERR_BAD_CODE;
            }
            for (;;) {
                int sym = huff_decode(r, lcode);
                if (sym < 0) return sym;
                if (sym < 256) {
                    if (!sink.room(1)) return ERR_OVERFLOW;
                    sink.put((uint8_t)sym);
                } else if (sym == 256) {
                    break;
                } else {
                    sym -= 257;
                    if (sym >= 29) return ERR_BAD_CODE;
                    const uint32_t len = lbase[sym] + br_bits(r, lext[sym]);
                    int ds = huff_decode(r, dcode);
                    if (ds < 0) return ds;
                    if (ds >= 30) return ERR_BAD_CODE;
                    const size_t dist = (size_t)dbase[ds] + br_bits(r, dext[ds]);
                    if (br_overrun(r)) return ERR_TRUNCATED;
                    if (!sink.reaches(dist)) return ERR_BAD_DISTANCE;
                    if (!sink.room(len)) return ERR_OVERFLOW;
                    sink.copy(dist, len);
                }
                if (br_overrun(r)) return ERR_TRUNCATED;
            }
        } else {
            return ERR_BAD_BLOCK;
        }
        if (final_block) break;   // with or without stop_at_final: the last piece of a stream ends with its final block
    }
    return OK;
}

// A whole zlib stream (RFC 1950: two header bytes, deflate data, Adler-32 -- the Adler-32 is checked separately).
// `window`: 32768 bytes of fast memory for the match history, or null (dst is read back directly).
LPNGD_HD inline int inflate_zlib(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* produced, InflateScratch& X, uint8_t* window = nullptr) {
    if (n < 2) return ERR_TRUNCATED;
    if ((src[0] & 0x0f) != 8 || (((uint32_t)src[0] << 8) | src[1]) % 31u != 0 || (src[1] & 0x20)) return ERR_BAD_BLOCK;
    BitReader r;
    br_init(r, src + 2, n - 2);
    int rc;
    if (window) {
        WindowSink sink{dst, window, cap, 0};
        rc = inflate_blocks(r, sink, true, X);
        *produced = sink.out;
    } else {
        DirectSink sink{dst, cap, 0};
        rc = inflate_blocks(r, sink, true, X);
        *produced = sink.out;
    }
    return rc;
}

// One piece of a file this library wrote: the blocks in src[0, n) must produce exactly `expect` bytes at dst[0 ..) and use
// the chunk up. (Matches never reach before the piece: png_core.h's runs start inside it.)
LPNGD_HD inline int inflate_piece(const uint8_t* src, size_t n, uint8_t* dst, size_t expect, InflateScratch& X) {
    BitReader r;
    br_init(r, src, n);
    DirectSink sink{dst, expect, 0};
    const int rc = inflate_blocks(r, sink, false, X);
    if (rc != OK) return rc;
    if (sink.out != expect) return ERR_LENGTH;
    return OK;
}

// ---------------------------------------------------------------------------------------------
// Adler-32 (RFC 1950): of a span, and the combination of two spans' sums.
// ---------------------------------------------------------------------------------------------
LPNGD_HD inline void adler_span(const uint8_t* p, size_t n, uint32_t& a, uint32_t& b) {   // a, b: sums of the span alone (a starts at 0)
    uint32_t sa = 0, sb = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) { sa += *p++; sb += sa; }
        sa %= ADLER_MOD;
        sb %= ADLER_MOD;
    }
    a = sa;
    b = sb;
}
// Running (A, B) (A starts at 1) extended by a span of `n` bytes whose own sums are (a, b).
LPNGD_HD inline void adler_append(uint32_t& A, uint32_t& B, uint32_t a, uint32_t b, size_t n) {
    B = (uint32_t)((B + (unsigned long long)(n % ADLER_MOD) * A + b) % ADLER_MOD);
    A = (A + a) % ADLER_MOD;
}

// ---------------------------------------------------------------------------------------------
// Scanline reconstruction (PNG specification 9.2), one pixel of `bpp` bytes packed into a word.
// ---------------------------------------------------------------------------------------------
LPNGD_HD inline uint32_t paeth_byte(uint32_t a, uint32_t b, uint32_t c) {
    const int p = (int)a + (int)b - (int)c;
    const int pa = p > (int)a ? p - (int)a : (int)a - p, pb = p > (int)b ? p - (int)b : (int)b - p, pc = p > (int)c ? p - (int)c : (int)c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// x: the filtered bytes of the pixel; a, b, c: the reconstructed left, upper and upper-left pixels (0 outside the image).
LPNGD_HD inline uint32_t unfilter_pixel(uint32_t ft, uint32_t x, uint32_t a, uint32_t b, uint32_t c, uint32_t bpp) {
    uint32_t out = 0;
    for (uint32_t k = 0; k < bpp; ++k) {
        const uint32_t sh = 8u * k, xv = (x >> sh) & 255u, av = (a >> sh) & 255u, bv = (b >> sh) & 255u, cv = (c >> sh) & 255u;
        uint32_t pred = 0;
        if (ft == 1) pred = av;
        else if (ft == 2) pred = bv;
        else if (ft == 3) pred = (av + bv) >> 1;
        else if (ft == 4) pred = paeth_byte(av, bv, cv);
        out |= ((xv + pred) & 255u) << sh;
    }
    return out;
}

// The same, branch free and with the channel loop unrolled (the wavefront kernel's step: every lane of a warp has its
// own row, hence its own filter type; branching on it serialised the five flavours and made a step 1400 cycles long).
// mA / mB / mV / mP: 0xff for the row's filter type (Sub / Up / Average / Paeth), else 0 -- at most one of them is set.
template <int BPP>
LPNGD_HD inline uint32_t unfilter_pixel_masked(uint32_t x, uint32_t a, uint32_t b, uint32_t c, uint32_t mA, uint32_t mB, uint32_t mV, uint32_t mP) {
    uint32_t out = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < BPP; ++k) {
        const int sh = 8 * k;
        const int xv = (int)((x >> sh) & 255u), av = (int)((a >> sh) & 255u), bv = (int)((b >> sh) & 255u), cv = (int)((c >> sh) & 255u);
        const int pa = bv > cv ? bv - cv : cv - bv;             // |p - a| = |b - c|
        const int pb = av > cv ? av - cv : cv - av;             // |p - b| = |a - c|
        const int t = av + bv - 2 * cv;
        const int pc = t < 0 ? -t : t;                          // |p - c|
        const int paeth = (pa <= pb && pa <= pc) ? av : (pb <= pc ? bv : cv);
        const uint32_t pred = ((uint32_t)av & mA) | ((uint32_t)bv & mB) | ((uint32_t)((av + bv) >> 1) & mV) | ((uint32_t)paeth & mP);
        out |= (((uint32_t)xv + pred) & 255u) << sh;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Container (host only).
// ---------------------------------------------------------------------------------------------
struct Info {
    uint32_t width = 0, height = 0;
    int color_type = 0;     // 0 grey, 2 RGB, 3 palette, 4 grey + alpha, 6 RGBA
    int channels = 0;       // samples per pixel in the file: 1, 3, 1, 2, 4
    int out_channels = 0;   // what to_rgb8 / to_rgba8 would be asked for by default: 3, or 4 when the file has transparency
    uint32_t palette[256];  // r | g << 8 | b << 16 | a << 24
    int npalette = 0;
    bool has_trns = false;
};

enum : int { P_OK = 0, P_NOT_PNG = 1, P_CORRUPT = 2, P_UNSUPPORTED = 3 };

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline uint32_t crc32_bytes(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool made = false;
    if (!made) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        made = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return ~crc;
}

// Walks the chunks; idat receives (offset, length) of every IDAT payload in file order. Every chunk's CRC is checked.
template <class Vec>
inline int png_parse(const uint8_t* f, size_t n, Info& info, Vec& idat) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8) return P_NOT_PNG;
    for (int i = 0; i < 8; ++i)
        if (f[i] != sig[i]) return P_NOT_PNG;
    size_t pos = 8;
    bool have_ihdr = false, have_iend = false;
    int bit_depth = 0, interlace = 0;
    for (int i = 0; i < 256; ++i) info.palette[i] = 0xff000000u;
    while (pos + 12 <= n && !have_iend) {
        const uint32_t len = be32(f + pos);
        if ((size_t)len > n - pos - 12) return P_CORRUPT;
        const uint8_t* type = f + pos + 4;
        const uint8_t* data = f + pos + 8;
        if (crc32_bytes(type, (size_t)len + 4) != be32(data + len)) return P_CORRUPT;
        const uint32_t tag = be32(type);
        if (tag == 0x49484452u) {   // IHDR
            if (len != 13 || have_ihdr) return P_CORRUPT;
            info.width = be32(data);
            info.height = be32(data + 4);
            bit_depth = data[8];
            info.color_type = data[9];
            interlace = data[12];
            if (data[10] != 0 || data[11] != 0) return P_CORRUPT;
            have_ihdr = true;
        } else if (!have_ihdr) {
            return P_CORRUPT;
        } else if (tag == 0x504C5445u) {   // PLTE
            if (len % 3 != 0 || len > 768) return P_CORRUPT;
            info.npalette = (int)(len / 3);
            for (int i = 0; i < info.npalette; ++i)
                info.palette[i] = (info.palette[i] & 0xff000000u) | data[3 * i] | ((uint32_t)data[3 * i + 1] << 8) | ((uint32_t)data[3 * i + 2] << 16);
        } else if (tag == 0x74524E53u) {   // tRNS
            if (info.color_type == 3) {
                if (len > 256) return P_CORRUPT;
                for (uint32_t i = 0; i < len; ++i) info.palette[i] = (info.palette[i] & 0x00ffffffu) | ((uint32_t)data[i] << 24);
                info.has_trns = true;
            } else if (info.color_type == 0 || info.color_type == 2) {
                return P_UNSUPPORTED;   // a colour key: image::to_rgba8 turns it into alpha; not done here
            }
        } else if (tag == 0x49444154u) {   // IDAT
            idat.push_back({pos + 8, (size_t)len});
        } else if (tag == 0x49454E44u) {   // IEND
            have_iend = true;
        } else if (!(type[0] & 0x20)) {
            return P_UNSUPPORTED;   // an unknown critical chunk
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !have_iend || idat.empty()) return P_CORRUPT;
    if (info.width == 0 || info.height == 0) return P_CORRUPT;
    if (bit_depth != 8 || interlace != 0) return P_UNSUPPORTED;
    switch (info.color_type) {
        case 0: info.channels = 1; info.out_channels = 3; break;
        case 2: info.channels = 3; info.out_channels = 3; break;
        case 3: info.channels = 1; info.out_channels = info.has_trns ? 4 : 3; if (info.npalette == 0) return P_CORRUPT; break;
        case 4: info.channels = 2; info.out_channels = 4; break;
        case 6: info.channels = 4; info.out_channels = 4; break;
        default: return P_CORRUPT;
    }
    return P_OK;
}
}  // namespace lpngd
