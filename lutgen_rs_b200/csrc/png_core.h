// png_core.h -- PNG encoding of an RGB8/RGBA8 image as a data-parallel pipeline (SURVEY 8(f) row 2:
// the `lut.save(path)` / `image.save(path)` calls of crates/cli/src/main.rs:767, 811, 839, 862, which
// after the kernels are where `lutgen generate` / `lutgen apply` spend their wall time).
//
// What is produced is an ordinary PNG: signature, IHDR, IDAT..., IEND; 8 bits per sample, no
// interlace, per-row adaptive filter (minimum sum of absolute residuals, the heuristic the `png`
// crate's Adaptive mode uses), one zlib stream. The zlib stream is cut into CHUNK-byte pieces of the
// filtered scanline stream; every piece is compressed independently by one CTA into one deflate block
// with its own dynamic Huffman code and ends on a byte boundary (an empty stored block, the same
// device `pigz` uses), so pieces concatenate without bit shifting and each travels in its own IDAT
// chunk with its own CRC. Matches are restricted to distance = bytes-per-pixel (runs of identical
// filtered pixels, which is what filtered CLUTs and photographs mostly offer; the `image` crate's
// default "fast" mode is of the same kind), which makes the tokenisation a function of run boundaries
// alone -- no sequential dependency, every thread tokenises its own 64 bytes.
//
// Everything in this header is plain integer C++ that compiles for the host as well: the per-thread
// phases are written against an executor (`par(f)`: run f(t) for every thread t of the CTA, then
// barrier), the CUDA kernels instantiate it with threadIdx/__syncthreads, and tests/png_sim.cpp
// instantiates it with a loop, so the unit tests here exercise the very code the GPU runs (there is
// no GPU in the development container). The host build is TEST infrastructure only; the product path
// (lutgen_cuda_encode_png*) has no CPU route.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LPNG_HD __host__ __device__ __forceinline__
#define LPNG_HDN __host__ __device__ inline
#else
#define LPNG_HD inline
#define LPNG_HDN inline
#endif

namespace lpng {

constexpr int SEG = 64;                  // stream bytes tokenised by one thread
constexpr int THREADS = 256;             // threads per CTA
constexpr int CHUNK = SEG * THREADS;     // stream bytes per deflate block
constexpr int MIN_RUN = 4;               // shortest run worth a match (classify() hard-wires it in its shifts)
constexpr int MAX_MATCH = 258;
constexpr int NLITLEN = 286;             // literal/length symbols in use
constexpr int NSYM = 288;
constexpr int EOB = 256;
constexpr int MAX_BITS = 15;
constexpr int OUT_BYTES = CHUNK + 32;    // zlib header + stored-block header + CHUNK, rounded up
constexpr int OUT_WORDS = OUT_BYTES / 4;
constexpr int F_BYTES = CHUNK + (CHUNK >> 6) * 4;  // one pad word per 64 bytes: thread stride 17 words
constexpr int CRC_SLICE = 68;            // bytes per thread in the CRC pass (17 words: conflict free)
constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr uint32_t ADLER_MOD = 65521u;
constexpr int BIG = 1 << 30;
constexpr int HIST_REPLICAS = 8;         // histogram copies (in the not yet used output buffer) against same-symbol contention
constexpr int HIST_PITCH = NSYM + 1;     // odd: the copies of one symbol sit in different banks
constexpr int PNG_HEADER_BYTES = 33;     // signature + IHDR chunk
constexpr int PNG_TRAILER_BYTES = 16 + 12;  // adler IDAT + IEND

static_assert(HIST_REPLICAS * HIST_PITCH <= OUT_WORDS, "histogram copies must fit in the output buffer");

LPNG_HD int pad(int o) { return o + ((o >> 6) << 2); }

LPNG_HD int clz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
LPNG_HD int ctz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
LPNG_HD int ctz32(uint32_t v) {  // v != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
LPNG_HD void atomic_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
LPNG_HD uint32_t atomic_inc(uint32_t* p) {  // returns the old value
#ifdef __CUDA_ARCH__
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}
LPNG_HD void atomic_max(int* p, int v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LPNG_HD void atomic_or(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// ---------------------------------------------------------------------------------------------
// CRC-32 (PNG / zlib polynomial, reflected). Polynomials over GF(2) are held bit-reflected: bit 31
// is the coefficient of x^0. crc_update is linear in (state, byte), so the state after a message is
// the XOR of the states of its slices, each multiplied by x^(8 * bytes that follow it).
// ---------------------------------------------------------------------------------------------
LPNG_HD uint32_t gf_mul(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
    for (int i = 0; i < 32; ++i) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? CRC_POLY : 0u);
    }
    return p;
}

struct Tables {
    uint32_t crc[256];
    uint32_t xs[8];  // x^(8 * CRC_SLICE * 2^j) mod P
};

inline uint32_t gf_xpow(unsigned long long n) {  // x^n mod P (host: table construction)
    uint32_t r = 0x80000000u, b = 0x40000000u;
    while (n) {
        if (n & 1) r = gf_mul(r, b);
        b = gf_mul(b, b);
        n >>= 1;
    }
    return r;
}

inline void make_tables(Tables& T) {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? CRC_POLY : 0u);
        T.crc[i] = c;
    }
    for (int j = 0; j < 8; ++j) T.xs[j] = gf_xpow((unsigned long long)(8 * CRC_SLICE) << j);
}

LPNG_HD uint32_t crc_bitwise(uint32_t state, uint32_t byte) {  // table-free (a handful of header bytes)
    state ^= byte;
    for (int k = 0; k < 8; ++k) state = (state >> 1) ^ ((state & 1u) ? CRC_POLY : 0u);
    return state;
}

// ---------------------------------------------------------------------------------------------
// Scanline filters (PNG specification, section 9).
// ---------------------------------------------------------------------------------------------
struct Image {
    const uint8_t* px;        // height x width x bpp, tightly packed
    const uint8_t* filters;   // height filter types (0..4), chosen by choose_filter
    uint32_t width, height, bpp, rb;  // rb = width * bpp
    unsigned long long stream_bytes;  // height * (rb + 1)
};

LPNG_HD int paeth(int a, int b, int c) {
    int p = a + b - c;
    int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Bytes [xbase, xbase + ndata) of a row under filter FT into the padded stream buffer at offset obase
// (prow: the row above, or null); thread t of the CTA takes every THREADS-th byte.
template <int FT>
LPNG_HD void filter_span(uint8_t* fb, int obase, const uint8_t* row, const uint8_t* prow, uint32_t xbase, int ndata, uint32_t bpp, int t) {
    for (int i = t; i < ndata; i += THREADS) {
        const uint32_t x = xbase + (uint32_t)i;
        const int cur = row[x];
        int pred = 0;
        if (FT == 1) {
            pred = x >= bpp ? row[x - bpp] : 0;
        } else if (FT == 2) {
            pred = prow ? prow[x] : 0;
        } else if (FT == 3) {
            const int a = x >= bpp ? row[x - bpp] : 0, b = prow ? prow[x] : 0;
            pred = (a + b) >> 1;
        } else if (FT == 4) {
            const bool left = x >= bpp;
            const int a = left ? row[x - bpp] : 0, b = prow ? prow[x] : 0, c = (left && prow) ? prow[x - bpp] : 0;
            pred = paeth(a, b, c);
        }
        fb[pad(obase + i)] = (uint8_t)(cur - pred);
    }
}

// |signed byte| of a residual given mod 256.
LPNG_HD uint32_t abs_residual(int r) {
    const uint32_t u = (uint32_t)r & 255u;
    return u < 256u - u ? u : 256u - u;
}

// Adds, for bytes first, first + step, ... of a row, the absolute (signed-byte) residual under each of the
// five filters (prow: the row above, or null).
LPNG_HD void row_costs(const uint8_t* row, const uint8_t* prow, uint32_t rb, uint32_t bpp, uint32_t first, uint32_t step, uint32_t cost[5]) {
    for (uint32_t x = first; x < rb; x += step) {
        const bool left = x >= bpp;
        const int cur = row[x];
        const int a = left ? row[x - bpp] : 0;
        const int b = prow ? prow[x] : 0;
        const int c = (left && prow) ? prow[x - bpp] : 0;
        cost[0] += abs_residual(cur);
        cost[1] += abs_residual(cur - a);
        cost[2] += abs_residual(cur - b);
        cost[3] += abs_residual(cur - ((a + b) >> 1));
        cost[4] += abs_residual(cur - paeth(a, b, c));
    }
}

LPNG_HD int pick_filter(const uint32_t cost[5]) {
    int best = 0;
    for (int ft = 1; ft < 5; ++ft)
        if (cost[ft] < cost[best]) best = ft;
    return best;
}

// ---------------------------------------------------------------------------------------------
// Deflate pieces (RFC 1951).
// ---------------------------------------------------------------------------------------------
LPNG_HD int ilog2(uint32_t v) {  // v > 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

LPNG_HD int length_symbol(int len, int& extra_bits, int& extra) {  // len in 3..258
    extra_bits = 0;
    extra = 0;
    if (len == MAX_MATCH) return 285;
    int l = len - 3;
    if (l < 8) return 257 + l;
    extra_bits = ilog2((uint32_t)l) - 2;
    extra = l & ((1 << extra_bits) - 1);
    return 257 + 4 * extra_bits + 4 + ((l >> extra_bits) - 4);
}

LPNG_HD uint32_t bit_reverse(uint32_t v, int n) {  // the low n bits of v, reversed (n in 1..32)
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    v = (v >> 16) | (v << 16);
    return v >> (32 - n);
#endif
}

struct ChunkMeta {
    uint32_t n_out;    // compressed bytes of this piece (what its IDAT chunk carries)
    uint32_t crc;      // CRC-32 of "IDAT" + those bytes
    uint32_t adler_a;  // sum of the piece's stream bytes mod 65521
    uint32_t adler_b;  // sum of (m - i) * byte_i mod 65521
    uint32_t m;        // stream bytes in the piece
};

// CTA-shared state of one piece: 45,600 bytes. (Five CTAs per SM need at most 45,568: each CTA's allocation is
// this + 1 KB rounded up to 128 bytes, and 5 x 46,720 is 128 bytes more than the SM's 228 KB -- ncu reports a
// limit of four. Moving 32 more bytes into Scratch would tip it; not done, not measured.)
struct Shared {
    uint32_t f[F_BYTES / 4];   // filtered stream bytes, padded layout (pad())
    uint32_t out[OUT_WORDS];   // compressed piece, LSB-first bit stream; before that: histogram copies and Scratch
    union {
        uint32_t hist[NSYM];     // symbol frequencies, until the code lengths are known
        uint32_t codelen[NSYM];  // then: canonical code, bit reversed (ready for the LSB-first stream) | length << 16 | header position << 20
    };
    uint8_t len[NSYM];
    unsigned long long eq[THREADS];  // bit k: stream byte lo+k equals the byte bpp before it
    unsigned long long ev[THREADS];  // bit k: a token starts at stream byte lo+k
    unsigned long long ms[THREADS];  // ... and that token is a match
    int sa[THREADS], sb[THREADS];    // run-boundary scans (kept until the last sweep)
    int bits[THREADS];               // Adler partial sums (b); later token bits per thread and their inclusive prefix sums
    uint32_t crc[THREADS];           // Adler partial sums (a); later scan scratch, then CRC partial states
    uint32_t nused;
    int nlit, stored, n_out;
    uint32_t header_bits, data_bits, chunk_a, chunk_b;
};

// What is only needed before the bit stream is written lives in the upper part of the output buffer
// (the histogram copies take the lower part); the buffer is cleared again before emission.
struct Scratch {
    int ta[THREADS], tb[THREADS];   // scan double buffers
    uint32_t node_freq[2 * NSYM];   // Huffman tree: frequencies, then depths
    uint16_t node_parent[2 * NSYM];
    uint16_t used[NSYM];            // used symbols, any order
    uint16_t order[NSYM];           // used symbols by ascending (frequency, symbol)
    uint32_t next_code[MAX_BITS + 1];  // first canonical code of each length (read by assign_codes, before the buffer is cleared)
};
static_assert(sizeof(uint32_t) * HIST_REPLICAS * HIST_PITCH + sizeof(Scratch) <= OUT_BYTES, "histogram copies + scratch must fit in the output buffer");
// five CTAs per SM: 5 x (round_up(sizeof(Shared), 128) + 1024 reserved per CTA) <= 228 KB  <=>  sizeof(Shared) <= 45568
static_assert(sizeof(Shared) <= 45568, "keep the shared state at the five-CTAs-per-SM size");

LPNG_HD Scratch& scratch(Shared& S) { return *reinterpret_cast<Scratch*>(S.out + HIST_REPLICAS * HIST_PITCH); }

// out |= v << bitpos (v at most 32 bits wide)
LPNG_HD void or_bits(uint32_t* out, uint32_t bitpos, uint32_t v) {
    uint32_t w = bitpos >> 5, sh = bitpos & 31u;
    atomic_or(&out[w], v << sh);
    if (sh && (v >> (32u - sh))) atomic_or(&out[w + 1], v >> (32u - sh));
}

struct HistSink {
    uint32_t* hist;
    LPNG_HD void lit(int b) { atomic_add(&hist[b], 1u); }
    LPNG_HD void match(int len) {
        int eb, ex;
        atomic_add(&hist[length_symbol(len, eb, ex)], 1u);
    }
};

struct CountSink {
    const uint8_t* len;
    uint32_t bits;
    LPNG_HD void lit(int b) { bits += len[b]; }
    LPNG_HD void match(int l) {
        int eb, ex;
        int s = length_symbol(l, eb, ex);
        bits += len[s] + eb + 1;
    }
};

// Appends to the bit stream from a given bit offset: the first and the last word a thread touches
// are shared with its neighbours (atomic OR into the zeroed buffer), the words between are its own.
struct EmitSink {
    uint32_t* out;
    const uint32_t* codelen;
    uint32_t acc;
    int nacc;
    uint32_t w;
    bool first;
    uint32_t dist_bit;
    LPNG_HD void begin(uint32_t bitpos) {
        acc = 0;
        nacc = (int)(bitpos & 31u);
        w = bitpos >> 5;
        first = true;
    }
    LPNG_HD void put(uint32_t v, int n) {  // the low n bits of v (n in 1..32, nothing above them)
        acc |= v << nacc;
        nacc += n;
        if (nacc >= 32) {
            if (first) {
                atomic_or(&out[w], acc);
                first = false;
            } else {
                out[w] = acc;
            }
            ++w;
            nacc -= 32;
            acc = nacc ? v >> (n - nacc) : 0u;
        }
    }
    LPNG_HD void end() {
        if (nacc > 0 && acc) atomic_or(&out[w], acc);
    }
    LPNG_HD void lit(int b) {
        const uint32_t cl = codelen[b];
        put(cl & 0xFFFFu, (int)((cl >> 16) & 15u));
    }
    LPNG_HD void match(int l) {  // length code, its extra bits and the one-bit distance code in one go (at most 21 bits)
        int eb, ex;
        const uint32_t cl = codelen[length_symbol(l, eb, ex)];
        const int n0 = (int)((cl >> 16) & 15u);
        put((cl & 0xFFFFu) | ((uint32_t)ex << n0) | (dist_bit << (n0 + eb)), n0 + eb + 1);
    }
};

// Tokens that START in thread t's 64 bytes. A maximal interval [s, e) of positions whose byte equals
// the byte bpp before it is a run; a run of at least MIN_RUN bytes is cut into matches of 258 bytes
// from its start (a tail shorter than 3 bytes goes out as literals), everything else is a literal.
// s and e come from the thread's own bit mask or, when the run crosses its segment's ends, from the
// scanned neighbours' masks (sa / sb), so no thread waits for another.
struct RunEnds {
    int s, e;
};

// The run that contains position lo+k (bit k of eq set).
LPNG_HD RunEnds run_of(const Shared& S, int t, int m, int k) {
    const int lo = t * SEG;
    int n = m - lo;
    if (n > SEG) n = SEG;
    const unsigned long long valid = n == 64 ? ~0ull : ((1ull << n) - 1ull);
    const unsigned long long zeros = ~S.eq[t] & valid;
    const unsigned long long below = zeros & ((1ull << k) - 1ull);
    const unsigned long long above = zeros >> k;
    RunEnds r;
    r.s = below ? lo + 64 - clz64(below) : (t > 0 ? S.sa[t - 1] : -1) + 1;
    if (above) {
        r.e = lo + k + ctz64(above);
    } else {
        int next0 = t + 1 < THREADS ? S.sb[t + 1] : BIG;
        r.e = next0 > m ? m : next0;
    }
    return r;
}

// True if the run containing position lo+k (bit k of eq set) starts exactly there.
LPNG_HD bool run_starts_at(const Shared& S, int t, int k) {
    return k > 0 ? !((S.eq[t] >> (k - 1)) & 1ull) : (t == 0 || !(S.eq[t - 1] >> 63));
}

// Where the matches of run [s, e) end: its last 1 or 2 bytes, if 258 does not divide the rest, are literals.
LPNG_HD int match_end(int s, int e) {
    const int rem = (e - s) % MAX_MATCH;
    return e - (rem < 3 ? rem : 0);
}

// S.ev[t] / S.ms[t]: which of thread t's 64 positions start a token, and which of those start a match.
// Positions in runs of at least MIN_RUN bytes are found with shifts of the thread's mask widened by
// three positions on either side (whether a position's run has four bytes depends on nothing farther);
// what is left is a literal. Each long run then contributes at most one match start to these 64 bytes.
LPNG_HD void classify(Shared& S, int t, int m) {
    const int lo = t * SEG;
    int n = m - lo;
    if (n <= 0) {
        S.ev[t] = 0;
        S.ms[t] = 0;
        return;
    }
    if (n > SEG) n = SEG;
    const unsigned long long E = S.eq[t];
    const unsigned long long valid = n == 64 ? ~0ull : ((1ull << n) - 1ull);
    const unsigned long long P3 = t > 0 ? S.eq[t - 1] >> 61 : 0ull;
    const unsigned long long N3 = t + 1 < THREADS ? (S.eq[t + 1] & 7ull) : 0ull;
    // window: bit i <-> position lo - 3 + i, 70 bits in (wl, wh)
    unsigned long long wl = P3 | (E << 3), wh = (E >> 61) | (N3 << 3);
    static_assert(MIN_RUN == 4, "the shifts below find runs of exactly four or more");
    // pairs, then fours: bit i set if positions i..i+3 are all in runs
    unsigned long long bl = wl & ((wl >> 1) | (wh << 63)), bh = wh & (wh >> 1);
    unsigned long long al = bl & ((bl >> 2) | (bh << 62)), ah = bh & (bh >> 2);
    // spread each four back over its positions
    unsigned long long cl = al | (al << 1), ch = ah | (ah << 1) | (al >> 63);
    unsigned long long ll = cl | (cl << 2), lh = ch | (ch << 2) | (cl >> 62);
    const unsigned long long longrun = ((ll >> 3) | (lh << 61)) & valid;
    unsigned long long ev = valid & ~longrun, ms = 0;
    unsigned long long rest = longrun;
    while (rest) {
        const int a = ctz64(rest);
        const unsigned long long inv = ~longrun >> a;  // bit 0 is clear
        const int b = inv ? a + ctz64(inv) : 64;
        if (b < 64 && run_starts_at(S, t, a)) {  // the run lies inside these 64 bytes: one match from its start
            ev |= 1ull << a;
            ms |= 1ull << a;
        } else {
            const RunEnds r = run_of(S, t, m, a);
            const int mend = match_end(r.s, r.e);
            const int pos = lo + a, stop = lo + b;
            const int p = r.s + ((pos - r.s + MAX_MATCH - 1) / MAX_MATCH) * MAX_MATCH;
            if (p < stop && p < mend) {
                ev |= 1ull << (p - lo);
                ms |= 1ull << (p - lo);
            }
            for (int q = pos > mend ? pos : mend; q < stop; ++q) ev |= 1ull << (q - lo);
        }
        rest = b >= 64 ? 0ull : (rest & ~((1ull << b) - 1ull));
    }
    S.ev[t] = ev;
    S.ms[t] = ms;
}

// Length of the match that starts at position lo+k.
LPNG_HD int match_length(const Shared& S, int t, int m, int k) {
    const unsigned long long rest = ~S.eq[t] >> k;  // bit 0 clear; bits past the data are set
    if (rest && run_starts_at(S, t, k)) return ctz64(rest);  // the whole run lies in these 64 bytes: one match
    const RunEnds r = run_of(S, t, m, k);
    const int left = match_end(r.s, r.e) - (t * SEG + k);
    return left < MAX_MATCH ? left : MAX_MATCH;
}

// The same tokens with all literals first and all matches after them: for the sinks that do not care about
// order (histogram, bit count) this keeps each loop body uniform across the lanes of a warp.
template <class Sink>
LPNG_HD void sweep_unordered(const Shared& S, int t, int m, Sink& sink) {
    const uint8_t* seg = reinterpret_cast<const uint8_t*>(S.f) + pad(t * SEG);
    const unsigned long long ev = S.ev[t], ms = S.ms[t];
    for (int h = 0; h < 2; ++h) {
        uint32_t e = (uint32_t)((ev & ~ms) >> (32 * h));
        while (e) {
            const int k = ctz32(e);
            e &= e - 1u;
            sink.lit(seg[k + 32 * h]);
        }
    }
    for (int h = 0; h < 2; ++h) {
        uint32_t e = (uint32_t)(ms >> (32 * h));
        while (e) {
            const int k = ctz32(e);
            e &= e - 1u;
            sink.match(match_length(S, t, m, k + 32 * h));
        }
    }
}

template <class Sink>
LPNG_HD void sweep(const Shared& S, int t, int m, Sink& sink) {
    const uint8_t* seg = reinterpret_cast<const uint8_t*>(S.f) + pad(t * SEG);  // a thread's 64 bytes are contiguous
    const unsigned long long ev = S.ev[t], ms = S.ms[t];
    for (int h = 0; h < 2; ++h) {
        uint32_t e = (uint32_t)(ev >> (32 * h));
        const uint32_t mh = (uint32_t)(ms >> (32 * h));
        while (e) {
            const int k = ctz32(e);
            e &= e - 1u;
            if ((mh >> k) & 1u)
                sink.match(match_length(S, t, m, k + 32 * h));
            else
                sink.lit(seg[k + 32 * h]);
        }
    }
}

// Used symbols, in any order, into X.used (S.nused and S.nlit start at 0 / 257).
LPNG_HD void collect_symbols(Shared& S, int t) {
    Scratch& X = scratch(S);
    for (int s = t; s < NLITLEN; s += THREADS)
        if (S.hist[s]) {
            X.used[atomic_inc(&S.nused)] = (uint16_t)s;
            if (s >= 257) atomic_max(&S.nlit, s + 1);
        }
}

// X.order = used symbols by ascending (frequency, symbol): thread t places the t-th used symbol by
// counting the used symbols that come before it.
LPNG_HD void rank_symbols(Shared& S, int t) {
    Scratch& X = scratch(S);
    const int n = (int)S.nused;
    for (int i = t; i < n; i += THREADS) {
        const int s = X.used[i];
        const uint32_t f = S.hist[s];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const int u = X.used[j];
            const uint32_t fu = S.hist[u];
            r += (fu < f || (fu == f && u < s)) ? 1 : 0;
        }
        X.order[r] = (uint16_t)s;
    }
}

// Code lengths (at most MAX_BITS) from S.hist and the first canonical code of every length; one thread.
// Huffman's algorithm with two queues over the frequency-sorted symbols (X.order), depths read off the
// parent links, over-long codes folded into MAX_BITS and the Kraft sum repaired by demoting shorter
// codes, lengths then handed out again in frequency order.
LPNG_HDN void build_code(Shared& S) {
    Scratch& X = scratch(S);
    const int n = (int)S.nused;
    for (int i = 0; i < n; ++i) X.node_freq[i] = S.hist[X.order[i]];
    int li = 0, ii = n, nn = n;
    while (nn < 2 * n - 1) {
        int a, b;
        if (li < n && (ii >= nn || X.node_freq[li] <= X.node_freq[ii])) a = li++; else a = ii++;
        if (li < n && (ii >= nn || X.node_freq[li] <= X.node_freq[ii])) b = li++; else b = ii++;
        X.node_freq[nn] = X.node_freq[a] + X.node_freq[b];
        X.node_parent[a] = (uint16_t)nn;
        X.node_parent[b] = (uint16_t)nn;
        ++nn;
    }
    // depths, root first (parents have larger indices than their children); node_freq is reused
    uint32_t cnt[MAX_BITS + 2];
    for (int i = 0; i <= MAX_BITS + 1; ++i) cnt[i] = 0;
    X.node_freq[2 * n - 2] = 0;
    for (int i = 2 * n - 3; i >= 0; --i) {
        uint32_t d = X.node_freq[X.node_parent[i]] + 1;
        X.node_freq[i] = d;
        if (i < n) ++cnt[d > MAX_BITS ? MAX_BITS : d];
    }
    uint32_t total = 0;
    for (int i = 1; i <= MAX_BITS; ++i) total += cnt[i] << (MAX_BITS - i);
    while (total > (1u << MAX_BITS)) {
        --cnt[MAX_BITS];
        for (int i = MAX_BITS - 1; i >= 1; --i)
            if (cnt[i]) {
                --cnt[i];
                cnt[i + 1] += 2;
                break;
            }
        --total;
    }
    int i = n - 1;  // most frequent symbol first, shortest length first
    for (int l = 1; l <= MAX_BITS; ++l)
        for (uint32_t c = 0; c < cnt[l]; ++c) S.len[X.order[i--]] = (uint8_t)l;
    uint32_t code = 0;
    cnt[0] = 0;
    X.next_code[0] = 0;
    for (int l = 1; l <= MAX_BITS; ++l) {
        code = (code + cnt[l - 1]) << 1;
        X.next_code[l] = code;
    }
}

// The block header (RFC 1951, 3.2.7) uses one fixed code for the code lengths: length 0 costs one bit, lengths
// 1..15 (and the unused zero-run symbol 18, which completes the code) five bits; symbols 16 and 17 are left out.
// Canonical codes of that code: 0 -> "0", L -> 15 + L in five bits. Most of the 257+ lengths are 0, so a header
// is about 70 bytes instead of the 150 a flat four-bit code would take.
constexpr uint32_t HEADER_FIXED_BITS = 17u + 19u * 3u;  // BFINAL, BTYPE, HLIT, HDIST, HCLEN, then 19 three-bit lengths

// Canonical codes: within one length, codes ascend with the symbol. Thread t places the t-th used symbol.
LPNG_HD void assign_codes(Shared& S, int t) {
    Scratch& X = scratch(S);
    const int n = (int)S.nused;
    for (int i = t; i < n; i += THREADS) {
        const int s = X.used[i];
        const int l = S.len[s];
        uint32_t before = 0, below = 0;
        for (int j = 0; j < n; ++j) {
            const int u = X.used[j];
            before += (S.len[u] == l && u < s) ? 1u : 0u;
            below += u < s ? 1u : 0u;
        }
        // where the symbol's code length goes in the block header: one bit per unused symbol before it, five per used one
        const uint32_t hpos = HEADER_FIXED_BITS + (uint32_t)s + 4u * below;
        S.codelen[s] = bit_reverse(X.next_code[l] + before, l) | ((uint32_t)l << 16) | (hpos << 20);
    }
}

// Executors run `f(t)` for t in 0..THREADS and then synchronise: see png.cu (CTA) and tests/png_sim.cpp (loop).

// One piece of the filtered stream -> one byte-aligned deflate block in `slot` (OUT_WORDS words) + its meta data.
template <class Exec>
LPNG_HDN void encode_chunk(Exec& ex, Shared& S, const Image& im, const Tables& T, uint32_t chunk, uint32_t nchunks, uint32_t* slot,
                                  ChunkMeta* meta) {
    const unsigned long long c0 = (unsigned long long)chunk * CHUNK;
    const unsigned long long left = im.stream_bytes - c0;
    const int m = left < (unsigned long long)CHUNK ? (int)left : CHUNK;
    const int bpp = (int)im.bpp;
    const bool final_chunk = chunk + 1 == nchunks;
    const uint32_t base_bits = chunk == 0 ? 16u : 0u;  // the zlib header rides in front of the first piece
    uint8_t* fb = reinterpret_cast<uint8_t*>(S.f);
    Scratch& X = scratch(S);
    const int dsym = bpp - 1;                // distance symbol of distance bpp (1..4 -> 0..3)
    const int dother = dsym == 0 ? 1 : 0;    // second distance code, so that the distance code is complete
    const int ndist = dsym + 1 > 2 ? dsym + 1 : 2;

    // 1. filtered stream bytes into shared memory (consecutive threads take consecutive bytes)
    ex.par([&](int t) {
        for (int w = t; w < OUT_WORDS; w += THREADS) S.out[w] = 0;
        for (int w = t; w < NSYM; w += THREADS) S.len[w] = 0;
        if (t == 0) {
            S.nused = 0;
            S.nlit = 257;
        }
        // row by row (filter type, row pointers and bounds are then the same for the whole CTA)
        const uint32_t stride = im.rb + 1;
        uint32_t y = (uint32_t)(c0 / stride);
        uint32_t col = (uint32_t)(c0 - (unsigned long long)y * stride);
        for (int o = 0; o < m; ++y, col = 0) {
            const int cnt = (int)(stride - col) < m - o ? (int)(stride - col) : m - o;
            const int ft = im.filters[y];
            const uint8_t* row = im.px + (size_t)y * im.rb;
            const uint8_t* prow = y > 0 ? row - im.rb : nullptr;
            const int lead = col == 0 ? 1 : 0;  // the row's filter-type byte opens this span
            if (lead && t == 0) fb[pad(o)] = (uint8_t)ft;
            const uint32_t xbase = col == 0 ? 0u : col - 1u;
            switch (ft) {
                case 0: filter_span<0>(fb, o + lead, row, prow, xbase, cnt - lead, im.bpp, t); break;
                case 1: filter_span<1>(fb, o + lead, row, prow, xbase, cnt - lead, im.bpp, t); break;
                case 2: filter_span<2>(fb, o + lead, row, prow, xbase, cnt - lead, im.bpp, t); break;
                case 3: filter_span<3>(fb, o + lead, row, prow, xbase, cnt - lead, im.bpp, t); break;
                default: filter_span<4>(fb, o + lead, row, prow, xbase, cnt - lead, im.bpp, t); break;
            }
            o += cnt;
        }
    });

    // 2. run structure of each thread's 64 bytes, Adler partial sums
    ex.par([&](int t) {
        const int lo = t * SEG;
        int n = m - lo;
        n = n < 0 ? 0 : (n > SEG ? SEG : n);
        unsigned long long eq = 0;
        uint32_t A = 0, B = 0;
        if (n == SEG) {
            // a full segment, four bytes at a time (little endian: byte k of the segment is byte k & 3 of word k >> 2)
            const uint32_t* w = S.f + (SEG / 4 + 1) * t;
            uint32_t prev = t > 0 ? w[-2] : 0u;  // stream bytes lo-4 .. lo-1 (the word before is the pad word)
            const int sh = 8 * bpp;
            for (int j = 0; j < SEG / 4; ++j) {
                const uint32_t x = w[j];
                const uint32_t y = bpp == 4 ? prev : ((prev >> (32 - sh)) | (x << sh));  // the bytes bpp before
                const uint32_t z = x ^ y;
                uint32_t zero = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu) >> 7;  // bit 8i set: byte i of z is zero
                zero = (zero | (zero >> 7) | (zero >> 14) | (zero >> 21)) & 15u;
                eq |= (unsigned long long)zero << (4 * j);
                const uint32_t pair = (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu);
                const uint32_t sum = (pair & 0xFFFFu) + (pair >> 16);
                A += sum;
                B += (uint32_t)(SEG - 4 * j) * sum - (((x >> 8) & 255u) + 2u * ((x >> 16) & 255u) + 3u * (x >> 24));
                prev = x;
            }
            if (t == 0) eq &= ~((1ull << bpp) - 1ull);  // nothing lies bpp before the first bytes of the piece
        } else {
            for (int k = 0; k < n; ++k) {
                int pos = lo + k;
                uint32_t v = fb[pad(pos)];
                if (pos >= bpp && v == fb[pad(pos - bpp)]) eq |= 1ull << k;
                A += v;
                B += (uint32_t)(n - k) * v;
            }
        }
        S.eq[t] = eq;
        const unsigned long long valid = n == 64 ? ~0ull : ((1ull << n) - 1ull);
        const unsigned long long zeros = ~eq & valid;
        S.sa[t] = zeros ? lo + 63 - clz64(zeros) : -1;   // last position that is not in a run
        S.sb[t] = zeros ? lo + ctz64(zeros) : BIG;       // first such position
        S.crc[t] = A;
        S.bits[t] = (int)B;
    });
    // inclusive scans: sa[t] = last non-run position at or before segment t, sb[t] = first at or after it
    for (int d = 1; d < THREADS; d <<= 1) {
        ex.par([&](int t) {
            int a = S.sa[t], b = S.sb[t];
            if (t - d >= 0 && S.sa[t - d] > a) a = S.sa[t - d];
            if (t + d < THREADS && S.sb[t + d] < b) b = S.sb[t + d];
            X.ta[t] = a;
            X.tb[t] = b;
        });
        ex.par([&](int t) {
            S.sa[t] = X.ta[t];
            S.sb[t] = X.tb[t];
        });
    }

    // 3. token starts, then symbol frequencies, counted into HIST_REPLICAS copies held in the (still zero) output buffer
    ex.par([&](int t) {
        classify(S, t, m);
        HistSink h{S.out + (t % HIST_REPLICAS) * HIST_PITCH};
        sweep_unordered(S, t, m, h);
    });
    ex.par([&](int t) {
        for (int s = t; s < NSYM; s += THREADS) {
            uint32_t f = s == EOB ? 1u : 0u;
            for (int r = 0; r < HIST_REPLICAS; ++r) f += S.out[r * HIST_PITCH + s];
            S.hist[s] = f;
        }
    });
    // 4. used symbols sorted by (frequency, symbol): every thread ranks its own symbols
    ex.par([&](int t) { collect_symbols(S, t); });
    ex.par([&](int t) { rank_symbols(S, t); });
    // 5. the code (one thread) while another thread folds the Adler partial sums
    ex.par([&](int t) {
        if (t == 0) build_code(S);
        if (t == 32) {
            unsigned long long a = 0, b = 0;
            for (int u = 0; u < THREADS; ++u) {
                int n = m - u * SEG;
                if (n <= 0) break;
                if (n > SEG) n = SEG;
                b += (unsigned long long)n * a + (uint32_t)S.bits[u];
                a += S.crc[u];
            }
            S.chunk_a = (uint32_t)(a % ADLER_MOD);
            S.chunk_b = (uint32_t)(b % ADLER_MOD);
        }
    });
    ex.par([&](int t) { assign_codes(S, t); });
    // 6. bits per thread, prefix sums
    ex.par([&](int t) {
        CountSink c{S.len, 0u};
        sweep_unordered(S, t, m, c);
        S.bits[t] = (int)c.bits;
        for (int w = t; w < OUT_WORDS; w += THREADS) S.out[w] = 0;  // histogram copies and scratch are done with: the output buffer again
    });
    for (int d = 1; d < THREADS; d <<= 1) {
        ex.par([&](int t) { S.crc[t] = (uint32_t)(S.bits[t] + (t - d >= 0 ? S.bits[t - d] : 0)); });
        ex.par([&](int t) { S.bits[t] = (int)S.crc[t]; });
    }
    // 7. dynamic block or stored block, whichever is shorter
    ex.par([&](int t) {
        if (t != 0) return;
        S.header_bits = HEADER_FIXED_BITS + (uint32_t)(S.nlit + ndist) + 4u * (S.nused + 2u);
        S.data_bits = (uint32_t)S.bits[THREADS - 1];
        uint32_t end = base_bits + S.header_bits + S.data_bits + S.len[EOB];
        uint32_t n_dyn = final_chunk ? (end + 7u) / 8u : (end + 3u + 7u) / 8u + 4u;
        uint32_t n_sto = base_bits / 8u + 5u + (uint32_t)m;
        S.stored = n_sto < n_dyn;
        S.n_out = (int)(S.stored ? n_sto : n_dyn);
    });
    // 8. the bit stream
    ex.par([&](int t) {
        uint8_t* ob = reinterpret_cast<uint8_t*>(S.out);
        if (S.stored) {
            const int hb = (int)(base_bits / 8u);
            if (t == 0) {
                if (chunk == 0) {
                    ob[0] = 0x78;
                    ob[1] = 0x01;
                }
                ob[hb] = final_chunk ? 1 : 0;
                ob[hb + 1] = (uint8_t)(m & 255);
                ob[hb + 2] = (uint8_t)(m >> 8);
                ob[hb + 3] = (uint8_t)(~m & 255);
                ob[hb + 4] = (uint8_t)((~m >> 8) & 255);
            }
            for (int o = t; o < m; o += THREADS) ob[hb + 5 + o] = fb[pad(o)];
            return;
        }
        const uint32_t hdr = base_bits;
        if (t == 0) {
            if (chunk == 0) or_bits(S.out, 0, 0x0178u);
            uint32_t h17 = (final_chunk ? 1u : 0u) | (2u << 1) | ((uint32_t)(S.nlit - 257) << 3) | ((uint32_t)(ndist - 1) << 8) | (15u << 13);
            or_bits(S.out, hdr, h17);
            // lengths of the code-length code in the order 16, 17, 18, 0, 8, 7, ...: 0, 0, 5, 1, then 5 for 1..15
            or_bits(S.out, hdr + 17u + 6u, 5u);
            or_bits(S.out, hdr + 17u + 9u, 1u);
            for (int j = 4; j < 19; ++j) or_bits(S.out, hdr + 17u + 3u * (uint32_t)j, 5u);
            // end of block, then (unless this is the last piece) an empty stored block to reach a byte boundary
            uint32_t eob = hdr + S.header_bits + S.data_bits;
            or_bits(S.out, eob, S.codelen[EOB] & 0xFFFFu);
            if (!final_chunk) {
                uint32_t bs = (eob + S.len[EOB] + 3u + 7u) / 8u;
                or_bits(S.out, 8u * (bs + 2u), 0xFFFFu);
            }
        }
        for (int q = t; q < S.nlit; q += THREADS)  // used literal/length symbols: their lengths (unused ones are single 0 bits)
            if (S.len[q]) or_bits(S.out, hdr + (S.codelen[q] >> 20), bit_reverse(15u + S.len[q], 5));
        if (t == 0) {  // the two distance codes of length 1
            const uint32_t dbase = hdr + HEADER_FIXED_BITS + (uint32_t)S.nlit + 4u * S.nused;
            const int d0 = dsym < dother ? dsym : dother, d1 = dsym < dother ? dother : dsym;
            or_bits(S.out, dbase + (uint32_t)d0, bit_reverse(16u, 5));
            or_bits(S.out, dbase + (uint32_t)d1 + 4u, bit_reverse(16u, 5));
        }
        EmitSink e;
        e.out = S.out;
        e.codelen = S.codelen;
        e.dist_bit = dsym > dother ? 1u : 0u;
        e.begin(hdr + S.header_bits + (t > 0 ? (uint32_t)S.bits[t - 1] : 0u));
        sweep(S, t, m, e);
        e.end();
    });
    // 9. CRC-32 of "IDAT" + piece. The register's 0xFFFFFFFF start equals complementing the first four
    // message bytes; slices are counted from the END of the message so every slice but the first is full
    // and the tree below multiplies by the same power of x at each level.
    ex.par([&](int t) {
        const uint8_t* ob = reinterpret_cast<const uint8_t*>(S.out);
        const int total = S.n_out + 4;
        int hi = total - CRC_SLICE * t, lo = hi - CRC_SLICE;
        if (lo < 0) lo = 0;
        uint32_t s = 0;
        for (int p = lo; p < hi; ++p) {
            uint32_t byte = p < 4 ? (uint32_t)(uint8_t)~"IDAT"[p] : ob[p - 4];
            s = T.crc[(s ^ byte) & 255u] ^ (s >> 8);
        }
        S.crc[t] = s;
    });
    for (int d = 1, j = 0; d < THREADS; d <<= 1, ++j) {
        ex.par([&](int t) {
            if ((t & (2 * d - 1)) == 0) {
                uint32_t earlier = S.crc[t + d];
                if (earlier) S.crc[t] ^= gf_mul(earlier, T.xs[j]);
            }
        });
    }
    // 10. out
    ex.par([&](int t) {
        for (int w = t; w < (S.n_out + 3) / 4; w += THREADS) slot[w] = S.out[w];
        if (t == 0) {
            meta->n_out = (uint32_t)S.n_out;
            meta->crc = ~S.crc[0];
            meta->adler_a = S.chunk_a;
            meta->adler_b = S.chunk_b;
            meta->m = (uint32_t)m;
        }
    });
}

// ---------------------------------------------------------------------------------------------
// Assembly: where each IDAT chunk goes, the Adler-32 of the whole stream, the fixed chunks.
// ---------------------------------------------------------------------------------------------
struct AdlerPart {
    uint32_t len, a, b;  // all mod 65521
};
LPNG_HD AdlerPart adler_join(AdlerPart x, AdlerPart y) {  // x first, then y
    AdlerPart r;
    r.len = (x.len + y.len) % ADLER_MOD;
    r.a = (x.a + y.a) % ADLER_MOD;
    r.b = (uint32_t)(((unsigned long long)x.b + y.b + (unsigned long long)y.len * x.a) % ADLER_MOD);
    return r;
}

struct LayoutShared {
    unsigned long long sum[THREADS], tsum[THREADS];
    AdlerPart ad[THREADS], tad[THREADS];
};

LPNG_HD void put_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24);
    p[1] = (uint8_t)(v >> 16);
    p[2] = (uint8_t)(v >> 8);
    p[3] = (uint8_t)v;
}

LPNG_HD int color_type(uint32_t bpp) { return bpp == 1 ? 0 : bpp == 2 ? 4 : bpp == 3 ? 2 : 6; }

// offsets[c] = file offset of piece c's IDAT chunk; writes signature, IHDR, the Adler IDAT and IEND; *total = file size.
template <class Exec>
LPNG_HDN void layout(Exec& ex, LayoutShared& L, const ChunkMeta* meta, uint32_t nchunks, uint32_t width, uint32_t height, uint32_t bpp,
                            unsigned long long* offsets, uint8_t* out, unsigned long long* total) {
    const uint32_t per = (nchunks + THREADS - 1) / THREADS;
    ex.par([&](int t) {
        unsigned long long s = 0;
        AdlerPart a{0, 0, 0};
        uint32_t c0 = (uint32_t)t * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
        for (uint32_t c = c0; c < c1; ++c) {
            s += 12ull + meta[c].n_out;
            a = adler_join(a, AdlerPart{meta[c].m % ADLER_MOD, meta[c].adler_a, meta[c].adler_b});
        }
        L.sum[t] = s;
        L.ad[t] = a;
    });
    for (int d = 1; d < THREADS; d <<= 1) {
        ex.par([&](int t) {
            L.tsum[t] = L.sum[t] + (t - d >= 0 ? L.sum[t - d] : 0ull);
            L.tad[t] = t - d >= 0 ? adler_join(L.ad[t - d], L.ad[t]) : L.ad[t];
        });
        ex.par([&](int t) {
            L.sum[t] = L.tsum[t];
            L.ad[t] = L.tad[t];
        });
    }
    ex.par([&](int t) {
        unsigned long long off = PNG_HEADER_BYTES + (t > 0 ? L.sum[t - 1] : 0ull);
        uint32_t c0 = (uint32_t)t * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
        for (uint32_t c = c0; c < c1; ++c) {
            offsets[c] = off;
            off += 12ull + meta[c].n_out;
        }
        if (t != 0) return;
        out[0] = 0x89; out[1] = 'P'; out[2] = 'N'; out[3] = 'G';
        out[4] = 0x0D; out[5] = 0x0A; out[6] = 0x1A; out[7] = 0x0A;
        uint8_t* p = out + 8;
        put_be32(p, 13);
        p[4] = 'I'; p[5] = 'H'; p[6] = 'D'; p[7] = 'R';
        put_be32(p + 8, width);
        put_be32(p + 12, height);
        p[16] = 8;
        p[17] = (uint8_t)color_type(bpp);
        p[18] = 0; p[19] = 0; p[20] = 0;
        uint32_t s = 0xFFFFFFFFu;
        for (int i = 4; i < 21; ++i) s = crc_bitwise(s, p[i]);
        put_be32(p + 21, ~s);
        // zlib trailer in an IDAT chunk of its own: Adler-32 of the whole filtered stream
        AdlerPart w = L.ad[THREADS - 1];
        uint32_t s1 = (1u + w.a) % ADLER_MOD, s2 = (w.len + w.b) % ADLER_MOD;
        p = out + PNG_HEADER_BYTES + L.sum[THREADS - 1];
        put_be32(p, 4);
        p[4] = 'I'; p[5] = 'D'; p[6] = 'A'; p[7] = 'T';
        put_be32(p + 8, (s2 << 16) | s1);
        s = 0xFFFFFFFFu;
        for (int i = 4; i < 12; ++i) s = crc_bitwise(s, p[i]);
        put_be32(p + 12, ~s);
        p += 16;
        put_be32(p, 0);
        p[4] = 'I'; p[5] = 'E'; p[6] = 'N'; p[7] = 'D';
        put_be32(p + 8, 0xAE426082u);
        *total = PNG_HEADER_BYTES + L.sum[THREADS - 1] + PNG_TRAILER_BYTES;
    });
}

// Piece c's IDAT chunk at its place: length, type, bytes, CRC.
template <class Exec>
LPNG_HDN void place_chunk(Exec& ex, const ChunkMeta& mt, const uint32_t* slot, uint8_t* dst) {
    ex.par([&](int t) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(slot);
        for (uint32_t o = (uint32_t)t; o < mt.n_out; o += THREADS) dst[8 + o] = src[o];
        if (t == 0) {
            put_be32(dst, mt.n_out);
            dst[4] = 'I'; dst[5] = 'D'; dst[6] = 'A'; dst[7] = 'T';
            put_be32(dst + 8 + mt.n_out, mt.crc);
        }
    });
}

// Upper bound of the file size (every piece stored).
inline unsigned long long png_bound(uint32_t width, uint32_t height, uint32_t bpp) {
    unsigned long long stream = (unsigned long long)height * ((unsigned long long)width * bpp + 1ull);
    unsigned long long nchunks = (stream + CHUNK - 1) / CHUNK;
    return PNG_HEADER_BYTES + stream + nchunks * (12ull + 5ull) + 2ull + PNG_TRAILER_BYTES;
}

}  // namespace lpng
