// png_sim.cpp -- host instantiation of lutgen_rs_b200/csrc/png_core.h for the CPU unit tests.
//
// TEST INFRASTRUCTURE ONLY. The PNG encoder's per-thread phases are written against an executor; the
// CUDA kernels (csrc/png.cu) run them with one CUDA thread per t and __syncthreads between phases,
// this file runs the same phases with a loop over t. Nothing in the product path links or calls this:
// tests/test_png_core.py builds it with g++ and checks what it writes with independent decoders
// (zlib, PIL). There is no GPU in the development container, so this is how the code the GPU runs
// gets exercised before it travels.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../lutgen_rs_b200/csrc/png_core.h"

namespace {
// Order in which the "threads" of a phase run: 0 ascending, 1 descending, 2 a different pseudo-random permutation in
// every phase. A phase whose outcome depends on this order has a race (a thread reading what another writes in the
// same phase, or two plain writes to one place) -- on the GPU the order is anybody's guess -- so the tests demand the
// same file for all three.
int g_order = 0;

struct LoopExec {
    unsigned long long state = 0x9E3779B97F4A7C15ull;
    template <class F>
    void par(F f) {
        if (g_order == 0) {
            for (int t = 0; t < lpng::THREADS; ++t) f(t);
        } else if (g_order == 1) {
            for (int t = lpng::THREADS; t-- > 0;) f(t);
        } else {
            int order[lpng::THREADS];
            for (int t = 0; t < lpng::THREADS; ++t) order[t] = t;
            for (int i = lpng::THREADS; i > 1; --i) {
                state ^= state << 13;
                state ^= state >> 7;
                state ^= state << 17;
                int j = (int)(state % (unsigned)i);
                int tmp = order[i - 1];
                order[i - 1] = order[j];
                order[j] = tmp;
            }
            for (int t = 0; t < lpng::THREADS; ++t) f(order[t]);
        }
    }
};
}  // namespace

extern "C" {

void lpng_sim_set_order(int mode) { g_order = mode; }

unsigned long long lpng_sim_bound(uint32_t width, uint32_t height, uint32_t bpp) { return lpng::png_bound(width, height, bpp); }

// Returns 0 and the file size in *len; -1 if `cap` is below the bound. filters_out (nullable): height bytes.
int lpng_sim_encode(const uint8_t* px, uint32_t width, uint32_t height, uint32_t bpp, uint8_t* out, unsigned long long cap, unsigned long long* len,
                    uint8_t* filters_out, uint32_t* stored_chunks) {
    using namespace lpng;
    if (cap < png_bound(width, height, bpp)) return -1;
    Tables T;
    make_tables(T);
    std::vector<uint8_t> filters(height);
    Image im{px, filters.data(), width, height, bpp, width * bpp, (unsigned long long)height * ((unsigned long long)width * bpp + 1ull)};
    for (uint32_t y = 0; y < height; ++y) {  // one warp per row on the GPU (k_png_choose_filters)
        uint32_t cost[5] = {0, 0, 0, 0, 0};
        const uint8_t* row = px + (size_t)y * im.rb;
        for (uint32_t lane = 0; lane < 32; ++lane) row_costs(row, y > 0 ? row - im.rb : nullptr, im.rb, bpp, lane, 32, cost);
        filters[y] = (uint8_t)pick_filter(cost);
    }
    if (filters_out) memcpy(filters_out, filters.data(), height);
    uint32_t nchunks = (uint32_t)((im.stream_bytes + CHUNK - 1) / CHUNK);
    std::vector<uint32_t> slots((size_t)nchunks * OUT_WORDS);
    std::vector<ChunkMeta> meta(nchunks);
    std::vector<unsigned long long> offsets(nchunks);
    LoopExec ex;
    Shared* S = new Shared;
    uint32_t stored = 0;
    for (uint32_t c = 0; c < nchunks; ++c) {
        memset(S, 0xA5, sizeof(Shared));  // shared memory starts as garbage
        encode_chunk(ex, *S, im, T, c, nchunks, slots.data() + (size_t)c * OUT_WORDS, &meta[c]);
        stored += S->stored ? 1 : 0;
    }
    delete S;
    if (stored_chunks) *stored_chunks = stored;
    LayoutShared* L = new LayoutShared;
    unsigned long long total = 0;
    layout(ex, *L, meta.data(), nchunks, width, height, bpp, offsets.data(), out, &total);
    delete L;
    for (uint32_t c = 0; c < nchunks; ++c) place_chunk(ex, meta[c], slots.data() + (size_t)c * OUT_WORDS, out + offsets[c]);
    *len = total;
    return 0;
}

// Code lengths and (bit-reversed) canonical codes for a histogram of the 286 literal/length symbols.
void lpng_sim_code(const uint32_t* hist, uint8_t* len, uint16_t* code) {
    using namespace lpng;
    Shared* S = new Shared;
    memset(S, 0xA5, sizeof(Shared));
    for (int s = 0; s < NSYM; ++s) {
        S->hist[s] = s < NLITLEN ? hist[s] : 0;
        S->len[s] = 0;
    }
    S->nused = 0;
    S->nlit = 257;
    for (int t = 0; t < THREADS; ++t) collect_symbols(*S, t);
    for (int t = 0; t < THREADS; ++t) rank_symbols(*S, t);
    build_code(*S);
    for (int t = 0; t < THREADS; ++t) assign_codes(*S, t);
    for (int s = 0; s < NLITLEN; ++s) {
        len[s] = S->len[s];
        code[s] = S->len[s] ? (uint16_t)(S->codelen[s] & 0xFFFFu) : 0;
    }
    delete S;
}

uint32_t lpng_sim_shared_bytes(void) { return (uint32_t)sizeof(lpng::Shared); }
}
