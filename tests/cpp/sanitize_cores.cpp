// sanitize_cores.cpp -- drives the host instantiations of the PNG encoder core (csrc/png_core.h through
// tests/png_sim.cpp) and of the palette-extraction kernel bodies (csrc/extract_core.h through
// tests/extract_sim.cpp) over a spread of geometries and contents, to be built with
// -fsanitize=address,undefined,bounds-strict (tests/test_host_sanitizers.py). compute-sanitizer is closed on the
// GPU pool, so this is the memory / undefined-behaviour check these cores get: out-of-range indices into the
// shared-state arrays, bad shifts, signed overflow. TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" {
unsigned long long lpng_sim_bound(uint32_t, uint32_t, uint32_t);
int lpng_sim_encode(const uint8_t*, uint32_t, uint32_t, uint32_t, uint8_t*, unsigned long long, unsigned long long*, uint8_t*, uint32_t*);
uint32_t lext_sim_points(const uint8_t*, size_t, uint64_t, uint32_t*, uint32_t*, uint8_t*);
uint32_t lext_sim_kmeans(uint32_t, const uint32_t*, const uint32_t*, const float*, uint32_t, uint32_t, uint64_t, float*);
}
static uint64_t rs = 88172645463325252ull;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
int main() {
    // PNG: geometry x content kinds
    const uint32_t shapes[][3] = {{1,1,1},{1,1,4},{7,300,3},{300,7,2},{64,64,3},{3,5461,3},{2,10923,3},{1,16385,1},{40,700,4},{1,40000,3},{200,1,1}};
    unsigned long long total = 0;
    for (auto& s : shapes) {
        uint32_t h = s[0], w = s[1], c = s[2];
        for (int kind = 0; kind < 5; ++kind) {
            std::vector<uint8_t> img((size_t)h * w * c);
            for (size_t i = 0; i < img.size(); ++i) {
                switch (kind) {
                    case 0: img[i] = (uint8_t)rnd(); break;
                    case 1: img[i] = (uint8_t)((rnd() % 3) * 50); break;
                    case 2: img[i] = (uint8_t)(i / 7); break;
                    case 3: img[i] = 200; break;
                    default: img[i] = (uint8_t)(((i / c) % w) + (i % c) * 40); break;
                }
            }
            unsigned long long cap = lpng_sim_bound(w, h, c), len = 0;
            std::vector<uint8_t> out(cap), filters(h);
            uint32_t stored = 0;
            if (lpng_sim_encode(img.data(), w, h, c, out.data(), cap, &len, filters.data(), &stored) != 0 || len > cap) { printf("FAIL\n"); return 1; }
            total += len;
        }
    }
    // extraction
    for (int t = 0; t < 6; ++t) {
        size_t npix = 500 + 3000 * t;
        std::vector<uint8_t> rgb(3 * npix);
        for (auto& v : rgb) v = (uint8_t)(t % 2 ? rnd() : (rnd() % 5) * 60);
        std::vector<uint32_t> keys(npix), weights(npix);
        std::vector<uint8_t> list(3 * npix);
        uint32_t n = lext_sim_points(rgb.data(), npix, t, keys.data(), weights.data(), list.data());
        std::vector<float> lab(3 * (size_t)n), out_lab(3 * 256);
        for (size_t i = 0; i < lab.size(); ++i) lab[i] = (float)list[i] / 255.0f - (i % 3 ? 0.5f : 0.0f);  // any coordinates will do here
        uint32_t k = lext_sim_kmeans(n, keys.data(), weights.data(), lab.data(), 1 + 40 * t, 5, t, out_lab.data());
        total += k;
    }
    printf("ok %llu\n", total);
    return 0;
}
