// C++ host-mirror test: drives include/lutgen.hpp (the crate-shaped API over the C ABI) and checks
// it against the CPU oracle, in the shape of the reference's own doctests (crates/lib/src/lib.rs:9-97)
// and CLI known-answer test (crates/cli/src/main.rs:1192-1226).
//
//   g++ -std=c++17 -Iinclude tests/cpp/host_mirror.cpp -o tests/cpp/_build/host_mirror \
//       -Llutgen_rs_b200 -llutgen_cuda -Loracle/_build -llutgen_oracle -Wl,-rpath,...
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "lutgen.hpp"

extern "C" {
// oracle/lutgen_oracle.c (test infrastructure)
struct lutgen_oracle_desc {
    int32_t kind, preserve;
    const uint8_t* palette;
    uint64_t palette_len;
    double param;
    uint64_t nearest;
    double lum_factor, mean, std_dev;
    uint64_t iterations, seed;
};
void* lutgen_oracle_remapper_create(const lutgen_oracle_desc*);
void lutgen_oracle_remapper_destroy(void*);
void lutgen_oracle_generate_lut(const void*, uint8_t level, uint8_t* out, int threads);
void lutgen_oracle_correct_image(uint8_t* rgba, size_t npix, const uint8_t* lut, uint8_t level, int threads);
void lutgen_oracle_identity(uint8_t level, uint8_t* out);
}

using namespace lutgen;
using interpolation::Palette;

static int failures = 0;
#define EXPECT(cond)                                                       \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static const Palette kMocha = {{0x11, 0x11, 0x1b}, {0x18, 0x18, 0x25}, {0x31, 0x32, 0x44}, {0x45, 0x47, 0x5a}, {0x58, 0x5b, 0x70},
                               {0xcd, 0xd6, 0xf4}, {0xf5, 0xe0, 0xdc}, {0xb4, 0xbe, 0xfe}, {0xf3, 0x8b, 0xa8}, {0xa6, 0xe3, 0xa1},
                               {0xfa, 0xb3, 0x87}, {0x89, 0xb4, 0xfa}, {0x94, 0xe2, 0xd5}, {0xf9, 0xe2, 0xaf}, {0xcb, 0xa6, 0xf7}};

static int max_abs_diff(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
    int m = 0;
    for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::abs((int)a[i] - (int)b[i]));
    return m;
}

int main() {
    // identity
    {
        RgbImage id = identity::generate(4);
        std::vector<uint8_t> want(id.data.size());
        lutgen_oracle_identity(4, want.data());
        EXPECT(id.width == 64 && id.height == 64);
        EXPECT(id.data == want);
        EXPECT(identity::detect_level(id) == 4);
        bool threw = false;
        try { identity::generate(1); } catch (const Panic&) { threw = true; }
        EXPECT(threw);
        threw = false;
        try { identity::detect_level(RgbImage(100, 100)); } catch (const Panic&) { threw = true; }
        EXPECT(threw);
    }
    // CLI known answer: level-2 NN CLUT of the RGB primaries sends #ff8000 to #ff0000
    {
        Palette prim = {{255, 0, 0}, {0, 255, 0}, {0, 0, 255}};
        interpolation::NearestNeighborRemapper nn(prim, 1.0, false);
        RgbImage lut = nn.generate_lut(2);
        uint8_t in[3] = {255, 128, 0}, out[3] = {0, 0, 0};
        identity::correct_pixel(in, lut, 2, out);
        EXPECT(out[0] == 255 && out[1] == 0 && out[2] == 0);
    }
    // doctest shape (lib.rs:42-71): Gaussian RBF CLUT, then correct an image with it
    {
        interpolation::GaussianRemapper g(kMocha, 96.0, 0, 1.0, false);
        RgbImage lut = g.par_generate_lut(8);
        lutgen_oracle_desc d{0, 0, kMocha.front().data(), kMocha.size(), 96.0, 0, 1.0, 0, 0, 0, 0};
        void* o = lutgen_oracle_remapper_create(&d);
        std::vector<uint8_t> want(lut.data.size());
        lutgen_oracle_generate_lut(o, 8, want.data(), 8);
        lutgen_oracle_remapper_destroy(o);
        EXPECT(max_abs_diff(lut.data, want) <= 1);

        RgbaImage img(640, 360);
        uint32_t s = 12345;
        for (auto& b : img.data) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
        std::vector<uint8_t> ref = img.data;
        lutgen_oracle_correct_image(ref.data(), img.pixels(), lut.data.data(), 8, 1);
        identity::correct_image(img, lut);
        EXPECT(img.data == ref);

        // direct remap (lib.rs:78-97): alpha untouched
        RgbaImage img2(64, 64);
        for (auto& b : img2.data) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
        std::vector<uint8_t> before = img2.data;
        g.par_remap_image(img2);
        bool alpha_same = true;
        for (size_t i = 3; i < before.size(); i += 4) alpha_same &= before[i] == img2.data[i];
        EXPECT(alpha_same);
    }
    // interrupts (lib.rs:149-170, mod.rs:36-42)
    {
        interpolation::ShepardRemapper sh(kMocha, 4.0, 8, 1.0, false);
        std::atomic<bool> abort{false};
        EXPECT(sh.par_generate_lut_with_interrupt(4, abort).has_value());
        abort = true;
        EXPECT(!sh.par_generate_lut_with_interrupt(4, abort).has_value());
        RgbaImage img(16, 16);
        std::vector<uint8_t> before = img.data;
        sh.remap_image_with_interrupt(img, abort);
        EXPECT(img.data == before);
    }
    // sampling: Normal::new(mean, inf).unwrap() panics
    {
        bool threw = false;
        try { interpolation::GaussianSamplingRemapper s(kMocha, 0.0, 1.0 / 0.0, 16, 1.0, 1, false); } catch (const Panic&) { threw = true; }
        EXPECT(threw);
        interpolation::GaussianSamplingRemapper ok(kMocha, 0.0, 20.0, 64, 1.0, 42080085, false);
        RgbImage lut = ok.generate_lut(3);
        EXPECT(lut.width == 27);
    }
    // png::encode: signature, IHDR geometry, IEND at the end (decoding is checked in tests/test_gpu_png.py)
    {
        RgbImage id = identity::generate(4);
        std::vector<uint8_t> file = png::encode(id);
        static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        EXPECT(file.size() > 57 && std::memcmp(file.data(), sig, 8) == 0);
        EXPECT(std::memcmp(file.data() + 12, "IHDR", 4) == 0 && file[18] == 0 && file[19] == 64 && file[22] == 0 && file[23] == 64);
        EXPECT(file[24] == 8 && file[25] == 2);
        EXPECT(std::memcmp(file.data() + file.size() - 8, "IEND", 4) == 0);
    }
    if (failures == 0) std::printf("host mirror OK\n");
    return failures ? 1 : 0;
}
