// One process, several GPUs: the crate-shaped API of include/lutgen.hpp, unchanged, with
// LUTGEN_CUDA_DEVICES naming the devices in the environment. Prints a hash per result; tests/test_cpp_host.py runs it
// on one device and on all of them and requires identical output: the sharding must be invisible to the caller
// (crates/lib/src/lib.rs:143-147 par_generate_lut and the frame loop of crates/cli/src/main.rs:878-888 are one call each).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "lutgen.hpp"

using namespace lutgen;
using interpolation::Palette;

static uint64_t fnv(const std::vector<uint8_t>& v) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : v) { h ^= b; h *= 1099511628211ull; }
    return h;
}

static const Palette kMocha = {{0x11, 0x11, 0x1b}, {0x18, 0x18, 0x25}, {0x31, 0x32, 0x44}, {0x45, 0x47, 0x5a}, {0x58, 0x5b, 0x70},
                               {0xcd, 0xd6, 0xf4}, {0xf5, 0xe0, 0xdc}, {0xb4, 0xbe, 0xfe}, {0xf3, 0x8b, 0xa8}, {0xa6, 0xe3, 0xa1},
                               {0xfa, 0xb3, 0x87}, {0x89, 0xb4, 0xfa}, {0x94, 0xe2, 0xd5}, {0xf9, 0xe2, 0xaf}, {0xcb, 0xa6, 0xf7}};

int main() {
    int in_use = 0;
    detail::ensure_init();
    lutgen_cuda_device_count_in_use(&in_use);
    std::printf("# devices in use: %d\n", in_use);
    {
        interpolation::GaussianRemapper g(kMocha, 128.0, 8, 1.0, false);
        for (int level : {16, 12, 9, 4}) {
            RgbImage lut = g.par_generate_lut((uint8_t)level);
            std::printf("gaussian level %d %016llx\n", level, (unsigned long long)fnv(lut.data));
        }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) (void)g.par_generate_lut(16);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / 5;
        std::printf("# level-16 CLUT into pageable host memory: %.3f ms\n", ms);
        std::atomic<bool> abort{true};
        std::printf("aborted %d\n", (int)!g.par_generate_lut_with_interrupt(16, abort).has_value());
    }
    {
        interpolation::NearestNeighborRemapper nn(kMocha, 1.0, false);
        std::printf("nearest level 16 %016llx\n", (unsigned long long)fnv(nn.generate_lut(16).data));
        interpolation::ShepardRemapper sh(kMocha, 4.0, 0, 1.2, true);
        std::printf("shepard level 13 %016llx\n", (unsigned long long)fnv(sh.generate_lut(13).data));
        interpolation::GaussianSamplingRemapper gs(kMocha, 0.0, 20.0, 128, 1.0, 42080085, false);
        std::printf("sampling level 12 %016llx\n", (unsigned long long)fnv(gs.par_generate_lut(12).data));
        interpolation::GaussianBlurRemapper gb(kMocha, 3.0, 1.0, false);
        std::printf("blur level 12 %016llx\n", (unsigned long long)fnv(gb.par_generate_lut(12).data));
    }
    {
        interpolation::GaussianRemapper g(kMocha, 96.0, 0, 1.0, false);
        RgbImage lut = g.par_generate_lut(8);
        std::vector<RgbaImage> frames;
        uint32_t s = 777;
        for (int f = 0; f < 13; ++f) {
            frames.emplace_back(640 + 16 * f, 360 + (f % 3));
            for (auto& b : frames.back().data) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
        }
        std::vector<RgbaImage*> ptrs;
        for (auto& f : frames) ptrs.push_back(&f);
        identity::correct_images(ptrs, lut);
        uint64_t h = 0;
        for (auto& f : frames) h = h * 31 + fnv(f.data);
        std::printf("correct_images 13 frames %016llx\n", (unsigned long long)h);
    }
    return 0;
}
