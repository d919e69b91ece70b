// lutgen.hpp -- C++ host mirror of the Rust crate `lutgen` (crates/lib of ozwaldorf/lutgen-rs 0.15.0)
// on top of the C ABI in lutgen_cuda.h.
//
// The reference's host language is Rust; no Rust toolchain exists in this build environment, so the
// compiled-language host layer above the C ABI is written in C++ (the Rust shim a maintainer would
// add is shown in INTEGRATION.md). Names, argument order and argument meaning follow the crate:
//
//   lutgen::identity::{generate, correct_pixel, correct_image, correct_image_with_level, detect_level}
//       crates/lib/src/identity.rs
//   lutgen::interpolation::{InterpolatedRemapper, GaussianRemapper, ShepardRemapper, LinearRemapper,
//                           NearestNeighborRemapper, GaussianSamplingRemapper}
//       crates/lib/src/interpolation/{mod,rbf,nearest_neighbor,gaussian_sample}.rs
//   GenerateLut (generate_lut, par_generate_lut, *_with_interrupt)      crates/lib/src/lib.rs:113-171
//
// Error behaviour: where the reference panics (invalid CLUT, level < 2, non-finite std_dev ...)
// these functions throw lutgen::Panic; CUDA / environment failures throw lutgen::Error. The
// `*_with_interrupt` functions take a std::atomic<bool> (the reference's Arc<AtomicBool>) and
// return std::nullopt / leave the image untouched when it is set, like the reference's `None`.
// Header only; link with liblutgen_cuda.so.
#pragma once
#include <array>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "lutgen_cuda.h"

namespace lutgen {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Panic : Error {
    using Error::Error;
};

// image::RgbImage / image::RgbaImage: tightly packed, row-major.
template <int CH>
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> data;
    Image() = default;
    Image(uint32_t w, uint32_t h) : width(w), height(h), data((size_t)w * h * CH) {}
    size_t pixels() const { return (size_t)width * height; }
    uint8_t* pixel(uint32_t x, uint32_t y) { return data.data() + ((size_t)y * width + x) * CH; }
    const uint8_t* pixel(uint32_t x, uint32_t y) const { return data.data() + ((size_t)y * width + x) * CH; }
};
using RgbImage = Image<3>;
using RgbaImage = Image<4>;

namespace detail {
// Returns true when the call reported LUTGEN_ABORTED.
inline bool check(int rc) {
    if (rc == LUTGEN_OK) return false;
    if (rc == LUTGEN_ABORTED) return true;
    std::string msg = lutgen_cuda_last_error();
    if (rc == LUTGEN_ERR_INVALID) throw Panic(msg);
    throw Error("lutgen_cuda error " + std::to_string(rc) + ": " + msg);
}
inline void ensure_init() {
    static const bool once = (check(lutgen_cuda_init(-1)), true);
    (void)once;
}
static_assert(sizeof(std::atomic<bool>) == 1, "abort flag is passed to the C ABI as one byte");
inline const volatile uint8_t* flag_ptr(const std::atomic<bool>* f) { return reinterpret_cast<const volatile uint8_t*>(f); }
}  // namespace detail

namespace identity {

// identity::generate(level) -> RgbImage                      identity.rs:7-34
inline RgbImage generate(uint8_t level) {
    detail::ensure_init();
    uint32_t side = (uint32_t)level * level * level;
    RgbImage out(level >= 2 && level <= 33 ? side : 1, level >= 2 && level <= 33 ? side : 1);
    detail::check(lutgen_cuda_identity(level, out.data.data()));
    return out;
}

// identity::detect_level(&RgbImage) -> u8                    identity.rs:85-99
inline uint8_t detect_level(const RgbImage& hald_clut) {
    uint8_t level = 0;
    detail::check(lutgen_cuda_detect_level(hald_clut.width, hald_clut.height, &level));
    return level;
}

// identity::correct_image_with_level(&mut RgbaImage, &RgbImage, level)   identity.rs:71-78
inline void correct_image_with_level(RgbaImage& image, const RgbImage& hald_clut, uint8_t level) {
    detail::ensure_init();
    uint32_t side = (uint32_t)level * level * level;
    if (level >= 2 && level <= 33 && (hald_clut.width != side || hald_clut.height != side))
        throw Panic("hald clut does not have level^3 x level^3 pixels (get_pixel out of bounds)");
    detail::check(lutgen_cuda_apply_hald(image.data.data(), image.pixels(), hald_clut.data.data(), level));
}

// identity::correct_image(&mut RgbaImage, &RgbImage)         identity.rs:62-65
inline void correct_image(RgbaImage& image, const RgbImage& hald_clut) {
    correct_image_with_level(image, hald_clut, detect_level(hald_clut));
}

// identity::correct_pixel(&[u8;3], &RgbImage, level) -> [u8;3]   identity.rs:40-52
inline void correct_pixel(const uint8_t input[3], const RgbImage& hald_clut, uint8_t level, uint8_t output[3]) {
    detail::ensure_init();
    const uint32_t n = (uint32_t)level * level * level;
    if (level >= 2 && level <= 33 && (hald_clut.width != n || hald_clut.height != n))   // the reference's get_pixel would panic
        throw Panic("hald clut is smaller than the level asks for (identity.rs:48-51 indexes out of bounds)");
    detail::check(lutgen_cuda_correct_pixels(input, 1, hald_clut.data.data(), level, output));
}

// One CLUT over many images (the CLI's frame loops, crates/cli/src/main.rs:852-886).
inline void correct_images(std::vector<RgbaImage*>& images, const RgbImage& hald_clut) {
    detail::ensure_init();
    uint8_t level = detect_level(hald_clut);
    std::vector<uint8_t*> ptrs;
    std::vector<size_t> sizes;
    for (RgbaImage* im : images) {
        ptrs.push_back(im->data.data());
        sizes.push_back(im->pixels());
    }
    detail::check(lutgen_cuda_apply_hald_batch(ptrs.data(), sizes.data(), ptrs.size(), hald_clut.data.data(), level));
}

}  // namespace identity

// `image.save(path)` for a .png path (the `image` crate's PNG writer as the reference's callers use it,
// crates/cli/src/main.rs:767, 811, 839, 862), encoded by the CUDA library.
namespace png {

template <int CH>
inline std::vector<uint8_t> encode(const Image<CH>& image) {
    detail::ensure_init();
    size_t bound = 0, len = 0;
    detail::check(lutgen_cuda_png_bound(image.width, image.height, CH, &bound));
    std::vector<uint8_t> file(bound);
    detail::check(lutgen_cuda_encode_png(image.data.data(), image.width, image.height, CH, file.data(), file.size(), &len));
    file.resize(len);
    return file;
}

template <int CH>
inline void save(const Image<CH>& image, const std::string& path) {
    std::vector<uint8_t> file = encode(image);
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error("cannot open " + path);
    size_t n = std::fwrite(file.data(), 1, file.size(), f);
    if (std::fclose(f) != 0 || n != file.size()) throw Error("cannot write " + path);
}

}  // namespace png

// `lutgen extract` (crates/cli/src/main.rs:772-815): the role of quantette's PalettePipeline with
// ColorSpace::Oklab and QuantizeMethod::kmeans() -- a deterministic k-means in Oklab, NOT quantette's palette
// (see lutgen_cuda_extract_palette in lutgen_cuda.h).
namespace extract {

inline std::vector<std::array<uint8_t, 3>> palette(const RgbImage& image, uint32_t color_count, uint32_t iterations = 16) {
    detail::ensure_init();
    std::vector<std::array<uint8_t, 3>> out(color_count ? color_count : 1);
    uint32_t n = 0;
    detail::check(lutgen_cuda_extract_palette(image.data.data(), image.pixels(), color_count, iterations, out.front().data(), &n));
    out.resize(n);
    return out;
}

}  // namespace extract

namespace interpolation {

// trait InterpolatedRemapper (interpolation/mod.rs:23-61) + the GenerateLut blanket impl (lib.rs:135-171).
// The sequential and par_ spellings are one GPU call; both exist so callers do not change.
class InterpolatedRemapper {
public:
    InterpolatedRemapper(const InterpolatedRemapper&) = delete;
    InterpolatedRemapper& operator=(const InterpolatedRemapper&) = delete;
    virtual ~InterpolatedRemapper() {
        if (handle_) lutgen_cuda_remapper_destroy(handle_);
    }

    // remap_pixel(&self, &mut Rgba<u8>)                       mod.rs:25
    void remap_pixel(uint8_t pixel[4]) const { detail::check(lutgen_cuda_remap_image(handle_, pixel, 1, nullptr)); }
    // remap_image / par_remap_image                           mod.rs:30-34, 46-50
    void remap_image(RgbaImage& image) const { detail::check(lutgen_cuda_remap_image(handle_, image.data.data(), image.pixels(), nullptr)); }
    void par_remap_image(RgbaImage& image) const { remap_image(image); }
    // remap_image_with_interrupt / par_...                    mod.rs:36-42, 54-60
    void remap_image_with_interrupt(RgbaImage& image, const std::atomic<bool>& abort) const {
        detail::check(lutgen_cuda_remap_image(handle_, image.data.data(), image.pixels(), detail::flag_ptr(&abort)));
    }
    void par_remap_image_with_interrupt(RgbaImage& image, const std::atomic<bool>& abort) const { remap_image_with_interrupt(image, abort); }

    // GenerateLut::generate_lut / par_generate_lut            lib.rs:136-147
    RgbImage generate_lut(uint8_t level) const { return *generate(level, nullptr, true); }
    RgbImage par_generate_lut(uint8_t level) const { return *generate(level, nullptr, false); }
    // GenerateLut::*_with_interrupt -> Option<RgbImage>       lib.rs:149-170
    std::optional<RgbImage> generate_lut_with_interrupt(uint8_t level, const std::atomic<bool>& abort) const { return generate(level, &abort, true); }
    std::optional<RgbImage> par_generate_lut_with_interrupt(uint8_t level, const std::atomic<bool>& abort) const { return generate(level, &abort, false); }

    const lutgen_remapper* handle() const { return handle_; }

protected:
    InterpolatedRemapper() = default;
    void create(const lutgen_remapper_desc& d) {
        detail::ensure_init();
        detail::check(lutgen_cuda_remapper_create(&d, &handle_));
    }
    static lutgen_remapper_desc desc(int kind, const std::vector<std::array<uint8_t, 3>>& palette) {
        lutgen_remapper_desc d{};
        d.kind = kind;
        d.palette = palette.empty() ? nullptr : palette.front().data();
        d.palette_len = palette.size();
        d.lum_factor = 1.0;
        return d;
    }
    lutgen_remapper* handle_ = nullptr;

private:
    std::optional<RgbImage> generate(uint8_t level, const std::atomic<bool>* abort, bool sequential = false) const {
        bool ok = level >= 2 && level <= 33;
        uint32_t side = ok ? (uint32_t)level * level * level : 1;
        RgbImage out(side, side);
        const volatile uint8_t* flag = abort ? detail::flag_ptr(abort) : nullptr;
        // the sequential spellings differ from the par_* ones for GaussianBlurRemapper only (hint chain across rows)
        int rc = sequential ? lutgen_cuda_generate_lut_sequential(handle_, level, out.data.data(), flag)
                            : lutgen_cuda_generate_lut(handle_, level, out.data.data(), flag);
        if (detail::check(rc)) return std::nullopt;
        return out;
    }
};

using Palette = std::vector<std::array<uint8_t, 3>>;

// GaussianRemapper::new(palette, shape, nearest, lum_factor, preserve_lum)        rbf.rs:171-180
class GaussianRemapper : public InterpolatedRemapper {
public:
    GaussianRemapper(const Palette& palette, double shape, size_t nearest, double lum_factor, bool preserve_lum) {
        auto d = desc(LUTGEN_GAUSSIAN, palette);
        d.param = shape; d.nearest = nearest; d.lum_factor = lum_factor; d.preserve = preserve_lum;
        create(d);
    }
};

// ShepardRemapper::new(palette, power, nearest, lum_factor, preserve_lum)         rbf.rs:161-169
class ShepardRemapper : public InterpolatedRemapper {
public:
    ShepardRemapper(const Palette& palette, double power, size_t nearest, double lum_factor, bool preserve_lum) {
        auto d = desc(LUTGEN_SHEPARD, palette);
        d.param = power; d.nearest = nearest; d.lum_factor = lum_factor; d.preserve = preserve_lum;
        create(d);
    }
};

// LinearRemapper::new(palette, nearest, lum_factor, preserve_lum)                 rbf.rs:153-159
class LinearRemapper : public InterpolatedRemapper {
public:
    LinearRemapper(const Palette& palette, size_t nearest, double lum_factor, bool preserve_lum) {
        auto d = desc(LUTGEN_LINEAR, palette);
        d.nearest = nearest; d.lum_factor = lum_factor; d.preserve = preserve_lum;
        create(d);
    }
};

// NearestNeighborRemapper::new(palette, lum_factor, preserve)                     nearest_neighbor.rs:19-37
class NearestNeighborRemapper : public InterpolatedRemapper {
public:
    NearestNeighborRemapper(const Palette& palette, double lum_factor, bool preserve) {
        auto d = desc(LUTGEN_NEAREST, palette);
        d.lum_factor = lum_factor; d.preserve = preserve;
        create(d);
    }
};

// GaussianSamplingRemapper::new(palette, mean, std_dev, iterations, lum_factor, seed, preserve)   gaussian_sample.rs:27-46
class GaussianSamplingRemapper : public InterpolatedRemapper {
public:
    GaussianSamplingRemapper(const Palette& palette, double mean, double std_dev, size_t iterations, double lum_factor, uint64_t seed,
                             bool preserve) {
        auto d = desc(LUTGEN_SAMPLING, palette);
        d.mean = mean; d.std_dev = std_dev; d.iterations = iterations; d.lum_factor = lum_factor; d.seed = seed; d.preserve = preserve;
        create(d);
    }
};

// GaussianBlurRemapper::new(palette, radius, lum_factor, preserve)                gaussian_blur.rs:34-53
// The CLI's and Studio's default. Like the reference type it implements GenerateLut only
// (gaussian_blur.rs:512-541): the remap_* members are not exposed. generate_lut* follow
// generate_lut_inner (hint carried across rows), par_generate_lut* follow par_generate_lut_inner.
class GaussianBlurRemapper : private InterpolatedRemapper {
public:
    GaussianBlurRemapper(const Palette& palette, double radius, double lum_factor, bool preserve) {
        auto d = desc(LUTGEN_BLUR, palette);
        d.param = radius; d.lum_factor = lum_factor; d.preserve = preserve;
        create(d);
    }
    using InterpolatedRemapper::generate_lut;
    using InterpolatedRemapper::par_generate_lut;
    using InterpolatedRemapper::generate_lut_with_interrupt;
    using InterpolatedRemapper::par_generate_lut_with_interrupt;
};

}  // namespace interpolation
}  // namespace lutgen
