// blur_launch.h -- host-side entry points of the GaussianBlurRemapper kernels (kernels_blur.cuh).
//
// Those kernels live in their own translation unit (blur.cu) compiled with --fmad=false: the
// reference sums `kw * src` with a rounded multiply and a rounded add, and nothing in this unit may
// be contracted. (For the packed f32x2 forms that flag is not enough -- ptxas fuses mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 regardless -- so the kernels add through blur_add2, see kernels_blur.cuh.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lutgen {

struct ColorTables;

constexpr int BLUR_ROW_WARPS = 8;
constexpr int BLUR_TR = 32, BLUR_TB = 32;   // k_blur_to_lut's tile: r x b cells of one g
constexpr int BLUR_MAX_DYN_SMEM = 226 * 1024;   // k_blur_axis: dynamic shared memory (the tiles) next to 16 bytes of static (the mbarriers)

// One blur pass along one axis of the [r][g][b][ch] volume (see k_blur_axis).
struct BlurAxis {
    uint32_t size, pstride, cols, pitch, cpl, run_stride, total_cols, groups, outer_stride;
    uint32_t bulk;   // every tile row is a 16-byte aligned run of floats on both sides: staged by cp.async.bulk (blur_launch_axis decides)
};

cudaError_t blur_set_attributes();
void blur_launch_palette(unsigned grid, cudaStream_t s, const uint8_t* pal_rgb4, uint32_t n, float lum, const float* decode, float* pal3, float* out3);
void blur_launch_nn_volume(unsigned grid, size_t smem, cudaStream_t s, float* colors, uint32_t size, uint32_t ch, const uint8_t* chan, const float* pal3,
                           const float* out3, uint32_t n, float lum, float threshold_sq, const ColorTables* tables, int stage_palette,
                           const uint32_t* end_prev, uint32_t* end_cur, uint32_t* changed);
void blur_launch_axis(unsigned grid, size_t smem, cudaStream_t s, const float* src, float* dst, const BlurAxis& P, uint32_t ntiles, const float* kernel,
                      int klen);
void blur_launch_to_lut(unsigned grid, cudaStream_t s, const float* colors, uint32_t size, uint32_t ch, const uint8_t* chan, uint8_t* out_rgb,
                        unsigned long long begin, unsigned long long count, uint32_t b_first, uint32_t tiles_r, uint32_t ntiles, const ColorTables* tables);

}  // namespace lutgen
