// blur.cu -- GaussianBlurRemapper kernels, compiled with --fmad=false (see blur_launch.h).
#include <cstdlib>

#include "blur_launch.h"
#include "kernels_blur.cuh"

namespace lutgen {

cudaError_t blur_set_attributes() {
    cudaError_t e = cudaFuncSetAttribute(k_blur_nn_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_blur_axis<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLUR_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_blur_axis<BLUR_PITCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLUR_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_blur_axis<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLUR_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_blur_axis<BLUR_PITCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLUR_MAX_DYN_SMEM);
}

void blur_launch_palette(unsigned grid, cudaStream_t s, const uint8_t* pal_rgb4, uint32_t n, float lum, const float* decode, float* pal3, float* out3) {
    k_blur_palette<<<grid, 128, 0, s>>>(pal_rgb4, n, lum, decode, pal3, out3);
}

void blur_launch_nn_volume(unsigned grid, size_t smem, cudaStream_t s, float* colors, uint32_t size, uint32_t ch, const uint8_t* chan, const float* pal3,
                           const float* out3, uint32_t n, float lum, float threshold_sq, const ColorTables* tables, int stage_palette,
                           const uint32_t* end_prev, uint32_t* end_cur, uint32_t* changed) {
    k_blur_nn_volume<<<grid, BLUR_ROW_WARPS * 32, smem, s>>>(colors, size, ch, chan, pal3, out3, n, lum, threshold_sq, tables, stage_palette, 1.0f,
                                                             end_prev, end_cur, changed);
}

void blur_launch_axis(unsigned grid, size_t smem, cudaStream_t s, const float* src, float* dst, const BlurAxis& P_in, uint32_t ntiles, const float* kernel,
                      int klen) {
    BlurAxis P = P_in;
    // bulk staging: rows of contiguous columns, every row start and length a multiple of 16 bytes on both sides
    // (the tile rows are: pitch * 4 and half * pitch * 4 are multiples of 16 when pitch is a multiple of 4)
    P.bulk = P.cpl == 0 && (P.cols & 3u) == 0 && (P.total_cols & 3u) == 0 && (P.pstride & 3u) == 0 && (P.outer_stride & 3u) == 0 && (P.pitch & 3u) == 0 &&
             (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && !getenv("LUTGEN_CUDA_BLUR_NO_BULK");
    if (P.pitch == BLUR_PITCH) {
        if (P.bulk) k_blur_axis<BLUR_PITCH, true><<<grid, 256, smem, s>>>(src, dst, P, ntiles, kernel, klen, 1.0f);
        else k_blur_axis<BLUR_PITCH, false><<<grid, 256, smem, s>>>(src, dst, P, ntiles, kernel, klen, 1.0f);
    } else {
        if (P.bulk) k_blur_axis<0, true><<<grid, 256, smem, s>>>(src, dst, P, ntiles, kernel, klen, 1.0f);
        else k_blur_axis<0, false><<<grid, 256, smem, s>>>(src, dst, P, ntiles, kernel, klen, 1.0f);
    }
}

void blur_launch_to_lut(unsigned grid, cudaStream_t s, const float* colors, uint32_t size, uint32_t ch, const uint8_t* chan, uint8_t* out_rgb,
                        unsigned long long begin, unsigned long long count, uint32_t b_first, uint32_t tiles_r, uint32_t ntiles, const ColorTables* tables) {
    if (ch == 2) k_blur_to_lut<2><<<grid, 256, 0, s>>>(colors, size, chan, out_rgb, begin, count, b_first, tiles_r, ntiles, tables);
    else k_blur_to_lut<3><<<grid, 256, 0, s>>>(colors, size, chan, out_rgb, begin, count, b_first, tiles_r, ntiles, tables);
}

}  // namespace lutgen
