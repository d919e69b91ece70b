// png_launch.h -- host-side entry points of the PNG encoder kernels (png.cu over png_core.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace lutgen {

// Upper bound of the encoded size of a width x height image with `bpp` bytes per pixel.
unsigned long long png_bound(uint32_t width, uint32_t height, uint32_t bpp);

// Device scratch the encoder needs for such an image (file size, row filters, per-piece slots, meta data, offsets).
size_t png_scratch_bytes(uint32_t width, uint32_t height, uint32_t bpp);

// One-time set-up (CRC tables in device memory, shared-memory attribute). Idempotent.
cudaError_t png_prepare();
void png_release();

// Encodes pixels_dev (height x width x bpp, tightly packed, device memory) into out_dev, which must hold
// png_bound() bytes. Four launches on `s`; the file size lands in the first 8 bytes of scratch_dev. Asynchronous.
cudaError_t png_encode_async(const uint8_t* pixels_dev, uint32_t width, uint32_t height, uint32_t bpp, uint8_t* out_dev, void* scratch_dev,
                             cudaStream_t s, unsigned* launches);

}  // namespace lutgen
