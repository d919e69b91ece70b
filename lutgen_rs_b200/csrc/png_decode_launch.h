// png_decode_launch.h -- host-side entry points of the PNG decoder kernels (png_decode.cu over png_decode_core.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace lutgen {

struct PngDecodePlan {
    uint32_t width, height;
    int color_type, channels;     // as in the file
    int out_channels;             // 1..4 as asked for (3 = to_rgb8, 4 = to_rgba8), or = channels (native)
    size_t zbytes;                // IDAT payloads, concatenated (one zlib stream)
    size_t stream_bytes;          // height * (1 + width * channels): the filtered scanlines
    uint32_t npieces;             // > 0: the chunk pattern of a file this library wrote (piece_off holds npieces + 1 offsets into z)
};

// Device scratch needed beside z and the output: the filtered stream, the unfiltered rows (when a conversion follows),
// Adler parts, band progress, status words.
size_t png_decode_scratch_bytes(const PngDecodePlan& plan);

// z_dev: the zlib stream; piece_off_dev: plan.npieces + 1 offsets (or nullptr); palette_dev: 256 x (r | g<<8 | b<<16 | a<<24)
// (colour type 3). out_dev receives height x width x out_channels bytes. status_host (page-locked, 4 ints) receives
// {inflate rc, adler ok, path (1 pieces, 2 serial), 0}; the call synchronises `s` once between inflate and the rest when it
// has to fall back from the piece path to the serial one. Returns the CUDA status; *launches = kernels launched.
cudaError_t png_decode_async(const PngDecodePlan& plan, const uint8_t* z_dev, const unsigned long long* piece_off_dev, const uint32_t* palette_dev,
                             uint8_t* out_dev, void* scratch_dev, int* status_host, cudaStream_t s, unsigned* launches);

// The IDAT payloads of a file in device memory into one contiguous stream: tab = nchunks + 1 pairs (offset in the file,
// offset in the stream), the last pair closing the stream.
void png_decode_gather(const uint8_t* file_dev, const unsigned long long* tab_dev, uint32_t nchunks, uint8_t* z_dev, cudaStream_t s);

}  // namespace lutgen
