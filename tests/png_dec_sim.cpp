// png_dec_sim.cpp -- host build of lutgen_rs_b200/csrc/png_decode_core.h for the CPU unit tests.
//
// TEST INFRASTRUCTURE ONLY: the same container parser, inflate (stream and piece mode), Adler-32 and per-pixel
// reconstruction the CUDA kernels of png_decode.cu run, driven by plain loops (reconstruction row by row, pixel by
// pixel: the order the wavefront kernel is equivalent to). tests/test_png_decode_core.py checks it against zlib and PIL.
// Nothing in the product path links or calls this.
#include <cstring>
#include <vector>

#include "../lutgen_rs_b200/csrc/png_decode_core.h"

namespace {
struct Span {
    size_t off, len;
};
}  // namespace

extern "C" {

// Raw zlib stream -> bytes. Returns the inflate status; *produced = bytes written.
int lpngd_sim_inflate(const uint8_t* z, unsigned long long n, uint8_t* dst, unsigned long long cap, unsigned long long* produced) {
    lpngd::InflateScratch X;
    size_t out = 0;
    int rc = lpngd::inflate_zlib(z, (size_t)n, dst, (size_t)cap, &out, X);
    *produced = out;
    return rc;
}

// info: width, height, colour type, channels, default output channels, pieces path used (1) or serial (2)
// Returns 0 ok, 1 not a PNG, 2 corrupt, 3 unsupported, < 0 inflate error, 10 Adler mismatch, 11 bad filter type.
int lpngd_sim_decode(const uint8_t* file, unsigned long long len, int out_channels, uint8_t* out, unsigned long long cap, unsigned* info6) {
    lpngd::Info info;
    std::vector<Span> idat;
    int prc = lpngd::png_parse(file, (size_t)len, info, idat);
    if (prc != lpngd::P_OK) return prc;
    const int oc = out_channels ? out_channels : (info.color_type == 3 ? info.out_channels : info.channels);
    info6[0] = info.width; info6[1] = info.height; info6[2] = (unsigned)info.color_type; info6[3] = (unsigned)info.channels; info6[4] = (unsigned)oc;
    if ((unsigned long long)info.width * info.height * (unsigned)oc > cap) return 2;
    std::vector<uint8_t> z;
    for (const Span& c : idat) z.insert(z.end(), file + c.off, file + c.off + c.len);
    const size_t rb = (size_t)info.width * info.channels, stream_bytes = (size_t)info.height * (rb + 1);
    std::vector<uint8_t> stream(stream_bytes);
    const size_t npieces = (stream_bytes + lpngd::PIECE - 1) / lpngd::PIECE;
    bool done = false;
    info6[5] = 2;
    if (idat.size() == npieces + 1 && idat.back().len == 4 && idat.front().len >= 2) {
        // the piece path, exactly as png_decode.cu tries it
        bool ok = true;
        size_t pos = 0;
        lpngd::InflateScratch X;
        for (size_t i = 0; i < npieces && ok; ++i) {
            const size_t o0 = i * (size_t)lpngd::PIECE, expect = stream_bytes - o0 < (size_t)lpngd::PIECE ? stream_bytes - o0 : (size_t)lpngd::PIECE;
            const size_t skip = i == 0 ? 2 : 0;
            ok = lpngd::inflate_piece(z.data() + pos + skip, idat[i].len - skip, stream.data() + o0, expect, X) == lpngd::OK;
            pos += idat[i].len;
        }
        if (ok) {
            uint32_t A = 1, B = 0, a, b;
            for (size_t o = 0; o < stream_bytes; o += 4096) {
                const size_t k = stream_bytes - o < 4096 ? stream_bytes - o : 4096;
                lpngd::adler_span(stream.data() + o, k, a, b);
                lpngd::adler_append(A, B, a, b, k);
            }
            ok = ((B << 16) | A) == lpngd::be32(z.data() + z.size() - 4);
        }
        done = ok;
        if (done) info6[5] = 1;
    }
    if (!done) {
        lpngd::InflateScratch X;
        size_t produced = 0;
        int rc = lpngd::inflate_zlib(z.data(), z.size(), stream.data(), stream_bytes, &produced, X);
        if (rc != lpngd::OK) return rc;
        if (produced != stream_bytes) return lpngd::ERR_LENGTH;
        uint32_t A = 1, B = 0, a, b;
        lpngd::adler_span(stream.data(), stream_bytes, a, b);
        lpngd::adler_append(A, B, a, b, stream_bytes);
        if (((B << 16) | A) != lpngd::be32(z.data() + z.size() - 4)) return 10;
    }
    std::vector<uint8_t> rows((size_t)info.height * rb);
    const uint32_t bpp = (uint32_t)info.channels;
    for (uint32_t y = 0; y < info.height; ++y) {
        const uint8_t* src = stream.data() + (size_t)y * (rb + 1);
        const uint32_t ft = src[0];
        if (ft > 4) return 11;
        ++src;
        uint8_t* dst = rows.data() + (size_t)y * rb;
        const uint8_t* up = y ? dst - rb : nullptr;
        uint32_t a = 0, c = 0;
        for (uint32_t x = 0; x < info.width; ++x) {
            uint32_t xv = 0, b = 0;
            for (uint32_t k = 0; k < bpp; ++k) {
                xv |= (uint32_t)src[(size_t)x * bpp + k] << (8 * k);
                if (up) b |= (uint32_t)up[(size_t)x * bpp + k] << (8 * k);
            }
            // the wavefront kernel's branch-free form on odd rows, the plain one on even rows: both must agree with PIL
            const uint32_t mA = ft == 1 ? 255u : 0u, mB = ft == 2 ? 255u : 0u, mV = ft == 3 ? 255u : 0u, mP = ft == 4 ? 255u : 0u;
            uint32_t cur;
            if (!(y & 1)) cur = lpngd::unfilter_pixel(ft, xv, a, b, c, bpp);
            else if (bpp == 1) cur = lpngd::unfilter_pixel_masked<1>(xv, a, b, c, mA, mB, mV, mP);
            else if (bpp == 2) cur = lpngd::unfilter_pixel_masked<2>(xv, a, b, c, mA, mB, mV, mP);
            else if (bpp == 3) cur = lpngd::unfilter_pixel_masked<3>(xv, a, b, c, mA, mB, mV, mP);
            else cur = lpngd::unfilter_pixel_masked<4>(xv, a, b, c, mA, mB, mV, mP);
            for (uint32_t k = 0; k < bpp; ++k) dst[(size_t)x * bpp + k] = (uint8_t)(cur >> (8 * k));
            a = cur;
            c = b;
        }
    }
    const size_t npix = (size_t)info.width * info.height;
    for (size_t i = 0; i < npix; ++i) {
        const uint8_t* p = rows.data() + i * bpp;
        uint32_t r, g, b, al = 255;
        if (info.color_type == 3) { uint32_t e = info.palette[p[0]]; r = e & 255; g = (e >> 8) & 255; b = (e >> 16) & 255; al = e >> 24; }
        else if (info.color_type == 0) { r = g = b = p[0]; }
        else if (info.color_type == 4) { r = g = b = p[0]; al = p[1]; }
        else if (info.color_type == 2) { r = p[0]; g = p[1]; b = p[2]; }
        else { r = p[0]; g = p[1]; b = p[2]; al = p[3]; }
        uint8_t* q = out + i * (size_t)oc;
        if (oc == 1) q[0] = (uint8_t)r;
        else if (oc == 2) { q[0] = (uint8_t)r; q[1] = (uint8_t)al; }
        else { q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b; if (oc == 4) q[3] = (uint8_t)al; }
    }
    return 0;
}

}  // extern "C"
