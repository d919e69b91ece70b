/*
 * lutgen_cuda.h -- C ABI of the B200 (sm_100a) implementation of lutgen's
 * Hald-CLUT generation / application hot path.
 *
 * This is the drop-in boundary: exactly the entry points a Rust `lutgen` crate
 * (crates/lib of ozwaldorf/lutgen-rs 0.15.0) would bind over FFI to move its hot
 * path to the GPU while keeping its public API. Each export cites the reference
 * interface it replaces (paths relative to the reference root). INTEGRATION.md
 * shows the Rust-side binding; include/lutgen.hpp is the C++ host mirror of the
 * crate API built on these symbols; lutgen_rs_b200/ is the Python (ctypes) one.
 *
 * Conventions
 *  - Plain pointers and sizes only. Pointers are HOST memory unless the
 *    function name ends in `_dev` (then: device memory of the current device).
 *  - Images are tightly packed, row-major, x fastest: RGB8 = 3 B/px (Hald CLUTs,
 *    `image::RgbImage`), RGBA8 = 4 B/px (`image::RgbaImage`). A level-L CLUT is
 *    L^3 x L^3 pixels = 3*L^6 bytes.
 *  - Return value: LUTGEN_OK (0), LUTGEN_ABORTED (1, the caller's abort flag was
 *    seen set -- the reference returns `None`), or a negative LUTGEN_ERR_*. The
 *    library never aborts the process where the reference panics; the message
 *    is available from lutgen_cuda_last_error() (thread local).
 *  - Multi-GPU, two ways. (1) ONE process, several GPUs -- what a CLI or Studio process is
 *    (crates/cli/src/main.rs:329-401): lutgen_cuda_init_devices, or LUTGEN_CUDA_DEVICES=0,1,..
 *    in the environment of an unchanged caller; lutgen_cuda_generate_lut deals the CLUT's
 *    row groups out to the GPUs, each copying its rows straight into the caller's image,
 *    lutgen_cuda_apply_hald_batch deals out the frames. Everything else runs on the first
 *    (primary) device. (2) One process per GPU (torch.distributed / MPI style): rows or
 *    tiles are sharded through the `_dev` entry points and exchanged by the kernel itself
 *    (lutgen_cuda_generate_lut_exchange_dev) or by the caller's collective.
 *  - There is no CPU fallback: without a usable sm_100 device every compute
 *    entry point returns LUTGEN_ERR_NO_DEVICE / LUTGEN_ERR_CUDA.
 *  - `stream` arguments are `cudaStream_t` passed as void*; 0 / NULL is CUDA's
 *    default stream, as in every CUDA API. `_dev` calls only enqueue work: they
 *    never synchronise, and results are ordered with the caller's other work on
 *    that stream. A remapper handle, and the apply entry points (which share
 *    one expanded-CLUT scratch buffer), must be driven from one stream at a time.
 */
#ifndef LUTGEN_CUDA_H
#define LUTGEN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LUTGEN_API
#else
#define LUTGEN_API __attribute__((visibility("default")))
#endif

/* ---- status codes ------------------------------------------------------- */
#define LUTGEN_OK 0
#define LUTGEN_ABORTED 1
#define LUTGEN_ERR_INVALID (-1)   /* bad argument (where the reference panics) */
#define LUTGEN_ERR_CUDA (-2)      /* CUDA runtime failure */
#define LUTGEN_ERR_NO_DEVICE (-3) /* no sm_100 device / library not initialised */
#define LUTGEN_ERR_UNSUPPORTED (-4)

/* ---- remapper kinds (crates/lib/src/interpolation/mod.rs:6-13) ----------- */
#define LUTGEN_GAUSSIAN 0 /* GaussianRemapper   rbf.rs:171-180  w = exp(-shape*d)      */
#define LUTGEN_SHEPARD 1  /* ShepardRemapper    rbf.rs:161-169  w = 1/sqrt(d)^power    */
#define LUTGEN_LINEAR 2   /* LinearRemapper     rbf.rs:153-159  w = d                  */
#define LUTGEN_NEAREST 3  /* NearestNeighborRemapper  nearest_neighbor.rs:11-61        */
#define LUTGEN_SAMPLING 4 /* GaussianSamplingRemapper gaussian_sample.rs:18-76         */
#define LUTGEN_BLUR 5     /* GaussianBlurRemapper     gaussian_blur.rs:26-542 (GenerateLut only;
                             param = radius; the CLI's and Studio's default algorithm)  */

/* Constructor arguments of every remapper, flattened.
 *   Gaussian/Shepard/Linear::new(palette, [shape|power], nearest, lum_factor, preserve_lum)  rbf.rs:126-140
 *   NearestNeighborRemapper::new(palette, lum_factor, preserve)                              nearest_neighbor.rs:19
 *   GaussianSamplingRemapper::new(palette, mean, std_dev, iterations, lum_factor, seed, preserve)  gaussian_sample.rs:27-35
 * Unused fields are ignored. The palette is copied (callers may free it). */
typedef struct lutgen_remapper_desc {
    int32_t kind;           /* LUTGEN_GAUSSIAN .. LUTGEN_SAMPLING */
    int32_t preserve;       /* preserve_lum / preserve (bool) */
    const uint8_t* palette; /* palette_len x [u8; 3] sRGB */
    uint64_t palette_len;
    double param;           /* Gaussian shape / Shepard power */
    uint64_t nearest;       /* 0 or >= palette_len: use every colour (rbf.rs:38) */
    double lum_factor;
    double mean, std_dev;   /* sampling */
    uint64_t iterations;    /* sampling */
    uint64_t seed;          /* sampling */
} lutgen_remapper_desc;

typedef struct lutgen_remapper lutgen_remapper; /* opaque; owns device copies */

/* ---- library ------------------------------------------------------------ */

/* Bind this process to CUDA device `device` (-1: $LUTGEN_CUDA_DEVICE, else
 * $LOCAL_RANK, else 0), create the stream, upload the constant tables.
 * Idempotent for the same device. Fails with LUTGEN_ERR_NO_DEVICE when there is
 * no device of compute capability 10.x. */
LUTGEN_API int lutgen_cuda_init(int device);
/* The same for a list of distinct devices (at most 8; devices[0] is the primary one): one context, stream set and
 * host worker thread per GPU. With more than one device lutgen_cuda_generate_lut and lutgen_cuda_apply_hald_batch
 * shard their work over all of them (GenerateLut::par_generate_lut, crates/lib/src/lib.rs:143-147, and the frame loop
 * of crates/cli/src/main.rs:878-888 are one call each in the reference: the sharding is invisible to the caller);
 * every other entry point uses the primary device. lutgen_cuda_init(-1) calls this with the devices named in
 * $LUTGEN_CUDA_DEVICES ("0,1,2,3") when that variable is set. */
LUTGEN_API int lutgen_cuda_init_devices(const int* devices, int ndevices);
/* Devices this process was initialised with (0 before lutgen_cuda_init*). */
LUTGEN_API int lutgen_cuda_device_count_in_use(int* count);
LUTGEN_API int lutgen_cuda_shutdown(void);
LUTGEN_API int lutgen_cuda_device_count(int* count);
LUTGEN_API const char* lutgen_cuda_last_error(void);
LUTGEN_API const char* lutgen_cuda_version(void);
/* Number of kernels this library has launched so far in this process. */
LUTGEN_API uint64_t lutgen_cuda_launch_count(void);

/* Pinned host buffers: host pointers handed to the non-_dev entry points are
 * copied by DMA directly when they come from here (or are otherwise
 * page-locked); pageable memory is staged through an internal pinned ring. */
LUTGEN_API int lutgen_cuda_host_alloc(size_t bytes, void** out);
LUTGEN_API int lutgen_cuda_host_free(void* p);

/* ---- identity.rs ---------------------------------------------------------- */

/* identity::generate(level) -> RgbImage            crates/lib/src/identity.rs:7-34
 * level in 2..=16 (the reference divides by zero below 2). out: 3*level^6 B. */
LUTGEN_API int lutgen_cuda_identity(uint8_t level, uint8_t* out_rgb);
LUTGEN_API int lutgen_cuda_identity_dev(uint8_t level, uint8_t* out_rgb_dev, void* stream);

/* identity::detect_level(&RgbImage) -> u8           identity.rs:85-99
 * LUTGEN_ERR_INVALID where the reference's assert_eq! panics. */
LUTGEN_API int lutgen_cuda_detect_level(uint32_t width, uint32_t height, uint8_t* level);

/* identity::correct_image_with_level(&mut RgbaImage, &RgbImage, level)   identity.rs:71-78
 * (= correct_pixel, identity.rs:40-52, per pixel). In place, alpha untouched,
 * no interpolation. npix may be 0. */
LUTGEN_API int lutgen_cuda_apply_hald(uint8_t* rgba, size_t npix, const uint8_t* lut_rgb, uint8_t level);
LUTGEN_API int lutgen_cuda_apply_hald_dev(uint8_t* rgba_dev, size_t npix, const uint8_t* lut_rgb_dev, uint8_t level,
                                          void* stream);
/* The CLI's frame loops (crates/cli/src/main.rs:852-886): one LUT, many images.
 * frames[i] has npix[i] RGBA8 pixels; uploads, kernels and downloads of
 * successive frames overlap on the copy engines. */
LUTGEN_API int lutgen_cuda_apply_hald_batch(uint8_t* const* frames, const size_t* npix, size_t nframes,
                                            const uint8_t* lut_rgb, uint8_t level);

/* identity::correct_pixel for n independent colours (identity.rs:40-52; the CLI's
 * `patch` calls it per colour, crates/cli/src/main.rs:1028-1042). in/out: n x [u8;3]. */
LUTGEN_API int lutgen_cuda_correct_pixels(const uint8_t* in_rgb, size_t n, const uint8_t* lut_rgb, uint8_t level,
                                          uint8_t* out_rgb);

/* ---- remappers ------------------------------------------------------------ */

LUTGEN_API int lutgen_cuda_remapper_create(const lutgen_remapper_desc* desc, lutgen_remapper** out);
LUTGEN_API int lutgen_cuda_remapper_destroy(lutgen_remapper* r);

/* GenerateLut::{generate_lut, par_generate_lut, *_with_interrupt}   crates/lib/src/lib.rs:113-171
 * identity -> Rgb->Rgba -> remap_image -> Rgba->Rgb fused into one pass; no
 * intermediate images. abort_flag may be NULL; when it reads non-zero before,
 * between or after the launches the call returns LUTGEN_ABORTED and out_rgb is
 * unspecified (the reference returns None, lib.rs:149-170). */
LUTGEN_API int lutgen_cuda_generate_lut(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb,
                                        const volatile uint8_t* abort_flag);

/* One process per GPU, CLUT wanted on the HOST: part `part_index` of `part_count` of the CLUT -- the row groups this
 * rank is dealt, round robin -- generated on this process's GPU and copied to their places in out_rgb, the WHOLE
 * image's base address (e.g. a buffer in shared memory that every rank has mapped, and page-locked with
 * cudaHostRegister for full PCIe speed). Every rank uses its own PCIe link; no GPU-to-GPU exchange is needed. The
 * image is complete when all ranks have returned. Same arithmetic as lutgen_cuda_generate_lut, so the bytes are those
 * of one GPU. (crates/lib/src/lib.rs:143-147 par_generate_lut, sharded by rows.) */
LUTGEN_API int lutgen_cuda_generate_lut_part(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb, uint32_t part_index, uint32_t part_count,
                                             const volatile uint8_t* abort_flag);

/* GenerateLut::generate_lut / generate_lut_with_interrupt, the SEQUENTIAL spellings (lib.rs:113-132). Identical to
 * lutgen_cuda_generate_lut for every remapper but GaussianBlurRemapper, whose sequential generate_lut_inner carries
 * the nearest-neighbour hint from the end of one (r, g) row into the next (gaussian_blur.rs:213) where
 * par_generate_lut_inner restarts it (l.339): the two differ in cells that lie within the early-exit threshold of
 * two palette colours. The par_* spellings bind lutgen_cuda_generate_lut. */
LUTGEN_API int lutgen_cuda_generate_lut_sequential(const lutgen_remapper* r, uint8_t level, uint8_t* out_rgb,
                                                   const volatile uint8_t* abort_flag);
/* Row shard of the same: rows [row_begin, row_end) of the level^3 x level^3 CLUT
 * are written at their place inside the full-size device image lut_rgb_dev
 * (3*level^6 B). This is the unit the multi-GPU path all-gathers. */
LUTGEN_API int lutgen_cuda_generate_lut_rows_dev(const lutgen_remapper* r, uint8_t level, uint8_t* lut_rgb_dev,
                                                 uint32_t row_begin, uint32_t row_end, void* stream);

/* Rows [row_begin, row_end) written to SEVERAL full-size CLUT images at once: luts_rgb_dev[0] is this
 * GPU's, the others are the same buffer of the other ranks mapped into this process (CUDA IPC below),
 * written over NVLink while the kernel computes. This is the fused form of "row shard + all-gather"
 * (reference loop: crates/lib/src/interpolation/mod.rs:46-50 over lib.rs:143-147): after a
 * barrier across the ranks every buffer holds the whole CLUT. RBF and NearestNeighbor remappers;
 * LUTGEN_ERR_UNSUPPORTED for the sampling and blur remappers (generate locally, then all-gather). */
LUTGEN_API int lutgen_cuda_generate_lut_rows_multi_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts,
                                                       uint32_t row_begin, uint32_t row_end, void* stream);

/* The same with BALANCED shards: this launch is shard `shard_index` of `shard_count` of the whole CLUT and
 * takes every shard_count-th tile hand-out of the warp-tile kernels (tiles differ a lot in cost; contiguous
 * row shards do not balance). Remappers that run on other kernels get the shard's contiguous rows. Every
 * destination must receive all shards (one launch per rank) before the CLUT is complete. */
LUTGEN_API int lutgen_cuda_generate_lut_shard_multi_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts,
                                                        uint32_t shard_index, uint32_t shard_count, void* stream);

/* The same with the destinations ALSO mapped as one NVSwitch multicast address (`multicast_lut`, e.g. from
 * torch's symmetric memory; NULL = none): whole tiles are stored once to that address and the switch
 * replicates them to every GPU, so a rank sends its share of the CLUT once instead of once per peer. Entries
 * that are stored byte by byte (tiles cut by the cube's edge, entries recomputed by the exact kernel) still
 * go to each of luts_rgb_dev[]. */
LUTGEN_API int lutgen_cuda_generate_lut_shard_mc_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts,
                                                     uint8_t* multicast_lut, uint32_t shard_index, uint32_t shard_count, void* stream);

/* The whole exchange in ONE launch: as lutgen_cuda_generate_lut_shard_mc_dev (shard `shard_index` of
 * `shard_count` = nluts, luts_rgb_dev[0] this GPU's copy), and the kernel's last thread block -- the one that
 * knows all of this GPU's stores to the peers have been issued -- publishes the rank's step number in every
 * rank's flag array and waits for all ranks' numbers in its own: when the kernel ends, the CLUT is complete
 * on this GPU. barrier_flags_dev_table: DEVICE array of nluts pointers, entry i = rank i's flag array of 16
 * u32 (zero before the first step, mapped into this process; slots 0..7 are written by the peers, word 8 is
 * the owner's step counter). Every rank must launch it once per step, all ranks' kernels running at the same
 * time (one GPU per rank). NULL: no barrier (the call above). Palettes of at most 32 colours, RBF /
 * NearestNeighbor remappers; others: LUTGEN_ERR_UNSUPPORTED (use rows + all-gather).
 * Mirrors GenerateLut::par_generate_lut (crates/lib/src/lib.rs:143-147) for one CLUT shared by G GPUs. */
LUTGEN_API int lutgen_cuda_generate_lut_exchange_dev(const lutgen_remapper* r, uint8_t level, uint8_t* const* luts_rgb_dev, int nluts,
                                                     uint8_t* multicast_lut, uint32_t shard_index, uint32_t shard_count,
                                                     unsigned* const* barrier_flags_dev_table, void* stream);

/* Barrier between the per-GPU processes through flags in peer memory. flags_dev_table: DEVICE array of
 * nranks pointers, entry i = rank i's flag array (>= nranks zero-initialised u32, mapped into this process);
 * epoch must grow by one per barrier. Orders this GPU's earlier writes to peer buffers before the peers'
 * later reads. Replaces the collective of the row-shard + all-gather scheme. */
LUTGEN_API int lutgen_cuda_peer_barrier_dev(unsigned* const* flags_dev_table, int nranks, int rank, unsigned epoch, void* stream);

/* Device memory that can be shared between the per-GPU processes of one node, and its CUDA IPC
 * handle (64 bytes) / mapping. Plumbing for the call above; no reference counterpart. */
LUTGEN_API int lutgen_cuda_device_alloc(size_t bytes, void** out);
LUTGEN_API int lutgen_cuda_device_free(void* p);
LUTGEN_API int lutgen_cuda_ipc_export(void* dev_ptr, unsigned char handle[64]);
LUTGEN_API int lutgen_cuda_ipc_import(const unsigned char handle[64], void** dev_ptr_out);
LUTGEN_API int lutgen_cuda_ipc_close(void* dev_ptr);

/* InterpolatedRemapper::{remap_image, par_remap_image, *_with_interrupt}
 * crates/lib/src/interpolation/mod.rs:23-61. RGBA8 in place, alpha untouched.
 * npix == 1 is InterpolatedRemapper::remap_pixel (mod.rs:25). */
LUTGEN_API int lutgen_cuda_remap_image(const lutgen_remapper* r, uint8_t* rgba, size_t npix,
                                       const volatile uint8_t* abort_flag);
LUTGEN_API int lutgen_cuda_remap_image_dev(const lutgen_remapper* r, uint8_t* rgba_dev, size_t npix, void* stream);

/* GaussianSamplingRemapper noise table: iterations x 4 f64 normals (r, g, b,
 * alpha order, gaussian_sample.rs:55-58), shared by every pixel because the
 * reference re-seeds its RNG per pixel (gaussian_sample.rs:52). `get` returns
 * the table the library drew with its own counter-based generator
 * (Philox4x32-10 + Box-Muller; the stream differs from rand's ChaCha12 +
 * ziggurat by design); `set` replaces it (test hook: lets a CPU oracle consume
 * the identical normals so results compare byte for byte). */
LUTGEN_API int lutgen_cuda_sampling_get_noise(const lutgen_remapper* r, double* noise_out);
LUTGEN_API int lutgen_cuda_sampling_set_noise(lutgen_remapper* r, const double* noise);

/* ---- PNG writer around the path -------------------------- */

/* `RgbImage::save(path)` / `RgbaImage::save(path)` with a .png path (crates/cli/src/main.rs:767, 811,
 * 839: `lut.save(&path)`; 862-871: `image.save(&path)`), i.e. the `image` crate's PNG encoder: the
 * bytes of a PNG file (8 bits per sample, colour type 0 / 4 / 2 / 6 for channels 1 / 2 / 3 / 4, no
 * interlace, adaptive row filters, one zlib stream split over several IDAT chunks) for a tightly
 * packed height x width x channels image. Lossless: any PNG reader returns the input pixels; the
 * bytes themselves differ from the `image` crate's, as two encoders' always do. The caller writes
 * them to the file.
 *
 * lutgen_cuda_png_bound: the largest size the encoder can produce for such an image.
 * lutgen_cuda_encode_png: host pixels in, host bytes out. *out_len receives the file size; if that
 *   exceeds `capacity`, nothing is written and LUTGEN_ERR_INVALID is returned (retry with *out_len,
 *   or pass lutgen_cuda_png_bound() up front).
 * lutgen_cuda_encode_png_dev: device pixels in, device bytes out, `capacity` must be at least
 *   lutgen_cuda_png_bound(); enqueues on `stream` and, unlike the other _dev calls, synchronises that
 *   stream once at the end to return the size in *out_len (host). */
LUTGEN_API int lutgen_cuda_png_bound(uint32_t width, uint32_t height, int channels, size_t* bound);
LUTGEN_API int lutgen_cuda_encode_png(const uint8_t* pixels, uint32_t width, uint32_t height, int channels, uint8_t* out,
                                      size_t capacity, size_t* out_len);
LUTGEN_API int lutgen_cuda_encode_png_dev(const uint8_t* pixels_dev, uint32_t width, uint32_t height, int channels,
                                          uint8_t* out_dev, size_t capacity, size_t* out_len, void* stream);

/* GenerateLut::par_generate_lut followed by `lut.save(path)` (crates/cli/src/main.rs:756-768, the
 * whole of `lutgen generate` after argument parsing): the CLUT is generated and encoded on the device
 * and only the PNG bytes cross PCIe. Same return values as lutgen_cuda_generate_lut; *out_len and
 * `capacity` as in lutgen_cuda_encode_png (bound: lutgen_cuda_png_bound(level^3, level^3, 3)). */
LUTGEN_API int lutgen_cuda_generate_lut_png(const lutgen_remapper* r, uint8_t level, uint8_t* out, size_t capacity,
                                            size_t* out_len, const volatile uint8_t* abort_flag);

/* identity::correct_image followed by `image.save(path)` (crates/cli/src/main.rs:852-871, what `lutgen apply`
 * does with a still image): the frame is uploaded once, corrected and encoded on the device, and only the PNG
 * bytes come back (an RGBA PNG, alpha as it went in). `rgba` is not modified. *out_len / `capacity` as in
 * lutgen_cuda_encode_png (bound: lutgen_cuda_png_bound(width, height, 4)). */
LUTGEN_API int lutgen_cuda_apply_hald_png(const uint8_t* rgba, uint32_t width, uint32_t height, const uint8_t* lut_rgb,
                                          uint8_t level, uint8_t* out, size_t capacity, size_t* out_len);

/* ---- PNG reader around the path (the decode half) -------------
 * `load_image` (crates/cli/src/main.rs:602-613: image::open(path), then to_rgba8() for the image and to_rgb8() for the
 * --hald-clut) for PNG files. The container is parsed on the host; inflate, scanline reconstruction and the channel
 * conversion run on the device. 8 bits per sample, colour types 0 / 2 / 3 / 4 / 6 (grey, RGB, palette with optional
 * tRNS, grey + alpha, RGBA), non-interlaced. Other flavours (16-bit samples, interlace, tRNS colour keys) return
 * LUTGEN_ERR_UNSUPPORTED -- keep the caller's decoder for those; a damaged file (signature, chunk CRC, deflate data,
 * Adler-32, filter type) returns LUTGEN_ERR_INVALID where the `image` crate returns an Err. Files written by this
 * library's own encoder (one IDAT chunk per 16 KB piece) are inflated in parallel, one thread per piece; a foreign
 * file's single deflate stream is inflated by one thread (correct, but no faster than a CPU). */

/* Geometry of a PNG file without decoding it: *channels as lutgen_cuda_decode_png(…, 0, …) would deliver them
 * (1, 2, 3 or 4; a palette image counts as 3, or 4 with tRNS); *has_alpha: the file carries transparency. */
LUTGEN_API int lutgen_cuda_png_info(const uint8_t* file, size_t len, uint32_t* width, uint32_t* height, int* channels, int* has_alpha);

/* PNG file bytes (host) -> pixels (host), tightly packed height x width x out_channels. out_channels: 3 = to_rgb8
 * (alpha dropped, grey replicated), 4 = to_rgba8 (alpha 255 where the file has none), 0 = as in the file, 1 / 2 only for
 * grey sources. `capacity` must hold width * height * channels bytes. */
LUTGEN_API int lutgen_cuda_decode_png(const uint8_t* file, size_t len, int out_channels, uint8_t* out, size_t capacity);
/* The same with the pixels left in device memory (the file bytes are still read from the host): enqueues on `stream`
 * and synchronises it (the inflate status has to be read back). */
LUTGEN_API int lutgen_cuda_decode_png_dev(const uint8_t* file, size_t len, int out_channels, uint8_t* out_dev, size_t capacity, void* stream);

/* `lutgen apply --hald-clut clut.png image.png` for PNG files, file to file (crates/cli/src/main.rs:832-871: load_image
 * twice, correct_image, image.save): both files are decoded on the device, the image is corrected and encoded there, and
 * only compressed bytes cross PCIe in either direction. The CLUT's level is detected from its size (identity.rs:85-99;
 * not a Hald CLUT -> LUTGEN_ERR_INVALID). The result is an RGBA PNG; *out_len / `capacity` as in lutgen_cuda_encode_png
 * (bound: lutgen_cuda_png_bound(width, height, 4) with the image's size from lutgen_cuda_png_info). */
LUTGEN_API int lutgen_cuda_apply_hald_png_file(const uint8_t* image_png, size_t image_len, const uint8_t* hald_png, size_t hald_len,
                                               uint8_t* out, size_t capacity, size_t* out_len);

/* ---- palette extraction next to the path ------------------- */

/* `lutgen extract` (crates/cli/src/main.rs:772-815): a palette of at most `color_count` (1..=256; the CLI's
 * argument is a u8) colours for an RGB8 image by k-means in Oklab -- the role of
 * quantette::PalettePipeline::{colorspace(Oklab), quantize_method(kmeans()), palette_size(n), palette_par()}.
 * NOT quantette's palette: that crate is not vendored under the reference and nothing there pins its output;
 * this is the deterministic k-means specified in csrc/extract_core.h (distinct colours weighted by pixel count,
 * weighted-maximin seeding, `iterations` (at most 1000) Lloyd passes in fixed-point integers), bit exact with its restatement
 * in oracle/. palette_rgb receives *count_out <= color_count colours (fewer when the image has fewer distinct
 * colours). */
LUTGEN_API int lutgen_cuda_extract_palette(const uint8_t* rgb, size_t npix, uint32_t color_count, uint32_t iterations,
                                           uint8_t* palette_rgb, uint32_t* count_out);

/* ---- oklab (third-party arithmetic on the path; oklab 1.1.2) --------------- */

/* oklab::srgb_to_oklab for n colours: in n x [u8;3], out n x [f32;3] (l, a, b). */
LUTGEN_API int lutgen_cuda_srgb_to_oklab(const uint8_t* rgb, size_t n, float* out_lab);
/* oklab::oklab_to_srgb for n colours: in n x [f32;3], out n x [u8;3]. */
LUTGEN_API int lutgen_cuda_oklab_to_srgb(const float* lab, size_t n, uint8_t* out_rgb);

/* ---- diagnostics ---------------------------------------------------------- */

/* Measured FP32 FMA throughput (TFLOP/s, FMA = 2) and MUFU throughput (T op/s)
 * of this GPU at the clocks it sustains: the roofline denominators for the
 * generation kernels (no counterpart in the reference). */
LUTGEN_API int lutgen_cuda_measure_peaks(double* fp32_tflops, double* mufu_tops);

/* The few-ulp f32 Oklab conversion the throughput kernels use for their tile bounds
 * (tests bound its distance to the exact lutgen_cuda_srgb_to_oklab). */
LUTGEN_API int lutgen_cuda_srgb_to_oklab_fast(const uint8_t* rgb, size_t n, float* out_lab);

/* With LUTGEN_CUDA_STATS=1 in the environment at init: per-tile histograms of the throughput
 * kernel's classification, out128[0..64) = number of UNCERTAIN colours, [64..128) = IN colours. */
LUTGEN_API int lutgen_cuda_debug_stats(uint32_t* out128, int reset);

/* With LUTGEN_CUDA_STATS=1: the timeline of the small-palette kernel's launches since the last reset, from the GPU's
 * nanosecond clock (low words): out4[0] = first thread block started, [1] = last warp out of tiles, [2] = last warp's
 * stores (to the peers too) complete, [3] = exchange barrier passed. Reset between launches to time one. */
LUTGEN_API int lutgen_cuda_debug_timeline(uint32_t* out4, int reset);

/* Measurement aid (tools/time_exchange.py): the exchange of lutgen_cuda_generate_lut_exchange_dev without its
 * arithmetic. The level's CLUT is taken as already present in luts_rgb_dev[0]; this GPU's share of it is sent to the
 * other destinations (or to the multicast address) and the barrier follows. pattern 0: whole strips (12 * level^2
 * contiguous bytes), 1: hand-outs of four tiles (96-byte rows). Needs level % 4 == 0 and 16-byte aligned buffers. */
LUTGEN_API int lutgen_cuda_debug_exchange_probe_dev(uint8_t level, uint8_t* const* luts_rgb_dev, int nluts, uint8_t* multicast_lut,
                                                    uint32_t shard_index, uint32_t shard_count, unsigned* const* barrier_flags_dev_table,
                                                    int pattern, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LUTGEN_CUDA_H */
