// microbench_gather.cu -- which gather path feeds the Hald-CLUT apply kernel fastest?
// Standalone experiment (not part of the library). Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/mb_gather tools/microbench_gather.cu
//   tools/_build/mb_gather
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lut_index(uint32_t p, uint32_t cs, uint32_t csm1) {
    uint32_t r = (p & 255u) * csm1 / 255u, g = ((p >> 8) & 255u) * csm1 / 255u, b = ((p >> 16) & 255u) * csm1 / 255u;
    return (b * cs + g) * cs + r;
}

// 0: pure in-place stream (HBM ceiling for this pattern)
__global__ void k_stream(uint4* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) { uint4 p = v[i]; p.x ^= 0x010101u; p.y ^= 0x010101u; p.z ^= 0x010101u; p.w ^= 0x010101u; v[i] = p; }
}

// A: LDG gather from a 4-byte-per-entry table
template <int UNROLL>
__global__ void __launch_bounds__(256) k_ldg(uint4* v, size_t n, const uint32_t* __restrict__ lut, uint32_t cs, uint32_t csm1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st * UNROLL) {
        uint4 p[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) if (i + u * st < n) p[u] = v[i + u * st];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) if (i + u * st < n) {
            uint32_t a = __ldg(lut + lut_index(p[u].x, cs, csm1)), b = __ldg(lut + lut_index(p[u].y, cs, csm1));
            uint32_t c = __ldg(lut + lut_index(p[u].z, cs, csm1)), d = __ldg(lut + lut_index(p[u].w, cs, csm1));
            p[u].x = (a & 0xffffffu) | (p[u].x & 0xff000000u); p[u].y = (b & 0xffffffu) | (p[u].y & 0xff000000u);
            p[u].z = (c & 0xffffffu) | (p[u].z & 0xff000000u); p[u].w = (d & 0xffffffu) | (p[u].w & 0xff000000u);
            v[i + u * st] = p[u];
        }
    }
}

// B: texture fetch (linear 1D texture over the same table)
__global__ void __launch_bounds__(256) k_tex1d(uint4* v, size_t n, cudaTextureObject_t tex, uint32_t cs, uint32_t csm1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        uint4 p = v[i];
        uint32_t a = tex1Dfetch<uint32_t>(tex, lut_index(p.x, cs, csm1)), b = tex1Dfetch<uint32_t>(tex, lut_index(p.y, cs, csm1));
        uint32_t c = tex1Dfetch<uint32_t>(tex, lut_index(p.z, cs, csm1)), d = tex1Dfetch<uint32_t>(tex, lut_index(p.w, cs, csm1));
        p.x = (a & 0xffffffu) | (p.x & 0xff000000u); p.y = (b & 0xffffffu) | (p.y & 0xff000000u);
        p.z = (c & 0xffffffu) | (p.z & 0xff000000u); p.w = (d & 0xffffffu) | (p.w & 0xff000000u);
        v[i] = p;
    }
}

// C: 3D texture (cudaArray, block-linear), point sampled with integer coords
__global__ void __launch_bounds__(256) k_tex3d(uint4* v, size_t n, cudaTextureObject_t tex, uint32_t csm1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        uint4 p = v[i];
        uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float r = (float)((w[k] & 255u) * csm1 / 255u), g = (float)(((w[k] >> 8) & 255u) * csm1 / 255u), b = (float)(((w[k] >> 16) & 255u) * csm1 / 255u);
            uchar4 t = tex3D<uchar4>(tex, r + 0.5f, g + 0.5f, b + 0.5f);
            w[k] = (uint32_t)t.x | ((uint32_t)t.y << 8) | ((uint32_t)t.z << 16) | (w[k] & 0xff000000u);
        }
        v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// D: whole table in shared memory (only when it fits: 4*cs^3 <= 200 KB, i.e. level <= 6)
__global__ void __launch_bounds__(1024) k_smem(uint4* v, size_t n, const uint32_t* __restrict__ lut, uint32_t entries, uint32_t cs, uint32_t csm1) {
    extern __shared__ uint32_t s_lut[];
    for (uint32_t j = threadIdx.x; j < entries; j += blockDim.x) s_lut[j] = lut[j];
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        uint4 p = v[i];
        uint32_t a = s_lut[lut_index(p.x, cs, csm1)], b = s_lut[lut_index(p.y, cs, csm1)];
        uint32_t c = s_lut[lut_index(p.z, cs, csm1)], d = s_lut[lut_index(p.w, cs, csm1)];
        p.x = (a & 0xffffffu) | (p.x & 0xff000000u); p.y = (b & 0xffffffu) | (p.y & 0xff000000u);
        p.z = (c & 0xffffffu) | (p.z & 0xff000000u); p.w = (d & 0xffffffu) | (p.w & 0xff000000u);
        v[i] = p;
    }
}

// E: CLUT-stationary. The level-8 CLUT (262144 entries) is split into NSLAB blue slabs of 3 B/entry
// held in shared memory; every CTA streams the WHOLE batch but rewrites only the pixels whose blue
// index falls in its slab. The batch is read NSLAB times (L2 absorbs the re-reads), written once.
template <int NSLAB>
__global__ void __launch_bounds__(1024) k_slab(uint32_t* px, size_t npix, const uint32_t* __restrict__ lut, uint32_t cs, uint32_t csm1) {
    extern __shared__ uint32_t s_raw[];
    uint8_t* s3 = reinterpret_cast<uint8_t*>(s_raw);
    const uint32_t slab = blockIdx.x % NSLAB, bper = cs / NSLAB, b_lo = slab * bper;
    const uint32_t per_slab = bper * cs * cs;
    for (uint32_t j = threadIdx.x; j < per_slab; j += blockDim.x) {
        uint32_t w = lut[b_lo * cs * cs + j];
        s3[3 * j] = (uint8_t)w; s3[3 * j + 1] = (uint8_t)(w >> 8); s3[3 * j + 2] = (uint8_t)(w >> 16);
    }
    __syncthreads();
    const size_t group = blockIdx.x / NSLAB, ngroups = gridDim.x / NSLAB;
    uint4* v = reinterpret_cast<uint4*>(px);
    size_t nvec = npix / 4;
    for (size_t i = group * blockDim.x + threadIdx.x; i < nvec; i += ngroups * blockDim.x) {
        uint4 p = v[i];
        uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t b = ((w[k] >> 16) & 255u) * csm1 / 255u;
            uint32_t bl = b - b_lo;
            if (bl < bper) {
                uint32_t r = (w[k] & 255u) * csm1 / 255u, g = ((w[k] >> 8) & 255u) * csm1 / 255u;
                const uint8_t* e = s3 + 3u * ((bl * cs + g) * cs + r);
                uint32_t o = (uint32_t)e[0] | ((uint32_t)e[1] << 8) | ((uint32_t)e[2] << 16) | (w[k] & 0xff000000u);
                px[4 * i + k] = o;
            }
        }
    }
}

// F: CLUT distributed over the shared memory of a thread-block cluster (DSMEM): CTA `rank` of the
// cluster holds entries [rank*per, (rank+1)*per); gathers read the owner's shared memory directly.
template <int CL>
__global__ void __launch_bounds__(1024) k_dsmem(uint4* v, size_t n, const uint32_t* __restrict__ lut, uint32_t entries, uint32_t cs, uint32_t csm1) {
    extern __shared__ uint32_t s_lut[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), per = entries / CL;
    for (uint32_t j = threadIdx.x; j < per; j += blockDim.x) s_lut[j] = lut[rank * per + j];
    cluster.sync();
    const uint32_t* base[CL];
#pragma unroll
    for (int c = 0; c < CL; ++c) base[c] = cluster.map_shared_rank(s_lut, c);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        uint4 p = v[i];
        uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t idx = lut_index(w[k], cs, csm1);
            uint32_t owner = idx / per, off = idx - owner * per;
            const uint32_t* b = cluster.map_shared_rank(s_lut, owner);
            w[k] = (b[off] & 0xffffffu) | (w[k] & 0xff000000u);
        }
        v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    cluster.sync();
}

template <int CL>
static void launch_dsmem(uint4* v, size_t n, const uint32_t* lut, uint32_t entries, uint32_t cs, int sms, int threads) {
    size_t smem = (size_t)entries / CL * 4;
    CK(cudaFuncSetAttribute(k_dsmem<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) CK(cudaFuncSetAttribute(k_dsmem<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((sms / CL) * CL); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_dsmem<CL>, v, n, lut, entries, cs, cs - 1));
}

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    int frames = argc > 1 ? atoi(argv[1]) : 32;
    const size_t W = 3840, H = 2160, npix = (size_t)frames * W * H;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, batch %d x 4K frames = %.1f MB\n", prop.name, sms, frames, npix * 4 / 1e6);
    std::vector<uint32_t> rnd(npix), smooth(npix);
    uint64_t s = 42080085;
    for (size_t i = 0; i < npix; i += 2) { uint64_t z = sm64(s); rnd[i] = (uint32_t)z; if (i + 1 < npix) rnd[i + 1] = (uint32_t)(z >> 32); }
    for (size_t i = 0; i < npix; ++i) {  // gradient + +-4 noise: photo-like locality
        size_t x = i % W, y = (i / W) % H; uint32_t nz = (uint32_t)(sm64(s) >> 40);
        int r = (int)(x * 255 / W) + (int)(nz & 7) - 4, g = (int)(y * 255 / H) + (int)((nz >> 3) & 7) - 4, b = (int)((x + y) * 255 / (W + H)) + (int)((nz >> 6) & 7) - 4;
        r = r < 0 ? 0 : r > 255 ? 255 : r; g = g < 0 ? 0 : g > 255 ? 255 : g; b = b < 0 ? 0 : b > 255 ? 255 : b;
        smooth[i] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | 0xff000000u;
    }
    uint32_t *d_px; CK(cudaMalloc(&d_px, npix * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int level : {8}) {
        uint32_t cs = level * level, entries = cs * cs * cs;
        std::vector<uint32_t> lut(entries);
        for (uint32_t j = 0; j < entries; ++j) lut[j] = (uint32_t)sm64(s) & 0xffffffu;
        uint32_t* d_lut; CK(cudaMalloc(&d_lut, (size_t)entries * 4)); CK(cudaMemcpy(d_lut, lut.data(), (size_t)entries * 4, cudaMemcpyHostToDevice));
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_lut; rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); rd.res.linear.sizeInBytes = (size_t)entries * 4;
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType; cudaTextureObject_t tex1 = 0;
        bool have_tex1 = ((size_t)entries <= (1u << 27)) && cudaCreateTextureObject(&tex1, &rd, &td, nullptr) == cudaSuccess;
        cudaArray_t arr = nullptr; cudaTextureObject_t tex3 = 0; bool have_tex3 = false;
        {
            cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
            if (cudaMalloc3DArray(&arr, &cd, make_cudaExtent(cs, cs, cs)) == cudaSuccess) {
                cudaMemcpy3DParms cp{}; cp.srcPtr = make_cudaPitchedPtr(lut.data(), cs * 4, cs, cs); cp.dstArray = arr; cp.extent = make_cudaExtent(cs, cs, cs); cp.kind = cudaMemcpyHostToDevice;
                CK(cudaMemcpy3D(&cp));
                cudaResourceDesc r3{}; r3.resType = cudaResourceTypeArray; r3.res.array.array = arr;
                cudaTextureDesc t3{}; t3.readMode = cudaReadModeElementType; t3.filterMode = cudaFilterModePoint; t3.addressMode[0] = t3.addressMode[1] = t3.addressMode[2] = cudaAddressModeClamp;
                have_tex3 = cudaCreateTextureObject(&tex3, &r3, &t3, nullptr) == cudaSuccess;
            } else cudaGetLastError();
        }
        for (int data = 0; data < 2; ++data) {
            const std::vector<uint32_t>& src = data ? smooth : rnd;
            auto run = [&](const char* name, auto launch) {
                float best = 1e9f;
                for (int rep = 0; rep < 4; ++rep) {
                    CK(cudaMemcpy(d_px, src.data(), npix * 4, cudaMemcpyHostToDevice));
                    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
                }
                printf("level %2d %-7s %-22s %8.3f ms  %7.1f Gpix/s  %7.1f GB/s(8B/px)\n", level, data ? "smooth" : "random", name, best, npix / best / 1e6, npix * 8 / best / 1e6);
            };
            size_t nvec = npix / 4;
            if (level == 6 && data == 0) run("stream (no gather)", [&] { k_stream<<<sms * 16, 256>>>((uint4*)d_px, nvec); });
            for (int mult : {64}) { char nm[64]; snprintf(nm, 64, "ldg x1 grid=%dxSM", mult); run(nm, [&] { k_ldg<1><<<sms * mult, 256>>>((uint4*)d_px, nvec, d_lut, cs, cs - 1); }); }
            run("ldg x2 grid=16xSM", [&] { k_ldg<2><<<sms * 16, 256>>>((uint4*)d_px, nvec, d_lut, cs, cs - 1); });
            run("ldg x4 grid=8xSM", [&] { k_ldg<4><<<sms * 8, 256>>>((uint4*)d_px, nvec, d_lut, cs, cs - 1); });
            if (have_tex1) run("tex1Dfetch", [&] { k_tex1d<<<sms * 16, 256>>>((uint4*)d_px, nvec, tex1, cs, cs - 1); });
            if (have_tex3) run("tex3D point", [&] { k_tex3d<<<sms * 16, 256>>>((uint4*)d_px, nvec, tex3, cs - 1); });
            if ((size_t)entries * 4 <= 200 * 1024) {
                CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, entries * 4));
                run("smem whole table", [&] { k_smem<<<sms, 1024, entries * 4>>>((uint4*)d_px, nvec, d_lut, entries, cs, cs - 1); });
            }
            if (level == 8) {
                run("dsmem cluster 8 x1024", [&] { launch_dsmem<8>((uint4*)d_px, nvec, d_lut, entries, cs, sms, 1024); });
                run("dsmem cluster 8 x512", [&] { launch_dsmem<8>((uint4*)d_px, nvec, d_lut, entries, cs, sms, 512); });
                run("dsmem cluster 16 x1024", [&] { launch_dsmem<16>((uint4*)d_px, nvec, d_lut, entries, cs, sms, 1024); });
            }
            if (level == 8 && false) {
                CK(cudaFuncSetAttribute(k_slab<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * entries / 4));
                run("slab x4 (stationary)", [&] { k_slab<4><<<(sms / 4) * 4, 1024, 3 * entries / 4>>>(d_px, npix, d_lut, cs, cs - 1); });
            }
        }
        if (have_tex1) cudaDestroyTextureObject(tex1);
        if (have_tex3) cudaDestroyTextureObject(tex3);
        if (arr) cudaFreeArray(arr);
        cudaFree(d_lut);
    }
    return 0;
}
