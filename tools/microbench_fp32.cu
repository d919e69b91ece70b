// FP32 issue / pipe throughput on this GPU for the instruction mixes the generation and blur kernels are made of
// (tools; not part of the library). For each mix: warp instructions per clock and SMSP, and lane-operations per clock
// and SM, at the SM clock measured by clock64 inside the kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_fp32 tools/microbench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
typedef float2 f2;

template <int MIX>
__global__ void __launch_bounds__(256) k_mix(float* sink, int iters, float a, float b, float one, unsigned long long* cycles) {
    f2 x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = make_float2(threadIdx.x + c, threadIdx.x - c); y[c] = make_float2(1.f + c, 2.f + c); }
    const f2 A = make_float2(a, a), B = make_float2(b, b), ONE = make_float2(one, one);
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MIX == 0) {          // scalar FFMA x2 (two independent scalars per chain)
                    x[c].x = fmaf(x[c].x, a, b); x[c].y = fmaf(x[c].y, a, b);
                } else if (MIX == 1) {   // FFMA2
                    x[c] = __ffma2_rn(x[c], A, B);
                } else if (MIX == 2) {   // blur: FMUL2 + FFMA2(s, 1, prod)
                    f2 p = __fmul2_rn(A, x[(c + 1) % CHAINS]);
                    x[c] = __ffma2_rn(x[c], ONE, p);
                } else if (MIX == 3) {   // scalar FMUL + FADD, not fused (two scalars per chain)
                    float p0 = __fmul_rn(a, x[(c + 1) % CHAINS].x), p1 = __fmul_rn(a, x[(c + 1) % CHAINS].y);
                    x[c].x = __fadd_rn(x[c].x, p0); x[c].y = __fadd_rn(x[c].y, p1);
                } else if (MIX == 4) {   // half packed, half scalar: chains 0..3 packed blur, 4..7 scalar blur
                    if (c < CHAINS / 2) {
                        f2 p = __fmul2_rn(A, x[(c + 1) % CHAINS]);
                        x[c] = __ffma2_rn(x[c], ONE, p);
                    } else {
                        float p0 = __fmul_rn(a, x[(c + 1) % CHAINS].x), p1 = __fmul_rn(a, x[(c + 1) % CHAINS].y);
                        x[c].x = __fadd_rn(x[c].x, p0); x[c].y = __fadd_rn(x[c].y, p1);
                    }
                } else if (MIX == 5) {   // FFMA2 + scalar FFMA 1:1 (instructions)
                    x[c] = __ffma2_rn(x[c], A, B);
                    y[c].x = fmaf(y[c].x, a, b);
                } else if (MIX == 6) {   // FFMA2 + FMNMX 1:1
                    x[c] = __ffma2_rn(x[c], A, B);
                    y[c].x = __uint_as_float(__funnelshift_l(__float_as_uint(y[c].x), __float_as_uint(x[c].y), 3));
                } else if (MIX == 7) {   // FFMA2 + integer ALU (LOP3) 1:1
                    x[c] = __ffma2_rn(x[c], A, B);
                    y[c].x = __uint_as_float(__float_as_uint(y[c].x) * 3u + __float_as_uint(x[c].x));
                } else if (MIX == 8) {   // scalar FMUL + FADD with packed loads' worth: FMUL2 + two scalar FADD
                    f2 p = __fmul2_rn(A, x[(c + 1) % CHAINS]);
                    x[c].x = __fadd_rn(x[c].x, p.x); x[c].y = __fadd_rn(x[c].y, p.y);
                } else if (MIX == 9) {   // two scalar FMUL + FFMA2-as-add
                    f2 p = make_float2(__fmul_rn(a, x[(c + 1) % CHAINS].x), __fmul_rn(a, x[(c + 1) % CHAINS].y));
                    x[c] = __ffma2_rn(x[c], ONE, p);
                } else if (MIX == 10) {  // FADD2 alone
                    x[c] = __fadd2_rn(x[c], B);
                } else if (MIX == 11) {  // FMUL2 alone
                    x[c] = __fmul2_rn(x[c], A);
                }
            }
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c].x + x[c].y + y[c].x + y[c].y;
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

struct Mix { const char* name; int instr_per_chain; int laneops_per_chain; };

template <int MIX>
void run(const Mix& m, float* sink, unsigned long long* cyc, int sms) {
    const int blocks = sms * 8, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    unsigned long long cycles = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_mix<MIX><<<blocks, 256>>>(sink, iters, 0.999f, 0.001f, 1.0f, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) { best = ms; cudaMemcpy(&cycles, cyc, 8, cudaMemcpyDeviceToHost); }
    }
    // per CTA-resident warp: the kernel runs 8 CTAs x 8 warps per SM = 16 warps per SMSP, in one wave
    const double winstr_per_warp = (double)iters * 4 * CHAINS * m.instr_per_chain;
    const double per_smsp = winstr_per_warp * 16;          // warp instructions one SMSP issues
    const double clk = best * 1e-3 * 1.965e9;               // nominal clocks of the whole launch
    printf("%-52s %8.3f ms  %6.3f warp-instr/clk/SMSP  %7.1f lane-ops/clk/SM  (one CTA's clock64 span: %.3f ms at 1.965 GHz)\n", m.name, best,
           per_smsp / clk, (double)iters * 4 * CHAINS * m.laneops_per_chain * 32 * 64 / clk, cycles / 1.965e6);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* sink;
    unsigned long long* cyc;
    cudaMalloc(&sink, 256);
    cudaMalloc(&cyc, 8);
    // lane-ops: one per scalar result (an FMA counts once)
    run<0>({"scalar FFMA x2", 2, 2}, sink, cyc, sms);
    run<1>({"FFMA2", 1, 2}, sink, cyc, sms);
    run<10>({"FADD2", 1, 2}, sink, cyc, sms);
    run<11>({"FMUL2", 1, 2}, sink, cyc, sms);
    run<2>({"blur step packed: FMUL2 + FFMA2(s,1,p)", 2, 4}, sink, cyc, sms);
    run<3>({"blur step scalar: 2 FMUL + 2 FADD", 4, 4}, sink, cyc, sms);
    run<4>({"blur step half packed, half scalar", 3, 4}, sink, cyc, sms);
    run<8>({"blur step: FMUL2 + 2 scalar FADD", 3, 4}, sink, cyc, sms);
    run<9>({"blur step: 2 scalar FMUL + FFMA2(s,1,p)", 3, 4}, sink, cyc, sms);
    run<5>({"FFMA2 + scalar FFMA 1:1", 2, 3}, sink, cyc, sms);
    run<6>({"FFMA2 + SHF (ALU pipe) 1:1", 2, 3}, sink, cyc, sms);
    run<7>({"FFMA2 + IMAD 1:1", 2, 3}, sink, cyc, sms);
    return 0;
}
