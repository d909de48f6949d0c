// Inner-loop probe for the network kernel (K2): measures the FP32 FMA rate the register-tiled
// conv loop reaches on one B200 as a function of (channels per warp, taps, warps per SM,
// scalar FFMA vs packed fma.rn.f32x2).  Not part of the product; build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o ffma_probe ffma_probe.cu
//   ./ffma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kRS = 136;

__device__ __forceinline__ void fma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

// One "layer": acc[4][COUT] += x[f+k] * w[ci][k][c], weights and activations in shared memory.
template <int COUT, int KW, int CIN, bool PACKED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(float* out, int iters) {
    extern __shared__ __align__(16) float smem[];
    constexpr int COUTP = (COUT + 3) & ~3;
    constexpr int NX4 = (KW + 3 + 3) / 4;
    float* sW = smem;                                 // [CIN][KW][COUTP]
    float* sX = smem + CIN * KW * COUTP;              // per warp: [CIN][kRS] (+16)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < CIN * KW * COUTP; i += blockDim.x) sW[i] = 1e-3f * (float)((i * 37) % 19 - 9);
    for (int i = threadIdx.x; i < nw * (CIN * kRS + 16); i += blockDim.x) sX[i] = 1e-2f * (float)((i * 13) % 23 - 11);
    __syncthreads();
    const float* inx = sX + warp * (CIN * kRS + 16) + 4 * lane;

    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        if constexpr (!PACKED) {
            float acc[4][COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c)
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[f][c] = 0.f;
#pragma unroll 1
            for (int ci = 0; ci < CIN; ++ci) {
                float x[4 * NX4];
                const float4* xp = reinterpret_cast<const float4*>(inx + ci * kRS);
#pragma unroll
                for (int i = 0; i < NX4; ++i) {
                    const float4 v = xp[i];
                    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
                }
                const float4* wp = reinterpret_cast<const float4*>(sW + ci * (KW * COUTP));
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    float w[COUTP];
#pragma unroll
                    for (int j = 0; j < COUTP / 4; ++j) {
                        const float4 v = wp[k * (COUTP / 4) + j];
                        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
                    }
#pragma unroll
                    for (int c = 0; c < COUT; ++c)
#pragma unroll
                        for (int f = 0; f < 4; ++f) acc[f][c] = fmaf(x[f + k], w[c], acc[f][c]);
                }
            }
#pragma unroll
            for (int c = 0; c < COUT; ++c)
#pragma unroll
                for (int f = 0; f < 4; ++f) tot += acc[f][c];
        } else {
            constexpr int CP = (COUT + 1) / 2;   // channel pairs
            float2 acc[4][CP];
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[f][c] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int ci = 0; ci < CIN; ++ci) {
                float2 x[4 * NX4];
                const float4* xp = reinterpret_cast<const float4*>(inx + ci * kRS);
#pragma unroll
                for (int i = 0; i < NX4; ++i) {
                    const float4 v = xp[i];
                    x[4 * i] = make_float2(v.x, v.x); x[4 * i + 1] = make_float2(v.y, v.y);
                    x[4 * i + 2] = make_float2(v.z, v.z); x[4 * i + 3] = make_float2(v.w, v.w);
                }
                const float4* wp = reinterpret_cast<const float4*>(sW + ci * (KW * COUTP));
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    float2 w[COUTP / 2];
#pragma unroll
                    for (int j = 0; j < COUTP / 4; ++j) {
                        const float4 v = wp[k * (COUTP / 4) + j];
                        w[2 * j] = make_float2(v.x, v.y); w[2 * j + 1] = make_float2(v.z, v.w);
                    }
#pragma unroll
                    for (int c = 0; c < CP; ++c)
#pragma unroll
                        for (int f = 0; f < 4; ++f) fma2(acc[f][c], x[f + k], w[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
                for (int f = 0; f < 4; ++f) tot += acc[f][c].x + acc[f][c].y;
        }
    }
    if (tot == 123.456f) out[0] = tot;
}


// Register-only issue-rate probes: NACC independent accumulator chains per thread.
template <bool PACKED, int NACC, int WSH = 2, int XMASK = 3>
__global__ void __launch_bounds__(1024) peak_probe(float* out, int iters, float a, float b) {
    float tot = 0.f;
    if constexpr (PACKED) {
        float2 acc[NACC];
        float2 w[4] = {make_float2(a, b), make_float2(b, a), make_float2(a + 1.f, b), make_float2(a, b + 1.f)};
        float x[4] = {a, b, a + b, a - b};
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = make_float2((float)(threadIdx.x + i), 1.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < NACC; ++i) fma2(acc[i], make_float2(x[(i + r) & XMASK], x[(i + r) & XMASK]), w[(i >> WSH) & 3]);
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i) tot += acc[i].x + acc[i].y;
    } else {
        float acc[NACC];
        float w[4] = {a, b, a + 1.f, b + 1.f};
        float x[4] = {a, b, a + b, a - b};
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = (float)(threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x[(i + r) & 3], w[(i >> 2) & 3], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i) tot += acc[i];
    }
    if (tot == 123.456f) out[0] = tot;
}

template <bool PACKED, int NACC, int WSH = 2, int XMASK = 3>
static void run_peak(int warps, int ctas_per_sm, int iters, int sms) {
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        peak_probe<PACKED, NACC, WSH, XMASK><<<sms * ctas_per_sm, warps * 32>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    const double flop = 2.0 * sms * ctas_per_sm * warps * 32.0 * iters * 8.0 * NACC * (PACKED ? 2 : 1);
    printf("register-only wsh %d xmask %d %s nacc %2d warps/SM %2d : %7.3f ms  %6.2f TFLOP/s (%s)\n", WSH, XMASK, PACKED ? "f32x2 " : "scalar", NACC,
           warps * ctas_per_sm, best, flop / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

// Mapping B: one warp = one frame; lanes 0-15 / 16-31 take the two channel halves, every lane owns
// 8 bins, so each weight pair feeds 8 packed FMAs.  COUT = channels per HALF.
template <int COUT, int KW, int CIN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe8(float* out, int iters) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NP = COUT / 2;
    constexpr int CIB = (KW * COUT + 3) & ~3;
    constexpr int NX4 = (8 + KW - 1 + 3 + 3) / 4;     // window of 8 + KW - 1 floats starting anywhere in a 16-byte word
    float* sW = smem;                                 // [2][CIN][CIB]
    float* sX = smem + 2 * CIN * CIB;                 // per warp: [CIN + 1][kRS] (+32)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * CIN * CIB; i += blockDim.x) sW[i] = 1e-3f * (float)((i * 37) % 19 - 9);
    for (int i = threadIdx.x; i < nw * ((CIN + 1) * kRS + 32); i += blockDim.x) sX[i] = 1e-2f * (float)((i * 13) % 23 - 11);
    __syncthreads();
    const int g = lane >> 4, l = lane & 15;
    const float* inx = sX + warp * ((CIN + 1) * kRS + 32) + 8 * l;
    const float* W = sW + g * CIN * CIB;

    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        float2 acc[8][NP];
#pragma unroll
        for (int c = 0; c < NP; ++c)
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f][c] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int ci = 0; ci < CIN; ++ci) {
            float x[4 * NX4];
            const float4* xp = reinterpret_cast<const float4*>(inx + ci * kRS);
#pragma unroll
            for (int i = 0; i < NX4; ++i) {
                const float4 v = xp[i];
                x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
            }
            const float4* wp = reinterpret_cast<const float4*>(W + ci * CIB);
#pragma unroll
            for (int k = 0; k < KW; ++k) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int P = k * NP + j;
                    const float4 q = wp[P >> 1];
                    const float2 w = (P & 1) ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
#pragma unroll
                    for (int f = 0; f < 8; ++f) fma2(acc[f][j], make_float2(x[f + k], x[f + k]), w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NP; ++c)
#pragma unroll
            for (int f = 0; f < 8; ++f) tot += acc[f][c].x + acc[f][c].y;
    }
    if (tot == 123.456f) out[0] = tot;
}

template <int COUT, int KW, int CIN, int WARPS>
static void run8(int iters, int sms) {
    constexpr int CIB = (KW * COUT + 3) & ~3;
    const size_t smem = (size_t)(2 * CIN * CIB + WARPS * ((CIN + 1) * kRS + 32)) * 4;
    cudaFuncSetAttribute(probe8<COUT, KW, CIN, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        probe8<COUT, KW, CIN, WARPS><<<sms, WARPS * 32, smem>>>(d, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double flop = 2.0 * sms * WARPS * 32.0 * iters * (double)CIN * KW * COUT * 8;
    printf("mapping B: %2d ch/half kw %2d cin %2d warps/SM %2d : %7.3f ms  %6.2f TFLOP/s  (%s, smem %zu)\n", COUT, KW, CIN, WARPS,
           best, flop / (best * 1e-3) / 1e12, cudaGetErrorString(e), smem);
    cudaFree(d);
}

template <int COUT, int KW, int CIN, bool PACKED, int WARPS>
static void run(int iters, int sms) {
    const int warps = WARPS;
    constexpr int COUTP = (COUT + 3) & ~3;
    const size_t smem = (size_t)(CIN * KW * COUTP + warps * (CIN * kRS + 16)) * 4;
    cudaFuncSetAttribute(probe<COUT, KW, CIN, PACKED, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        probe<COUT, KW, CIN, PACKED, WARPS><<<sms, warps * 32, smem>>>(d, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    // useful FMAs: executed packed lanes for padded channel pairs are not counted
    const double flop = 2.0 * sms * warps * 32.0 * iters * (double)CIN * KW * COUT * 4;
    printf("cout %2d kw %2d cin %2d %s warps/SM %2d : %7.3f ms  %6.2f TFLOP/s useful  (%s)\n", COUT, KW, CIN,
           PACKED ? "f32x2 " : "scalar", warps, best, flop / (best * 1e-3) / 1e12, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    const int it = 200;
    run8<14, 11, 23, 4>(it, sms);
    run8<14, 11, 23, 8>(it, sms);
    run8<12, 7, 21, 4>(it, sms);
    run8<12, 7, 21, 8>(it, sms);
    run8<8, 5, 14, 4>(it, sms);
    run8<8, 5, 14, 8>(it, sms);
    run8<6, 7, 10, 8>(it, sms);
    run8<10, 5, 15, 8>(it, sms);
    run<14, 11, 23, true, 8>(it, sms);
    run<8, 5, 14, true, 8>(it, sms);
    return 0;
    for (int w : {4, 8, 16, 32}) {
        run_peak<false, 32>(w, 1, 4096, sms);
        run_peak<true, 32>(w, 1, 4096, sms);
    }
    run_peak<false, 16>(8, 8, 4096, sms);
    run_peak<true, 16>(8, 8, 4096, sms);
    run_peak<true, 16, 3, 3>(8, 8, 4096, sms);
    run_peak<true, 16, 4, 3>(8, 8, 4096, sms);
    run_peak<true, 16, 0, 0>(8, 8, 4096, sms);
    run_peak<true, 16, 4, 0>(8, 8, 4096, sms);
    run_peak<true, 16, 1, 3>(8, 8, 4096, sms);
    return 0;
    // today's structure: one warp per sub-partition, all channels of the layer
    run<23, 7, 21, false, 4>(it, sms);
    run<23, 7, 21, true, 4>(it, sms);
    run<24, 7, 21, false, 4>(it, sms);
    run<24, 7, 21, true, 4>(it, sms);
    run<12, 7, 21, false, 4>(it, sms);
    run<12, 7, 21, true, 4>(it, sms);
    // channel halves, two / four warps per sub-partition
    run<12, 7, 21, false, 8>(it, sms);
    run<12, 7, 21, true, 8>(it, sms);
    run<12, 7, 21, false, 16>(it, sms);
    run<12, 7, 21, true, 16>(it, sms);
    run<6, 7, 21, false, 16>(it, sms);
    run<6, 7, 21, true, 16>(it, sms);
    run<24, 7, 21, false, 8>(it, sms);
    run<24, 7, 21, true, 8>(it, sms);
    // small and large taps
    run<8, 5, 14, false, 8>(it, sms);
    run<8, 5, 14, true, 8>(it, sms);
    run<14, 11, 23, false, 8>(it, sms);
    run<14, 11, 23, true, 8>(it, sms);
    run<16, 5, 14, true, 8>(it, sms);
    run<16, 5, 14, false, 8>(it, sms);
    return 0;
}
