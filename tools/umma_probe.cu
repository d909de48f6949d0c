// tcgen05 probe for the network kernel's tensor-core variant ("Tensor cores").
//
// Question 1 (check): can a conv1d along frequency run as an implicit GEMM WITHOUT an im2col copy,
//   by giving every filter tap its own shared-memory matrix descriptor whose start address is the
//   activation array shifted by one row (16 bytes)?  Layout under test, K-major, no swizzle:
//   activations  A[cg][row][q]   (q = 4 tf32 or 8 f16 channels = one 16-byte chunk; cg = chunk index)
//   weights      B[tap][cg][n][q]
//   SBO (8-row group stride) = 128 B makes rows linear at 16 B, LBO = stride between the two K chunks.
//   D[r][n] = sum_tap sum_k A[r + tap][k] * B[tap][n][k] is compared with a CPU loop.
// Question 2 (rate): cycles per tcgen05.mma (SS mode, M = 128) against N for kind::tf32 (K = 8) and
//   kind::f16 (K = 16) when A is re-read from shared memory for every instruction, alone and
//   with the other warps streaming 16-byte shared-memory stores (the epilogue's traffic).
// Question 3 (ld): tcgen05.ld throughput for the epilogue (32x32b.x32, four warps).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o umma_probe umma_probe.cu
//   ./umma_probe check | rate | ld
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded wait: returns false on timeout instead of hanging the box
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>   // 0: f16, 2: tf32
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
            "l"(a), "l"(b), "r"(idesc), "r"(acc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
            "l"(a), "l"(b), "r"(idesc), "r"(acc)
            : "memory");
}
// K-major, no swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
// layout: 0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B; base_off: bits 49..51
__host__ __device__ inline uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0,
                                              uint32_t base_off = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: F32 accumulate, A/B format fmt (0 f16, 1 bf16, 2 tf32), both K-major
__host__ __device__ inline uint32_t instr_desc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// check: kTaps-tap conv, K = 2 chunks per instruction, M = 128, N = 32, in three layouts
//   layout 0 (none)   : A[cg][row][16 B], B[tap][cg][n][16 B]; SBO 128, LBO = chunk-plane stride
//   layout 2 (SW128)  : A[row][128 B], B[tap][n][128 B], 16-byte chunk c of a row stored at chunk
//                       c ^ (row & 7) (row = absolute 128-byte line of the 1024-byte aligned array);
//                       a tap shift moves the start address by one 128-byte line
// ------------------------------------------------------------------------------------------------
constexpr int kRows = 160;   // activation rows held (tile reads rows tap .. tap + 127)
constexpr int kN = 32;
constexpr int kTaps = 5;

struct CheckArgs {
    const uint32_t* a;   // byte image of A (a_bytes)
    const uint32_t* b;   // byte image of B (b_bytes)
    float* d;            // [128][kN]
    int a_bytes, b_bytes;
    int a_tap, a_lbo, a_sbo;   // bytes added to the A start address per tap; descriptor fields
    int b_tap, b_lbo, b_sbo;
    int layout;          // descriptor layout type
    int use_base_off;    // 1: base_offset field = (start >> 7) & 7
    int* status;
};

template <int KIND>
__global__ void __launch_bounds__(128, 1) check_kernel(CheckArgs p) {
    extern __shared__ uint32_t sm_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t s_tmem;
    const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
    uint32_t* sA = sm_raw + (base - smem_u32(sm_raw)) / 4;
    uint32_t* sB = sA + ((p.a_bytes + 1023) & ~1023) / 4;
    for (int i = threadIdx.x; i < p.a_bytes / 4; i += blockDim.x) sA[i] = p.a[i];
    for (int i = threadIdx.x; i < p.b_bytes / 4; i += blockDim.x) sB[i] = p.b[i];
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 32);
    fence_async_smem();   // generic-proxy writes of sA / sB -> visible to the tensor core's async proxy
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(KIND, 128, kN);
        for (int t = 0; t < kTaps; ++t) {
            const uint32_t a_addr = smem_u32(sA) + t * p.a_tap;
            const uint32_t b_addr = smem_u32(sB) + t * p.b_tap;
            const uint64_t da = smem_desc(a_addr, p.a_lbo, p.a_sbo, p.layout, p.use_base_off ? (a_addr >> 7) : 0);
            const uint64_t db = smem_desc(b_addr, p.b_lbo, p.b_sbo, p.layout, p.use_base_off ? (b_addr >> 7) : 0);
            umma<KIND>(tm, da, db, idesc, t > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    const bool ok = mbar_wait(smem_u32(&bar), 0);
    fence_after();
    if (!ok) {
        if (threadIdx.x == 0) *p.status = 1;
    } else {
        uint32_t v[32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) p.d[(warp * 32 + lane) * kN + j] = __uint_as_float(v[j]);
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

template <int KIND>
static int run_check(int layout, int use_base_off) {
    constexpr int Q = KIND == 2 ? 4 : 8;   // elements per 16-byte chunk
    constexpr int K = 2 * Q;
    constexpr int ES = KIND == 2 ? 4 : 2;
    std::vector<float> X((size_t)kRows * K), W((size_t)kTaps * kN * K);
    srand(7);
    for (auto& v : X) v = (float)(rand() % 17 - 8) * 0.125f;
    for (auto& v : W) v = (float)(rand() % 13 - 6) * 0.25f;
    CheckArgs args{};
    if (layout == 0) {
        args.a_bytes = 2 * kRows * 16; args.b_bytes = kTaps * 2 * kN * 16;
        args.a_tap = 16; args.a_lbo = kRows * 16; args.a_sbo = 128;
        args.b_tap = 2 * kN * 16; args.b_lbo = kN * 16; args.b_sbo = 128;
    } else {
        args.a_bytes = kRows * 128; args.b_bytes = kTaps * kN * 128;
        args.a_tap = 128; args.a_lbo = 16; args.a_sbo = 1024;
        args.b_tap = kN * 128; args.b_lbo = 16; args.b_sbo = 1024;
    }
    args.layout = layout; args.use_base_off = use_base_off;
    std::vector<uint8_t> ha(args.a_bytes, 0), hb(args.b_bytes, 0);
    auto put = [&](uint8_t* dst, float v) {
        if (KIND == 2) {
            memcpy(dst, &v, 4);
        } else {
            __half h = __float2half(v);
            memcpy(dst, &h, 2);
        }
    };
    for (int r = 0; r < kRows; ++r)
        for (int k = 0; k < K; ++k) {
            const int c = k / Q, q = k % Q;
            const size_t off = layout == 0 ? ((size_t)(c * kRows + r) * 16 + q * ES) : ((size_t)r * 128 + ((c ^ (r & 7)) * 16) + q * ES);
            put(&ha[off], X[(size_t)r * K + k]);
        }
    for (int t = 0; t < kTaps; ++t)
        for (int n = 0; n < kN; ++n)
            for (int k = 0; k < K; ++k) {
                const int c = k / Q, q = k % Q;
                const int line = t * kN + n;   // kN * 128 = 4096: every tap's tile starts 1024-aligned
                const size_t off = layout == 0 ? ((size_t)((t * 2 + c) * kN + n) * 16 + q * ES)
                                               : ((size_t)line * 128 + ((c ^ (line & 7)) * 16) + q * ES);
                put(&hb[off], W[((size_t)t * kN + n) * K + k]);
            }
    uint32_t *da, *db;
    float* dd;
    int* ds;
    CK(cudaMalloc(&da, ha.size()));
    CK(cudaMalloc(&db, hb.size()));
    CK(cudaMalloc(&dd, 128 * kN * 4));
    CK(cudaMalloc(&ds, 4));
    CK(cudaMemset(ds, 0, 4));
    CK(cudaMemset(dd, 0, 128 * kN * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    args.a = da; args.b = db; args.d = dd; args.status = ds;
    const size_t smem = ((args.a_bytes + 1023) & ~1023) + args.b_bytes + 2048;
    CK(cudaFuncSetAttribute(check_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    check_kernel<KIND><<<1, 128, smem>>>(args);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("check kind %d layout %d: kernel failed: %s\n", KIND, layout, cudaGetErrorString(e));
        return 2;
    }
    std::vector<float> D(128 * kN);
    int st;
    CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < kN; ++n) {
            double ref = 0;
            for (int t = 0; t < kTaps; ++t)
                for (int k = 0; k < K; ++k) ref += (double)tf32_trunc(X[(size_t)(r + t) * K + k]) * tf32_trunc(W[((size_t)t * kN + n) * K + k]);
            const double err = fabs(ref - D[r * kN + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) ++bad;
        }
    printf("check kind %s layout %d base_off %d: timeout %d, max |err| %.3g, mismatches %d / %d  -> %s\n",
           KIND == 2 ? "tf32" : "f16", layout, use_base_off, st, maxerr, bad, 128 * kN, (st == 0 && bad == 0) ? "OK" : "WRONG");
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(ds);
    return (st == 0 && bad == 0) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// rate: NI instructions back to back, A descriptor cycling over 8 tap shifts, B over 8 tiles
// ------------------------------------------------------------------------------------------------
struct RateArgs {
    int n;          // N of the instruction
    int m;          // 64 or 128
    int ni;         // instructions per measurement
    int store_warps;   // warps 1.. that stream 16-byte shared stores while the MMAs run
    int layout;     // 0: no swizzle (chunk planes), 2: 128-byte swizzle (row = 128 B)
    int same_a;     // 1: every instruction reads the same A tile (no tap / tile cycling)
    int commit_every;   // > 0: tcgen05.commit (to a barrier nobody waits on) after every so many MMAs
    int fence_every;    // > 0: tcgen05.fence::after_thread_sync after every so many MMAs
    long long* cycles;   // per CTA
    int* status;
};

constexpr int kRateABytes = 1152 * 128;      // 144 KB: 1152 rows of 128 B, or 8 chunk planes of 1152 x 16 B
constexpr int kRateBBytes = 8 * 256 * 32;    // 64 KB: 8 tap tiles of 256 rows x 2 chunks

template <int KIND>
__global__ void __launch_bounds__(256, 1) rate_kernel(RateArgs p) {
    extern __shared__ uint32_t sm_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ __align__(8) unsigned long long bar2;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
    uint32_t* sA = sm_raw + (base - smem_u32(sm_raw)) / 4;
    uint32_t* sB = sA + kRateABytes / 4;
    if (threadIdx.x == 0) mbar_init(smem_u32(&bar2), 1);
    uint32_t* sS = sB + kRateBBytes / 4;   // 16 KB scratch for the store warps
    for (int i = threadIdx.x; i < (kRateABytes + kRateBBytes + 16384) / 4; i += blockDim.x) sA[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_stop = 0;
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t idesc = instr_desc(KIND, p.m, p.n);
        // descriptors of 16 instructions (8 taps x 2 accumulator tiles) built before the timed loop,
        // so that the issuing thread does nothing but tcgen05.mma
        uint64_t da[16], db[16];
        uint32_t dd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int tap = p.same_a ? 0 : (i & 7);
            const int tile = p.same_a ? 0 : (i >> 3);
            if (p.layout == 0) {
                da[i] = smem_desc(smem_u32(sA) + (tile * 128 + tap) * 16, 1152 * 16, 128);
                db[i] = smem_desc(smem_u32(sB) + tap * (2 * 256 * 16), 256 * 16, 128);
            } else {
                da[i] = smem_desc(smem_u32(sA) + (tile * 128 + tap) * 128, 16, 1024, 2);
                db[i] = smem_desc(smem_u32(sB), 16, 1024, 2);   // 256 rows x 128 B = 32 KB tile
            }
            dd[i] = tm + (uint32_t)(tile * 256);
        }
        const long long t0 = clock64();
        if (elect_one()) {
            int since_c = 0, since_f = 0;
            for (int i = 0; i < p.ni; i += 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    umma<KIND>(dd[j], da[j], db[j], idesc, 1);
                    if (p.commit_every > 0 && ++since_c == p.commit_every) { umma_commit(smem_u32(&bar2)); since_c = 0; }
                    if (p.fence_every > 0 && ++since_f == p.fence_every) { fence_after(); since_f = 0; }
                }
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        const bool ok = mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) {
            s_stop = 1;
            if (!ok) *p.status = 1;
            p.cycles[blockIdx.x] = t1 - t0;
        }
    } else if (warp >= 1 && warp <= p.store_warps) {
        // epilogue-like traffic: each lane stores 16 bytes, consecutive lanes consecutive addresses
        uint4* dst = reinterpret_cast<uint4*>(sS) + ((warp - 1) & 3) * 256 + (threadIdx.x & 31);
        uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
        while (!s_stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dst[j * 32] = v;
                v.x += 1;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int KIND>
static void run_rate(int grid) {
    const size_t smem = kRateABytes + kRateBBytes + 16384 + 1024;
    CK(cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* dc;
    int* ds;
    CK(cudaMalloc(&dc, grid * 8));
    CK(cudaMalloc(&ds, 4));
    CK(cudaMemset(ds, 0, 4));
    const int ns[] = {8, 16, 32, 64, 128, 256};
    const int KE = KIND == 2 ? 8 : 16;
    for (int layout = 0; layout <= 2; layout += 2)
        for (int m = 64; m <= 128; m += 64)
            for (int same = 0; same < 2; ++same)
                for (int sw = 0; sw <= 4; sw += 4)
                    for (int n : ns) {
                        if ((same || sw) && (n != 32 && n != 128)) continue;
                        RateArgs a{n, m, 4096, sw, layout, same, 0, 0, dc, ds};
                        rate_kernel<KIND><<<grid, 256, smem>>>(a);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) {
                            printf("rate kind %d N %d: kernel failed: %s\n", KIND, n, cudaGetErrorString(e));
                            return;
                        }
                        std::vector<long long> c(grid);
                        int st;
                        CK(cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
                        long long mx = 0;
                        for (auto v : c) mx = v > mx ? v : mx;
                        const double cyc = (double)mx / 4096;
                        printf("rate %s grid %3d layout %d M %3d same_a %d store_warps %d N %3d: %7.1f cycles/instr, %7.0f MAC/cycle/SM%s\n",
                               KIND == 2 ? "tf32" : "f16 ", grid, layout, m, same, sw, n, cyc, (double)m * KE * n / cyc,
                               st ? "  (TIMEOUT)" : "");
                    }
    cudaFree(dc);
    cudaFree(ds);
}

// ------------------------------------------------------------------------------------------------
// ld: tcgen05.ld 32x32b.x32, four warps, 64 rounds of 8 loads (all 256 of 512 columns)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ld_kernel(long long* cycles, float* sink, int batch) {
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    uint32_t v[32];
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < 64; ++r) {
        for (int c = 0; c < 8; ++c) {
            tmem_ld32(tm + c * 32, v);
            if (batch == 1 || (c % batch) == batch - 1) tmem_wait_ld();
            if (batch == 1) acc += __uint_as_float(v[r & 31]);
        }
        tmem_wait_ld();
        acc += __uint_as_float(v[r & 31]);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s_tmem, 512);
}

static void run_sync_sweep() {
    const size_t smem = kRateABytes + kRateBBytes + 16384 + 1024;
    CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* dc;
    int* ds;
    CK(cudaMalloc(&dc, 8));
    CK(cudaMalloc(&ds, 4));
    CK(cudaMemset(ds, 0, 4));
    for (int n : {32, 64})
        for (int ce : {0, 32, 16, 8, 4, 2})
            for (int fe : {0, 8}) {
                RateArgs a{n, 128, 4096, 0, 0, 0, ce, fe, dc, ds};
                rate_kernel<0><<<1, 256, smem>>>(a);
                CK(cudaDeviceSynchronize());
                long long c;
                CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
                printf("sync f16 N %3d commit every %2d MMAs, fence every %d: %6.1f cycles/MMA\n", n, ce, fe, (double)c / 4096);
            }
}

static void run_ld() {
    long long* dc;
    float* dsink;
    CK(cudaMalloc(&dc, 8));
    CK(cudaMalloc(&dsink, 128 * 4));
    for (int batch : {1, 2, 4, 8}) {
        ld_kernel<<<1, 128>>>(dc, dsink, batch);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        const double bytes = 64.0 * 8 * 128 * 32 * 4;
        printf("ld 32x32b.x32, wait every %d loads: %lld cycles, %.1f bytes/cycle/SM, %.1f cycles per 128x32 tile\n", batch, c,
               bytes / c, (double)c / (64 * 8));
    }
}

// ------------------------------------------------------------------------------------------------
// lat: cost of the synchronisation primitives the MMA-issuing thread executes between two tiles
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out) {
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int flag;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar[0]), 1);
        mbar_init(smem_u32(&bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        flag = 1;
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 32);
    fence_before();
    __syncthreads();
    fence_after();
    if (threadIdx.x == 0) {
        // complete phase 0 of bar[0] so that waits on parity 0 succeed at once
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        long long t0 = clock64();
        for (int i = 0; i < 64; ++i) mbar_wait(smem_u32(&bar[0]), 0);
        long long t1 = clock64();
        out[0] = (t1 - t0) / 64;   // try_wait on a completed phase
        t0 = clock64();
        for (int i = 0; i < 64; ++i) fence_after();
        t1 = clock64();
        out[1] = (t1 - t0) / 64;   // tcgen05.fence::after_thread_sync
        t0 = clock64();
        int acc = 0;
        for (int i = 0; i < 64; ++i) acc += flag;
        t1 = clock64();
        out[2] = (t1 - t0) / 64 + (acc == 12345);   // volatile shared load
        // commit with nothing outstanding -> wait for it
        t0 = clock64();
        for (int i = 0; i < 16; ++i) {
            umma_commit(smem_u32(&bar[1]));
            mbar_wait(smem_u32(&bar[1]), i & 1);
        }
        t1 = clock64();
        out[3] = (t1 - t0) / 16;   // commit + wait round trip
        t0 = clock64();
        for (int i = 0; i < 16; ++i) umma_commit(smem_u32(&bar[1]));
        t1 = clock64();
        out[4] = (t1 - t0) / 16;   // commit issue cost
        t0 = clock64();
        for (int i = 0; i < 64; ++i) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        t1 = clock64();
        out[5] = (t1 - t0) / 64;   // mbarrier.arrive
        t0 = clock64();
        for (int i = 0; i < 64; ++i) fence_async_smem();
        t1 = clock64();
        out[6] = (t1 - t0) / 64;   // fence.proxy.async.shared::cta
        t0 = clock64();
        int okc = 0;
        for (int i = 0; i < 64; ++i) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"(1u) : "memory");
            okc += ok;
            if (!ok) break;
        }
        t1 = clock64();
        out[7] = (t1 - t0) / 64 + (okc == 12345);   // mbarrier.test_wait (dependent chain)
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s_tmem, 32);
}

static void run_lat() {
    long long* d;
    CK(cudaMalloc(&d, 64));
    lat_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    long long h[8];
    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
    const char* names[8] = {"mbarrier.try_wait (phase already complete)", "tcgen05.fence::after_thread_sync", "ld.volatile.shared",
                            "tcgen05.commit + wait (nothing outstanding)", "tcgen05.commit (issue only)", "mbarrier.arrive",
                            "fence.proxy.async.shared::cta", "mbarrier.test_wait (phase already complete)"};
    for (int i = 0; i < 8; ++i) printf("lat %-48s %5lld cycles\n", names[i], h[i]);
}

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "check";
    if (!strcmp(mode, "check")) {
        const int kind = argc > 2 ? atoi(argv[2]) : 2;
        const int layout = argc > 3 ? atoi(argv[3]) : 0;
        const int bo = argc > 4 ? atoi(argv[4]) : 0;
        return kind == 2 ? run_check<2>(layout, bo) : run_check<0>(layout, bo);
    }
    if (!strcmp(mode, "rate")) {
        const int grid = argc > 2 ? atoi(argv[2]) : 1;
        run_rate<2>(grid);
        run_rate<0>(grid);
        return 0;
    }
    if (!strcmp(mode, "sync")) {
        run_sync_sweep();
        return 0;
    }
    if (!strcmp(mode, "lat")) {
        run_lat();
        return 0;
    }
    if (!strcmp(mode, "ld")) {
        run_ld();
        return 0;
    }
    printf("usage: umma_probe check [kind 0|2] [layout 0|2] [base_off 0|1] | rate [grid] | ld\n");
    return 1;
}
