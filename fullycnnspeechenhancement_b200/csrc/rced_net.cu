// K2: the fused R-CED / CR-CED network kernel (sm_100a).
//
// Replaces FullyCNNTester.test_step / sess.run(pred) of the reference
// (model_utils/tester.py:85-90) for the graphs of model_utils/model.py.
//
// Design:
//  * persistent grid, one CTA of 4 warps per SM, one warp per SM sub-partition;
//  * every warp is an independent FRAME PIPELINE: it takes one spectrogram frame through
//    all 10/16 layers; lane l owns frequency bins 4l..4l+3 for ALL output channels of the
//    current layer (register tile 4 x cout, FP32 FFMA), bin 128 is a small extra phase with
//    lane == output channel;
//  * all BN-folded weights (about 130 KB) sit in shared memory for the whole kernel, brought
//    in once per CTA with bulk async copies (cp.async.bulk + mbarrier);
//  * layer activations live in a per-warp shared-memory slot and are overwritten in place
//    (the whole layer output is held in registers before the first store), so no block-level
//    barrier exists anywhere in the layer loop -- only __syncwarp;
//  * encoder outputs needed later by decoder skip connections are parked in TENSOR MEMORY
//    (tcgen05.st), thread-private, and loaded back (tcgen05.ld) straight into the accumulator
//    registers of the consuming decoder layer.  Global memory only sees the input magnitude
//    rows and the output rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rced_arch.cuh"
#include "rced_internal.h"

namespace rced {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- tensor memory (tcgen05) ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc512(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a), "=f"(b), "=f"(c), "=f"(d)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float& a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(a) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Registers written by tcgen05.ld may only be read after tcgen05.wait::ld; routing them through
// an (empty) volatile asm placed after the wait pins that order for the compiler.
__device__ __forceinline__ void reg_fence4(float& a, float& b, float& c, float& d) {
    asm volatile("" : "+f"(a), "+f"(b), "+f"(c), "+f"(d)::"memory");
}
__device__ __forceinline__ void reg_fence1(float& a) { asm volatile("" : "+f"(a)::"memory"); }

// ------------------------------------------------------------------------------------------
// where skip tensors are parked: tensor memory (default) or a per-warp global scratch area
// ------------------------------------------------------------------------------------------
template <bool TM>
struct SkipStore;

template <>
struct SkipStore<true> {
    uint32_t base;   // tensor-memory address of column 0 in this warp's lane quadrant
    __device__ __forceinline__ void st4(int col, float a, float b, float c, float d) const { tmem_st4(base + col, a, b, c, d); }
    __device__ __forceinline__ void st1(int col, float a) const { tmem_st1(base + col, a); }
    __device__ __forceinline__ void ld4(int col, float& a, float& b, float& c, float& d) const { tmem_ld4(base + col, a, b, c, d); }
    __device__ __forceinline__ void ld1(int col, float& a) const { tmem_ld1(base + col, a); }
    __device__ __forceinline__ void wait_ld() const { tmem_wait_ld(); }
    __device__ __forceinline__ void wait_st() const { tmem_wait_st(); }
};

template <>
struct SkipStore<false> {
    float* base;   // scratch + (global warp id) * 512 * 32 + lane ; column stride 32 floats
    __device__ __forceinline__ void st4(int col, float a, float b, float c, float d) const {
        base[(col + 0) * 32] = a; base[(col + 1) * 32] = b; base[(col + 2) * 32] = c; base[(col + 3) * 32] = d;
    }
    __device__ __forceinline__ void st1(int col, float a) const { base[col * 32] = a; }
    __device__ __forceinline__ void ld4(int col, float& a, float& b, float& c, float& d) const {
        a = base[(col + 0) * 32]; b = base[(col + 1) * 32]; c = base[(col + 2) * 32]; d = base[(col + 3) * 32];
    }
    __device__ __forceinline__ void ld1(int col, float& a) const { a = base[col * 32]; }
    __device__ __forceinline__ void wait_ld() const {}
    __device__ __forceinline__ void wait_st() const {}
};

// ------------------------------------------------------------------------------------------
// one conv_bn_relu layer (all but the last) for one frame, executed by one warp
// ------------------------------------------------------------------------------------------
template <int ARCH, bool TM, int LI>
__device__ __forceinline__ void conv_layer(const float* __restrict__ sW, float* __restrict__ slot, const int lane,
                                           const SkipStore<TM>& sk) {
    constexpr LSpec S = spec(ARCH, LI);
    constexpr int NL = num_layers(ARCH);
    constexpr int CIN = cin_eff(ARCH, LI);
    constexpr int COUT = S.cout;
    constexpr int COUTP = pad4(COUT);
    constexpr int KW = S.kw;
    constexpr int PADL = (KW - 1) / 2;
    constexpr bool WIDEWIN = PADL > 4;                 // window of 20 floats instead of 12
    constexpr int NX4 = WIDEWIN ? 5 : 3;
    constexpr int XB = (WIDEWIN ? 8 : 4) - PADL;       // x[XB + f + k] is bin 4l+f+k-PADL
    constexpr bool OUT_WIDE = (LI == NL - 2);          // feeds the (1,129) layer
    constexpr bool PRE_ADD = (S.add >= 0) && !S.after; // skip pre-loaded into the accumulators
    static_assert(PADL <= 8, "window loader covers SAME pads up to 8");
    static_assert(COUT <= 32, "tail phase maps output channels to lanes");

    const float* __restrict__ W = sW + packed_w_off(ARCH, LI);
    const float* __restrict__ B = sW + packed_b_off(ARCH, LI);
    const float* __restrict__ in0 = slot + (LI == 0 ? stage_row(ARCH) * kRS : 0);
    const float* __restrict__ inx = in0 + (WIDEWIN ? 0 : 4) + 4 * lane;
    const int cl = lane < COUTP ? lane : COUTP - 1;    // tail phase: this lane's output channel

    float acc[4][COUT];
    float tacc;

    if constexpr (PRE_ADD) {
        constexpr int col = skip_col_base(ARCH, S.add);
#pragma unroll
        for (int c = 0; c < COUT; ++c) sk.ld4(col + 4 * c, acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        sk.ld1(col + 4 * COUT, tacc);
        sk.wait_ld();
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            reg_fence4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
            const float b = B[c];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f][c] += b;
        }
        reg_fence1(tacc);
        tacc += B[cl];
    } else {
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            const float b = B[c];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f][c] = b;
        }
        tacc = B[cl];
    }

    // ---- main phase: bins 4l..4l+3, all output channels ----------------------------------
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
        float x[NX4 * 4];
        const float4* xp = reinterpret_cast<const float4*>(inx + ci * kRS);
#pragma unroll
        for (int i = 0; i < NX4; ++i) {
            const float4 v = xp[i];
            x[4 * i + 0] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        const float4* wp = reinterpret_cast<const float4*>(W + ci * (KW * COUTP));
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            float w[COUTP];
#pragma unroll
            for (int j = 0; j < COUTP / 4; ++j) {
                const float4 v = wp[k * (COUTP / 4) + j];   // warp-uniform address: broadcast
                w[4 * j + 0] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[f][c] = fmaf(x[XB + f + k], w[c], acc[f][c]);
            }
        }
    }

    // ---- tail phase: bin 128, lane == output channel ---------------------------------------
    {
        const float* xt = in0 + kRowBin0 + 128 - PADL;   // bins 128-PADL .. 128 (taps beyond hit zeros)
        const float* wt = W + cl;
#pragma unroll 2
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int k = 0; k <= PADL; ++k)
                tacc = fmaf(xt[ci * kRS + k], wt[(ci * KW + k) * COUTP], tacc);
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------
    float tsk = 0.f;
    if constexpr (S.add >= 0 && S.after) {   // V3: relu first, then add the skip (no second relu)
        constexpr int col = skip_col_base(ARCH, S.add);
        sk.ld1(col + 4 * COUT, tsk);
    }
    __syncwarp();   // every lane has finished reading this layer's input (it is overwritten below)

    if constexpr (OUT_WIDE) {
        // the (1,129) layer reads rows of stride kWS with 64 zeros either side: clear, then fill
        float4* z = reinterpret_cast<float4*>(slot);
        for (int i = lane; i < wide_floats(ARCH) / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
    }

#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        float v0 = acc[0][c], v1 = acc[1][c], v2 = acc[2][c], v3 = acc[3][c];
        if constexpr (S.relu) {
            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
        }
        if constexpr (S.add >= 0 && S.after) {
            constexpr int col = skip_col_base(ARCH, S.add);
            float s0, s1, s2, s3;
            sk.ld4(col + 4 * c, s0, s1, s2, s3);
            sk.wait_ld();
            reg_fence4(s0, s1, s2, s3);
            v0 += s0; v1 += s1; v2 += s2; v3 += s3;
        }
        if constexpr (S.save >= 0) sk.st4(skip_col_base(ARCH, S.save) + 4 * c, v0, v1, v2, v3);
        float* o = OUT_WIDE ? slot + c * kWS + kWideBin0 + 4 * lane : slot + c * kRS + kRowBin0 + 4 * lane;
        *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
    }
    {
        float t = tacc;
        if constexpr (S.relu) t = fmaxf(t, 0.f);
        if constexpr (S.add >= 0 && S.after) {
            sk.wait_ld();
            reg_fence1(tsk);
            t += tsk;
        }
        if constexpr (S.save >= 0) sk.st1(skip_col_base(ARCH, S.save) + 4 * COUT, t);
        if (lane < COUT) {
            float* o = OUT_WIDE ? slot + lane * kWS + kWideBin0 + 128 : slot + lane * kRS + kRowBin0 + 128;
            *o = t;
        }
    }
    if constexpr (S.save >= 0) sk.wait_st();
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// the (1,129) output layer: cout = 1, no BN / ReLU; reads the wide layout, writes global memory
// ------------------------------------------------------------------------------------------
template <int ARCH>
__device__ __forceinline__ void final_layer(const float* __restrict__ sW, const float* __restrict__ slot, const int lane,
                                            float* __restrict__ out_row) {
    constexpr int LI = num_layers(ARCH) - 1;
    constexpr int CIN = spec(ARCH, LI).cin;
    const float* __restrict__ Wf = sW + packed_w_off(ARCH, LI);
    const float bias = sW[packed_b_off(ARCH, LI)];

    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
        // out bin 4l+j, tap k reads wide offset 4l + j + k
        const float4* row = reinterpret_cast<const float4*>(slot + ci * kWS + 4 * lane);
        const float4* w4 = reinterpret_cast<const float4*>(Wf + ci * kFinalKP);
        float4 xa = row[0];
#pragma unroll 8
        for (int q = 0; q < 32; ++q) {
            const float4 xb = row[q + 1];
            const float4 w = w4[q];
            a0 = fmaf(xa.x, w.x, a0); a0 = fmaf(xa.y, w.y, a0); a0 = fmaf(xa.z, w.z, a0); a0 = fmaf(xa.w, w.w, a0);
            a1 = fmaf(xa.y, w.x, a1); a1 = fmaf(xa.z, w.y, a1); a1 = fmaf(xa.w, w.z, a1); a1 = fmaf(xb.x, w.w, a1);
            a2 = fmaf(xa.z, w.x, a2); a2 = fmaf(xa.w, w.y, a2); a2 = fmaf(xb.x, w.z, a2); a2 = fmaf(xb.y, w.w, a2);
            a3 = fmaf(xa.w, w.x, a3); a3 = fmaf(xb.x, w.y, a3); a3 = fmaf(xb.y, w.z, a3); a3 = fmaf(xb.z, w.w, a3);
            xa = xb;
        }
        const float w128 = Wf[ci * kFinalKP + 128];
        a0 = fmaf(xa.x, w128, a0); a1 = fmaf(xa.y, w128, a1); a2 = fmaf(xa.z, w128, a2); a3 = fmaf(xa.w, w128, a3);
    }
    // bin 128: taps 0..64 over wide offsets 128..192, split across lanes, shuffle-reduced
    float t = 0.f;
#pragma unroll 2
    for (int ci = 0; ci < CIN; ++ci) {
        const float* r = slot + ci * kWS + 128;
        const float* w = Wf + ci * kFinalKP;
        t = fmaf(r[lane], w[lane], t);
        t = fmaf(r[lane + 32], w[lane + 32], t);
        if (lane == 0) t = fmaf(r[64], w[64], t);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);

    float* o = out_row + 4 * lane;
    o[0] = a0 + bias; o[1] = a1 + bias; o[2] = a2 + bias; o[3] = a3 + bias;
    if (lane == 0) out_row[128] = t + bias;
}

template <int ARCH, bool TM, int LI>
__device__ __forceinline__ void run_conv_layers(const float* sW, float* slot, int lane, const SkipStore<TM>& sk) {
    if constexpr (LI < num_layers(ARCH) - 1) {
        conv_layer<ARCH, TM, LI>(sW, slot, lane, sk);
        run_conv_layers<ARCH, TM, LI + 1>(sW, slot, lane, sk);
    }
}

// ------------------------------------------------------------------------------------------
// frame bookkeeping
// ------------------------------------------------------------------------------------------
struct FrameLoc {
    long long lo, hi;   // rows [lo, hi) of the utterance that owns the frame
};

__device__ __forceinline__ FrameLoc locate(const long long* __restrict__ row_off, int n_utt, long long g) {
    int a = 0, b = n_utt;   // invariant: row_off[a] <= g < row_off[b]
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (__ldg(row_off + m) <= g) a = m; else b = m;
    }
    FrameLoc f;
    f.lo = __ldg(row_off + a);
    f.hi = __ldg(row_off + a + 1);
    return f;
}

// Stage the 8 input rows (frames g-3 .. g+4, zeros outside the utterance) of frame g into rows
// stage_row .. stage_row+7 of the slot with asynchronous 4-byte copies.
template <int ARCH>
__device__ __forceinline__ void prefetch_frame(const NetParams& p, long long g, float* slot, int lane) {
    constexpr int SR = stage_row(ARCH);
    const FrameLoc loc = locate(p.row_off, p.n_utt, g);
    // halo offsets 1..7 of the 9 rows touched (offset 0 is bin 128 of the row before)
    for (int i = lane; i < 9 * 7; i += 32) slot[(SR + i / 7) * kRS + 1 + (i % 7)] = 0.f;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const long long r = g + dt - 3;
        float* dst = slot + (SR + dt) * kRS + kRowBin0;
        if (r >= loc.lo && r < loc.hi) {
            const float* src = p.in + r * (long long)kBins;
            const uint32_t d = smem_u32(dst + 4 * lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async4(d + 4 * j, src + 4 * lane + j);
            if (lane == 0) cp_async4(smem_u32(dst + 128), src + 128);
        } else {
            *reinterpret_cast<float4*>(dst + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane == 0) dst[128] = 0.f;
        }
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int ARCH, bool TM>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) rced_net_kernel(const NetParams p) {
    extern __shared__ __align__(128) float smem[];
    constexpr int PK = packed_count(ARCH);
    constexpr int SLOT = slot_floats(ARCH);
    constexpr int SR = stage_row(ARCH);
    float* sW = smem;
    float* slots = smem + pad4(PK);
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* slot = slots + warp * SLOT;

    // ---- weights -> shared memory with bulk async copies -----------------------------------
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        constexpr uint32_t total = pad4(PK) * 4u;
        mbar_expect_tx(bar, total);
        constexpr uint32_t CH = 16384;
        for (uint32_t o = 0; o < total; o += CH) {
            const uint32_t n = total - o < CH ? total - o : CH;
            bulk_g2s(smem_u32(sW) + o, reinterpret_cast<const char*>(p.packed) + o, n, bar);
        }
    }
    // meanwhile: clear this warp's slot (all halos must read as zero)
    for (int i = lane; i < SLOT / 4; i += 32) reinterpret_cast<float4*>(slot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    SkipStore<TM> sk;
    if constexpr (TM) {
        if (warp == 0) tmem_alloc512(smem_u32(&s_tmem));
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        sk.base = s_tmem + ((uint32_t)(warp & 3) << 21);   // lane field (bits 31:16) = 32 * (warp % 4)
    } else {
        sk.base = p.skip_scratch + ((size_t)blockIdx.x * kWarpsPerCta + warp) * (512 * 32) + lane;
        __syncthreads();
    }
    mbar_wait(bar, 0);

    const long long stride = (long long)gridDim.x * kWarpsPerCta;
    long long g = (long long)blockIdx.x * kWarpsPerCta + warp;
    __syncwarp();
    if (g < p.total_rows) prefetch_frame<ARCH>(p, g, slot, lane);

    for (; g < p.total_rows; g += stride) {
        cp_async_wait_all();
        __syncwarp();
        run_conv_layers<ARCH, TM, 0>(sW, slot, lane, sk);
        // the wide layout sits below row SR: the next frame's input can land while the last layer runs
        const long long gn = g + stride;
        if (gn < p.total_rows) prefetch_frame<ARCH>(p, gn, slot, lane);
        final_layer<ARCH>(sW, slot, lane, p.out + g * (long long)kBins);
        __syncwarp();
        // the wide layout overwrote the halos of rows 0..SR: restore their zeros
        for (int i = lane; i < (SR + 1) * 7; i += 32) slot[(i / 7) * kRS + 1 + (i % 7)] = 0.f;
        __syncwarp();
    }

    if constexpr (TM) {
        tmem_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc512(s_tmem);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int ARCH>
constexpr size_t net_smem_bytes() {
    return (size_t)(pad4(packed_count(ARCH)) + kWarpsPerCta * slot_floats(ARCH)) * sizeof(float);
}

template <int ARCH, bool TM>
static cudaError_t launch_net_t(const NetParams& p, int num_sms, cudaStream_t stream) {
    constexpr size_t smem = net_smem_bytes<ARCH>();
    static_assert(smem <= 227 * 1024, "weights + activation slots must fit one SM's shared memory");
    static bool configured = false;   // per template instance; the attribute is per-device-per-function
    (void)configured;
    cudaError_t e = cudaFuncSetAttribute(rced_net_kernel<ARCH, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long ctas = (p.total_rows + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas > num_sms) ctas = num_sms;
    if (ctas < 1) return cudaSuccess;
    rced_net_kernel<ARCH, TM><<<(unsigned)ctas, kWarpsPerCta * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_net(int arch, bool skip_in_tmem, const NetParams& p, int num_sms, cudaStream_t stream) {
    switch (arch * 2 + (skip_in_tmem ? 1 : 0)) {
        case 2: return launch_net_t<1, false>(p, num_sms, stream);
        case 3: return launch_net_t<1, true>(p, num_sms, stream);
        case 4: return launch_net_t<2, false>(p, num_sms, stream);
        case 5: return launch_net_t<2, true>(p, num_sms, stream);
        case 6: return launch_net_t<3, false>(p, num_sms, stream);
        case 7: return launch_net_t<3, true>(p, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

size_t net_smem_bytes_rt(int arch) {
    return arch == 1 ? net_smem_bytes<1>() : arch == 2 ? net_smem_bytes<2>() : net_smem_bytes<3>();
}

}  // namespace rced
