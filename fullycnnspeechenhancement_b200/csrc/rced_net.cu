// K2: the fused R-CED / CR-CED network kernel (sm_100a).
//
// Replaces FullyCNNTester.test_step / sess.run(pred) of the reference
// (model_utils/tester.py:85-90) for the graphs of model_utils/model.py.
//
// Design:
//  * persistent grid, one CTA of 8 warps per SM.  Each SM sub-partition runs one FRAME PIPELINE
//    made of the two warps w and w+4: they take one spectrogram frame through all 10/16 layers
//    and split every layer's output channels in halves (part = w / 4).  Two warps per scheduler
//    hide each other's shared-memory and pipeline latencies;
//  * lane l owns frequency bins 4l..4l+3 for its part's channels; the register tile is
//    4 bins x (channels / 2) PAIRS and the inner loop is the packed FP32 FMA of Blackwell
//    (fma.rn.f32x2 -> FFMA2: the activation is broadcast, the weight pair comes straight out of
//    an LDS.128, the accumulator pair is a 64-bit register) -- half the issue slots and
//    register-file reads of scalar FFMA per flop, so loads issue in the FMA pipe's shadow;
//    bin 128 is a small extra phase with lane == output channel;
//  * all BN-folded weights (about 145 KB, laid out per part) sit in shared memory for the whole
//    kernel, brought in once per CTA with bulk async copies (cp.async.bulk + mbarrier);
//  * layer activations live in a per-frame shared-memory slot and are overwritten in place (the
//    whole layer output is held in registers before the first store); the two warps of a frame
//    meet at a 64-thread named barrier before and after the stores -- there is no CTA-wide
//    barrier anywhere in the layer loop;
//  * encoder outputs needed later by decoder skip connections are parked in TENSOR MEMORY
//    (tcgen05.st), thread-private, and loaded back (tcgen05.ld) straight into the accumulator
//    registers of the consuming decoder layer.  Global memory only sees the input magnitude
//    rows and the output rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rced_arch.cuh"
#include "rced_internal.h"
#include "rced_slots.cuh"

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
// packed FP32 pairs (fma.rn.f32x2 -> FFMA2) and the 64-thread frame barrier
// ------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d += {x, x} * w      (ptxas encodes the duplicated activation as a broadcast .F32 operand)
__device__ __forceinline__ void fma2_bcast(u64& d, float x, u64 w) {
    const u64 xx = pack2(x, x);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(xx), "l"(w));
}
// d += a * b, both packed
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
// the kSplit warps of one frame pipeline (named barrier 1 + frame slot, 32 * kSplit threads)
__device__ __forceinline__ void frame_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * kSplit) : "memory");
}

// ------------------------------------------------------------------------------------------
// one conv_bn_relu layer (all but the last) for one frame; executed by the kSplit warps of the
// frame, `part` selects this warp's output channels [part * CH, part * CH + CH)
// ------------------------------------------------------------------------------------------
template <int ARCH, bool TM, int LI>
__device__ __forceinline__ void conv_layer(const float* __restrict__ sW, float* __restrict__ slot, const int lane,
                                           const int part, const int bar_id, const SkipStore<TM>& sk) {
    constexpr LSpec S = spec(ARCH, LI);
    constexpr int NL = num_layers(ARCH);
    constexpr int CIN = cin_eff(ARCH, LI);
    constexpr int COUT = S.cout;
    constexpr int CH = ch_part(ARCH, LI);              // channels of this warp (even)
    constexpr int NP = CH / 2;                         // channel pairs
    constexpr int CIB = ci_block(ARCH, LI);            // weight floats per input channel per part
    constexpr int KW = S.kw;
    constexpr int PADL = (KW - 1) / 2;
    constexpr bool WIDEWIN = PADL > 4;                 // window of 20 floats instead of 12
    constexpr int NX4 = WIDEWIN ? 5 : 3;
    constexpr int XB = (WIDEWIN ? 8 : 4) - PADL;       // x[XB + f + k] is bin 4l+f+k-PADL
    constexpr bool OUT_WIDE = (LI == NL - 2);          // feeds the (1,129) layer
    constexpr bool PRE_ADD = (S.add >= 0) && !S.after; // skip pre-loaded into the accumulators
    constexpr bool POST_ADD = (S.add >= 0) && S.after; // V3: added after the ReLU
    // Both parts run the SAME instruction stream with run-time offsets: they share one
    // scheduler's instruction cache.  (Measured alternatives, tools/k2_bench: per-part
    // instantiations with an uneven 7+6 pair split, or predicating off the padded pair of the
    // last part, both cost 6-8% -- instruction-cache misses / predicated FFMA2 still take the pipe.)
    static_assert(PADL <= 7, "window loader covers SAME pads up to 7");
    static_assert(CH <= 32, "tail phase maps output channels to lanes");

    const float* __restrict__ W = sW + packed_w_off(ARCH, LI) + part * (CIN * CIB);
    const float* __restrict__ B = sW + packed_b_off(ARCH, LI) + part * pad4(CH);
    const float* __restrict__ in0 = slot + (LI == 0 ? stage_row(ARCH) * kRS : 0);
    const float* __restrict__ inx = in0 + (WIDEWIN ? 0 : 4) + 4 * lane;
    const int cbase = part * CH;                       // first output channel of this warp
    const int cl = lane < CH ? lane : CH - 1;          // tail phase: this lane's channel (local)
    const int col_add = S.add >= 0 ? skip_col_base(ARCH, S.add >= 0 ? S.add : 0) + part * (4 * CH + 1) : 0;
    const int col_save = S.save >= 0 ? skip_col_base(ARCH, S.save >= 0 ? S.save : 0) + part * (4 * CH + 1) : 0;

    u64 acc[4][NP];      // acc[f][j] = channels (2j, 2j+1) of bin 4l+f
    float tacc;

    if constexpr (PRE_ADD) {
        float s[4][CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) sk.ld4(col_add + 4 * c, s[0][c], s[1][c], s[2][c], s[3][c]);
        sk.ld1(col_add + 4 * CH, tacc);
        sk.wait_ld();
#pragma unroll
        for (int c = 0; c < CH; ++c) reg_fence4(s[0][c], s[1][c], s[2][c], s[3][c]);
        reg_fence1(tacc);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const float b0 = B[2 * j], b1 = B[2 * j + 1];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f][j] = pack2(s[f][2 * j] + b0, s[f][2 * j + 1] + b1);
        }
        tacc += B[cl];
    } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const u64 b = pack2(B[2 * j], B[2 * j + 1]);
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f][j] = b;
        }
        tacc = B[cl];
    }

    // ---- main phase: bins 4l..4l+3, this warp's channel pairs -------------------------------
    // Software pipeline over the input channels: the activation window and the first NPRE
    // weight words of channel ci+1 are loaded while channel ci is being multiplied.  (Row CIN
    // exists in the slot; what is read from it is never used.)  Loop bodies of up to 140 packed
    // FMAs are unrolled twice; larger unrolls lose to the instruction cache.  Measured and
    // dropped (tools/k2_bench.cu): no prefetch (-0.5 %), deeper weight prefetch, f-outer loop
    // order (ptxas re-schedules the packed FMAs into weight-reuse order whatever the source says),
    // broadcast LDS instead of SHFL for bin 128, holding part 1 back to de-phase the two warps
    // (all within +-1 %), unroll 3 / 4 (-8 % / -11 %).
    constexpr int NQ = CIB / 4;                        // 16-byte weight words per input channel
    constexpr int NPRE = NQ < 2 ? NQ : 2;
    constexpr int BODY = KW * NP * 4;                  // packed FMAs per input channel
    constexpr int UNR = BODY <= 140 ? 2 : 1;
    const float* wt = W + cl;                          // bin-128 phase: this lane's channel
    float4 xn[NX4];
    ulonglong2 wn[NPRE];
    {
        const ulonglong2* wq = reinterpret_cast<const ulonglong2*>(W);
#pragma unroll
        for (int i = 0; i < NPRE; ++i) wn[i] = wq[i];
    }
    // Everything above (bias / skip pre-load, first weight words) is independent of the layer
    // input, so it overlaps the partner warp's stores of the previous layer; from here on the
    // input rows are read.
    frame_bar(bar_id);   // the previous layer's output (or the staged frame) is complete and visible
    {
        const float4* xp = reinterpret_cast<const float4*>(inx);
#pragma unroll
        for (int i = 0; i < NX4; ++i) xn[i] = xp[i];
    }
#pragma unroll UNR
    for (int ci = 0; ci < CIN; ++ci) {
        float x[NX4 * 4];
        ulonglong2 wc[NPRE];
#pragma unroll
        for (int i = 0; i < NX4; ++i) {
            x[4 * i + 0] = xn[i].x; x[4 * i + 1] = xn[i].y; x[4 * i + 2] = xn[i].z; x[4 * i + 3] = xn[i].w;
        }
#pragma unroll
        for (int i = 0; i < NPRE; ++i) wc[i] = wn[i];
        {
            const float4* xp = reinterpret_cast<const float4*>(inx + (ci + 1) * kRS);
#pragma unroll
            for (int i = 0; i < NX4; ++i) xn[i] = xp[i];
        }
        // [kw][CH] weights of this input channel as one flat run of pairs; warp-uniform
        // addresses (broadcast LDS.128 = two pairs), a pair never straddles a 16-byte word
        const ulonglong2* wq = reinterpret_cast<const ulonglong2*>(W + ci * CIB);
#pragma unroll
        for (int k = 0; k < KW; ++k) {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int P = k * NP + j;
                const ulonglong2 q = (P >> 1) < NPRE ? wc[(P >> 1) < NPRE ? (P >> 1) : 0] : wq[P >> 1];
                const u64 w = (P & 1) ? q.y : q.x;
#pragma unroll
                for (int f = 0; f < 4; ++f) fma2_bcast(acc[f][j], x[XB + f + k], w);
            }
        }
        // bin 128 (lane == output channel): its PADL+1 taps read bins 128-PADL..128, which lane 31
        // holds in its window; the scalar FMAs ride in the issue slots the packed FMAs leave free
#pragma unroll
        for (int k = 0; k <= PADL; ++k) {
            const float xs = __shfl_sync(0xffffffffu, x[XB + 4 + k], 31);
            tacc = fmaf(xs, wt[ci * CIB + k * CH], tacc);
        }
#pragma unroll
        for (int i = 0; i < NPRE; ++i) wn[i] = wq[NQ + i];   // first words of channel ci+1
    }

    // ---- epilogue ----------------------------------------------------------------------------
    float tsk = 0.f;
    if constexpr (POST_ADD) sk.ld1(col_add + 4 * CH, tsk);   // V3: relu first, then add the skip (no second relu)
    frame_bar(bar_id);   // every lane of the frame has finished reading this layer's input (overwritten below)

    if constexpr (OUT_WIDE) {
        // the (1,129) layer reads rows of stride kWS with 64 zeros either side: clear, then fill
        float4* z = reinterpret_cast<float4*>(slot);
        for (int i = part * 32 + lane; i < wide_floats(ARCH) / 4; i += 32 * kSplit) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        frame_bar(bar_id);
    }

#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float v[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float lo, hi;
            unpack2(acc[f][c >> 1], lo, hi);
            v[f] = (c & 1) ? hi : lo;
            if constexpr (S.relu) v[f] = fmaxf(v[f], 0.f);
        }
        if constexpr (POST_ADD) {
            float s0, s1, s2, s3;
            sk.ld4(col_add + 4 * c, s0, s1, s2, s3);
            sk.wait_ld();
            reg_fence4(s0, s1, s2, s3);
            v[0] += s0; v[1] += s1; v[2] += s2; v[3] += s3;
        }
        if constexpr (S.save >= 0) sk.st4(col_save + 4 * c, v[0], v[1], v[2], v[3]);
        if (cbase + c < COUT) {   // channels beyond cout are the zero padding of the last part
            float* o = OUT_WIDE ? slot + (cbase + c) * kWS + kWideBin0 + 4 * lane : slot + (cbase + c) * kRS + kRowBin0 + 4 * lane;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    {
        float t = tacc;
        if constexpr (S.relu) t = fmaxf(t, 0.f);
        if constexpr (POST_ADD) {
            sk.wait_ld();
            reg_fence1(tsk);
            t += tsk;
        }
        if constexpr (S.save >= 0) sk.st1(col_save + 4 * CH, t);
        if (lane < CH && cbase + lane < COUT) {
            float* o = OUT_WIDE ? slot + (cbase + lane) * kWS + kWideBin0 + 128 : slot + (cbase + lane) * kRS + kRowBin0 + 128;
            *o = t;
        }
    }
    if constexpr (S.save >= 0) sk.wait_st();
    // no barrier here: the consumer (next layer / final layer) waits right before it reads
}

// ------------------------------------------------------------------------------------------
// the (1,129) output layer: cout = 1, no BN / ReLU; reads the wide layout, writes global memory.
// The parts split the INPUT channels; part 1 hands its partial sums to part 0 through `comb`.
// Packed pairs run along the taps: bin b pairs taps (k, k+1) with b + k even, so that both the
// activation pair and the weight pair are aligned 64-bit words -- even bins read W, odd bins
// read S (S[t] = W[t+1]) and take their tap 0 as a scalar FMA, even bins their tap 128.
// ------------------------------------------------------------------------------------------
template <int ARCH>
__device__ __forceinline__ void final_layer(const float* __restrict__ sW, float* __restrict__ slot, const int lane,
                                            const int part, const int bar_id, float* __restrict__ out_row) {
    constexpr int LI = num_layers(ARCH) - 1;
    constexpr int CIN = spec(ARCH, LI).cin;
    constexpr int CPART = (CIN + kSplit - 1) / kSplit;
    const float* __restrict__ Wf = sW + packed_w_off(ARCH, LI);
    const float* __restrict__ Sf = Wf + CIN * kFinalKP;
    const float bias = sW[packed_b_off(ARCH, LI)];
    float* comb = slot + combine_off(ARCH);
    const int c0 = part * CPART;
    const int c1 = c0 + CPART < CIN ? c0 + CPART : CIN;

    frame_bar(bar_id);   // the wide layout written by the last conv layer is complete
    u64 a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;   // (even-tap, odd-tap) partial sums of bins 4l..4l+3
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;     // the unpaired taps
#pragma unroll 1
    for (int ci = c0; ci < c1; ++ci) {
        // out bin 4l+j, tap k reads wide offset 4l + j + k
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(slot + ci * kWS + 4 * lane);
        const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(Wf + ci * kFinalKP);
        const ulonglong2* s4 = reinterpret_cast<const ulonglong2*>(Sf + ci * kFinalKP);
        ulonglong2 xa = row[0];
        {
            float x0, x1, x2, x3;
            unpack2(xa.x, x0, x1);
            unpack2(xa.y, x2, x3);
            const float w0 = Wf[ci * kFinalKP];
            l1 = fmaf(x1, w0, l1);
            l3 = fmaf(x3, w0, l3);
        }
#pragma unroll 8
        for (int q = 0; q < 32; ++q) {
            const ulonglong2 xb = row[q + 1];
            const ulonglong2 w = w4[q];
            const ulonglong2 s = s4[q];
            fma2(a0, xa.x, w.x); fma2(a0, xa.y, w.y);
            fma2(a2, xa.y, w.x); fma2(a2, xb.x, w.y);
            fma2(a1, xa.y, s.x); fma2(a1, xb.x, s.y);
            fma2(a3, xb.x, s.x); fma2(a3, xb.y, s.y);
            xa = xb;
        }
        {
            float x0, x1, x2, x3;
            unpack2(xa.x, x0, x1);
            unpack2(xa.y, x2, x3);
            const float w128 = Wf[ci * kFinalKP + 128];
            l0 = fmaf(x0, w128, l0);
            l2 = fmaf(x2, w128, l2);
        }
    }
    // bin 128: taps 0..64 over wide offsets 128..192, split across lanes, shuffle-reduced
    float t = 0.f;
#pragma unroll 2
    for (int ci = c0; ci < c1; ++ci) {
        const float* r = slot + ci * kWS + 128;
        const float* w = Wf + ci * kFinalKP;
        t = fmaf(r[lane], w[lane], t);
        t = fmaf(r[lane + 32], w[lane + 32], t);
        if (lane == 0) t = fmaf(r[64], w[64], t);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);

    float e, o, r0, r1, r2, r3;
    unpack2(a0, e, o); r0 = (e + o) + l0;
    unpack2(a1, e, o); r1 = (e + o) + l1;
    unpack2(a2, e, o); r2 = (e + o) + l2;
    unpack2(a3, e, o); r3 = (e + o) + l3;

    if (part != 0) {
        *reinterpret_cast<float4*>(comb + 4 * lane) = make_float4(r0, r1, r2, r3);
        if (lane == 0) comb[128] = t;
    }
    frame_bar(bar_id);   // partial sums are in `comb`; nobody reads the wide layout any more
    if (part == 0) {
        const float4 c = *reinterpret_cast<const float4*>(comb + 4 * lane);
        float* op = out_row + 4 * lane;
        op[0] = (r0 + c.x) + bias; op[1] = (r1 + c.y) + bias; op[2] = (r2 + c.z) + bias; op[3] = (r3 + c.w) + bias;
        if (lane == 0) out_row[128] = (t + comb[128]) + bias;
    }
}

template <int ARCH, bool TM, int LI>
__device__ __forceinline__ void run_conv_layers(const float* sW, float* slot, int lane, int part, int bar_id,
                                                const SkipStore<TM>& sk) {
    if constexpr (LI < num_layers(ARCH) - 1) {
        conv_layer<ARCH, TM, LI>(sW, slot, lane, part, bar_id, sk);
        run_conv_layers<ARCH, TM, LI + 1>(sW, slot, lane, part, bar_id, sk);
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
// stage_row .. stage_row+7 of the slot with asynchronous 4-byte copies; each part of the frame
// brings in 8 / kSplit of the rows.
template <int ARCH>
__device__ __forceinline__ void prefetch_frame(const NetParams& p, long long g, float* slot, int lane, int part) {
    constexpr int SR = stage_row(ARCH);
    constexpr int RPP = 8 / kSplit;
    const FrameLoc loc = locate(p.row_off, p.n_utt, g);
    // halo offsets 1..7 of the 9 rows touched (offset 0 is bin 128 of the row before)
    for (int i = part * 32 + lane; i < 9 * 7; i += 32 * kSplit) slot[(SR + i / 7) * kRS + 1 + (i % 7)] = 0.f;
#pragma unroll
    for (int d = 0; d < RPP; ++d) {
        const int dt = part * RPP + d;
        const long long r = g + dt - 3;
        float* dst = slot + (SR + dt) * kRS + kRowBin0;
        if (r >= loc.lo && r < loc.hi) {
            const float* src = p.in + r * (long long)kBins;
            const uint32_t da = smem_u32(dst + 4 * lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async4(da + 4 * j, src + 4 * lane + j);
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
    __shared__ int s_slot;

    // fall-back role behind the tensor-core variant: nothing to do unless its guard tripped
    if (p.guard != nullptr && p.guard[0] <= 0x477FE000u /* 65504.0f */ && p.guard[1] == 0u) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fs = warp & (kFramesPerCta - 1);   // frame slot == SM sub-partition == tensor-memory lane quadrant
    const int part = warp / kFramesPerCta;       // which share of every layer's output channels
    const int bar_id = 1 + fs;
    float* slot = slots + fs * SLOT;

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
    // meanwhile: clear this frame's slot (all halos must read as zero)
    for (int i = part * 32 + lane; i < SLOT / 4; i += 32 * kSplit)
        reinterpret_cast<float4*>(slot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    SkipStore<TM> sk;
    if constexpr (TM) {
        if (warp == 0) tmem_alloc512(smem_u32(&s_tmem));
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        sk.base = s_tmem + ((uint32_t)fs << 21);   // lane field (bits 31:16) = 32 * (warp % 4)
    } else {
        // launches of one handle may overlap (several streams): the CTA claims a region of the scratch
        if (threadIdx.x == 0) s_slot = scratch_slot_acquire(p.slot_busy, p.n_slots);
        __syncthreads();
        if (s_slot < 0) __trap();   // every region busy: a kernel was killed between claim and release
        sk.base = p.skip_scratch + ((size_t)s_slot * kFramesPerCta + fs) * (512 * 32) + lane;
    }
    mbar_wait(bar, 0);

    const long long stride = (long long)gridDim.x * kFramesPerCta;
    long long g = (long long)blockIdx.x * kFramesPerCta + fs;
    if (g < p.total_rows) prefetch_frame<ARCH>(p, g, slot, lane, part);

    for (; g < p.total_rows; g += stride) {
        cp_async_wait_all();   // this thread's share of the input rows has landed; the first layer's
                               // barrier makes both shares and the restored halos visible
        run_conv_layers<ARCH, TM, 0>(sW, slot, lane, part, bar_id, sk);
        // the wide layout sits below row SR: the next frame's input can land while the last layer runs
        const long long gn = g + stride;
        if (gn < p.total_rows) prefetch_frame<ARCH>(p, gn, slot, lane, part);
        final_layer<ARCH>(sW, slot, lane, part, bar_id, p.out + g * (long long)kBins);
        // the wide layout overwrote the halos of rows 0..SR: restore their zeros
        for (int i = part * 32 + lane; i < (SR + 1) * 7; i += 32 * kSplit) slot[(i / 7) * kRS + 1 + (i % 7)] = 0.f;
    }

    if constexpr (TM) {
        tmem_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc512(s_tmem);
    } else {
        __syncthreads();
        if (threadIdx.x == 0) scratch_slot_release(p.slot_busy, s_slot);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int ARCH>
constexpr size_t net_smem_bytes() {
    return (size_t)(pad4(packed_count(ARCH)) + kFramesPerCta * slot_floats(ARCH)) * sizeof(float);
}

template <int ARCH, bool TM>
static cudaError_t launch_net_t(const NetParams& p, int num_sms, cudaStream_t stream) {
    constexpr size_t smem = net_smem_bytes<ARCH>();
    static_assert(smem <= 227 * 1024, "weights + activation slots must fit one SM's shared memory");
    static_assert(slot_floats(ARCH) % 4 == 0 && pad4(packed_count(ARCH)) % 4 == 0, "16-byte aligned slots");
    cudaError_t e = cudaFuncSetAttribute(rced_net_kernel<ARCH, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long ctas = (p.total_rows + kFramesPerCta - 1) / kFramesPerCta;
    if (ctas > num_sms) ctas = num_sms;
    if (ctas < 1) return cudaSuccess;
    rced_net_kernel<ARCH, TM><<<(unsigned)ctas, kWarpsPerCta * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_net(int arch, bool skip_in_tmem, const NetParams& p, int num_sms, cudaStream_t stream) {
#ifdef RCED_BENCH_ONLY_V2   // tools/k2_bench.cu: one instantiation keeps experiment builds short
    return arch == 2 && skip_in_tmem ? launch_net_t<2, true>(p, num_sms, stream) : cudaErrorInvalidValue;
#else
    switch (arch * 2 + (skip_in_tmem ? 1 : 0)) {
        case 2: return launch_net_t<1, false>(p, num_sms, stream);
        case 3: return launch_net_t<1, true>(p, num_sms, stream);
        case 4: return launch_net_t<2, false>(p, num_sms, stream);
        case 5: return launch_net_t<2, true>(p, num_sms, stream);
        case 6: return launch_net_t<3, false>(p, num_sms, stream);
        case 7: return launch_net_t<3, true>(p, num_sms, stream);
    }
    return cudaErrorInvalidValue;
#endif
}

size_t net_smem_bytes_rt(int arch) {
    return arch == 1 ? net_smem_bytes<1>() : arch == 2 ? net_smem_bytes<2>() : net_smem_bytes<3>();
}

}  // namespace rced
