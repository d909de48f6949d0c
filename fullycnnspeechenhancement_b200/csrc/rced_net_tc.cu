// K2-TC: tensor-core variant of the fused R-CED / CR-CED network kernel (sm_100a, tcgen05).
//
// Same contract as rced_net_kernel (rced_net.cu): replaces FullyCNNTester.test_step /
// sess.run(pred) of the reference (model_utils/tester.py:85-90) for the graphs of
// model_utils/model.py, one fused kernel, activations never leave the SM.
//
// Design (rced_tc.cuh):
//  * a CTA takes a batch of 7 consecutive spectrogram frames through all layers.  The (frame, bin)
//    rows are flattened with a stride of 136 rows per frame, so that the 7 zero rows between two
//    frames are the SAME-padding halo of both; 8 row tiles of M = 128 cover the batch.  (Launches too
//    small to give every SM two such batches take fewer frames per batch -- TcParams.fb / nt --, so
//    that the latency of one utterance is the time of a 2- or 3-tile batch, not of an 8-tile one);
//  * every conv layer is an implicit GEMM without im2col: D[row][cout] += A_tap[row][cin] *
//    W_tap[cout][cin], one tcgen05.mma (kind::f16, K = 16) per pair of (tap, 8-channel group)
//    chunks, the A descriptor of a tap being the activation plane shifted by (tap - pad) rows;
//  * FP32 accuracy from FP16 tensor cores by an error-compensated split: activations and
//    weights are stored as hi + lo FP16 pairs and every K step issues A_hi x [Whi | Wlo]
//    (N = 2 NP) and A_lo x Whi (N = NP) into FP32 accumulators in tensor memory (two column blocks per
//    tile: hi*Whi and the two cross products, which carry the residuals' scale).  The residuals are
//    stored times 2^11, the weights of a step times a power of two and every frame in its own
//    power-of-two scaled domain (rced_tc.cuh, "Range"), so that the FP16 operands stay in their normal
//    range for inputs and weights of any magnitude;
//  * warp roles (rced_tc.cuh): warps 0, 3, 5 issue the MMAs (one elected lane each, row tiles of the
//    global (step, tile) sequence round robin, all descriptor arithmetic in the uniform datapath from
//    constant-memory tables), warp 1 streams the next layers' weight tiles from L2 with bulk async
//    copies into a double buffer, warp 2 is the dependency scout (waits on the mbarriers of every
//    (step, tile) in issue order and publishes a counter), warp 4 prefetches the next batch's input
//    rows, warps 6-21 are four epilogue groups (tcgen05.ld -> bias, skip, ReLU -> hi/lo split ->
//    st.shared of the next layer's planes).  Layers overlap tile by tile: the epilogue of (layer,
//    tile t) starts when the MMAs of tiles t-1, t, t+1 have completed (they read tile t's rows, the
//    neighbours as halo; planes are updated in place) and the MMAs of (layer+1, t) start when the
//    epilogues of tiles t-1..t+1 are done.  The boundary between two batches is such a layer boundary
//    too: the output layer's epilogue of row tile t stages the next batch's first-layer rows of tile t;
//  * the (1,129) output layer runs "taps in N with row-shifted accumulation" (rced_tc.cuh): five MMAs
//    per row tile whose A descriptors are shifted by -64 .. +64 rows accumulate E[r][n] =
//    sum_i X[r + 32 (i - 2)] . W[32 i + n] into 32 columns, and the epilogue forms out[b] =
//    sum_n E[b + n][n] with one shuffle per column (fixed order).  The shifts would reach the
//    neighbouring frames, so the last conv layer writes an even-frames-only and an odd-frames-only
//    copy of its output and each parity has its own accumulator columns;
//  * skip tensors go to FP32 rows in a region of a global scratch (L2 resident) that the CTA claims
//    when it starts (rced_slots.cuh: one scratch per handle serves any number of streams);
//  * range guard: FP16 overflows beyond 65504.  The kernel records the largest |activation| it
//    stored (in the frames' scaled domains) and whether an input was not finite; rced_forward
//    re-runs the call with the FP32 FFMA kernel when the guard tripped.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "rced_internal.h"
#include "rced_slots.cuh"
#include "rced_tc.cuh"

namespace rced {
namespace tc {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
#ifdef RCED_TC_DIAG_RELAXED
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
// Bounded wait: a protocol error must end the kernel with an error flag, not hang the GPU.  Once
// any wait has timed out every other wait of the grid gives up at its next check of the flag.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* err, int code) {
    // fast path: the non-blocking test costs a fraction of try_wait (tools/umma_probe lat: 168 cycles
    // for a try_wait on a phase that is already complete)
    if (mbar_test(bar, parity)) return;
    for (int it = 0; it < (1 << 21); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((it & 255) == 255 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) return;
    }
    atomicMax(err, (unsigned int)code);
}
__device__ __forceinline__ void st_release(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// Both barriers complete: the two tests of a round are in flight together, so the wake-up costs one
// test latency (~160 cycles) after the later arrival instead of two.
#ifndef RCED_TC_EPIWAIT
#define RCED_TC_EPIWAIT 1   // 1 (default) blocking try_wait after one test of both; experiment: 0 spin on test_wait, 2 test_wait + nanosleep
#endif
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Three barriers complete (two of them may be the same): the tests of a round are in flight together,
// so the common case costs one test latency (~160 cycles); what is not complete yet is waited for
// with the blocking try_wait, which does not take issue slots from the MMA-issuing warps.
__device__ __forceinline__ void mbar_wait3(uint32_t bar_a, uint32_t bar_b, uint32_t bar_c, uint32_t parity, unsigned int* err, int code) {
#if RCED_TC_EPIWAIT == 1
    bool a = mbar_test(bar_a, parity);
    bool b = mbar_test(bar_b, parity);
    bool c = mbar_test(bar_c, parity);
    if (a && b && c) return;
    for (int it = 0; it < (1 << 21); ++it) {
        if (!a) a = mbar_try(bar_a, parity);
        if (!b) b = mbar_try(bar_b, parity);
        if (!c) c = mbar_try(bar_c, parity);
        if (a && b && c) return;
        if ((it & 63) == 63 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) return;
    }
#else
    for (int it = 0; it < (1 << 22); ++it) {
        const bool a = mbar_test(bar_a, parity);
        const bool b = mbar_test(bar_b, parity);
        const bool c = mbar_test(bar_c, parity);
        if (a && b && c) return;
#if RCED_TC_EPIWAIT == 2
        __nanosleep(RCED_TC_EPISLEEP);
#endif
        if ((it & 255) == 255 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) return;
    }
#endif
    atomicMax(err, (unsigned int)code);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers written by tcgen05.ld may only be read after tcgen05.wait::ld (see rced_net.cu)
__device__ __forceinline__ void reg_fence8(float (&v)[8]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])::"memory");
}
// skip scratch traffic with an L2 eviction-priority hint (RCED_TC_SKIPHINT: experiment switch)
#ifndef RCED_TC_MAXINFLIGHT
#define RCED_TC_MAXINFLIGHT 0   // experiment switch: row tiles being issued concurrently (0: no limit; a limit of 1 / 2 costs 2x / 25 %)
#endif
#ifndef RCED_TC_SKIPHINT
#define RCED_TC_SKIPHINT 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
#if RCED_TC_SKIPHINT == 2   // the hinted instruction forms with the default priority (experiment)
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_hint4(const float4* p, uint64_t pol) {
#if RCED_TC_SKIPHINT == 3   // untested candidate (round 2): skip rows are read once -- keep them out of the L1
    (void)pol;
    return __ldcg(p);
#elif RCED_TC_SKIPHINT
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_hint4(float4* p, const float4 v, uint64_t pol) {
#if RCED_TC_SKIPHINT == 3
    (void)pol;
    __stcg(p, v);
#elif RCED_TC_SKIPHINT
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_hint1(float* p, const float v, uint64_t pol) {
#if RCED_TC_SKIPHINT == 1 || RCED_TC_SKIPHINT == 2
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
// Skip tensors by bulk copy (experiment switch RCED_TC_SKIP_BULK=1, off): a saving layer's rows go from the FP16
// hi / lo planes to the global scratch with cp.async.bulk issued AFTER the tile has been released -- the async proxy
// reads the planes, no thread has a global store outstanding when the next proxy fence (MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC) runs.  Measured on B200 (tools/k2tc_experiments.sh, V2, 254,976 frames): 15.30 ms against
// 13.59 ms for the store form (FP32 rows written from the registers in front of the fence; no skip traffic at
// all: 13.16 ms) -- 1,088 512-byte copies per batch are slower than the L2 round trip they avoid.
#ifndef RCED_TC_SKIP_BULK
#define RCED_TC_SKIP_BULK 0
#endif
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (version 1): start and LBO in the low word,
// SBO = 128 bytes (rows of an 8-row group are 16 bytes apart, groups follow each other) in the high
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t make_desc(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }
// instruction descriptor: FP16 A/B (format 0), FP32 accumulate, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// x = hi + 2^-11 lo as two FP16 pairs (the residual is stored times 2^11: it has the magnitude of x
// itself and stays a normal FP16 number whenever hi is one); element a goes to the low half
constexpr float kLoScale = (float)(1 << kLoShift), kLoInv = 1.f / (float)(1 << kLoShift);
// packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): the epilogue is bound by its issue slots, and
// a packed instruction does the work of two
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    float ra, rb;
    unpack2(mul2(sub2(pack2(a, b), pack2(hf.x, hf.y)), pack2(kLoScale, kLoScale)), ra, rb);
    const __half2 l = __floats2half2_rn(ra, rb);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// largest |value| of the packed FP16 pairs seen so far (the range guard: an overflow is an infinity here)
__device__ __forceinline__ void track_max(uint32_t& amax2, uint32_t h) {
    asm("{\n\t.reg .b32 t;\n\tabs.f16x2 t, %1;\n\tmax.f16x2 %0, %0, t;\n\t}" : "+r"(amax2) : "r"(h));
}
// rows that are not (frame, bin) rows of the batch -- the halo rows between frames and everything
// behind the last frame -- are stored as zeros whatever was computed for them
__device__ __forceinline__ void store_split8(unsigned char* act, int cg, int q, const float (&v)[8], bool valid, uint32_t& amax2) {
    uint4 h, l;
    split2(v[0], v[1], h.x, l.x);
    split2(v[2], v[3], h.y, l.y);
    split2(v[4], v[5], h.z, l.z);
    split2(v[6], v[7], h.w, l.w);
    track_max(amax2, h.x);
    track_max(amax2, h.y);
    track_max(amax2, h.z);
    track_max(amax2, h.w);
    uint4* ph = reinterpret_cast<uint4*>(act) + cg * kPlane16 + q;
    if (valid) {
        ph[0] = h;
        ph[kLo16] = l;
    } else {
        ph[0] = make_uint4(0, 0, 0, 0);
        ph[kLo16] = make_uint4(0, 0, 0, 0);
    }
}

// The last conv layer feeds the row-shifted output layer: its rows are stored twice, in planes
// (0, 1) if the row's frame index is even and in planes (2, 3) if it is odd, zeros in the other pair
__device__ __forceinline__ void store_split8_parity(unsigned char* act, int cg, int q, const float (&v)[8], bool valid, int odd,
                                                    uint32_t& amax2) {
    uint4 h, l;
    split2(v[0], v[1], h.x, l.x);
    split2(v[2], v[3], h.y, l.y);
    split2(v[4], v[5], h.z, l.z);
    split2(v[6], v[7], h.w, l.w);
    track_max(amax2, h.x);
    track_max(amax2, h.y);
    track_max(amax2, h.z);
    track_max(amax2, h.w);
    if (!valid) h = l = make_uint4(0, 0, 0, 0);
    uint4* pe = reinterpret_cast<uint4*>(act) + cg * kPlane16 + q;   // even-frame copy
    uint4* po = pe + 2 * kPlane16;                                   // odd-frame copy
    const uint4 z = make_uint4(0, 0, 0, 0);
    pe[0] = odd ? z : h;
    pe[kLo16] = odd ? z : l;
    po[0] = odd ? h : z;
    po[kLo16] = odd ? l : z;
}

// Per-step issue parameters and the units' A-descriptor words in CONSTANT memory: the issuing thread
// indexes them with warp-uniform loop counters, so the loads and all descriptor arithmetic run in the
// uniform datapath (LDCU / UIADD3) and land in the uniform registers tcgen05.mma takes its operands
// from -- no per-unit R2UR moves, no per-step prologue.  a[s][u] = (LBO << 16) + first-chunk offset in
// 16-byte units relative to row kLead of plane 0 (may be negative: the sum with the plane address is not).
template <int ARCH>
struct IssueTab {
    int a[n_steps(ARCH) * kTabStride];
    int nu[n_steps(ARCH)], np[n_steps(ARCH)], tile16[n_steps(ARCH)], rows[n_steps(ARCH)];
    constexpr IssueTab() : a{}, nu{}, np{}, tile16{}, rows{} {
        for (int s = 0; s < n_steps(ARCH); ++s) {
            nu[s] = step_units(ARCH, s);
#ifdef RCED_TC_DIAG_PACKEST   // diagnosis only (wrong results): the units a tap-packed partial channel group would save (V2)
            if (ARCH == 2) {
                const int cheap[16] = {0, 2, 0, 0, 0, 2, 0, 0, 3, 0, 0, 2, 0, 0, 0, 0};
                const int more[16] = {0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 2, 0};
                nu[s] -= cheap[s] + (RCED_TC_DIAG_PACKEST > 1 ? more[s] : 0);
            }
#endif
            np[s] = step_np(ARCH, s);
            tile16[s] = step_tile_bytes(ARCH, s) >> 4;
            rows[s] = step_tile_rows(ARCH, s);
            for (int u = 0; u < kTabStride; ++u) {
                const int uu = u < step_units(ARCH, s) ? u : 0;
                a[s * kTabStride + u] = unit_lbo16(ARCH, s, uu) * 65536 + kLead + unit_off16(ARCH, s, uu);
            }
        }
    }
};
template <int ARCH>
__constant__ IssueTab<ARCH> c_issue = IssueTab<ARCH>();

struct TcParams {
    const unsigned char* wimg;   // weight image (global): per step, per unit, [2][rows][8] halfs
    const float* bias;           // [n_steps + 1][32]: biases per step, then the steps' c1 and the largest |bias| (rced_tc.cuh)
    const float* in;             // mag  [total_rows][129]
    float* out;                  // pred [total_rows][129]
    const long long* row_off;    // [n_utt + 1]
    int n_utt;
    long long total_rows;
    float* skip;                 // scratch for the skip tensors: n_slots regions of skip_floats_per_cta
    unsigned int* slot_busy;     // [n_slots] claim words of the regions (rced_slots.cuh)
    int n_slots;
    unsigned int* flags;         // [0] bits of the largest |activation| stored as FP16 (scaled domain; >= inf: an input was not finite), [1] protocol error
    long long* trace;            // development aid (RCED_TC_TRACE): clock64 stamps of CTA 0's second batch, or null
    // Launch shape: frames per CTA batch (kFB, or fewer when the launch is too small to give every SM a full batch: the
    // latency of one utterance or of a streaming block is the time of ONE batch, which shrinks with its row tiles) and the
    // row tiles that hold them (ceil(fb * 136 / 128)); uniform over the launch, so the (step, tile) sequence stays regular
    int fb, nt;
};
// trace slots: [step][tile][event]; events: 0 MMA issue begins, 1 MMA issued (commit), 2 epilogue
// past its waits, 3 accumulator in registers, 4 epilogue done (arrive); tile 0 only: 5 / 6 before /
// after the wait for the step's weights
// The development trace (RCED_TC_TRACE=<file>, tools/tc_trace_report.py) is compiled in only with -DRCED_TC_TRACING=1:
// the kernel is sensitive to the size of its hot loops -- the twelve predicated stamps cost 2 % (13.34 -> 13.05 ms).
#ifndef RCED_TC_TRACING
#define RCED_TC_TRACING 0
#endif
__device__ __forceinline__ void stamp(long long* trace, bool on, int s, int t, int ev) {
#if RCED_TC_TRACING
    if (on) trace[(s * kTiles + t) * kTraceEvents + ev] = clock64();
#endif
}

// what the epilogue of a conv step does, as data: one code body serves every layer (the
// per-layer template instantiations of the first version were ~100 KB of code and missed the
// instruction cache at the start of every step)
struct EpiStep {
    int np;        // accumulator column of the hi x Wlo block
    int cg;        // groups of 8 output channels
    int relu;
    int add;       // 0 none, 1 before the ReLU, 2 after it (V3)
    int add_base;  // first 8-channel group of the skip slot added
    int save_base; // first 8-channel group of the skip slot saved (-1: none)
    int last;      // the last conv layer: even / odd frame copies for the output layer
    float skip_r;  // power of two between the domain of the skip tensor added and this step's domain
};
static_assert(sizeof(EpiStep) == 32, "EpiStep layout");

struct Ctx {
    unsigned char* smem;
    unsigned char* act;
    uint32_t bars;      // shared address of the barrier block
    uint32_t tm;        // tensor-memory base
    const float* bias;  // shared copy
    const float* scale; // power-of-two scales of the batch's frames ([8], shared; written by the prefetch warp)
    const EpiStep* epi; // shared copy
    const long long* bnd;
    unsigned int* err;
    long long* trace;
    bool tracing;
    int lane, quad, grp, et, nf, nt;
    float* skip;
    float* outp;        // [2][kRows]: the two partial sums of every output row of the (1,129) layer
    long long g0;
    uint64_t pol_last, pol_first;   // L2 eviction-priority policies (skip scratch / streamed output)
};

__device__ __forceinline__ uint32_t bar_addr(const Ctx& c, int slot) { return c.bars + 8u * slot; }

// utterance [lo, hi) that owns global row g
__device__ __forceinline__ void locate(const long long* __restrict__ row_off, int n_utt, long long g, long long& lo, long long& hi) {
    int a = 0, b = n_utt;
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (__ldg(row_off + m) <= g) a = m; else b = m;
    }
    lo = __ldg(row_off + a);
    hi = __ldg(row_off + a + 1);
}

// ------------------------------------------------------------------------------------------
// epilogue of one conv layer for one row tile (runtime-parameterised, see EpiStep)
// ------------------------------------------------------------------------------------------
// AFTER: the model adds its skip tensors behind the ReLU (CR-CED V3) instead of in front of it (R-CED V1 / V2); a model has
// one kind or the other, so the body carries only one of the two forms
template <bool AFTER>
__device__ __forceinline__ void epi_conv_tile(const Ctx& c, const int s, const int t, const uint32_t par, uint32_t& amax2) {
    EpiStep e = c.epi[s];
#ifdef RCED_TC_DIAG_NOSKIP   // diagnosis only (wrong results): no skip traffic
    e.add = 0;
    e.save_base = -1;
#endif
#ifdef RCED_TC_DIAG_NOSAVE
    e.save_base = -1;
#endif
#ifdef RCED_TC_DIAG_NOADD
    e.add = 0;
#endif
    const int r = t * 128 + c.quad * 32 + c.lane;   // row in tile space
    const int fi = r / kFS, b = r - fi * kFS;
    const bool valid = fi < c.nf && b < kBins;
    const float sf = c.scale[fi];                    // the frame's scale: the (pre-scaled) biases enter its domain times sf
    const u64 sf2 = pack2(sf, sf), loinv2 = pack2(kLoInv, kLoInv), r2 = pack2(e.skip_r, e.skip_r);
    const uint32_t ta = c.tm + ((uint32_t)(c.quad * 32) << 16) + (uint32_t)(t * kAccCols);
    // Skip tensors: in the CTA's region of the global scratch (L2), [group of 8 channels][half][row][16 bytes], so
    // that a warp's 16-byte accesses cover whole sectors.  The row is saved and added by the same warp.  Bulk form:
    // half 0 = the row's 8 FP16 hi values, half 1 = its 8 scaled residuals (copies of the plane rows); store form:
    // the two halves are FP32 channels 0-3 and 4-7.
    const float4* sp = reinterpret_cast<const float4*>(c.skip) + ((size_t)e.add_base * 2 * kRows + r);
    float4* dp = reinterpret_cast<float4*>(c.skip) + ((size_t)(e.save_base < 0 ? 0 : e.save_base) * 2 * kRows + r);
    // the skip row does not depend on the accumulator: the L2 latency of its first group hides behind
    // the wait for the MMAs
#ifdef RCED_TC_DIAG_VALIDROWS   // measured: predicating the skip traffic on valid rows costs more than the 12 % of bytes it saves
    const bool do_add = e.add != 0 && valid, do_save = e.save_base >= 0 && valid;
#else
    const bool do_add = e.add != 0, do_save = e.save_base >= 0;
#endif
    float4 sk0 = make_float4(0.f, 0.f, 0.f, 0.f), sk1 = sk0;
#if RCED_TC_SKIP_BULK
    if (do_add) {
        // the copies that saved the tensor have completed and their writes are visible to this warp (nothing is
        // pending after the first adding layer of a batch); the loads bypass the L1, which the copies do not update
        if (c.lane == 0) bulk_wait_all();
        __syncwarp();
        sk0 = __ldcg(sp);
        sk1 = __ldcg(sp + kRows);
    }
#else
    if (do_add) {
        sk0 = ld_hint4(sp, c.pol_last);
        sk1 = ld_hint4(sp + kRows, c.pol_last);
    }
#endif
    // The planes are updated in place: this tile's rows are read by its own MMAs and, as halo, by the
    // MMAs of both neighbour tiles.  The tiles are issued by different threads, each committing in its
    // own order, so all three commits are waited for.
#ifdef RCED_TC_DIAG_WAIT2   // diagnosis only (unsafe): without the commit of the tile before
    mbar_wait3(bar_addr(c, kBarAccFull + t), bar_addr(c, kBarAccFull + (t + 1 < c.nt ? t + 1 : t - 1)),
               bar_addr(c, kBarAccFull + t), par, c.err, 100 + s);
#else
    mbar_wait3(bar_addr(c, kBarAccFull + t), bar_addr(c, kBarAccFull + (t + 1 < c.nt ? t + 1 : t)),
               bar_addr(c, kBarAccFull + (t > 0 ? t - 1 : t)), par, c.err, 100 + s);
#endif
    fence_after();
    stamp(c.trace, c.tracing && c.quad == 0 && c.lane == 0, s, t, 2);
#if RCED_TC_SKIP_BULK
    // the bulk copies that read this warp's plane rows (the skip tensor a previous step saved) have finished reading
    if (c.lane == 0) bulk_wait_read_all();
    __syncwarp();
#endif

    float d1[8], d2[8];
    tmem_ld8(ta, d1);
    tmem_ld8(ta + e.np, d2);
    const float* bias = c.bias + s * 32;
#pragma unroll 1
    for (int g = 0; g < e.cg; ++g) {
        tmem_wait_ld();
        reg_fence8(d1);
        reg_fence8(d2);
        if (g == 0) stamp(c.trace, c.tracing && c.quad == 0 && c.lane == 0, s, t, 3);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + g * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + g * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#if RCED_TC_SKIP_BULK
        float ss[8];   // hi + 2^-11 lo'
        {
            const uint32_t hh[4] = {__float_as_uint(sk0.x), __float_as_uint(sk0.y), __float_as_uint(sk0.z), __float_as_uint(sk0.w)};
            const uint32_t ll[4] = {__float_as_uint(sk1.x), __float_as_uint(sk1.y), __float_as_uint(sk1.z), __float_as_uint(sk1.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&ll[i]));
                unpack2(fma2(pack2(lf.x, lf.y), loinv2, pack2(hf.x, hf.y)), ss[2 * i], ss[2 * i + 1]);
            }
        }
#else
        const float ss[8] = {sk0.x, sk0.y, sk0.z, sk0.w, sk1.x, sk1.y, sk1.z, sk1.w};
#endif
        float v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // hi*Whi + 2^-11 (hi*Wlo' + lo'*Whi) + bias, two channels per instruction
            u64 x = fma2(pack2(d2[2 * i], d2[2 * i + 1]), loinv2, pack2(d1[2 * i], d1[2 * i + 1]));
            x = fma2(pack2(bb[2 * i], bb[2 * i + 1]), sf2, x);
            if (!AFTER && e.add != 0) x = fma2(pack2(ss[2 * i], ss[2 * i + 1]), r2, x);
            float xa, xb;
            unpack2(x, xa, xb);
            if (e.relu) {
                xa = fmaxf(xa, 0.f);
                xb = fmaxf(xb, 0.f);
            }
            if (AFTER && e.add != 0) {
                xa = fmaf(ss[2 * i], e.skip_r, xa);
                xb = fmaf(ss[2 * i + 1], e.skip_r, xb);
            }
            v[2 * i] = xa;       // (the range guard looks at the FP16 values stored below; halo rows hold
            v[2 * i + 1] = xb;   //  finite sums of their neighbours: harmless)
        }
        // next group: accumulator columns and skip values are in flight while this one is stored
        if (g + 1 < e.cg) {
            tmem_ld8(ta + (g + 1) * 8, d1);
            tmem_ld8(ta + e.np + (g + 1) * 8, d2);
            if (do_add) {
#if RCED_TC_SKIP_BULK
                sk0 = __ldcg(sp + (size_t)(g + 1) * 2 * kRows);
                sk1 = __ldcg(sp + (size_t)(g + 1) * 2 * kRows + kRows);
#else
                sk0 = ld_hint4(sp + (size_t)(g + 1) * 2 * kRows, c.pol_last);
                sk1 = ld_hint4(sp + (size_t)(g + 1) * 2 * kRows + kRows, c.pol_last);
#endif
            }
        }
#if !RCED_TC_SKIP_BULK && !defined(RCED_TC_SAVE_DEFERRED)   // the skip row is stored from the registers, in front of the fence
        if (do_save) {
            st_hint4(dp + (size_t)g * 2 * kRows, make_float4(v[0], v[1], v[2], v[3]), c.pol_last);
            st_hint4(dp + (size_t)g * 2 * kRows + kRows, make_float4(v[4], v[5], v[6], v[7]), c.pol_last);
        }
#endif
        if (e.last) store_split8_parity(c.act, g, kLead + r, v, valid, fi & 1, amax2);
        else store_split8(c.act, g, kLead + r, v, valid, amax2);
    }
    fence_before();       // tcgen05.ld of this accumulator ordered before the barrier
    fence_async_smem();   // plane writes visible to the tensor core (async proxy)
    __syncwarp();
    if (c.lane == 0) mbar_arrive(bar_addr(c, kBarActReady + t));
    stamp(c.trace, c.tracing && c.quad == 0 && c.lane == 0, s, t, 4);
#if RCED_TC_SKIP_BULK
    if (do_save && c.lane == 0) {
        // this warp's 32 rows of every hi and lo plane of the layer: 512 contiguous bytes each
        const int r0 = t * 128 + c.quad * 32;
        const uint32_t src = smem_u32(c.act) + (uint32_t)(kLead + r0) * 16u;
        unsigned char* dst = reinterpret_cast<unsigned char*>(c.skip) + ((size_t)e.save_base * 2 * kRows + r0) * 16;
#pragma unroll 1
        for (int g = 0; g < e.cg; ++g) {
            bulk_s2g(dst + (size_t)g * 2 * kRows * 16, src + (uint32_t)(g * kPlane16) * 16u, 512u);
            bulk_s2g(dst + ((size_t)g * 2 + 1) * kRows * 16, src + (uint32_t)(g * kPlane16 + kLo16) * 16u, 512u);
        }
        bulk_commit();
    }
#elif defined(RCED_TC_SAVE_DEFERRED)
    // Experiment (measured slower: 14.46 against 13.92 ms, the extra instructions cost more than the MEMBAR
    // of the proxy fence waiting for the stores): the skip tensor of the tile written AFTER the tile has
    // been released, its values read back from the planes -- this thread's own row, which nobody overwrites
    // before this same thread does in the next step -- as hi + 2^-11 lo.
    if (do_save) {
        const uint4* ph = reinterpret_cast<const uint4*>(c.act) + (kLead + r);
#pragma unroll 1
        for (int g = 0; g < e.cg; ++g) {
            const uint4 h = ph[g * kPlane16], l = ph[g * kPlane16 + kLo16];
            const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
            float v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&ll[i]));
                v[2 * i] = fmaf(lf.x, kLoInv, hf.x);
                v[2 * i + 1] = fmaf(lf.y, kLoInv, hf.y);
            }
            st_hint4(dp + (size_t)g * 2 * kRows, make_float4(v[0], v[1], v[2], v[3]), c.pol_last);
            st_hint4(dp + (size_t)g * 2 * kRows + kRows, make_float4(v[4], v[5], v[6], v[7]), c.pol_last);
        }
    }
#endif
}

// Epilogue of the (1,129) layer for one row tile.  E[row r][n] (32 columns per frame parity) belongs
// to output row r - n: the warp sums its 32 x 32 block along the diagonals with one shuffle per
// column -- lane m collects the diagonals r - n = r0 + m (columns n < 32 - m, from lanes m + n) and
// r0 + m - 32 (the other columns) -- and stores the two sums: every output row receives exactly one
// sum from the warp that owns its 32-row block (outp[0]) and one from the next block (outp[1]).
// E rows that belong to no frame of the parity, and diagonals that leave the frame, only ever
// reach rows of the other parity or halo rows, which are not stored.
// Batch boundary (RCED_TC_BOUNDARY, default 1): the epilogue of the output layer's row tile t also stages the NEXT
// batch's first-layer input for the rows of tile t -- an in-place update of plane 0 like every conv epilogue (it
// therefore waits for the commits of tiles t-1, t, t+1: their MMAs read tile t's rows) -- so that the boundary
// between two batches is an ordinary layer boundary: the next batch's first layer starts on the early row tiles
// while the output layer still runs on the late ones.  (0: the round-1 form -- all rows staged between a group's
// two output-layer tiles, behind final_done, the first layer behind in_ready of all sixteen warps.)
#ifndef RCED_TC_BOUNDARY
#define RCED_TC_BOUNDARY 1
#endif
struct NextIn {       // what the staging of the next batch needs (prefetched by warp 4)
    const float* ib;  // its kFB + 7 input rows
    const int* tm8;   // per frame: mask of the time taps inside its utterance
    const float* fsc; // per frame: power-of-two scale
    int nf;           // frames of the next batch (0: there is none)
};
template <int ARCH>
__device__ __forceinline__ void epi_final_tile(const Ctx& c, const int t, const uint32_t par, const NextIn& nx, uint32_t& amax2) {
    constexpr int s = num_layers(ARCH) - 1;
#if RCED_TC_BOUNDARY
    if (nx.nf > 0)
        mbar_wait3(bar_addr(c, kBarAccFull + t), bar_addr(c, kBarAccFull + (t + 1 < c.nt ? t + 1 : t)),
                   bar_addr(c, kBarAccFull + (t > 0 ? t - 1 : t)), par, c.err, 300);
    else
#endif
        mbar_wait(bar_addr(c, kBarAccFull + t), par, c.err, 300);
    fence_after();
    stamp(c.trace, c.tracing && c.quad == 0 && c.lane == 0, s, t, 2);
    const int r0 = t * 128 + c.quad * 32;
    const uint32_t ta = c.tm + ((uint32_t)(c.quad * 32) << 16) + (uint32_t)(t * kAccCols);
    // rows this lane stores to: own block / previous block
    const int ra = r0 + c.lane, rb = ra - 32;
    const int fa = ra / kFS, fb = rb >= 0 ? rb / kFS : 0;
    const bool va = fa < c.nf && ra - fa * kFS < kBins;
    const bool vb = rb >= 0 && fb < c.nf && rb - fb * kFS < kBins;
#pragma unroll 1
    for (int odd = 0; odd < 2; ++odd) {
        if (odd && (t == 0 || (c.nt == kTiles && t == kTiles - 1))) break;   // no odd frame reaches the first row tile, nor the last one of a full batch
        float v[kFinalN];
#pragma unroll
        for (int g = 0; g < kFinalN / 8; ++g) {
            float d[8];
            tmem_ld8(ta + odd * kFinalN + g * 8, d);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[g * 8 + e] = d[e];
        }
        tmem_wait_ld();
#pragma unroll
        for (int n = 0; n < kFinalN; ++n) asm volatile("" : "+f"(v[n])::"memory");
        float tot = 0.f, own = 0.f;
#pragma unroll
        for (int n = 0; n < kFinalN; ++n) {
            const float x = __shfl_sync(0xffffffffu, v[n], c.lane + n);   // source lane (m + n) mod 32
            tot += x;
            if (c.lane + n < 32) own += x;
        }
        if (va && (fa & 1) == odd) c.outp[ra] = own;
        if (vb && (fb & 1) == odd) c.outp[kRows + rb] = tot - own;
    }
    if (c.tracing && c.quad == 0 && c.lane == 0) stamp(c.trace, true, s, t, 3);
#if RCED_TC_BOUNDARY
    if (nx.nf > 0) {
        // first-layer input of the next batch, this thread's row: "channel" = time tap, rows g-3 .. g+4 of the utterance
        const int r = t * 128 + c.quad * 32 + c.lane;
        const int fi = r / kFS, b = r - fi * kFS;
        float v[8];
        if (fi < nx.nf && b < kBins) {
            const int m = nx.tm8[fi];
            const float sfi = nx.fsc[fi];   // into the frame's scaled domain (a power of two: exact)
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) v[tt] = (m >> tt) & 1 ? nx.ib[(fi + tt) * kInStride + b] * sfi : 0.f;
        } else {
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) v[tt] = 0.f;
        }
        store_split8(c.act, 0, kLead + r, v, true, amax2);
        fence_async_smem();   // plane writes visible to the tensor core (async proxy)
    }
#endif
    fence_before();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(bar_addr(c, kBarActReady + t));
    stamp(c.trace, c.tracing && c.quad == 0 && c.lane == 0, s, t, 4);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
#ifndef RCED_TC_UNROLL_UNITS
#define RCED_TC_UNROLL_UNITS 2   // units per iteration of the issue loop (experiment switch: 1 and 4 measured no better)
#endif
constexpr int kUnrollUnits = RCED_TC_UNROLL_UNITS;
template <int ARCH>
__global__ void __launch_bounds__(kThreads, 1) rced_net_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NS = n_steps(ARCH);
    constexpr int NL = num_layers(ARCH);
    __shared__ uint32_t s_tmem;
    __shared__ int s_slot, s_slot_owned;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler (uniform datapath)
    unsigned char* act = smem + smem_act_off;
    float* s_bias = reinterpret_cast<float*>(smem + smem_bias_off(ARCH));
    long long* bnd = reinterpret_cast<long long*>(smem + smem_bnd_off(ARCH));
    const uint32_t bars = smem_u32(smem + smem_bar_off(ARCH));
    unsigned int* err = p.flags + 1;

    // ---- one-time setup ----------------------------------------------------------------------
    for (int i = threadIdx.x; i < (kFrontPad + kActBytes) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < bias_floats(ARCH); i += kThreads) s_bias[i] = p.bias[i];
    // frame scales of the batches in flight (ring of kScaleRing batches x 8 frames) and the row maxima of the
    // batch being prefetched, behind the time-tap masks
    float* s_scale = reinterpret_cast<float*>(smem + smem_bnd_off(ARCH) + 64);
    uint32_t* s_rmax = reinterpret_cast<uint32_t*>(smem + smem_bnd_off(ARCH) + 192);
    if (threadIdx.x < kScaleRing * 8) s_scale[threadIdx.x] = 1.f;
    EpiStep* s_epi = reinterpret_cast<EpiStep*>(smem + smem_epi_off(ARCH));
    if (threadIdx.x < NL - 1) {
        const int li = threadIdx.x;
        const LSpec sp = spec(ARCH, li);
        EpiStep e;
        e.np = step_np(ARCH, li);
        e.cg = (sp.cout + 7) / 8;
        e.relu = sp.relu;
        e.add = sp.add < 0 ? 0 : (sp.after ? 2 : 1);
        e.add_base = sp.add < 0 ? 0 : skip_c8_base(ARCH, sp.add);
        e.save_base = sp.save < 0 ? -1 : skip_c8_base(ARCH, sp.save);
        e.last = li == NL - 2 ? 1 : 0;
        e.skip_r = p.bias[NS * 32 + li];
        s_epi[li] = e;
    }
    if (threadIdx.x == 32) {   // (warp 1; warp 0 allocates the tensor memory meanwhile)
        int slot = scratch_slot_acquire(p.slot_busy, p.n_slots);
        s_slot_owned = slot >= 0;
        if (slot < 0) {   // cannot happen in a healthy context: report, the FP32 kernel recomputes the call
            atomicMax(p.flags + 1, 7u);
            slot = (int)(blockIdx.x % (unsigned int)p.n_slots);
        }
        s_slot = slot;
    }
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile uint32_t*>(smem + smem_bar_off(ARCH) + 8 * kFlagSlot) = 0u;
        *reinterpret_cast<volatile uint32_t*>(smem + smem_bar_off(ARCH) + 8 * kNextInSlot) = 0u;
        *reinterpret_cast<volatile uint32_t*>(smem + smem_bar_off(ARCH) + 8 * kIssuedSlot) = 0u;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTiles; ++i) {
            mbar_init(bars + 8 * (kBarAccFull + i), 1);
            mbar_init(bars + 8 * (kBarActReady + i), 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars + 8 * (kBarWFull + i), 1);
            mbar_init(bars + 8 * (kBarWFree + i), kIssuers);   // every issuing thread commits
        }
        mbar_init(bars + 8 * kBarInReady, kEpiWarps);
        mbar_init(bars + 8 * kBarFinalDone, kIssuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    const int FB = p.fb, NT = p.nt;   // frames and row tiles per batch of this launch
    const long long NB = (p.total_rows + FB - 1) / FB;

    if (warp == 0 || warp == 3 || (warp >= 5 && warp < kCtrlWarps)) {
        // ================= MMA issue =================
        // kIssuers issuing threads (one elected lane of warps 0, 3 and 5) take the row tiles of the
        // global (step, tile) sequence round robin: the tensor pipe accepts only a couple of
        // instructions ahead of execution, so the work of one thread between two tiles (dependency
        // check, descriptor set-up, commit) or between two steps (the step prologue) would leave the
        // pipe idle; with three threads and 8 tiles per step the step boundaries of the threads fall
        // at different times, and one thread prepares while the others issue.
        // A unit is a constant-memory load, three uniform adds and the MMA pair (see IssueTab).
        const int iss = warp == 0 ? 0 : (warp == 3 ? 1 : warp - 3);
        if (elect_one()) {
            const uint32_t flag = bars + 8 * kFlagSlot;
            uint32_t seen = 0;   // last value read from the scout's counter
#if RCED_TC_MAXINFLIGHT > 0
            uint32_t issued_seen = 0;   // last value read from the count of tiles whose issue is complete
#endif
            const uint32_t a16_0 = smem_u32(act) >> 4;                                   // plane 0, in 16-byte units
            const uint32_t w16_0 = smem_u32(smem + smem_w_off(ARCH, 0)) >> 4;            // weight buffer 0
            const uint32_t w16_step = (uint32_t)(smem_w_off(ARCH, 1) - smem_w_off(ARCH, 0)) >> 4;
            uint32_t it = 0;
            for (long long batch = blockIdx.x; batch < NB; batch += gridDim.x, ++it) {
                const bool tr = p.trace != nullptr && blockIdx.x == 0 && it == 1;
#pragma unroll 1
                for (int s = 0; s < NS; ++s) {
                    const uint32_t k = it * NS + s;
                    const int wb = k & 1;
                    const int nu = c_issue<ARCH>.nu[s], np = c_issue<ARCH>.np[s];
                    const bool fin = s == NL - 1;
                    const uint32_t id_a = idesc_f16(fin ? np : 2 * np), id_b = idesc_f16((np + 15) & ~15);   // (N of the second instruction: rced_tc.cuh)
                    const int* ta = c_issue<ARCH>.a + s * kTabStride;
                    // B descriptor of unit 0 (unit u is tile16 * u further): steps alternate between the two weight
                    // buffers and n_steps is even, so step s always uses buffer s & 1
                    const uint32_t tile16 = (uint32_t)c_issue<ARCH>.tile16[s];
                    const uint32_t ub0 = ((w16_0 + (uint32_t)(s & 1) * w16_step) & 0x3FFFu) | ((uint32_t)c_issue<ARCH>.rows[s] << 16);
                    const int t_first = (iss + kIssuers - (int)((k * (uint32_t)NT) % kIssuers)) % kIssuers;
#pragma unroll 1
                    for (int t = t_first; t < NT; t += kIssuers) {
                        // the scout (warp 2) has waited on this tile's mbarriers and published its index: a
                        // shared-memory load costs ~30 cycles where an mbarrier test costs ~160 (umma_probe
                        // lat), and it is only needed when the last value seen does not cover this tile
                        stamp(p.trace, tr, s, t, 5);
                        const uint32_t need = k * (uint32_t)NT + t + 1;
                        for (int spin = 0; seen < need; ++spin) {
                            seen = ld_acquire(flag);
                            if ((spin & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) break;
                        }
                        fence_after();
#if RCED_TC_MAXINFLIGHT > 0
                        // At most RCED_TC_MAXINFLIGHT tiles are being issued at a time.  Tiles issued concurrently
                        // share the pipe and finish together; three threads that finish together also prepare
                        // their next tiles together and the pipe idles meanwhile (a stable mode: it persists
                        // once entered).  Holding one thread back -- prepared, dependency cleared -- keeps the
                        // finish times staggered.
                        {
                            const uint32_t gidx = k * (uint32_t)NT + t;   // global tile index
                            if (gidx >= (uint32_t)RCED_TC_MAXINFLIGHT) {
                                const uint32_t want = gidx - (uint32_t)RCED_TC_MAXINFLIGHT + 1u;
                                for (int spin = 0; issued_seen < want; ++spin) {
                                    issued_seen = ld_acquire(bars + 8 * kIssuedSlot);
                                    if ((spin & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) break;
                                }
                            }
                        }
#endif
                        stamp(p.trace, tr, s, t, 0);
                        const uint32_t d = tm + (uint32_t)(t * kAccCols);
                        const uint32_t toff = 128u * t;   // never carries out of the 14-bit start field
                        if (!fin) {
#pragma unroll kUnrollUnits
                            for (int u = 0; u < nu; ++u) {
                                const uint32_t ua = a16_0 + (uint32_t)ta[u] + toff;
                                const uint64_t db = make_desc(ub0 + (uint32_t)u * tile16);
                                umma_f16(d, make_desc(ua), db, id_a, u > 0);            // hi x [Whi | Wlo'] -> columns [0, 2 NP)
                                umma_f16(d + np, make_desc(ua + kLo16), db, id_b, 1);   // lo' x Whi         -> columns [NP, 2 NP): both carry 2^11
                                if (u == 0) stamp(p.trace, tr, s, t, 6);
                            }
                        } else {
                            // output layer: per frame parity (even-frame copy in planes 0-1, odd-frame copy two planes
                            // further, 32 accumulator columns each) five row-shifted blocks of three products
#pragma unroll 1
                            for (int odd = 0; odd < 2; ++odd) {
                                if (odd && (t == 0 || (NT == kTiles && t == kTiles - 1))) break;   // no odd frame reaches these row tiles
                                const uint32_t dd = d + (uint32_t)(odd * kFinalN);
                                const uint32_t po = toff + (uint32_t)(odd * 2 * kPlane16);
#pragma unroll
                                for (int u = 0; u < kFinalShifts; ++u) {
                                    const uint64_t dbh = make_desc(ub0 + (uint32_t)u * tile16);                      // Whi        rows 0..31
                                    const uint64_t dbl = make_desc(ub0 + (uint32_t)u * tile16 + (uint32_t)np);       // Wlo        rows 32..63
                                    const uint64_t dbs = make_desc(ub0 + (uint32_t)u * tile16 + 2u * (uint32_t)np);  // Whi 2^-11  rows 64..95
                                    const uint32_t ua = a16_0 + (uint32_t)ta[u] + po;
                                    umma_f16(dd, make_desc(ua), dbh, id_b, u > 0);
                                    umma_f16(dd, make_desc(ua + kLo16), dbs, id_b, 1);   // the activations' residual carries 2^11
                                    umma_f16(dd, make_desc(ua), dbl, id_b, 1);
                                }
                            }
                        }
                        stamp(p.trace, tr, s, t, 7);
                        umma_commit(bars + 8 * (kBarAccFull + t));
#if RCED_TC_MAXINFLIGHT > 0
                        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(bars + 8 * kIssuedSlot) : "memory");
#endif
                        stamp(p.trace, tr, s, t, 1);
                    }
                    umma_commit(bars + 8 * (kBarWFree + wb));
                    if (s == NL - 1) umma_commit(bars + 8 * kBarFinalDone);   // the planes are free for the next batch's input
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= weight producer =================
        // streams every step's weight tiles from L2 into the double buffer, two steps ahead
        uint32_t it = 0;
        for (long long batch = blockIdx.x; batch < NB; batch += gridDim.x, ++it) {
#pragma unroll 1
            for (int s = 0; s < NS; ++s) {
                const uint32_t k = it * NS + s;
                const int wb = k & 1;
                if (k >= 2) mbar_wait(bars + 8 * (kBarWFree + wb), ((k >> 1) - 1) & 1, err, 5);
                if (lane == 0) {
                    const uint32_t bytes = (uint32_t)step_w_bytes(ARCH, s);
                    const uint32_t full = bars + 8 * (kBarWFull + wb);
                    mbar_expect_tx(full, bytes);
                    const unsigned char* src = p.wimg + step_w_off(ARCH, s);
                    const uint32_t dst = smem_u32(smem + smem_w_off(ARCH, wb));
                    for (uint32_t o = 0; o < bytes; o += 16384u) bulk_g2s(dst + o, src + o, bytes - o < 16384u ? bytes - o : 16384u, full);
                }
                __syncwarp();
            }
        }
    } else if (warp == 4) {
        // ================= input prefetch =================
        // While a batch runs, fetches what the NEXT one needs at its start: the utterance bounds
        // of its frames (a binary search over row_off: ten dependent L2 round trips) and its
        // kFB + 7 input rows, into the half of the double buffer the batch before last used.
        float* inbuf = reinterpret_cast<float*>(smem + smem_in_off(ARCH));
        const uint32_t flag = bars + 8 * kFlagSlot;
        uint32_t itn = 0;   // local index of the batch being prefetched
        const uint32_t bref = __float_as_uint(s_bias[NS * 32 + kBiasRefSlot]);   // largest |bias| (non-negative: orders like its bits)
        bool bad_input = false;
        for (long long batch = blockIdx.x; batch < NB; batch += gridDim.x, ++itn) {
            if (itn >= 2) {
                // buffer itn & 1 was read by the staging of batch itn - 2: wait until the scout has cleared every row
                // tile of that batch's first layer (the staging of its last tile has then been released; round-1 form:
                // its counter has passed the input barrier, i.e. the batch's first tile)
                const uint32_t need = (itn - 2) * NS * (uint32_t)NT + (RCED_TC_BOUNDARY ? (uint32_t)NT : 1u);
                for (int spin = 0; ld_acquire(flag) < need; ++spin) {
                    __nanosleep(200);
                    if ((spin & 255) == 255 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) break;
                }
            }
            const long long g0 = batch * FB;
            // per frame of the batch: which of its 8 time taps (rows g-3 .. g+4) lie inside its utterance
            int* tm8 = reinterpret_cast<int*>(bnd) + (itn & 1) * 8;
            int m = 0;
            if (lane < FB) {
                long long lo = 0, hi = 0;
                const long long g = g0 + lane;
                if (g < p.total_rows) locate(p.row_off, p.n_utt, g, lo, hi);
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) m |= (g + tt - 3 >= lo && g + tt - 3 < hi) ? (1 << tt) : 0;
                tm8[lane] = m;
            }
            float* ib = inbuf + (itn & 1) * kInRows * kInStride;
            for (int j = 0; j < FB + 7; ++j) {
                const long long src = g0 - 3 + j;
                const bool ok = src >= 0 && src < p.total_rows;
                uint32_t mb = 0;   // largest |x| of the row as a bit pattern (NaN orders above infinity)
                for (int b = lane; b < kBins; b += 32) {
                    const float x = ok ? __ldg(p.in + src * kBins + b) : 0.f;
                    mb = max(mb, __float_as_uint(x) & 0x7fffffffu);
                    ib[j * kInStride + b] = x;
                }
                mb = __reduce_max_sync(0xffffffffu, mb);
                if (lane == 0) s_rmax[j] = mb;
            }
            __syncwarp();
            // The frame's scale: a power of two that puts max(|input rows the frame reads|, largest |bias|) into
            // [2^(kFrameTop-1), 2^kFrameTop).  Frames without input and a bias-free model: 1.
            if (lane < 8) {
                uint32_t fm = 0;
#pragma unroll
                for (int tt = 0; tt < 8; ++tt)
                    if ((m >> tt) & 1) fm = max(fm, s_rmax[lane + tt]);
                float sc = 1.f;
                if (fm >= 0x7f800000u) {
                    bad_input = true;   // infinity or NaN: the FP32 kernel recomputes the call
                } else {
                    const uint32_t ref = max(fm, bref);
                    if (ref != 0u) {
                        int k = 127 + kFrameTop - 1 - (int)(ref >> 23);
                        if (k < -100) { k = -100; bad_input = true; }   // beyond 2^103: out of any sensible range
                        if (k > 100) k = 100;
                        sc = __uint_as_float((uint32_t)(127 + k) << 23);
                    }
                }
                s_scale[(itn & (kScaleRing - 1)) * 8 + lane] = sc;
            }
            __syncwarp();
            // a counter, not an mbarrier: this warp may be two batches ahead of the epilogue warps,
            // which a phase parity could not tell apart
            if (lane == 0) st_release(bars + 8 * kNextInSlot, itn + 1);
        }
        if (bad_input) atomicMax(p.flags, 0x7f800000u);
    } else if (warp == 2) {
        // ================= dependency scout =================
        // Waits, in the MMA thread's issue order, on the mbarriers every (step, tile) depends on and
        // publishes the number of tiles cleared for issue.
        if (elect_one()) {
            const uint32_t flag = bars + 8 * kFlagSlot;
            uint32_t done = 0;
            uint32_t it = 0;
            for (long long batch = blockIdx.x; batch < NB; batch += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < NS; ++s) {
                    const uint32_t k = it * NS + s;
                    const int wb = k & 1;
#pragma unroll 1
                    for (int t = 0; t < NT; ++t) {
                        if (t == 0) {
                            // first layer: its input is staged (the first batch: by all epilogue warps up front; later
                            // batches with RCED_TC_BOUNDARY: by the output layer's epilogues, covered by act_ready below)
                            if (s == 0 && (!RCED_TC_BOUNDARY || it == 0)) mbar_wait(bars + 8 * kBarInReady, it & 1, err, 1);
                            mbar_wait(bars + 8 * (kBarWFull + wb), (k >> 1) & 1, err, 2);
                            if (k > 0) mbar_wait(bars + 8 * (kBarActReady + 0), (k - 1) & 1, err, 3);
                        }
                        // (step 0 of a batch: the output layer's epilogues of the batch before have read the accumulators)
                        if (k > 0 && t + 1 < NT) mbar_wait(bars + 8 * (kBarActReady + t + 1), (k - 1) & 1, err, 4);
                        st_release(flag, ++done);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kCtrlWarps) {
        // ================= epilogue groups =================
        Ctx c;
        c.smem = smem;
        c.act = act;
        c.bars = bars;
        c.tm = tm;
        c.bias = s_bias;
        c.epi = s_epi;
        c.bnd = bnd;
        c.err = err;
        c.trace = p.trace;
        c.tracing = false;
        c.lane = lane;
        c.quad = warp & 3;
        c.grp = (warp - kCtrlWarps) >> 2;   // four consecutive warps cover the four lane quadrants
        c.nt = NT;
        c.et = (warp - kCtrlWarps) * 32 + lane;
        c.skip = p.skip + (size_t)s_slot * skip_floats_per_cta(ARCH);
        c.outp = reinterpret_cast<float*>(smem + smem_out_off(ARCH));
        c.pol_last = l2_policy_evict_last();
        c.pol_first = l2_policy_evict_first();
        uint32_t amax2 = 0u;   // packed FP16 pair: largest |activation| stored
        const float bias_f = s_bias[(NL - 1) * 32];
        const float c1_f = s_bias[NS * 32 + NL - 1];   // inverse weight scale of the output layer
        // Stages the first layer's input of batch sb (local index sit): "channel" = time tap, rows g-3 .. g+4
        // of the utterance, from the block the prefetch warp has loaded.  Plane 0 must be free.
        auto stage_input = [&](const long long sb, const uint32_t sit) {
            const long long sg0 = sb * FB;
            const long long left = p.total_rows - sg0;
            const int snf = left < FB ? (int)left : FB;
            for (int spin = 0; ld_acquire(bars + 8 * kNextInSlot) <= sit; ++spin)
                if ((spin & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) break;
            const int* tm8 = reinterpret_cast<const int*>(bnd) + (sit & 1) * 8;
            const float* ib = reinterpret_cast<const float*>(smem + smem_in_off(ARCH)) + (sit & 1) * kInRows * kInStride;
            const float* fsc = s_scale + (sit & (kScaleRing - 1)) * 8;
#pragma unroll 1
            for (int r = c.et; r < NT * 128; r += 32 * kEpiWarps) {
                const int fi = r / kFS, b = r - fi * kFS;
                float v[8];
                if (fi < snf && b < kBins) {
                    const int m = tm8[fi];
                    const float sfi = fsc[fi];   // into the frame's scaled domain (a power of two: exact)
#pragma unroll
                    for (int tt = 0; tt < 8; ++tt) v[tt] = (m >> tt) & 1 ? ib[(fi + tt) * kInStride + b] * sfi : 0.f;   // input row fi + tt of the block
                } else {
#pragma unroll
                    for (int tt = 0; tt < 8; ++tt) v[tt] = 0.f;
                }
                store_split8(act, 0, kLead + r, v, true, amax2);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * kBarInReady);
        };
        if ((long long)blockIdx.x < NB) stage_input(blockIdx.x, 0);
        uint32_t it = 0;
        for (long long batch = blockIdx.x; batch < NB; batch += gridDim.x, ++it) {
            const long long g0 = batch * FB;
            const long long left = p.total_rows - g0;
            c.nf = left < FB ? (int)left : FB;
            c.g0 = g0;
            c.scale = s_scale + (it & (kScaleRing - 1)) * 8;
            c.tracing = p.trace != nullptr && blockIdx.x == 0 && it == 1;
            const uint32_t k0 = it * NS;
#pragma unroll 1
            for (int s = 0; s < NL - 1; ++s) {
                const uint32_t par = (k0 + s) & 1;
#pragma unroll 1
                for (int t = c.grp; t < NT; t += kGroups) epi_conv_tile<ARCH == 3>(c, s, t, par, amax2);
            }
            // Output layer.  Between this warp's two row tiles the next batch's input is staged: plane 0 is
            // free as soon as the output layer's MMAs have completed, so the first layer of the next
            // batch starts on the row tiles whose accumulators have been read while the epilogues of the
            // later ones still run (the scout holds every tile back until its accumulator is free).
            {
                const uint32_t par = (k0 + NL - 1) & 1;
                static_assert(kTiles == 2 * kGroups, "two row tiles per epilogue group");
                NextIn nx{nullptr, nullptr, nullptr, 0};
#if RCED_TC_BOUNDARY
                if (batch + gridDim.x < NB) {
                    const uint32_t sit = it + 1;
                    const long long left_n = p.total_rows - (batch + gridDim.x) * FB;
                    for (int spin = 0; ld_acquire(bars + 8 * kNextInSlot) <= sit; ++spin)   // the prefetch warp has the batch
                        if ((spin & 1023) == 1023 && *reinterpret_cast<volatile unsigned int*>(err) != 0u) break;
                    nx.ib = reinterpret_cast<const float*>(smem + smem_in_off(ARCH)) + (sit & 1) * kInRows * kInStride;
                    nx.tm8 = reinterpret_cast<const int*>(bnd) + (sit & 1) * 8;
                    nx.fsc = s_scale + (sit & (kScaleRing - 1)) * 8;
                    nx.nf = left_n < FB ? (int)left_n : FB;
                }
                if (c.grp < NT) epi_final_tile<ARCH>(c, c.grp, par, nx, amax2);
#else
                if (c.grp < NT) epi_final_tile<ARCH>(c, c.grp, par, nx, amax2);
                if (batch + gridDim.x < NB) {
                    mbar_wait(bars + 8 * kBarFinalDone, it & 1, err, 6);
                    stage_input(batch + gridDim.x, it + 1);
                }
#endif
                if (c.grp + kGroups < NT) epi_final_tile<ARCH>(c, c.grp + kGroups, par, nx, amax2);
            }
            epi_bar();   // both partial sums of every output row are stored
            for (int i = c.et; i < c.nf * kBins; i += 32 * kEpiWarps) {
                const int fi = i / kBins, b = i - fi * kBins;
                const int row = fi * kFS + b;
                // back from the frame's scaled domain: 1 / s of a power of two by exponent arithmetic
                const float inv = c1_f * __uint_as_float(0x7f000000u - __float_as_uint(c.scale[fi]));
                st_hint1(p.out + (g0 + fi) * kBins + b, fmaf(c.outp[row] + c.outp[kRows + row], inv, bias_f), c.pol_first);
            }
        }
#if RCED_TC_SKIP_BULK
        if (lane == 0) bulk_wait_all();   // no bulk copy may still read this CTA's shared memory when it exits
#endif
        // range guard: non-negative floats order like their bit patterns (an FP16 overflow is an infinity)
        const float2 am = __half22float2(*reinterpret_cast<const __half2*>(&amax2));
        uint32_t amax = max(__float_as_uint(am.x), __float_as_uint(am.y));
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0) atomicMax(p.flags, amax);
    }

    fence_before();
    __syncthreads();
    if (threadIdx.x == 32 && s_slot_owned) scratch_slot_release(p.slot_busy, s_slot);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// ------------------------------------------------------------------------------------------
// host side: weight image, launch
// ------------------------------------------------------------------------------------------
// w = hi + lo as FP16 numbers; lo_scaled = 2^11 lo and hi_down = 2^-11 hi are the forms that meet the
// scaled residual of the activations (rced_tc.cuh, "Range")
static inline void split_half(float w, uint16_t& hi, uint16_t& lo, uint16_t& lo_scaled, uint16_t& hi_down) {
    const __half h = __float2half_rn(w);
    const float r = w - __half2float(h);
    const __half l = __float2half_rn(r);
    const __half ls = __float2half_rn(ldexpf(r, kLoShift));
    const __half hd = __float2half_rn(ldexpf(__half2float(h), -kLoShift));
    memcpy(&hi, &h, 2);
    memcpy(&lo, &l, 2);
    memcpy(&lo_scaled, &ls, 2);
    memcpy(&hi_down, &hd, 2);
}

}  // namespace tc

int tc_image_bytes(int arch) { return tc::w_image_bytes(arch); }
int tc_bias_floats(int arch) { return tc::bias_floats(arch); }
size_t tc_skip_floats_per_cta(int arch) { return tc::skip_floats_per_cta(arch); }
int tc_smem_bytes(int arch) { return tc::smem_total(arch); }

// folded: canonical BN-folded weights (rced_folded_weight_count); img: tc_image_bytes; bias: tc_bias_floats
void tc_pack_weights(int arch, const float* folded, unsigned char* img, float* bias) {
    using namespace tc;
    const int nl = num_layers(arch), ns = n_steps(arch);
    memset(img, 0, (size_t)w_image_bytes(arch));
    for (int i = 0; i < bias_floats(arch); ++i) bias[i] = 0.f;
    float bias_ref = 0.f;
    // Power-of-two weight scales (rced_tc.cuh, "Range").  kw[s]: exponent the weights of step s are scaled by;
    // dom[s]: exponent of the domain the step's OUTPUT lives in (on top of the frame's scale) = sum of kw up to s.
    // Conv steps keep kw = 0 while their largest weight lies in [2^-7, 2^3) -- their FP16 image is then as precise as
    // it gets and the epilogue needs no multiplication; layers outside (BN folds with extreme gamma / variance) are
    // brought to [2^-2, 2^-1) and the domain of everything behind them shifts with them.  The output layer is always
    // scaled to [2^12, 2^13): its 2^-11 copy of Whi must stay a normal FP16 number.
    int kw[kMaxLayers] = {}, dom[kMaxLayers] = {};
    for (int s = 0; s < ns; ++s) {
        const int li = step_layer(arch, s);
        const LSpec sp = spec(arch, li);
        const float* k = folded + folded_off(arch, li);
        float wmax = 0.f;
        for (size_t i = 0; i < (size_t)sp.kh * sp.kw * sp.cin * sp.cout; ++i) wmax = fmaxf(wmax, fabsf(k[i]));
        int ex = 0;
        if (wmax > 0.f) frexpf(wmax, &ex);   // wmax = m 2^ex, m in [0.5, 1)
        if (wmax > 0.f) {
            if (is_final(arch, s)) kw[s] = kStepWeightTop - ex;
            else if (ex < -6 || ex > 3) kw[s] = -1 - ex;
        }
        if (kw[s] > 100) kw[s] = 100;
        if (kw[s] < -100) kw[s] = -100;
        dom[s] = (s > 0 ? dom[s - 1] : 0) + kw[s];
    }
    uint16_t* im = reinterpret_cast<uint16_t*>(img);
    for (int s = 0; s < ns; ++s) {
        const int li = step_layer(arch, s);
        const LSpec sp = spec(arch, li);
        const float* k = folded + folded_off(arch, li);   // [kh][kw][cin][cout]
        const float* b = k + (size_t)sp.kh * sp.kw * sp.cin * sp.cout;
        const int rows = step_tile_rows(arch, s), np = step_np(arch, s), nc = step_chunks(arch, s);
        uint16_t* base = im + step_w_off(arch, s) / 2;
        const int kwe = kw[s];
        if (is_final(arch, s)) {
            bias[ns * 32 + s] = ldexpf(1.f, -dom[s]);   // accumulator -> true output, before the division by the frame's scale
        } else {
            // the skip tensor a step adds was saved in the domain of the step that produced it
            float r = 1.f;
            if (sp.add >= 0)
                for (int j = 0; j < s; ++j)
                    if (spec(arch, j).save == sp.add) r = ldexpf(1.f, dom[s] - dom[j]);
            bias[ns * 32 + s] = r;
        }
        for (int u = 0; u < step_units(arch, s); ++u)
            for (int cc = 0; cc < 2; ++cc) {
                const int ch = is_final(arch, s) ? cc : 2 * u + cc;
                if (ch >= nc) continue;
                for (int n = 0; n < rows; ++n)
                    for (int e = 0; e < 8; ++e) {
                        float w = 0.f;
                        const int blk = n / np;   // 0: Whi, 1: residual (conv steps: times 2^11), 2 (output layer): Whi 2^-11
                        const int nn = n - blk * np;
                        if (is_final(arch, s)) {
                            // unit u = block of 32 taps: tap j = 32 u + nn, chunk = channel group
                            const int tap = u * kFinalN + nn, ci = 8 * cc + e;
                            if (tap < sp.kw && ci < sp.cin) w = k[((size_t)tap * sp.cin + ci) * sp.cout];
                        } else {
                            const int g = ch / sp.kw, j = ch % sp.kw, ci = 8 * g + e;
                            if (nn < sp.cout && ci < cin_eff(arch, s)) {
                                // first layer: "channel" ci is the time tap (kh index), cin == 1
                                w = s == 0 ? k[(((size_t)ci * sp.kw + j) * sp.cin + 0) * sp.cout + nn]
                                           : k[(((size_t)0 * sp.kw + j) * sp.cin + ci) * sp.cout + nn];
                            }
                        }
                        uint16_t hi, lo, lo_scaled, hi_down;
                        split_half(ldexpf(w, kwe), hi, lo, lo_scaled, hi_down);
                        base[((size_t)u * 2 * rows + (size_t)cc * rows + n) * 8 + e] =
                            blk == 0 ? hi : (blk == 2 ? hi_down : (is_final(arch, s) ? lo : lo_scaled));
                    }
            }
        // biases enter the domain of their step; the output layer's bias is added outside the scaled domain
        if (is_final(arch, s)) bias[s * 32] = b[0];
        else
            for (int o = 0; o < sp.cout; ++o) {
                bias[s * 32 + o] = ldexpf(b[o], dom[s]);
                bias_ref = fmaxf(bias_ref, fabsf(bias[s * 32 + o]));
            }
    }
    bias[ns * 32 + kBiasRefSlot] = bias_ref;
}

template <int ARCH>
static cudaError_t launch_tc_t(const tc::TcParams& p, int num_sms, size_t persist_bytes, cudaStream_t stream) {
    constexpr int smem = tc::smem_total(ARCH);
    static_assert(smem <= 227 * 1024, "activation planes + weight double buffer must fit one SM's shared memory");
    static bool attr_set[64] = {};   // per instantiation and device; the attribute is sticky
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tc::rced_net_tc_kernel<ARCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    long long ctas = (p.total_rows + p.fb - 1) / p.fb;
    if (ctas > num_sms) ctas = num_sms;
    if (ctas < 1) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (persist_bytes > 0) {
        // the skip scratch is re-written by every batch and read back a few layers later: keep it in the
        // part of the L2 set aside for persisting lines (cudaLimitPersistingL2CacheSize, set by the caller)
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = p.skip;
        attr[0].val.accessPolicyWindow.num_bytes = persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, tc::rced_net_tc_kernel<ARCH>, p);
}

int tc_trace_slots(int arch) { return RCED_TC_TRACING ? tc::n_steps(arch) * tc::kTiles * tc::kTraceEvents : 0; }   // 0: trace compiled out

cudaError_t launch_net_tc(int arch, const NetParams& np, const unsigned char* wimg, const float* bias, float* skip,
                          unsigned int* slot_busy, int n_slots, size_t persist_bytes, unsigned int* flags, long long* trace,
                          int num_sms, cudaStream_t stream) {
    tc::TcParams p;
    p.wimg = wimg;
    p.bias = bias;
    p.in = np.in;
    p.out = np.out;
    p.row_off = np.row_off;
    p.n_utt = np.n_utt;
    p.total_rows = np.total_rows;
    p.skip = skip;
    p.slot_busy = slot_busy;
    p.n_slots = n_slots;
    p.flags = flags;
    p.trace = trace;
    // launch shape: full batches of kFB frames when every SM gets at least two of them; smaller launches spread their
    // frames over the SMs (fb = ceil(rows / SMs)), RCED_TC_FB overrides (experiments)
    static const int fb_env = getenv("RCED_TC_FB") ? atoi(getenv("RCED_TC_FB")) : 0;
    int fb = tc::kFB;
    if (np.total_rows < 2LL * tc::kFB * num_sms) {
        fb = (int)((np.total_rows + num_sms - 1) / num_sms);
        if (fb < 1) fb = 1;
        if (fb > tc::kFB) fb = tc::kFB;
    }
    if (fb_env >= 1 && fb_env <= tc::kFB) fb = fb_env;
    p.fb = fb;
    p.nt = (fb * tc::kFS + 127) / 128;
    switch (arch) {
        case 1: return launch_tc_t<1>(p, num_sms, persist_bytes, stream);
        case 2: return launch_tc_t<2>(p, num_sms, persist_bytes, stream);
        case 3: return launch_tc_t<3>(p, num_sms, persist_bytes, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace rced
