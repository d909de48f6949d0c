// Warp-level 128-point complex FFT shared by the STFT (K1) and iSTFT (K3) kernels.
//
// One warp transforms one 128-point sequence held 4 values per lane: a radix-4 butterfly in
// registers, a twiddle, then a 32-point decimation-in-frequency FFT across the lanes made of
// five shfl_xor stages.  tools/fft_emulator.py is the numpy lane-by-lane model of this file.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rced {

// tables filled by upload_tables() (rced_tables.cu), copied to shared memory by the kernels
struct FftTables {
    float2 tw256[256];     // e^{-2 pi i j / 256}
    float2 tw512[132];     // e^{+ pi i k / 256}, k = 0..128 (pre-twiddle of the odd output samples)
    float ham[256];        // np.hamming(256)
    float inv_ham[256];    // 1 / np.hamming(256)
};
// one copy per translation unit (no relocatable device code needed); each TU that uses the
// tables exports an upload function built on upload_tables_local()
static __device__ FftTables g_tables;

static inline cudaError_t upload_tables_local() {
    static FftTables host;   // ~5 KB, filled in double precision
    const double pi = 3.14159265358979323846;
    for (int j = 0; j < 256; ++j) {
        host.tw256[j] = make_float2((float)cos(2.0 * pi * j / 256.0), (float)(-sin(2.0 * pi * j / 256.0)));
        const double w = 0.54 - 0.46 * cos(2.0 * pi * j / 255.0);   // np.hamming(256), symmetric
        host.ham[j] = (float)w;
        host.inv_ham[j] = (float)(1.0 / w);
    }
    for (int k = 0; k < 132; ++k) {
        const int kk = k > 128 ? 128 : k;
        host.tw512[k] = make_float2((float)cos(pi * kk / 256.0), (float)sin(pi * kk / 256.0));
    }
    return cudaMemcpyToSymbol(g_tables, &host, sizeof(FftTables));
}

// complex add / subtract as ONE packed FP32 instruction (Blackwell FADD2): both kernels are bound by their issue slots
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long pa, pb, pd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
    return d;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long pa, pb, pd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
    return d;
}
// The FFT leaves lane l with outputs 4 bitrev5(l) + b: stored as they are, the 32 lanes of a store hit 4 banks.
// Two padding slots per 16 entries spread them over all banks, for the 16-byte stores of the transform's output as
// well as for the consecutive 8-byte loads that follow.
__device__ __forceinline__ int zpad(int e) { return e + 2 * (e >> 4); }
constexpr int kZPad = 128 + 2 * 8;   // entries of a padded 128-point buffer
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 shfl_xor2(float2 a, int m) {
    return make_float2(__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m));
}

// The eight twiddles of a transform depend on the lane only.  Read from tw256 itself their indices are strided
// (2 b lane, (lane & (s - 1)) 128 / s): up to 8 lanes of a half-warp on one bank pair, 62 shared-memory wavefronts
// per transform where 16 would do.  Each CTA therefore keeps a copy ordered [twiddle][lane] - consecutive lanes,
// consecutive 8-byte entries: rows 0..2 the twiddles after the radix-4 step (b = 1..3), rows 3..7 the stages
// s = 16..1.
constexpr int kLaneTw = 8 * 32;
__device__ __forceinline__ int lane_tw_index(int row, int lane) {
    if (row < 3) return 2 * lane * (row + 1);   // <= 186
    const int s = 16 >> (row - 3);
    return (lane & (s - 1)) * (128 / s);
}
// Rows 3..7 hold the twiddle only for the lanes that multiply by it (lane & s); the others hold 1, so that a stage is
// the same instructions on every lane: t = o -+ v, v = t * w.
__device__ __forceinline__ void fill_lane_tw(float2* lane_tw, const float2* tw256) {   // before a __syncthreads()
    for (int i = threadIdx.x; i < kLaneTw; i += blockDim.x) {
        const int row = i >> 5, lane = i & 31;
        const bool one = row >= 3 && (lane & (16 >> (row - 3))) == 0;
        lane_tw[i] = one ? make_float2(1.f, 0.f) : tw256[lane_tw_index(row, lane)];
    }
}
// a * (s, s) + b as one packed FP32 instruction (FFMA2)
__device__ __forceinline__ float2 cfma_s(float2 a, float s, float2 b) {
    unsigned long long pa, ps, pb, pd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pd) : "l"(pa), "l"(ps), "l"(pb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
    return d;
}
// HBM -> L2 prefetch of the line holding p (no register, no scoreboard)
__device__ __forceinline__ void prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// In : v[a] = z[lane + 32 a]                       (natural order)
// Out: v[b] = Z[4 * bitrev5(lane) + b]             (INV: unscaled inverse transform)
template <bool INV>
__device__ __forceinline__ void fft128_warp(float2 (&v)[4], const int lane, const float2* __restrict__ lane_tw) {
    const float2 s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]);
    const float2 s2 = cadd(v[1], v[3]), s3 = csub(v[1], v[3]);
    // forward: -i*s3 = (s3.y, -s3.x); inverse: +i*s3 = (-s3.y, s3.x)
    const float2 r3 = INV ? make_float2(-s3.y, s3.x) : make_float2(s3.y, -s3.x);
    v[0] = cadd(s0, s2);
    v[1] = cadd(s1, r3);
    v[2] = csub(s0, s2);
    v[3] = csub(s1, r3);
#pragma unroll
    for (int b = 1; b < 4; ++b) {
        float2 w = lane_tw[32 * (b - 1) + lane];
        if (INV) w.y = -w.y;
        v[b] = cmul(v[b], w);
    }
    int row = 3;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1, ++row) {
        // lanes with bit s: (o - v) * w; the others: (o + v) * 1 - exactly o + v.  The last stage's twiddle is 1 on
        // every lane.
        const float sg = (lane & s) != 0 ? -1.f : 1.f;
        float2 w = lane_tw[32 * row + lane];
        if (INV) w.y = -w.y;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float2 t = cfma_s(v[b], sg, shfl_xor2(v[b], s));
            v[b] = s > 1 ? cmul(t, w) : t;
        }
    }
}

// ---- two transforms at once ---------------------------------------------------------------------------------------
// K3 (irfft_n = 512) inverse-transforms two sequences per segment with the same twiddles.  Held as packed pairs --
// x[b] = (Re a[b], Re b[b]), y[b] = (Im a[b], Im b[b]) -- every arithmetic step of the two transforms is ONE packed
// instruction with the lane's twiddle as broadcast scalar (FFMA2 / FMUL2 with an .F32 operand): 10 instead of 14
// instructions per butterfly pair, the shuffles stay.  Same operations in the same order as fft128_warp: same bits.
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_pack(float lo, float hi) { pk2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void pk_unpack(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) { pk2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b) { pk2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) { pk2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) { pk2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2 pk_shfl_xor(pk2 v, int m) {
    float lo, hi;
    pk_unpack(v, lo, hi);
    return pk_pack(__shfl_xor_sync(0xffffffffu, lo, m), __shfl_xor_sync(0xffffffffu, hi, m));
}
// (x, y) *= (wx + i wy) for both transforms: x' = x wx - y wy, y' = x wy + y wx, rounded like cmul()
__device__ __forceinline__ void pk_cmul(pk2& x, pk2& y, float wx, float wy) {
    const pk2 bx = pk_pack(wx, wx), by = pk_pack(wy, wy), nby = pk_pack(-wy, -wy);
    const pk2 nx = pk_fma(x, bx, pk_mul(y, nby));
    y = pk_fma(x, by, pk_mul(y, bx));
    x = nx;
}
template <bool INV>
__device__ __forceinline__ void fft128_warp_pair(pk2 (&x)[4], pk2 (&y)[4], const int lane, const float2* __restrict__ lane_tw) {
    const pk2 s0x = pk_add(x[0], x[2]), s0y = pk_add(y[0], y[2]), s1x = pk_sub(x[0], x[2]), s1y = pk_sub(y[0], y[2]);
    const pk2 s2x = pk_add(x[1], x[3]), s2y = pk_add(y[1], y[3]), s3x = pk_sub(x[1], x[3]), s3y = pk_sub(y[1], y[3]);
    x[0] = pk_add(s0x, s2x); y[0] = pk_add(s0y, s2y);
    x[2] = pk_sub(s0x, s2x); y[2] = pk_sub(s0y, s2y);
    if (INV) {   // +i s3 = (-s3.y, s3.x)
        x[1] = pk_sub(s1x, s3y); y[1] = pk_add(s1y, s3x);
        x[3] = pk_add(s1x, s3y); y[3] = pk_sub(s1y, s3x);
    } else {     // -i s3 = (s3.y, -s3.x)
        x[1] = pk_add(s1x, s3y); y[1] = pk_sub(s1y, s3x);
        x[3] = pk_sub(s1x, s3y); y[3] = pk_add(s1y, s3x);
    }
#pragma unroll
    for (int b = 1; b < 4; ++b) {
        const float2 w = lane_tw[32 * (b - 1) + lane];
        pk_cmul(x[b], y[b], w.x, INV ? -w.y : w.y);
    }
    int row = 3;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1, ++row) {
        const float sg = (lane & s) != 0 ? -1.f : 1.f;
        const pk2 sg2 = pk_pack(sg, sg);
        const float2 w = lane_tw[32 * row + lane];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            x[b] = pk_fma(x[b], sg2, pk_shfl_xor(x[b], s));
            y[b] = pk_fma(y[b], sg2, pk_shfl_xor(y[b], s));
            if (s > 1) pk_cmul(x[b], y[b], w.x, INV ? -w.y : w.y);
        }
    }
}

__device__ __forceinline__ int bitrev5(int lane) { return (int)(__brev((unsigned)lane) >> 27); }

}  // namespace rced
