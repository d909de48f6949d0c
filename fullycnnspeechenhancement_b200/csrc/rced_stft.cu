// K1: pre-emphasis + framing + Hamming window + 256-point real FFT + magnitude + unit phase.
//
// Replaces AudioFeature.compute_spectrogram / power_spectrum / divide_phase
// (data_utils/audio_feature.py:22-115) and the zero padding of DataLoader.padding_batch
// (data_utils/data_loader.py:198-209).  One warp per spectrogram row; a CTA of 8 warps walks
// kRowsPerCta consecutive rows so the 5 KB of tables are staged in shared memory once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rced_fft.cuh"
#include "rced_internal.h"

namespace rced {

constexpr int kStftWarps = 8;
constexpr int kRowsPerCta = 64;        // rows a CTA walks in a large launch (tables staged once per 64 rows)
constexpr int kRowsPerCtaSmall = 8;    // small launches (one utterance, a streaming block): one row per warp, more CTAs

__device__ __forceinline__ long long frames_of(long long L) {   // audio_feature.py:67-70
    const long long d = L >= 256 ? L - 256 : 256 - L;
    return (d + 127) / 128 + 1;
}

__global__ void __launch_bounds__(kStftWarps * 32, 4) rced_stft_kernel(const StftParams p, const int rows_per_cta) {
    __shared__ float2 s_htw[128];          // W256^k / 2, k < 128: the split step's twiddles
    __shared__ float s_ham[256];
    __shared__ float2 s_ltw[kLaneTw];
    __shared__ __align__(16) float2 s_z[kStftWarps][kZPad];

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        if (i < 128) s_htw[i] = make_float2(0.5f * g_tables.tw256[i].x, 0.5f * g_tables.tw256[i].y);
        s_ham[i] = g_tables.ham[i];
    }
    fill_lane_tw(s_ltw, g_tables.tw256);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long g0 = (long long)blockIdx.x * rows_per_cta;
    long long g1 = g0 + rows_per_cta;
    if (g1 > p.total_rows) g1 = p.total_rows;

    // utterance of the first row by binary search; later rows walk forward
    int u = 0;
    {
        int a = 0, b = p.n_utt;
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (__ldg(p.row_off + m) <= g0) a = m; else b = m;
        }
        u = a;
    }
    long long lo = __ldg(p.row_off + u), hi = __ldg(p.row_off + u + 1);

    for (long long g = g0 + warp; g < g1; g += kStftWarps) {
        while (g >= hi) { ++u; lo = hi; hi = __ldg(p.row_off + u + 1); }
        const long long t = g - lo;
        const long long L = __ldg(p.wav_len + u);
        float* mag = p.mag + g * 129;
        float2* ph = p.phase ? p.phase + g * 129 : nullptr;

        if (t >= frames_of(L)) {   // batch padding row: X = 0 -> |X| = 0, exp(j*angle(0)) = 1
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                mag[lane + 32 * a] = 0.f;
                if (ph) ph[lane + 32 * a] = make_float2(1.f, 0.f);
            }
            if (lane == 0) { mag[128] = 0.f; if (ph) ph[128] = make_float2(1.f, 0.f); }
            continue;
        }

        const long long woff = __ldg(p.wav_off + u);
        const float* __restrict__ s = p.wav + woff;
        // the row this warp transforms next (kStftWarps frames on): its second half is new to the CTA (the first half
        // is the neighbouring warp's second), 512 bytes = 4 lines; warp 0 has no such neighbour.  Measured on
        // configs[1]: 0.166 -> 0.160 ms into L2; the same into L1 0.162; the same in K3 gains nothing.
        if (g + kStftWarps < g1 && g + kStftWarps < hi && 128 * (t + kStftWarps) + 256 <= L) {
            const float* nx = s + 128 * (t + kStftWarps);
            if (lane < 4) prefetch_line(nx + 128 + 32 * lane);
            else if (warp == 0 && lane < 8) prefetch_line(nx + 32 * (lane - 4));
        }
        float2 v[4];
        if (((woff | (long long)(uintptr_t)p.wav >> 2) & 1) == 0 && 128 * t + 256 <= L) {
            // Frame inside the signal, sample pairs 8-byte aligned: four coalesced 8-byte loads per lane (256
            // contiguous bytes per warp instruction); the sample in front of a pair comes from the neighbouring
            // lane, only lane 0 reads the one in front of the frame.
            const float2* __restrict__ s2 = reinterpret_cast<const float2*>(s + 128 * t);
            float2 q[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) q[a] = __ldg(s2 + lane + 32 * a);
            float edge = 0.f;
            if (lane == 0 && t > 0) edge = __ldg(s + 128 * t - 1);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int n = lane + 32 * a;
                float sm = __shfl_up_sync(0xffffffffu, q[a].y, 1);
                const float wrap = __shfl_sync(0xffffffffu, a > 0 ? q[a > 0 ? a - 1 : 0].y : edge, a > 0 ? 31 : 0);
                if (lane == 0) sm = wrap;
                // audio_feature.py:54 in float32 without contraction
                const float e0 = (a == 0 && lane == 0 && t == 0) ? q[a].x : __fsub_rn(q[a].x, __fmul_rn(0.97f, sm));
                const float e1 = __fsub_rn(q[a].y, __fmul_rn(0.97f, q[a].x));
                const float2 hw = *reinterpret_cast<const float2*>(s_ham + 2 * n);
                v[a] = make_float2(e0 * hw.x, e1 * hw.y);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int n = lane + 32 * a;
                const long long i0 = 128 * t + 2 * n;
                const float sm = (i0 >= 1 && i0 - 1 < L) ? __ldg(s + i0 - 1) : 0.f;
                const float s0 = (i0 < L) ? __ldg(s + i0) : 0.f;
                const float s1 = (i0 + 1 < L) ? __ldg(s + i0 + 1) : 0.f;
                // audio_feature.py:54 in float32 without contraction; zero padding is appended AFTER
                // the emphasis (audio_feature.py:71-73), so samples >= L are exactly 0
                float e0 = (i0 == 0) ? s0 : __fsub_rn(s0, __fmul_rn(0.97f, sm));
                float e1 = __fsub_rn(s1, __fmul_rn(0.97f, s0));
                if (i0 >= L) e0 = 0.f;
                if (i0 + 1 >= L) e1 = 0.f;
                v[a] = make_float2(e0 * s_ham[2 * n], e1 * s_ham[2 * n + 1]);
            }
        }
        fft128_warp<false>(v, lane, s_ltw);
        const int k0 = zpad(4 * bitrev5(lane));   // (a run of four never crosses a padding step)
        *reinterpret_cast<float4*>(&s_z[warp][k0]) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        *reinterpret_cast<float4*>(&s_z[warp][k0 + 2]) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
        __syncwarp();
        // split step: X_k = E_k + W^k O_k from Z_k and Z_{128-k}, E = (Z_k + conj Z_{128-k}) / 2, O = -i (Z_k - conj
        // Z_{128-k}) / 2, with the halves folded into the table (s_htw = W^k / 2: scaling by 1/2 is exact, so the bits
        // are those of the unfolded form); |X| and X / |X| with one reciprocal square root (FLT_MIN below it: no
        // subnormal argument, |X| = 0 stays 0)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int k = lane + 32 * a;
            const float2 zk = s_z[warp][zpad(k)];
            const float2 zn = cconj(s_z[warp][zpad((128 - k) & 127)]);
            const float2 sm = cadd(zk, zn), d = csub(zk, zn);
            const float2 h = s_htw[k];
            const float2 x = make_float2(fmaf(0.5f, sm.x, fmaf(h.x, d.y, h.y * d.x)), fmaf(0.5f, sm.y, fmaf(-h.x, d.x, h.y * d.y)));
            const float m2 = fmaf(x.x, x.x, x.y * x.y);
            float inv;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(m2, 1.17549435e-38f)));
            mag[k] = m2 * inv;
            if (ph) ph[k] = make_float2(m2 > 0.f ? x.x * inv : 1.f, x.y * inv);   // exp(j angle(0)) = 1
        }
        if (lane == 0) {   // bin 128 is real: X_128 = Re Z_0 - Im Z_0
            const float2 z0 = s_z[warp][0];
            const float x = z0.x - z0.y;
            mag[128] = fabsf(x);
            if (ph) ph[128] = make_float2(x < 0.f ? -1.f : 1.f, 0.f);   // exp(j angle(X)): +-1 (1 at X = 0)
        }
        __syncwarp();
    }
}

cudaError_t upload_tables_stft() { return upload_tables_local(); }

cudaError_t launch_stft(const StftParams& p, cudaStream_t stream) {
    if (p.total_rows <= 0) return cudaSuccess;
    // latency of a small launch is the rows one warp walks: spread them when the GPU would not be full anyway
    const int rows_per_cta = p.total_rows < 148LL * 2 * kRowsPerCta ? kRowsPerCtaSmall : kRowsPerCta;
    const long long ctas = (p.total_rows + rows_per_cta - 1) / rows_per_cta;
    rced_stft_kernel<<<(unsigned)ctas, kStftWarps * 32, 0, stream>>>(p, rows_per_cta);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rced
