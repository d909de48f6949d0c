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
constexpr int kRowsPerCta = 64;

__device__ __forceinline__ long long frames_of(long long L) {   // audio_feature.py:67-70
    const long long d = L >= 256 ? L - 256 : 256 - L;
    return (d + 127) / 128 + 1;
}

__global__ void __launch_bounds__(kStftWarps * 32) rced_stft_kernel(const StftParams p) {
    __shared__ float2 s_tw[256];
    __shared__ float s_ham[256];
    __shared__ float2 s_z[kStftWarps][128];

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_tw[i] = g_tables.tw256[i];
        s_ham[i] = g_tables.ham[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long g0 = (long long)blockIdx.x * kRowsPerCta;
    long long g1 = g0 + kRowsPerCta;
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

        const float* __restrict__ s = p.wav + __ldg(p.wav_off + u);
        float2 v[4];
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
        fft128_warp<false>(v, lane, s_tw);
        const int k0 = 4 * bitrev5(lane);
#pragma unroll
        for (int b = 0; b < 4; ++b) s_z[warp][k0 + b] = v[b];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const int k = a < 4 ? lane + 32 * a : 128;
            if (a == 4 && lane != 0) break;
            const float2 zk = s_z[warp][k & 127];
            const float2 zn = cconj(s_z[warp][(128 - k) & 127]);
            const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
            const float2 d = csub(zk, zn);
            const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);   // -i/2 * d
            const float2 x = cadd(e, cmul(s_tw[k], o));
            const float m = sqrtf(fmaf(x.x, x.x, x.y * x.y));
            mag[k] = m;
            if (ph) ph[k] = m > 0.f ? make_float2(x.x / m, x.y / m) : make_float2(1.f, 0.f);
        }
        __syncwarp();
    }
}

cudaError_t upload_tables_stft() { return upload_tables_local(); }

cudaError_t launch_stft(const StftParams& p, cudaStream_t stream) {
    if (p.total_rows <= 0) return cudaSuccess;
    const long long ctas = (p.total_rows + kRowsPerCta - 1) / kRowsPerCta;
    rced_stft_kernel<<<(unsigned)ctas, kStftWarps * 32, 0, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rced
