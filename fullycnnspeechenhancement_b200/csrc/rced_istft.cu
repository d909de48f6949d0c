// K3: noisy-phase reconstruction.
//
// Replaces AudioReBuild.rebuild_audio (model_utils/utils.py:171-183):
//   pred * phase -> np.fft.irfft(., nfft)[:256] -> / hamming(256) -> half-frame concatenation
//   (de_frame, utils.py:139-147: NOT an overlap-add) -> de-emphasis y[i] = x[i] + 0.97 y[i-1]
//   (utils.py:104-113) -> truncate to the signal length.
//
// Output sample s of an utterance lives in 128-sample SEGMENT j = s / 128: segment 0 is the
// first half of frame 0, segment j >= 1 the second half of frame j-1.  A CTA of 8 warps owns a
// run of `chunk_segs` segments of one utterance; each warp inverse-transforms one segment at a
// time (two 128-point complex FFTs give the 256 needed samples of the 512-point irfft, see
// tools/fft_emulator.py), the de-emphasis recurrence is a weighted scan across the CTA with a
// running carry.  A chunk that does not start the utterance first replays the 8 previous
// segments to rebuild its carry: 0.97^1024 ~ 3e-14 is below float32 resolution.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rced_fft.cuh"
#include "rced_internal.h"

namespace rced {

constexpr int kIstftWarps = 8;

struct ScanConsts {
    float a1, a2, a3, a4;        // 0.97^(1..4)
    float r1, r2, r4, r8, r16;   // 0.97^(4 d)
    float a128;                  // 0.97^128
};

template <bool N512>
__global__ void __launch_bounds__(kIstftWarps * 32) rced_istft_kernel(const IstftParams p, const ScanConsts sc) {
    __shared__ float2 s_htw[128];          // W256^{-k} / 2, k < 128
    __shared__ float2 s_tw512[132];
    __shared__ __align__(16) float s_iham[256];   // 1 / (hamming * irfft length): the transform's scale folded in (a power of two)
    __shared__ float2 s_ltw[kLaneTw];
    __shared__ float2 s_y[kIstftWarps][132];
    __shared__ __align__(16) float2 s_ze[kIstftWarps][kZPad];
    __shared__ __align__(16) float2 s_zo[kIstftWarps][kZPad];
    __shared__ float s_wend[2][kIstftWarps];

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        if (i < 128) s_htw[i] = make_float2(0.5f * g_tables.tw256[i].x, -0.5f * g_tables.tw256[i].y);
        s_iham[i] = g_tables.inv_ham[i] * (N512 ? 1.f / 256.f : 1.f / 128.f);
        if (i < 132) s_tw512[i] = g_tables.tw512[i];
    }
    fill_lane_tw(s_ltw, g_tables.tw256);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int u = blockIdx.x;
    const long long lo = __ldg(p.row_off + u), hi = __ldg(p.row_off + u + 1);
    const long long rows = hi - lo;
    const long long nseg = rows + 1;
    const long long j0 = (long long)blockIdx.y * p.chunk_segs;
    if (j0 >= nseg || rows <= 0) return;
    long long j1 = j0 + p.chunk_segs;
    if (j1 > nseg) j1 = nseg;
    const long long out_len = __ldg(p.out_len + u);
    if (j0 * 128 >= out_len) return;   // nothing of this chunk survives the truncation
    float* __restrict__ out = p.out + __ldg(p.out_off + u);

    const float a4lane = (float)pow(0.97, 4.0 * lane);
    float carry = 0.f;
    int buf = 0;

    for (long long jg = (j0 >= kIstftWarps ? j0 - kIstftWarps : 0); jg < j1; jg += kIstftWarps) {
        const long long j = jg + warp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < j1 && j * 128 < out_len) {
            const long long row = lo + (j > 0 ? j - 1 : 0);
            const int half = j > 0 ? 1 : 0;
            const float* pr = p.pred + row * 129;
            const float2* ph = p.phase + row * 129;
            // Y = pred * phase (utils.py:119-126)
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int k = lane + 32 * a;
                const float m = __ldg(pr + k);
                const float2 q = __ldg(ph + k);
                s_y[warp][k] = make_float2(m * q.x, m * q.y);
            }
            if (lane == 0) {
                const float m = __ldg(pr + 128);
                const float2 q = __ldg(ph + 128);
                s_y[warp][128] = make_float2(m * q.x, m * q.y);
            }
            __syncwarp();
            const float2 y0 = s_y[warp][0], y128 = s_y[warp][128];
            // Hermitian half-spectra A (even output samples) and A' (odd output samples, only for
            // irfft_n = 512) -> Z_k = E_k + i O_k -> 128-point inverse FFT
            float2 ze[4], zo[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int k = lane + 32 * a;           // 0..127
                const int kn = 128 - k;                // 128..1
                float2 ak = s_y[warp][k], an = s_y[warp][kn];
                float2 bk = ak, bn = an;
                if (N512) {
                    bk = cmul(ak, s_tw512[k]);
                    bn = cmul(an, s_tw512[kn]);
                }
                if (k == 0) {
                    ak = make_float2(y0.x, 0.f);
                    an = make_float2(N512 ? 2.f * y128.x : y128.x, 0.f);
                    bk = ak;
                    bn = make_float2(-2.f * y128.y, 0.f);
                }
                // Z = E + i O, E = (A_k + conj A_{128-k}) / 2, O = (A_k - conj A_{128-k}) / 2 * W256^{-k}; the halves
                // are folded into the table (s_htw = W256^{-k} / 2: scaling by 1/2 is exact, same bits)
                const float2 hw = s_htw[k];
                {
                    const float2 cn = cconj(an);
                    const float2 sm = cadd(ak, cn), o = cmul(csub(ak, cn), hw);
                    ze[a] = make_float2(fmaf(0.5f, sm.x, -o.y), fmaf(0.5f, sm.y, o.x));
                }
                if (N512) {
                    const float2 cn = cconj(bn);
                    const float2 sm = cadd(bk, cn), o = cmul(csub(bk, cn), hw);
                    zo[a] = make_float2(fmaf(0.5f, sm.x, -o.y), fmaf(0.5f, sm.y, o.x));
                }
            }
            fft128_warp<true>(ze, lane, s_ltw);
            if (N512) fft128_warp<true>(zo, lane, s_ltw);
            // of the 128 outputs only the 32 (irfft_n = 512) or 64 behind this segment are read back: the lanes that
            // hold them stage them
            const int br = bitrev5(lane);
            const int k0 = zpad(4 * br);   // (a run of four never crosses a padding step)
            if ((N512 ? br >> 3 : br >> 4) == half) {
                *reinterpret_cast<float4*>(&s_ze[warp][k0]) = make_float4(ze[0].x, ze[0].y, ze[1].x, ze[1].y);
                *reinterpret_cast<float4*>(&s_ze[warp][k0 + 2]) = make_float4(ze[2].x, ze[2].y, ze[3].x, ze[3].y);
                if (N512) {
                    *reinterpret_cast<float4*>(&s_zo[warp][k0]) = make_float4(zo[0].x, zo[0].y, zo[1].x, zo[1].y);
                    *reinterpret_cast<float4*>(&s_zo[warp][k0 + 2]) = make_float4(zo[2].x, zo[2].y, zo[3].x, zo[3].y);
                }
            }
            __syncwarp();
            const int m0 = 128 * half + 4 * lane;      // position inside the 256-sample frame
            if (N512) {
                // y[4n..4n+3] = (Re ze[n], Re zo[n], Im ze[n], Im zo[n]) / 256
                const int n = 32 * half + lane;
                const float2 e = s_ze[warp][zpad(n)], o = s_zo[warp][zpad(n)];
                v = make_float4(e.x, o.x, e.y, o.y);
            } else {
                // y[2n] = Re z[n] / 128, y[2n+1] = Im z[n] / 128
                const int n = 64 * half + 2 * lane;
                const float2 e = s_ze[warp][zpad(n)], o = s_ze[warp][zpad(n + 1)];
                v = make_float4(e.x, e.y, o.x, o.y);
            }
            // scale and de-window (utils.py:128-137)
            const float4 ih = *reinterpret_cast<const float4*>(s_iham + m0);
            v.x *= ih.x; v.y *= ih.y; v.z *= ih.z; v.w *= ih.w;
            __syncwarp();
        }

        // ---- de-emphasis: weighted scan over the 8 x 128 samples of this group ---------------
        const float p0 = v.x;
        const float p1 = fmaf(sc.a1, p0, v.y);
        const float p2 = fmaf(sc.a1, p1, v.z);
        const float p3 = fmaf(sc.a1, p2, v.w);
        float e = p3;
        float t;
        t = __shfl_up_sync(0xffffffffu, e, 1);  if (lane >= 1)  e = fmaf(sc.r1, t, e);
        t = __shfl_up_sync(0xffffffffu, e, 2);  if (lane >= 2)  e = fmaf(sc.r2, t, e);
        t = __shfl_up_sync(0xffffffffu, e, 4);  if (lane >= 4)  e = fmaf(sc.r4, t, e);
        t = __shfl_up_sync(0xffffffffu, e, 8);  if (lane >= 8)  e = fmaf(sc.r8, t, e);
        t = __shfl_up_sync(0xffffffffu, e, 16); if (lane >= 16) e = fmaf(sc.r16, t, e);
        float eprev = __shfl_up_sync(0xffffffffu, e, 1);
        if (lane == 0) eprev = 0.f;
        if (lane == 31) s_wend[buf][warp] = e;
        __syncthreads();
        float c = carry, cw = carry;
#pragma unroll
        for (int w = 0; w < kIstftWarps; ++w) {
            if (w == warp) cw = c;
            c = fmaf(sc.a128, c, s_wend[buf][w]);
        }
        carry = c;
        buf ^= 1;
        const float cin = fmaf(a4lane, cw, eprev);
        if (j >= j0 && j < j1) {
            const long long s0 = j * 128 + 4 * lane;
            const float o0 = fmaf(sc.a1, cin, p0), o1 = fmaf(sc.a2, cin, p1);
            const float o2 = fmaf(sc.a3, cin, p2), o3 = fmaf(sc.a4, cin, p3);
            if (s0 + 3 < out_len && ((uintptr_t)(out + s0) & 15) == 0) {
                *reinterpret_cast<float4*>(out + s0) = make_float4(o0, o1, o2, o3);
            } else {
                if (s0 + 0 < out_len) out[s0 + 0] = o0;
                if (s0 + 1 < out_len) out[s0 + 1] = o1;
                if (s0 + 2 < out_len) out[s0 + 2] = o2;
                if (s0 + 3 < out_len) out[s0 + 3] = o3;
            }
        }
    }
}

cudaError_t upload_tables_istft() { return upload_tables_local(); }

cudaError_t launch_istft(const IstftParams& p, long long max_rows_per_utt, cudaStream_t stream) {
    if (p.n_utt <= 0 || max_rows_per_utt <= 0) return cudaSuccess;
    ScanConsts sc;
    const double a = 0.97;
    sc.a1 = (float)a; sc.a2 = (float)(a * a); sc.a3 = (float)(a * a * a); sc.a4 = (float)(a * a * a * a);
    sc.r1 = (float)pow(a, 4); sc.r2 = (float)pow(a, 8); sc.r4 = (float)pow(a, 16);
    sc.r8 = (float)pow(a, 32); sc.r16 = (float)pow(a, 64);
    sc.a128 = (float)pow(a, 128);
    const long long nseg = max_rows_per_utt + 1;
    const long long chunks = (nseg + p.chunk_segs - 1) / p.chunk_segs;
    if (chunks > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)p.n_utt, (unsigned)chunks);
    if (p.irfft_n == 512) rced_istft_kernel<true><<<grid, kIstftWarps * 32, 0, stream>>>(p, sc);
    else                  rced_istft_kernel<false><<<grid, kIstftWarps * 32, 0, stream>>>(p, sc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rced
