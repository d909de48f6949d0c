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
    __shared__ __align__(16) float2 s_ze[N512 ? 1 : kIstftWarps][kZPad];   // irfft_n = 256: the transform's output, padded
    // irfft_n = 512: the 32 kept outputs of both transforms, [b][lane's block] with rows of 10 slots: the 8 lanes of a store
    // write 128 contiguous bytes, the 8 lanes of a quarter-warp load hit slots that differ modulo 8 (no bank conflicts)
    constexpr int kZzRow = 10;
    __shared__ __align__(16) float4 s_zz[N512 ? kIstftWarps : 1][4 * kZzRow];
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
            // irfft_n = 512) -> Z_k = E_k + i O_k -> 128-point inverse FFT.
            // Z = E + i O, E = (A_k + conj A_{128-k}) / 2, O = (A_k - conj A_{128-k}) / 2 * W256^{-k}; the halves are
            // folded into the table (s_htw = W256^{-k} / 2: scaling by 1/2 is exact, same bits)
            const int m0 = 128 * half + 4 * lane;      // position inside the 256-sample frame
            const int br = bitrev5(lane);              // the transform leaves lane l with outputs 4 bitrev5(l) + b
            if (N512) {
                // the two sequences run as packed pairs from here to the staging buffer: x = (Re Z, Re Z'), y = (Im Z, Im Z')
                pk2 px[4], py[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int k = lane + 32 * a;           // 0..127
                    const int kn = 128 - k;                // 128..1
                    float2 ak = s_y[warp][k], an = s_y[warp][kn];
                    float2 bk = cmul(ak, s_tw512[k]), bn = cmul(an, s_tw512[kn]);
                    if (k == 0) {
                        ak = make_float2(y0.x, 0.f);
                        an = make_float2(2.f * y128.x, 0.f);
                        bk = ak;
                        bn = make_float2(-2.f * y128.y, 0.f);
                    }
                    const float2 hw = s_htw[k];
                    const pk2 ax = pk_pack(ak.x, bk.x), ay = pk_pack(ak.y, bk.y), nx = pk_pack(an.x, bn.x), ny = pk_pack(an.y, bn.y);
                    // sm = A_k + conj A_n, d = A_k - conj A_n, o = d * hw, Z = (sm.x / 2 - o.y, sm.y / 2 + o.x)
                    const pk2 smx = pk_add(ax, nx), smy = pk_sub(ay, ny), dx = pk_sub(ax, nx), dy = pk_add(ay, ny);
                    const pk2 ox = pk_fma(dx, pk_pack(hw.x, hw.x), pk_mul(dy, pk_pack(-hw.y, -hw.y)));
                    const pk2 noy = pk_fma(dx, pk_pack(-hw.y, -hw.y), pk_mul(dy, pk_pack(-hw.x, -hw.x)));
                    px[a] = pk_fma(smx, pk_pack(0.5f, 0.5f), noy);
                    py[a] = pk_fma(smy, pk_pack(0.5f, 0.5f), ox);
                }
                fft128_warp_pair<true>(px, py, lane, s_ltw);
                // of the 128 outputs the segment keeps the 32 with index 32 half + lane: the 8 lanes that hold them stage
                // them as (Re z, Re z', Im z, Im z') = y[4n .. 4n+3] * 256
                if ((br >> 3) == half) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        float4 q;
                        pk_unpack(px[b], q.x, q.y);
                        pk_unpack(py[b], q.z, q.w);
                        s_zz[warp][kZzRow * b + (br & 7)] = q;
                    }
                }
                __syncwarp();
                v = s_zz[warp][kZzRow * (lane & 3) + (lane >> 2)];
            } else {
                float2 ze[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int k = lane + 32 * a;
                    const int kn = 128 - k;
                    float2 ak = s_y[warp][k], an = s_y[warp][kn];
                    if (k == 0) {
                        ak = make_float2(y0.x, 0.f);
                        an = make_float2(y128.x, 0.f);
                    }
                    const float2 hw = s_htw[k];
                    const float2 cn = cconj(an);
                    const float2 sm = cadd(ak, cn), o = cmul(csub(ak, cn), hw);
                    ze[a] = make_float2(fmaf(0.5f, sm.x, -o.y), fmaf(0.5f, sm.y, o.x));
                }
                fft128_warp<true>(ze, lane, s_ltw);
                // the segment keeps the 64 outputs 64 half + 2 lane (+ 1): y[2n] = Re z[n] / 128, y[2n+1] = Im z[n] / 128
                const int k0 = zpad(4 * br);   // (a run of four never crosses a padding step)
                if ((br >> 4) == half) {
                    *reinterpret_cast<float4*>(&s_ze[warp][k0]) = make_float4(ze[0].x, ze[0].y, ze[1].x, ze[1].y);
                    *reinterpret_cast<float4*>(&s_ze[warp][k0 + 2]) = make_float4(ze[2].x, ze[2].y, ze[3].x, ze[3].y);
                }
                __syncwarp();
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
