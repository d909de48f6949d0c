// Host-buffer entry points of the C ABI: rced_enhance_host / rced_enhance_host_async / rced_host_sync.
//
// They replace the body of the reference's batch loop -- parse_audio -> power_spectrum / divide_phase ->
// sess.run -> rebuild_audio (model_utils/tester.py:104-113, infer.py:54-71) -- for callers that hold their
// waveforms in HOST memory (numpy arrays): the library owns the device side.  A call is cut into chunks of
// utterances; chunk i runs on stream i % n_streams as
//     H2D waveforms -> H2D metadata -> K1 (STFT) -> K2 (network) -> K3 (reconstruction) -> D2H waveforms
// so that the copies of one chunk overlap the kernels of the others.  Everything the chunks need on the
// device (waveform in / out, spectrogram workspaces, offset tables) belongs to the handle and only grows;
// no allocation and no host synchronisation happens in steady state except the back-pressure on the
// metadata ring (the host may run at most kMetaRing chunks ahead of a stream).
//
// Range guard of the tensor-core network kernel: the device-pointer ABI queues the FP32 kernel behind every
// tensor-core launch (it returns at once unless the guard tripped).  Between the persistent launches of different
// streams that kernel -- 148 CTAs that need most of an SM's shared memory -- has to wait for whole CTAs of the
// neighbouring chunk to retire, which holds back the chunk's reconstruction and download.  The host pipeline
// therefore copies each launch's guard words to page-locked memory instead and looks at them when it
// synchronises: a chunk whose guard tripped (activations beyond the FP16 range, non-finite input: rare) is
// recomputed from the caller's buffers with the FP32 kernel before the call returns.
//
// Caller's buffers: any host memory works (cudaMemcpyAsync); page-locked memory (cudaHostAlloc /
// cudaHostRegister / torch pin_memory) is what makes the copies asynchronous and the chunks overlap.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/rced.h"
#include "rced_handle.h"
#include "rced_internal.h"

namespace rced {

constexpr int kMaxStreams = 4;
constexpr int kMetaRing = 4;   // chunks a stream's metadata staging can hold
constexpr int kPadGap = 16;    // samples: smaller gaps between consecutive outputs count as padding and may be overwritten

struct StreamCtx {
    cudaStream_t stream = nullptr;
    float *d_wav = nullptr, *d_out = nullptr;           // chunk-local waveform in / out
    float *ws_mag = nullptr, *ws_phase = nullptr, *ws_pred = nullptr;
    unsigned char* d_meta = nullptr;                    // offset tables of the chunk in flight
    unsigned char* h_meta[kMetaRing] = {};              // pinned staging of the tables
    cudaEvent_t meta_used[kMetaRing] = {};              // recorded behind the H2D copy that read h_meta[i]
    size_t cap_wav = 0, cap_out = 0, cap_rows = 0, cap_meta = 0;
    unsigned int uses = 0;
};

constexpr int kPendingMax = 1024;   // chunks whose guard words wait for the next synchronisation

struct PendingChunk {   // what is needed to recompute a chunk whose range guard tripped
    const float* wav;
    const int64_t* wav_off;
    const int32_t* wav_len;
    int c0, c1, irfft_n;
    float* out;
    const int64_t* out_off;
    const int32_t* out_len;
};

struct HostPipe {
    int n_streams = 3;
    // target spectrogram rows per chunk.  A synchronous call is cut finely (its first upload and last download
    // overlap nothing, so they should be short); asynchronous calls pipeline behind each other and prefer few,
    // large chunks (every chunk boundary is a gap in which the small kernels wait for persistent CTAs to retire).
    int64_t chunk_rows = 49152, chunk_rows_async = 262144;
    StreamCtx s[kMaxStreams];
    unsigned int next_stream = 0;   // round robin across calls: consecutive small calls use different streams
    unsigned int* h_flags = nullptr;   // page-locked [kPendingMax][2]: guard words of the pending tensor-core launches
    std::vector<PendingChunk> pending;
    bool recomputing = false;
};

static size_t grow(size_t need) { return need + need / 4 + 256; }

template <class T>
static cudaError_t ensure(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) {
        cudaError_t e = cudaFree(p);   // (synchronises the device: only while the buffers still grow)
        if (e != cudaSuccess) return e;
        p = nullptr;
        cap = 0;
    }
    const size_t n = grow(need);
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

void host_pipe_destroy(HostPipe* p) {
    if (!p) return;
    for (StreamCtx& c : p->s) {
        if (c.stream) cudaStreamSynchronize(c.stream);
        cudaFree(c.d_wav);
        cudaFree(c.d_out);
        cudaFree(c.ws_mag);
        cudaFree(c.ws_phase);
        cudaFree(c.ws_pred);
        cudaFree(c.d_meta);
        for (int i = 0; i < kMetaRing; ++i) {
            if (c.h_meta[i]) cudaFreeHost(c.h_meta[i]);
            if (c.meta_used[i]) cudaEventDestroy(c.meta_used[i]);
        }
        if (c.stream) cudaStreamDestroy(c.stream);
    }
    if (p->h_flags) cudaFreeHost(p->h_flags);
    delete p;
}

static int pipe_get(rced_handle* h, HostPipe** out) {
    if (!h->pipe) {
        HostPipe* p = new HostPipe();
        for (int i = 0; i < kMaxStreams; ++i) {
            cudaError_t e = cudaStreamCreateWithFlags(&p->s[i].stream, cudaStreamNonBlocking);
            for (int j = 0; j < kMetaRing && e == cudaSuccess; ++j) e = cudaEventCreateWithFlags(&p->s[i].meta_used[j], cudaEventDisableTiming);
            if (e != cudaSuccess) {
                host_pipe_destroy(p);
                return cuda_fail(e, "host pipeline: stream / event creation");
            }
        }
        cudaError_t e = cudaHostAlloc(&p->h_flags, sizeof(unsigned int) * 2 * kPendingMax, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            host_pipe_destroy(p);
            return cuda_fail(e, "host pipeline: pinned guard words");
        }
        p->pending.reserve(kPendingMax);
        h->pipe = p;
    }
    *out = h->pipe;
    return RCED_OK;
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// one chunk [c0, c1) of the call on stream context c
static int run_chunk(rced_handle* h, StreamCtx& c, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int c0, int c1,
                     int irfft_n, float* out, const int64_t* out_off, const int32_t* out_len) {
    const int n = c1 - c0;
    // sample ranges of the chunk in the caller's buffers, frame counts
    int64_t in_lo = INT64_MAX, in_hi = 0, o_lo = INT64_MAX, o_hi = 0, rows = 0, max_rows = 0;
    bool out_contiguous = true;
    for (int u = c0; u < c1; ++u) {
        in_lo = std::min(in_lo, wav_off[u]);
        in_hi = std::max(in_hi, wav_off[u] + wav_len[u]);
        o_lo = std::min(o_lo, out_off[u]);
        o_hi = std::max(o_hi, out_off[u] + out_len[u]);
        // gaps of fewer than kPadGap samples between consecutive outputs are alignment padding (include/rced.h)
        if (u > c0 && (out_off[u] < out_off[u - 1] + out_len[u - 1] || out_off[u] - (out_off[u - 1] + out_len[u - 1]) >= kPadGap))
            out_contiguous = false;
        const int64_t t = rced_num_frames(wav_len[u]);
        rows += t;
        max_rows = std::max(max_rows, t);
    }
    // tables: wav_off[n] | row_off[n+1] | out_off[n] (int64), wav_len[n] | out_len[n] (int32)
    const size_t o_wav_off = 0, o_row_off = o_wav_off + 8 * (size_t)n, o_out_off = o_row_off + 8 * (size_t)(n + 1);
    const size_t o_wav_len = o_out_off + 8 * (size_t)n, o_out_len = align16(o_wav_len + 4 * (size_t)n);
    const size_t meta_bytes = align16(o_out_len + 4 * (size_t)n);

    cudaError_t e;
    if ((e = ensure(c.d_wav, c.cap_wav, (size_t)(in_hi - in_lo))) != cudaSuccess) return cuda_fail(e, "host pipeline: waveform buffer");
    if ((e = ensure(c.d_out, c.cap_out, (size_t)(o_hi - o_lo))) != cudaSuccess) return cuda_fail(e, "host pipeline: output buffer");
    if ((size_t)rows > c.cap_rows) {
        cudaFree(c.ws_mag);     // (cudaFree synchronises the device: queued work that uses the old buffers has finished)
        cudaFree(c.ws_phase);
        cudaFree(c.ws_pred);
        c.ws_mag = c.ws_phase = c.ws_pred = nullptr;
        c.cap_rows = 0;
        const size_t nr = grow((size_t)rows);
        if ((e = cudaMalloc(&c.ws_mag, nr * RCED_FREQ_BINS * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&c.ws_phase, nr * RCED_FREQ_BINS * 2 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&c.ws_pred, nr * RCED_FREQ_BINS * sizeof(float))) != cudaSuccess)
            return cuda_fail(e, "host pipeline: spectrogram workspaces");
        c.cap_rows = nr;
    }
    if (meta_bytes > c.cap_meta) {
        cudaStreamSynchronize(c.stream);   // the old tables may still be read by queued work
        const size_t nb = (grow(meta_bytes) + 255) & ~(size_t)255;   // every ring slot starts 256-byte aligned
        cudaFree(c.d_meta);
        c.d_meta = nullptr;
        for (int i = 0; i < kMetaRing; ++i) {
            if (c.h_meta[i]) cudaFreeHost(c.h_meta[i]);
            c.h_meta[i] = nullptr;
        }
        c.cap_meta = 0;
        if ((e = cudaMalloc(&c.d_meta, nb * kMetaRing)) != cudaSuccess) return cuda_fail(e, "host pipeline: table buffer");
        for (int i = 0; i < kMetaRing; ++i)
            if ((e = cudaHostAlloc(&c.h_meta[i], nb, cudaHostAllocDefault)) != cudaSuccess) return cuda_fail(e, "host pipeline: pinned table staging");
        c.cap_meta = nb;
    }
    const int slot = (int)(c.uses++ % kMetaRing);
    // back-pressure: the copy that read this staging slot kMetaRing chunks ago has run
    if ((e = cudaEventSynchronize(c.meta_used[slot])) != cudaSuccess) return cuda_fail(e, "host pipeline: table staging");
    unsigned char* hm = c.h_meta[slot];
    unsigned char* dm = c.d_meta + (size_t)slot * c.cap_meta;
    int64_t* m_wav_off = reinterpret_cast<int64_t*>(hm + o_wav_off);
    int64_t* m_row_off = reinterpret_cast<int64_t*>(hm + o_row_off);
    int64_t* m_out_off = reinterpret_cast<int64_t*>(hm + o_out_off);
    int32_t* m_wav_len = reinterpret_cast<int32_t*>(hm + o_wav_len);
    int32_t* m_out_len = reinterpret_cast<int32_t*>(hm + o_out_len);
    int64_t r = 0;
    for (int i = 0; i < n; ++i) {
        const int u = c0 + i;
        m_wav_off[i] = wav_off[u] - in_lo;
        m_out_off[i] = out_off[u] - o_lo;
        m_wav_len[i] = wav_len[u];
        m_out_len[i] = out_len[u];
        m_row_off[i] = r;
        r += rced_num_frames(wav_len[u]);
    }
    m_row_off[n] = r;

    if ((e = cudaMemcpyAsync(c.d_wav, wav + in_lo, (size_t)(in_hi - in_lo) * sizeof(float), cudaMemcpyHostToDevice, c.stream)) != cudaSuccess)
        return cuda_fail(e, "host pipeline: H2D waveforms");
    if ((e = cudaMemcpyAsync(dm, hm, meta_bytes, cudaMemcpyHostToDevice, c.stream)) != cudaSuccess) return cuda_fail(e, "host pipeline: H2D tables");
    cudaEventRecord(c.meta_used[slot], c.stream);
    const int64_t* d_wav_off = reinterpret_cast<const int64_t*>(dm + o_wav_off);
    const int64_t* d_row_off = reinterpret_cast<const int64_t*>(dm + o_row_off);
    const int64_t* d_out_off = reinterpret_cast<const int64_t*>(dm + o_out_off);
    const int32_t* d_wav_len = reinterpret_cast<const int32_t*>(dm + o_wav_len);
    const int32_t* d_out_len = reinterpret_cast<const int32_t*>(dm + o_out_len);
    int rc = rced_stft(h, c.d_wav, d_wav_off, d_wav_len, d_row_off, n, rows, c.ws_mag, c.ws_phase, c.stream);
    if (rc != RCED_OK) return rc;
    HostPipe* pipe = h->pipe;
    unsigned int* d_flags = nullptr;
    rc = forward_impl(h, c.ws_mag, d_row_off, n, rows, c.ws_pred, c.stream, pipe->recomputing ? nullptr : &d_flags);
    if (rc != RCED_OK) return rc;
    if (d_flags) {   // tensor-core launch: its guard words travel to the host behind it, the chunk is remembered
        e = cudaMemcpyAsync(pipe->h_flags + 2 * pipe->pending.size(), d_flags, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream);
        if (e != cudaSuccess) return cuda_fail(e, "host pipeline: D2H guard words");
        pipe->pending.push_back(PendingChunk{wav, wav_off, wav_len, c0, c1, irfft_n, out, out_off, out_len});
    }
    rc = rced_istft(h, c.ws_pred, c.ws_phase, d_row_off, n, max_rows, irfft_n, c.d_out, d_out_off, d_out_len, c.stream);
    if (rc != RCED_OK) return rc;
    if (out_contiguous) {
        e = cudaMemcpyAsync(out + o_lo, c.d_out, (size_t)(o_hi - o_lo) * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    } else {   // gaps between the outputs belong to the caller: copy utterance by utterance
        e = cudaSuccess;
        for (int u = c0; u < c1 && e == cudaSuccess; ++u)
            e = cudaMemcpyAsync(out + out_off[u], c.d_out + (out_off[u] - o_lo), (size_t)out_len[u] * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    }
    return e == cudaSuccess ? RCED_OK : cuda_fail(e, "host pipeline: D2H waveforms");
}

}  // namespace rced

using namespace rced;

extern "C" {

int rced_host_config(rced_handle* h, int n_streams, int64_t chunk_rows) {
    if (!h) return fail(RCED_ERR_ARG, "null handle");
    if (n_streams < 1 || n_streams > kMaxStreams) return fail(RCED_ERR_ARG, "n_streams must be 1.." + std::to_string(kMaxStreams));
    if (chunk_rows < 1) return fail(RCED_ERR_ARG, "chunk_rows must be positive");
    DeviceGuard guard(h->device);
    HostPipe* p = nullptr;
    const int rc = pipe_get(h, &p);
    if (rc != RCED_OK) return rc;
    p->n_streams = n_streams;
    p->chunk_rows = chunk_rows;
    p->chunk_rows_async = chunk_rows;
    return RCED_OK;
}

static int host_sync_impl(rced_handle* h) {
    HostPipe* p = h->pipe;
    if (!p) return RCED_OK;
    for (int i = 0; i < kMaxStreams; ++i) {
        cudaError_t e = cudaStreamSynchronize(p->s[i].stream);
        if (e != cudaSuccess) return cuda_fail(e, "rced_host_sync");
    }
    // range guard of the tensor-core launches since the last synchronisation (see the head of this file)
    int rc = RCED_OK;
    const std::vector<PendingChunk> pend = p->pending;
    p->pending.clear();
    for (size_t i = 0; i < pend.size() && rc == RCED_OK; ++i) {
        const unsigned int amax = p->h_flags[2 * i], perr = p->h_flags[2 * i + 1];
        if (amax <= 0x477FE000u /* 65504.0f */ && perr == 0u) continue;
        const PendingChunk& q = pend[i];
        const int variant = h->variant;
        h->variant = RCED_VARIANT_FFMA;
        p->recomputing = true;
        rc = run_chunk(h, p->s[0], q.wav, q.wav_off, q.wav_len, q.c0, q.c1, q.irfft_n, q.out, q.out_off, q.out_len);
        p->recomputing = false;
        h->variant = variant;
        if (rc == RCED_OK) {
            cudaError_t e = cudaStreamSynchronize(p->s[0].stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "rced_host_sync (FP32 recomputation)");
        }
    }
    return rc;
}

static int enhance_host_impl(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int n_utt, int irfft_n,
                             float* out, const int64_t* out_off, const int32_t* out_len, bool async) {
    if (!h) return fail(RCED_ERR_ARG, "null handle");
    if (n_utt < 0) return fail(RCED_ERR_ARG, "negative size");
    if (irfft_n != 512 && irfft_n != 256) return fail(RCED_ERR_ARG, "irfft_n must be 512 or 256");
    if (n_utt == 0) return RCED_OK;
    if (!wav || !wav_off || !wav_len || !out || !out_off || !out_len) return fail(RCED_ERR_ARG, "null pointer");
    for (int u = 0; u < n_utt; ++u) {
        if (wav_len[u] < 1) return fail(RCED_ERR_ARG, "every utterance needs at least one sample");
        if (out_len[u] < 0 || (int64_t)out_len[u] > (rced_num_frames(wav_len[u]) + 1) * RCED_FRAME_HOP)
            return fail(RCED_ERR_ARG, "out_len exceeds (frames + 1) * 128");
        if (wav_off[u] < 0 || out_off[u] < 0) return fail(RCED_ERR_ARG, "negative offset");
    }
    DeviceGuard guard(h->device);
    HostPipe* p = nullptr;
    int rc = pipe_get(h, &p);
    if (rc != RCED_OK) return rc;
    // chunks of about `target_rows` spectrogram rows; the first and the last chunk of a long synchronous call are a
    // quarter of that: nothing overlaps the first upload and the last download
    const int64_t chunk_rows = async ? p->chunk_rows_async : p->chunk_rows;
    std::vector<int> bounds;
    bounds.push_back(0);
    int64_t total_rows = 0;
    for (int u = 0; u < n_utt; ++u) total_rows += rced_num_frames(wav_len[u]);
    const bool many = !async && total_rows > 2 * chunk_rows;
    int64_t acc = 0, done = 0;
    for (int u = 0; u < n_utt; ++u) {
        acc += rced_num_frames(wav_len[u]);
        int64_t target = chunk_rows;
        if (many && (bounds.size() == 1 || total_rows - done - acc < chunk_rows / 4)) target = chunk_rows / 4;
        if (acc >= target || u == n_utt - 1) {
            bounds.push_back(u + 1);
            done += acc;
            acc = 0;
        }
    }
    for (size_t i = 0; i + 1 < bounds.size(); ++i) {
        if (p->pending.size() >= (size_t)kPendingMax) {   // the guard-word staging is full: drain
            rc = host_sync_impl(h);
            if (rc != RCED_OK) return rc;
        }
        StreamCtx& c = p->s[p->next_stream++ % (unsigned int)p->n_streams];
        rc = run_chunk(h, c, wav, wav_off, wav_len, bounds[i], bounds[i + 1], irfft_n, out, out_off, out_len);
        if (rc != RCED_OK) return rc;
    }
    return RCED_OK;
}

int rced_enhance_host_async(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int n_utt, int irfft_n,
                            float* out, const int64_t* out_off, const int32_t* out_len) {
    return enhance_host_impl(h, wav, wav_off, wav_len, n_utt, irfft_n, out, out_off, out_len, true);
}

int rced_host_sync(rced_handle* h) {
    if (!h) return fail(RCED_ERR_ARG, "null handle");
    if (!h->pipe) return RCED_OK;
    DeviceGuard guard(h->device);
    return host_sync_impl(h);
}

int rced_enhance_host(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int n_utt, int irfft_n, float* out,
                      const int64_t* out_off, const int32_t* out_len) {
    const int rc = enhance_host_impl(h, wav, wav_off, wav_len, n_utt, irfft_n, out, out_off, out_len, false);
    if (rc != RCED_OK) return rc;
    return rced_host_sync(h);
}

}  // extern "C"
