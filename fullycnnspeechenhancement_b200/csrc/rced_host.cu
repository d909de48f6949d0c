// Host-buffer entry points of the C ABI: rced_enhance_host / rced_enhance_host_async / rced_host_sync.
//
// They replace the body of the reference's batch loop -- parse_audio -> power_spectrum / divide_phase ->
// sess.run -> rebuild_audio (model_utils/tester.py:104-113, infer.py:54-71) -- for callers that hold their
// waveforms in HOST memory (numpy arrays): the library owns the device side.  A call is cut into chunks of
// utterances that flow through three streams,
//     copy-in:  H2D waveforms, H2D offset tables
//     compute:  K1 (STFT) -> K2 (network) -> K3 (reconstruction), chunk after chunk, exactly the sequence of the
//               device-pointer path
//     copy-out: D2H guard words, D2H waveforms
// coupled by events, over a ring of buffer sets (waveform in / out, spectrogram workspaces, offset tables; they
// belong to the handle and only grow).  The kernels of different chunks never compete for SMs -- a first version
// that gave every chunk its own stream lost 4 % to small kernels waiting for the persistent network CTAs of the
// neighbouring chunk to retire -- and the copies of one chunk overlap the kernels of the others.  No allocation
// and no host synchronisation happens in steady state except the back-pressure of the ring (the host runs at
// most kSets chunks ahead of the copy-in stream).
//
// Caller's buffers: any host memory works (cudaMemcpyAsync); page-locked memory (cudaHostAlloc /
// cudaHostRegister / torch pin_memory) is what makes the copies asynchronous and the chunks overlap.
//
// Relay (rced_host_set_relay): on multi-GPU boxes some GPUs reach host memory through a shared, slower link than
// others (measured on the 8 x B200 box of this project: 7.8 GB/s per direction for four of the GPUs against > 20 GB/s
// for the other four).  A handle on such a GPU can move its waveforms through a PEER GPU of the
// same process instead: H2D into a staging buffer on the relay device, then cudaMemcpyPeerAsync over NVLink into the
// handle's own buffer (and the same way back).  The relay's copy engines do the work; the handle's kernels and the
// relay device's own users are not involved.
//
// Range guard of the tensor-core network kernel: the device-pointer ABI queues the FP32 kernel behind every
// tensor-core launch (it returns at once unless the guard tripped).  The host pipeline copies each launch's guard
// words to page-locked memory instead and looks at them when it synchronises: a chunk whose guard tripped
// (activations beyond the FP16 range, non-finite input: rare) is recomputed from the caller's buffers with the FP32
// kernel before the call returns.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/rced.h"
#include "rced_handle.h"
#include "rced_internal.h"

namespace rced {

constexpr int kSets = 4;            // buffer sets in flight
constexpr int kPadGap = 16;         // samples: smaller gaps between consecutive outputs count as padding and may be overwritten
constexpr int kPendingMax = 1024;   // chunks whose guard words wait for the next synchronisation

struct BufferSet {
    float *d_wav = nullptr, *d_out = nullptr;           // chunk-local waveform in / out
    float *ws_mag = nullptr, *ws_phase = nullptr, *ws_pred = nullptr;
    unsigned char* d_meta = nullptr;                    // offset tables of the chunk
    unsigned char* h_meta = nullptr;                    // page-locked staging of the tables
    size_t cap_wav = 0, cap_out = 0, cap_rows = 0, cap_meta = 0;
    cudaEvent_t uploaded = nullptr;    // copy-in:  the set's waveforms and tables are on the device (h_meta may be refilled)
    cudaEvent_t computed = nullptr;    // compute:  K1-K3 done (d_wav / d_meta may be overwritten, d_out may be downloaded)
    cudaEvent_t downloaded = nullptr;  // copy-out: d_out has left (the next K3 of this set may write it)
    // relay mode: staging on the relay device and the events of its two streams
    float *r_in = nullptr, *r_out = nullptr;
    size_t cap_r_in = 0, cap_r_out = 0;
    cudaEvent_t r_uploaded = nullptr, r_downloaded = nullptr;
};

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
    // target spectrogram rows per chunk.  A synchronous call is cut finely (its first upload and last download
    // overlap nothing, so they should be short); asynchronous calls pipeline behind each other and take larger
    // chunks (fewer kernel boundaries).
    int64_t chunk_rows = 32768, chunk_rows_async = 131072;
    cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
    BufferSet set[kSets];
    unsigned int next_set = 0;
    unsigned int* h_flags = nullptr;   // page-locked [kPendingMax][2]: guard words of the pending tensor-core launches
    std::vector<PendingChunk> pending;
    bool recomputing = false;
    int relay = -1;                    // device whose copy engines carry the waveforms (-1: the handle's own)
    int relay_made_on = -1;            // device the relay streams / events / staging were created on
    cudaStream_t r_in = nullptr, r_out = nullptr;   // streams on the relay device
};

static size_t grow(size_t need) { return need + need / 4 + 256; }

void host_pipe_destroy(HostPipe* p) {
    if (!p) return;
    for (cudaStream_t s : {p->s_in, p->s_compute, p->s_out})
        if (s) cudaStreamSynchronize(s);
    for (BufferSet& b : p->set) {
        cudaFree(b.d_wav);
        cudaFree(b.d_out);
        cudaFree(b.ws_mag);
        cudaFree(b.ws_phase);
        cudaFree(b.ws_pred);
        cudaFree(b.d_meta);
        if (b.h_meta) cudaFreeHost(b.h_meta);
        for (cudaEvent_t e : {b.uploaded, b.computed, b.downloaded, b.r_uploaded, b.r_downloaded})
            if (e) cudaEventDestroy(e);
        cudaFree(b.r_in);    // (cudaFree finds the owning device by itself)
        cudaFree(b.r_out);
    }
    for (cudaStream_t s : {p->r_in, p->r_out})
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    for (cudaStream_t s : {p->s_in, p->s_compute, p->s_out})
        if (s) cudaStreamDestroy(s);
    if (p->h_flags) cudaFreeHost(p->h_flags);
    delete p;
}

static int pipe_get(rced_handle* h, HostPipe** out) {
    if (!h->pipe) {
        HostPipe* p = new HostPipe();
        cudaError_t e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
        for (BufferSet& b : p->set)
            for (cudaEvent_t* ev : {&b.uploaded, &b.computed, &b.downloaded})
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc(&p->h_flags, sizeof(unsigned int) * 2 * kPendingMax, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            host_pipe_destroy(p);
            return cuda_fail(e, "host pipeline: streams / events / pinned guard words");
        }
        p->pending.reserve(kPendingMax);
        h->pipe = p;
    }
    *out = h->pipe;
    return RCED_OK;
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// (Re)allocation of a set's buffers.  cudaFree synchronises the device, so queued work that uses the old buffers
// has finished; it only happens while the buffers still grow.
template <class T>
static cudaError_t ensure(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t n = grow(need);
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

// one chunk [c0, c1) of the call through buffer set b
// `inline_copies`: a synchronous call that is a single chunk (one utterance, a streaming block) puts its copies on the
// compute stream as well -- nothing could overlap them, and every hand-over between streams costs a few microseconds
static int run_chunk(rced_handle* h, HostPipe* pipe, BufferSet& b, const float* wav, const int64_t* wav_off, const int32_t* wav_len,
                     int c0, int c1, int irfft_n, float* out, const int64_t* out_off, const int32_t* out_len, bool inline_copies = false) {
    const int n = c1 - c0;
    const cudaStream_t s_in = inline_copies ? pipe->s_compute : pipe->s_in;
    const cudaStream_t s_out = inline_copies ? pipe->s_compute : pipe->s_out;
    // sample ranges of the chunk in the caller's buffers, frame counts
    int64_t in_lo = INT64_MAX, in_hi = 0, o_lo = INT64_MAX, o_hi = 0, rows = 0, max_rows = 0, sum_in = 0, sum_out = 0;
    bool out_contiguous = true;
    for (int u = c0; u < c1; ++u) {
        sum_in += wav_len[u];
        sum_out += out_len[u];
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
    // Utterances that lie close together in the caller's buffers (the normal case: a packed batch) travel as ONE copy
    // per direction and keep their relative offsets on the device.  Scattered utterances (the span is more than twice
    // the samples) are copied one by one into a compact device layout, so that the device buffers follow the samples,
    // not the caller's address span.
    const bool compact_in = in_hi - in_lo > 2 * sum_in + 4096;
    const bool compact_out = o_hi - o_lo > 2 * sum_out + 4096;
    if (compact_in) {
        in_lo = 0;
        in_hi = sum_in + 4 * (int64_t)n;   // every utterance starts 16-byte aligned
    }
    if (compact_out) {
        o_lo = 0;
        o_hi = sum_out + 4 * (int64_t)n;
        out_contiguous = false;
    }
    // tables: wav_off[n] | row_off[n+1] | out_off[n] (int64), wav_len[n] | out_len[n] (int32)
    const size_t o_wav_off = 0, o_row_off = o_wav_off + 8 * (size_t)n, o_out_off = o_row_off + 8 * (size_t)(n + 1);
    const size_t o_wav_len = o_out_off + 8 * (size_t)n, o_out_len = align16(o_wav_len + 4 * (size_t)n);
    const size_t meta_bytes = align16(o_out_len + 4 * (size_t)n);

    cudaError_t e;
    // the host may refill this set's table staging once its previous upload has run (back-pressure: the host is at
    // most kSets chunks ahead of the copy-in stream)
    if ((e = cudaEventSynchronize(b.uploaded)) != cudaSuccess) return cuda_fail(e, "host pipeline: table staging");
    if ((e = ensure(b.d_wav, b.cap_wav, (size_t)(in_hi - in_lo))) != cudaSuccess) return cuda_fail(e, "host pipeline: waveform buffer");
    if ((e = ensure(b.d_out, b.cap_out, (size_t)(o_hi - o_lo))) != cudaSuccess) return cuda_fail(e, "host pipeline: output buffer");
    if ((size_t)rows > b.cap_rows) {
        cudaFree(b.ws_mag);
        cudaFree(b.ws_phase);
        cudaFree(b.ws_pred);
        b.ws_mag = b.ws_phase = b.ws_pred = nullptr;
        b.cap_rows = 0;
        const size_t nr = grow((size_t)rows);
        if ((e = cudaMalloc(&b.ws_mag, nr * RCED_FREQ_BINS * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&b.ws_phase, nr * RCED_FREQ_BINS * 2 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&b.ws_pred, nr * RCED_FREQ_BINS * sizeof(float))) != cudaSuccess)
            return cuda_fail(e, "host pipeline: spectrogram workspaces");
        b.cap_rows = nr;
    }
    if (meta_bytes > b.cap_meta) {
        cudaDeviceSynchronize();   // the old tables may still be read by queued work
        const size_t nb = (grow(meta_bytes) + 255) & ~(size_t)255;
        cudaFree(b.d_meta);
        if (b.h_meta) cudaFreeHost(b.h_meta);
        b.d_meta = nullptr;
        b.h_meta = nullptr;
        b.cap_meta = 0;
        if ((e = cudaMalloc(&b.d_meta, nb)) != cudaSuccess) return cuda_fail(e, "host pipeline: table buffer");
        if ((e = cudaHostAlloc(&b.h_meta, nb, cudaHostAllocDefault)) != cudaSuccess) return cuda_fail(e, "host pipeline: pinned table staging");
        b.cap_meta = nb;
    }
    unsigned char* hm = b.h_meta;
    unsigned char* dm = b.d_meta;
    int64_t* m_wav_off = reinterpret_cast<int64_t*>(hm + o_wav_off);
    int64_t* m_row_off = reinterpret_cast<int64_t*>(hm + o_row_off);
    int64_t* m_out_off = reinterpret_cast<int64_t*>(hm + o_out_off);
    int32_t* m_wav_len = reinterpret_cast<int32_t*>(hm + o_wav_len);
    int32_t* m_out_len = reinterpret_cast<int32_t*>(hm + o_out_len);
    int64_t r = 0, pos_in = 0, pos_out = 0;
    for (int i = 0; i < n; ++i) {
        const int u = c0 + i;
        m_wav_off[i] = compact_in ? pos_in : wav_off[u] - in_lo;
        m_out_off[i] = compact_out ? pos_out : out_off[u] - o_lo;
        pos_in += (wav_len[u] + 3) / 4 * 4;
        pos_out += (out_len[u] + 3) / 4 * 4;
        m_wav_len[i] = wav_len[u];
        m_out_len[i] = out_len[u];
        m_row_off[i] = r;
        r += rced_num_frames(wav_len[u]);
    }
    m_row_off[n] = r;

    // ---- copy-in: behind the kernels that read this set's previous waveforms and tables
    // development aid (tools/e2e_sweep.py): RCED_HOST_NOCOPY=1 leaves the waveform copies out to see what they cost
    // ("in" / "out": only that direction is left out)
    static const char* no_copy_env = getenv("RCED_HOST_NOCOPY");
    static const bool no_copy_in = no_copy_env && strcmp(no_copy_env, "out") != 0;
    static const bool no_copy_out = no_copy_env && strcmp(no_copy_env, "in") != 0;
    const bool no_copy = no_copy_in;
    const bool relay = pipe->relay >= 0 && !pipe->recomputing && !inline_copies && !compact_in && !compact_out;
    const size_t in_bytes = (size_t)(in_hi - in_lo) * sizeof(float), out_bytes = (size_t)(o_hi - o_lo) * sizeof(float);
    cudaStreamWaitEvent(s_in, b.computed, 0);
    if (relay && !no_copy) {
        // host -> relay device (its link to the host) -> this device (NVLink), on the relay's copy-in stream
        cudaSetDevice(pipe->relay);
        e = cudaSuccess;
        if (in_bytes > b.cap_r_in) {
            cudaFree(b.r_in);
            b.r_in = nullptr;
            b.cap_r_in = 0;
            if ((e = cudaMalloc(&b.r_in, grow(in_bytes))) == cudaSuccess) b.cap_r_in = grow(in_bytes);
        }
        if (e == cudaSuccess && out_bytes > b.cap_r_out) {
            cudaFree(b.r_out);
            b.r_out = nullptr;
            b.cap_r_out = 0;
            if ((e = cudaMalloc(&b.r_out, grow(out_bytes))) == cudaSuccess) b.cap_r_out = grow(out_bytes);
        }
        if (e == cudaSuccess) {
            cudaStreamWaitEvent(pipe->r_in, b.computed, 0);   // the kernels that read d_wav before have finished
            e = cudaMemcpyAsync(b.r_in, wav + in_lo, in_bytes, cudaMemcpyHostToDevice, pipe->r_in);
        }
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(b.d_wav, h->device, b.r_in, pipe->relay, in_bytes, pipe->r_in);
        if (e == cudaSuccess) e = cudaEventRecord(b.r_uploaded, pipe->r_in);
        cudaSetDevice(h->device);
        if (e != cudaSuccess) return cuda_fail(e, "host pipeline: H2D waveforms through the relay device");
    } else if (!no_copy && compact_in) {
        e = cudaSuccess;
        for (int i = 0; i < n && e == cudaSuccess; ++i)
            e = cudaMemcpyAsync(b.d_wav + m_wav_off[i], wav + wav_off[c0 + i], (size_t)wav_len[c0 + i] * sizeof(float), cudaMemcpyHostToDevice, s_in);
        if (e != cudaSuccess) return cuda_fail(e, "host pipeline: H2D waveforms");
    } else if (!no_copy &&
        (e = cudaMemcpyAsync(b.d_wav, wav + in_lo, in_bytes, cudaMemcpyHostToDevice, s_in)) != cudaSuccess)
        return cuda_fail(e, "host pipeline: H2D waveforms");
    if ((e = cudaMemcpyAsync(dm, hm, meta_bytes, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return cuda_fail(e, "host pipeline: H2D tables");
    cudaEventRecord(b.uploaded, s_in);

    // ---- compute: behind the upload, and behind the download of what this set's d_out held before
    const int64_t* d_wav_off = reinterpret_cast<const int64_t*>(dm + o_wav_off);
    const int64_t* d_row_off = reinterpret_cast<const int64_t*>(dm + o_row_off);
    const int64_t* d_out_off = reinterpret_cast<const int64_t*>(dm + o_out_off);
    const int32_t* d_wav_len = reinterpret_cast<const int32_t*>(dm + o_wav_len);
    const int32_t* d_out_len = reinterpret_cast<const int32_t*>(dm + o_out_len);
    cudaStreamWaitEvent(pipe->s_compute, b.uploaded, 0);
    cudaStreamWaitEvent(pipe->s_compute, b.downloaded, 0);
    if (pipe->relay >= 0) {   // (events that were never recorded are complete)
        cudaStreamWaitEvent(pipe->s_compute, b.r_uploaded, 0);
        cudaStreamWaitEvent(pipe->s_compute, b.r_downloaded, 0);
    }
    int rc = rced_stft(h, b.d_wav, d_wav_off, d_wav_len, d_row_off, n, rows, b.ws_mag, b.ws_phase, pipe->s_compute);
    if (rc != RCED_OK) return rc;
    unsigned int* d_flags = nullptr;
    rc = forward_impl(h, b.ws_mag, d_row_off, n, rows, b.ws_pred, pipe->s_compute, pipe->recomputing ? nullptr : &d_flags);
    if (rc != RCED_OK) return rc;
    rc = rced_istft(h, b.ws_pred, b.ws_phase, d_row_off, n, max_rows, irfft_n, b.d_out, d_out_off, d_out_len, pipe->s_compute);
    if (rc != RCED_OK) return rc;
    cudaEventRecord(b.computed, pipe->s_compute);

    // ---- copy-out
    cudaStreamWaitEvent(s_out, b.computed, 0);
    if (d_flags) {   // tensor-core launch: its guard words travel to the host, the chunk is remembered
        e = cudaMemcpyAsync(pipe->h_flags + 2 * pipe->pending.size(), d_flags, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s_out);
        if (e != cudaSuccess) return cuda_fail(e, "host pipeline: D2H guard words");
        pipe->pending.push_back(PendingChunk{wav, wav_off, wav_len, c0, c1, irfft_n, out, out_off, out_len});
    }
    e = cudaSuccess;
    if (no_copy_out) {
    } else if (out_contiguous && relay) {
        // this device -> relay device (NVLink) -> host, on the relay's copy-out stream
        cudaSetDevice(pipe->relay);
        cudaStreamWaitEvent(pipe->r_out, b.computed, 0);
        e = cudaMemcpyPeerAsync(b.r_out, pipe->relay, b.d_out, h->device, out_bytes, pipe->r_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + o_lo, b.r_out, out_bytes, cudaMemcpyDeviceToHost, pipe->r_out);
        if (e == cudaSuccess) e = cudaEventRecord(b.r_downloaded, pipe->r_out);
        cudaSetDevice(h->device);
    } else if (out_contiguous) {
        e = cudaMemcpyAsync(out + o_lo, b.d_out, out_bytes, cudaMemcpyDeviceToHost, s_out);
    } else {   // gaps between the outputs belong to the caller: copy utterance by utterance
        for (int u = c0; u < c1 && e == cudaSuccess; ++u)
            e = cudaMemcpyAsync(out + out_off[u], b.d_out + m_out_off[u - c0], (size_t)out_len[u] * sizeof(float), cudaMemcpyDeviceToHost, s_out);
    }
    cudaEventRecord(b.downloaded, s_out);
    return e == cudaSuccess ? RCED_OK : cuda_fail(e, "host pipeline: D2H waveforms");
}

static int host_sync_impl(rced_handle* h) {
    HostPipe* p = h->pipe;
    if (!p) return RCED_OK;
    for (cudaStream_t s : {p->s_in, p->s_compute, p->s_out, p->r_in, p->r_out}) {
        if (!s) continue;
        cudaError_t e = cudaStreamSynchronize(s);
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
        rc = run_chunk(h, p, p->set[0], q.wav, q.wav_off, q.wav_len, q.c0, q.c1, q.irfft_n, q.out, q.out_off, q.out_len);
        p->recomputing = false;
        h->variant = variant;
        if (rc == RCED_OK) {
            cudaError_t e = cudaStreamSynchronize(p->s_out);
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
    // chunks of about `chunk_rows` spectrogram rows; the first and the last chunk of a long synchronous call are a
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
        BufferSet& b = p->set[p->next_set++ % (unsigned int)kSets];
        rc = run_chunk(h, p, b, wav, wav_off, wav_len, bounds[i], bounds[i + 1], irfft_n, out, out_off, out_len,
                       !async && bounds.size() == 2);
        if (rc != RCED_OK) return rc;
    }
    return RCED_OK;
}

}  // namespace rced

using namespace rced;

extern "C" {

int rced_host_config(rced_handle* h, int64_t chunk_rows, int64_t chunk_rows_async) {
    if (!h) return fail(RCED_ERR_ARG, "null handle");
    if (chunk_rows < 1 || chunk_rows_async < 1) return fail(RCED_ERR_ARG, "chunk sizes must be positive");
    DeviceGuard guard(h->device);
    HostPipe* p = nullptr;
    const int rc = pipe_get(h, &p);
    if (rc != RCED_OK) return rc;
    p->chunk_rows = chunk_rows;
    p->chunk_rows_async = chunk_rows_async;
    return RCED_OK;
}

int rced_host_set_relay(rced_handle* h, int relay_device) {
    if (!h) return fail(RCED_ERR_ARG, "null handle");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return fail(RCED_ERR_CUDA, "no CUDA device");
    if (relay_device >= ndev || relay_device == h->device) return fail(RCED_ERR_ARG, "relay must be another visible device (or -1)");
    DeviceGuard guard(h->device);
    HostPipe* p = nullptr;
    int rc = pipe_get(h, &p);
    if (rc != RCED_OK) return rc;
    rc = host_sync_impl(h);   // nothing in flight while the route changes
    if (rc != RCED_OK) return rc;
    if (relay_device < 0) {
        p->relay = -1;
        return RCED_OK;
    }
    if (p->r_in && p->relay_made_on != relay_device) {   // a different relay than before: its streams, events and staging go
        for (BufferSet& b : p->set) {
            cudaFree(b.r_in);
            cudaFree(b.r_out);
            b.r_in = b.r_out = nullptr;
            b.cap_r_in = b.cap_r_out = 0;
            for (cudaEvent_t* ev : {&b.r_uploaded, &b.r_downloaded})
                if (*ev) {
                    cudaEventDestroy(*ev);
                    *ev = nullptr;
                }
        }
        cudaStreamDestroy(p->r_in);
        cudaStreamDestroy(p->r_out);
        p->r_in = p->r_out = nullptr;
    }
    int can = 0;
    cudaDeviceCanAccessPeer(&can, h->device, relay_device);
    if (!can) return fail(RCED_ERR_STATE, "the relay device is not a peer of the handle's device");
    cudaError_t e = cudaDeviceEnablePeerAccess(relay_device, 0);   // own device -> relay
    if (e == cudaErrorPeerAccessAlreadyEnabled) e = cudaGetLastError(), e = cudaSuccess;
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    cudaSetDevice(relay_device);
    e = cudaDeviceEnablePeerAccess(h->device, 0);                  // relay -> own device
    if (e == cudaErrorPeerAccessAlreadyEnabled) e = cudaGetLastError(), e = cudaSuccess;
    if (e == cudaSuccess && !p->r_in) e = cudaStreamCreateWithFlags(&p->r_in, cudaStreamNonBlocking);
    if (e == cudaSuccess && !p->r_out) e = cudaStreamCreateWithFlags(&p->r_out, cudaStreamNonBlocking);
    for (BufferSet& b : p->set)
        for (cudaEvent_t* ev : {&b.r_uploaded, &b.r_downloaded})
            if (e == cudaSuccess && !*ev) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(e, "relay set-up");
    p->relay = relay_device;
    p->relay_made_on = relay_device;
    return RCED_OK;
}

int rced_host_link_probe(int device, size_t bytes, int iters, double* h2d_gbs, double* d2h_gbs) {
    if (!h2d_gbs || !d2h_gbs || bytes == 0 || iters < 1) return fail(RCED_ERR_ARG, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(RCED_ERR_CUDA, "no such CUDA device");
    DeviceGuard guard(device);
    void *h_a = nullptr, *h_b = nullptr, *d_a = nullptr, *d_b = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, f0 = nullptr, f1 = nullptr;
    cudaError_t e = cudaHostAlloc(&h_a, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&h_b, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&d_a, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_b, bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    for (cudaEvent_t* ev : {&e0, &e1, &f0, &f1})
        if (e == cudaSuccess) e = cudaEventCreate(ev);
    if (e == cudaSuccess) {
        memset(h_a, 0, bytes);
        // both directions at once, like the pipeline in steady state; one warm-up pass
        for (int pass = 0; pass < 2 && e == cudaSuccess; ++pass) {
            cudaEventRecord(e0, s0);
            cudaEventRecord(f0, s1);
            for (int i = 0; i < iters; ++i) {
                cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s0);
                cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s1);
            }
            cudaEventRecord(e1, s0);
            cudaEventRecord(f1, s1);
            e = cudaStreamSynchronize(s0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s1);
        }
        if (e == cudaSuccess) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, e0, e1);
            cudaEventElapsedTime(&b, f0, f1);
            *h2d_gbs = (double)bytes * iters / (a * 1e-3) / 1e9;
            *d2h_gbs = (double)bytes * iters / (b * 1e-3) / 1e9;
        }
    }
    for (cudaEvent_t ev : {e0, e1, f0, f1})
        if (ev) cudaEventDestroy(ev);
    if (s0) cudaStreamDestroy(s0);
    if (s1) cudaStreamDestroy(s1);
    cudaFree(d_a);
    cudaFree(d_b);
    if (h_a) cudaFreeHost(h_a);
    if (h_b) cudaFreeHost(h_b);
    return e == cudaSuccess ? RCED_OK : cuda_fail(e, "rced_host_link_probe");
}

int rced_host_alloc(size_t bytes, int write_combined, void** out) {
    if (!out || bytes == 0) return fail(RCED_ERR_ARG, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    return e == cudaSuccess ? RCED_OK : cuda_fail(e, "rced_host_alloc");
}

int rced_host_free(void* p) {
    if (!p) return RCED_OK;
    cudaError_t e = cudaFreeHost(p);
    return e == cudaSuccess ? RCED_OK : cuda_fail(e, "rced_host_free");
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
