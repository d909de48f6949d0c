/*
 * rced.h -- C ABI of librced_b200.so: the B200 (sm_100a) enhancement forward path of
 * phecda-xu/FullyCNNSpeechEnhancement.
 *
 * The reference has no FFI of its own (it is pure Python); the seams this ABI sits
 * behind are the Python call sites listed per function below (file:line relative to
 * the reference tree).  The Python host classes in fullycnnspeechenhancement_b200/
 * keep the reference's class / method names and call these entry points via ctypes.
 *
 * Conventions
 *  - every pointer marked DEVICE is a CUDA device pointer on the handle's device
 *    (in Python: torch.Tensor.data_ptr()); everything else is host memory;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all
 *    calls are stream-ordered and never synchronise the device;
 *  - return value 0 = success, negative = error; rced_last_error() gives the text;
 *  - ragged batches are described by ONE device array `row_off[n_utt+1]` (int64):
 *    utterance u owns spectrogram rows [row_off[u], row_off[u+1]) of the
 *    row-major [rows][129] arrays `mag`, `phase`, `pred`.  The dense padded layout
 *    of the reference ([N][T_max][129], data_utils/data_loader.py:198-209) is the
 *    special case row_off[u] = u*T_max;
 *  - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef RCED_H_
#define RCED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCED_FREQ_BINS 129   /* nfft 256 -> 129 bins   (data_utils/data_loader.py:59) */
#define RCED_FRAME_LEN 256   /* 32 ms @ 8 kHz          (Work/<recipe>/cfg/<x>.cfg)  */
#define RCED_FRAME_HOP 128   /* 16 ms @ 8 kHz                                          */

#define RCED_ARCH_V1 1       /* FullyCNNSEModel   (model_utils/model.py:6-29)  */
#define RCED_ARCH_V2 2       /* FullyCNNSEModelV2 (model_utils/model.py:32-61) */
#define RCED_ARCH_V3 3       /* FullyCNNSEModelV3 (model_utils/model.py:64-96) */

#define RCED_VARIANT_FFMA 0  /* FP32 FFMA network kernel (rced_net.cu): what a new handle runs until    */
                             /* rced_set_variant is called; exact FP32, the fall-back of the variant below */
#define RCED_VARIANT_TC   1  /* tcgen05 tensor-core kernel, FP16 x3 error-compensated split: what the   */
                             /* Python engine and bench.py select by default (3x faster, 1e-6 of float64) */

#define RCED_OK 0
#define RCED_ERR_ARG   (-1)
#define RCED_ERR_CUDA  (-2)
#define RCED_ERR_STATE (-3)

typedef struct rced_handle rced_handle;

/* ---- pure host helpers (no GPU needed) ------------------------------------------- */

/* ABI version of this header (checked by the Python loader). */
int rced_abi_version(void);

/* Last error message of the calling thread ("" if none). */
const char* rced_last_error(void);

/* Number of STFT frames for an n_samples-long signal:
 * ceil(|L-256|/128 + 1)   (data_utils/audio_feature.py:67-70, note the abs()). */
int64_t rced_num_frames(int64_t n_samples);

/* Number of floats rced_create expects in `folded` for `arch`: for every layer in
 * execution order, the BN-folded HWIO kernel [kh][kw][cin][cout] followed by the
 * folded bias [cout].  Returns -1 for an unknown arch. */
int64_t rced_folded_weight_count(int arch);

/* Number of layers of `arch`, and the shape of layer `i` (kh,kw,cin,cout). */
int rced_num_layers(int arch);
int rced_layer_shape(int arch, int layer, int* kh, int* kw, int* cin, int* cout);

/* Size (floats) of the packed shared-memory weight image for `arch`, and the packing
 * itself (host only; used by rced_create, exported so that tests can check the layout
 * the network kernel reads without a GPU). */
int64_t rced_packed_weight_count(int arch);
int rced_pack_weights(int arch, const float* folded, size_t n_folded, float* packed, size_t n_packed);

/* Layout constants of the network kernel for `arch` (development / test aid): writes
 * out[0]=stage_row, out[1]=slot_floats, out[2]=wide_floats, out[3]=shared-memory bytes,
 * out[4]=tensor-memory columns used, then per layer i: out[8+4i]=packed weight offset,
 * out[9+4i]=packed bias offset, out[10+4i]=skip column base of the slot it saves (-1),
 * out[11+4i]=skip column base of the slot it adds (-1).  `n` must be >= 8 + 4*layers. */
int rced_debug_layout(int arch, int64_t* out, int n);

/* Multiply-accumulates per frame of `arch` (valid taps only when valid_only != 0);
 * the roofline numerator of SURVEY.md section 8(d). */
int64_t rced_mac_per_frame(int arch, int valid_only);

/* ---- handle ---------------------------------------------------------------------- */

/* Replaces graph construction + checkpoint restore (model_utils/tester.py:67-83,36-39;
 * infer.py:36-52): uploads the BN-folded weights of one model to `device`.
 * `folded` is host memory laid out as described at rced_folded_weight_count(). */
int rced_create(int arch, const float* folded, size_t n_folded, int device, rced_handle** out);
void rced_destroy(rced_handle* h);
int rced_arch(const rced_handle* h);
int rced_device(const rced_handle* h);

/* FFMA kernel only.  1 (default): skip-connection tensors are parked in Tensor Memory (tcgen05.st/ld);
 * 0: they go to a scratch in global memory (L2 resident) whose regions the CTAs claim when they start,
 * so launches that overlap on several streams are safe.  Both are exercised by the parity tests. */
int rced_set_skip_in_tmem(rced_handle* h, int enable);

/* Network kernel behind rced_forward / rced_enhance (same call sites as K2 below:
 * model_utils/tester.py:85-90, infer.py:62-65).
 *   RCED_VARIANT_FFMA (a new handle's setting): register-tiled FP32 FFMA kernel.
 *   RCED_VARIANT_TC: implicit-GEMM kernel on the 5th-generation tensor cores.  Activations and
 *     weights are split into FP16 hi + lo pairs (22 significant bits) and multiplied as
 *     hi*Whi + hi*Wlo + lo*Whi with FP32 accumulation; max relative error vs the float64 oracle
 *     is stated in tests/test_gpu_tc.py.  The arithmetic is scale invariant: residuals are stored
 *     times 2^11, every step's weights and every frame's activations live in power-of-two scaled
 *     domains (csrc/rced_tc.cuh), so inputs and weights of any finite magnitude keep FP32-like
 *     accuracy.  What remains is growth INSIDE the network: the kernel records the largest
 *     |activation| it stored (scaled domain, limit 65504: 4000 x the frame's reference magnitude) and
 *     whether an input was not finite, and, stream-ordered, the FFMA kernel recomputes the call when
 *     that guard tripped.  Refused (RCED_ERR_STATE) only if a folded weight is not finite.
 *     Side effect on the device: the first switch to RCED_VARIANT_TC sets cudaLimitPersistingL2CacheSize (a device-wide
 *     limit) to hold the kernel's skip scratch, and its launches carry an access-policy window over that scratch, so that
 *     other kernels and the host copies do not push it out of the L2; RCED_TC_L2_PERSIST=0 in the environment leaves the
 *     limit alone. */
int rced_set_variant(rced_handle* h, int variant);
int rced_variant(const rced_handle* h);
/* Largest |activation| stored by the last tensor-core launch (in the frames' scaled domains; infinity:
 * overflow or a non-finite input -> the FFMA kernel recomputed the call) and its protocol-error code
 * (synchronises the device; diagnostics and tests only). */
int rced_tc_status(rced_handle* h, float* max_abs, unsigned int* protocol_error);

/* Host-side packing of the tensor-core kernel's weight image (FP16 hi/lo B-operand tiles per
 * layer, tap and 8-channel group) and FP32 bias table; exported so that tests can check the layout
 * and arithmetic without a GPU.  rced_tc_layout writes: out[0]=steps, [1]=units, [2]=image bytes,
 * [3]=shared-memory bytes, [4]=plane stride (16-byte units), [5]=lead rows, [6]=rows per frame,
 * [7]=frames per batch, [8]=row tiles, [9]=offset of the lo planes (16-byte units), [10]=taps per
 * row-shifted block of the output layer, [11]=skip scratch floats per region, [12]=row-shifted blocks,
 * [13]=zero rows in front of plane 0, [14..15]=0; then per step 6 values (units, first unit, NP, tile
 * bytes, image offset, is_final) and per unit 2 values (start offset and LBO of the A descriptor in
 * 16-byte units).  n >= 16 + 6*steps + 2*units. */
int64_t rced_tc_image_bytes(int arch);
int64_t rced_tc_bias_count(int arch);
int rced_tc_pack_weights(int arch, const float* folded, size_t n_folded, void* image, size_t image_bytes,
                         float* bias, size_t n_bias);
int rced_tc_layout(int arch, int64_t* out, int n);

/* ---- the three kernels of the path ----------------------------------------------- */

/* K1. Replaces AudioParser.parse_audio -> AudioFeature.compute_spectrogram +
 * power_spectrum + divide_phase (data_utils/data_loader.py:54-61,
 * data_utils/audio_feature.py:22-115) and the zero padding of
 * DataLoader.padding_batch (data_utils/data_loader.py:198-209).
 *   wav      DEVICE float32, all utterances concatenated
 *   wav_off  DEVICE int64[n_utt]   first sample of utterance u in `wav`
 *   wav_len  DEVICE int32[n_utt]   its length L_u (>= 1)
 *   row_off  DEVICE int64[n_utt+1] see above; rows beyond rced_num_frames(L_u)
 *                                  are written as padding (mag 0, phase 1+0j)
 *   total_rows = row_off[n_utt] (host copy, sizes the launch)
 *   mag      DEVICE float32 [total_rows][129]   linear magnitude |X|
 *   phase    DEVICE float32 [total_rows][129][2] X/|X| (re,im), may be NULL */
int rced_stft(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len,
              const int64_t* row_off, int n_utt, int64_t total_rows,
              float* mag, float* phase, void* stream);

/* K2. Replaces FullyCNNTester.test_step / sess.run(self.pred, ...)
 * (model_utils/tester.py:85-90, infer.py:62-65): pred = Model(mag), one fused kernel.
 * Frames of different utterances never interact; time taps outside an utterance's
 * rows read zeros (TF 'SAME' padding, model_utils/module.py:27).
 *   mag, pred DEVICE float32 [total_rows][129] (must not alias) */
int rced_forward(rced_handle* h, const float* mag, const int64_t* row_off, int n_utt,
                 int64_t total_rows, float* pred, void* stream);

/* K3. Replaces AudioReBuild.rebuild_audio (model_utils/utils.py:171-183):
 * pred*phase -> irfft(irfft_n)[:256] -> /hamming -> half-frame concatenation ->
 * de-emphasis -> truncate to out_len[u].  irfft_n is 512 (the shipped default,
 * model_utils/utils.py:94) or 256.
 *   out      DEVICE float32, utterance u written at out + out_off[u], out_len[u] samples
 *            (out_len[u] <= (rows_u+1)*128) */
int rced_istft(rced_handle* h, const float* pred, const float* phase, const int64_t* row_off,
               int n_utt, int64_t max_rows_per_utt, int irfft_n,
               float* out, const int64_t* out_off, const int32_t* out_len, void* stream);

/* K1 -> K2 -> K3 on one stream with caller-provided workspaces
 * (mag, pred: [total_rows][129] floats; phase: [total_rows][129][2] floats).
 * Replaces the body of FullyCNNTester.test's batch loop (model_utils/tester.py:104-113)
 * and InferenceEngine.denoise (infer.py:54-71) up to the wav write. */
int rced_enhance(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len,
                 const int64_t* row_off, int n_utt, int64_t total_rows, int64_t max_rows_per_utt,
                 int irfft_n, float* ws_mag, float* ws_phase, float* ws_pred,
                 float* out, const int64_t* out_off, const int32_t* out_len, void* stream);

/* ---- host-buffer entry points ------------------------------------------------------- */

/* The batch loop of the reference with HOST waveforms in and out: replaces
 * parse_audio -> power_spectrum / divide_phase -> sess.run(pred) -> rebuild_audio
 * (model_utils/tester.py:104-113, infer.py:54-71) for callers that hold numpy arrays.  The library owns
 * the device side: the call is cut into chunks of utterances that flow through a copy-in, a compute and
 * a copy-out stream over a ring of buffer sets (the copies of one chunk overlap the kernels of the
 * others; the kernels run in the order of the device-pointer path); device buffers belong to the
 * handle and only grow.
 *   wav      HOST float32, utterances concatenated (gaps allowed); page-locked memory makes the
 *            copies asynchronous (pageable memory works, without overlap)
 *   wav_off  HOST int64[n_utt] first sample of utterance u;  wav_len HOST int32[n_utt] (>= 1)
 *   out      HOST float32; utterance u is written to out + out_off[u], out_len[u] samples
 *            (out_len[u] <= (rced_num_frames(wav_len[u]) + 1) * 128; the reference truncates to
 *            len(clean_sig[u]), model_utils/utils.py:181-182).  Gaps of fewer than 16 samples between
 *            consecutive outputs are treated as alignment padding and may be overwritten; larger gaps
 *            are left untouched (the outputs are then copied one by one)
 * rced_enhance_host returns when `out` is complete.  rced_enhance_host_async returns once the work
 * is queued -- the buffers AND the offset / length arrays must stay valid and `out` is complete after
 * rced_host_sync(h); consecutive async calls pipeline behind each other.  With the tensor-core
 * variant the range guard is evaluated at the synchronisation: a chunk whose guard tripped is
 * recomputed with the FP32 kernel before rced_host_sync / rced_enhance_host returns.
 * rced_host_config: target spectrogram rows per chunk for the synchronous call (default 32768; the
 * first and the last chunk of a long call are a quarter of that, because nothing overlaps the first
 * upload and the last download) and for asynchronous calls (default 131072). */
int rced_enhance_host(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int n_utt,
                      int irfft_n, float* out, const int64_t* out_off, const int32_t* out_len);
int rced_enhance_host_async(rced_handle* h, const float* wav, const int64_t* wav_off, const int32_t* wav_len, int n_utt,
                            int irfft_n, float* out, const int64_t* out_off, const int32_t* out_len);
int rced_host_sync(rced_handle* h);
int rced_host_config(rced_handle* h, int64_t chunk_rows, int64_t chunk_rows_async);

/* Route the waveform copies of the host-buffer calls through a PEER GPU of this process: host -> relay device (its own
 * link to the host) -> NVLink -> the handle's device, and back the same way.  For boxes on which some GPUs reach host
 * memory through a slower or shared link than others (measure first: bench.py does).  relay_device = -1 restores the
 * direct route.  Synchronises the handle's host pipeline.  RCED_ERR_STATE if the two devices are not peers. */
int rced_host_set_relay(rced_handle* h, int relay_device);

/* Bandwidth of `device`'s link to page-locked host memory with both directions busy (GB/s each), `iters` copies of `bytes`
 * per direction.  Run on all GPUs of a box at the same time it shows which of them share a slower link. */
int rced_host_link_probe(int device, size_t bytes, int iters, double* h2d_gbs, double* d2h_gbs);

/* Page-locked host memory for the buffers above (cudaHostAlloc), for callers that have no other way to pin memory.
 * write_combined != 0: write-combined memory -- faster for the device to read, very slow for the CPU to read: for
 * INPUT buffers the host only writes. */
int rced_host_alloc(size_t bytes, int write_combined, void** out);
int rced_host_free(void* p);

/* Element-wise |X| and X/|X| of `n` complex64 values (X == 0 -> phase 1+0j).  Replaces
 * AudioFeature.power_spectrum / divide_phase (data_utils/audio_feature.py:101-115) when the
 * caller already holds a complex spectrogram (model_utils/tester.py:104-105, infer.py:57-60).
 *   X DEVICE float32[n][2]; mag DEVICE float32[n] or NULL; phase DEVICE float32[n][2] or NULL */
int rced_mag_phase(int device, const float* X, int64_t n, float* mag, float* phase, void* stream);

/* Energy sums of the SDR score of n_utt (reference, estimate) pairs: sums[u] = { sum ref^2,
 * sum (est - ref)^2 } over len[u] samples, accumulated in float64.  The score of
 * SDR.sdr (model_utils/utils.py:68-78; used at model_utils/tester.py:136-139) is
 * 10*log10(sums[u][0] / (sums[u][1] + FLT_EPSILON)).
 *   ref, est  DEVICE float32, concatenated signals;  ref_off, est_off  DEVICE int64[n_utt]
 *   len       DEVICE int32[n_utt];  max_len = max(len) (grid sizing);  sums DEVICE float64[n_utt][2] */
int rced_sdr_sums(int device, const float* ref, const int64_t* ref_off, const float* est, const int64_t* est_off,
                  const int32_t* len, int n_utt, int64_t max_len, double* sums, void* stream);

/* ---- measurement helpers --------------------------------------------------------- */

/* Dense FP32 FFMA microbenchmark (independent register chains, one CTA set per SM).
 * Writes the achieved TFLOP/s (2 flop per FFMA) to *tflops; used as the measured
 * denominator of the network kernel's roofline. */
int rced_ffma_peak(int device, int iters, double* tflops);

/* Round-trips a pattern through Tensor Memory with the same tcgen05.st/ld shapes the
 * network kernel uses; returns 0 when every value came back intact. */
int rced_selftest_tmem(int device);

/* Number of kernels this library has launched since load (all handles). */
int64_t rced_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RCED_H_ */
