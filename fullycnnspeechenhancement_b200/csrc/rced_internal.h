// Internal declarations shared by the translation units of librced_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rced {

struct NetParams {
    const float* packed;        // packed weight image (global), pad4(packed_count) floats
    const float* in;            // mag  [total_rows][129]
    float* out;                 // pred [total_rows][129]
    const long long* row_off;   // [n_utt + 1]
    int n_utt;
    long long total_rows;
    float* skip_scratch;        // only used when skips are not parked in tensor memory: n_slots regions of
                                // kFramesPerCta * 512 * 32 floats and their claim words (rced_slots.cuh)
    unsigned int* slot_busy;
    int n_slots;
    // When not null: flags written by the tensor-core kernel that ran before on the same stream
    // ([0] bits of the largest |activation| it stored as FP16, [1] protocol error).  The FFMA
    // kernel then only runs (and overwrites `out`) if the FP16 range was exceeded or an error
    // was reported.
    const unsigned int* guard;
};

struct StftParams {
    const float* wav;
    const long long* wav_off;
    const int* wav_len;
    const long long* row_off;
    int n_utt;
    long long total_rows;
    float* mag;      // [rows][129]
    float2* phase;   // [rows][129] or nullptr
};

struct IstftParams {
    const float* pred;       // [rows][129]
    const float2* phase;     // [rows][129]
    const long long* row_off;
    int n_utt;
    int irfft_n;             // 512 or 256
    int chunk_segs;          // 128-sample segments per CTA
    float* out;
    const long long* out_off;
    const int* out_len;
};

cudaError_t launch_net(int arch, bool skip_in_tmem, const NetParams& p, int num_sms, cudaStream_t stream);
size_t net_smem_bytes_rt(int arch);
// tensor-core variant (rced_net_tc.cu)
int tc_image_bytes(int arch);
int tc_bias_floats(int arch);
size_t tc_skip_floats_per_cta(int arch);
int tc_smem_bytes(int arch);
void tc_pack_weights(int arch, const float* folded, unsigned char* img, float* bias);
int tc_trace_slots(int arch);
// skip: n_slots regions of tc_skip_floats_per_cta floats, slot_busy: their claim words (zero-initialised);
// persist_bytes > 0: the launch carries an L2 access-policy window (persisting) over the first persist_bytes of skip
cudaError_t launch_net_tc(int arch, const NetParams& p, const unsigned char* wimg, const float* bias, float* skip,
                          unsigned int* slot_busy, int n_slots, size_t persist_bytes, unsigned int* flags, long long* trace,
                          int num_sms, cudaStream_t stream);
cudaError_t launch_stft(const StftParams& p, cudaStream_t stream);
cudaError_t launch_istft(const IstftParams& p, long long max_rows_per_utt, cudaStream_t stream);
cudaError_t upload_tables_stft();    // twiddles + window tables of K1 -> current device
cudaError_t upload_tables_istft();   // same for K3
cudaError_t run_ffma_peak(int iters, int num_sms, double* tflops);
cudaError_t run_tmem_selftest(int* mismatches);

void count_launch(int n = 1);

}  // namespace rced
