// The handle behind the C ABI (include/rced.h), shared by rced_api.cu and rced_host.cu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

namespace rced {
struct HostPipe;   // rced_host.cu: streams, staging and workspaces of the host-buffer entry points
void host_pipe_destroy(HostPipe* p);
int fail(int code, const std::string& msg);          // records the calling thread's error text, returns code
int cuda_fail(cudaError_t e, const char* what);
}  // namespace rced

struct rced_handle;
namespace rced {
int forward_impl(rced_handle* h, const float* mag, const int64_t* row_off, int n_utt, int64_t total_rows, float* pred, void* stream,
                 unsigned int** deferred_flags);
}  // namespace rced

constexpr unsigned int kFlagRing = 4096;   // guard-flag pairs handed out round robin, one per tensor-core launch

struct rced_handle {
    int arch;
    int device;
    int num_sms;
    bool skip_in_tmem;
    float* d_packed;
    // FFMA kernel with skips in global memory (rced_set_skip_in_tmem(h, 0)): num_sms regions + claim words
    float* d_scratch;
    unsigned int* d_scratch_busy;
    // tensor-core variant (rced_net_tc.cu), allocated by rced_set_variant(h, RCED_VARIANT_TC)
    int variant;
    std::vector<float> folded;
    unsigned char* d_tc_img;
    float* d_tc_bias;
    // The tensor-core kernel parks skip tensors in a global scratch of num_sms regions; a CTA claims a
    // region when it starts (rced_slots.cuh), so launches that overlap on different streams share the
    // one scratch.  Every launch reports through its own pair of guard-flag words, taken round robin
    // from a ring (stream-ordered memset in front of the launch; a pair is reused after kFlagRing launches).
    float* d_tc_skip;
    unsigned int* d_tc_busy;
    unsigned int* d_tc_flags;
    std::atomic<unsigned int> tc_launches;
    std::atomic<unsigned int*> last_tc_flags;
    size_t tc_persist_bytes;       // > 0: launches carry an L2 access-policy window over the scratch
    const char* trace_path;        // RCED_TC_TRACE (development aid), read once
    rced::HostPipe* pipe;          // created by the first rced_enhance_host call
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

