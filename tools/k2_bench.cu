// Timing harness for the network kernel (K2) alone: compiles csrc/rced_net.cu into this
// translation unit (so that -D experiment switches apply), fills the packed weight image with
// small random numbers and times the BASELINE configs[1] launch (1024 x 249 frames).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -I../fullycnnspeechenhancement_b200/csrc -o k2_bench k2_bench.cu && ./k2_bench [arch]
#include "../fullycnnspeechenhancement_b200/csrc/rced_net.cu"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace rced {
void count_launch(int) {}
}  // namespace rced

int main(int argc, char** argv) {
    using namespace rced;
    const int arch = argc > 1 ? atoi(argv[1]) : 2;
    const int n_utt = 1024, T = 249;
    const long long rows = (long long)n_utt * T;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int pk = arch == 1 ? pad4(packed_count(1)) : arch == 2 ? pad4(packed_count(2)) : pad4(packed_count(3));
    std::vector<float> w(pk);
    srand(1);
    for (auto& v : w) v = 0.05f * ((float)rand() / RAND_MAX - 0.5f);
    std::vector<float> x((size_t)rows * kBins);
    for (auto& v : x) v = (float)rand() / RAND_MAX;
    std::vector<long long> ro(n_utt + 1);
    for (int i = 0; i <= n_utt; ++i) ro[i] = (long long)i * T;
    float *dw, *dx, *dy;
    long long* dro;
    cudaMalloc(&dw, pk * 4);
    cudaMalloc(&dx, x.size() * 4);
    cudaMalloc(&dy, x.size() * 4);
    cudaMalloc(&dro, ro.size() * 8);
    cudaMemcpy(dw, w.data(), pk * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dro, ro.data(), ro.size() * 8, cudaMemcpyHostToDevice);
    NetParams p;
    p.packed = dw; p.in = dx; p.out = dy; p.row_off = dro; p.n_utt = n_utt; p.total_rows = rows; p.skip_scratch = nullptr; p.guard = nullptr;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        cudaError_t e = launch_net(arch, true, p, prop.multiProcessorCount, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    const double mac = arch == 1 ? (double)mac_per_frame(1, true) : arch == 2 ? (double)mac_per_frame(2, true) : (double)mac_per_frame(3, true);
    printf("%s arch %d: %.3f ms, %.2f TFLOP/s valid-tap\n", argc > 2 ? argv[2] : "", arch, best, 2.0 * mac * rows / (best * 1e-3) / 1e12);
    return 0;
}
