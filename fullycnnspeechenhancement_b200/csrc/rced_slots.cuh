// Scratch slots for the persistent network kernels.
//
// Both network kernels park skip tensors in a per-CTA region of a global scratch buffer when they do
// not fit on chip.  Launches of one handle may overlap on the GPU (several streams), so the region
// cannot be chosen by blockIdx: a CTA claims a free slot when it starts and returns it when it ends.
// At most one CTA of these kernels is resident per SM (each needs more than half of the SM's shared
// memory), so num_sms slots always suffice and the first probe -- the SM's own id -- almost always
// succeeds: the whole scratch is num_sms regions, whatever the number of streams.
#pragma once
#include <cuda_runtime.h>

namespace rced {

// returns the claimed slot or -1 (every slot stayed busy: only possible after a kernel was killed
// between claim and release; the caller reports a protocol error and the FP32 kernel recomputes)
__device__ __forceinline__ int scratch_slot_acquire(unsigned int* busy, int n) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    int slot = (int)(smid % (unsigned int)n);
    for (int tries = 0; tries < 64 * n; ++tries) {
        if (atomicCAS(busy + slot, 0u, 1u) == 0u) {
            __threadfence();   // the previous owner's stores precede its release, ours follow the claim
            return slot;
        }
        slot = slot + 1 < n ? slot + 1 : 0;
        if (tries % n == n - 1) __nanosleep(2000);
    }
    return -1;
}

// called by one thread after a CTA-wide barrier behind the last access to the region
__device__ __forceinline__ void scratch_slot_release(unsigned int* busy, int slot) {
    __threadfence();
    atomicExch(busy + slot, 0u);
}

}  // namespace rced
