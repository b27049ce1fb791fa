// "Last block done" helper: the block that arrives last at the end of a kernel finishes the reduction of the partials
// all blocks wrote, instead of a second launch (the one-block finalize kernels were 111 launches / 0.8 ms of a training
// step, most of them on the latency-bound critical path of the deep levels).
//
// Tickets: a zero-initialised device array, one slot per launch taken round-robin on the host; the last block resets
// its slot, so a slot is clean again long before the ring wraps (4096 slots > the launches of a whole training step;
// CUDA-graph replays of the same launch are ordered on their stream).
#pragma once
#include <atomic>

#include "common.cuh"

namespace lastblk {

constexpr int kSlots = 4096;
static __device__ unsigned g_tickets[kSlots];

// n consecutive slots (one per output tile of a split-K launch); never wraps inside a request
static inline unsigned* next_tickets(unsigned count);
static inline unsigned* next_ticket() { return next_tickets(1u); }
static inline unsigned* next_tickets(unsigned count) {
    static std::atomic<unsigned> n{0};
    static unsigned* base[64] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (base[dev] == nullptr) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, g_tickets);
        base[dev] = static_cast<unsigned*>(p);
    }
    if (count > (unsigned)kSlots) return nullptr;
    for (;;) {
        unsigned cur = n.load();
        unsigned start = cur % kSlots;
        unsigned adv = count;
        if (start + count > (unsigned)kSlots) { adv += kSlots - start; start = 0; }    // skip the tail of the ring
        if (n.compare_exchange_weak(cur, cur + adv)) return base[dev] + start;
    }
}

// true in EVERY thread of the block that arrived last (of `total` blocks).  All global writes of all blocks made before
// the call are visible to that block afterwards (read them with __ldcg).
__device__ __forceinline__ bool arrive(unsigned* ticket, unsigned total) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == total - 1u);
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    const bool last = s_last != 0;
    if (last) __threadfence();
    return last;
}

}  // namespace lastblk
