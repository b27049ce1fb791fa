// Library-wide device status block: ONE sticky error word + the debug record of the first bounded wait that timed out
// in any tcgen05 kernel (tc_common.cuh).  The kernels get the block's device address as an ordinary parameter; the
// fused loss and the sliding-window finalize kernels read word 0 and turn a non-zero value into NaN results, so a
// pipeline fault can never pass silently through training or inference even if the host never asks.
#include "common.cuh"

namespace {
__device__ int g_status[FCD_STATUS_INTS];
}

int* fcd_status_dev() {
    static int* ptr[64] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (ptr[dev] == nullptr) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_status) != cudaSuccess) return nullptr;
        ptr[dev] = static_cast<int*>(p);
    }
    return ptr[dev];
}

// Copies the status block of the current device to host_out[FCD_STATUS_INTS] (may be NULL), optionally clears it, and
// returns word 0.  Synchronises the device: a test / debug / end-of-epoch call, not a per-step one.
FCD_API int fcd_status(int* host_out, int clear) {
    int buf[FCD_STATUS_INTS];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(buf, g_status, sizeof(buf)) != cudaSuccess) return -1;
    if (host_out != nullptr)
        for (int i = 0; i < FCD_STATUS_INTS; ++i) host_out[i] = buf[i];
    if (clear && buf[0] != 0) {
        int zero[FCD_STATUS_INTS] = {0};
        cudaMemcpyToSymbol(g_status, zero, sizeof(zero));
    }
    return buf[0];
}

// Device address of the status block (word 0 = error word), for hosts that want to read it asynchronously.
FCD_API int fcd_status_device_ptr(void** out) {
    int* p = fcd_status_dev();
    if (out == nullptr || p == nullptr) return -1;
    *out = p;
    return 0;
}

// The five per-kernel accessors of round 1 are kept as aliases of the shared word: value (and clear) of the error word.
static int status_word_and_clear() { return fcd_status(nullptr, 1); }
FCD_API int fcd_tc_error(void) { return status_word_and_clear(); }
FCD_API int fcd_tcf_error(void) { return status_word_and_clear(); }
FCD_API int fcd_gemm_tc_error(void) { return status_word_and_clear(); }
FCD_API int fcd_wgrad_tc_error(void) { return status_word_and_clear(); }
FCD_API int fcd_wgrad_gemm_tc_error(void) { return status_word_and_clear(); }
