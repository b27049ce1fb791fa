// Micro-benchmark of tcgen05.mma (kind::f16, bf16 operands from shared memory in the no-swizzle K-major layout the conv
// kernels use): cycles per instruction as a function of M, N and the number of issuing warps.  Not on the product
// path and not part of libfcd_b200.so (tools/umma_bench.py compiles it on demand) -- it exists because ncu's tensor-pipe "cycles active" counters turned out to be work counters on this part
// (DESIGN.md 3.1), so the per-instruction cost model has to be measured directly.  tools/umma_bench.py drives it.
#include "../fcd_b200/csrc/tc_common.cuh"

namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int M, int N, int iters, int nissue, int same_acc,
                                                            long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];     // 64 KB of zeros: operands
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ WaitCtx wctx;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid == 0) { wait_ctx_init(&wctx, nullptr, 9); mbar_init(smem_u32(&bar), (uint32_t)nissue); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    long long t0 = 0;
    if (warp < nissue) {
        const uint32_t idesc = umma_idesc(M, N, 0, 0);
        // A: [k/8][128 rows][8]  (LBO 2048, SBO 128) at offset 0;  B: [k/8][256 rows][8] (LBO 4096, SBO 128) at 8 KB
        const uint64_t ad = umma_desc(smem_u32(smem), 2048, 128);
        const uint64_t bd = umma_desc(smem_u32(smem) + 8192, 4096, 128);
        const uint32_t d = tmem_base + (same_acc ? 0 : warp * 128);   // own accumulator per issuing warp (N <= 128 then)
        __syncwarp();
        t0 = clock64();
        if (lane == 0) {
            for (int i = 0; i < iters; ++i) umma_f16(d, ad, bd, idesc, 1u);
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0, &wctx, 9);
    const long long t1 = clock64();
    if (warp == 0 && lane == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}
}  // namespace

// cycles[0] <- SM cycles for `iters` back-to-back tcgen05.mma (M x N x 16) from each of `nissue` warps (1..4) of CTA 0;
// `ctas` CTAs run the same loop concurrently (one per SM).  same_acc != 0: all warps accumulate into one TMEM tile.
extern "C" __attribute__((visibility("default"))) int fcd_umma_bench(int M, int N, int iters, int nissue, int same_acc, int ctas, long long* cycles,
                           cudaStream_t stream) {
    if (!(M == 64 || M == 128) || N % 16 || N < 16 || N > 256 || nissue < 1 || nissue > 4 || iters < 1 || ctas < 1)
        return -1;
    if (!same_acc && N > 128) return -1;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        configured = true;
    }
    umma_bench_kernel<<<ctas, 128, 65536, stream>>>(M, N, iters, nissue, same_acc, cycles);
    return (int)cudaGetLastError();
}

int* fcd_status_dev() { return nullptr; }   // stand-alone tool: no status block
