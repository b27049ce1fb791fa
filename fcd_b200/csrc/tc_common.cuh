// Blackwell (sm_100a) building blocks shared by the tcgen05 kernels: mbarrier, TMA tensor-map loads, TMEM allocation /
// loads / stores, UMMA shared-memory + instruction descriptors, tcgen05.mma / commit.
//
// Every wait is BOUNDED: a pipeline bug must not hang the GPU.  A wait that exceeds kWaitCycles sets the CTA-wide
// `dead` flag (all later waits fall through at once) and the library-wide status block (csrc/status.cu): the sticky
// error word the loss / sliding-window finalize kernels turn into NaN, plus a debug record naming the wait
// (fcd_status).
#pragma once
#include "common.cuh"

namespace tc {

constexpr long long kWaitCycles = 400000000LL;   // ~0.2 s at 1.9 GHz; a healthy wait here is < 100 us

// Per-CTA wait context (shared memory).  `dead`: a wait of this CTA timed out, all later waits fall through.
// `status`: the library-wide device status block (csrc/status.cu, fcd_status): word 0 is the sticky error word
// (kernel id << 24 | wait site << 16 | CTA), words 1.. the debug record of the FIRST timed-out wait, including a copy of
// `prog` -- progress counters every role of the CTA keeps up to date, so the record shows where each role stood.
constexpr int kProgInts = 32;
struct WaitCtx {
    int dead;
    int kernel_id;
    int* status;
    int prog[kProgInts];
};
__device__ __forceinline__ void wait_ctx_init(WaitCtx* c, int* status, int kernel_id) {   // one thread, before the CTA sync
    c->dead = 0;
    c->kernel_id = kernel_id;
    c->status = status;
    for (int i = 0; i < kProgInts; ++i) c->prog[i] = -1;
}
// progress words freeze once a wait of the CTA has timed out (the other roles then fall through their waits and run
// ahead: the record must show where they stood, not where they got to afterwards)
__device__ __forceinline__ void prog_set(WaitCtx* c, int i, int v) {
    if (reinterpret_cast<volatile int*>(&c->dead)[0] == 0) reinterpret_cast<volatile int*>(c->prog)[i] = v;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait.  `code` identifies the call site in the error word; `info` is free-form (item / counter) for the record.
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, WaitCtx* ctx, int code, int info) {
    reinterpret_cast<volatile int*>(&ctx->dead)[0] = 1;
    int* st = ctx->status;
    if (st == nullptr) return;
    const int word = (ctx->kernel_id << 24) | (code << 16) | (int)(blockIdx.x & 0xffff);
    if (atomicCAS(st, 0, word) == 0) {
        st[1] = ctx->kernel_id; st[2] = code; st[3] = (int)blockIdx.x; st[4] = (int)threadIdx.x;
        st[5] = (int)bar; st[6] = (int)parity; st[7] = info; st[8] = (int)gridDim.x; st[9] = (int)blockIdx.y;
        for (int i = 0; i < kProgInts; ++i) st[16 + i] = reinterpret_cast<volatile int*>(ctx->prog)[i];
        __threadfence();
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, WaitCtx* ctx, int code, int info = 0) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (reinterpret_cast<volatile int*>(&ctx->dead)[0]) return;
        if (clock64() - t0 > kWaitCycles) {
            mbar_timeout(bar, parity, ctx, code, info);
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------------- TMA (tensor maps)
// cp.async.bulk.tensor: ONE thread moves a whole box of a tiled tensor map into shared memory; out-of-bounds elements
// (negative or too-large coordinates: the halo of a conv tile) arrive as zeros and still count towards the transaction
// bytes the mbarrier expects.
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------------------------------------- TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {    // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), columns [col, col+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// store 16 consecutive fp32 columns of this thread's TMEM lane (used to zero accumulators)
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr),
        "r"(z)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts (cute/atom/mma_traits_sm100.hpp):
//   K-major :  ((8,m),(8,2)) : ((16 B, SBO),(2 B, LBO))   core matrix = 8 rows x 16 B, contiguous 128 B
//   MN-major:  ((8,m),(8,k)) : ((2 B, SBO),(16 B, LBO))   core matrix = 8 k-rows x 16 B of 8 contiguous MN elements
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type (0 = none)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
// K-major operand tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 64 bf16 (128 B), 8-row groups 1024 B apart
// (SBO), 16-byte chunks XOR-swizzled by the row index inside a group; the tile base must be 1024-byte aligned.  LBO is
// not used by swizzled K-major layouts (1 by convention); a 16-element K step advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (1u << 16);
    const uint32_t hi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// advance the start address of a descriptor by `bytes` (multiple of 16; no carry out of the 14-bit field allowed)
__device__ __forceinline__ uint64_t umma_desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute/arch/mma_sm100_desc.hpp InstrDescriptor):
// c_format F32 (bit 4) | a_format BF16 (bit 7) | b_format BF16 (bit 10) | a_major bit 15 | b_major bit 16 |
// N>>3 at [17,23) | M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Warp-convergent variants: EVERY lane executes the call (so the descriptor arithmetic stays warp-uniform and ptxas
// keeps it on the uniform datapath -- UTCHMMA reads its descriptors from uniform registers), the instruction itself
// is predicated on `leader` (non-zero in exactly one lane).
__device__ __forceinline__ void umma_f16_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint32_t bar, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
        "r"(leader)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all tcgen05.mma issued so far by this thread complete -> one arrive on `bar` (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

}  // namespace tc
