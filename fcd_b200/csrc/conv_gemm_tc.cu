// 3x3x3 stride-1 pad-1 Conv3d (forward and data gradient) for the DEEP levels -- 64..512 channels on 32^3 .. 4^3 voxels --
// as a split-K GEMM on tcgen05.mma / TMEM (reference conv_blocks.py:393-416 at the encoder levels 3-6, the decoder
// blocks above them and TransformerBlock.conv51, conv_blocks.py:56).
//
//   C[m][n] = sum_tap sum_k A[src(m, tap)][k] * Wp[tap][n][k]         m = output voxel, src = tap-shifted voxel (or zero)
//
// Why a second kernel: conv_tc(f).cu keeps all 27 weight taps in shared memory and marches down long columns -- right for
// 16-64 channels on 64^3..128^3 voxels, impossible for 256x256 weights and pointless on a 4^3 volume.  Here the weights
// stream through the pipeline, the tile is 128 output voxels x BN (64..256) output channels, and the (tap, 64-channel
// chunk) loop is split over gridDim.z so that 4^3 x batch 2 = ONE voxel tile still fills the GPU; fp32 partials are
// added in a fixed order by fcd_splitk_reduce.  With N >= 128 an M=128 instruction carries 64-128 cycles of math for
// the same operand fetch that starves the N=16 layers, so this kernel is the one that runs near the tensor roofline.
//
// Two feeds for the same MMA / drain code:
//  * TMA (conv_gemm_tma_kernel, the default whenever 128 consecutive output voxels form a box of the volume, i.e. W, H, D
//    powers of two as on the 128^3 patch): ONE thread per CTA issues, per pipeline stage, one cp.async.bulk.tensor.5d of
//    the tap-shifted [b][d][h][w][64 ch] HALO TILE of the NDHWC activations -- coordinates outside the volume arrive as
//    zeros, which is the conv's zero padding -- and one .3d of the [BN][64] weight tile, both 128B-swizzled K-major
//    (the layout the UMMA descriptor names), completing on the stage's mbarrier by transaction bytes;
//  * cp.async (conv_gemm_tc_kernel: ragged volumes such as the 20^3 / 10^3 levels of the 160^3 patch): warps 0-3 gather
//    the A rows with zero-fill and copy the B rows 16 B at a time into the UMMA no-swizzle K-major layout [k/8][row][8].
// Warp 4 issues the MMAs, then warps 0-3 drain TMEM.
#include <cuda.h>      // CUtensorMap + enums only: cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint

#include "last_block.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int NPRODW = 4;
constexpr int NTHREADS = 32 * (NPRODW + 1);
constexpr int DEPTH = 2;                  // cp.async groups in flight per producer lane

struct GemmTcParams {
    const bf16* A; long long lda;
    const bf16* Wp;                       // [27][N][K] bf16 (fcd_pack_weight layout)
    bf16* C; long long ldc;               // ksplit == 1
    float* ws;                            // ksplit > 1: [ksplit][M][N] fp32
    int Bn, D, H, W, K, N, M, mode, ksplit;
    int* status;
    unsigned* tickets;                    // ksplit > 1: one slot per output tile (in-kernel reduction by the last CTA)
};

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int NST = BN >= 256 ? 4 : 5;
    static constexpr int LBO_A = BM * 16, LBO_B = BN * 16, SBO = 128;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr int SMEM = NST * STAGE + 1024;
};

// drain: warps 0-3, thread = output row; then (split-K) the CTA that wrote the LAST partial of the tile reduces
template <int BN, int TMEM_COLS>
__device__ __forceinline__ void drain_tile(const GemmTcParams& p, uint32_t DONE, WaitCtx* ctx, uint32_t tmem_base, int m0,
                                           int n0, int warp, int lane, int tid) {
    mbar_wait(DONE, 0, ctx, 3);
    tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const int m = m0 + row;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(trow + c0, v);
            tmem_wait_ld();
            if (m < p.M) {
                if (p.ksplit > 1) {
                    float4* dst = reinterpret_cast<float4*>(p.ws + ((long long)blockIdx.z * p.M + m) * p.N + n0 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                             __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                } else {
                    float f[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
                    bf16* dst = p.C + (long long)m * p.ldc + n0 + c0;
                    st8(dst, pack8(f));
                    st8(dst + 8, pack8(f + 8));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPRODW) tmem_dealloc<TMEM_COLS>(tmem_base);
    // split-K: the CTA that wrote the LAST partial of this output tile sums the ksplit partials in the fixed order
    // z = 0, 1, ... (deterministic) and writes the bf16 rows -- no fcd_splitk_reduce launch
    if (p.ksplit > 1 && p.tickets != nullptr &&
        lastblk::arrive(p.tickets + (blockIdx.y * gridDim.x + blockIdx.x), (unsigned)p.ksplit)) {
        constexpr int CH = BM * (BN / 8);
        for (int idx = tid; idx < CH; idx += NTHREADS) {
            const int row = idx / (BN / 8), c8 = idx % (BN / 8);
            const int m = m0 + row, n = n0 + c8 * 8;
            if (m >= p.M) continue;
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = 0.f;
            for (int z = 0; z < p.ksplit; ++z) {
                const float4* src = reinterpret_cast<const float4*>(p.ws + ((long long)z * p.M + m) * p.N + n);
                const float4 u = __ldcg(src), v = __ldcg(src + 1);
                a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += v.x; a[5] += v.y; a[6] += v.z; a[7] += v.w;
            }
            st8(p.C + (long long)m * p.ldc + n, pack8(a));
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) conv_gemm_tc_kernel(const GemmTcParams p) {
    using K = Cfg<BN>;
    constexpr int NST = K::NST;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * K::STAGE);
    // bars: [0,NST) FULL | [NST,2NST) EMPTY | [2NST] DONE
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t DONE = bar0 + 8u * (2 * NST);
    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 3);
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 32 * NPRODW); mbar_init(EMPTY(s), 1); }
        mbar_init(DONE, 1);
        fence_barrier_init();
    }
    if (warp == NPRODW) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kchunks = p.K / BK;
    const int nk_all = 27 * kchunks;
    const int it0 = (int)((long long)nk_all * blockIdx.z / p.ksplit);
    const int nk = (int)((long long)nk_all * (blockIdx.z + 1) / p.ksplit) - it0;

    if (warp < NPRODW) {
        // ===================================================================== producers
        // A: thread t owns output row m0 + t: eight 16 B pieces (64 channels) per stage, contiguous in global memory
        const int row = tid;                                  // 0..127
        const int m = m0 + row;
        int x = m % p.W, r1 = m / p.W;
        int y = r1 % p.H; r1 /= p.H;
        int z = r1 % p.D;
        const int b = r1 / p.D;
        const bool row_ok = m < p.M;
        const uint32_t smem_u = smem_u32(smem);
        uint32_t signaled = 0;
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL(signaled % NST)); ++signaled; }
        };
        for (int i = 0; i < nk; ++i) {
            const int it = it0 + i;
            const int tap = it / kchunks, kc = it - tap * kchunks;
            const int s = i % NST;
            mbar_wait(EMPTY(s), ((i / NST) & 1u) ^ 1u, ctx, 1);
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            // forward: src = m + tap - 1; data gradient (stride 1): src = m + 1 - tap
            const int sz = p.mode == 0 ? z + kd - 1 : z + 1 - kd;
            const int sy = p.mode == 0 ? y + kh - 1 : y + 1 - kh;
            const int sx = p.mode == 0 ? x + kw - 1 : x + 1 - kw;
            const bool ok = row_ok && sz >= 0 && sz < p.D && sy >= 0 && sy < p.H && sx >= 0 && sx < p.W;
            const bf16* src = ok ? p.A + ((((long long)b * p.D + sz) * p.H + sy) * p.W + sx) * p.lda + kc * BK : p.A;
            const uint32_t sa = smem_u + s * K::STAGE + row * 16;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) cp_async16(sa + c8 * K::LBO_A, src + c8 * 8, ok);
            // B: BN rows x 8 pieces, 128 threads
            const bf16* wsrc = p.Wp + ((long long)tap * p.N + n0) * p.K + kc * BK;
            const uint32_t sb = smem_u + s * K::STAGE + K::A_BYTES;
#pragma unroll
            for (int j = 0; j < BN * 8 / 128; ++j) {
                const int idx = tid + j * 128;
                const int n = idx >> 3, c8 = idx & 7;
                cp_async16(sb + c8 * K::LBO_B + n * 16, wsrc + (long long)n * p.K + c8 * 8, true);
            }
            cp_async_commit();
            if (i + 1 >= DEPTH) {
                cp_async_wait<DEPTH - 1>();
                fence_proxy_async();
                flush_to(i + 2 - DEPTH);
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        flush_to(nk);
    } else {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = umma_idesc(BM, BN, 0, 0);
        constexpr uint32_t HI = ((K::SBO >> 4) & 0x3fffu) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_A >> 4) << 16);
        const uint32_t b_lo0 = (((smem_u32(smem) + K::A_BYTES) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_B >> 4) << 16);
        for (int i = 0; i < nk; ++i) {
            const int s = i % NST;
            mbar_wait(FULL(s), (i / NST) & 1u, ctx, 2);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint32_t a_lo = a_lo0 + ((s * K::STAGE + kk * 2 * K::LBO_A) >> 4);
                    const uint32_t b_lo = b_lo0 + ((s * K::STAGE + kk * 2 * K::LBO_B) >> 4);
                    umma_f16(tmem_base, ((uint64_t)HI << 32) | a_lo, ((uint64_t)HI << 32) | b_lo, idesc,
                             (i | kk) ? 1u : 0u);
                }
                umma_commit(EMPTY(s));
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(DONE);
        __syncwarp();
    }

    drain_tile<BN, K::TMEM_COLS>(p, DONE, ctx, tmem_base, m0, n0, warp, lane, tid);
}

// ------------------------------------------------------------------------------------------------ TMA-fed variant
struct TmaBox { int bw, bh, bd, bb; };

template <int BN>
struct TCfg {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;   // multiples of 1024
    static constexpr int NST = BN >= 256 ? 4 : 6;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr int SMEM = NST * STAGE + 1024 /* alignment slack */ + 1024 /* barriers, TMEM slot, wait context */;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) conv_gemm_tma_kernel(const GemmTcParams p,
                                                                    const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const TmaBox box) {
    using K = TCfg<BN>;
    constexpr int NST = K::NST;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* aux = smem_raw + (tiles - smem_u32(smem_raw)) + NST * K::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
    // bars: [0,NST) FULL | [NST,2NST) EMPTY | [2NST] DONE
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t DONE = bar0 + 8u * (2 * NST);
    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 3);
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(DONE, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == NPRODW) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kchunks = p.K / BK;
    const int nk_all = 27 * kchunks;
    const int it0 = (int)((long long)nk_all * blockIdx.z / p.ksplit);
    const int nk = (int)((long long)nk_all * (blockIdx.z + 1) / p.ksplit) - it0;

    if (warp == 0) {
        // ===================================================================== producer: one thread, two TMAs per stage
        if (lane == 0) {
            // the tile is the box [b0, b0+bb) x [z0, z0+bd) x [y0, y0+bh) x [0, W) of the volume (m0 is box-aligned)
            int r = m0 / p.W;
            const int y0 = r % p.H; r /= p.H;
            const int z0 = r % p.D;
            const int b0 = r / p.D;
            const uint32_t a_bytes = (uint32_t)(box.bw * box.bh * box.bd * box.bb) * (BK * 2);
            for (int i = 0; i < nk; ++i) {
                const int it = it0 + i;
                const int tap = it / kchunks, kc = it - tap * kchunks;
                const int s = i % NST;
                mbar_wait(EMPTY(s), ((i / NST) & 1u) ^ 1u, ctx, 1, i);
                const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                // forward: src = m + tap - 1; data gradient (stride 1): src = m + 1 - tap
                const int dz = p.mode == 0 ? kd - 1 : 1 - kd;
                const int dy = p.mode == 0 ? kh - 1 : 1 - kh;
                const int dx = p.mode == 0 ? kw - 1 : 1 - kw;
                mbar_expect_tx(FULL(s), a_bytes + K::B_BYTES);
                tma_load_5d(tiles + s * K::STAGE, &tmA, FULL(s), kc * BK, dx, y0 + dy, z0 + dz, b0);
                tma_load_3d(tiles + s * K::STAGE + K::A_BYTES, &tmB, FULL(s), kc * BK, n0, tap);
            }
        }
        __syncwarp();
    } else if (warp == NPRODW) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = umma_idesc(BM, BN, 0, 0);
        const uint64_t a_d0 = umma_desc_sw128(tiles), b_d0 = umma_desc_sw128(tiles + K::A_BYTES);
        for (int i = 0; i < nk; ++i) {
            const int s = i % NST;
            mbar_wait(FULL(s), (i / NST) & 1u, ctx, 2, i);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    const uint32_t off = (uint32_t)(s * K::STAGE + kk * 32);
                    umma_f16(tmem_base, umma_desc_add(a_d0, off), umma_desc_add(b_d0, off), idesc, (i | kk) ? 1u : 0u);
                }
                umma_commit(EMPTY(s));
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(DONE);
        __syncwarp();
    }
    drain_tile<BN, K::TMEM_COLS>(p, DONE, ctx, tmem_base, m0, n0, warp, lane, tid);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

int g_use_tma = 1;

// 128 consecutive output voxels = the box [bb][bd][bh][W] of the [B][D][H][W] volume?
bool tma_box(int Bn, int D, int H, int W, TmaBox* box) {
    int rem = BM;
    if (W > rem || rem % W) return false;
    box->bw = W; rem /= W;
    if (rem >= H) { if (rem % H) return false; box->bh = H; rem /= H; } else { if (H % rem) return false; box->bh = rem; rem = 1; }
    if (rem >= D) { if (rem % D) return false; box->bd = D; rem /= D; } else { if (D % rem) return false; box->bd = rem; rem = 1; }
    if (rem > Bn) return false;
    box->bb = rem;
    return true;
}

template <int BN>
int launch_tma(const GemmTcParams& p, const TmaBox& box, cudaStream_t stream) {
    using K = TCfg<BN>;
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return -2;
    CUtensorMap tmA, tmB;
    {   // activations: [B][D][H][W][lda] bf16, innermost first; box = 64 channels x the voxel box
        const cuuint64_t dims[5] = {(cuuint64_t)p.K, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.Bn};
        const cuuint64_t row = (cuuint64_t)p.lda * 2;
        const cuuint64_t strides[4] = {row, row * p.W, row * p.W * p.H, row * p.W * p.H * p.D};
        const cuuint32_t bx[5] = {(cuuint32_t)BK, (cuuint32_t)box.bw, (cuuint32_t)box.bh, (cuuint32_t)box.bd,
                                  (cuuint32_t)box.bb};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(p.A), dims, strides, bx, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    {   // packed weights: [27][N][K] bf16
        const cuuint64_t dims[3] = {(cuuint64_t)p.K, (cuuint64_t)p.N, 27};
        const cuuint64_t strides[2] = {(cuuint64_t)p.K * 2, (cuuint64_t)p.K * 2 * p.N};
        const cuuint32_t bx[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(p.Wp), dims, strides, bx, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_gemm_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, p.N / BN, p.ksplit);
    conv_gemm_tma_kernel<BN><<<grid, NTHREADS, K::SMEM, stream>>>(p, tmA, tmB, box);
    return (int)cudaGetLastError();
}

template <int BN>
int launch(const GemmTcParams& p, cudaStream_t stream) {
    using K = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, p.N / BN, p.ksplit);
    conv_gemm_tc_kernel<BN><<<grid, NTHREADS, K::SMEM, stream>>>(p);
    return (int)cudaGetLastError();
}

int pick_bn(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 0)); }

}  // namespace

namespace {
// TMA feed: the N tile shrinks (256 -> 128 -> 64) while the grid cannot fill the GPU even at the largest useful split --
// on the 4^3 level (ONE 128-voxel tile) 512 output channels then run as 8 tiles x 27 splits instead of 2 x 27 CTAs
int pick_bn_tma(long long M, int K, int N) {
    int bn = pick_bn(N);
    const long long mt = (M + BM - 1) / BM;
    const int nk = 27 * (K / BK);
    const int ks_cap = nk / 8 > 0 ? nk / 8 : 1;
    while (bn > 64 && mt * (N / bn) * ks_cap < fcd_num_sms()) bn /= 2;
    return bn;
}
int ksplit_for(long long M, int K, int N, int bn) {
    const long long tiles = ((M + BM - 1) / BM) * (N / bn);
    const int nk = 27 * (K / BK);
    long long ks = (2LL * fcd_num_sms() + tiles - 1) / tiles;
    if (ks > nk / 8) ks = nk / 8;         // >= 8 pipeline iterations per CTA: below that the prologue/drain and the
                                          // fp32 partial traffic cost more than the extra CTAs bring (measured)
    if (ks < 1) ks = 1;
    while (ks > 1 && ks * M * N * 4 > (256LL << 20)) --ks;
    return (int)ks;
}
}  // namespace

// Split factor for fcd_conv_gemm_tc on (M voxels, K in, N out channels) with the cp.async feed; 0 = shape not taken
// (K % 64, N % 64).
FCD_API int fcd_conv_gemm_tc_ksplit(long long M, int K, int N) {
    if (M < 1 || K % BK || pick_bn(N) == 0) return 0;
    // Measured against the split-K mma.sync kernel (tools/time_deep_conv.py): this kernel wins from ~1k output voxels
    // and >= 128 input channels on (1.2-2x at 16^3 / 8^3 and at inference batch sizes); on the 4^3 level (M = 128) and
    // with 64 input channels the gathered A tile per tap makes it producer-bound and it only ties or loses.
    if (M < 1024 || K < 128) return 0;
    return ksplit_for(M, K, N, pick_bn(N));
}

// The same for a concrete volume: where the TMA feed applies (fcd_conv_gemm_tc_tma_ok) the producer limit is gone
// (tools/time_gemm_feeds.py: 1.6-3x the cp.async feed, bit-identical), so 64 input channels and the single-tile 4^3 level
// are taken too.
FCD_API int fcd_conv_gemm_tc_ksplit_vol(int Bn, int D, int H, int W, int K, int N) {
    const long long M = (long long)Bn * D * H * W;
    if (!fcd_conv_gemm_tc_tma_ok(Bn, D, H, W)) return fcd_conv_gemm_tc_ksplit(M, K, N);
    if (M < BM || K % BK || pick_bn(N) == 0) return 0;
    return ksplit_for(M, K, N, pick_bn_tma(M, K, N));
}

// A: NDHWC bf16 rows (pitch lda >= K); Wp: packed bf16 [27][N][K]; mode 0 forward / 1 data gradient (Wp then holds the
// transposed weights, as for fcd_igemm mode 1).  ksplit == 1: bf16 rows into C (pitch ldc); ksplit > 1: fp32 partials
// into ws[ksplit][M][N]: for ksplit == 2 summed into the bf16 rows of C by the CTA that finishes an output tile last,
// for ksplit > 2 to be finished by fcd_splitk_reduce.
FCD_API int fcd_conv_gemm_tc(const void* A, long long lda, const void* Wp, void* C, long long ldc, float* ws, int Bn,
                             int D, int H, int W, int K, int N, int mode, int ksplit, cudaStream_t stream) {
    const int bn = pick_bn(N);
    if (bn == 0 || K % BK || lda % 8 || ldc % 8 || ksplit < 1 || (ksplit > 1 && ws == nullptr)) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)Wp & 15) || ((uintptr_t)C & 15)) return -1;
    GemmTcParams p;
    p.A = (const bf16*)A; p.lda = lda; p.Wp = (const bf16*)Wp; p.C = (bf16*)C; p.ldc = ldc; p.ws = ws;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W; p.K = K; p.N = N; p.mode = mode; p.ksplit = ksplit;
    const long long M = (long long)Bn * D * H * W;
    if (M > 0x7fffffffLL) return -1;
    p.M = (int)M; p.status = fcd_status_dev();
    // in-kernel reduction only for ksplit == 2 (a single CTA summing many partial tiles is a serial tail: measured
    // 1.8x slower kernels at ksplit 8-27); otherwise the caller launches fcd_splitk_reduce
    p.tickets = ksplit == 2 ? lastblk::next_tickets((unsigned)(((M + BM - 1) / BM) * (N / bn))) : nullptr;
    TmaBox box;
    if (g_use_tma && tma_box(Bn, D, H, W, &box)) {
        const int bt = pick_bn_tma(M, K, N);
        if (ksplit == 2) p.tickets = lastblk::next_tickets((unsigned)(((M + BM - 1) / BM) * (N / bt)));
        const int rc = bt == 256 ? launch_tma<256>(p, box, stream)
                                 : (bt == 128 ? launch_tma<128>(p, box, stream) : launch_tma<64>(p, box, stream));
        if (rc != -2) return rc;              // -2: no tensor-map encoder in this driver -> the cp.async feed
        if (ksplit == 2) p.tickets = lastblk::next_tickets((unsigned)(((M + BM - 1) / BM) * (N / bn)));
    }
    if (bn == 256) return launch<256>(p, stream);
    if (bn == 128) return launch<128>(p, stream);
    return launch<64>(p, stream);
}

// 1 (default): volumes whose 128-voxel tiles are boxes take the TMA-fed kernel; 0: always the cp.async feed (A/B timing,
// tests of both feeds).  Returns the previous setting.
FCD_API int fcd_conv_gemm_tc_use_tma(int on) {
    const int prev = g_use_tma;
    if (on >= 0) g_use_tma = on ? 1 : 0;
    return prev;
}
// 1 if fcd_conv_gemm_tc would feed this volume by TMA
FCD_API int fcd_conv_gemm_tc_tma_ok(int Bn, int D, int H, int W) {
    TmaBox box;
    return (g_use_tma && encode_tiled_fn() != nullptr && tma_box(Bn, D, H, W, &box)) ? 1 : 0;
}
