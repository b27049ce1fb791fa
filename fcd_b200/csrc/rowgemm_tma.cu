// Pointwise contractions on LARGE volumes as a persistent TMA + tcgen05 kernel:
//   * 1x1x1 Conv3d / nn.Linear rows (conv_blocks.py:420-437 conv3 of UnetResBlock, :57 conv8, :225 qkvv;
//     ms_dsa_net.py:215 patch_embedding) and their data gradients:          C[m][n] = sum_k A[m][k] W[n][k] (+ bias[n])
//   * ConvTranspose3d k2 s2 (conv_blocks.py:640-649) forward: the same GEMM with N = 8 * Cq columns, column group j
//     scattered to the fine voxel (2z + j/4, 2y + (j/2)%2, 2x + j%2) of the (concat) output buffer;
//   * its data gradient: dX[m][ci] = sum_{tap, co} dY[2m + tap][co] W[ci][co][tap] -- eight taps, each a STRIDED TMA box
//     (elementStrides 2 on the fine grid picks every other voxel), accumulated into one TMEM tile.
// These layers move 0.2-0.5 GB for 1-4 GFLOP (10 FLOP/B): they are HBM-bound, the tensor pipe idles.  What the kernel is
// built for is bytes in flight with no per-element instructions: ONE thread per SM keeps an 8-24-stage ring of 128-row
// [rows][K] tiles filled with cp.async.bulk.tensor (swizzled K-major, the layout the UMMA descriptor names; rows beyond M
// arrive as zeros), the whole packed weight matrix sits in shared memory for the lifetime of the CTA, ONE thread issues
// K/16 tcgen05.mma per tile (and tap) into one of two TMEM accumulators, and four warps drain the other accumulator:
// tcgen05.ld -> (+bias) -> bf16 -> 16-byte stores of whole channel rows.
//
// Roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-5 epilogue (TMEM lane quarter =
// warp % 4).  Barriers: FULL/EMPTY per ring stage, TFULL/TEMPTY per accumulator, WFULL for the weights.
#include <cuda.h>      // CUtensorMap + enums only: cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int NTHREADS = 192;

struct RowGemmParams {
    bf16* C; long long ldc;
    const float* bias;
    int M, N;                 // rows, columns (N = BN)
    int mode;                 // 0 plain rows, 1 k2s2 scatter (N = 8 * Cq), 2 k2s2 data gradient (8 strided taps)
    int Cq;                   // mode 1: channels per column group
    int D, H, W;              // modes 1, 2: the COARSE grid (rows m = voxels of [B][D][H][W])
    int bw, bh, bd, bb;       // mode 2: the 128-row tile as a box of the coarse grid
    int ntiles;
    int* status;
};

template <int BK, int BN>
struct RCfg {
    static constexpr int A_BYTES = BM * BK * 2;                   // 4 / 8 / 16 KB per stage
    static constexpr int NST = BK == 64 ? 8 : (BK == 32 ? 14 : 24);     // 96-128 KB of row tiles in flight per SM
    static constexpr int W_TAP = BN * BK * 2;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators
    static constexpr uint32_t SBO = 8 * BK * 2;                   // 8 rows of one swizzle span
    static constexpr uint32_t LAYOUT = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);   // SWIZZLE_128B / 64B / 32B
    static int smem(int taps) { return NST * A_BYTES + ((taps * W_TAP + 1023) & ~1023) + 2048; }
};

template <int BK, int BN>
__global__ void __launch_bounds__(NTHREADS, 1) rowgemm_tma_kernel(const RowGemmParams p,
                                                                  const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW, int taps) {
    using K = RCfg<BK, BN>;
    constexpr int NST = K::NST;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;             // swizzled tiles: 1024-byte aligned
    const uint32_t wbase = ring + NST * K::A_BYTES;
    const uint32_t w_bytes = (uint32_t)taps * K::W_TAP;
    unsigned char* aux = smem_raw + (ring - smem_u32(smem_raw)) + NST * K::A_BYTES + ((w_bytes + 1023u) & ~1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
    // bars: [0,NST) FULL | [NST,2NST) EMPTY | [2NST, 2NST+2) TFULL | [2NST+2, 2NST+4) TEMPTY | [2NST+4] WFULL
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 5);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    auto TFULL = [&](int a) { return bar0 + 8u * (2 * NST + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * NST + 2 + a); };
    const uint32_t WFULL = bar0 + 8u * (2 * NST + 4);
    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 6);
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), 128); }
        mbar_init(WFULL, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ===================================================================== TMA producer (one thread)
        if (lane == 0) {
            mbar_expect_tx(WFULL, w_bytes);
            for (int t = 0; t < taps; ++t) tma_load_3d(wbase + t * K::W_TAP, &tmW, WFULL, 0, 0, t);
            int j = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = blockIdx.x + it * gridDim.x;
                int x0 = 0, y0 = 0, z0 = 0, b0 = 0;
                if (p.mode == 2) {                      // tile = box [b0..][z0..][y0..][0, W) of the coarse grid
                    int r = tile * BM / p.W;
                    y0 = r % p.H; r /= p.H;
                    z0 = r % p.D;
                    b0 = r / p.D;
                }
                for (int t = 0; t < taps; ++t, ++j) {
                    const int s = j % NST;
                    mbar_wait(EMPTY(s), ((j / NST) & 1u) ^ 1u, ctx, 1, j);
                    mbar_expect_tx(FULL(s), K::A_BYTES);
                    if (p.mode == 2)
                        tma_load_5d(ring + s * K::A_BYTES, &tmA, FULL(s), 0, 2 * x0 + (t & 1), 2 * y0 + ((t >> 1) & 1),
                                    2 * z0 + (t >> 2), b0);
                    else
                        tma_load_2d(ring + s * K::A_BYTES, &tmA, FULL(s), 0, tile * BM);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (one thread)
        constexpr uint32_t idesc = umma_idesc(BM, BN, 0, 0);
        constexpr uint32_t HI = ((K::SBO >> 4) & 0x3fffu) | (1u << 14) | (K::LAYOUT << 29);
        auto desc = [&](uint32_t saddr) { return ((uint64_t)HI << 32) | (((saddr >> 4) & 0x3fffu) | (1u << 16)); };
        mbar_wait(WFULL, 0, ctx, 2, -1);
        int j = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int a = it & 1;
            mbar_wait(TEMPTY(a), (((it >> 1) & 1u) ^ 1u), ctx, 3, it);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
            for (int t = 0; t < taps; ++t, ++j) {
                const int s = j % NST;
                mbar_wait(FULL(s), (j / NST) & 1u, ctx, 4, j);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk)
                        umma_f16(d_tmem, desc(ring + s * K::A_BYTES + kk * 32), desc(wbase + t * K::W_TAP + kk * 32),
                                 idesc, (t | kk) ? 1u : 0u);
                    umma_commit(EMPTY(s));
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(TFULL(a));
            __syncwarp();
        }
    } else {
        // ===================================================================== epilogue: 4 warps, thread = output row
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        for (int it = 0; it < my_tiles; ++it) {
            const int a = it & 1;
            const int tile = blockIdx.x + it * gridDim.x;
            const int m = tile * BM + row;
            mbar_wait(TFULL(a), (it >> 1) & 1u, ctx, 5, it);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
            long long fine0 = 0;
            if (p.mode == 1) {
                int mm = m;
                const int x = mm % p.W; mm /= p.W;
                const int y = mm % p.H; mm /= p.H;
                const int z = mm % p.D; mm /= p.D;
                fine0 = (((long long)mm * (2 * p.D) + 2 * z) * (2 * p.H) + 2 * y) * (2 * p.W) + 2 * x;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                tmem_wait_ld();
                if (m < p.M) {
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                    bf16* dst;
                    int nb = c0;
                    if (p.mode == 1) {
                        const int tap = c0 / p.Cq;
                        nb = c0 - tap * p.Cq;
                        const long long vox = fine0 + ((long long)(tap >> 2) * (2 * p.H) + ((tap >> 1) & 1)) * (2 * p.W) +
                                              (tap & 1);
                        dst = p.C + vox * p.ldc + nb;
                    } else {
                        dst = p.C + (long long)m * p.ldc + c0;
                    }
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + nb + i);
                    }
                    st8(dst, pack8(f));
                    st8(dst + 8, pack8(f + 8));
                }
            }
            tc_fence_before();
            mbar_arrive(TEMPTY(a));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<K::TMEM_COLS>(tmem_base);
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

bool coarse_box(int Bn, int D, int H, int W, int* bw, int* bh, int* bd, int* bb) {
    int rem = BM;
    if (W > rem || rem % W || 2 * W > 256) return false;
    *bw = W; rem /= W;
    if (rem >= H) { if (rem % H) return false; *bh = H; rem /= H; } else { if (H % rem) return false; *bh = rem; rem = 1; }
    if (rem >= D) { if (rem % D) return false; *bd = D; rem /= D; } else { if (D % rem) return false; *bd = rem; rem = 1; }
    if (rem > Bn) return false;
    *bb = rem;
    return true;
}

template <int BK, int BN>
int launch(RowGemmParams p, const void* A, long long lda, const void* Wp, int taps, int Bn, cudaStream_t stream) {
    using K = RCfg<BK, BN>;
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return -2;
    const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUtensorMap tmA, tmW;
    if (p.mode == 2) {
        // the FINE grid [B][2D][2H][2W][lda], traversed with element stride 2: one box = the 128 coarse voxels' taps
        const cuuint64_t dims[5] = {(cuuint64_t)BK, (cuuint64_t)2 * p.W, (cuuint64_t)2 * p.H, (cuuint64_t)2 * p.D,
                                    (cuuint64_t)Bn};
        const cuuint64_t row = (cuuint64_t)lda * 2;
        const cuuint64_t strides[4] = {row, row * 2 * p.W, row * 4 * p.W * p.H, row * 8 * p.W * p.H * p.D};
        const cuuint32_t bx[5] = {(cuuint32_t)BK, (cuuint32_t)(2 * p.bw), (cuuint32_t)(2 * p.bh), (cuuint32_t)(2 * p.bd),
                                  (cuuint32_t)p.bb};
        const cuuint32_t es[5] = {1, 2, 2, 2, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(A), dims, strides, bx, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    } else {
        const cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)p.M};
        const cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
        const cuuint32_t bx[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
        const cuuint32_t es[2] = {1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dims, strides, bx, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    {   // packed weights [taps][N][K] bf16
        const cuuint64_t dims[3] = {(cuuint64_t)BK, (cuuint64_t)BN, (cuuint64_t)taps};
        const cuuint64_t strides[2] = {(cuuint64_t)BK * 2, (cuuint64_t)BK * 2 * BN};
        const cuuint32_t bx[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(Wp), dims, strides, bx, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -2;
    }
    const int smem = K::smem(taps);
    static int configured = 0;
    if (configured < smem) {
        if (cudaFuncSetAttribute(rowgemm_tma_kernel<BK, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return (int)cudaGetLastError();
        configured = smem;
    }
    int grid = fcd_num_sms();
    if (grid > p.ntiles) grid = p.ntiles;
    rowgemm_tma_kernel<BK, BN><<<grid, NTHREADS, smem, stream>>>(p, tmA, tmW, taps);
    return (int)cudaGetLastError();
}

template <int BK>
int launch_bn(const RowGemmParams& p, const void* A, long long lda, const void* Wp, int taps, int Bn, cudaStream_t st) {
    switch (p.N) {
        case 16: return launch<BK, 16>(p, A, lda, Wp, taps, Bn, st);
        case 32: return launch<BK, 32>(p, A, lda, Wp, taps, Bn, st);
        case 64: return launch<BK, 64>(p, A, lda, Wp, taps, Bn, st);
        case 128: return launch<BK, 128>(p, A, lda, Wp, taps, Bn, st);
        case 256: return launch<BK, 256>(p, A, lda, Wp, taps, Bn, st);
    }
    return -1;
}

bool kn_ok(int K, int N) {
    return (K == 16 || K == 32 || K == 64) && (N == 16 || N == 32 || N == 64 || N == 128 || N == 256);
}

}  // namespace

// 1 if fcd_rowgemm takes this contraction: mode 0 / 1: M rows of K (padded) channels -> N columns; mode 2: the data
// gradient of a k2 s2 transposed conv on the coarse grid (Bn, D, H, W) with K = Cq fine channels per tap, N = coarse
// channels.  Small problems stay on the mma.sync kernel (a persistent CTA per SM needs tiles to stream).
FCD_API int fcd_rowgemm_ok(int mode, int Bn, int D, int H, int W, long long M, int K, int N) {
    if (!kn_ok(K, N) || encode_tiled_fn() == nullptr || M < 16384 || M > 0x7fffffffLL) return 0;
    if (mode == 2) {
        int bw, bh, bd, bb;
        if ((long long)Bn * D * H * W != M || !coarse_box(Bn, D, H, W, &bw, &bh, &bd, &bb)) return 0;
        if (8 * N * K * 2 > 96 * 1024) return 0;
    } else if (mode == 1) {
        if ((long long)Bn * D * H * W != M) return 0;
    } else if (mode != 0) {
        return 0;
    }
    return 1;
}

// A: bf16 rows of pitch lda (mode 2: the FINE grid [Bn][2D][2H][2W] rows, K channels used); Wp: packed bf16 [taps][N][K]
// (fcd_pack_weight layout; taps = 1, mode 2: 8); C: bf16 rows of pitch ldc (mode 1: the fine (concat) buffer, Cq channels
// per voxel written); bias: fp32 [N] (mode 1: [Cq]) or NULL.
FCD_API int fcd_rowgemm(int mode, const void* A, long long lda, const void* Wp, void* C, long long ldc, const float* bias,
                        int Bn, int D, int H, int W, long long M, int K, int N, int Cq, cudaStream_t stream) {
    if (!fcd_rowgemm_ok(mode, Bn, D, H, W, M, K, N) || lda % 8 || ldc % 8 || lda < K) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)Wp & 15) || ((uintptr_t)C & 15)) return -1;
    if (mode == 1 && (Cq < 16 || Cq % 16 || N != 8 * Cq)) return -1;
    RowGemmParams p;
    p.C = (bf16*)C; p.ldc = ldc; p.bias = bias; p.M = (int)M; p.N = N; p.mode = mode; p.Cq = Cq > 0 ? Cq : 16;
    p.D = D; p.H = H; p.W = W; p.bw = p.bh = p.bd = p.bb = 0;
    if (mode == 2 && !coarse_box(Bn, D, H, W, &p.bw, &p.bh, &p.bd, &p.bb)) return -1;
    p.ntiles = (int)((M + BM - 1) / BM);
    p.status = fcd_status_dev();
    const int taps = mode == 2 ? 8 : 1;
    if (K == 16) return launch_bn<16>(p, A, lda, Wp, taps, Bn, stream);
    if (K == 32) return launch_bn<32>(p, A, lda, Wp, taps, Bn, stream);
    return launch_bn<64>(p, A, lda, Wp, taps, Bn, stream);
}
