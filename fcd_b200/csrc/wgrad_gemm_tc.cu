// Weight gradient of the 3x3x3 stride-1 pad-1 convs of the DEEP levels (>= 64 channels on <= 32^3 voxels) as a batch of
// tcgen05 GEMMs -- autograd of the nn.Conv3d modules of conv_blocks.py:393-416 at encoder levels 3-6, the decoder blocks
// above them and TransformerBlock.conv51 (conv_blocks.py:56).
//
//   dW[tap][n][k] = sum_m dY[m][n] * X[src(m, tap)][k]                 m = output voxel, src = tap-shifted voxel (or zero)
//
// One CTA owns (tap, 128 input channels, BN <= 256 output channels, a slice of the voxels): D[k][n] += A^T B with the
// voxels as the GEMM K dimension, 16 per tcgen05.mma.  Both operands are staged by cp.async as [ch/8][voxel][8 ch] --
// X rows gathered with the tap shift (zero-fill outside the volume), dY rows straight -- which is the UMMA no-swizzle
// MN-major canonical layout for both (8 channels = 16 contiguous bytes, next voxel 16 B further), so nothing is ever
// transposed.  M = 128 x N = BN instructions run at the dense rate for BN >= 128 (measured cost model):
// this is where the wide layers differ from fcd_wgrad3_tc, which folds kd into M for 16-32 channel layers.
// Partials part[msplit][27][Np][Kp] are added by fcd_wgrad_reduce in a fixed order.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BKV = 64;                   // voxels per pipeline stage (4 instructions of K = 16)
constexpr int BMK = 128;                  // input-channel tile = UMMA M
constexpr int NPRODW = 4, NTHREADS = 32 * (NPRODW + 1), DEPTH = 2;

struct WgemmParams {
    const bf16* X; long long ldx;
    const bf16* dY; long long ldy;
    float* part;
    int Bn, D, H, W, Kp, Np, M, ktiles, ntiles, msplit, mper;
    int* status;
};

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = BKV * BMK * 2, B_BYTES = BKV * BN * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int NST = BN >= 256 ? 4 : 5;
    static constexpr int SBO = BKV * 16, LBO = 128;           // channel-octet stride, 8-voxel-group stride
    static constexpr int SMEM = NST * STAGE + 1024;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_gemm_tc_kernel(const WgemmParams p) {
    using K = Cfg<BN>;
    constexpr int NST = K::NST;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * K::STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t DONE = bar0 + 8u * (2 * NST);
    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 5);
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 32 * NPRODW); mbar_init(EMPTY(s), 1); }
        mbar_init(DONE, 1);
        fence_barrier_init();
    }
    if (warp == NPRODW) tmem_alloc<(BN < 32 ? 32 : BN)>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int job = blockIdx.x;
    const int nt = job % p.ntiles; job /= p.ntiles;
    const int kt = job % p.ktiles; job /= p.ktiles;
    const int tap = job;                                       // 0..26
    const int m_begin = blockIdx.y * p.mper;
    const int m_end = min(p.M, m_begin + p.mper);
    const int nstage = (m_end - m_begin + BKV - 1) / BKV;      // >= 1 by construction of mper / msplit
    const int kch = min(BMK, p.Kp - kt * BMK);                 // real channels of this k tile (64 or 128)

    if (warp < NPRODW) {
        // ===================================================================== producers: thread = (voxel row, half)
        const int r = tid & 63, half = tid >> 6;
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const uint32_t smem_u = smem_u32(smem);
        uint32_t signaled = 0;
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL(signaled % NST)); ++signaled; }
        };
        for (int i = 0; i < nstage; ++i) {
            const int s = i % NST;
            mbar_wait(EMPTY(s), ((i / NST) & 1u) ^ 1u, ctx, 1);
            const int m = m_begin + i * BKV + r;
            const bool row_ok = m < m_end;
            int x = m % p.W, q = m / p.W;
            int y = q % p.H; q /= p.H;
            int z = q % p.D;
            const int b = q / p.D;
            const int sz = z + kd - 1, sy = y + kh - 1, sx = x + kw - 1;
            const bool src_ok = row_ok && sz >= 0 && sz < p.D && sy >= 0 && sy < p.H && sx >= 0 && sx < p.W;
            const bf16* xs = src_ok ? p.X + ((((long long)b * p.D + sz) * p.H + sy) * p.W + sx) * p.ldx + kt * BMK : p.X;
            const uint32_t sa = smem_u + s * K::STAGE + r * 16;
#pragma unroll
            for (int j = 0; j < BMK / 16; ++j) {               // this half's channel octets of the X row
                const int c8 = half * (BMK / 16) + j;
                const bool ok = src_ok && c8 * 8 < kch;
                cp_async16(sa + c8 * K::SBO, ok ? xs + c8 * 8 : p.X, ok);
            }
            const bf16* ys = p.dY + (long long)(row_ok ? m : 0) * p.ldy + nt * BN;
            const uint32_t sb = smem_u + s * K::STAGE + K::A_BYTES + r * 16;
#pragma unroll
            for (int j = 0; j < BN / 16; ++j) {
                const int c8 = half * (BN / 16) + j;
                cp_async16(sb + c8 * K::SBO, ys + c8 * 8, row_ok);
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
        flush_to(nstage);
    } else {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = umma_idesc(BMK, BN, 1, 1);                 // both operands MN-major
        constexpr uint32_t HI = ((K::SBO >> 4) & 0x3fffu) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO >> 4) << 16);
        const uint32_t b_lo0 = (((smem_u32(smem) + K::A_BYTES) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO >> 4) << 16);
        for (int i = 0; i < nstage; ++i) {
            const int s = i % NST;
            mbar_wait(FULL(s), (i / NST) & 1u, ctx, 2);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kk = 0; kk < BKV / 16; ++kk) {            // 16 voxels = two 8-voxel groups = 256 B further
                    const uint32_t off = (s * K::STAGE + kk * 256) >> 4;
                    umma_f16(tmem_base, ((uint64_t)HI << 32) | (a_lo0 + off), ((uint64_t)HI << 32) | (b_lo0 + off), idesc,
                             (i | kk) ? 1u : 0u);
                }
                umma_commit(EMPTY(s));
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(DONE);
        __syncwarp();
    }

    // ===================================================================== drain: TMEM lane = input channel k of the tile
    mbar_wait(DONE, 0, ctx, 3);
    tc_fence_after();
    if (warp < 4) {
        const int k = warp * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* out = p.part + (((long long)blockIdx.y * 27 + tap) * p.Np + (long long)nt * BN) * p.Kp + kt * BMK + k;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(trow + c0, v);                           // .sync.aligned: every lane takes part
            tmem_wait_ld();
            if (k < kch) {
#pragma unroll
                for (int q = 0; q < 16; ++q) out[(long long)(c0 + q) * p.Kp] = __uint_as_float(v[q]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPRODW) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
}

template <int BN>
int launch(const WgemmParams& p, cudaStream_t stream) {
    using K = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(wgrad_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
        configured = true;
    }
    dim3 grid(27 * p.ktiles * p.ntiles, p.msplit);
    wgrad_gemm_tc_kernel<BN><<<grid, NTHREADS, K::SMEM, stream>>>(p);
    return (int)cudaGetLastError();
}

int pick_bn(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 0)); }

void plan(long long M, int Kp, int Np, int& msplit, int& mper) {
    const int base = 27 * ((Kp + BMK - 1) / BMK) * (Np / pick_bn(Np));
    long long ms = (2LL * fcd_num_sms() + base - 1) / base;
    const long long maxs = (M + 4 * BKV - 1) / (4 * BKV);       // >= 4 pipeline stages per CTA
    if (ms > maxs) ms = maxs;
    if (ms < 1) ms = 1;
    long long per = (M + ms - 1) / ms;
    per = (per + BKV - 1) / BKV * BKV;
    ms = (M + per - 1) / per;
    msplit = (int)ms; mper = (int)per;
}

}  // namespace

// Number of partial dW buffers fcd_wgrad_gemm_tc writes (0: shape not taken -- Kp, Np must be multiples of 64).
FCD_API int fcd_wgrad_gemm_tc_nsplit(long long M, int Kp, int Np) {
    if (M < 64 || M > 0x7fffffffLL || Kp % 64 || pick_bn(Np) == 0) return 0;
    // measured (tools/time_deep_wgrad.py, batch 2): 2-4.6x faster than the previous paths on every deep shape except
    // 64 -> 64 channels, where half of the 128-row M tile is padding and the kd-folded fcd_wgrad3_tc slices win
    if (Kp == 64 && Np == 64) return 0;
    int ms, per;
    plan(M, Kp, Np, ms, per);
    return ms;
}

// X: conv input rows (pitch ldx >= Kp), dY: output-gradient rows (pitch ldy >= Np), both NDHWC bf16 of the same volume;
// part: fcd_wgrad_gemm_tc_nsplit() x [27][Np][Kp] fp32, finished by fcd_wgrad_reduce.
FCD_API int fcd_wgrad_gemm_tc(const void* X, long long ldx, const void* dY, long long ldy, float* part, int Bn, int D,
                              int H, int W, int Kp, int Np, cudaStream_t stream) {
    const int bn = pick_bn(Np);
    const long long M = (long long)Bn * D * H * W;
    if (bn == 0 || Kp % 64 || ldx % 8 || ldy % 8 || M < 64 || M > 0x7fffffffLL) return -1;
    if (((uintptr_t)X & 15) || ((uintptr_t)dY & 15)) return -1;
    WgemmParams p;
    p.X = (const bf16*)X; p.ldx = ldx; p.dY = (const bf16*)dY; p.ldy = ldy; p.part = part;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W; p.Kp = Kp; p.Np = Np; p.M = (int)M;
    p.ktiles = (Kp + BMK - 1) / BMK; p.ntiles = Np / bn; p.status = fcd_status_dev();
    plan(M, Kp, Np, p.msplit, p.mper);
    if (bn == 256) return launch<256>(p, stream);
    if (bn == 128) return launch<128>(p, stream);
    return launch<64>(p, stream);
}
