// Weight gradient of the 3x3x3 stride-1 pad-1 convs on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators):
// autograd of the nn.Conv3d modules built by get_conv_layer (reference conv_blocks.py:393-416) and MONAI ResBlock
// (segresnet_dsa.py:102).
//
//   dW[tap][n][k] = sum_{b,z,y,x} U[b,z,y,x,n] * S[b, z+kd-1, y+kh-1, x+kw-1, k]       tap = (kd*3+kh)*3+kw
//
// S is the operand that is read shifted (the conv input X, CS channels), U the unshifted one (the output gradient dY,
// CU channels); CS, CU in {16, 32} -- wider layers are cut into 32-channel slices by the caller (row pitch and
// channel offset are free parameters).
//
// GEMM view: D[(kd,k)][n] += A[(kd,k)][voxel] * B[voxel][n], K = 16 voxels per tcgen05.mma.  Both operands are read
// MN-major straight from the staging layout conv_tc.cu uses, [ch/8][voxel][8 ch] (8 channels = 16 contiguous bytes, the
// next voxel 16 B further, the next channel octet one plane further) -- exactly the UMMA no-swizzle MN-major canonical
// layout.  A tap shift (kh,kw) is again just a start-address offset inside the 18x10 halo plane.  The three kd taps are
// FOLDED INTO M: the halo planes z-1, z, z+1 sit in three consecutive ring slots, so "next M-group" (SBO) walks from
// one plane's channel octets into the next plane's -- one M=64 (CS=16) or M=128 (CS=32) instruction covers three taps.
// The ring has exactly DL+2 slots = one work item (a 16x8 tile over DL planes plus its two halo planes), so the
// three-slot window never wraps.  Planes outside the volume are zero-filled by the producers.
//
// The 9 (kh,kw) accumulators, [(kd,k)][n] fp32 each, stay in TMEM for the CTA's whole life (persistent CTA, all its
// work items); at the end each CTA writes ONE partial dW, and fcd_wgrad_reduce sums the <= 148 partials in a fixed
// order (deterministic, no atomics).
//
// The three kw taps are FOLDED INTO N (round 2): the U tile is staged three times, copy kw shifted by kw-1 voxels along W
// (B_kw[v] = U[v - (kw-1)], zeros outside the volume), the copies' channel octets consecutive at the N-group stride -- so
// one instruction with N = 3 CU and A = the centre column S[v + (kd-1, kh-1, 0)] accumulates sum_v U[v-(kw-1)] S[v+..] =
// sum_v' U[v'] S[v' + (kd-1, kh-1, kw-1)] for kw = 0..2 at once (every (v', tap) pair is still counted exactly once over
// the volume: by the tile that holds v = v' + kw - 1).  24 instead of 72 tcgen05.mma per plane; an M64 x N48 instruction
// costs 28 cycles against 3 x 23 for three N16 ones (section 3.1's instruction table).  The TMEM columns of warp kh are
// [kw][n] as before, so the dump is unchanged.
//
// Warps: 0-1 S producers (cp.async 16 B pieces, as conv_tc.cu), 2-4 U producers (one per kw copy), 5-7 MMA issuers (warp kh
// owns taps (kh, 0..2) and their accumulators), then warps 0-2 dump TMEM.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TH = 16, TW = 8, HH = TH + 2, HW = TW + 2, HV = HH * HW;
constexpr int NPS = 2;                    // S producer warps (4 measured no faster)
constexpr int NMMA = 3;
constexpr int NPU = 3;                    // U producer warps: one per kw copy
constexpr int NTHREADS = 32 * (NPS + NPU + NMMA);
constexpr int DEPTH = 3;                  // cp.async groups in flight per producer lane

struct WgradTcParams {
    const bf16* S; long long lds;
    const bf16* U; long long ldu;
    float* part;                          // [grid][27][ldn][ldk] fp32
    int ldn, ldk, n_off, k_off;
    int nks;                              // blockIdx.y = slice: k slice (y % nks) of CS channels, n slice (y / nks) of CU
    int Bn, D, H, W, nht, nwt, nseg, nitems;
    int* status;
};

template <int CS, int CU, int DL>
struct Cfg {
    static constexpr int NS = DL + 2;                       // S ring slots = planes of one item
    static constexpr int PS_BYTES = HV * CS * 2;            // [CS/8][180][8]
    static constexpr int PU_COPY = TH * TW * CU * 2;        // one kw copy: [CU/8][128][8]
    static constexpr int PU_BYTES = 3 * PU_COPY;            // [kw][CU/8][128][8]
    static constexpr int M = CS == 16 ? 64 : 128;           // 3 kd taps x CS rows (+ one ignored plane's worth)
    static constexpr int SBO_A = HV * 16, LBO_A = HW * 16;  // M-group (channel octet / next plane), K-group (next h row)
    static constexpr int SBO_B = TH * TW * 16, LBO_B = TW * 16;
    static constexpr int TCOLS = 9 * CU;
    static constexpr int TMEM_COLS = TCOLS <= 256 ? 256 : 512;
    // the ignored 4th plane of the last window reads past the S ring: the U ring sits right behind it
    // U ring: with kw folded into N a plane is 24 instructions (~0.35 us): the ring and the producer's cp.async depth must
    // cover ~2 us of load latency.  As many slots as fit beside the S ring in one CTA's shared memory (<= 12), DU planes
    // of cp.async in flight.
    // (measured, 16->16 @128^3: 12 slots / 6 planes in flight with ONE CTA per SM 0.367 ms, 4 slots with two CTAs 0.260 ms)
    static constexpr int NU = 4;
    static constexpr int DU = 2;     // (3: 0.187 -> 0.173 ms stand-alone, nothing in the step)
    static constexpr int SMEM = NS * PS_BYTES + NU * PU_BYTES + 1024;
    static constexpr int CTAS_PER_SM = (2 * SMEM <= 226 * 1024 && 2 * TMEM_COLS <= 512) ? 2 : 1;   // issue-bound: interleave two CTAs
    static_assert(PS_BYTES % 128 == 0 && PU_BYTES % 128 == 0, "alignment");
};

struct Item { int n, h0, w0, d0; };
__device__ __forceinline__ Item decode(const WgradTcParams& p, int item, int DL) {
    Item it;
    int wt = item % p.nwt; item /= p.nwt;
    int ht = item % p.nht; item /= p.nht;
    int seg = item % p.nseg; item /= p.nseg;
    it.n = item; it.h0 = ht * TH; it.w0 = wt * TW; it.d0 = seg * DL;
    return it;
}

template <int CS, int CU, int DL>
__global__ void __launch_bounds__(NTHREADS, Cfg<CS, CU, DL>::CTAS_PER_SM) wgrad3_tc_kernel(const WgradTcParams p) {
    using K = Cfg<CS, CU, DL>;
    constexpr int NS = K::NS, NU = K::NU;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sring = smem;
    unsigned char* uring = smem + NS * K::PS_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * K::PS_BYTES + NU * K::PU_BYTES);
    // bars: [0,NS) FULL_S | [NS,2NS) EMPTY_S | [2NS,2NS+NU) FULL_U | [2NS+NU,2NS+2NU) EMPTY_U | [2NS+2NU] DONE
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2 * NU + 1);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL_S = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY_S = [&](int s) { return bar0 + 8u * (NS + s); };
    auto FULL_U = [&](int s) { return bar0 + 8u * (2 * NS + s); };
    auto EMPTY_U = [&](int s) { return bar0 + 8u * (2 * NS + NU + s); };
    const uint32_t DONE = bar0 + 8u * (2 * NS + 2 * NU);

    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 4);
        for (int s = 0; s < NS; ++s) { mbar_init(FULL_S(s), 32 * NPS); mbar_init(EMPTY_S(s), NMMA); }
        for (int s = 0; s < NU; ++s) { mbar_init(FULL_U(s), 32 * NPU); mbar_init(EMPTY_U(s), NMMA); }
        mbar_init(DONE, NMMA);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long plane_s = (long long)p.H * p.W * p.lds, plane_u = (long long)p.H * p.W * p.ldu;
    // all channel slices of one conv in ONE launch (they were one launch each: 16-CTA grids, serial on the side stream)
    const int ksl = blockIdx.y % p.nks, nsl = blockIdx.y / p.nks;
    const bf16* const Sp = p.S + ksl * CS;
    const bf16* const Up = p.U + nsl * CU;
    const int n_off = p.n_off + nsl * CU, k_off = p.k_off + ksl * CS;

    if (warp < NPS) {
        // ===================================================================== S producers: halo planes d0-1 .. d0+DL
        constexpr int C8 = CS / 8, VS = 32 * NPS / C8, NJ = (HV + VS - 1) / VS;
        const int pt = warp * 32 + lane;
        const int c8 = pt % C8, v0 = pt / C8;
        const uint32_t ring_u = smem_u32(sring);
        uint32_t seq = 0, signaled = 0;
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL_S(signaled % NS)); ++signaled; }
        };
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item, DL);
            int off[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int v = v0 + j * VS;
                const int hh = v / HW, ww = v - hh * HW;
                const int h = it.h0 - 1 + hh, w = it.w0 - 1 + ww;
                const bool ok = v < HV && h >= 0 && h < p.H && w >= 0 && w < p.W;
                off[j] = ok ? (int)(((long long)h * p.W + w) * p.lds) + c8 * 8 : -1;
            }
            for (int i = 0; i < NS; ++i, ++seq) {
                const int s = seq % NS;                     // == i
                const uint32_t ph = (seq / NS) & 1u;
                mbar_wait(EMPTY_S(s), ph ^ 1u, ctx, 1);
                const int pl = it.d0 - 1 + i;
                const bool inside = pl >= 0 && pl < p.D;    // planes outside the volume are zero-filled
                const bf16* plane = Sp + ((long long)it.n * p.D + (inside ? pl : 0)) * plane_s;
                const uint32_t dst0 = ring_u + s * K::PS_BYTES + c8 * K::SBO_A + v0 * 16;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (j < NJ - 1 || v0 + j * VS < HV) {
                        const bool ok = inside && off[j] >= 0;
                        cp_async16(dst0 + j * VS * 16, ok ? plane + off[j] : Sp, ok);
                    }
                }
                cp_async_commit();
                if (seq + 1 >= DEPTH) {
                    cp_async_wait<DEPTH - 1>();
                    fence_proxy_async();
                    flush_to(seq + 2 - DEPTH);
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        flush_to(seq);
    } else if (warp < NPS + NPU) {
        // ===================================================================== U producers: planes d0 .. d0+DL-1; warp kw
        // stages copy kw of every plane (the U tile shifted by kw - 1 voxels along W)
        const int kw = warp - NPS;
        constexpr int C8 = CU / 8, VS = 32 / C8, NJ = (TH * TW) / VS;
        const int c8 = lane % C8, v0 = lane / C8;
        const uint32_t ring_u = smem_u32(uring);
        uint32_t seq = 0, signaled = 0;
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL_U(signaled % NU)); ++signaled; }
        };
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item, DL);
            const bf16* col = Up + (((long long)it.n * p.D + it.d0) * p.H + it.h0) * p.W * p.ldu + (long long)it.w0 * p.ldu + c8 * 8;
            for (int j = 0; j < DL; ++j, ++seq) {
                const int s = seq % NU;
                const uint32_t ph = (seq / NU) & 1u;
                mbar_wait(EMPTY_U(s), ph ^ 1u, ctx, 5);
                const bf16* plane = col + (long long)j * plane_u;
                const uint32_t dst0 = ring_u + s * K::PU_BYTES + c8 * K::SBO_B + v0 * 16;
#pragma unroll
                for (int q = 0; q < NJ; ++q) {
                    const int v = v0 + q * VS;               // tile voxel: hh = v / 8, ww = v % 8
                    const int ws = (v & 7) - (kw - 1);       // copy kw holds U shifted by kw - 1 voxels along W
                    const bool ok = it.h0 + (v >> 3) < p.H && it.w0 + ws >= 0 && it.w0 + ws < p.W;       // zero outside
                    cp_async16(dst0 + kw * K::PU_COPY + q * VS * 16,
                               ok ? plane + ((long long)(v >> 3) * p.W + ws) * p.ldu : Up, ok);
                }
                cp_async_commit();
                if (seq + 1 >= K::DU) {                     // DU planes of copies in flight
                    cp_async_wait<K::DU - 1>();
                    fence_proxy_async();
                    flush_to(seq + 2 - K::DU);
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        flush_to(seq);
    } else {
        // ===================================================================== MMA issuers: warp kh owns taps (kh, kw=0..2)
        const int kh = warp - NPS - NPU;
        constexpr uint32_t idesc = umma_idesc(K::M, 3 * CU, 1, 1);       // both operands MN-major; N = (kw, n)
        constexpr uint32_t A_HI = ((K::SBO_A >> 4) & 0x3fffu) | (1u << 14);
        constexpr uint32_t B_HI = ((K::SBO_B >> 4) & 0x3fffu) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(sring) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_A >> 4) << 16);
        const uint32_t b_lo0 = ((smem_u32(uring) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_B >> 4) << 16);
        uint32_t useq = 0, nitem = 0, first = 0;            // first == 0: the accumulators are still uninitialised
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++nitem) {
            const uint32_t sph = nitem & 1u;                 // every S slot is used exactly once per item
            int waited = 0;
            for (int j = 0; j < DL; ++j, ++useq) {
                while (waited < j + 3) { mbar_wait(FULL_S(waited), sph, ctx, 2); ++waited; }
                const int us = useq % NU;
                mbar_wait(FULL_U(us), (useq / NU) & 1u, ctx, 6);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_pl = a_lo0 + j * (K::PS_BYTES >> 4) + ((kh * HW * 16) >> 4);
                    const uint32_t b_pl = b_lo0 + us * (K::PU_BYTES >> 4);
                    const uint32_t d_tmem = tmem_base + kh * 3 * CU;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {         // 16 voxels = tile rows 2ks, 2ks+1; centre column of the halo
                        const uint32_t a_lo = a_pl + (((2 * ks * HW + 1) * 16) >> 4);
                        const uint32_t b_lo = b_pl + ((2 * ks * TW * 16) >> 4);
                        umma_f16(d_tmem, ((uint64_t)A_HI << 32) | a_lo, ((uint64_t)B_HI << 32) | b_lo, idesc,
                                 ks ? 1u : first);
                    }
                    umma_commit(EMPTY_S(j));                 // window slides: plane j is done (for this warp)
                    if (j == DL - 1) { umma_commit(EMPTY_S(DL)); umma_commit(EMPTY_S(DL + 1)); }
                    umma_commit(EMPTY_U(us));
                }
                __syncwarp();
                first = 1;
            }
        }
        if (lane == 0) umma_commit(DONE);
        __syncwarp();
    }

    // ===================================================================== dump: one partial dW per CTA
    mbar_wait(DONE, 0, ctx, 7);
    tc_fence_after();
    __syncthreads();
    if (warp < 3) {
        const int kd = warp;                                  // TMEM quadrant = kd tap, lane = input channel k
        float* out = p.part + (long long)blockIdx.x * 27 * p.ldn * p.ldk;
#pragma unroll 1
        for (int khw = 0; khw < 9; ++khw) {
            const int t = kd * 9 + khw;
            uint32_t v[CU];
#pragma unroll
            for (int c0 = 0; c0 < CU; c0 += 16)              // .sync.aligned: all 32 lanes take part
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + khw * CU + c0, v + c0);
            tmem_wait_ld();
            if (lane < CS) {
#pragma unroll
                for (int n = 0; n < CU; ++n)
                    out[((long long)t * p.ldn + n_off + n) * p.ldk + k_off + lane] = __uint_as_float(v[n]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<K::TMEM_COLS>(tmem_base);
}

template <int CS, int CU, int DL>
int launch(const WgradTcParams& p, int grid, int nslices, cudaStream_t stream) {
    using K = Cfg<CS, CU, DL>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(wgrad3_tc_kernel<CS, CU, DL>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
        configured = true;
    }
    wgrad3_tc_kernel<CS, CU, DL><<<dim3(grid, nslices), NTHREADS, K::SMEM, stream>>>(p);
    return (int)cudaGetLastError();
}

int pick_dl(int Bn, int D, int H, int W) {
    if (H < 1 || W < 1) return 0;
    const long long cols = (long long)Bn * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
    if (D % 8 == 0 && cols * (D / 8) >= 2 * fcd_num_sms()) return 8;
    if (D % 4 == 0) return 4;
    return 0;
}

}  // namespace

// Number of partial dW buffers (= CTAs) fcd_wgrad3_tc writes for this volume; 0: shape not supported.
FCD_API int fcd_wgrad3_tc_nsplit(int Bn, int D, int H, int W) {
    const int dl = pick_dl(Bn, D, H, W);
    if (dl == 0) return 0;
    const long long items = (long long)Bn * ((H + TH - 1) / TH) * ((W + TW - 1) / TW) * (D / dl);
    return (int)(items < 2LL * fcd_num_sms() ? items : 2LL * fcd_num_sms());   // two CTAs per SM when they fit
}

// part[nsplit][27][ldn][ldk] fp32 (fcd_wgrad_reduce layout): this call fills the [n_off, n_off+nns*CU) x
// [k_off, k_off+nks*CS) block of every tap in every partial, one CTA row (blockIdx.y) per CU x CS slice.
// S: shifted operand rows (pitch lds, already offset to channel k_off), U: unshifted operand rows (pitch ldu, offset to
// channel n_off).  CS, CU in {16, 32}.
FCD_API int fcd_wgrad3_tc(const void* S, long long lds, const void* U, long long ldu, float* part, int ldn, int ldk,
                          int n_off, int k_off, int nns, int nks, int Bn, int D, int H, int W, int CS, int CU,
                          cudaStream_t stream) {
    const int dl = pick_dl(Bn, D, H, W);
    if (dl == 0 || lds % 8 || ldu % 8 || ((uintptr_t)S & 15) || ((uintptr_t)U & 15)) return -1;
    if (!(CS == 16 || CS == 32) || !(CU == 16 || CU == 32)) return -1;
    if (nns < 1 || nks < 1 || (long long)nns * nks > 65535) return -1;
    if (n_off + nns * CU > ldn || k_off + nks * CS > ldk) return -1;
    WgradTcParams p;
    p.S = (const bf16*)S; p.lds = lds; p.U = (const bf16*)U; p.ldu = ldu; p.part = part;
    p.ldn = ldn; p.ldk = ldk; p.n_off = n_off; p.k_off = k_off; p.nks = nks;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W; p.nht = (H + TH - 1) / TH; p.nwt = (W + TW - 1) / TW; p.nseg = D / dl;
    p.nitems = Bn * p.nht * p.nwt * p.nseg; p.status = fcd_status_dev();
    const int grid = p.nitems < 2 * fcd_num_sms() ? p.nitems : 2 * fcd_num_sms();   // == fcd_wgrad3_tc_nsplit
#define FCD_WG_CASE(A, B, L) if (CS == A && CU == B && dl == L) return launch<A, B, L>(p, grid, nns * nks, stream)
    FCD_WG_CASE(16, 16, 8); FCD_WG_CASE(16, 32, 8); FCD_WG_CASE(32, 16, 8); FCD_WG_CASE(32, 32, 8);
    FCD_WG_CASE(16, 16, 4); FCD_WG_CASE(16, 32, 4); FCD_WG_CASE(32, 16, 4); FCD_WG_CASE(32, 32, 4);
#undef FCD_WG_CASE
    return -1;
}
