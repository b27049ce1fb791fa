// 3x3x3 stride-1 pad-1 Conv3d as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM), fed by halo planes staged once in shared memory -- the FLOP-heavy layers of the hot path (reference
// conv_blocks.py:393-416 via get_conv_layer; MONAI ResBlock convs segresnet_dsa.py:102; their data gradients).
//
//   out[b, z, y, x, n] = sum_{kd,kh,kw,k} in[b, z+kd-1, y+kh-1, x+kw-1, k] * Wp[tap][n][k]        (tap = (kd*3+kh)*3+kw)
//
// Decomposition.  A work item is a column of output planes: batch b, a 16(h) x 8(w) tile, planes [d0, d0+DL).
// A persistent CTA (one per SM, 9 warps) marches down the column:
//   warps 0-1 producers: one halo plane (18 x 10 voxels x CIN) per step into a ring of NST stages, stored as
//           [CIN/8][18*10 voxels][8 ch] -- the UMMA no-swizzle K-major canonical layout (core matrix = 8 consecutive
//           voxels x 16 B).  The channel de-interleave makes every element a 16 B piece, so the feed is 16 B
//           cp.async (zero-fill for the out-of-range h/w halo = the conv padding; out-of-range planes are simply not
//           multiplied), completion handed to the MMA warp through an mbarrier after a generic->async proxy fence.
//           (Measured on B200: the same box through a 5-D TMA tensor map is request-rate bound -- one 16 B inner row
//           per ~6.5 cycles, 3300 cycles per Cin=16 plane against an 860-cycle MMA budget -- see DESIGN.md.)
//   warps 2-4 MMA issuers (warp kd owns input plane z+kd-1): per output plane 27 * CIN/16 tcgen05.mma of M=128 voxels
//           x N=COUT x K=16 into three TMEM accumulators.  A tap
//           shift is nothing but a different start address of the A descriptor inside the halo plane
//           ((kh*10 + kw) * 16 B), so each halo plane is read from HBM/L2 once and reused by all 27 taps (9 per
//           plane x 3 output planes).  All 27 weight taps stay resident in shared memory.  Two TMEM accumulators
//           ping-pong so the epilogue of plane z overlaps the MMAs of plane z+1.
//   warps 5-8 epilogue: tcgen05.ld (one voxel row of COUT fp32 per thread) -> bf16 -> 16 B stores into the NDHWC
//           output (any row pitch: writes straight into concat buffers), plus the per-(b, channel) sum / sum of
//           squares of the ROUNDED outputs for the InstanceNorm that follows every conv (conv_blocks.py:439-452),
//           so the separate statistics pass over the tensor disappears.
#include <cstdlib>

#include "last_block.cuh"
#include "norm_fin.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TH = 16, TW = 8;            // output tile per plane (UMMA M = 128 voxels)
constexpr int HH = TH + 2, HW = TW + 2;   // halo plane
constexpr int HV = HH * HW;               // 180 voxels

constexpr int NPROD = 2;                  // producer warps
constexpr int NMMA = 3;                   // MMA-issuing warps: one per kd tap plane, each with its own accumulator
constexpr int NTHREADS = 32 * (NPROD + NMMA + 4);

struct ConvTcParams {
    const bf16* A;
    long long lda;
    const float* Wf;    // fp32 parameter: element (tap t, out n, in k) at Wf[n*sn + k*sk + t*st]
    long long sn, sk, st;
    int Nr, Kr;         // real (unpadded) channel counts
    int kseg, ksegpad;  // input channels arrive in concat segments of kseg real channels padded to ksegpad
    int nsg, nsgpad;    // same for the output channels (data gradient of a conv that read a concat buffer)
    bf16* C;
    long long ldc;
    float* part;        // nullptr or [B][nchunk][2][COUT] fp32, nchunk = nht*nwt*nseg
    const float* bias;  // nullptr or [COUT] fp32, added before the bf16 rounding
    int Bn, D, H, W;
    int nht, nwt, nseg, DL, nitems;
    int flip;           // 1: use tap 26-t (data gradient of a stride-1 conv = correlation with the mirrored kernel)
    int* status;
    NormFin fin;        // fin.mean != nullptr: the last CTA turns the fused partials into mean / rstd
    unsigned* ticket;
};

template <int CIN, int COUT>
struct Cfg {
    static constexpr int W_BYTES = 27 * CIN * COUT * 2;
    static constexpr int TAP_BYTES = CIN * COUT * 2;
    static constexpr int PLANE_BYTES = HV * CIN * 2;
    static constexpr int LBO_A = HV * 16, SBO_A = HW * 16;     // K-chunk stride, 8-voxel-group (= next h row) stride
    static constexpr int LBO_B = COUT * 16, SBO_B = 128;
    static constexpr int BUDGET = 220 * 1024 - W_BYTES;
    static constexpr int NST = BUDGET / PLANE_BYTES >= 6 ? 6 : BUDGET / PLANE_BYTES;
    // One issuing warp was the limiter (ptxas wraps every UTCHMMA in an ELECT loop; ncu showed the MMA warp busy issuing,
    // never waiting), so three warps issue -- warp kd owns the 9 taps of input plane z+kd-1 and its own TMEM
    // accumulator; the epilogue adds the (valid) three: 0.276 -> 0.223 ms for 16->16 @128^3.  Measured: splitting the
    // taps of ONE issuing warp over 4 or 8 accumulators changes nothing (the accumulate chain is not the limiter).
    static constexpr int TCOLS = 2 * NMMA * COUT;
    static constexpr int TMEM_COLS = TCOLS <= 32 ? 32 : (TCOLS <= 64 ? 64 : (TCOLS <= 128 ? 128 : (TCOLS <= 256 ? 256 : 512)));
    static_assert(TCOLS <= 512, "TMEM columns");
    static constexpr int SMEM = W_BYTES + NST * PLANE_BYTES + 2560;   // + barriers, tmem slot, stats scratch
    // Small-Cout layers are limited by the latency-bound single-warp roles (one epilogue warp per SMSP), so two CTAs
    // share an SM when shared memory and registers allow: their roles interleave (0.191 -> 0.166 ms for 16->16).
    static constexpr int CTAS_PER_SM = (COUT == 16 && 2 * SMEM <= 226 * 1024) ? 2 : 1;
    static_assert(NST >= 4, "need >= 4 halo-plane stages");
    static_assert(PLANE_BYTES % 128 == 0 && W_BYTES % 128 == 0, "alignment");
};

struct Item {
    int n, h0, w0, d0, d1, p_lo, p_hi, chunk;
};
__device__ __forceinline__ Item decode(const ConvTcParams& p, int item) {
    Item it;
    int wt = item % p.nwt; item /= p.nwt;
    int ht = item % p.nht; item /= p.nht;
    int seg = item % p.nseg; item /= p.nseg;
    it.n = item;
    it.h0 = ht * TH; it.w0 = wt * TW;
    it.d0 = seg * p.DL; it.d1 = min(it.d0 + p.DL, p.D);
    it.p_lo = max(it.d0 - 1, 0); it.p_hi = min(it.d1, p.D - 1);
    it.chunk = (seg * p.nht + ht) * p.nwt + wt;
    return it;
}

template <int CIN, int COUT, bool STATS, bool FLIP>
__global__ void __launch_bounds__(NTHREADS, Cfg<CIN, COUT>::CTAS_PER_SM) conv3_tc_kernel(const ConvTcParams p) {
    using K = Cfg<CIN, COUT>;
    constexpr int NST = K::NST;
    constexpr int DEPTH = NST >= 6 ? 3 : NST - 3;   // cp.async groups in flight per producer lane (>= 1)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* wsm = smem;
    unsigned char* ring = smem + K::W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::W_BYTES + NST * K::PLANE_BYTES);
    // bars[0..NST) full, [NST..2NST) empty, [2NST..2NST+2) tmem_full, [2NST+2..2NST+4) tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    float* red = reinterpret_cast<float*>(ctx + 1);              // [4 warps][2][COUT] (STATS)  <= 2 KB for COUT 64
    float* sbias = red + 8 * (COUT <= 32 ? COUT : 0);            // (no statistics for COUT 64: red is unused there)

    const int tid = threadIdx.x, lane = tid & 31;
    // shfl from a fixed lane: provably warp-uniform, so the role branches and everything inside the MMA role stay on
    // the uniform datapath (UTCHMMA takes its descriptors from uniform registers)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    auto TFULL = [&](int b) { return bar0 + 8u * (2 * NST + b); };
    auto TEMPTY = [&](int b) { return bar0 + 8u * (2 * NST + 2 + b); };

    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 32 * NPROD); mbar_init(EMPTY(s), NMMA); }
        for (int b = 0; b < 2; ++b) { mbar_init(TFULL(b), NMMA); mbar_init(TEMPTY(b), 4); }
        fence_barrier_init();
    }
    if (warp == NPROD) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));   // (warp NPROD also deallocates)
    // weights: fp32 parameter (any strides) -> bf16 smem [tap][k/8][n][8]  (K-major no-swizzle core matrices:
    // 8 couts x 16 B); padded / out-of-segment channels are zero.  No separate pack kernel, no packed copy in HBM.
    {
        constexpr int CHUNKS = 27 * COUT * (CIN / 8);
        for (int i = tid; i < CHUNKS; i += NTHREADS) {
            const int t = i % 27;                               // tap fastest: the parameter is tap-contiguous
            const int np_ = (i / 27) % COUT;
            const int c8 = i / (27 * COUT);
            const int ns = np_ / p.nsgpad, nw = np_ % p.nsgpad;
            const int n = ns * p.nsg + nw;
            const bool nok = nw < p.nsg && n < p.Nr;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kp = c8 * 8 + j;
                const int ks = kp / p.ksegpad, kwi = kp % p.ksegpad;
                const int k = ks * p.kseg + kwi;
                f[j] = (nok && kwi < p.kseg && k < p.Kr) ? __ldg(p.Wf + n * p.sn + k * p.sk + t * p.st) : 0.f;
            }
            *reinterpret_cast<bf16x8*>(wsm + t * K::TAP_BYTES + c8 * K::LBO_B + np_ * 16) = pack8(f);
        }
    }
    if (tid < COUT) sbias[tid] = p.bias != nullptr ? p.bias[tid] : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPROD) {
        // ===================================================================== producers (cp.async, 16 B pieces)
        constexpr int C8 = CIN / 8;               // 16 B pieces per voxel
        constexpr int VS = 32 * NPROD / C8;       // voxels covered per pass of all producer lanes
        constexpr int NJ = (HV + VS - 1) / VS;
        const int pt = warp * 32 + lane;
        const int c8 = pt % C8, v0 = pt / C8;
        const uint32_t ring_u = smem_u32(ring);
        uint32_t seq = 0, signaled = 0;           // planes issued / planes handed to the MMA warp (per lane)
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL(signaled % NST)); ++signaled; }
        };
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item);
            // per-lane source offsets of my NJ pieces inside a plane (the same for every plane of the column):
            // element offset from the plane base, or -1 for the zero-filled halo outside the volume
            int off[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int v = v0 + j * VS;
                const int hh = v / HW, ww = v - hh * HW;
                const int h = it.h0 - 1 + hh, w = it.w0 - 1 + ww;
                const bool ok = v < HV && h >= 0 && h < p.H && w >= 0 && w < p.W;
                off[j] = ok ? (int)(((long long)h * p.W + w) * p.lda) + c8 * 8 : -1;
            }
            const long long plane_elems = (long long)p.H * p.W * p.lda;
            for (int pl = it.p_lo; pl <= it.p_hi; ++pl, ++seq) {
                const int s = seq % NST;
                const uint32_t ph = (seq / NST) & 1u;
                // Slot s frees when the MMAs that read plane seq-NST are done; those need planes <= seq-NST+2, all
                // handed over already because hand-over lags issue by DEPTH-1 <= NST-3 planes.  So a plain wait cannot
                // deadlock, and (unlike draining first) it keeps DEPTH planes in flight in steady state.
                mbar_wait(EMPTY(s), ph ^ 1u, ctx, 1);
                const bf16* plane = p.A + ((long long)it.n * p.D + pl) * plane_elems;
                const uint32_t dst0 = ring_u + s * K::PLANE_BYTES + c8 * K::LBO_A + v0 * 16;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if ((NJ - 1) * VS + 32 * NPROD / C8 <= HV || j < NJ - 1 || v0 + j * VS < HV) {
                        const bool ok = off[j] >= 0;
                        cp_async16(dst0 + j * VS * 16, ok ? plane + off[j] : p.A, ok);
                    }
                }
                cp_async_commit();
                if (seq + 1 >= DEPTH) {            // planes up to seq-(DEPTH-1) have landed for this lane
                    cp_async_wait<DEPTH - 1>();
                    fence_proxy_async();
                    flush_to(seq + 2 - DEPTH);
                }
                static_assert(NST >= DEPTH + 1 + 2, "ring too shallow for the hand-over lag");
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        flush_to(seq);
    } else if (warp < NPROD + NMMA) {
        // ===================================================================== MMA issuers: warp kd multiplies input
        // plane z+kd-1 (9 taps x CIN/16 k-steps) into accumulator kd of the output plane's TMEM buffer.  Every
        // warp walks ALL loaded planes of an item in order (waiting FULL, arriving EMPTY) so the barrier phases stay
        // in step even for the planes it never reads, and arrives on TFULL once per output plane.
        {
            const int kd = warp - NPROD;
            constexpr uint32_t idesc = umma_idesc(128, COUT, 0, 0);
            constexpr uint32_t A_HI = ((K::SBO_A >> 4) & 0x3fffu) | (1u << 14);
            constexpr uint32_t B_HI = ((K::SBO_B >> 4) & 0x3fffu) | (1u << 14);
            constexpr int TAP16 = K::TAP_BYTES >> 4;
            const uint32_t a_lo0 = ((smem_u32(ring) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_A >> 4) << 16);
            const uint32_t b_lo0 = ((smem_u32(wsm) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_B >> 4) << 16);
            const uint32_t b_kd = b_lo0 + (FLIP ? 26 - kd * 9 : kd * 9) * TAP16;
            uint32_t seq_base = 0, odc = 0;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                const Item it = decode(p, item);
                const int nload = it.p_hi - it.p_lo + 1;
                int done = 0;                                  // planes of this item waited for + released by this warp
                auto pass_planes = [&](int upto) {             // walk (without reading) planes [done, upto)
                    while (done < upto) {
                        const uint32_t sq = seq_base + done;
                        mbar_wait(FULL(sq % NST), (sq / NST) & 1u, ctx, 2);
                        if (lane == 0) mbar_arrive(EMPTY(sq % NST));
                        ++done;
                    }
                };
                for (int od = it.d0; od < it.d1; ++od, ++odc) {
                    const int buf = odc & 1;
                    const uint32_t uph = (odc >> 1) & 1u;
                    const int pl = od + kd - 1;
                    const bool valid = pl >= 0 && pl < p.D;
                    if (valid) {
                        pass_planes(pl - it.p_lo);             // planes before mine that I never read (item start)
                        const uint32_t sq = seq_base + done;   // == plane pl
                        mbar_wait(FULL(sq % NST), (sq / NST) & 1u, ctx, 2);
                        mbar_wait(TEMPTY(buf), uph ^ 1u, ctx, 3);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (buf * NMMA + kd) * COUT;
                        const uint32_t a_pl = a_lo0 + (sq % NST) * (K::PLANE_BYTES >> 4);
                        if (lane == 0) {
#pragma unroll
                            for (int khw = 0; khw < 9; ++khw) {
                                const int kh = khw / 3, kw = khw % 3;
#pragma unroll
                                for (int kc = 0; kc < CIN / 16; ++kc) {
                                    const uint32_t a_lo = a_pl + (((kh * HW + kw) * 16 + kc * 2 * K::LBO_A) >> 4);
                                    const uint32_t b_lo = b_kd + (FLIP ? -khw * TAP16 : khw * TAP16) +
                                                          ((kc * 2 * K::LBO_B) >> 4);
                                    umma_f16(d_tmem, ((uint64_t)A_HI << 32) | a_lo, ((uint64_t)B_HI << 32) | b_lo,
                                             idesc, (khw | kc) ? 1u : 0u);
                                }
                            }
                            umma_commit(EMPTY(sq % NST));      // my reads of plane pl are done when these MMAs are
                            umma_commit(TFULL(buf));
                        }
                        __syncwarp();
                        ++done;
                    } else {
                        // zero-padding plane: nothing to add, but the epilogue still counts NMMA arrivals.  Wait for
                        // the buffer first: an early arrive would be counted into the previous phase of TFULL(buf).
                        mbar_wait(TEMPTY(buf), uph ^ 1u, ctx, 3);
                        if (lane == 0) mbar_arrive(TFULL(buf));
                    }
                }
                pass_planes(nload);                            // planes after my last read (item end)
                seq_base += nload;
            }
        }
    } else {
        // ===================================================================== epilogue (4 warps)
        const int q = warp & 3;                   // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;              // accumulator row = voxel (hh, ww) of the tile
        const int hh = r >> 3, ww = r & 7;
        uint32_t odc = 0;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item);
            float s1[STATS ? COUT : 1], s2[STATS ? COUT : 1];
            if (STATS) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) s1[c] = s2[c] = 0.f;
            }
            for (int od = it.d0; od < it.d1; ++od, ++odc) {
                const int buf = odc & 1;
                const uint32_t uph = (odc >> 1) & 1u;
                mbar_wait(TFULL(buf), uph, ctx, 4);
                tc_fence_after();
                // accumulator kd is valid iff input plane od+kd-1 exists (kd = 1 always)
                float v[COUT];
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (NMMA * COUT);
                {
                    uint32_t t[COUT];
#pragma unroll
                    for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16(trow + COUT + c0, t + c0);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < COUT; ++k) v[k] = __uint_as_float(t[k]);
                    if (od > 0) {
#pragma unroll
                        for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16(trow + c0, t + c0);
                        tmem_wait_ld();
#pragma unroll
                        for (int k = 0; k < COUT; ++k) v[k] += __uint_as_float(t[k]);
                    }
                    if (od < p.D - 1) {
#pragma unroll
                        for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16(trow + 2 * COUT + c0, t + c0);
                        tmem_wait_ld();
#pragma unroll
                        for (int k = 0; k < COUT; ++k) v[k] += __uint_as_float(t[k]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(buf));
                bf16* dst = p.C + ((((long long)it.n * p.D + od) * p.H + it.h0 + hh) * p.W + it.w0 + ww) * p.ldc;
                if (!(it.h0 + hh < p.H && it.w0 + ww < p.W)) continue;     // ragged edge tile: H % 16 or W % 8 != 0
#pragma unroll
                for (int c0 = 0; c0 < COUT; c0 += 8) {
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = v[c0 + k] + sbias[c0 + k];
                    const bf16x8 pk = pack8(f);
                    st8(dst + c0, pk);
                    if (STATS) {
                        float g[8];
                        unpack8(pk, g);
#pragma unroll
                        for (int k = 0; k < 8; ++k) { s1[c0 + k] += g[k]; s2[c0 + k] = fmaf(g[k], g[k], s2[c0 + k]); }
                    }
                }
            }
            if (STATS) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
                    if (lane == 0) { red[(q * 2 + 0) * COUT + c] = a; red[(q * 2 + 1) * COUT + c] = b; }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int e = tid - 32 * (NPROD + NMMA);   // 0..127 over the epilogue threads
                if (e < 2 * COUT) {
                    const float t = red[e] + red[2 * COUT + e] + red[4 * COUT + e] + red[6 * COUT + e];
                    const long long nchunk = (long long)p.nht * p.nwt * p.nseg;
                    p.part[((long long)it.n * nchunk + it.chunk) * 2 * COUT + e] = t;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPROD) tmem_dealloc<K::TMEM_COLS>(tmem_base);
    if (STATS && p.fin.mean != nullptr && lastblk::arrive(p.ticket, gridDim.x))
        norm_finalize_block(p.part, p.fin, reinterpret_cast<double*>(smem));      // all shared memory is free by now
}

template <int CIN, int COUT, bool STATS, bool FLIP>
int launch2(const ConvTcParams& p, cudaStream_t stream) {
    using K = Cfg<CIN, COUT>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv3_tc_kernel<CIN, COUT, STATS, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             K::SMEM);
        configured = true;
    }
    const int grid = min(p.nitems, fcd_num_sms() * K::CTAS_PER_SM);
    conv3_tc_kernel<CIN, COUT, STATS, FLIP><<<grid, NTHREADS, K::SMEM, stream>>>(p);
    return (int)cudaGetLastError();
}

template <int CIN, int COUT>
int launch(const ConvTcParams& p, cudaStream_t stream) {
    if (p.part != nullptr) {
        if constexpr (COUT <= 32) return p.flip ? launch2<CIN, COUT, true, true>(p, stream)
                                                : launch2<CIN, COUT, true, false>(p, stream);
        else return -1;
    }
    return p.flip ? launch2<CIN, COUT, false, true>(p, stream) : launch2<CIN, COUT, false, false>(p, stream);
}

bool shape_ok(int D, int H, int W, int K, int N) {
    if (H < 1 || W < 1 || D < 1) return false;   // H, W need not be multiples of the 16 x 8 tile: edge tiles are masked
    if (!(K == 16 || K == 32 || K == 64) || !(N == 16 || N == 32 || N == 64)) return false;
    if (K == 64 && N == 64) return false;       // 27 taps of 64x64 weights do not fit next to the halo ring
    return true;
}

}  // namespace

// Number of d-segments the conv column of (B, D, H, W) is cut into (0 = shape not supported by the tcgen05 kernel).
// The fused statistics buffer has nchunk = (H/16)*(W/8)*nseg entries per sample.
FCD_API int fcd_conv3_tc_nseg(int Bn, int D, int H, int W, int K, int N) {
    if (!shape_ok(D, H, W, K, N)) return 0;
    // Minimise  rounds x (planes per item + 3)  over power-of-two segment counts with >= 4 planes per segment, where
    // rounds = ceil(items / resident CTAs) and "+ 3" = the two halo planes an item loads beyond its outputs plus its
    // pipeline fill.  Resident CTAs: two per SM for the kd-folded Cout = 16 configurations whose shared memory allows it
    // (conv_tcf.cu Cfg), else one.  (Measured, batch 2 @128^3 16 -> 16: 1 segment 109 us, 4 segments 127 us -- the
    // first model, rounds over the SM count x (planes + 1), preferred 4.)
    // Round 1 additionally forbade several segments together with several items per CTA because that regime hit rare
    // time-outs; the cause (an aliased FULL-barrier parity wait, conv_tcf.cu) is fixed and tests/test_gpu_conv_stress.py
    // covers the regime.  FCD_NSEG_RESTRICTED=1 restores the old rule for A/B runs.
    static const bool restricted = getenv("FCD_NSEG_RESTRICTED") != nullptr;
    const int cols = Bn * ((H + TH - 1) / TH) * ((W + TW - 1) / TW), sms = fcd_num_sms();
    int per_sm = 1;
    if (N == 16 && K <= 64) {
        const int w_bytes = 27 * K * N * 2, plane = 180 * K * 2;
        int nst = (220 * 1024 - w_bytes) / plane;
        if (nst > 6) nst = 6;
        if (2 * (w_bytes + nst * plane + 2560) <= 226 * 1024) per_sm = 2;
    }
    const long long ctas = (long long)sms * per_sm;
    int nseg = 1;
    long long best = -1;
    for (int c = 1; c == 1 || D / c >= 4; c *= 2) {
        const int dl = (D + c - 1) / c;
        if ((D + dl - 1) / dl != c) continue;
        if (restricted && c > 1 && (long long)cols * c > sms) continue;
        const long long cost = (((long long)cols * c + ctas - 1) / ctas) * (dl + 3);
        if (best < 0 || cost < best) { best = cost; nseg = c; }
    }
    return nseg;
}

// Replaces F.conv3d(x, w, padding=1) for 3x3x3 stride-1 convs (conv_blocks.py:393-416 conv1/conv2; dgrad with flip=1).
// A: NDHWC bf16 rows of pitch lda (>= K); C: NDHWC bf16 rows of pitch ldc; K, N: padded channel counts (16/32/64).
// Wf: the fp32 parameter itself, element (tap, n, k) at Wf[n*sn + k*sk + tap*st] for n < Nr, k < Kr (with the
// concat-segment maps of fcd_pack_weight); it is converted to bf16 while being staged into shared memory.
// part: optional [Bn][nchunk][2][N] fp32 partial (sum, sum of squares) of the rounded outputs.
FCD_API int fcd_conv3_tc(const void* A, long long lda, const float* Wf, int Nr, int Kr, long long sn, long long sk,
                         long long st, int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, float* part,
                         const float* bias, int Bn, int D, int H, int W, int K, int N, int flip, int nseg, float* mean,
                         float* rstd,
                          int norm_mode, float eps, float* running_mean, float* running_var, int crun, float momentum,
                          cudaStream_t stream) {
    if (!shape_ok(D, H, W, K, N) || nseg < 1 || lda % 8 || ldc % 8 || lda < K || ldc < N) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)C & 15) || kseg < 1 || ksegpad < 1 || nsg < 1 || nsgpad < 1) return -1;
    ConvTcParams p;
    p.A = (const bf16*)A; p.lda = lda;
    p.Wf = Wf; p.sn = sn; p.sk = sk; p.st = st; p.Nr = Nr; p.Kr = Kr;
    p.kseg = kseg; p.ksegpad = ksegpad; p.nsg = nsg; p.nsgpad = nsgpad;
    p.C = (bf16*)C; p.ldc = ldc; p.part = part; p.bias = bias;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W;
    p.nht = (H + TH - 1) / TH; p.nwt = (W + TW - 1) / TW; p.nseg = nseg; p.DL = (D + nseg - 1) / nseg;
    p.nseg = (D + p.DL - 1) / p.DL;
    if (p.nseg != nseg) return -1;              // caller sizes `part` with nseg: must be exact
    p.nitems = Bn * p.nht * p.nwt * p.nseg; p.status = fcd_status_dev();
    // optional: the last CTA finishes the fused statistics (mean / rstd of the norm that follows the conv); the caller
    // asks fcd_norm_fin_fold first -- partials too large for one CTA are finished by fcd_norm_finalize
    if (mean != nullptr && (part == nullptr || !fin_fold(Bn, p.nht * p.nwt * p.nseg, 2 * N))) return -1;
    p.fin = NormFin{mean, rstd, running_mean, running_var, Bn, N, p.nht * p.nwt * p.nseg,
                    norm_mode, crun, (long long)D * H * W, eps, momentum};
    p.ticket = p.fin.mean != nullptr ? lastblk::next_ticket() : nullptr;
    p.flip = flip;
#define FCD_TC_CASE(CI, CO) if (K == CI && N == CO) return launch<CI, CO>(p, stream)
    FCD_TC_CASE(16, 16); FCD_TC_CASE(16, 32); FCD_TC_CASE(32, 16); FCD_TC_CASE(32, 32);
    FCD_TC_CASE(64, 32); FCD_TC_CASE(32, 64); FCD_TC_CASE(16, 64); FCD_TC_CASE(64, 16);
#undef FCD_TC_CASE
    return -1;
}
