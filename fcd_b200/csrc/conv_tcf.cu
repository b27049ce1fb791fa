// kd-FOLDED variant of the tcgen05 conv3x3x3 (conv_tc.cu) for Cout <= 32 -- the layers that hold most of the FLOPs.
//
// Observation: a halo-plane A tile with a fixed in-plane shift (kh,kw) is needed by THREE output planes -- plane p feeds
// output z = p-1 with tap kd = 2, z = p with kd = 1, z = p+1 with kd = 0.  The time of these small-N kernels follows the
// NUMBER of tcgen05.mma instructions, not their N (an M=128 x N=16 x K=16 instruction carries 8 cycles of math but
// costs several times that in fixed work: A-tile fetch, issue), so the three are issued as ONE
// instruction of N = 3*Cout whose B operand is [W(2,kh,kw) | W(1,kh,kw) | W(0,kh,kw)] and whose D columns are the
// accumulators of three CONSECUTIVE output planes, kept adjacent in a ring of R TMEM slots.  9*Cin/16 instructions per
// plane instead of 27*Cin/16, same A traffic: measured 1.4x (16->16) to 1.6x (32->16) faster.
//
// Consequences: every instruction accumulates (a slot sees its first tap together with older slots' later taps), so the
// epilogue hands a slot back ZEROED (tcgen05.st) after reading it; a slot is complete once the plane two below it has
// been multiplied; slot ranges that wrap around the ring are issued as two instructions.  The three issuing warps take
// turns on planes and own one accumulator set each (the epilogue adds the three sets).
// (Measured: ONE issuing warp with one accumulator set -- lighter epilogue, 3 CTAs per SM -- is 1.5x SLOWER:
// 0.189 vs 0.120 ms for 16->16 @128^3; the single issue stream / accumulate chain becomes the limiter again.)
// Producers, halo-plane ring, weight staging from the fp32 parameter, fused statistics: as in conv_tc.cu.
#include <cstdlib>

#include "last_block.cuh"
#include "norm_fin.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TH = 16, TW = 8, HH = TH + 2, HW = TW + 2, HV = HH * HW;
constexpr int NPROD = 2, NMMA = 3;
constexpr int NTHREADS = 32 * (NPROD + NMMA + 4);
constexpr int R = 5;                      // accumulator ring slots (3 accumulating + slack for the epilogue)
// Plane-done barriers: more than the furthest the MMA warps can run ahead of the epilogue's wait pointer -- every plane
// (loaded or padding) first waits for TEMPTY of the outputs it touches, which keeps the warps within R = 5 outputs,
// i.e. within 5 + 2 planes per item boundary crossed (<= 4 boundaries with 1-plane items) + 2 = 15 planes -- so a
// barrier can never complete two phases before the epilogue has looked at the first.
constexpr int NPB = 18;

struct ConvTcfParams {
    const bf16* A; long long lda;
    const float* Wf; long long sn, sk, st;
    int Nr, Kr, kseg, ksegpad, nsg, nsgpad;
    bf16* C; long long ldc;
    float* part;
    const float* bias;   // nullptr or [COUT] fp32, added before the bf16 rounding (sub-pixel convs, conv_blocks.py:727-735)
    int Bn, D, H, W, nht, nwt, nseg, DL, nitems;
    int* status;
    int accumulate;      // 1: C += result (bf16 read-modify-write): K-sliced data gradients of convs with > 64 output channels
    NormFin fin;        // fin.mean != nullptr: the last CTA turns the fused partials into mean / rstd
    unsigned* ticket;
    int dbg_delay_ns;                         // reproducer switch FCD_TCF_PRODUCER_DELAY_NS: slow the producers down
};

template <int CIN, int COUT>
struct Cfg {
    static constexpr int W_BYTES = 27 * CIN * COUT * 2;          // [khw][CIN/8][t = 2-kd][COUT][8]
    static constexpr int KHW_BYTES = 3 * CIN * COUT * 2;
    static constexpr int PLANE_BYTES = HV * CIN * 2;
    static constexpr int LBO_A = HV * 16, SBO_A = HW * 16;
    static constexpr int LBO_B = 3 * COUT * 16, SBO_B = 128;
    static constexpr int BUDGET = 220 * 1024 - W_BYTES;
    static constexpr int NST = BUDGET / PLANE_BYTES >= 6 ? 6 : BUDGET / PLANE_BYTES;
    static constexpr int TCOLS = 3 * R * COUT;                   // [kh set][slot][COUT]
    static constexpr int TMEM_COLS = TCOLS <= 256 ? 256 : 512;
    static constexpr int SMEM = W_BYTES + NST * PLANE_BYTES + 2560;
    static constexpr int CTAS_PER_SM = (COUT == 16 && 2 * SMEM <= 226 * 1024 && 2 * TMEM_COLS <= 512) ? 2 : 1;
    static_assert(NST >= 4 && TCOLS <= 512, "resources");   // NST > NMMA: see the FULL wait of the MMA warps
};

struct Item { int n, h0, w0, d0, d1, p_lo, p_hi, chunk; };
__device__ __forceinline__ Item decode(const ConvTcfParams& p, int item) {
    Item it;
    int wt = item % p.nwt; item /= p.nwt;
    int ht = item % p.nht; item /= p.nht;
    int seg = item % p.nseg; item /= p.nseg;
    it.n = item; it.h0 = ht * TH; it.w0 = wt * TW;
    it.d0 = seg * p.DL; it.d1 = min(it.d0 + p.DL, p.D);
    it.p_lo = max(it.d0 - 1, 0); it.p_hi = min(it.d1, p.D - 1);
    it.chunk = (seg * p.nht + ht) * p.nwt + wt;
    return it;
}

template <int CIN, int COUT, bool STATS, bool FLIP>
__global__ void __launch_bounds__(NTHREADS, Cfg<CIN, COUT>::CTAS_PER_SM) conv3_tcf_kernel(const ConvTcfParams p) {
    using K = Cfg<CIN, COUT>;
    constexpr int NST = K::NST;
    constexpr int DEPTH = NST >= 6 ? 3 : NST - 3;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* wsm = smem;
    unsigned char* ring = smem + K::W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::W_BYTES + NST * K::PLANE_BYTES);
    // bars: [0,NST) FULL | [NST,2NST) EMPTY | [2NST,2NST+NPB) PDONE | [2NST+NPB,2NST+NPB+R) TEMPTY
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + NPB + R);
    WaitCtx* ctx = reinterpret_cast<WaitCtx*>(tmem_slot + 4);
    float* red = reinterpret_cast<float*>(ctx + 1);
    float* sbias = red + 8 * COUT;
    // ctx->prog (debug record of a timed-out wait): [0] / [26] producer warp 0 / 1 plane seq, [1] / [27] its item;
    // [2+me] MMA warp's plane counter g, [5+me] its item, [8+me] (wait site << 24 | sq or zc; site 7 = issued g);
    // [11+q] epilogue outputs done, [15+q] plane it waits for (| 0x40000000: past the waits), [19+q] its item;
    // [23] nitems, [24] nseg, [25] DL

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    auto PDONE = [&](int s) { return bar0 + 8u * (2 * NST + s); };
    auto TEMPTY = [&](int s) { return bar0 + 8u * (2 * NST + NPB + s); };

    if (tid == 0) {
        wait_ctx_init(ctx, p.status, 2);
        ctx->prog[23] = p.nitems; ctx->prog[24] = p.nseg; ctx->prog[25] = p.DL;
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 32 * NPROD); mbar_init(EMPTY(s), 1); }
        for (int s = 0; s < NPB; ++s) mbar_init(PDONE(s), 1);
        for (int s = 0; s < R; ++s) mbar_init(TEMPTY(s), 4);
        fence_barrier_init();
    }
    if (warp == NPROD) tmem_alloc<K::TMEM_COLS>(smem_u32(tmem_slot));
    // weights: fp32 parameter -> bf16 smem [khw][k/8][t][n][8], t = 2 - kd (output planes ascending);
    // FLIP (data gradient): the tap read is the mirrored one
    {
        constexpr int CHUNKS = 27 * COUT * (CIN / 8);
        for (int i = tid; i < CHUNKS; i += NTHREADS) {
            const int tap = i % 27;                              // logical tap of the correlation being computed
            const int np_ = (i / 27) % COUT;
            const int c8 = i / (27 * COUT);
            const int src_tap = FLIP ? 26 - tap : tap;
            const int kd = tap / 9, khw = tap % 9;
            const int ns = np_ / p.nsgpad, nw = np_ % p.nsgpad;
            const int n = ns * p.nsg + nw;
            const bool nok = nw < p.nsg && n < p.Nr;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kp = c8 * 8 + j;
                const int ks = kp / p.ksegpad, kwi = kp % p.ksegpad;
                const int k = ks * p.kseg + kwi;
                f[j] = (nok && kwi < p.kseg && k < p.Kr) ? __ldg(p.Wf + n * p.sn + k * p.sk + src_tap * p.st) : 0.f;
            }
            *reinterpret_cast<bf16x8*>(wsm + khw * K::KHW_BYTES + c8 * K::LBO_B + ((2 - kd) * COUT + np_) * 16) = pack8(f);
        }
    }
    if (tid < COUT) sbias[tid] = p.bias != nullptr ? p.bias[tid] : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NPROD) {
        // ===================================================================== producers (identical to conv_tc.cu)
        constexpr int C8 = CIN / 8, VS = 32 * NPROD / C8, NJ = (HV + VS - 1) / VS;
        const int pt = warp * 32 + lane;
        const int c8 = pt % C8, v0 = pt / C8;
        const uint32_t ring_u = smem_u32(ring);
        uint32_t seq = 0, signaled = 0;
        auto flush_to = [&](uint32_t upto) {
            while (signaled < upto) { mbar_arrive(FULL(signaled % NST)); ++signaled; }
        };
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item);
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
                if (lane == 0) { prog_set(ctx, warp == 0 ? 0 : 26, (int)seq); prog_set(ctx, warp == 0 ? 1 : 27, item); }
                mbar_wait(EMPTY(s), ((seq / NST) & 1u) ^ 1u, ctx, 1, (int)seq);
                if (p.dbg_delay_ns > 0) __nanosleep((unsigned)p.dbg_delay_ns);
                const bf16* plane = p.A + ((long long)it.n * p.D + pl) * plane_elems;
                const uint32_t dst0 = ring_u + s * K::PLANE_BYTES + c8 * K::LBO_A + v0 * 16;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (j < NJ - 1 || v0 + j * VS < HV) {
                        const bool ok = off[j] >= 0;
                        cp_async16(dst0 + j * VS * 16, ok ? plane + off[j] : p.A, ok);
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
    } else if (warp < NPROD + NMMA) {
        // ===================================================================== MMA issuers: the three warps take turns on
        // PLANES: one warp pays a plane's waits / commits / bookkeeping once and issues all
        // 9*CIN/16 instructions of it, while the other two work on the neighbouring planes.  Each warp accumulates
        // into its own set of R slots (concurrent warps never share an accumulator); the epilogue adds the three sets.
        const int me = warp - NPROD;
        constexpr uint32_t A_HI = ((K::SBO_A >> 4) & 0x3fffu) | (1u << 14);
        constexpr uint32_t B_HI = ((K::SBO_B >> 4) & 0x3fffu) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(ring) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_A >> 4) << 16);
        const uint32_t b_lo0 = ((smem_u32(wsm) >> 4) & 0x3fffu) | ((uint32_t)(K::LBO_B >> 4) << 16);
        const uint32_t d_set = tmem_base + me * R * COUT;
        // n_out consecutive output slots starting at `slot`, whose first B row block is t0
        auto issue = [&](uint32_t a_pl, int slot, int t0, int n_out) {
            const uint32_t idesc = umma_idesc(128, n_out * COUT, 0, 0);
            const uint32_t d = d_set + slot * COUT;
            const uint32_t b_t = b_lo0 + ((t0 * COUT * 16) >> 4);
#pragma unroll
            for (int khw = 0; khw < 9; ++khw) {
                const int kh = khw / 3, kw = khw % 3;
#pragma unroll
                for (int kc = 0; kc < CIN / 16; ++kc) {
                    const uint32_t a_lo = a_pl + (((kh * HW + kw) * 16 + kc * 2 * K::LBO_A) >> 4);
                    const uint32_t b_lo = b_t + ((khw * K::KHW_BYTES + kc * 2 * K::LBO_B) >> 4);
                    umma_f16(d, ((uint64_t)A_HI << 32) | a_lo, ((uint64_t)B_HI << 32) | b_lo, idesc, 1u);
                }
            }
        };
        uint32_t seq_base = 0, zc0 = 0, pc0 = 0;               // loaded-plane sequence / output counter / plane counter
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item);
            const int nload = it.p_hi - it.p_lo + 1, DLi = it.d1 - it.d0;
            for (int i = 0; i <= DLi + 1; ++i) {               // input plane z = d0 - 1 + i feeds outputs i-2, i-1, i
                const uint32_t g = pc0 + i;
                const int pl = it.d0 - 1 + i;
                const bool valid = pl >= 0 && pl < p.D;
                // LOADED planes take turns by their position in the halo ring (sq % 3), zero-padding planes by the plane
                // counter.  A parity wait is only meaningful while the waiter is at most one phase behind the barrier:
                // waiting for plane sq on slot sq % NST presumes plane sq - NST has been loaded.  With turns by sq a
                // warp's consecutive FULL waits are exactly 3 ring positions apart, so the plane it consumed last
                // (sq - 3) vouches for plane sq - NST (NST >= 4 > 3, planes are handed over in order).  Round 1 took
                // turns by the plane COUNTER g: at an item boundary with exactly one padding plane (several d-segments
                // AND several items per CTA) the warp owning the padding plane then jumped 5 ring positions, one more
                // than the 4 stages of the 64 -> 32 configuration; "phase sq / NST done" was answered by plane
                // sq - 2 NST (same parity), the warp multiplied a stale plane and its early EMPTY arrival left the
                // producer one phase out of step -- the rare bad tiles and 0.2 s time-outs of sharded inference.
                const uint32_t sq = valid ? seq_base + (uint32_t)(pl - it.p_lo) : 0u;
                if ((int)((valid ? sq : g) % NMMA) != me) continue;
                const int j_lo = max(i - 2, 0), j_hi = min(i, DLi - 1);
                if (lane == 0) { prog_set(ctx, 2 + me, (int)g); prog_set(ctx, 5 + me, item); }
                if (valid) {
                    if (lane == 0) prog_set(ctx, 8 + me, (2 << 24) | (int)(sq & 0xffffff));
                    mbar_wait(FULL(sq % NST), (sq / NST) & 1u, ctx, 2, item);
                }
                // every output this plane touches must have been handed back (zeroed) by the epilogue; this also
                // keeps a warp from running more than R outputs ahead, i.e. from lapping the PDONE phases
                for (int j = j_lo; j <= j_hi; ++j) {
                    const uint32_t zc = zc0 + j;
                    if (lane == 0) prog_set(ctx, 8 + me, (3 << 24) | (int)(zc & 0xffffff));
                    mbar_wait(TEMPTY(zc % R), (zc / R) & 1u, ctx, 3, item);
                }
                tc_fence_after();
                if (lane == 0) {
                    if (valid) {
                        const uint32_t a_pl = a_lo0 + (sq % NST) * (K::PLANE_BYTES >> 4);
                        const int t_lo = j_lo - (i - 2), n_out = j_hi - j_lo + 1;
                        const int s_lo = (int)((zc0 + j_lo) % R);
                        const int n1 = min(n_out, R - s_lo);
                        issue(a_pl, s_lo, t_lo, n1);
                        if (n_out > n1) issue(a_pl, 0, t_lo + n1, n_out - n1);      // the range wraps around the ring
                        umma_commit(EMPTY(sq % NST));
                        umma_commit(PDONE(g % NPB));
                        prog_set(ctx, 8 + me, (7 << 24) | (int)(g & 0xffffff));
                    } else {
                        mbar_arrive(PDONE(g % NPB));           // zero-padding plane: nothing to add
                    }
                }
                __syncwarp();
            }
            seq_base += nload;
            zc0 += DLi;
            pc0 += DLi + 2;
        }
    } else {
        // ===================================================================== epilogue (4 warps)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int hh = r >> 3, ww = r & 7;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        // all accumulators start zeroed; completing phase 0 of every TEMPTY hands the slots to the MMA warps
#pragma unroll 1
        for (int c = 0; c < K::TCOLS; c += 16) tmem_st16_zero(trow + c);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            for (int s = 0; s < R; ++s) mbar_arrive(TEMPTY(s));
        }
        uint32_t zc = 0, pc0 = 0, pwaited = 0;              // outputs done / plane counter at item start / planes waited
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode(p, item);
            float s1[STATS ? COUT : 1], s2[STATS ? COUT : 1];
            if (STATS) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) s1[c] = s2[c] = 0.f;
            }
            for (int od = it.d0; od < it.d1; ++od, ++zc) {
                const int slot = zc % R;
                // output j = od - d0 is complete when planes j, j+1, j+2 of this item have been multiplied
                const uint32_t need = pc0 + (uint32_t)(od - it.d0) + 3;
                if (lane == 0) { prog_set(ctx, 11 + q, (int)zc); prog_set(ctx, 19 + q, item); }
                while (pwaited < need) {
                    if (lane == 0) prog_set(ctx, 15 + q, (int)pwaited);
                    mbar_wait(PDONE(pwaited % NPB), (pwaited / NPB) & 1u, ctx, 4, (int)pwaited);
                    ++pwaited;
                }
                if (lane == 0) prog_set(ctx, 15 + q, (int)pwaited | 0x40000000);      // past the waits of this output
                tc_fence_after();
                uint32_t t[3][COUT];
#pragma unroll
                for (int set = 0; set < 3; ++set)
#pragma unroll
                    for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16(trow + (set * R + slot) * COUT + c0, t[set] + c0);
                tmem_wait_ld();
#pragma unroll
                for (int set = 0; set < 3; ++set)
#pragma unroll
                    for (int c0 = 0; c0 < COUT; c0 += 16) tmem_st16_zero(trow + (set * R + slot) * COUT + c0);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(slot));
                bf16* dst = p.C + ((((long long)it.n * p.D + od) * p.H + it.h0 + hh) * p.W + it.w0 + ww) * p.ldc;
                if (!(it.h0 + hh < p.H && it.w0 + ww < p.W)) continue;     // ragged edge tile: H % 16 or W % 8 != 0
#pragma unroll
                for (int c0 = 0; c0 < COUT; c0 += 8) {
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        f[k] = ((__uint_as_float(t[0][c0 + k]) + __uint_as_float(t[1][c0 + k])) + __uint_as_float(t[2][c0 + k])) +
                               sbias[c0 + k];
                    if (p.accumulate) {
                        float o[8];
                        unpack8(ld8(dst + c0), o);
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] += o[k];
                    }
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
                const int e = tid - 32 * (NPROD + NMMA);
                if (e < 2 * COUT) {
                    const float tt = red[e] + red[2 * COUT + e] + red[4 * COUT + e] + red[6 * COUT + e];
                    const long long nchunk = (long long)p.nht * p.nwt * p.nseg;
                    p.part[((long long)it.n * nchunk + it.chunk) * 2 * COUT + e] = tt;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            pc0 += (uint32_t)(it.d1 - it.d0) + 2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPROD) tmem_dealloc<K::TMEM_COLS>(tmem_base);
    if (STATS && p.fin.mean != nullptr && lastblk::arrive(p.ticket, gridDim.x))
        norm_finalize_block(p.part, p.fin, reinterpret_cast<double*>(smem));      // all shared memory is free by now
}

template <int CIN, int COUT, bool STATS, bool FLIP>
int launch2(const ConvTcfParams& p, cudaStream_t stream) {
    using K = Cfg<CIN, COUT>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv3_tcf_kernel<CIN, COUT, STATS, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
        configured = true;
    }
    const int grid = min(p.nitems, fcd_num_sms() * K::CTAS_PER_SM);
    conv3_tcf_kernel<CIN, COUT, STATS, FLIP><<<grid, NTHREADS, K::SMEM, stream>>>(p);
    return (int)cudaGetLastError();
}

template <int CIN, int COUT>
int launch(const ConvTcfParams& p, int flip, cudaStream_t stream) {
    if (p.part != nullptr)
        return flip ? launch2<CIN, COUT, true, true>(p, stream) : launch2<CIN, COUT, true, false>(p, stream);
    return flip ? launch2<CIN, COUT, false, true>(p, stream) : launch2<CIN, COUT, false, false>(p, stream);
}

}  // namespace

// Same contract as fcd_conv3_tc (see conv_tc.cu / the header); takes K in {16,32,64} x N in {16,32}.  Returns -1 for
// anything else so the caller can fall back to fcd_conv3_tc.
FCD_API int fcd_conv3_tcf(const void* A, long long lda, const float* Wf, int Nr, int Kr, long long sn, long long sk,
                          long long st, int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, float* part,
                          const float* bias, int accumulate, int Bn, int D, int H, int W, int K, int N, int flip, int nseg,
                          float* mean, float* rstd,
                          int norm_mode, float eps, float* running_mean, float* running_var, int crun, float momentum,
                          cudaStream_t stream) {
    if (H < 1 || W < 1 || D < 1 || nseg < 1 || lda % 8 || ldc % 8 || lda < K || ldc < N) return -1;
    if (!(K == 16 || K == 32 || K == 64) || !(N == 16 || N == 32)) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)C & 15) || kseg < 1 || ksegpad < 1 || nsg < 1 || nsgpad < 1) return -1;
    ConvTcfParams p;
    p.A = (const bf16*)A; p.lda = lda; p.Wf = Wf; p.sn = sn; p.sk = sk; p.st = st; p.Nr = Nr; p.Kr = Kr;
    p.kseg = kseg; p.ksegpad = ksegpad; p.nsg = nsg; p.nsgpad = nsgpad;
    p.C = (bf16*)C; p.ldc = ldc; p.part = part; p.bias = bias; p.accumulate = accumulate;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W;
    p.nht = (H + TH - 1) / TH; p.nwt = (W + TW - 1) / TW; p.DL = (D + nseg - 1) / nseg;
    p.nseg = (D + p.DL - 1) / p.DL;
    if (p.nseg != nseg) return -1;
    p.nitems = Bn * p.nht * p.nwt * p.nseg;
    p.status = fcd_status_dev();
    // optional: the last CTA finishes the fused statistics (mean / rstd of the norm that follows the conv); the caller
    // asks fcd_norm_fin_fold first -- partials too large for one CTA are finished by fcd_norm_finalize
    if (mean != nullptr && (part == nullptr || !fin_fold(Bn, p.nht * p.nwt * p.nseg, 2 * N))) return -1;
    p.fin = NormFin{mean, rstd, running_mean, running_var, Bn, N, p.nht * p.nwt * p.nseg,
                    norm_mode, crun, (long long)D * H * W, eps, momentum};
    p.ticket = p.fin.mean != nullptr ? lastblk::next_ticket() : nullptr;
    static const int dbg_delay = getenv("FCD_TCF_PRODUCER_DELAY_NS") ? atoi(getenv("FCD_TCF_PRODUCER_DELAY_NS")) : 0;
    p.dbg_delay_ns = dbg_delay;
#define FCD_TCF_CASE(CI, CO) if (K == CI && N == CO) return launch<CI, CO>(p, flip, stream)
    FCD_TCF_CASE(16, 16); FCD_TCF_CASE(32, 16); FCD_TCF_CASE(64, 16);
    FCD_TCF_CASE(16, 32); FCD_TCF_CASE(32, 32); FCD_TCF_CASE(64, 32);
#undef FCD_TCF_CASE
    return -1;
}
