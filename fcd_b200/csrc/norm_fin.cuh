// Shared by elementwise.cu and the tcgen05 conv kernels (fused statistics epilogue): turning the per-chunk partial rows
// part[B][nchunk][V][C] (V = 2: sum, sum of squares; V = 3: the three backward sums) into per-(b, c) results.
//
// Two forms.  (1) `*_warp`: one warp per (b, c), for a stand-alone finalize kernel spread over many blocks -- used when
// the partials are large.  (2) `*_block`: ONE block (the last one of the producing kernel to finish, last_block.cuh)
// does everything; the column sums are then taken by all its threads with independent 16-byte loads (a warp-per-pair
// loop in a single block is a chain of dependent L2 latencies: measured 2x slower than the separate launch), in a fixed
// order (deterministic), in fp64, staged through 16 KB of shared memory, one channel tile at a time.
#pragma once
#include <cstdlib>

#include "common.cuh"

constexpr int kFinScratchDoubles = 2048;           // 16 KB: k-split exchange [0, 1024) + tile totals [1024, 2048)
constexpr long long kFinFoldBytes = 128 << 10;     // partials up to this size are finished by the last block

static inline bool fin_fold(int B, int nchunk, int L) {
    static const long long limit = getenv("FCD_FIN_FOLD_BYTES") ? atoll(getenv("FCD_FIN_FOLD_BYTES")) : kFinFoldBytes;
    return (long long)B * nchunk * L * 4 <= limit && B * 3 * 4 <= 1024;
}

// mode 0 instance (per b,c), 1 batch (per c over b), 2 group-of-2-channels (per b, c/2).
// Writes mean[b][c], rstd[b][c]; for batch mode also updates running stats (momentum, unbiased var) if given.
struct NormFin {
    float* mean; float* rstd;
    float* running_mean; float* running_var;
    int B, C, nchunk, mode, crun;
    long long S;
    float eps, momentum;
};

__device__ __forceinline__ void norm_fin_write(const NormFin& f, int b, int c, double s, double q, double n) {
    const double m = s / n;
    double var = q / n - m * m;
    if (var < 0.0) var = 0.0;
    f.mean[b * f.C + c] = (float)m;
    f.rstd[b * f.C + c] = (float)(1.0 / sqrt(var + (double)f.eps));
    if (f.mode == 1 && b == 0 && f.running_mean != nullptr && c < f.crun) {
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)m;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
    }
}

// (1) one WARP per (b, c), warps w0, w0 + nw, ...: the lanes stride over the chunk partials
__device__ __forceinline__ void norm_finalize_warp(const float* __restrict__ part, const NormFin& f, int w0, int nw) {
    const int lane = threadIdx.x & 31;
    const int B = f.B, C = f.C, nchunk = f.nchunk;
    for (int i = w0; i < B * C; i += nw) {
        const int b = i / C, c = i % C;
        double s = 0.0, q = 0.0, n = 0.0;
        auto add = [&](int bb, int cc) {
            const float* o = part + (long long)bb * nchunk * 2 * C;
            for (int k = lane; k < nchunk; k += 32) {
                s += __ldcg(o + (long long)k * 2 * C + cc);
                q += __ldcg(o + (long long)k * 2 * C + C + cc);
            }
            n += (double)f.S;
        };
        if (f.mode == 0) {
            add(b, c);
        } else if (f.mode == 1) {
            for (int bb = 0; bb < B; ++bb) add(bb, c);
        } else {
            add(b, c & ~1);
            add(b, c | 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) norm_fin_write(f, b, c, s, q, n);
    }
}

// Column sums of channel tile [c0, c0 + CT) over the nchunk rows: tot[(b * V + v) * CT + cc], cc < CT (CT % 4 == 0).
// Every thread of the block must call it; ends with a __syncthreads().
template <int V>
__device__ __forceinline__ void fin_colsum_tile(const float* __restrict__ part, int B, int nchunk, int C, int c0, int CT,
                                                double* scratch) {
    const int nthr = min((int)blockDim.x, 256), tid = threadIdx.x;   // workers (the exchange area holds 256 x 4 doubles)
    const bool worker = tid < nthr;
    const int Q = CT / 4, T = B * V * Q;               // 16-byte column tasks
    double* xch = scratch;
    double* tot = scratch + 1024;
    const long long rs = (long long)V * C;             // floats per partial row
    int G = 1;                                         // k-split: G threads per task when there are few tasks
    if (T < nthr) { G = nthr / T; if (G > nchunk) G = nchunk; }
    __syncthreads();                                   // scratch may alias memory the caller was still using
    if (G > 1) {
        const int t = tid % T, g = tid / T;
        const int b = t / (V * Q), v = (t / Q) % V, q = t % Q;
        if (worker && g < G) {
            const float* col = part + (long long)b * nchunk * rs + (long long)v * C + c0 + q * 4;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
            for (int k = g; k < nchunk; k += G) {
                const float4 x = __ldcg(reinterpret_cast<const float4*>(col + (long long)k * rs));
                a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
            }
            double* o = xch + (long long)(g * T + t) * 4;
            o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
        }
        __syncthreads();
        if (tid < T) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int gg = 0; gg < G; ++gg) {           // fixed order
                const double* o = xch + (long long)(gg * T + tid) * 4;
                a0 += o[0]; a1 += o[1]; a2 += o[2]; a3 += o[3];
            }
            double* d = tot + (long long)(b * V + v) * CT + q * 4;
            d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
        }
    } else {
        for (int t = worker ? tid : T; t < T; t += nthr) {
            const int b = t / (V * Q), v = (t / Q) % V, q = t % Q;
            const float* col = part + (long long)b * nchunk * rs + (long long)v * C + c0 + q * 4;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
            for (int k = 0; k < nchunk; ++k) {
                const float4 x = __ldcg(reinterpret_cast<const float4*>(col + (long long)k * rs));
                a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
            }
            double* d = tot + (long long)(b * V + v) * CT + q * 4;
            d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int fin_tile_channels(int B, int V, int C) {
    int ct = (1024 / (B * V)) & ~3;
    return ct < C ? ct : C;
}

// (2) the whole finalize by ONE block (all threads call it).  scratch: kFinScratchDoubles doubles of shared memory.
__device__ __forceinline__ void norm_finalize_block(const float* __restrict__ part, const NormFin& f, double* scratch) {
    const int B = f.B, C = f.C;
    const int CT0 = fin_tile_channels(B, 2, C);
    const double* tot = scratch + 1024;
    for (int c0 = 0; c0 < C; c0 += CT0) {
        const int CT = min(CT0, C - c0);
        fin_colsum_tile<2>(part, B, f.nchunk, C, c0, CT, scratch);
        for (int i = threadIdx.x; i < B * CT; i += blockDim.x) {
            const int b = i / CT, cc = i % CT;
            double s = 0.0, q = 0.0, n = 0.0;
            auto add = [&](int bb, int c2) {
                s += tot[(long long)(bb * 2 + 0) * CT + c2];
                q += tot[(long long)(bb * 2 + 1) * CT + c2];
                n += (double)f.S;
            };
            if (f.mode == 0) add(b, cc);
            else if (f.mode == 1) { for (int bb = 0; bb < B; ++bb) add(bb, cc); }
            else { add(b, cc & ~1); add(b, cc | 1); }
            norm_fin_write(f, b, c0 + cc, s, q, n);
        }
    }
}
