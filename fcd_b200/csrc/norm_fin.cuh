// Shared by elementwise.cu and the tcgen05 conv kernels (fused statistics epilogue): turning the per-chunk partial
// (sum, sum of squares) rows part[B][nchunk][2][C] into mean / rstd, run by the LAST block of the producing kernel.
#pragma once
#include "common.cuh"

// mode 0 instance (per b,c), 1 batch (per c over b), 2 group-of-2-channels (per b, c/2).
// Writes mean[b][c], rstd[b][c]; for batch mode also updates running stats (momentum, unbiased var) if given.
// One WARP per (b, c), warps `w0`, `w0 + nw`, ...: the lanes stride over the chunk partials (fixed order =>
// deterministic), fp64 accumulation.
struct NormFin {
    float* mean; float* rstd;
    float* running_mean; float* running_var;
    int B, C, nchunk, mode, crun;
    long long S;
    float eps, momentum;
};
__device__ __forceinline__ void norm_finalize_body(const float* __restrict__ part, const NormFin& f, int w0, int nw) {
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
        if (lane != 0) continue;
        const double m = s / n;
        double var = q / n - m * m;
        if (var < 0.0) var = 0.0;
        f.mean[i] = (float)m;
        f.rstd[i] = (float)(1.0 / sqrt(var + (double)f.eps));
        if (f.mode == 1 && b == 0 && f.running_mean != nullptr && c < f.crun) {
            const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
            f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)m;
            f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
        }
    }
}

