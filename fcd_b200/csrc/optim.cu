// AdamW over ALL parameter tensors of a model in ONE launch (train_utils.py:63-71 builds torch.optim.AdamW(lr 1e-4,
// weight_decay 1e-5); train.py:382 steps it).  torch's fused implementation chunks the ~244 tensors of MS_DSA_NET into
// 10 multi_tensor_apply launches (0.34 ms per step); here a device job table (pointer quadruples + element counts) lets
// one grid walk every tensor: 43.5 M parameters x (4 reads + 3 writes) x 4 B = 1.2 GB, HBM-bound.
// Arithmetic = torch.optim.AdamW (decoupled weight decay, bias-corrected, eps added after the square root), fp32.
#include "common.cuh"

namespace {

struct AdamJob {            // 48 bytes, mirrored by fcd_b200/optim.py
    float* p; const float* g; float* m; float* v;
    long long n;            // elements
    long long blk0;         // first block of this tensor
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = ADAM_THREADS * 16;      // elements per block

__global__ void __launch_bounds__(ADAM_THREADS) adamw_multi_kernel(const AdamJob* __restrict__ jobs, int njobs,
                                                                   const float* __restrict__ step, double lr, double beta1,
                                                                   double beta2, float eps, double wd) {
    __shared__ AdamJob job;
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        int lo = 0, hi = njobs - 1;                 // last job whose blk0 <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].blk0 <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        job = jobs[lo];
        const double t = (double)step[0];
        const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
        s_step_size = (float)(lr / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const long long base = ((long long)blockIdx.x - job.blk0) * ADAM_CHUNK;
    // scalar factors formed in double and rounded once, as torch does with its Python-float hyper-parameters
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt, decay = (float)(1.0 - lr * wd);
    const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), b2 = (float)beta2;
    const bool vec = ((((uintptr_t)job.p | (uintptr_t)job.g | (uintptr_t)job.m | (uintptr_t)job.v) & 15) == 0);
    auto upd = [&](float& p, float g, float& m, float& v) {
        p *= decay;
        m = m + (g - m) * omb1;                               // exp_avg.lerp_(grad, 1 - beta1)
        v = v * b2 + omb2 * (g * g);                          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p -= step_size * (m / denom);
    };
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long i = base + ((long long)r * ADAM_THREADS + threadIdx.x) * 4;
        if (i >= job.n) break;
        if (vec && i + 4 <= job.n) {
            float4 p = *reinterpret_cast<float4*>(job.p + i), m = *reinterpret_cast<float4*>(job.m + i),
                   v = *reinterpret_cast<float4*>(job.v + i);
            const float4 g = *reinterpret_cast<const float4*>(job.g + i);
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(job.p + i) = p;
            *reinterpret_cast<float4*>(job.m + i) = m;
            *reinterpret_cast<float4*>(job.v + i) = v;
        } else {
            for (long long j = i; j < min(i + 4, job.n); ++j) upd(job.p[j], job.g[j], job.m[j], job.v[j]);
        }
    }
}

}  // namespace

// elements one block updates (the host sizes blk0 / nblocks with it)
FCD_API int fcd_adamw_chunk(void) { return ADAM_CHUNK; }

// jobs: device array of njobs {p, g, m, v, n, blk0} records (48 bytes each, blk0 ascending from 0), nblocks = total
// blocks; step: device fp32 scalar holding the 1-based step count of THIS update (the host increments it first).
FCD_API int fcd_adamw_multi(const void* jobs, int njobs, int nblocks, const float* step, double lr, double beta1,
                            double beta2, double eps, double weight_decay, cudaStream_t stream) {
    if (njobs < 1 || nblocks < 1 || jobs == nullptr || step == nullptr) return -1;
    adamw_multi_kernel<<<nblocks, ADAM_THREADS, 0, stream>>>(static_cast<const AdamJob*>(jobs), njobs, step, lr, beta1,
                                                             beta2, (float)eps, weight_decay);
    FCD_LAUNCH_CHECK();
}
