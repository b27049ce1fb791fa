// Fused segmentation loss: Dice | generalized Dice (+ CE | + sigmoid-focal) (+ total variation), forward and backward.
//
// Replaces, for a 2-class prediction, the ~15 ATen kernels of the reference's CombinedLoss.forward
// (get_loss.py:24-39): MONAI DiceLoss / DiceCELoss / DiceFocalLoss built at get_loss.py:46-78
// (include_background=False, to_onehot_y=True, softmax=True, batch=True, smooth 1e-5), GeneralizedDiceLoss /
// GeneralizedDiceFocalLoss built at get_loss.py:79-93 (include_background=True: both channels, class weights
// 1/G_c^2 | 1/G_c | 1 from the label counts) and compute_total_variation_loss / dilate_mask (get_loss.py:100-165).
// pred: fp32 NCDHW [B,2,D,H,W] logits; target: fp32 [B,1,D,H,W] in {0,1}.  All reductions are fp32 per thread,
// combined in double by a single finalize block; nothing is read back to the host.
#include "common.cuh"

struct LossCfg {
    int kind;        // 0 Dice, 1 DiceCE, 2 DiceFocal, 3 GeneralizedDice, 4 GeneralizedDiceFocal
    float lambda_dice, lambda_2, w_bg, w_fg, gamma;
    int squared, jaccard;
    int w_type;      // kinds 3, 4: class weight 0 = 1/G^2 ("square"), 1 = 1/G ("simple"), 2 = 1 ("uniform")
    float smooth_nr, smooth_dr;
    float tv_w;
    int tv_norm, tv_exclude;
};

namespace {

constexpr int LOSS_BLOCKS = 592;   // 4 x 148
constexpr int LOSS_THREADS = 256;

// streaming (read-once) 16-byte fp32 load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float softplusf(float x) { return x > 0.f ? x + log1pf(__expf(-x)) : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// MONAI sigmoid focal term of one logit x with label t in {0,1}: exp(gamma * logsigmoid(-x (2t-1))) * bce(x, t)
__device__ __forceinline__ float focal_term(float x, float t, float gamma) {
    const float bce = x - x * t + softplusf(-x);                // x - x t - logsigmoid(x)
    const float invp = -softplusf(x * (2.f * t - 1.f));         // logsigmoid(-x (2t-1))
    return __expf(gamma * invp) * bce;
}
__device__ __forceinline__ float focal_grad(float x, float t, float gamma) {    // d focal_term / dx
    const float sgn = 2.f * t - 1.f;
    const float u = -x * sgn;
    const float bce = x - x * t + softplusf(-x);
    const float e = __expf(gamma * -softplusf(-u));
    return e * (gamma * (-sgn) * sigmoidf_(-u) * bce + (sigmoidf_(x) - t));
}

// border mask of get_loss.py:141-150: 1 where the clipped 5^3 window holds both a 0 and a 1 of gt.
__global__ void tv_mask_kernel(const float* __restrict__ gt, unsigned char* __restrict__ keep, int B, int D, int H,
                               int W) {
    const long long total = (long long)B * D * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int x = (int)(r % W); r /= W;
        int y = (int)(r % H); r /= H;
        int z = (int)(r % D);
        long long b = r / D;
        bool any1 = false, any0 = false;
        for (int dz = -2; dz <= 2; ++dz) {
            int zz = z + dz; if (zz < 0 || zz >= D) continue;
            for (int dy = -2; dy <= 2; ++dy) {
                int yy = y + dy; if (yy < 0 || yy >= H) continue;
                for (int dx = -2; dx <= 2; ++dx) {
                    int xx = x + dx; if (xx < 0 || xx >= W) continue;
                    float g = gt[((b * D + zz) * H + yy) * W + xx];
                    any1 |= (g > 0.f);
                    any0 |= (g < 1.f);
                }
            }
        }
        keep[i] = (any1 && any0) ? 0 : 1;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_fwd_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ target, long long S, int B,
                                                                LossCfg cfg, const unsigned char* __restrict__ keep,
                                                                float* __restrict__ pbuf, float* __restrict__ part) {
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // I, P, G, CE numerator, CE denominator, focal sum
    const long long total = (long long)B * S;
    auto voxel = [&](float l0, float l1, float t, long long i) {
        const float y = t == 1.f ? 1.f : 0.f;
        const float d = l1 - l0;
        const float p = sigmoidf_(d);
        a[0] += p * y;
        a[1] += (cfg.squared && cfg.kind < 3) ? p * p : p;
        a[2] += y;
        if (cfg.kind == 1) {
            const float w = y > 0.f ? cfg.w_fg : cfg.w_bg;
            a[3] += w * softplusf(y > 0.f ? -d : d);
            a[4] += w;
        } else if (cfg.kind == 2) {
            a[5] += focal_term(l1, y, cfg.gamma);
        } else if (cfg.kind == 4) {                                  // include_background: both one-hot channels
            a[5] += focal_term(l1, y, cfg.gamma) + focal_term(l0, 1.f - y, cfg.gamma);
        }
        if (pbuf) pbuf[i] = keep ? (keep[i] ? p : 0.f) : p;
    };
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gn = (long long)gridDim.x * blockDim.x;
    if ((S & 3) == 0 && (((uintptr_t)pred | (uintptr_t)target) & 15) == 0) {
        // four consecutive voxels per thread: 16-byte loads of both logit planes and of the label (3 independent 16 B
        // loads in flight per thread instead of 3 scalar ones; one 64-bit division per 4 voxels)
        // two such groups per iteration: six independent 16 B loads in flight per thread before the first
        // transcendental (the kernel moves 100 MB in ~15 us at the HBM roofline: latency, not arithmetic, is what it hides)
        const long long S4 = S >> 2, n4 = (long long)B * S4;
        for (long long j = gtid; j < n4; j += 2 * gn) {
            const long long j2 = j + gn;
            const bool two = j2 < n4;
            const long long b = j / S4, s = (j - b * S4) << 2;
            const long long b2 = two ? j2 / S4 : b, s2 = two ? (j2 - b2 * S4) << 2 : s;
            const float4 l0 = ldg_stream4(pred + (b * 2) * S + s);
            const float4 l1 = ldg_stream4(pred + (b * 2 + 1) * S + s);
            const float4 t = ldg_stream4(target + b * S + s);
            const float4 m0 = ldg_stream4(pred + (b2 * 2) * S + s2);
            const float4 m1 = ldg_stream4(pred + (b2 * 2 + 1) * S + s2);
            const float4 u = ldg_stream4(target + b2 * S + s2);
            const long long i = b * S + s;
            voxel(l0.x, l1.x, t.x, i); voxel(l0.y, l1.y, t.y, i + 1); voxel(l0.z, l1.z, t.z, i + 2); voxel(l0.w, l1.w, t.w, i + 3);
            if (two) {
                const long long i2 = b2 * S + s2;
                voxel(m0.x, m1.x, u.x, i2); voxel(m0.y, m1.y, u.y, i2 + 1); voxel(m0.z, m1.z, u.z, i2 + 2);
                voxel(m0.w, m1.w, u.w, i2 + 3);
            }
        }
    } else {
        for (long long i = gtid; i < total; i += gn) {
            const long long b = i / S, s = i - b * S;
            voxel(pred[(b * 2) * S + s], pred[(b * 2 + 1) * S + s], target[i], i);
        }
    }
    __shared__ float sh[(LOSS_THREADS / 32) * 6];
    block_sum<6>(a, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) part[blockIdx.x * 8 + k] = a[k];
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) tv_fwd_kernel(const float* __restrict__ pbuf, int B, int D, int H,
                                                              int W, int norm, float* __restrict__ part) {
    float a[3] = {0.f, 0.f, 0.f};
    const long long total = (long long)B * D * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int x = (int)(r % W); r /= W;
        int y = (int)(r % H); r /= H;
        int z = (int)(r % D);
        const float p = pbuf[i];
        if (z + 1 < D) { float dd = pbuf[i + (long long)H * W] - p; a[0] += norm == 1 ? fabsf(dd) : dd * dd; }
        if (y + 1 < H) { float dd = pbuf[i + W] - p; a[1] += norm == 1 ? fabsf(dd) : dd * dd; }
        if (x + 1 < W) { float dd = pbuf[i + 1] - p; a[2] += norm == 1 ? fabsf(dd) : dd * dd; }
    }
    __shared__ float sh[(LOSS_THREADS / 32) * 3];
    block_sum<3>(a, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 4 + 0] = a[0];
        part[blockIdx.x * 4 + 1] = a[1];
        part[blockIdx.x * 4 + 2] = a[2];
    }
}

// res[0]=total, [1]=dice N, [2]=dice Dn, [3]=CE weight sum, [4..6]=tv_z,tv_y,tv_x, [7]=dice, [8]=ce|focal, [9]=tv,
// [10],[11] = generalized-Dice class weights w_bg, w_fg
__global__ void loss_finalize_kernel(const float* __restrict__ part, const float* __restrict__ tvpart, int nblk,
                                     LossCfg cfg, int B, int D, int H, int W, float* __restrict__ res,
                                     const int* __restrict__ status) {
    __shared__ double sh[9];
    if (threadIdx.x < 9) {
        double s = 0.0;
        if (threadIdx.x < 6) {
            for (int k = 0; k < nblk; ++k) s += part[k * 8 + threadIdx.x];
        } else if (tvpart != nullptr) {
            for (int k = 0; k < nblk; ++k) s += tvpart[k * 4 + (threadIdx.x - 6)];
        }
        sh[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double I = sh[0], P = sh[1], G = sh[2];
        const double vox = (double)B * D * H * W;
        double N, Dn, gw[2] = {0.0, 0.0};
        if (cfg.kind >= 3) {
            // MONAI GeneralizedDiceLoss.forward with batch=True over the channels (1-p, p) / (1-y, y): the background
            // sums follow from the foreground ones.  w = 1/G^2 | 1/G | 1; an infinite weight (empty class) is replaced
            // by the largest finite one.
            const double Ic[2] = {vox - P - G + I, I}, Gc[2] = {vox - G, G}, Pc[2] = {vox - P, P};
            bool inf[2];
            double wmax = 0.0;
            for (int c = 0; c < 2; ++c) {
                inf[c] = cfg.w_type != 2 && Gc[c] == 0.0;
                if (!inf[c]) gw[c] = cfg.w_type == 0 ? 1.0 / (Gc[c] * Gc[c]) : (cfg.w_type == 1 ? 1.0 / Gc[c] : 1.0);
                // the reference computes w in fp32 (ground_o.float()): round the same way
                gw[c] = (double)(float)gw[c];
                if (!inf[c]) wmax = fmax(wmax, gw[c]);
            }
            for (int c = 0; c < 2; ++c) if (inf[c]) gw[c] = wmax;
            N = 2.0 * (Ic[0] * gw[0] + Ic[1] * gw[1]) + cfg.smooth_nr;
            Dn = (Gc[0] + Pc[0]) * gw[0] + (Gc[1] + Pc[1]) * gw[1] + cfg.smooth_dr;
        } else {
            N = 2.0 * I + cfg.smooth_nr;
            double base = G + P;
            if (cfg.jaccard) base = 2.0 * (base - I);
            Dn = base + cfg.smooth_dr;
        }
        const double dice = 1.0 - N / Dn;
        double second = 0.0;
        if (cfg.kind == 1) second = sh[3] / sh[4];
        else if (cfg.kind == 2) second = sh[5] / vox;
        else if (cfg.kind == 4) second = sh[5] / (2.0 * vox);
        double tv[3] = {0.0, 0.0, 0.0}, tvsum = 0.0;
        if (tvpart != nullptr) {
            const double cnt[3] = {(double)B * (D - 1) * H * W, (double)B * D * (H - 1) * W, (double)B * D * H * (W - 1)};
            for (int k = 0; k < 3; ++k) {
                tv[k] = cfg.tv_norm == 1 ? sh[6 + k] / cnt[k] : sqrt(sh[6 + k] / cnt[k] + 1e-10);
                tvsum += tv[k];
            }
        }
        double total = ((cfg.kind == 0 || cfg.kind == 3) ? dice : cfg.lambda_dice * dice + cfg.lambda_2 * second) +
                       cfg.tv_w * tvsum;
        // a tcgen05 pipeline wait timed out somewhere upstream (status word, csrc/status.cu): the activations this loss
        // was computed from may contain a bad tile -- report NaN instead of a plausible number
        if (status != nullptr && *reinterpret_cast<const volatile int*>(status) != 0) total = nan("");
        res[0] = (float)total; res[1] = (float)N; res[2] = (float)Dn; res[3] = (float)sh[4];
        res[4] = (float)tv[0]; res[5] = (float)tv[1]; res[6] = (float)tv[2];
        res[7] = (float)dice; res[8] = (float)second; res[9] = (float)tvsum;
        res[10] = (float)gw[0]; res[11] = (float)gw[1];
    }
}

// dpred (fp32 NCDHW) = gout * dLoss/dpred
__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ target, int B, int D, int H,
                                                                int W, LossCfg cfg,
                                                                const unsigned char* __restrict__ keep,
                                                                const float* __restrict__ pbuf,
                                                                const float* __restrict__ res,
                                                                const float* __restrict__ gout,
                                                                float* __restrict__ dpred) {
    const long long S = (long long)D * H * W;
    const long long total = (long long)B * S;
    const float go = gout[0];
    const float N = res[1], Dn = res[2], cew = res[3];
    const float ldice = (cfg.kind == 0 || cfg.kind == 3) ? 1.f : cfg.lambda_dice;
    const float gw0 = res[10], gw1 = res[11];
    const float cntz = (float)((double)B * (D - 1) * H * W), cnty = (float)((double)B * D * (H - 1) * W),
                cntx = (float)((double)B * D * H * (W - 1));
    // gradient pair (d/dl0, d/dl1) of one voxel
    auto voxel = [&](float l0, float l1, float t, long long i, long long s, float& g0, float& g1) {
        const float y = t == 1.f ? 1.f : 0.f;
        const float d = l1 - l0;
        const float p = sigmoidf_(d);
        // dice: f = 1 - N/Dn
        float dN, dDn;                                                   // d numerator / dp, d denominator / dp
        if (cfg.kind >= 3) {
            dN = 2.f * (gw1 * y - gw0 * (1.f - y));
            dDn = gw1 - gw0;
        } else {
            const float dP = cfg.squared ? 2.f * p : 1.f;
            dN = 2.f * y;
            dDn = cfg.jaccard ? 2.f * (dP - y) : dP;
        }
        float gp = ldice * (-(dN * Dn - N * dDn) / (Dn * Dn));           // dL/dp
        if (pbuf != nullptr) {                                           // total variation on p' = p * keep
            long long r = s;
            const int x = (int)(r % W); r /= W;
            const int yy = (int)(r % H);
            const int z = (int)(r / H);
            const float pc = pbuf[i];
            float gt = 0.f;
            auto term = [&](float diff, float cnt, float tvv) -> float {
                // d/d(diff) of mean|diff| or sqrt(mean diff^2 + eps)
                if (cfg.tv_norm == 1) return (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) / cnt;
                return diff / (cnt * tvv);
            };
            if (z > 0) gt += term(pc - pbuf[i - (long long)H * W], cntz, res[4]);
            if (z + 1 < D) gt -= term(pbuf[i + (long long)H * W] - pc, cntz, res[4]);
            if (yy > 0) gt += term(pc - pbuf[i - W], cnty, res[5]);
            if (yy + 1 < H) gt -= term(pbuf[i + W] - pc, cnty, res[5]);
            if (x > 0) gt += term(pc - pbuf[i - 1], cntx, res[6]);
            if (x + 1 < W) gt -= term(pbuf[i + 1] - pc, cntx, res[6]);
            if (keep != nullptr && !keep[i]) gt = 0.f;
            gp += cfg.tv_w * gt;
        }
        float gd = gp * p * (1.f - p);                                   // through p = sigmoid(l1 - l0)
        float g1_extra = 0.f, g0_extra = 0.f;
        if (cfg.kind == 1) {
            const float w = y > 0.f ? cfg.w_fg : cfg.w_bg;
            gd += cfg.lambda_2 * w * (p - y) / cew;
        } else if (cfg.kind == 2) {
            g1_extra = cfg.lambda_2 * focal_grad(l1, y, cfg.gamma) / (float)total;
        } else if (cfg.kind == 4) {
            g1_extra = cfg.lambda_2 * focal_grad(l1, y, cfg.gamma) / (2.f * (float)total);
            g0_extra = cfg.lambda_2 * focal_grad(l0, 1.f - y, cfg.gamma) / (2.f * (float)total);
        }
        g0 = go * (g0_extra - gd);
        g1 = go * (gd + g1_extra);
    };
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gn = (long long)gridDim.x * blockDim.x;
    if ((S & 3) == 0 && (((uintptr_t)pred | (uintptr_t)target | (uintptr_t)dpred) & 15) == 0) {
        const long long S4 = S >> 2;            // four consecutive voxels per thread, 16-byte loads and stores
        for (long long j = gtid; j < (long long)B * S4; j += gn) {
            const long long b = j / S4, s = (j - b * S4) << 2;
            const float4 l0 = *reinterpret_cast<const float4*>(pred + (b * 2) * S + s);
            const float4 l1 = *reinterpret_cast<const float4*>(pred + (b * 2 + 1) * S + s);
            const float4 t = *reinterpret_cast<const float4*>(target + b * S + s);
            const long long i = b * S + s;
            float4 g0, g1;
            voxel(l0.x, l1.x, t.x, i, s, g0.x, g1.x); voxel(l0.y, l1.y, t.y, i + 1, s + 1, g0.y, g1.y);
            voxel(l0.z, l1.z, t.z, i + 2, s + 2, g0.z, g1.z); voxel(l0.w, l1.w, t.w, i + 3, s + 3, g0.w, g1.w);
            *reinterpret_cast<float4*>(dpred + (b * 2) * S + s) = g0;
            *reinterpret_cast<float4*>(dpred + (b * 2 + 1) * S + s) = g1;
        }
    } else {
        for (long long i = gtid; i < total; i += gn) {
            const long long b = i / S, s = i - b * S;
            float g0, g1;
            voxel(pred[(b * 2) * S + s], pred[(b * 2 + 1) * S + s], target[i], i, s, g0, g1);
            dpred[(b * 2) * S + s] = g0;
            dpred[(b * 2 + 1) * S + s] = g1;
        }
    }
}

// F.mse_loss(net_input, x_vae) of the VAE branch (segresnet_dsa.py:357)
__global__ void __launch_bounds__(LOSS_THREADS) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               long long n, float* __restrict__ part) {
    float s[1] = {0.f};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        s[0] = fmaf(d, d, s[0]);
    }
    __shared__ float sh[LOSS_THREADS / 32];
    block_sum<1>(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

__global__ void mse_finalize_kernel(const float* __restrict__ part, int nblk, long long n, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < nblk; ++k) s += part[k];
        out[0] = (float)(s / (double)n);
    }
}

__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               const float* __restrict__ gout, float* __restrict__ da) {
    const float g = gout[0] * 2.f / (float)n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        da[i] = g * (a[i] - b[i]);
}

}  // namespace

FCD_API int fcd_mse_fwd(const float* a, const float* b, long long n, float* part, float* out, cudaStream_t st) {
    mse_fwd_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(a, b, n, part);
    mse_finalize_kernel<<<1, 32, 0, st>>>(part, LOSS_BLOCKS, n, out);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_mse_bwd(const float* a, const float* b, long long n, const float* gout, float* da, cudaStream_t st) {
    mse_bwd_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(a, b, n, gout, da);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_loss_blocks() { return LOSS_BLOCKS; }

// Workspace: part >= LOSS_BLOCKS*8 floats, tvpart >= LOSS_BLOCKS*4 floats (or null), pbuf B*S floats (or null
// when tv_w == 0), keep B*S bytes (only when tv_exclude), res >= 16 floats.
FCD_API int fcd_loss_fwd(const float* pred, const float* target, int B, int D, int H, int W, int kind,
                         float lambda_dice, float lambda_2, float w_bg, float w_fg, float gamma, int squared,
                         int jaccard, int w_type, float smooth_nr, float smooth_dr, float tv_w, int tv_norm, int tv_exclude,
                         unsigned char* keep, float* pbuf, float* part, float* tvpart, float* res, cudaStream_t st) {
    LossCfg cfg{kind, lambda_dice, lambda_2, w_bg, w_fg, gamma, squared, jaccard, w_type, smooth_nr, smooth_dr, tv_w, tv_norm,
                tv_exclude};
    const long long S = (long long)D * H * W;
    const bool tv = tv_w > 0.f;
    if (tv && pbuf == nullptr) return -1;
    if (tv && tv_exclude) {
        if (keep == nullptr) return -1;
        tv_mask_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(target, keep, B, D, H, W);
    }
    loss_fwd_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(pred, target, S, B, cfg, (tv && tv_exclude) ? keep : nullptr,
                                                          tv ? pbuf : nullptr, part);
    if (tv) tv_fwd_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(pbuf, B, D, H, W, tv_norm, tvpart);
    loss_finalize_kernel<<<1, 32, 0, st>>>(part, tv ? tvpart : nullptr, LOSS_BLOCKS, cfg, B, D, H, W, res,
                                           fcd_status_dev());
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_loss_bwd(const float* pred, const float* target, int B, int D, int H, int W, int kind,
                         float lambda_dice, float lambda_2, float w_bg, float w_fg, float gamma, int squared,
                         int jaccard, int w_type, float smooth_nr, float smooth_dr, float tv_w, int tv_norm, int tv_exclude,
                         const unsigned char* keep, const float* pbuf, const float* res, const float* gout,
                         float* dpred, cudaStream_t st) {
    LossCfg cfg{kind, lambda_dice, lambda_2, w_bg, w_fg, gamma, squared, jaccard, w_type, smooth_nr, smooth_dr, tv_w, tv_norm,
                tv_exclude};
    const bool tv = tv_w > 0.f;
    loss_bwd_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, st>>>(pred, target, B, D, H, W, cfg,
                                                          (tv && tv_exclude) ? keep : nullptr, tv ? pbuf : nullptr, res,
                                                          gout, dpred);
    FCD_LAUNCH_CHECK();
}
