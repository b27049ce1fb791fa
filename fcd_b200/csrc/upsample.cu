// Pixel-shuffle x2 + ConstantPad3d((1,0,1,0,1,0)) + AvgPool3d(2, stride 1): the tail of MONAI SubpixelUpsample
// (UpSample(mode="pixelshuffle"), reference conv_blocks.py:727-735 and segresnet_dsa.py:133-141),
// fused into one bandwidth-bound pass, optionally with the residual add `up(x) + skip` of SegResNet.decode
// (segresnet_dsa.py:217) or writing straight into the left half of a concat buffer (conv_blocks.py:771).
//
// The preceding 3x3x3 conv is run with its output channels re-ordered to n' = tap*Cq + c (tap = i*4+j*2+k, the
// sub-lattice offset MONAI's pixelshuffle assigns to channel c*8+tap), so every 8-channel chunk is one 16-byte load.
#include "common.cuh"

namespace {

// out[b,Z,Y,X,c] = 1/8 * sum_{a,b',c' in {0,1}} S[Z-a, Y-b', X-c'],  S[Z',..] = src[Z'/2,..][tap(Z'%2,..)*Cq + c]
__global__ void ps_blur_fwd_kernel(const bf16* __restrict__ src, long long lds, const bf16* __restrict__ skip,
                                   long long ldk, bf16* __restrict__ out, long long ldo, int B, int D, int H, int W,
                                   int Cq) {
    const int C8 = Cq / 8;
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * Do * Ho * Wo * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long r = i / C8;
        const int X = (int)(r % Wo); r /= Wo;
        const int Y = (int)(r % Ho); r /= Ho;
        const int Z = (int)(r % Do);
        const long long b = r / Do;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int zz = Z - (t >> 2), yy = Y - ((t >> 1) & 1), xx = X - (t & 1);
            if (zz < 0 || yy < 0 || xx < 0) continue;
            const int tap = ((zz & 1) << 2) | ((yy & 1) << 1) | (xx & 1);
            const long long vox = ((b * D + (zz >> 1)) * H + (yy >> 1)) * W + (xx >> 1);
            float f[8];
            unpack8(ld8(src + vox * lds + tap * Cq + c8 * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
        const long long ov = ((b * Do + Z) * Ho + Y) * Wo + X;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= 0.125f;
        if (skip) {
            float f[8];
            unpack8(ld8(skip + ov * ldk + c8 * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
        st8(out + ov * ldo + c8 * 8, pack8(acc));
    }
}

// dsrc[b,z,y,x, tap*Cq + c] = 1/8 * sum_{a,b',c'} dout[2z+i+a, 2y+j+b', 2x+k+c']   (in range)
__global__ void ps_blur_bwd_kernel(const bf16* __restrict__ dout, long long lddo, bf16* __restrict__ dsrc,
                                   long long ldds, int B, int D, int H, int W, int Cq) {
    const int C8 = Cq / 8;
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * D * H * W * 8 * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long r = i / C8;
        const int tap = (int)(r % 8); r /= 8;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H); r /= H;
        const int z = (int)(r % D);
        const long long b = r / D;
        const int Z0 = 2 * z + (tap >> 2), Y0 = 2 * y + ((tap >> 1) & 1), X0 = 2 * x + (tap & 1);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int zz = Z0 + (t >> 2), yy = Y0 + ((t >> 1) & 1), xx = X0 + (t & 1);
            if (zz >= Do || yy >= Ho || xx >= Wo) continue;
            float f[8];
            unpack8(ld8(dout + (((b * Do + zz) * Ho + yy) * Wo + xx) * lddo + c8 * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= 0.125f;
        const long long vox = ((b * D + z) * H + y) * W + x;
        st8(dsrc + vox * ldds + tap * Cq + c8 * 8, pack8(acc));
    }
}

// ---- UpSample(mode="nontrainable"): nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)
// (MONAI UpSample, conv_blocks.py:727-735 with upsample_mode="nontrainable"; segresnet_dsa.py:133-141).
// PyTorch's source index for output index o: s = max((o + 0.5) / 2 - 0.5, 0); i0 = floor(s), i1 = min(i0 + 1, n - 1),
// weight of i1 = s - i0.
__device__ __forceinline__ void tri_src(int o, int n, int& i0, int& i1, float& w1) {
    const float s = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.f);
    i0 = (int)s;
    i1 = min(i0 + 1, n - 1);
    w1 = s - (float)i0;
}

__global__ void trilinear_up_fwd_kernel(const bf16* __restrict__ src, long long lds, const bf16* __restrict__ skip,
                                        long long ldk, bf16* __restrict__ out, long long ldo, int B, int D, int H, int W,
                                        int C) {
    const int C8 = C / 8;
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * Do * Ho * Wo * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long r = i / C8;
        const int X = (int)(r % Wo); r /= Wo;
        const int Y = (int)(r % Ho); r /= Ho;
        const int Z = (int)(r % Do);
        const long long b = r / Do;
        int z0, z1, y0, y1, x0, x1;
        float wz, wy, wx;
        tri_src(Z, D, z0, z1, wz);
        tri_src(Y, H, y0, y1, wy);
        tri_src(X, W, x0, x1, wx);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int zz = (t & 4) ? z1 : z0, yy = (t & 2) ? y1 : y0, xx = (t & 1) ? x1 : x0;
            const float w = ((t & 4) ? wz : 1.f - wz) * ((t & 2) ? wy : 1.f - wy) * ((t & 1) ? wx : 1.f - wx);
            if (w == 0.f) continue;
            float f[8];
            unpack8(ld8(src + (((b * D + zz) * H + yy) * W + xx) * lds + c8 * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, f[k], acc[k]);
        }
        const long long ov = ((b * Do + Z) * Ho + Y) * Wo + X;
        if (skip) {
            float f[8];
            unpack8(ld8(skip + ov * ldk + c8 * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
        st8(out + ov * ldo + c8 * 8, pack8(acc));
    }
}

// dsrc[z,y,x] = sum over the output voxels whose interpolation footprint contains (z,y,x), with their weights
__global__ void trilinear_up_bwd_kernel(const bf16* __restrict__ dout, long long lddo, bf16* __restrict__ dsrc,
                                        long long ldds, int B, int D, int H, int W, int C) {
    const int C8 = C / 8;
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * D * H * W * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long r = i / C8;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H); r /= H;
        const int z = (int)(r % D);
        const long long b = r / D;
        // per axis: the (at most 4) output indices 2i-1 .. 2i+2 and the weight each gives to source index i
        float wz[4], wy[4], wx[4];
        auto axis = [](int i, int n, int no, float* w) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o = 2 * i - 1 + k;
                w[k] = 0.f;
                if (o < 0 || o >= no) continue;
                int i0, i1;
                float w1;
                tri_src(o, n, i0, i1, w1);
                if (i0 == i) w[k] += 1.f - w1;
                if (i1 == i) w[k] += w1;
            }
        };
        axis(z, D, Do, wz);
        axis(y, H, Ho, wy);
        axis(x, W, Wo, wx);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int a = 0; a < 4; ++a) {
            if (wz[a] == 0.f) continue;
            for (int bb = 0; bb < 4; ++bb) {
                if (wy[bb] == 0.f) continue;
                for (int c = 0; c < 4; ++c) {
                    const float w = wz[a] * wy[bb] * wx[c];
                    if (w == 0.f) continue;
                    const long long ov = ((b * Do + (2 * z - 1 + a)) * Ho + (2 * y - 1 + bb)) * Wo + (2 * x - 1 + c);
                    float f[8];
                    unpack8(ld8(dout + ov * lddo + c8 * 8), f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, f[k], acc[k]);
                }
            }
        }
        st8(dsrc + (((b * D + z) * H + y) * W + x) * ldds + c8 * 8, pack8(acc));
    }
}

inline int ps_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 8LL * fcd_num_sms();
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

FCD_API int fcd_ps_blur_fwd(const void* src, long long lds, const void* skip, long long ldk, void* out, long long ldo,
                            int B, int D, int H, int W, int Cq, cudaStream_t st) {
    if (Cq % 8) return -1;
    const long long total = (long long)B * 8 * D * H * W * (Cq / 8);
    ps_blur_fwd_kernel<<<ps_grid(total), 256, 0, st>>>((const bf16*)src, lds, (const bf16*)skip, ldk, (bf16*)out, ldo,
                                                      B, D, H, W, Cq);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_ps_blur_bwd(const void* dout, long long lddo, void* dsrc, long long ldds, int B, int D, int H, int W,
                            int Cq, cudaStream_t st) {
    if (Cq % 8) return -1;
    const long long total = (long long)B * 8 * D * H * W * (Cq / 8);
    ps_blur_bwd_kernel<<<ps_grid(total), 256, 0, st>>>((const bf16*)dout, lddo, (bf16*)dsrc, ldds, B, D, H, W, Cq);
    FCD_LAUNCH_CHECK();
}

// nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False) on channels-last rows (C % 8 == 0), optionally
// + skip (SegResNet.decode, segresnet_dsa.py:217) or into the left half of a concat buffer (ldo > C, conv_blocks.py:771).
FCD_API int fcd_trilinear_up_fwd(const void* src, long long lds, const void* skip, long long ldk, void* out,
                                 long long ldo, int B, int D, int H, int W, int C, cudaStream_t st) {
    if (C % 8) return -1;
    const long long total = (long long)B * 8 * D * H * W * (C / 8);
    trilinear_up_fwd_kernel<<<ps_grid(total), 256, 0, st>>>((const bf16*)src, lds, (const bf16*)skip, ldk, (bf16*)out,
                                                           ldo, B, D, H, W, C);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_trilinear_up_bwd(const void* dout, long long lddo, void* dsrc, long long ldds, int B, int D, int H,
                                 int W, int C, cudaStream_t st) {
    if (C % 8) return -1;
    const long long total = (long long)B * D * H * W * (C / 8);
    trilinear_up_bwd_kernel<<<ps_grid(total), 256, 0, st>>>((const bf16*)dout, lddo, (bf16*)dsrc, ldds, B, D, H, W, C);
    FCD_LAUNCH_CHECK();
}
