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
