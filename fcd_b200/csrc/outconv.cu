// Output head: 1x1x1 Conv3d with bias from C (<=64, channels-last bf16) to Co (<=4) classes, written as fp32
// NCDHW logits -- UnetOutBlock (ms_dsa_net.py:362), BaseUNet.final_conv (ms_dsa_net.py:82) and SegResNet's
// conv_final 1x1 (segresnet_dsa.py:192).  HBM-bound: 2*C bytes read + 4*Co bytes written per voxel.
#include "common.cuh"

namespace {

constexpr int OC_MAX_CO = 4;
constexpr int OC_MAX_C = 64;

__global__ void __launch_bounds__(256) outconv_fwd_kernel(const bf16* __restrict__ x, long long ld,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out, long long S, long long total, int C,
                                                          int Co) {
    __shared__ float sw[OC_MAX_CO * OC_MAX_C];
    __shared__ float sb[OC_MAX_CO];
    for (int i = threadIdx.x; i < Co * C; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < Co) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        float acc[OC_MAX_CO];
#pragma unroll
        for (int o = 0; o < OC_MAX_CO; ++o) acc[o] = o < Co ? sb[o] : 0.f;
        for (int c0 = 0; c0 < C; c0 += 8) {
            float f[8];
            unpack8(ld8(x + v * ld + c0), f);
#pragma unroll
            for (int o = 0; o < OC_MAX_CO; ++o) {
                if (o < Co) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[o] = fmaf(f[k], sw[o * C + c0 + k], acc[o]);
                }
            }
        }
        const long long b = v / S, s = v - b * S;
#pragma unroll
        for (int o = 0; o < OC_MAX_CO; ++o)
            if (o < Co) out[(b * Co + o) * S + s] = acc[o];
    }
}

// dx[v][c] = sum_o dout[o][v] w[o][c];  partial dW / dbias per block: part[blk][Co][C+1]
__global__ void __launch_bounds__(256) outconv_bwd_kernel(const bf16* __restrict__ x, long long ld,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ dout, bf16* __restrict__ dx,
                                                          long long lddx, float* __restrict__ part, long long S,
                                                          long long total, int C, int Co) {
    __shared__ float sw[OC_MAX_CO * OC_MAX_C];
    __shared__ float red[8][OC_MAX_CO * (OC_MAX_C + 1)];
    for (int i = threadIdx.x; i < Co * C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each warp accumulates dW for its voxels: lane handles channels lane and lane+32
    float dw[OC_MAX_CO][2], db[OC_MAX_CO];
#pragma unroll
    for (int o = 0; o < OC_MAX_CO; ++o) { dw[o][0] = dw[o][1] = 0.f; db[o] = 0.f; }
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = blockIdx.x * (long long)(blockDim.x >> 5) + warp;
    // a warp processes 32 consecutive voxels per iteration: phase 1 (thread per voxel) dx, phase 2 (thread per
    // channel) dW via shuffles of dout
    for (long long v0 = gw * 32; v0 < total; v0 += nwarps * 32) {
        const long long v = v0 + lane;
        float g[OC_MAX_CO];
#pragma unroll
        for (int o = 0; o < OC_MAX_CO; ++o) g[o] = 0.f;
        if (v < total) {
            const long long b = v / S, s = v - b * S;
#pragma unroll
            for (int o = 0; o < OC_MAX_CO; ++o)
                if (o < Co) g[o] = dout[(b * Co + o) * S + s];
            for (int c0 = 0; c0 < C; c0 += 8) {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float a = 0.f;
#pragma unroll
                    for (int o = 0; o < OC_MAX_CO; ++o)
                        if (o < Co) a = fmaf(g[o], sw[o * C + c0 + k], a);
                    f[k] = a;
                }
                st8(dx + v * lddx + c0, pack8(f));
            }
        }
#pragma unroll
        for (int o = 0; o < OC_MAX_CO; ++o) db[o] += g[o];
        const long long nv = min((long long)32, total - v0);
        for (int j = 0; j < (int)nv; ++j) {
            float gj[OC_MAX_CO];
#pragma unroll
            for (int o = 0; o < OC_MAX_CO; ++o) gj[o] = __shfl_sync(0xffffffffu, g[o], j);
            const bf16* row = x + (v0 + j) * ld;
            const float x0 = lane < C ? __bfloat162float(row[lane]) : 0.f;
            const float x1 = lane + 32 < C ? __bfloat162float(row[lane + 32]) : 0.f;
#pragma unroll
            for (int o = 0; o < OC_MAX_CO; ++o) {
                dw[o][0] = fmaf(gj[o], x0, dw[o][0]);
                dw[o][1] = fmaf(gj[o], x1, dw[o][1]);
            }
        }
    }
    const int W1 = C + 1;
#pragma unroll
    for (int o = 0; o < OC_MAX_CO; ++o) {
        if (o < Co) {
            if (lane < C) red[warp][o * W1 + lane] = dw[o][0];
            if (lane + 32 < C) red[warp][o * W1 + lane + 32] = dw[o][1];
            float s = warp_sum(db[o]);
            if (lane == 0) red[warp][o * W1 + C] = s;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Co * W1; i += blockDim.x) {
        float s = 0.f;
        for (int wq = 0; wq < 8; ++wq) s += red[wq][i];
        part[(long long)blockIdx.x * Co * W1 + i] = s;
    }
}

// Fast path for the shape every model of the hot path has (C = 16 feature channels, Co = 2 classes): thread per voxel,
// the 2x16 dW and 2 dbias partial sums live in registers for the thread's whole voxel range (no per-voxel shuffles),
// one warp-shuffle + shared-memory reduction per block at the end.  Same part[] layout as the generic kernel.
__global__ void __launch_bounds__(256) outconv_bwd_c16o2_kernel(const bf16* __restrict__ x, long long ld,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ dout,
                                                                bf16* __restrict__ dx, long long lddx,
                                                                float* __restrict__ part, long long S,
                                                                long long total) {
    constexpr int C = 16, Co = 2, W1 = C + 1;
    __shared__ float red[8][Co * W1];
    float w0[C], w1[C], a0[C], a1[C], b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { w0[c] = w[c]; w1[c] = w[C + c]; a0[c] = a1[c] = 0.f; }
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        const long long b = v / S, s = v - b * S;
        const float g0 = dout[(b * Co) * S + s], g1 = dout[(b * Co + 1) * S + s];
        float f[C], o[C];
        unpack8(ld8_stream(x + v * ld), f);
        unpack8(ld8_stream(x + v * ld + 8), f + 8);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            o[c] = fmaf(g0, w0[c], g1 * w1[c]);
            a0[c] = fmaf(g0, f[c], a0[c]);
            a1[c] = fmaf(g1, f[c], a1[c]);
        }
        b0 += g0; b1 += g1;
        st8(dx + v * lddx, pack8(o));
        st8(dx + v * lddx + 8, pack8(o + 8));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float s0 = warp_sum(a0[c]), s1 = warp_sum(a1[c]);
        if (lane == 0) { red[warp][c] = s0; red[warp][W1 + c] = s1; }
    }
    b0 = warp_sum(b0); b1 = warp_sum(b1);
    if (lane == 0) { red[warp][C] = b0; red[warp][W1 + C] = b1; }
    __syncthreads();
    for (int i = threadIdx.x; i < Co * W1; i += blockDim.x) {
        float t = 0.f;
        for (int wq = 0; wq < 8; ++wq) t += red[wq][i];
        part[(long long)blockIdx.x * Co * W1 + i] = t;
    }
}

__global__ void outconv_bwd_reduce_kernel(const float* __restrict__ part, int nblk, int C, int Co,
                                          float* __restrict__ dw, float* __restrict__ db) {
    const int W1 = C + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Co * W1; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nblk; ++k) s += part[(long long)k * Co * W1 + i];
        const int o = i / W1, c = i % W1;
        if (c < C) dw[o * C + c] = (float)s;
        else if (db) db[o] = (float)s;
    }
}

}  // namespace

FCD_API int fcd_outconv_blocks() { return 2 * fcd_num_sms(); }

// w: fp32 [Co][C] (C = true channel count, C % 8 == 0 after padding by the caller), bias fp32 [Co] or null.
FCD_API int fcd_outconv_fwd(const void* x, long long ld, const float* w, const float* bias, float* out, int B,
                            long long S, int C, int Co, cudaStream_t st) {
    if (C % 8 || C > OC_MAX_C || Co > OC_MAX_CO || Co < 1) return -1;
    const long long total = (long long)B * S;
    long long g = (total + 255) / 256;
    if (g > 8LL * fcd_num_sms()) g = 8LL * fcd_num_sms();
    outconv_fwd_kernel<<<(int)g, 256, 0, st>>>((const bf16*)x, ld, w, bias, out, S, total, C, Co);
    FCD_LAUNCH_CHECK();
}

// part: fcd_outconv_blocks()*Co*(C+1) floats.  dw fp32 [Co][C], db fp32 [Co] (or null), dx bf16 rows of stride lddx.
FCD_API int fcd_outconv_bwd(const void* x, long long ld, const float* w, const float* dout, void* dx, long long lddx,
                            float* part, float* dw, float* db, int B, long long S, int C, int Co, cudaStream_t st) {
    if (C % 8 || C > OC_MAX_C || Co > OC_MAX_CO || Co < 1) return -1;
    const long long total = (long long)B * S;
    const int nblk = fcd_outconv_blocks();
    if (C == 16 && Co == 2)
        outconv_bwd_c16o2_kernel<<<nblk, 256, 0, st>>>((const bf16*)x, ld, w, dout, (bf16*)dx, lddx, part, S, total);
    else
        outconv_bwd_kernel<<<nblk, 256, 0, st>>>((const bf16*)x, ld, w, dout, (bf16*)dx, lddx, part, S, total, C, Co);
    outconv_bwd_reduce_kernel<<<1, 256, 0, st>>>(part, nblk, C, Co, dw, db);
    FCD_LAUNCH_CHECK();
}
