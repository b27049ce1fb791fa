// Sliding-window whole-volume inference support (monai.inferers.sliding_window_inference as called at reference
// train.py:148-165 / seg_fcd_test.py:37-54, mode='constant'):
//   gather  : cut sw_batch_size windows out of the fp32 NCDHW volume straight into the model's channels-last bf16
//             input batch (replaces the torch slicing + torch.cat + dtype casts);
//   blend   : out[win] += pred, fp32, one launch per window so the summation order is the reference's window order;
//   finalize: out /= count (count = product of per-axis coverage counts, importance map == 1), and optionally the
//             label map of train.py:185,209-211 (softmax >= 0.5 per channel) / get_transforms.py:142-154 (argmax).
#include "common.cuh"

namespace {

struct WinStarts {
    int z[8], y[8], x[8];
};

__global__ void sw_gather_kernel(const float* __restrict__ vol, bf16* __restrict__ dst, int C, int Cp, int D, int H,
                                 int W, int r0, int r1, int r2, int pz, int py, int px, WinStarts ws, int nwin) {
    // vol: [C][D][H][W] of ONE image; window voxel (z,y,x) reads vol[.., ws.z+z-pz, ..] with zero padding (images
    // smaller than the roi are padded symmetrically by MONAI; pz/py/px are the low-side pads)
    const long long per = (long long)r0 * r1 * r2;
    const long long total = per * nwin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / per);
        long long r = i - (long long)j * per;
        const int x = (int)(r % r2); r /= r2;
        const int y = (int)(r % r1);
        const int z = (int)(r / r1);
        const int sz = ws.z[j] + z - pz, sy = ws.y[j] + y - py, sx = ws.x[j] + x - px;
        const bool in = sz >= 0 && sz < D && sy >= 0 && sy < H && sx >= 0 && sx < W;
        const long long so = ((long long)sz * H + sy) * W + sx;
        for (int c0 = 0; c0 < Cp; c0 += 8) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = (in && c0 + k < C) ? vol[(long long)(c0 + k) * D * H * W + so] : 0.f;
            st8(dst + i * Cp + c0, pack8(f));
        }
    }
}

// out[c*sc + (z0+z)*sz + (y0+y)*Wp + (x0+x)] += pred[c][z][y][x]   (pred: one window; out: the padded accumulation
// volume, channel-major [C][Dp][Hp][Wp] (sc = Dp*Hp*Wp, sz = Hp*Wp) or plane-major [Dp][C][Hp][Wp] (sc = Hp*Wp,
// sz = C*Hp*Wp: the layout whose D-slabs are contiguous for the multi-GPU reduce-scatter).  VEC: 4 x-voxels per thread.
template <int VEC>
__global__ void sw_blend_kernel(const float* __restrict__ pred, float* __restrict__ out, int C, int r0, int r1, int r2,
                                int Wp, long long sc, long long sz, int z0, int y0, int x0) {
    const int r2v = r2 / VEC;
    const long long per = (long long)r0 * r1 * r2v;
    const long long total = per * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / per);
        long long r = i - (long long)c * per;
        const int x = (int)(r % r2v) * VEC; r /= r2v;
        const int y = (int)(r % r1);
        const int z = (int)(r / r1);
        const long long o = (long long)c * sc + (long long)(z0 + z) * sz + (long long)(y0 + y) * Wp + (x0 + x);
        if (VEC == 4) {
            const float4 pv = *reinterpret_cast<const float4*>(pred + i * 4);
            float4 ov = *reinterpret_cast<float4*>(out + o);
            ov.x += pv.x; ov.y += pv.y; ov.z += pv.z; ov.w += pv.w;
            *reinterpret_cast<float4*>(out + o) = ov;
        } else {
            out[o] += pred[i];
        }
    }
}

// For unpadded voxel (z,y,x), z in [z_lo, z_hi):  v[c] = acc[c*sc + (z+pz-acc_z0)*sz + (y+py)*Wp + (x+px)] /
// (cz[z+pz] * cy[y+py] * cx[x+px]);  dst (optional) [C][out_planes][H][W] at plane z-out_z0;  label (optional):
//   mode 1: label_f[c][..] = softmax(v)[c] >= 0.5 (float {0,1}, C channels);  mode 2: label_u8[..] = argmax_c (uint8)
__global__ void sw_finalize_kernel(const float* __restrict__ acc, const int* __restrict__ cz,
                                   const int* __restrict__ cy, const int* __restrict__ cx, float* __restrict__ dst,
                                   float* __restrict__ label_f, unsigned char* __restrict__ label_u8, int C, int H,
                                   int W, int Wp, long long sc, long long sz, int pz, int py, int px, int z_lo, int z_hi,
                                   int acc_z0, int out_z0, int out_planes, int mode, const int* __restrict__ status) {
    const long long total = (long long)(z_hi - z_lo) * H * W;
    const long long S = (long long)out_planes * H * W;
    // a tcgen05 pipeline wait timed out upstream (status word, csrc/status.cu): poison the result (NaN logits, label 255)
    const bool bad = status != nullptr && *reinterpret_cast<const volatile int*>(status) != 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int z = (int)(r / H) + z_lo;
        const float cnt = (float)(cz[z + pz] * cy[y + py] * cx[x + px]);   // true division below: bit-exact with `out / cnt`
        const long long so = (long long)(z + pz - acc_z0) * sz + (long long)(y + py) * Wp + (x + px);
        const long long o = ((long long)(z - out_z0) * H + y) * W + x;
        float v[8];
        float mx = -INFINITY;
        int am = 0;
        for (int c = 0; c < C; ++c) {
            v[c] = bad ? NAN : acc[(long long)c * sc + so] / cnt;
            if (dst != nullptr) dst[(long long)c * S + o] = v[c];
            if (v[c] > mx) { mx = v[c]; am = c; }
        }
        if (mode == 1) {
            float den = 0.f;
            for (int c = 0; c < C; ++c) den += expf(v[c] - mx);
            for (int c = 0; c < C; ++c)
                label_f[(long long)c * S + o] = bad ? NAN : ((expf(v[c] - mx) / den >= 0.5f) ? 1.f : 0.f);
        } else if (mode == 2) {
            label_u8[o] = bad ? (unsigned char)255 : (unsigned char)am;
        }
    }
}

inline int sw_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 8LL * fcd_num_sms();
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

// starts: host arrays of nwin (<= 8) window origins in the PADDED volume frame.
FCD_API int fcd_sw_gather(const float* vol, void* dst, int C, int Cp, int D, int H, int W, int r0, int r1, int r2,
                          int pz, int py, int px, const int* starts_zyx, int nwin, cudaStream_t st) {
    if (nwin < 1 || nwin > 8 || Cp % 8) return -1;
    WinStarts ws;
    for (int j = 0; j < nwin; ++j) { ws.z[j] = starts_zyx[3 * j]; ws.y[j] = starts_zyx[3 * j + 1]; ws.x[j] = starts_zyx[3 * j + 2]; }
    const long long total = (long long)nwin * r0 * r1 * r2;
    sw_gather_kernel<<<sw_grid(total), 256, 0, st>>>(vol, (bf16*)dst, C, Cp, D, H, W, r0, r1, r2, pz, py, px, ws, nwin);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_sw_blend(const float* pred, float* out, int C, int r0, int r1, int r2, int Wp, long long sc,
                         long long sz, int z0, int y0, int x0, cudaStream_t st) {
    const bool vec = r2 % 4 == 0 && x0 % 4 == 0 && Wp % 4 == 0 && sc % 4 == 0 && sz % 4 == 0 &&
                     ((uintptr_t)pred & 15) == 0 && ((uintptr_t)out & 15) == 0;
    const long long total = (long long)C * r0 * r1 * (vec ? r2 / 4 : r2);
    if (vec) sw_blend_kernel<4><<<sw_grid(total), 256, 0, st>>>(pred, out, C, r0, r1, r2, Wp, sc, sz, z0, y0, x0);
    else sw_blend_kernel<1><<<sw_grid(total), 256, 0, st>>>(pred, out, C, r0, r1, r2, Wp, sc, sz, z0, y0, x0);
    FCD_LAUNCH_CHECK();
}

// cz/cy/cx: device int arrays (padded-frame per-axis window coverage counts).  Processes unpadded planes z in
// [z_lo, z_hi); acc plane 0 is padded plane acc_z0; dst / label buffers hold out_planes planes starting at plane out_z0.
FCD_API int fcd_sw_finalize(const float* acc, const int* cz, const int* cy, const int* cx, float* dst, float* label_f,
                            void* label_u8, int C, int H, int W, int Wp, long long sc, long long sz, int pz, int py,
                            int px, int z_lo, int z_hi, int acc_z0, int out_z0, int out_planes, int mode,
                            cudaStream_t st) {
    if (C > 8 || z_lo < 0 || z_hi <= z_lo || z_lo < out_z0 || z_hi > out_z0 + out_planes || z_lo + pz < acc_z0) return -1;
    const long long total = (long long)(z_hi - z_lo) * H * W;
    sw_finalize_kernel<<<sw_grid(total), 256, 0, st>>>(acc, cz, cy, cx, dst, label_f, (unsigned char*)label_u8, C, H, W,
                                                      Wp, sc, sz, pz, py, px, z_lo, z_hi, acc_z0, out_z0, out_planes,
                                                      mode, fcd_status_dev());
    FCD_LAUNCH_CHECK();
}
