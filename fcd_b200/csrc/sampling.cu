// On-device patch sampling and augmentation: the part of the reference's MONAI training pipeline that
// runs per patch -- RandCropByPosNegLabeld(pos=1, neg=1, num_samples), RandFlipd x 3 (p = 0.5 each), RandShiftIntensityd
// (offsets 0.1, p = 0.5) and RandGaussianNoised (std 0.1, p = 0.5) (get_transforms.py:63-84) -- on a whole pre-processed
// volume that is already resident in HBM, with no host round trip and no host random numbers: every decision is a
// counter-based hash of (seed, sample, purpose), also written to a small `meta` record so that a test (or a debugger)
// can reproduce the crop exactly.  RandRotated (get_transforms.py:75) needs MONAI's affine-grid conventions, which the
// oracle cannot pin without a MONAI install: not built.
//
//   fcd_fg_block_counts : foreground (label > 0) voxels per block of 4096 voxels
//   fcd_pick_centers    : per sample: foreground or background (p = pos / (pos + neg); the other class when one is empty),
//                         the r-th voxel of that class in flat order (r uniform), centre clamped so the patch fits
//                         (MONAI correct_crop_centers: start = centre - roi/2 in [0, dim - roi]); flips, shift, noise std
//   fcd_crop_augment    : out[s][c][z][y][x] = img[c][z0 + flip(z)]... + shift_s + std_s * N(0,1), label likewise (no
//                         intensity change); fp32 NCDHW outputs = what the reference's DataLoader hands to train.py:371
#include "common.cuh"

namespace {

constexpr int kBlockVox = 4096;
constexpr int kMeta = 12;      // per sample: z0, y0, x0, flip bits, shift, noise std, picked class (1 fg / 0 bg), rank, cz, cy, cx, 0

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {     // "lowbias32" integer hash
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// uniform in [0, 1) with 24 bits, a pure function of (seed, a, b, c)
__host__ __device__ __forceinline__ float u01(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = mix32((uint32_t)seed ^ 0x9e3779b9U);
    h = mix32(h ^ (uint32_t)(seed >> 32));
    h = mix32(h ^ (a * 0x85ebca6bU + 0x1234567U));
    h = mix32(h ^ (b * 0xc2b2ae35U + 0x89abcdefU));
    h = mix32(h ^ (c * 0x27d4eb2fU + 0x0f1e2d3cU));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256) fg_block_counts_kernel(const float* __restrict__ label, long long V,
                                                              int* __restrict__ counts) {
    const long long base = (long long)blockIdx.x * kBlockVox;
    int n = 0;
    for (int i = threadIdx.x; i < kBlockVox; i += 256) {
        const long long v = base + i;
        if (v < V && label[v] > 0.f) ++n;
    }
    __shared__ int sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        counts[blockIdx.x] = t;
    }
}

// one block; thread s < S decides sample s
__global__ void __launch_bounds__(256) pick_centers_kernel(const float* __restrict__ label, const int* __restrict__ counts,
                                                           int nb, int D, int H, int W, int rd, int rh, int rw, int S,
                                                           unsigned long long seed, float pos_ratio, float flip_p,
                                                           float shift_max, float shift_p, float noise_std,
                                                           float noise_p, float* __restrict__ meta) {
    const long long V = (long long)D * H * W;
    __shared__ long long s_fg;
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int b = 0; b < nb; ++b) t += counts[b];
        s_fg = t;
    }
    __syncthreads();
    const long long F = s_fg, G = V - F;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        bool fg = u01(seed, s, 0, 0) < pos_ratio;
        if (F == 0) fg = false;
        if (G == 0) fg = true;
        const long long n = fg ? F : G;
        // rank from two 24-bit draws (volumes have more than 2^24 voxels)
        const double u = ((double)u01(seed, s, 1, 0) + (double)u01(seed, s, 2, 0) * (1.0 / 16777216.0));
        long long r = (long long)(u * (double)n);
        if (r >= n) r = n - 1;
        // block holding the r-th voxel of the class, then the voxel inside it
        long long acc = 0;
        int b = 0;
        for (; b < nb; ++b) {
            const long long inb = min((long long)kBlockVox, V - (long long)b * kBlockVox);
            const long long c = fg ? counts[b] : inb - counts[b];
            if (acc + c > r) break;
            acc += c;
        }
        long long v = (long long)b * kBlockVox, left = r - acc;
        for (;; ++v) {
            const bool is = label[v] > 0.f;
            if (is == fg) {
                if (left == 0) break;
                --left;
            }
        }
        int cx = (int)(v % W), cy = (int)((v / W) % H), cz = (int)(v / ((long long)W * H));
        auto start = [](int c, int roi, int dim) {
            int s0 = c - roi / 2;
            return s0 < 0 ? 0 : (s0 > dim - roi ? dim - roi : s0);
        };
        float* m = meta + (long long)s * kMeta;
        m[0] = (float)start(cz, rd, D); m[1] = (float)start(cy, rh, H); m[2] = (float)start(cx, rw, W);
        int flips = 0;
        for (int a = 0; a < 3; ++a)
            if (u01(seed, s, 3, a) < flip_p) flips |= 1 << a;
        m[3] = (float)flips;
        m[4] = (u01(seed, s, 4, 0) < shift_p) ? (2.f * u01(seed, s, 4, 1) - 1.f) * shift_max : 0.f;
        m[5] = (u01(seed, s, 5, 0) < noise_p) ? u01(seed, s, 5, 1) * noise_std : 0.f;
        m[6] = fg ? 1.f : 0.f;
        m[7] = (float)r;
        m[8] = (float)cz; m[9] = (float)cy; m[10] = (float)cx; m[11] = 0.f;
    }
}

// grid: (x-chunks, rd * rh rows, S); each thread 4 consecutive x of one output row, all channels
__global__ void __launch_bounds__(128) crop_augment_kernel(const float* __restrict__ img, const float* __restrict__ label,
                                                           int C, int D, int H, int W, int rd, int rh, int rw,
                                                           const float* __restrict__ meta, unsigned long long seed,
                                                           float* __restrict__ out_img, float* __restrict__ out_lab) {
    const int s = blockIdx.z;
    const float* m = meta + (long long)s * kMeta;
    const int z0 = (int)m[0], y0 = (int)m[1], x0 = (int)m[2], flips = (int)m[3];
    const float shift = m[4], nstd = m[5];
    const int row = blockIdx.y, z = row / rh, y = row % rh;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= rw) return;
    // RandFlipd(spatial_axis=a) reverses axis a of the CROPPED patch
    const int sz = z0 + ((flips & 1) ? rd - 1 - z : z);
    const int sy = y0 + ((flips & 2) ? rh - 1 - y : y);
    const long long V = (long long)D * H * W, P = (long long)rd * rh * rw;
    const long long src_row = ((long long)sz * H + sy) * W;
    const long long dst = ((long long)z * rh + y) * rw + x4;
    for (int c = 0; c <= C; ++c) {                       // c == C: the label
        const float* src = (c < C ? img + (long long)c * V : label) + src_row;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x4 + i;
            const int sx = x0 + ((flips & 4) ? rw - 1 - x : x);
            v[i] = x < rw ? src[sx] : 0.f;
        }
        if (c < C) {
            if (nstd > 0.f) {
                // one Box-Muller pair per two voxels; counter = (sample, channel, voxel pair)
                const uint32_t vp = (uint32_t)((dst >> 1) & 0xffffffffu);
#pragma unroll
                for (int i = 0; i < 4; i += 2) {
                    const float u1 = u01(seed, 0x80000000u | s, 16 + c, vp + (i >> 1));
                    const float u2 = u01(seed, 0x80000000u | s, 1024 + c, vp + (i >> 1));
                    const float rr = sqrtf(-2.f * __logf(1.f - u1)), th = 6.28318530718f * u2;
                    v[i] += nstd * rr * __cosf(th);
                    v[i + 1] += nstd * rr * __sinf(th);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] += shift;
        }
        float* o = (c < C ? out_img + ((long long)s * C + c) * P : out_lab + (long long)s * P) + dst;
        if (x4 + 3 < rw && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int i = 0; i < 4 && x4 + i < rw; ++i) o[i] = v[i];
        }
    }
}

}  // namespace

FCD_API int fcd_sampling_block_voxels(void) { return kBlockVox; }
FCD_API int fcd_sampling_meta_floats(void) { return kMeta; }

// counts[(V + 4095) / 4096] = foreground (label > 0) voxels per block of the flat label volume
FCD_API int fcd_fg_block_counts(const float* label, long long V, int* counts, cudaStream_t st) {
    if (V < 1) return -1;
    const int nb = (int)((V + kBlockVox - 1) / kBlockVox);
    fg_block_counts_kernel<<<nb, 256, 0, st>>>(label, V, counts);
    return (int)cudaGetLastError();
}

// meta[S][12] from (label, counts): crop starts, flip bits (bit a = spatial axis a), intensity shift, noise std, class
FCD_API int fcd_pick_centers(const float* label, const int* counts, int D, int H, int W, int rd, int rh, int rw, int S,
                             unsigned long long seed, float pos_ratio, float flip_p, float shift_max, float shift_p,
                             float noise_std, float noise_p, float* meta, cudaStream_t st) {
    if (rd > D || rh > H || rw > W || S < 1 || rd < 1 || rh < 1 || rw < 1) return -1;
    const long long V = (long long)D * H * W;
    const int nb = (int)((V + kBlockVox - 1) / kBlockVox);
    pick_centers_kernel<<<1, 256, 0, st>>>(label, counts, nb, D, H, W, rd, rh, rw, S, seed, pos_ratio, flip_p, shift_max,
                                           shift_p, noise_std, noise_p, meta);
    return (int)cudaGetLastError();
}

// img [C][D][H][W] fp32, label [D][H][W] fp32 -> out_img [S][C][rd][rh][rw], out_lab [S][1][rd][rh][rw] (fp32, NCDHW)
FCD_API int fcd_crop_augment(const float* img, const float* label, int C, int D, int H, int W, int rd, int rh, int rw,
                             int S, const float* meta, unsigned long long seed, float* out_img, float* out_lab,
                             cudaStream_t st) {
    if (rd > D || rh > H || rw > W || S < 1 || C < 1 || (long long)rd * rh > 65535) return -1;
    dim3 grid((rw + 511) / 512, rd * rh, S);
    crop_augment_kernel<<<grid, 128, 0, st>>>(img, label, C, D, H, W, rd, rh, rw, meta, seed, out_img, out_lab);
    return (int)cudaGetLastError();
}
