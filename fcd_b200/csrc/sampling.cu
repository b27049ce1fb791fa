// On-device patch sampling and augmentation: the part of the reference's MONAI training pipeline that
// runs per patch -- RandCropByPosNegLabeld(pos=1, neg=1, num_samples), RandFlipd x 3 (p = 0.5 each), RandRotated(range_y =
// pi/2, bilinear / nearest, p = 0.5), RandShiftIntensityd (offsets 0.1, p = 0.5), RandGaussianNoised (std 0.1, p = 0.5),
// RandCoarseDropoutd(holes=5, spatial_size=16^3, fill 0) and GridMaskd (utils/gridmask.py:8-72) (get_transforms.py:45-89)
// -- on a whole pre-processed volume that is already resident in HBM, with no host round trip and no host random numbers:
// every decision is a counter-based hash of (seed, sample, purpose), also written to a small `meta` record so that a test
// (or a debugger) can reproduce the patch exactly.
//
//   fcd_fg_block_counts : foreground (label > 0) voxels per block of 4096 voxels
//   fcd_pick_centers    : per sample: foreground or background (p = pos / (pos + neg); the other class when one is empty),
//                         the r-th voxel of that class in flat order (r uniform), centre clamped so the patch fits
//                         (MONAI correct_crop_centers: start = centre - roi/2 in [0, dim - roi]); flips, rotation angle,
//                         shift, noise std, hole corners, grid period / phases
//   fcd_crop_augment    : one gather pass per patch: crop -> flips -> rotation about spatial axis 1 around the patch centre
//                         (src = c + R (p - c); image bilinear, label nearest, border padding: MONAI Rotate with
//                         keep_size [RECALLED]) -> + std_s * N(0,1) + shift_s -> holes = 0 -> * grid mask; the label gets
//                         the spatial transforms only; fp32 NCDHW outputs = what the reference's DataLoader hands to
//                         train.py:371.  The interpolation is written with explicit round-to-nearest multiplies and adds
//                         (no FMA contraction) so that the numpy oracle reproduces it bit for bit.
#include "common.cuh"

namespace {

constexpr int kBlockVox = 4096;
constexpr int kMaxHoles = 8;
// per sample: 0-2 z0, y0, x0; 3 flip bits; 4 shift; 5 noise std; 6 picked class (1 fg / 0 bg); 7 rank; 8-10 cz, cy, cx;
// 11 rotated (0/1); 12 cos; 13 sin; 14 angle; 15 holes applied; 16.. hole corners (z, y, x) x kMaxHoles; 40 grid mask on;
// 41 period d; 42 stripe width l; 43-45 phases (axis 0, 1, 2); 46 inverted; 47 reserved
constexpr int kMeta = 48;
constexpr int kHole0 = 16, kGrid0 = 40;

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
    if (base + kBlockVox <= V && (reinterpret_cast<uintptr_t>(label) & 15) == 0) {
        // whole block inside the volume: four independent 16-byte loads per thread
        const float4* p = reinterpret_cast<const float4*>(label + base);
        float4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = __ldg(p + k * 256 + threadIdx.x);
#pragma unroll
        for (int k = 0; k < 4; ++k) n += (q[k].x > 0.f) + (q[k].y > 0.f) + (q[k].z > 0.f) + (q[k].w > 0.f);
    } else {
        for (int i = threadIdx.x; i < kBlockVox; i += 256) {
            const long long v = base + i;
            if (v < V && label[v] > 0.f) ++n;
        }
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

// one WARP per sample (8 per block): the lanes scan the per-block counts and then the voxels of the chosen block 32 at a
// time (ballot + popcount), which finds the same r-th voxel of the class in flat order as a serial walk; lane 0 writes
__global__ void __launch_bounds__(256) pick_centers_kernel(const float* __restrict__ label, const int* __restrict__ counts,
                                                           int nb, int D, int H, int W, int rd, int rh, int rw, int S,
                                                           unsigned long long seed, float pos_ratio, float flip_p,
                                                           float shift_max, float shift_p, float noise_std,
                                                           float noise_p, float rot_p, float rot_range, float cd_p,
                                                           int holes, int hz, int hy, int hx, float grid_p, int d1,
                                                           int d2, double grid_ratio, int grid_invert,
                                                           float* __restrict__ meta) {
    constexpr unsigned kFull = 0xffffffffu;
    const long long V = (long long)D * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ long long s_part[8];
    {
        long long t = 0;
        for (int b = threadIdx.x; b < nb; b += 256) t += counts[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
        if (lane == 0) s_part[warp] = t;
    }
    __syncthreads();
    long long F = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) F += s_part[w];
    const long long G = V - F;
    const int s = blockIdx.x * 8 + warp;
    if (s >= S) return;
    {
        bool fg = u01(seed, s, 0, 0) < pos_ratio;
        if (F == 0) fg = false;
        if (G == 0) fg = true;
        const long long n = fg ? F : G;
        // rank from two 24-bit draws (volumes have more than 2^24 voxels)
        const double u = ((double)u01(seed, s, 1, 0) + (double)u01(seed, s, 2, 0) * (1.0 / 16777216.0));
        long long r = (long long)(u * (double)n);
        if (r >= n) r = n - 1;
        // block holding the r-th voxel of the class: first block with (voxels of the class before it) + (its own) > r.
        // Two levels: every lane sums a contiguous chunk of the blocks (independent loads), one warp scan picks the
        // chunk, the warp scans that chunk 32 blocks at a time.
        auto class_count = [&](int bi) -> long long {
            if (bi >= nb) return 0;
            const long long inb = min((long long)kBlockVox, V - (long long)bi * kBlockVox);
            return fg ? (long long)counts[bi] : inb - counts[bi];
        };
        auto warp_scan = [&](long long c) {            // inclusive prefix over the lanes
            long long incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += up;
            }
            return incl;
        };
        const int chunk = (nb + 31) / 32;              // blocks per lane
        long long mine = 0;
        for (int j = 0; j < chunk; ++j) mine += class_count(lane * chunk + j);
        long long incl = warp_scan(mine);
        const int lsel = __ffs(__ballot_sync(kFull, incl > r)) - 1;        // r < n: some lane passes
        long long acc = __shfl_sync(kFull, incl - mine, lsel);             // voxels of the class before the chunk
        int b = -1;
        for (int b0 = lsel * chunk; b < 0; b0 += 32) {
            const int bi = b0 + lane;
            const long long c = bi < (lsel + 1) * chunk ? class_count(bi) : 0;
            incl = warp_scan(c);
            const long long total = __shfl_sync(kFull, incl, 31);
            if (acc + total > r) {
                const int l = __ffs(__ballot_sync(kFull, acc + incl > r)) - 1;
                b = b0 + l;
                acc += __shfl_sync(kFull, incl - c, l);
            } else {
                acc += total;
            }
        }
        // ... then the voxel inside it, the same way: 128 consecutive voxels per lane, then 32 at a time
        const long long base = (long long)b * kBlockVox;
        int left = (int)(r - acc);
        constexpr int kPerLane = kBlockVox / 32;
        int cnt_lane = 0;
        {
            const long long v0 = base + (long long)lane * kPerLane;
            if (v0 + kPerLane <= V && (reinterpret_cast<uintptr_t>(label) & 15) == 0) {
                const float4* p = reinterpret_cast<const float4*>(label + v0);
#pragma unroll 8
                for (int j = 0; j < kPerLane / 4; ++j) {
                    const float4 q = __ldg(p + j);
                    cnt_lane += ((q.x > 0.f) == fg) + ((q.y > 0.f) == fg) + ((q.z > 0.f) == fg) + ((q.w > 0.f) == fg);
                }
            } else {
                for (int j = 0; j < kPerLane; ++j)
                    cnt_lane += (v0 + j < V) && ((label[v0 + j] > 0.f) == fg);
            }
        }
        int incl_v = cnt_lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(kFull, incl_v, o);
            if (lane >= o) incl_v += up;
        }
        const int vsel = __ffs(__ballot_sync(kFull, incl_v > left)) - 1;
        left -= __shfl_sync(kFull, incl_v - cnt_lane, vsel);
        long long v = base;
        for (int i0 = vsel * kPerLane; i0 < (vsel + 1) * kPerLane; i0 += 32) {
            const long long vi = base + i0 + lane;
            const bool match = vi < V && ((label[vi] > 0.f) == fg);
            unsigned mm = __ballot_sync(kFull, match);
            const int cnt = __popc(mm);
            if (left < cnt) {
                for (int k = 0; k < left; ++k) mm &= mm - 1;         // drop the `left` lowest matches
                v = base + i0 + (__ffs(mm) - 1);
                break;
            }
            left -= cnt;
        }
        if (lane != 0) return;
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
        m[8] = (float)cz; m[9] = (float)cy; m[10] = (float)cx;
        for (int i = 11; i < kMeta; ++i) m[i] = 0.f;
        // RandRotated(range_y): angle ~ U(-range, range) about spatial axis 1
        if (u01(seed, s, 6, 0) < rot_p) {
            const float ang = (2.f * u01(seed, s, 6, 1) - 1.f) * rot_range;
            m[11] = 1.f; m[12] = cosf(ang); m[13] = sinf(ang); m[14] = ang;
        }
        // RandCoarseDropout: `holes` boxes of (hz, hy, hx), corner uniform in [0, dim - size]
        if (u01(seed, s, 7, 0) < cd_p) {
            m[15] = (float)holes;
            const int dims[3] = {rd, rh, rw}, hs[3] = {hz, hy, hx};
            for (int h = 0; h < holes; ++h)
                for (int a = 0; a < 3; ++a) {
                    const int span = dims[a] - hs[a] + 1;
                    int c0 = (int)(u01(seed, s, 7, 1 + 3 * h + a) * (float)span);
                    m[kHole0 + 3 * h + a] = (float)(c0 > span - 1 ? span - 1 : c0);
                }
        }
        // GridMask (utils/gridmask.py:20-45): period d in [d1, d2), stripe width ceil(d * ratio), one phase per axis
        if (u01(seed, s, 9, 0) < grid_p) {
            int d = d1 + (int)(u01(seed, s, 9, 1) * (float)(d2 - d1));
            if (d > d2 - 1) d = d2 - 1;
            m[kGrid0] = 1.f; m[kGrid0 + 1] = (float)d; m[kGrid0 + 2] = (float)(int)ceil((double)d * grid_ratio);
            for (int a = 0; a < 3; ++a) {
                int st = (int)(u01(seed, s, 9, 2 + a) * (float)d);
                m[kGrid0 + 3 + a] = (float)(st > d - 1 ? d - 1 : st);
            }
            m[kGrid0 + 6] = grid_invert ? 1.f : 0.f;
        }
    }
}

// value of the cropped + flipped patch at patch coordinates (q0, q1, q2)
__device__ __forceinline__ float patch_at(const float* __restrict__ src, int H, int W, int z0, int y0, int x0, int flips,
                                          int rd, int rh, int rw, int q0, int q1, int q2) {
    const int sz = z0 + ((flips & 1) ? rd - 1 - q0 : q0);
    const int sy = y0 + ((flips & 2) ? rh - 1 - q1 : q1);
    const int sx = x0 + ((flips & 4) ? rw - 1 - q2 : q2);
    return __ldg(src + ((sz * H + sy) * W + sx));          // D * H * W < 2^31 (checked by the entry point)
}

// grid: (x-chunks, row groups, S); block = rpb rows x tpr threads (tpr = threads a row needs, a multiple of 32), each thread
// 4 consecutive x of one output row, all channels
__global__ void __launch_bounds__(128, 6) crop_augment_kernel(const float* __restrict__ img, const float* __restrict__ label,
                                                           int C, int D, int H, int W, int rd, int rh, int rw,
                                                           const float* __restrict__ meta, unsigned long long seed,
                                                           int hz, int hy, int hx, int hh, int tpr,
                                                           float* __restrict__ out_img, float* __restrict__ out_lab) {
    const int s = blockIdx.z;
    const float* m = meta + (long long)s * kMeta;
    const int z0 = (int)m[0], y0 = (int)m[1], x0 = (int)m[2], flips = (int)m[3];
    const float shift = m[4], nstd = m[5];
    const bool rot = m[11] != 0.f;
    const float cs = m[12], sn = m[13];
    const int nholes = (int)m[15];
    const bool grid = m[kGrid0] != 0.f;
    const int row = blockIdx.y * (128 / tpr) + threadIdx.x / tpr, z = row / rh, y = row % rh;
    const int x4 = (blockIdx.x * tpr + threadIdx.x % tpr) * 4;
    if (x4 >= rw || row >= rd * rh) return;
    const long long V = (long long)D * H * W, P = (long long)rd * rh * rw;
    const long long dst = ((long long)z * rh + y) * rw + x4;

    // rotation about spatial axis 1: source coordinates on axes 0 and 2 (border padding).  They are recomputed per channel
    // (a dozen flops) instead of being kept in registers across the channel loop: occupancy matters more here.
    const float c0 = (float)(rd - 1) * 0.5f, c2 = (float)(rw - 1) * 0.5f;
    const float e0 = (float)z - c0;
    // coarse-dropout holes and grid mask of the 4 voxels (image channels only)
    float keep[4] = {1.f, 1.f, 1.f, 1.f};
    bool hole[4] = {false, false, false, false};
    for (int h = 0; h < nholes; ++h) {
        const int bz = (int)m[kHole0 + 3 * h], by = (int)m[kHole0 + 3 * h + 1], bx = (int)m[kHole0 + 3 * h + 2];
        if (z >= bz && z < bz + hz && y >= by && y < by + hy) {
#pragma unroll
            for (int i = 0; i < 4; ++i) hole[i] |= (x4 + i >= bx && x4 + i < bx + hx);
        }
    }
    if (grid) {
        // utils/gridmask.py:34-66: a cube of edge hh, stripes [d i + st, d i + st + l) zeroed on each axis, centre-cropped
        const int d = (int)m[kGrid0 + 1], l = (int)m[kGrid0 + 2];
        const int st0 = (int)m[kGrid0 + 3], st1 = (int)m[kGrid0 + 4], st2 = (int)m[kGrid0 + 5];
        const bool inv = m[kGrid0 + 6] != 0.f;
        auto striped = [d, l](int p, int st) { int r = (p - st) % d; if (r < 0) r += d; return r < l; };
        const bool zy = striped(z + (hh - rd) / 2, st0) || striped(y + (hh - rh) / 2, st1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool masked = zy || striped(x4 + i + (hh - rw) / 2, st2);
            keep[i] = (masked != inv) ? 0.f : 1.f;
        }
    }

    for (int c = 0; c <= C; ++c) {                       // c == C: the label
        const float* src = (c < C ? img + (long long)c * V : label);
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x4 + i;
            if (x >= rw) { v[i] = 0.f; continue; }
            if (!rot) {
                // RandFlipd(spatial_axis=a) reverses axis a of the CROPPED patch
                v[i] = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, z, y, x);
                continue;
            }
            const float e2 = (float)x - c2;
            float s0 = __fadd_rn(c0, __fadd_rn(__fmul_rn(cs, e0), __fmul_rn(sn, e2)));
            float s2 = __fadd_rn(c2, __fsub_rn(__fmul_rn(cs, e2), __fmul_rn(sn, e0)));
            s0 = fminf(fmaxf(s0, 0.f), (float)(rd - 1));
            s2 = fminf(fmaxf(s2, 0.f), (float)(rw - 1));
            if (c == C) {                                 // label: nearest, round half to even (grid_sample)
                v[i] = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, (int)rintf(s0), y, (int)rintf(s2));
                continue;
            }
            const float f0 = floorf(s0), f2 = floorf(s2);
            const float w0 = __fsub_rn(s0, f0), w2 = __fsub_rn(s2, f2);
            const int i0 = (int)f0, i2 = (int)f2;
            const int i0b = min(i0 + 1, rd - 1), i2b = min(i2 + 1, rw - 1);
            const float v00 = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, i0, y, i2);
            const float v01 = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, i0, y, i2b);
            const float v10 = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, i0b, y, i2);
            const float v11 = patch_at(src, H, W, z0, y0, x0, flips, rd, rh, rw, i0b, y, i2b);
            const float u2 = __fsub_rn(1.f, w2), u0 = __fsub_rn(1.f, w0);
            const float a = __fadd_rn(__fmul_rn(v00, u2), __fmul_rn(v01, w2));
            const float b = __fadd_rn(__fmul_rn(v10, u2), __fmul_rn(v11, w2));
            v[i] = __fadd_rn(__fmul_rn(a, u0), __fmul_rn(b, w0));
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
            for (int i = 0; i < 4; ++i) {
                v[i] += shift;
                if (hole[i]) v[i] = 0.f;                 // RandCoarseDropout(fill_value=0)
                if (grid) v[i] = __fmul_rn(v[i], keep[i]);   // img * mask
            }
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

// meta[S][fcd_sampling_meta_floats()] from (label, counts): crop starts, flip bits (bit a = spatial axis a), rotation,
// intensity shift, noise std, class, hole corners, grid mask (layout at the top of this file)
FCD_API int fcd_pick_centers(const float* label, const int* counts, int D, int H, int W, int rd, int rh, int rw, int S,
                             unsigned long long seed, float pos_ratio, float flip_p, float shift_max, float shift_p,
                             float noise_std, float noise_p, float rot_p, float rot_range, float cd_p, int holes, int hz,
                             int hy, int hx, float grid_p, int d1, int d2, double grid_ratio, int grid_invert,
                             float* meta, cudaStream_t st) {
    if (rd > D || rh > H || rw > W || S < 1 || rd < 1 || rh < 1 || rw < 1) return -1;
    if (holes < 0 || holes > kMaxHoles || hz < 1 || hy < 1 || hx < 1 || hz > rd || hy > rh || hx > rw) return -1;
    if (d1 < 1 || d2 <= d1 || grid_ratio < 0.0) return -1;
    const long long V = (long long)D * H * W;
    const int nb = (int)((V + kBlockVox - 1) / kBlockVox);
    pick_centers_kernel<<<(S + 7) / 8, 256, 0, st>>>(label, counts, nb, D, H, W, rd, rh, rw, S, seed, pos_ratio, flip_p, shift_max,
                                           shift_p, noise_std, noise_p, rot_p, rot_range, cd_p, holes, hz, hy, hx,
                                           grid_p, d1, d2, grid_ratio, grid_invert, meta);
    return (int)cudaGetLastError();
}

// img [C][D][H][W] fp32, label [D][H][W] fp32 -> out_img [S][C][rd][rh][rw], out_lab [S][1][rd][rh][rw] (fp32, NCDHW);
// (hz, hy, hx) = the hole size fcd_pick_centers was given, hh = ceil(sqrt(rd^2 + rh^2 + rw^2)) (utils/gridmask.py:31)
FCD_API int fcd_crop_augment(const float* img, const float* label, int C, int D, int H, int W, int rd, int rh, int rw,
                             int S, const float* meta, unsigned long long seed, int hz, int hy, int hx, int hh,
                             float* out_img, float* out_lab, cudaStream_t st) {
    if (rd > D || rh > H || rw > W || S < 1 || S > 65535 || C < 1 || (long long)D * H * W > 0x7fffffffLL) return -1;
    if (hh < rd || hh < rh || hh < rw) return -1;
    // threads per row: what rw needs (4 voxels each), rounded up to whole warps, at most the block
    int tpr = (((rw + 3) / 4 + 31) / 32) * 32;
    if (tpr > 128) tpr = 128;
    if (tpr == 96) tpr = 128;                            // rows per block must divide the block
    const int rpb = 128 / tpr;
    const long long gy = ((long long)rd * rh + rpb - 1) / rpb;
    if (gy > 65535) return -1;
    dim3 grid((rw + 4 * tpr - 1) / (4 * tpr), (unsigned)gy, S);
    crop_augment_kernel<<<grid, 128, 0, st>>>(img, label, C, D, H, W, rd, rh, rw, meta, seed, hz, hy, hx, hh, tpr,
                                              out_img, out_lab);
    return (int)cudaGetLastError();
}
