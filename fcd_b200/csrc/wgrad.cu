// Weight-gradient kernels for the conv family (mma.sync bf16, fp32 accumulate, deterministic split-M reduction).
//
//   dWp[split][t][n][k] = sum_{m in split} Q[m][n] * P[src(m, t)][k]
//
// For a forward conv (reference conv_blocks.py:393-437 via autograd) Q = dY on the output grid, P = X on the input
// grid with src = m*stride + tap - pad.  For ConvTranspose3d k2 s2 (conv_blocks.py:640-649) the roles swap:
// Q = X on the coarse grid, P = dY on the fine grid, src = 2m + tap (i.e. the wgrad of a k2 s2 p0 conv).
// One CTA owns a 16(n) x 16(k) output block for up to TP taps; its 8 warps split the voxel (contraction) axis and
// are reduced through shared memory at the end, so no atomics are needed.  fcd_wgrad_reduce sums the splits and
// writes the fp32 gradient in the parameter's own (PyTorch) layout.
#include "common.cuh"

struct WGradParams {
    const bf16* Q; long long ldq;
    const bf16* P; long long ldp;
    float* out;              // [nsplit][T][Np][Kp]
    int Bn, Ds, Hs, Ws, Dm, Hm, Wm;
    int Np, Kp, kd, kh, kw, stride, pad;
    int M, nsplit, tiles_per_split;
};

namespace {

constexpr int WBM = 128;   // voxels per stage (8 warps x 16)

__device__ __forceinline__ int wswz(int row, int chunk) { return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4); }

// G: 16-voxel groups per warp per stage.  The pointwise launches on big volumes (G = 4) were bound by their two
// __syncthreads per 128-voxel stage (two ldmatrix + two mma per warp between barriers), not by the loads.
template <int TP, int G>
__global__ void __launch_bounds__(256) wgrad_kernel(const WGradParams p) {
    constexpr int WT = WBM * G;               // voxels per stage
    constexpr int Q_BYTES = WT * 32;
    constexpr int STAGE_BYTES = Q_BYTES * (1 + TP);
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.kd * p.kh * p.kw;
    const int npass = (T + TP - 1) / TP;
    const int nb = p.Np / 16, kb = p.Kp / 16;
    int unit = blockIdx.y;
    const int pass = unit % npass; unit /= npass;
    const int kblk = unit % kb;
    const int nblk = unit / kb;
    (void)nb;
    const int t0 = pass * TP;
    const int ntap = min(TP, T - t0);
    const int split = blockIdx.x;
    const int tile0 = split * p.tiles_per_split;
    const int ntiles_total = (p.M + WT - 1) / WT;
    const int tile1 = min(tile0 + p.tiles_per_split, ntiles_total);

    // each thread loads one 16B chunk of Q and of every tap's P tile: row = tid/2, chunk = tid&1
    const int lrow = tid >> 1, lch = tid & 1;

    auto load_stage = [&](int tile, int slot) {
        const uint32_t sq = sbase + slot * STAGE_BYTES;
#pragma unroll
        for (int gq = 0; gq < G; ++gq) {
            const int row = gq * WBM + lrow;
            const int m = tile * WT + row;
            const bool ok = m < p.M;
            const bf16* qsrc = p.Q;
            if (ok) qsrc += (long long)m * p.ldq + nblk * 16 + lch * 8;
            cp_async16(sq + wswz(row, lch), qsrc, ok);
            int mm = ok ? m : 0;
            const int x = mm % p.Wm; mm /= p.Wm;
            const int y = mm % p.Hm; mm /= p.Hm;
            const int z = mm % p.Dm; mm /= p.Dm;
            const long long base = (long long)mm * p.Ds * p.Hs * p.Ws;
#pragma unroll
            for (int j = 0; j < TP; ++j) {
                if (j < ntap) {
                    const int t = t0 + j;
                    const int tx = t % p.kw, ty = (t / p.kw) % p.kh, tz = t / (p.kw * p.kh);
                    const int sz = z * p.stride + tz - p.pad, sy = y * p.stride + ty - p.pad, sx = x * p.stride + tx - p.pad;
                    const bool v = ok && sz >= 0 && sz < p.Ds && sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws;
                    const bf16* src = p.P;
                    if (v) src += (base + ((long long)sz * p.Hs + sy) * p.Ws + sx) * p.ldp + kblk * 16 + lch * 8;
                    cp_async16(sq + Q_BYTES * (1 + j) + wswz(row, lch), src, v);
                }
            }
        }
    };

    float acc[TP][2][4];
#pragma unroll
    for (int j = 0; j < TP; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j][i][k] = 0.f;

    if (tile0 < tile1) load_stage(tile0, 0);
    cp_async_commit();
    for (int tile = tile0; tile < tile1; ++tile) {
        const int slot = (tile - tile0) & 1;
        if (tile + 1 < tile1) load_stage(tile + 1, slot ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t sq = sbase + slot * STAGE_BYTES;
#pragma unroll
        for (int gq = 0; gq < G; ++gq) {
            // A fragment: Q^T block (16 n x 16 voxels of this warp)
            uint32_t a[4];
            {
                int v = gq * WBM + warp * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = (lane >> 3) & 1;
                ldmatrix_x4_trans(a[0], a[1], a[2], a[3], sq + wswz(v, ch));
            }
#pragma unroll
            for (int j = 0; j < TP; ++j) {
                if (j < ntap) {
                    uint32_t b0, b1, b2, b3;
                    int v = gq * WBM + warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    int ch = lane >> 4;
                    ldmatrix_x4_trans(b0, b1, b2, b3, sq + Q_BYTES * (1 + j) + wswz(v, ch));
                    mma_bf16_16816(acc[j][0], a, b0, b1);
                    mma_bf16_16816(acc[j][1], a, b2, b3);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- cross-warp reduction, one tap at a time: red[warp][16][16]
    float* red = reinterpret_cast<float*>(smem);
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
        if (j < ntap) {
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                int col = ni * 8 + tq * 2;
                red[warp * 256 + g * 16 + col] = acc[j][ni][0];
                red[warp * 256 + g * 16 + col + 1] = acc[j][ni][1];
                red[warp * 256 + (g + 8) * 16 + col] = acc[j][ni][2];
                red[warp * 256 + (g + 8) * 16 + col + 1] = acc[j][ni][3];
            }
        }
        __syncthreads();
        if (j < ntap) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w * 256 + tid];
            const int n = nblk * 16 + (tid >> 4), k = kblk * 16 + (tid & 15);
            p.out[(((long long)split * T + (t0 + j)) * p.Np + n) * p.Kp + k] = s;
        }
        __syncthreads();
    }
}

template <int TP, int G>
int launch_wgrad(const WGradParams& p, cudaStream_t stream) {
    constexpr int pipe = 2 * WBM * G * 32 * (1 + TP);
    constexpr int red = 8 * 256 * 4;
    constexpr int smem = pipe > red ? pipe : red;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(wgrad_kernel<TP, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    const int T = p.kd * p.kh * p.kw;
    const int npass = (T + TP - 1) / TP;
    dim3 grid(p.nsplit, (p.Np / 16) * (p.Kp / 16) * npass);
    wgrad_kernel<TP, G><<<grid, 256, smem, stream>>>(p);
    return (int)cudaGetLastError();
}

// out[n*sn + kmap(k)*sk + t*st] = sum_split part[split][t][n][k];  kmap undoes the concat-segment padding.
// Block = (256/ng) outputs x ng split groups (ng = 1, 2, 4, 8 chosen from nsplit): group g adds splits g, g+ng, ...
// (coalesced rows, 4 loads in flight), the group sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                           int nsplit, int T, int N, int K, int Np, int Kp,
                                                           long long sn, long long sk, long long st, int kseg,
                                                           int ksegpad, int accumulate, int ng) {
    __shared__ float sh[256];
    const int per = 256 / ng;
    const int li = threadIdx.x % per, g = threadIdx.x / per;
    const long long total = (long long)T * N * K;
    const long long sstride = (long long)T * Np * Kp;
    for (long long i0 = blockIdx.x * (long long)per; i0 < total; i0 += (long long)gridDim.x * per) {
        const long long i = i0 + li;
        float s = 0.f;
        long long o = 0;
        if (i < total) {
            const int k = (int)(i % K);
            const long long r = i / K;
            const int n = (int)(r % N);
            const int t = (int)(r / N);
            const int kp = (k / kseg) * ksegpad + (k % kseg);
            const float* src = part + ((long long)t * Np + n) * Kp + kp;
            o = n * sn + k * sk + t * st;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int sp = g;
            for (; sp + 3 * ng < nsplit; sp += 4 * ng) {
                a0 += src[sp * sstride];
                a1 += src[(sp + ng) * sstride];
                a2 += src[(sp + 2 * ng) * sstride];
                a3 += src[(sp + 3 * ng) * sstride];
            }
            for (; sp < nsplit; sp += ng) a0 += src[sp * sstride];
            s = (a0 + a1) + (a2 + a3);
        }
        if (ng == 1) {
            if (i < total) out[o] = accumulate ? out[o] + s : s;
            continue;
        }
        sh[threadIdx.x] = s;
        __syncthreads();
        if (g == 0 && i < total) {
            float tot = 0.f;
            for (int q = 0; q < ng; ++q) tot += sh[q * per + li];
            out[o] = accumulate ? out[o] + tot : tot;
        }
        __syncthreads();
    }
}

// dst[t][n][kp] (bf16, zero padded) = src[n*sn + k*sk + t*st]
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int T, int N, int K, int Np,
                                   int Kp, long long sn, long long sk, long long st, int kseg, int ksegpad, int nseg,
                                   int nsegpad) {
    const long long total = (long long)T * Np * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int kp = (int)(i % Kp);
        long long r = i / Kp;
        int np_ = (int)(r % Np);
        int t = (int)(r / Np);
        float v = 0.f;
        int seg = kp / ksegpad, within = kp % ksegpad;
        int k = seg * kseg + within;
        int nsg = np_ / nsegpad, nwithin = np_ % nsegpad;
        int n = nsg * nseg + nwithin;
        if (nwithin < nseg && n < N && within < kseg && k < K) v = src[n * sn + k * sk + t * st];
        dst[i] = __float2bfloat16(v);
    }
}

// One launch packs MANY weights: the host keeps a table of pack jobs (one per conv layer and layout) and every block
// finds its job by binary search over the jobs' first-block indices.  Same arithmetic as pack_weight_kernel.
struct PackJob {
    const float* src;
    bf16* dst;
    long long sn, sk, st, total;
    int T, N, K, Np, Kp, kseg, ksegpad, nseg, nsegpad, blk0;
    int pad_[2];
};
static_assert(sizeof(PackJob) == 96, "PackJob layout is mirrored by fcd_b200/ops.py");
constexpr int PACK_NB = 8, PACK_KB = 64;      // one block packs an 8 (n) x 64 (k) tile of every tap

__device__ __forceinline__ uint32_t pack_bf16x2_raw(bf16 lo, bf16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Tile body for a compile-time tap count (index math by constants; the per-row segment maps come from shared memory).
template <int T>
__device__ __forceinline__ void pack_tile(const PackJob& j, int n0, int k0, const int* nmap, const int* kmap, bf16* tile) {
    constexpr int TP = T | 1;                           // odd pitch: conflict-free in both passes
    constexpr int NE = PACK_NB * PACK_KB * T;
    const float* __restrict__ src = j.src;
    const int sn = (int)j.sn, sk = (int)j.sk, st = (int)j.st;
    // Interior tiles of 3x3x3 weights (all 8 n and 64 k real and contiguous in the source: the bulk of the bytes) are
    // read as float4 with almost no index math; ragged / segmented tiles take the generic element loop below.
    bool fast = false;
    if constexpr (T == 27) {
        fast = st == 1 && kmap[0] >= 0 && kmap[PACK_KB - 1] == kmap[0] + PACK_KB - 1 && nmap[0] >= 0 &&
               nmap[PACK_NB - 1] == nmap[0] + PACK_NB - 1;
        if (fast && sk == T) {                          // k next to the taps: 8 runs of 64 * 27 floats
            fast = (sn % 4) == 0 && ((reinterpret_cast<uintptr_t>(src + (long long)nmap[0] * sn + kmap[0] * T)) & 15) == 0;
            if (fast) {
                constexpr int RUN = PACK_KB * T;        // == the shared-memory pitch of one n (TP == T)
                for (int nn = 0; nn < PACK_NB; ++nn) {
                    const float* base = src + (long long)(nmap[0] + nn) * sn + kmap[0] * T;
                    for (int i = threadIdx.x * 4; i < RUN; i += 1024) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(base + i));
                        uint2 o;
                        o.x = pack_bf16x2_raw(__float2bfloat16(v.x), __float2bfloat16(v.y));
                        o.y = pack_bf16x2_raw(__float2bfloat16(v.z), __float2bfloat16(v.w));
                        *reinterpret_cast<uint2*>(tile + nn * RUN + i) = o;
                    }
                }
            }
        } else if (fast && sn == T) {                   // n next to the taps: 64 runs of 8 * 27 floats
            fast = (sk % 4) == 0 && ((reinterpret_cast<uintptr_t>(src + (long long)kmap[0] * sk + nmap[0] * T)) & 15) == 0;
            if (fast) {
                constexpr int RUN4 = PACK_NB * T / 4;   // float4 per run (216 / 4)
                for (int idx = threadIdx.x; idx < PACK_KB * RUN4; idx += 256) {
                    const int kk = idx / RUN4, j0 = (idx % RUN4) * 4;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)(kmap[0] + kk) * sk + nmap[0] * T + j0));
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = j0 + u, nn = jj / T, t = jj % T;
                        tile[(nn * PACK_KB + kk) * TP + t] = __float2bfloat16(vv[u]);
                    }
                }
            }
        } else {
            fast = false;
        }
    }
    if (fast) {
        // tile filled above
    } else if (j.sn < j.sk) {                           // n is the faster source axis (data-gradient packs)
        for (int e = threadIdx.x; e < NE; e += 256) {
            const int t = e % T, r = e / T;
            const int nn = r % PACK_NB, kk = r / PACK_NB;
            const int n = nmap[nn], k = kmap[kk];
            const float v = (n >= 0 && k >= 0) ? __ldg(src + n * sn + k * sk + t * st) : 0.f;
            tile[(nn * PACK_KB + kk) * TP + t] = __float2bfloat16(v);
        }
    } else {
        for (int e = threadIdx.x; e < NE; e += 256) {
            const int t = e % T, r = e / T;
            const int kk = r % PACK_KB, nn = r / PACK_KB;
            const int n = nmap[nn], k = kmap[kk];
            const float v = (n >= 0 && k >= 0) ? __ldg(src + n * sn + k * sk + t * st) : 0.f;
            tile[(nn * PACK_KB + kk) * TP + t] = __float2bfloat16(v);
        }
    }
    __syncthreads();
    // 8 consecutive k per thread: one 16-byte store (Kp % 8 == 0, so an 8-group is entirely inside or outside)
    constexpr int NG = PACK_NB * (PACK_KB / 8) * T;
    for (int g = threadIdx.x; g < NG; g += 256) {
        const int k8 = g % (PACK_KB / 8), r = g / (PACK_KB / 8);
        const int nn = r % PACK_NB, t = r / PACK_NB;
        const int np_ = n0 + nn, kp = k0 + k8 * 8;
        if (np_ >= j.Np || kp >= j.Kp) continue;
        const bf16* tp = tile + (nn * PACK_KB + k8 * 8) * TP + t;
        uint4 o;
        o.x = pack_bf16x2_raw(tp[0], tp[TP]);
        o.y = pack_bf16x2_raw(tp[2 * TP], tp[3 * TP]);
        o.z = pack_bf16x2_raw(tp[4 * TP], tp[5 * TP]);
        o.w = pack_bf16x2_raw(tp[6 * TP], tp[7 * TP]);
        *reinterpret_cast<uint4*>(j.dst + ((long long)t * j.Np + np_) * j.Kp + kp) = o;
    }
}

// The source is fp32 [.., T] with the tap index fastest and either k (forward packs: sk == T) or n (data-gradient
// packs: sn == T) next.  A block reads its tile as long contiguous runs in whichever order the source has,
// transposes through shared memory and writes 128-byte runs of the packed [T][Np][Kp] bf16 layout.  (The first
// version computed ~6 runtime integer divisions per element and was instruction-bound: 362 us for MS_DSA_NET's
// 87 M packed elements, at the head of every step.)
__global__ void __launch_bounds__(256) pack_weight_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
    __shared__ __align__(16) bf16 tile[27 * PACK_NB * PACK_KB];
    __shared__ int kmap[PACK_KB], nmap[PACK_NB];
    __shared__ PackJob sj;
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {                                   // last job with blk0 <= blockIdx.x
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].blk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    if (threadIdx.x < sizeof(PackJob) / 4)
        reinterpret_cast<int*>(&sj)[threadIdx.x] = reinterpret_cast<const int*>(jobs + lo)[threadIdx.x];
    __syncthreads();
    const PackJob& j = sj;
    const int kt = (j.Kp + PACK_KB - 1) / PACK_KB;
    const int b = blockIdx.x - j.blk0;
    const int n0 = (b / kt) * PACK_NB, k0 = (b % kt) * PACK_KB;
    if (threadIdx.x < PACK_KB) {                        // padded k -> source k (or -1): concat-segment map
        const int kp = k0 + threadIdx.x;
        const int seg = kp / j.ksegpad, within = kp % j.ksegpad;
        const int k = seg * j.kseg + within;
        kmap[threadIdx.x] = (kp < j.Kp && within < j.kseg && k < j.K) ? k : -1;
    } else if (threadIdx.x < PACK_KB + PACK_NB) {
        const int nn = threadIdx.x - PACK_KB, np_ = n0 + nn;
        const int nsg = np_ / j.nsegpad, nwithin = np_ % j.nsegpad;
        const int n = nsg * j.nseg + nwithin;
        nmap[nn] = (np_ < j.Np && nwithin < j.nseg && n < j.N) ? n : -1;
    }
    __syncthreads();
    switch (j.T) {
        case 27: pack_tile<27>(j, n0, k0, nmap, kmap, tile); break;
        case 8: pack_tile<8>(j, n0, k0, nmap, kmap, tile); break;
        case 1: pack_tile<1>(j, n0, k0, nmap, kmap, tile); break;
        default: break;                                 // the host only builds jobs with T in {1, 8, 27}
    }
}

}  // namespace

// jobs: device array of `njobs` PackJob records (96 bytes each, see above), nblocks = sum over jobs of
// ceil(Np / 8) * ceil(Kp / 64); T in {1, 8, 27}; Kp % 8 == 0 and dst 16-byte aligned.  Replaces one fcd_pack_weight launch per layer and layout with one launch per forward.
FCD_API int fcd_pack_weight_batched(const void* jobs, int njobs, int nblocks, cudaStream_t stream) {
    if (njobs < 1 || nblocks < 1) return -1;
    pack_weight_batched_kernel<<<nblocks, 256, 0, stream>>>((const PackJob*)jobs, njobs);
    FCD_LAUNCH_CHECK();
}

// Weight gradient of Conv3d / ConvTranspose3d / Linear (autograd of the modules cited in igemm.cu).
// `part` must hold nsplit*T*Np*Kp floats.  tiles_per_split = ceil(ceil(M/128)/nsplit).
FCD_API int fcd_wgrad(const void* Q, long long ldq, const void* P, long long ldp, float* part, int Bn, int Ds, int Hs,
                      int Ws, int Dm, int Hm, int Wm, int Np, int Kp, int kd, int kh, int kw, int stride, int pad,
                      int nsplit, cudaStream_t stream) {
    if (Np % 16 != 0 || Kp % 16 != 0 || ldq % 8 != 0 || ldp % 8 != 0 || nsplit < 1) return -1;
    WGradParams p;
    p.Q = (const bf16*)Q; p.ldq = ldq; p.P = (const bf16*)P; p.ldp = ldp; p.out = part;
    p.Bn = Bn; p.Ds = Ds; p.Hs = Hs; p.Ws = Ws; p.Dm = Dm; p.Hm = Hm; p.Wm = Wm;
    p.Np = Np; p.Kp = Kp; p.kd = kd; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad;
    long long M = (long long)Bn * Dm * Hm * Wm;
    if (M <= 0 || M > 0x7fffffffLL) return -1;
    p.M = (int)M;
    const int T = kd * kh * kw;
    const int G = fcd_wgrad_group(M, T);
    int ntiles = (p.M + WBM * G - 1) / (WBM * G);
    if (nsplit > ntiles) return -1;
    p.nsplit = nsplit;
    p.tiles_per_split = (ntiles + nsplit - 1) / nsplit;
    if (T == 1) return G == 4 ? launch_wgrad<1, 4>(p, stream) : launch_wgrad<1, 1>(p, stream);
    if (T == 8) return launch_wgrad<8, 1>(p, stream);
    return launch_wgrad<9, 1>(p, stream);
}

// 16-voxel groups per warp per pipeline stage of fcd_wgrad: 4 for one-tap (1x1x1 conv / linear) gradients on at least 2^18
// voxels, else 1.  A stage holds 128 * group voxels: callers size nsplit <= ceil(M / (128 * group)).
FCD_API int fcd_wgrad_group(long long M, int T) { return (T == 1 && M >= (1LL << 18)) ? 4 : 1; }

FCD_API int fcd_wgrad_reduce(const float* part, float* out, int nsplit, int T, int N, int K, int Np, int Kp,
                             long long sn, long long sk, long long st, int kseg, int ksegpad, int accumulate,
                             cudaStream_t stream) {
    long long total = (long long)T * N * K;
    // few outputs and many splits: several split groups per output; many outputs: one thread per output
    int ng = 1;
    while (ng < 8 && ng * 2 <= nsplit / 4 && total * ng < 256LL * 4 * fcd_num_sms()) ng *= 2;
    const int per = 256 / ng;
    int blocks = (int)((total + per - 1) / per);
    if (blocks > 16 * fcd_num_sms()) blocks = 16 * fcd_num_sms();
    if (blocks < 1) blocks = 1;
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(part, out, nsplit, T, N, K, Np, Kp, sn, sk, st, kseg, ksegpad,
                                                    accumulate, ng);
    FCD_LAUNCH_CHECK();
}

// fp32 parameter (PyTorch layout, described by strides) -> bf16 packed [T][Np][Kp], K contiguous, zero padded.
FCD_API int fcd_pack_weight(const float* src, void* dst, int T, int N, int K, int Np, int Kp, long long sn,
                            long long sk, long long st, int kseg, int ksegpad, int nseg, int nsegpad,
                            cudaStream_t stream) {
    long long total = (long long)T * Np * Kp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    if (blocks < 1) blocks = 1;
    pack_weight_kernel<<<blocks, 256, 0, stream>>>(src, (bf16*)dst, T, N, K, Np, Kp, sn, sk, st, kseg, ksegpad, nseg,
                                                   nsegpad);
    FCD_LAUNCH_CHECK();
}
