// Dual self-attention (DSA, sa_type='parallel') and the token-wise ops around it, forward and backward.
//
// Reference: networks/ms_dsa_net/conv_blocks.py -- TransformerBlock.forward 69-90 (pos_embed add, LayerNorm,
// x + gamma * DSA(LN(x))) and DSA.forward 328-355.  All of it is linear in the token count N: the spatial branch
// projects K and V_SA from N to P tokens with the learned EF[N,P]; the channel branch is a c x c Gram matrix.
// So the whole block is a pair of streaming passes over the tokens (a reduction pass and an apply pass), HBM-bound,
// and the [N,P] attention map (8.4 M elements per block at level 3) is never materialised.
//
// Layouts: tokens are channels-last rows r = b*N + n (bf16, row stride ld, true channel count C = h*c, h heads);
// qkvv rows hold [q | k | v_CA | v_SA], each C wide, head-major (conv_blocks.py:330-332).
// The reference's line 353 reinterprets x_SA's [B,c,h,N] memory as [B,N,C] (a scramble); we store
// x_SA directly in that flat [c][h][N] order (tsa) so that the "scramble" is a plain flat read.
#include "common.cuh"

namespace {

// counter-based Bernoulli keep-mask for the spatial-attention dropout (attn_drop_2, conv_blocks.py:352):
// the same (seed, element) always gives the same bit, so backward regenerates the forward's mask.
__device__ __forceinline__ bool sa_keep(unsigned long long seed, unsigned long long elem, uint32_t thresh) {
    unsigned long long z = seed + elem * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (uint32_t)(z >> 32) >= thresh;
}

// acc[0..P) += s * row[0..P) and dot(v, row) with `row` a 16-byte aligned shared-memory row read as float4: every lane
// reads the same address (broadcast), and a scalar read per FMA made these kernels LDS-issue bound (1 LDS : 1 FFMA).
template <int P>
__device__ __forceinline__ void fma_row(float* acc, float s, const float* row) {
#pragma unroll
    for (int p = 0; p < P; p += 4) {
        const float4 k = *reinterpret_cast<const float4*>(row + p);
        acc[p] = fmaf(s, k.x, acc[p]);
        acc[p + 1] = fmaf(s, k.y, acc[p + 1]);
        acc[p + 2] = fmaf(s, k.z, acc[p + 2]);
        acc[p + 3] = fmaf(s, k.w, acc[p + 3]);
    }
}
template <int P>
__device__ __forceinline__ float dot_row(const float* v, const float* row) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int p = 0; p < P; p += 4) {
        const float4 k = *reinterpret_cast<const float4*>(row + p);
        a0 = fmaf(v[p], k.x, a0);
        a1 = fmaf(v[p + 1], k.y, a1);
        a0 = fmaf(v[p + 2], k.z, a0);
        a1 = fmaf(v[p + 3], k.w, a1);
    }
    return a0 + a1;
}

// Per-token row segments (one head's CH channels) as 16-byte accesses: a thread's segment is contiguous but the
// threads of a warp sit on different rows, so a scalar access touches 32 sectors for 2-4 bytes each (and a scalar
// bf16 store is a partial-sector write); CH % 8 == 0 on every MS_DSA_NET level.
template <int CH>
__device__ __forceinline__ void ld_seg_bf16(const bf16* p, float* out) {
    if constexpr (CH % 8 == 0) {
#pragma unroll
        for (int k = 0; k < CH; k += 8) unpack8(ld8(p + k), out + k);
    } else {
#pragma unroll
        for (int k = 0; k < CH; ++k) out[k] = __bfloat162float(p[k]);
    }
}
template <int CH>
__device__ __forceinline__ void st_seg_bf16(bf16* p, const float* in) {
    if constexpr (CH % 8 == 0) {
#pragma unroll
        for (int k = 0; k < CH; k += 8) st8(p + k, pack8(in + k));
    } else {
#pragma unroll
        for (int k = 0; k < CH; ++k) p[k] = __float2bfloat16(in[k]);
    }
}
template <int CH>
__device__ __forceinline__ void ld_seg_f32(const float* p, float* out) {
    if constexpr (CH % 4 == 0) {
#pragma unroll
        for (int k = 0; k < CH; k += 4) *reinterpret_cast<float4*>(out + k) = *reinterpret_cast<const float4*>(p + k);
    } else {
#pragma unroll
        for (int k = 0; k < CH; ++k) out[k] = p[k];
    }
}
template <int CH>
__device__ __forceinline__ void st_seg_f32(float* p, const float* in) {
    if constexpr (CH % 4 == 0) {
#pragma unroll
        for (int k = 0; k < CH; k += 4) *reinterpret_cast<float4*>(p + k) = *reinterpret_cast<const float4*>(in + k);
    } else {
#pragma unroll
        for (int k = 0; k < CH; ++k) p[k] = in[k];
    }
}

// block-cooperative global -> shared copy of n floats (n % 4 == 0, both 16-byte aligned) as float4
__device__ __forceinline__ void stage_f4(float* dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
}

__device__ __forceinline__ float group_sum(float v, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------ LayerNorm (+pos)
// t = x + pos ; ln = LayerNorm_C(t) * w + b.   One thread per (row, 8-channel chunk); LPT = Cp/8 lanes per row.
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ pos,
                              const float* __restrict__ w, const float* __restrict__ bvec, bf16* __restrict__ t,
                              long long ldt, bf16* __restrict__ ln, long long ldl, float* __restrict__ mean,
                              float* __restrict__ rstd, long long rows, int N, int C, int LPT, float eps) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long row = gid / LPT;
    const int chunk = (int)(gid % LPT);
    const bool live = row < rows;
    if (!live) row = rows - 1;
    const int n = (int)(row % N);
    float v[8];
    unpack8(ld8(x + row * ldx + chunk * 8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = chunk * 8 + k;
        if (c < C) {
            if (pos) v[k] += pos[(long long)n * C + c];
            v[k] = __bfloat162float(__float2bfloat16(v[k]));
        } else {
            v[k] = 0.f;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    s = group_sum(s, LPT);
    const float m = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = chunk * 8 + k;
        if (c < C) q += (v[k] - m) * (v[k] - m);
    }
    q = group_sum(q, LPT);
    const float r = rsqrtf(q / C + eps);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = chunk * 8 + k;
        o[k] = c < C ? (v[k] - m) * r * w[c] + bvec[c] : 0.f;
    }
    if (live) {
        st8(t + row * ldt + chunk * 8, pack8(v));
        st8(ln + row * ldl + chunk * 8, pack8(o));
        if (chunk == 0) { mean[row] = m; rstd[row] = r; }
    }
}

// dx = dt_direct + LN-backward(dln);  dpos[n][c] = sum_b dx[b][n][c];  per-block partials of dw, db.
// One thread per (n, chunk), looping over the batch.  part: [gridDim.x][2][Cp].
__global__ void ln_bwd_kernel(const bf16* __restrict__ dln, long long lddl, const bf16* __restrict__ dtd,
                              long long lddt, const bf16* __restrict__ t, long long ldt,
                              const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ w, bf16* __restrict__ dx, long long lddx,
                              float* __restrict__ dpos, float* __restrict__ part, int B, int N, int C, int LPT) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = gid / LPT;
    const int chunk = (int)(gid % LPT);
    const bool live = n < N;
    if (!live) n = N - 1;
    float wv[8], accp[8], accw[8], accb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = chunk * 8 + k;
        wv[k] = c < C ? w[c] : 0.f;
        accp[k] = accw[k] = accb[k] = 0.f;
    }
    for (int b = 0; b < B; ++b) {
        const long long row = (long long)b * N + n;
        float g[8], tv[8], d[8], xh[8];
        unpack8(ld8(dln + row * lddl + chunk * 8), g);
        unpack8(ld8(t + row * ldt + chunk * 8), tv);
        unpack8(ld8(dtd + row * lddt + chunk * 8), d);
        const float m = mean[row], r = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = chunk * 8 + k;
            xh[k] = c < C ? (tv[k] - m) * r : 0.f;
            if (live) { accw[k] = fmaf(g[k], xh[k], accw[k]); accb[k] += (c < C ? g[k] : 0.f); }
            g[k] *= wv[k];
            s1 += g[k];
            s2 = fmaf(g[k], xh[k], s2);
        }
        s1 = group_sum(s1, LPT) / C;
        s2 = group_sum(s2, LPT) / C;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = chunk * 8 + k;
            o[k] = c < C ? r * (g[k] - s1 - xh[k] * s2) + d[k] : 0.f;
            accp[k] += o[k];
        }
        if (live) st8(dx + row * lddx + chunk * 8, pack8(o));
    }
    if (live && dpos) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = chunk * 8 + k;
            if (c < C) dpos[n * C + c] = accp[k];
        }
    }
    __shared__ float sh[256 * 16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sh[threadIdx.x * 16 + k] = accw[k]; sh[threadIdx.x * 16 + 8 + k] = accb[k]; }
    __syncthreads();
    if ((int)threadIdx.x < LPT) {          // thread `chunk` == threadIdx.x sums its column over the block's rows
        float a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = 0.f;
        for (int r = threadIdx.x; r < (int)blockDim.x; r += LPT)
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] += sh[r * 16 + k];
        const int Cp = LPT * 8;
        float* o = part + (long long)blockIdx.x * 2 * Cp;
#pragma unroll
        for (int k = 0; k < 8; ++k) { o[threadIdx.x * 8 + k] = a[k]; o[Cp + threadIdx.x * 8 + k] = a[8 + k]; }
    }
}

// out0[c] = sum_k part[k][0][c], out1[c] = sum_k part[k][1][c]   (c < C).  One warp per column: lanes stride over
// the nblk partial rows, fp64 accumulation, fixed shuffle order (deterministic).
__global__ void __launch_bounds__(256) pair_colsum_kernel(const float* __restrict__ part, int nblk, int Cp, int C,
                                                          float* __restrict__ out0, float* __restrict__ out1) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
    for (int k = lane; k < nblk; k += 32) { a += part[(long long)k * 2 * Cp + c]; b += part[(long long)k * 2 * Cp + Cp + c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        if (out0) out0[c] = (float)a;
        if (out1) out1[c] = (float)b;
    }
}

// ------------------------------------------------------------------------------------------ DSA forward
// Partial layout per (b, tile): [Sqq: C][Skk: C][G: h*c*c][KP|VP: 2C*P]
__device__ __forceinline__ long long dsa_osize(int C, int c, int P) { return 2LL * C + (long long)C * c + 2LL * C * P; }

__global__ void __launch_bounds__(256) dsa_reduce_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                         const float* __restrict__ EF, float* __restrict__ part,
                                                         int N, int C, int c, int P, int TN) {
    extern __shared__ __align__(16) float sm[];
    const int W3 = 3 * C;                       // [q | k | v_SA] per token
    float* sQ = sm;                             // [TN][3C]
    float* sE = sm + (long long)TN * W3;        // [TN][P]
    const int b = blockIdx.y, tile = blockIdx.x;
    const int n0 = tile * TN;
    // token rows as 16-byte loads (C % 8 == 0 is checked by the host wrapper; ldq % 8 == 0)
    const int W8 = W3 / 8;
    for (int i = threadIdx.x; i < TN * W8; i += blockDim.x) {
        const int n = i / W8, col = (i % W8) * 8;
        float v[8];
        if (n0 + n < N) {
            const int src = col < 2 * C ? col : col + C;        // skip v_CA (third slot)
            unpack8(ld8(qkvv + ((long long)b * N + n0 + n) * ldq + src), v);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
        }
        *reinterpret_cast<float4*>(&sQ[n * W3 + col]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(&sQ[n * W3 + col + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    const int P4 = P / 4;
    for (int i = threadIdx.x; i < TN * P4; i += blockDim.x) {
        const int n = i / P4;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n < N) e = *reinterpret_cast<const float4*>(&EF[(long long)(n0 + n) * P + (i % P4) * 4]);
        *reinterpret_cast<float4*>(&sE[i * 4]) = e;
    }
    __syncthreads();
    const int ntiles = gridDim.x;
    float* out = part + ((long long)b * ntiles + tile) * dsa_osize(C, c, P);
    // Outputs as 4 x 4 register tiles (two float4 reads per 16 FMAs; one scalar + one float4 read per 4 FMAs made the
    // kernel LDS-bound): [k | v_SA] rows x P columns, then the per-head q k^T blocks when c % 4 == 0, then the norms.
    const bool g_tiled = (c % 4) == 0;
    const int c4 = c / 4;
    const int n_kv = (2 * C / 4) * P4;
    const int n_g = g_tiled ? (C / c) * c4 * c4 : C * c;
    const int n_s = 2 * C;
    // gridDim.z > 1 (small N, few token tiles): the blocks of one tile share the token tile and split the outputs
    for (int it = blockIdx.z * blockDim.x + threadIdx.x; it < n_kv + n_g + n_s; it += blockDim.x * gridDim.z) {
        if (it < n_kv || (g_tiled && it < n_kv + n_g)) {
            int offA, offB, ldo, strideB;
            float* o;
            if (it < n_kv) {
                const int r4 = it / P4, p4 = it % P4;
                offA = C + 4 * r4; offB = TN * W3 + 4 * p4; strideB = P;           // B operand lives in sE
                o = out + 2 * C + (long long)C * c + (long long)(4 * r4) * P + 4 * p4; ldo = P;
            } else {
                const int g = it - n_kv;
                const int j4 = g % c4, i4 = (g / c4) % c4, hd = g / (c4 * c4);
                offA = hd * c + 4 * i4; offB = C + hd * c + 4 * j4; strideB = W3;
                o = out + 2 * C + (long long)hd * c * c + (long long)(4 * i4) * c + 4 * j4; ldo = c;
            }
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
#pragma unroll 4
            for (int n = 0; n < TN; ++n) {
                const float4 av = *reinterpret_cast<const float4*>(&sm[n * W3 + offA]);
                const float4 bv = *reinterpret_cast<const float4*>(&sm[offB + n * strideB]);
                const float ar[4] = {av.x, av.y, av.z, av.w}, bc[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(ar[r], bc[q], acc[r][q]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<float4*>(&o[(long long)r * ldo]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        } else if (it < n_kv + n_g) {
            const int g = it - n_kv;
            const int j = g % c, i = (g / c) % c, hd = g / (c * c);
            float a = 0.f;
            for (int n = 0; n < TN; ++n) a = fmaf(sQ[n * W3 + hd * c + i], sQ[n * W3 + C + hd * c + j], a);
            out[2 * C + g] = a;
        } else {
            const int r = it - n_kv - n_g;
            float a = 0.f;
            for (int n = 0; n < TN; ++n) a = fmaf(sQ[n * W3 + r], sQ[n * W3 + r], a);
            out[r] = a;
        }
    }
}

// In-place first level of the tile sums: for each of `nsets` partial sets (rows of O floats, `ntiles` rows per set),
// row g (g < G) becomes the sum of rows g, g+G, g+2G, ...  Column j of row g is read and written by one thread only,
// so this is race-free, deterministic, coalesced, and leaves the finalize kernels a G-row loop instead of ntiles.
constexpr int kTileGroups = 8;
__global__ void __launch_bounds__(128) tile_group_sum_kernel(float* __restrict__ part, int ntiles, long long O) {
    const long long j = blockIdx.x * 128LL + threadIdx.x;
    const int g = blockIdx.y;
    if (j >= O || g >= ntiles) return;
    float* base = part + (long long)blockIdx.z * ntiles * O + j;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = g;
    for (; k + 3 * kTileGroups < ntiles; k += 4 * kTileGroups) {
        a0 += base[(long long)k * O];
        a1 += base[(long long)(k + kTileGroups) * O];
        a2 += base[(long long)(k + 2 * kTileGroups) * O];
        a3 += base[(long long)(k + 3 * kTileGroups) * O];
    }
    for (; k < ntiles; k += kTileGroups) a0 += base[(long long)k * O];
    base[(long long)g * O] = (a0 + a1) + (a2 + a3);
}
static inline void tile_group_sum(float* part, int nsets, int ntiles, long long O, cudaStream_t st) {
    if (ntiles <= kTileGroups) return;
    dim3 grid((unsigned)((O + 127) / 128), kTileGroups, nsets);
    tile_group_sum_kernel<<<grid, 128, 0, st>>>(part, ntiles, O);
}
static inline int tile_rows_after_sum(int ntiles) { return ntiles <= kTileGroups ? ntiles : kTileGroups; }

// grid (h, B).  Sums the tile partials and produces: inv_nq/inv_nk [B][C], Ghat [B][h][c][c], A = softmax rows,
// KP/VP [B][2][C][P] (k projection then v_SA projection).
__global__ void __launch_bounds__(256) dsa_finalize_kernel(const float* __restrict__ part, int ntiles, int nsum,
                                                           const float* __restrict__ temperature,
                                                           const float* __restrict__ ca_scale,
                                                           float* __restrict__ inv_n, float* __restrict__ Ghat,
                                                           float* __restrict__ A, float* __restrict__ Ad,
                                                           float* __restrict__ KV, int C, int c, int P) {
    extern __shared__ __align__(16) float sm[];
    float* sG = sm;               // [c][c]
    float* sN = sm + c * c;       // [2][c]
    const int hd = blockIdx.x, b = blockIdx.y;
    const long long O = dsa_osize(C, c, P);
    const float* base = part + (long long)b * ntiles * O;
    auto tsum = [&](long long off) {           // rows [0, nsum) hold the group sums (tile_group_sum_kernel)
        float s = 0.f;
        for (int k = 0; k < nsum; ++k) s += base[(long long)k * O + off];
        return s;
    };
    // the K/V projection copy is split over the gridDim.z blocks of this (head, sample); block z == 0 also does the
    // small-matrix part (on the deep levels this kernel is a link of a latency chain: 8 blocks took ~35 us)
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < 2 * c * P; i += blockDim.x * gridDim.z) {
        const int which = i / (c * P), rem = i % (c * P);
        const int j = rem / P, p = rem % P;
        const long long r = (long long)which * C + hd * c + j;
        KV[((long long)b * 2 * C + r) * P + p] = tsum(2LL * C + (long long)C * c + r * P + p);
    }
    if (blockIdx.z != 0) return;
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) {
        const int which = i / c, j = i % c;
        const float s = tsum((long long)which * C + hd * c + j);
        const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);          // F.normalize eps (conv_blocks.py:342-343)
        sN[i] = inv;
        inv_n[((long long)b * 2 + which) * C + hd * c + j] = inv;
    }
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) sG[i] = tsum(2LL * C + (long long)hd * c * c + i);
    __syncthreads();
    const float tau = temperature[hd];
    // softmax rows: one warp per row, lanes over the columns
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < c; i += nwarp) {
        float mx = -INFINITY;
        for (int j = lane; j < c; j += 32) {
            const float g = sG[i * c + j] * sN[i] * sN[c + j];
            sG[i * c + j] = g;
            mx = fmaxf(mx, g * tau);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int j = lane; j < c; j += 32) den += __expf(sG[i * c + j] * tau - mx);
        den = warp_sum(den);
        const long long o = (((long long)b * gridDim.x + hd) * c + i) * c;
        for (int j = lane; j < c; j += 32) {
            Ghat[o + j] = sG[i * c + j];
            const float a = __expf(sG[i * c + j] * tau - mx) / den;
            A[o + j] = a;
            Ad[o + j] = ca_scale ? a * ca_scale[o + j] : a;     // attn_drop (conv_blocks.py:347)
        }
    }
}

// thread per (token, head): spatial softmax over P in registers, channel attention apply.
// xca [B*N][C] fp32 (head-merged), tsa [B][c][h][N] fp32 (the reference's scrambled memory order).
template <int CH, int P>
__global__ void __launch_bounds__(128) dsa_apply_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                        const float* __restrict__ inv_n, const float* __restrict__ A,
                                                        const float* __restrict__ KV,
                                                        const float* __restrict__ temperature2,
                                                        float* __restrict__ xca, float* __restrict__ tsa, int N, int C,
                                                        float drop_scale, uint32_t drop_thresh,
                                                        unsigned long long seed0,
                                                        const long long* __restrict__ seed_dev) {
    // seed_dev: device step counter (ticks once per training forward) mixed into the seed, so a captured CUDA graph
    // draws a NEW mask on every replay while forward and backward of one step still agree
    const unsigned long long seed = seed0 + (seed_dev ? (unsigned long long)(*seed_dev) * 0xD1B54A32D192ED03ULL : 0ULL);
    extern __shared__ __align__(16) float sm[];
    float* sKP = sm;                    // [CH][P]
    float* sVP = sm + CH * P;           // [CH][P]
    float* sA = sm + 2 * CH * P;        // [CH][CH]
    float* sIq = sA + CH * CH;          // [CH]
    const int hd = blockIdx.y, b = blockIdx.z, H = gridDim.y;
    stage_f4(sKP, KV + ((long long)b * 2 * C + hd * CH) * P, CH * P);
    stage_f4(sVP, KV + ((long long)b * 2 * C + C + hd * CH) * P, CH * P);
    if constexpr (CH % 2 == 0) stage_f4(sA, A + ((long long)b * H + hd) * CH * CH, CH * CH);
    else for (int i = threadIdx.x; i < CH * CH; i += blockDim.x) sA[i] = A[((long long)b * H + hd) * CH * CH + i];
    for (int i = threadIdx.x; i < CH; i += blockDim.x) sIq[i] = inv_n[(long long)b * 2 * C + hd * CH + i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const bf16* row = qkvv + ((long long)b * N + n) * ldq;
    const float tau2 = temperature2[hd];
    float lg[P];
#pragma unroll
    for (int p = 0; p < P; ++p) lg[p] = 0.f;
    constexpr int CK = CH % 8 == 0 ? 8 : CH;              // channels per 16-byte row segment
    for (int j0 = 0; j0 < CH; j0 += CK) {
        float q[CK];
        ld_seg_bf16<CK>(row + hd * CH + j0, q);
#pragma unroll
        for (int jj = 0; jj < CK; ++jj) fma_row<P>(lg, q[jj] * sIq[j0 + jj], sKP + (j0 + jj) * P);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < P; ++p) { lg[p] *= tau2; mx = fmaxf(mx, lg[p]); }
    float den = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) { lg[p] = __expf(lg[p] - mx); den += lg[p]; }
    float inv = 1.f / den;
    if (drop_thresh) {
        const unsigned long long e0 = (((unsigned long long)b * H + hd) * N + n) * P;
#pragma unroll
        for (int p = 0; p < P; ++p) lg[p] = sa_keep(seed, e0 + p, drop_thresh) ? lg[p] : 0.f;
        inv *= drop_scale;
    }
    for (int j = 0; j < CH; ++j) {
        const float a = dot_row<P>(lg, sVP + j * P);
        tsa[(((long long)b * CH + j) * H + hd) * N + n] = a * inv;
    }
    float v[CH];
    ld_seg_bf16<CH>(row + 2 * C + hd * CH, v);
    for (int i0 = 0; i0 < CH; i0 += CK) {
        float o[CK];
#pragma unroll
        for (int ii = 0; ii < CK; ++ii) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < CH; ++j) a = fmaf(sA[(i0 + ii) * CH + j], v[j], a);
            o[ii] = a;
        }
        st_seg_f32<CK>(xca + ((long long)b * N + n) * C + hd * CH + i0, o);
    }
}

// y = t + gamma * (xca + tsa_flat)   (conv_blocks.py:77).  thread per (row, chunk).
__global__ void dsa_combine_kernel(const bf16* __restrict__ t, long long ldt, const float* __restrict__ gamma,
                                   const float* __restrict__ xca, const float* __restrict__ tsa, bf16* __restrict__ y,
                                   long long ldy, long long rows, int N, int C, int LPT) {
    const long long total = rows * LPT;
    for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
         gid += (long long)gridDim.x * blockDim.x) {
        const long long row = gid / LPT;
        const int chunk = (int)(gid % LPT);
        const long long b = row / N, n = row % N;
        float v[8];
        unpack8(ld8(t + row * ldt + chunk * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = chunk * 8 + k;
            if (ch < C) v[k] += gamma[ch] * (xca[row * C + ch] + tsa[b * N * C + n * C + ch]);
            else v[k] = 0.f;
        }
        st8(y + row * ldy + chunk * 8, pack8(v));
    }
}

// ------------------------------------------------------------------------------------------ DSA backward
// dgamma partials: part[blk][0][c] = sum_rows dy * (xca + tsa_flat).
__global__ void dsa_dgamma_kernel(const bf16* __restrict__ dy, long long lddy, const float* __restrict__ xca,
                                  const float* __restrict__ tsa, float* __restrict__ part, long long rows, int N,
                                  int C, int LPT, float* __restrict__ dtemp, float* __restrict__ dtemp2, int H) {
    // first kernel of fcd_dsa_bwd: also clears the two accumulators dsa_bwd_finalize_kernel adds into (saves the
    // caller two fill launches per block on a latency-bound chain)
    if (blockIdx.x == 0 && (int)threadIdx.x < H) { dtemp[threadIdx.x] = 0.f; dtemp2[threadIdx.x] = 0.f; }
    const int chunk = threadIdx.x % LPT, r0 = threadIdx.x / LPT, rstep = blockDim.x / LPT;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (long long row = (long long)blockIdx.x * rstep + r0; row < rows; row += (long long)gridDim.x * rstep) {
        const long long b = row / N, n = row % N;
        float g[8];
        unpack8(ld8(dy + row * lddy + chunk * 8), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = chunk * 8 + k;
            if (ch < C) acc[k] = fmaf(g[k], xca[row * C + ch] + tsa[b * N * C + n * C + ch], acc[k]);
        }
    }
    __shared__ float sh[256 * 8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[threadIdx.x * 8 + k] = acc[k];
    __syncthreads();
    if ((int)threadIdx.x < LPT) {
        const int Cp = LPT * 8;
        float* o = part + (long long)blockIdx.x * 2 * Cp;
        for (int k = 0; k < 8; ++k) {
            float s = 0.f;
            for (int r = threadIdx.x; r < (int)blockDim.x; r += LPT) s += sh[r * 8 + k];
            o[threadIdx.x * 8 + k] = s;
            o[Cp + threadIdx.x * 8 + k] = 0.f;
        }
    }
}

// Backward partial layout per (b, hd, tile): [dVP: c*P][dKP: c*P][dA: c*c][R1: c][dT2: 1]
__device__ __forceinline__ long long dsa_bsize(int c, int P) { return 2LL * c * P + (long long)c * c + c + 1; }

// Phase 1: thread per token (recompute the spatial softmax, form dlogits, write dqhat_sa); phase 2: the block
// reduces the outer products over its TN tokens.
template <int CH, int P, int NT>
__global__ void __launch_bounds__(NT) dsa_bwd_reduce_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                            const bf16* __restrict__ dy, long long lddy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ inv_n,
                                                            const float* __restrict__ KV,
                                                            const float* __restrict__ temperature2,
                                                            float* __restrict__ dqh, float* __restrict__ part, int N,
                                                            int C, float drop_scale, uint32_t drop_thresh,
                                                            unsigned long long seed0,
                                                            const long long* __restrict__ seed_dev) {
    const unsigned long long seed = seed0 + (seed_dev ? (unsigned long long)(*seed_dev) * 0xD1B54A32D192ED03ULL : 0ULL);
    constexpr int TN = 64;
    constexpr int ROW = (2 * P + 5 * CH + 1 + 3) & ~3;      // multiple of 4 floats: rows are float4-addressable
    extern __shared__ __align__(16) float sm[];
    float* sKP = sm;                        // [CH][P]
    float* sVP = sm + CH * P;               // [CH][P]
    float* sT = sm + 2 * CH * P;            // [TN][ROW]: a[P] | dlog[P] | dxs[CH] | qh[CH] | dxca[CH] | vca[CH] | r1[CH] | dt2
    const int tile = blockIdx.x, hd = blockIdx.y, b = blockIdx.z, H = gridDim.y, ntiles = gridDim.x;
    stage_f4(sKP, KV + ((long long)b * 2 * C + hd * CH) * P, CH * P);
    stage_f4(sVP, KV + ((long long)b * 2 * C + C + hd * CH) * P, CH * P);
    __syncthreads();
    // phase 1 runs on the first TN threads (one token each); all NT threads take part in the phase-2 reduction
    // (NT = 64 when there are enough token tiles to fill the GPU, 256 on the deep levels with a handful of blocks)
    const bool tok = threadIdx.x < TN;
    const int n = tile * TN + threadIdx.x;
    float* my = sT + (tok ? threadIdx.x : 0) * ROW;
    if (tok && n < N) {
        const bf16* row = qkvv + ((long long)b * N + n) * ldq;
        const float tau2 = temperature2[hd];
        // live per-thread arrays: lg[P] and da[P] only
        float lg[P];
#pragma unroll
        for (int p = 0; p < P; ++p) lg[p] = 0.f;
        // the token's row segments arrive as 16-byte loads and are parked in its shared-memory row; the channel loops
        // below stay rolled (unrolling them lets the compiler hoist every KP/VP row: 255 registers + KBs of spills)
        constexpr int CK = CH % 8 == 0 ? 8 : CH;          // channels per 16-byte row segment
        for (int j0 = 0; j0 < CH; j0 += CK) {
            float q8[CK], g8[CK], v8[CK];
            ld_seg_bf16<CK>(row + hd * CH + j0, q8);
            ld_seg_bf16<CK>(dy + ((long long)b * N + n) * lddy + hd * CH + j0, g8);
            ld_seg_bf16<CK>(row + 2 * C + hd * CH + j0, v8);
#pragma unroll
            for (int jj = 0; jj < CK; ++jj) {
                const int j = j0 + jj;
                my[2 * P + CH + j] = q8[jj] * inv_n[(long long)b * 2 * C + hd * CH + j];
                my[2 * P + 2 * CH + j] = gamma[hd * CH + j] * g8[jj];
                my[2 * P + 3 * CH + j] = v8[jj];
            }
        }
#pragma unroll 1
        for (int j = 0; j < CH; ++j) fma_row<P>(lg, my[2 * P + CH + j], sKP + j * P);
        float mx = -INFINITY;
#pragma unroll
        for (int p = 0; p < P; ++p) { lg[p] *= tau2; mx = fmaxf(mx, lg[p]); }
        float den = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) { lg[p] = __expf(lg[p] - mx); den += lg[p]; }
        const float inv = 1.f / den;
#pragma unroll
        for (int p = 0; p < P; ++p) lg[p] *= inv;                       // a[p]
        float da[P];
#pragma unroll
        for (int p = 0; p < P; ++p) da[p] = 0.f;
#pragma unroll 1
        for (int j = 0; j < CH; ++j) {
            // gradient of the scrambled x_SA element: flat index f -> (n', ch')
            const long long f = ((long long)j * H + hd) * N + n;
            const long long n2 = f / C;
            const int ch2 = (int)(f % C);
            const float dxs = gamma[ch2] * __bfloat162float(dy[((long long)b * N + n2) * lddy + ch2]);
            my[2 * P + j] = dxs;
            fma_row<P>(da, dxs, sVP + j * P);
        }
        float s = 0.f;
        const unsigned long long e0 = (((unsigned long long)b * H + hd) * N + n) * P;
#pragma unroll
        for (int p = 0; p < P; p += 4) {
            float4 kept;
            float* kv = reinterpret_cast<float*>(&kept);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float kf = drop_thresh ? (sa_keep(seed, e0 + p + u, drop_thresh) ? drop_scale : 0.f) : 1.f;
                da[p + u] *= kf;
                s = fmaf(lg[p + u], da[p + u], s);
                kv[u] = lg[p + u] * kf;                                   // dropped-out attention row (for dVP)
            }
            *reinterpret_cast<float4*>(my + p) = kept;
        }
#pragma unroll
        for (int p = 0; p < P; p += 4) {
            float4 d4;
            float* dv = reinterpret_cast<float*>(&d4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dv[u] = lg[p + u] * (da[p + u] - s);
                da[p + u] = dv[u];                                        // reuse as dlog
            }
            *reinterpret_cast<float4*>(my + P + p) = d4;
        }
        // dt2 = sum_p dlog[p] * raw[p] with raw[p] = sum_j qh[j] KP[j][p]  ==  sum_j qh[j] * (sum_p dlog[p] KP[j][p])
        float dt2 = 0.f;
        constexpr int C4 = CH % 4 == 0 ? 4 : 1;           // dqh leaves as float4 stores
#pragma unroll 1
        for (int j0 = 0; j0 < CH; j0 += C4) {
            float o[C4];
#pragma unroll
            for (int jj = 0; jj < C4; ++jj) {
                const int j = j0 + jj;
                float a = dot_row<P>(da, sKP + j * P);
                const float qh = my[2 * P + CH + j];
                dt2 = fmaf(qh, a, dt2);
                a *= tau2;
                o[jj] = a;
                my[2 * P + 4 * CH + j] = a * qh;
            }
            st_seg_f32<C4>(dqh + ((long long)b * N + n) * C + hd * CH + j0, o);
        }
        my[2 * P + 5 * CH] = dt2;
    } else if (tok) {
        for (int i = 0; i < ROW; ++i) my[i] = 0.f;
    }
    __syncthreads();
    float* out = part + (((long long)b * H + hd) * ntiles + tile) * dsa_bsize(CH, P);
    const float tau2 = temperature2[hd];
    constexpr int n_vp = CH * P, n_a = CH * CH;
    if constexpr (CH % 4 == 0) {
        // outer-product reductions over the TN tokens as 4 x 4 register tiles: two float4 reads per 16 FMAs
        // (the scalar version issued two LDS per FMA and dominated the kernel)
        constexpr int JT = CH / 4, PT = P / 4, NVP = JT * PT, NA = JT * JT;
        for (int ti = threadIdx.x; ti < 2 * NVP + NA; ti += blockDim.x) {
            int offA, offB, ldo;
            float scale = 1.f;
            float* o;
            if (ti < NVP) {                                // dVP[j][p] = sum a[p] dxs[j]
                const int j4 = ti / PT, p4 = ti % PT;
                offA = 2 * P + 4 * j4; offB = 4 * p4; o = out + 4 * j4 * P + 4 * p4; ldo = P;
            } else if (ti < 2 * NVP) {                     // dKP[j][p] = tau2 sum qh[j] dlog[p]
                const int u = ti - NVP, j4 = u / PT, p4 = u % PT;
                offA = 2 * P + CH + 4 * j4; offB = P + 4 * p4; o = out + n_vp + 4 * j4 * P + 4 * p4; ldo = P;
                scale = tau2;
            } else {                                       // dA[i][j] = sum dxca[i] vca[j]
                const int u = ti - 2 * NVP, i4 = u / JT, j4 = u % JT;
                offA = 2 * P + 2 * CH + 4 * i4; offB = 2 * P + 3 * CH + 4 * j4; o = out + 2 * n_vp + 4 * i4 * CH + 4 * j4;
                ldo = CH;
            }
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 8
            for (int q = 0; q < TN; ++q) {
                const float4 av = *reinterpret_cast<const float4*>(sT + q * ROW + offA);
                const float4 bv = *reinterpret_cast<const float4*>(sT + q * ROW + offB);
                const float ar[4] = {av.x, av.y, av.z, av.w}, bc[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], bc[c], acc[r][c]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[r * ldo + c] = acc[r][c] * scale;
        }
        for (int it = threadIdx.x; it < CH + 1; it += blockDim.x) {        // R1[j], dt2
            float a = 0.f;
            for (int q = 0; q < TN; ++q) a += sT[q * ROW + 2 * P + 4 * CH + it];
            out[2 * n_vp + n_a + it] = a;
        }
    } else {
        for (int it = threadIdx.x; it < 2 * n_vp + n_a + CH + 1; it += blockDim.x) {
            float a = 0.f;
            if (it < n_vp) {                                   // dVP[j][p] = sum a[p] dxs[j]
                const int j = it / P, p = it % P;
                for (int q = 0; q < TN; ++q) a = fmaf(sT[q * ROW + p], sT[q * ROW + 2 * P + j], a);
            } else if (it < 2 * n_vp) {                        // dKP[j][p] = tau2 sum qh[j] dlog[p]
                const int j = (it - n_vp) / P, p = (it - n_vp) % P;
                for (int q = 0; q < TN; ++q) a = fmaf(sT[q * ROW + 2 * P + CH + j], sT[q * ROW + P + p], a);
                a *= tau2;
            } else if (it < 2 * n_vp + n_a) {                  // dA[i][j] = sum dxca[i] vca[j]
                const int i = (it - 2 * n_vp) / CH, j = (it - 2 * n_vp) % CH;
                for (int q = 0; q < TN; ++q) a = fmaf(sT[q * ROW + 2 * P + 2 * CH + i], sT[q * ROW + 2 * P + 3 * CH + j], a);
            } else {                                           // R1[j], dt2
                const int j = it - 2 * n_vp - n_a;
                for (int q = 0; q < TN; ++q) a += sT[q * ROW + 2 * P + 4 * CH + j];
            }
            out[it] = a;
        }
    }
}

// grid (h, B): sums tile partials; softmax backward of the channel attention; emits
// dKV [B][2][C][P] (dKP then dVP), dGhat [B][h][c][c], rqk [B][2][C] and accumulates dtemperature / dtemperature2.
__global__ void __launch_bounds__(256) dsa_bwd_finalize_kernel(const float* __restrict__ part, int ntiles, int nsum,
                                                               const float* __restrict__ temperature,
                                                               const float* __restrict__ Ghat,
                                                               const float* __restrict__ A,
                                                               const float* __restrict__ ca_scale,
                                                               float* __restrict__ dKV,
                                                               float* __restrict__ dGhat, float* __restrict__ rqk,
                                                               float* __restrict__ dtemp, float* __restrict__ dtemp2,
                                                               int C, int c, int P) {
    extern __shared__ __align__(16) float sm[];
    float* sdA = sm;                 // [c][c] -> dGhat
    float* sR1 = sm + c * c;         // [c]
    __shared__ float sred[256];
    const int hd = blockIdx.x, b = blockIdx.y, H = gridDim.x;
    const long long O = dsa_bsize(c, P);
    const float* base = part + ((long long)b * H + hd) * ntiles * O;
    auto tsum = [&](long long off) {
        float s = 0.f;
        for (int k = 0; k < nsum; ++k) s += base[(long long)k * O + off];
        return s;
    };
    // the dKV copy is split over the gridDim.z blocks; block z == 0 also does the small-matrix part
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < 2 * c * P; i += blockDim.x * gridDim.z) {
        const int which = i / (c * P), rem = i % (c * P);          // partial: 0 dVP, 1 dKP
        const int j = rem / P, p = rem % P;
        const int dst_which = which == 0 ? 1 : 0;                   // dKV: 0 dKP, 1 dVP
        dKV[(((long long)b * 2 + dst_which) * C + hd * c + j) * P + p] = tsum(i);
    }
    if (blockIdx.z != 0) return;
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) {
        const float v = tsum(2LL * c * P + i);
        sdA[i] = ca_scale ? v * ca_scale[((long long)b * H + hd) * c * c + i] : v;
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) sR1[i] = tsum(2LL * c * P + (long long)c * c + i);
    float t2 = 0.f;
    if (threadIdx.x == 0) t2 = tsum(O - 1);
    __syncthreads();
    const float tau = temperature[hd];
    const long long o = ((long long)b * H + hd) * c * c;
    float dtau = 0.f;
    // softmax backward rows: one warp per row, lanes over the columns (coalesced A / Ghat reads)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < c; i += nwarp) {
        float s = 0.f;
        for (int j = lane; j < c; j += 32) s = fmaf(A[o + i * c + j], sdA[i * c + j], s);
        s = warp_sum(s);
        float rq = 0.f;
        for (int j = lane; j < c; j += 32) {
            const float dS = A[o + i * c + j] * (sdA[i * c + j] - s);     // d(tau * Ghat)
            const float gh = Ghat[o + i * c + j];
            dtau = fmaf(dS, gh, dtau);
            const float dg = tau * dS;
            sdA[i * c + j] = dg;
            dGhat[o + i * c + j] = dg;
            rq = fmaf(dg, gh, rq);
        }
        rq = warp_sum(rq);
        if (lane == 0) rqk[((long long)b * 2) * C + hd * c + i] = rq + sR1[i];
    }
    __syncthreads();
    // rk[j] = sum_i dGhat[i][j] Ghat[i][j]: the row range is split over blockDim / c thread groups
    {
        const int j = threadIdx.x % c, k = threadIdx.x / c, K = blockDim.x / c;     // c <= 256 (host check)
        float rk = 0.f;
        if (k < K)
            for (int i = k; i < c; i += K) rk = fmaf(sdA[i * c + j], Ghat[o + i * c + j], rk);
        __shared__ float sp[256];              // [K][c] partials, K * c <= 256
        if (k < K) sp[k * c + j] = rk;
        __syncthreads();
        if ((int)threadIdx.x < c) {
            float r = 0.f;
            for (int kk = 0; kk < K; ++kk) r += sp[kk * c + threadIdx.x];
            rqk[((long long)b * 2 + 1) * C + hd * c + threadIdx.x] = r;
        }
    }
    sred[threadIdx.x] = dtau;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)blockDim.x; ++i) s += sred[i];
        atomicAdd(&dtemp[hd], s);
        atomicAdd(&dtemp2[hd], t2);
    }
}

// thread per (token, head): gradients of q, k, v_CA, v_SA -> dqkvv rows (bf16).  SPLIT (deep levels, a few hundred
// tokens): CH/8 adjacent threads share a token and each produces one 8-channel segment of every output, which cuts the
// serial FMA chain of the latency-bound small-N launches by CH/8.
template <int CH, int P, bool SPLIT>
__global__ void __launch_bounds__(128) dsa_bwd_apply_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                            const bf16* __restrict__ dy, long long lddy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ EF,
                                                            const float* __restrict__ inv_n,
                                                            const float* __restrict__ A,
                                                            const float* __restrict__ dGhat,
                                                            const float* __restrict__ rqk,
                                                            const float* __restrict__ dKV,
                                                            const float* __restrict__ dqh, bf16* __restrict__ dqkvv,
                                                            long long lddq, int N, int C) {
    extern __shared__ __align__(16) float sm[];
    float* sdKP = sm;                     // [CH][P]
    float* sdVP = sm + CH * P;            // [CH][P]
    float* sA = sm + 2 * CH * P;          // [CH][CH]
    float* sdG = sA + CH * CH;            // [CH][CH]
    float* sV = sdG + CH * CH;            // inv_nq | inv_nk | rq | rk : [4][CH]
    const int hd = blockIdx.y, b = blockIdx.z, H = gridDim.y;
    stage_f4(sdKP, dKV + ((long long)b * 2 * C + hd * CH) * P, CH * P);
    stage_f4(sdVP, dKV + ((long long)b * 2 * C + C + hd * CH) * P, CH * P);
    if constexpr (CH % 2 == 0) {
        stage_f4(sA, A + ((long long)b * H + hd) * CH * CH, CH * CH);
        stage_f4(sdG, dGhat + ((long long)b * H + hd) * CH * CH, CH * CH);
    } else {
        for (int i = threadIdx.x; i < CH * CH; i += blockDim.x) {
            sA[i] = A[((long long)b * H + hd) * CH * CH + i];
            sdG[i] = dGhat[((long long)b * H + hd) * CH * CH + i];
        }
    }
    for (int i = threadIdx.x; i < CH; i += blockDim.x) {
        sV[i] = inv_n[(long long)b * 2 * C + hd * CH + i];
        sV[CH + i] = inv_n[(long long)b * 2 * C + C + hd * CH + i];
        sV[2 * CH + i] = rqk[(long long)b * 2 * C + hd * CH + i];
        sV[3 * CH + i] = rqk[(long long)b * 2 * C + C + hd * CH + i];
    }
    __syncthreads();
    constexpr int NSEG = (SPLIT && CH % 8 == 0) ? CH / 8 : 1;     // threads per token
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) / NSEG;
    const int seg = threadIdx.x % NSEG;
    if (n >= N) return;
    const long long r = (long long)b * N + n;
    const bf16* row = qkvv + r * ldq;
    bf16* drow = dqkvv + r * lddq;
    constexpr int CK = CH % 8 == 0 ? 8 : CH;              // channels per 16-byte row segment
    float qh[CH], kh[CH];
    ld_seg_bf16<CH>(row + hd * CH, qh);
    ld_seg_bf16<CH>(row + C + hd * CH, kh);
#pragma unroll
    for (int j = 0; j < CH; ++j) { qh[j] *= sV[j]; kh[j] *= sV[CH + j]; }
    // dq, one row segment at a time (the segment's own q values are re-read: no dynamic register indexing)
    const int c_lo = NSEG > 1 ? seg * CK : 0, c_hi = NSEG > 1 ? c_lo + CK : CH;      // this thread's output channels
    for (int i0 = c_lo; i0 < c_hi; i0 += CK) {
        float o[CK], q8[CK];
        ld_seg_f32<CK>(dqh + r * C + hd * CH + i0, o);
        ld_seg_bf16<CK>(row + hd * CH + i0, q8);
#pragma unroll
        for (int ii = 0; ii < CK; ++ii) {
            const int i = i0 + ii;
            float a = o[ii];
#pragma unroll
            for (int j = 0; j < CH; ++j) a = fmaf(sdG[i * CH + j], kh[j], a);
            o[ii] = (a - q8[ii] * sV[i] * sV[2 * CH + i]) * sV[i];
        }
        st_seg_bf16<CK>(drow + hd * CH + i0, o);
    }
    {
        float ef[P];
        ld_seg_f32<P>(EF + (long long)n * P, ef);
        for (int j0 = c_lo; j0 < c_hi; j0 += CK) {
            float o[CK], f[CK], k8[CK];
            ld_seg_bf16<CK>(row + C + hd * CH + j0, k8);
#pragma unroll
            for (int jj = 0; jj < CK; ++jj) {
                const int j = j0 + jj;
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < CH; ++i) a = fmaf(sdG[i * CH + j], qh[i], a);
                a = (a - k8[jj] * sV[CH + j] * sV[3 * CH + j]) * sV[CH + j];
                o[jj] = a + dot_row<P>(ef, sdKP + j * P);                 // dk (+ its EF term)
                f[jj] = dot_row<P>(ef, sdVP + j * P);                     // dv_SA
            }
            st_seg_bf16<CK>(drow + C + hd * CH + j0, o);
            st_seg_bf16<CK>(drow + 3 * C + hd * CH + j0, f);
        }
    }
    // dv_CA[j] = sum_i A[i][j] dxca[i]   (reuse qh as dxca)
    ld_seg_bf16<CH>(dy + r * lddy + hd * CH, qh);
#pragma unroll
    for (int i = 0; i < CH; ++i) qh[i] *= gamma[hd * CH + i];
    for (int j0 = c_lo; j0 < c_hi; j0 += CK) {
        float o[CK];
#pragma unroll
        for (int jj = 0; jj < CK; ++jj) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < CH; ++i) a = fmaf(sA[i * CH + j0 + jj], qh[i], a);
            o[jj] = a;
        }
        st_seg_bf16<CK>(drow + 2 * C + hd * CH + j0, o);
    }
}

// Small-N variant (deep levels: N <= 1024 tokens but up to 256 channels): thread per (token, 4 p's, channel part);
// the 8 parts of one output live in 8 adjacent lanes, split the channel loop and are added by xor-shuffles.
__global__ void __launch_bounds__(256) dsa_bwd_ef_small_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                               const float* __restrict__ dKV,
                                                               float* __restrict__ dEF, int B, int N, int C, int P) {
    const int P4 = P / 4;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int part = (int)(gid & 7);
    const long long o = gid >> 3;
    const long long n = o / P4;
    const int p = (int)(o % P4) * 4;
    const bool live = n < N;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        for (int b = 0; b < B; ++b) {
            const float* dkp = dKV + (long long)b * 2 * C * P;
            const float* dvp = dkp + (long long)C * P;
            const bf16* row = qkvv + ((long long)b * N + n) * ldq;
            for (int c0 = part * 8; c0 < C; c0 += 64) {
                float kf[8], vf[8];
                unpack8(ld8(row + C + c0), kf);
                unpack8(ld8(row + 3 * C + c0), vf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(&dkp[(long long)(c0 + j) * P + p]);
                    const float4 y = *reinterpret_cast<const float4*>(&dvp[(long long)(c0 + j) * P + p]);
                    a.x = fmaf(kf[j], x.x, fmaf(vf[j], y.x, a.x));
                    a.y = fmaf(kf[j], x.y, fmaf(vf[j], y.y, a.y));
                    a.z = fmaf(kf[j], x.z, fmaf(vf[j], y.z, a.z));
                    a.w = fmaf(kf[j], x.w, fmaf(vf[j], y.w, a.w));
                }
            }
        }
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, d);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, d);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, d);
        a.w += __shfl_xor_sync(0xffffffffu, a.w, d);
    }
    if (live && part == 0) *reinterpret_cast<float4*>(&dEF[n * P + p]) = a;
}

// dEF[n][p] = sum_b sum_ch ( k[b,n,ch] dKP[b][ch][p] + v_SA[b,n,ch] dVP[b][ch][p] ).  Thread per (4 tokens, 4 p's):
// k / v_SA come in as 16 B (8-channel) loads, every dKP/dVP float4 (L1-resident) feeds 4 tokens.
__global__ void __launch_bounds__(256) dsa_bwd_ef_kernel(const bf16* __restrict__ qkvv, long long ldq,
                                                         const float* __restrict__ dKV, float* __restrict__ dEF,
                                                         int B, int N, int C, int P) {
    const int P4 = P / 4;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n0 = (gid / P4) * 4;
    const int p = (int)(gid % P4) * 4;
    if (n0 >= N) return;
    float4 a[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
        const float* dkp = dKV + (long long)b * 2 * C * P;
        const float* dvp = dkp + (long long)C * P;
        for (int c0 = 0; c0 < C; c0 += 8) {
            float kf[4][8], vf[4][8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const long long n = n0 + t < N ? n0 + t : N - 1;
                const bf16* row = qkvv + ((long long)b * N + n) * ldq;
                unpack8(ld8(row + C + c0), kf[t]);
                unpack8(ld8(row + 3 * C + c0), vf[t]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(&dkp[(long long)(c0 + j) * P + p]);
                const float4 y = *reinterpret_cast<const float4*>(&dvp[(long long)(c0 + j) * P + p]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    a[t].x = fmaf(kf[t][j], x.x, fmaf(vf[t][j], y.x, a[t].x));
                    a[t].y = fmaf(kf[t][j], x.y, fmaf(vf[t][j], y.y, a[t].y));
                    a[t].z = fmaf(kf[t][j], x.z, fmaf(vf[t][j], y.z, a[t].z));
                    a[t].w = fmaf(kf[t][j], x.w, fmaf(vf[t][j], y.w, a[t].w));
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (n0 + t < N) *reinterpret_cast<float4*>(&dEF[(n0 + t) * P + p]) = a[t];
}

template <int CH, int P>
int launch_apply(const bf16* qkvv, long long ldq, const float* inv_n, const float* A, const float* KV,
                 const float* t2, float* xca, float* tsa, int B, int N, int C, int H, float ds, uint32_t dth,
                 unsigned long long seed, const long long* seed_dev, cudaStream_t st) {
    const int smem = (2 * CH * P + CH * CH + CH) * 4;
    static bool conf = false;
    if (!conf) { cudaFuncSetAttribute(dsa_apply_kernel<CH, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); conf = true; }
    dim3 grid((N + 127) / 128, H, B);
    dsa_apply_kernel<CH, P><<<grid, 128, smem, st>>>(qkvv, ldq, inv_n, A, KV, t2, xca, tsa, N, C, ds, dth, seed, seed_dev);
    return (int)cudaGetLastError();
}

template <int CH, int P>
int launch_bwd_reduce(const bf16* qkvv, long long ldq, const bf16* dy, long long lddy, const float* gamma,
                      const float* inv_n, const float* KV, const float* t2, float* dqh, float* part, int B, int N,
                      int C, int H, float ds, uint32_t dth, unsigned long long seed, const long long* seed_dev,
                      cudaStream_t st) {
    const int smem = (2 * CH * P + 64 * ((2 * P + 5 * CH + 1 + 3) & ~3)) * 4;
    static bool conf = false;
    if (!conf) {
        cudaFuncSetAttribute(dsa_bwd_reduce_kernel<CH, P, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(dsa_bwd_reduce_kernel<CH, P, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        conf = true;
    }
    dim3 grid((N + 63) / 64, H, B);
    if ((long long)grid.x * H * B >= 2LL * fcd_num_sms())
        dsa_bwd_reduce_kernel<CH, P, 64><<<grid, 64, smem, st>>>(qkvv, ldq, dy, lddy, gamma, inv_n, KV, t2, dqh, part, N,
                                                                C, ds, dth, seed, seed_dev);
    else
        dsa_bwd_reduce_kernel<CH, P, 256><<<grid, 256, smem, st>>>(qkvv, ldq, dy, lddy, gamma, inv_n, KV, t2, dqh, part,
                                                                  N, C, ds, dth, seed, seed_dev);
    return (int)cudaGetLastError();
}

template <int CH, int P>
int launch_bwd_apply(const bf16* qkvv, long long ldq, const bf16* dy, long long lddy, const float* gamma,
                     const float* EF, const float* inv_n, const float* A, const float* dGhat, const float* rqk,
                     const float* dKV, const float* dqh, bf16* dqkvv, long long lddq, int B, int N, int C, int H,
                     cudaStream_t st) {
    const int smem = (2 * CH * P + 2 * CH * CH + 4 * CH) * 4;
    static bool conf = false;
    if (!conf) {
        cudaFuncSetAttribute(dsa_bwd_apply_kernel<CH, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(dsa_bwd_apply_kernel<CH, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        conf = true;
    }
    // few tokens (deep levels): one thread per (token, 8-channel segment)
    constexpr int NSEG = CH % 8 == 0 ? CH / 8 : 1;
    if (NSEG > 1 && (long long)N * H * B < 64LL * fcd_num_sms()) {
        dim3 grid((N * NSEG + 127) / 128, H, B);
        dsa_bwd_apply_kernel<CH, P, true><<<grid, 128, smem, st>>>(qkvv, ldq, dy, lddy, gamma, EF, inv_n, A, dGhat, rqk,
                                                                   dKV, dqh, dqkvv, lddq, N, C);
    } else {
        dim3 grid((N + 127) / 128, H, B);
        dsa_bwd_apply_kernel<CH, P, false><<<grid, 128, smem, st>>>(qkvv, ldq, dy, lddy, gamma, EF, inv_n, A, dGhat, rqk,
                                                                    dKV, dqh, dqkvv, lddq, N, C);
    }
    return (int)cudaGetLastError();
}

#define DSA_DISPATCH(FN, ...)                                                    \
    do {                                                                         \
        if (P == 64) {                                                           \
            switch (c) {                                                         \
                case 2: return FN<2, 64>(__VA_ARGS__);                           \
                case 4: return FN<4, 64>(__VA_ARGS__);                           \
                case 8: return FN<8, 64>(__VA_ARGS__);                           \
                case 16: return FN<16, 64>(__VA_ARGS__);                         \
                case 32: return FN<32, 64>(__VA_ARGS__);                         \
                default: return -1;                                              \
            }                                                                    \
        } else if (P == 32) {                                                    \
            switch (c) {                                                         \
                case 2: return FN<2, 32>(__VA_ARGS__);                           \
                case 4: return FN<4, 32>(__VA_ARGS__);                           \
                case 8: return FN<8, 32>(__VA_ARGS__);                           \
                case 16: return FN<16, 32>(__VA_ARGS__);                         \
                case 32: return FN<32, 32>(__VA_ARGS__);                         \
                case 64: return FN<64, 32>(__VA_ARGS__);                         \
                default: return -1;                                              \
            }                                                                    \
        }                                                                        \
        return -1;                                                               \
    } while (0)

// blocks per (head, sample) of the finalize kernels: ~512 copied K/V projection elements per block
int finalize_zsplit(int c, int P) {
    int z = (2 * c * P + 511) / 512;
    return z < 1 ? 1 : (z > 16 ? 16 : z);
}

int dsa_tile_tokens(int C, int P) {
    int tn = 64;
    while (tn > 8 && (long long)tn * (3 * C + P) * 4 > 60 * 1024) tn >>= 1;
    return tn;
}

}  // namespace

// ---- C ABI -----------------------------------------------------------------------------------------------------
// TransformerBlock.forward lines 72-77 up to LayerNorm: t = x + pos_embed, ln = LayerNorm(t) (conv_blocks.py:75-77).
FCD_API int fcd_ln_fwd(const void* x, long long ldx, const float* pos, const float* w, const float* b, void* t,
                       long long ldt, void* ln, long long ldl, float* mean, float* rstd, long long rows, int N, int C,
                       int Cp, float eps, cudaStream_t st) {
    const int LPT = Cp / 8;
    if (Cp % 8 || LPT > 32 || (LPT & (LPT - 1)) || C > Cp) return -1;
    const long long total = rows * LPT;
    const long long grid = (total + 255) / 256;
    ln_fwd_kernel<<<(unsigned)grid, 256, 0, st>>>((const bf16*)x, ldx, pos, w, b, (bf16*)t, ldt, (bf16*)ln, ldl, mean,
                                                  rstd, rows, N, C, LPT, eps);
    FCD_LAUNCH_CHECK();
}

FCD_API int fcd_ln_bwd_blocks(int N, int Cp) { return (int)(((long long)N * (Cp / 8) + 255) / 256); }

// part: fcd_ln_bwd_blocks(N,Cp)*2*Cp floats.  dpos fp32 [N][C] (or null), dw/db fp32 [C].
FCD_API int fcd_ln_bwd(const void* dln, long long lddl, const void* dtd, long long lddt, const void* t, long long ldt,
                       const float* mean, const float* rstd, const float* w, void* dx, long long lddx, float* dpos,
                       float* part, float* dw, float* db, int B, int N, int C, int Cp, cudaStream_t st) {
    const int LPT = Cp / 8;
    if (Cp % 8 || LPT > 32 || (LPT & (LPT - 1)) || C > Cp) return -1;
    const int nblk = fcd_ln_bwd_blocks(N, Cp);
    ln_bwd_kernel<<<nblk, 256, 0, st>>>((const bf16*)dln, lddl, (const bf16*)dtd, lddt, (const bf16*)t, ldt, mean, rstd,
                                        w, (bf16*)dx, lddx, dpos, part, B, N, C, LPT);
    pair_colsum_kernel<<<(C + 7) / 8, 256, 0, st>>>(part, nblk, Cp, C, dw, db);
    FCD_LAUNCH_CHECK();
}

static int dsa_tiles(int N, int C, int P) { const int tn = dsa_tile_tokens(C, P); return (N + tn - 1) / tn; }
FCD_API int fcd_dsa_fwd_part_floats(int B, int N, int C, int H, int P) {
    return (int)((long long)B * dsa_tiles(N, C, P) * (2LL * C + (long long)C * (C / H) + 2LL * C * P));
}
FCD_API int fcd_dsa_bwd_part_floats(int B, int N, int C, int H, int P) {
    const int c = C / H;
    return (int)((long long)B * H * ((N + 63) / 64) * (2LL * c * P + (long long)c * c + c + 1));
}

// Bernoulli keep-mask scaled by 1/(1-p) for the O(B*C) channel dropouts (nn.Dropout3d conv_blocks.py:57, attn_drop
// conv_blocks.py:347): one launch instead of torch's rand / compare / cast / divide chain.
__global__ void keep_scale_kernel(float* __restrict__ out, int n, float scale, uint32_t thresh, unsigned long long seed0,
                                  const long long* __restrict__ seed_dev) {
    const unsigned long long seed = seed0 + (seed_dev ? (unsigned long long)(*seed_dev) * 0xD1B54A32D192ED03ULL : 0ULL);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sa_keep(seed, (unsigned long long)i, thresh) ? scale : 0.f;
}

static inline void drop_params(float p, float& scale, uint32_t& thresh) {
    if (p <= 0.f) { scale = 1.f; thresh = 0; return; }
    scale = 1.f / (1.f - p);
    double t = (double)p * 4294967296.0;
    thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    if (thresh == 0) thresh = 1;
}

FCD_API int fcd_keep_scale(float* out, int n, float p, long long seed, const long long* seed_dev, cudaStream_t st) {
    if (n < 1 || p < 0.f || p >= 1.f) return -1;
    float ds; uint32_t dth;
    drop_params(p, ds, dth);
    keep_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, n, ds, dth, (unsigned long long)seed, seed_dev);
    FCD_LAUNCH_CHECK();
}

// DSA.forward (conv_blocks.py:328-355) + `x + gamma * dsa` (conv_blocks.py:77).
// qkvv: bf16 rows [B*N][ldq] = Linear(LN(t)); EF fp32 [N][P]; temperature/temperature2 fp32 [H]; gamma fp32 [C].
// Saved for backward: inv_n [B][2][C], Ghat/A [B][H][c][c], KV [B][2][C][P], xca [B*N][C], tsa [B][c][H][N].
FCD_API int fcd_dsa_fwd(const void* qkvv, long long ldq, const float* EF, const float* temperature,
                        const float* temperature2, const float* gamma, const void* t, long long ldt, void* y,
                        long long ldy, const float* ca_scale, float sa_drop, long long seed, const long long* seed_dev,
                        float* part, float* inv_n,
                        float* Ghat, float* A, float* Ad, float* KV, float* xca, float* tsa, int B, int N, int C,
                        int Cp, int H, int P, cudaStream_t st) {
    if (C % H || P % 4 || Cp % 8 || C % 8) return -1;
    const int c = C / H;
    const int LPT = Cp / 8;
    if (LPT > 32 || (LPT & (LPT - 1))) return -1;
    const int tn = dsa_tile_tokens(C, P);
    const int ntiles = (N + tn - 1) / tn;
    const int smem_r = tn * (3 * C + P) * 4;
    static int conf_r = 0;
    if (smem_r > conf_r) {
        cudaFuncSetAttribute(dsa_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r > 48 * 1024 ? smem_r : 48 * 1024);
        conf_r = smem_r;
    }
    int zsplit = 1;
    {
        const long long items = (2LL * C / 4) * (P / 4) + ((c % 4) ? (long long)C * c : (long long)C * c / 16) + 2LL * C;
        while ((long long)ntiles * B * zsplit < 2LL * fcd_num_sms() && zsplit * 256LL < items) zsplit *= 2;
    }
    dsa_reduce_kernel<<<dim3(ntiles, B, zsplit), 256, smem_r, st>>>((const bf16*)qkvv, ldq, EF, part, N, C, c, P, tn);
    float ds; uint32_t dth;
    drop_params(sa_drop, ds, dth);
    tile_group_sum(part, B, ntiles, 2LL * C + (long long)C * c + 2LL * C * P, st);
    dsa_finalize_kernel<<<dim3(H, B, finalize_zsplit(c, P)), 256, (c * c + 2 * c) * 4, st>>>(part, ntiles, tile_rows_after_sum(ntiles), temperature,
                                                                     ca_scale, inv_n, Ghat, A, Ad, KV, C, c, P);
    {
        auto run = [&]() -> int {
            DSA_DISPATCH(launch_apply, (const bf16*)qkvv, ldq, inv_n, Ad, KV, temperature2, xca, tsa, B, N, C, H, ds, dth,
                         (unsigned long long)seed, seed_dev, st);
        };
        int rc = run();
        if (rc != 0) return rc;
    }
    const long long rows = (long long)B * N;
    long long grid = (rows * LPT + 255) / 256;
    if (grid > 8LL * fcd_num_sms()) grid = 8LL * fcd_num_sms();
    dsa_combine_kernel<<<(unsigned)grid, 256, 0, st>>>((const bf16*)t, ldt, gamma, xca, tsa, (bf16*)y, ldy, rows, N, C, LPT);
    FCD_LAUNCH_CHECK();
}

// dEF [N][P] = sum over samples and channels of k (x) dKP + v_SA (x) dVP: a PARAMETER gradient that nothing else in
// the backward pass reads, so the host may launch it on a side stream (dKV from fcd_dsa_bwd must stay alive until then).
FCD_API int fcd_dsa_bwd_ef(const void* qkvv, long long ldq, const float* dKV, float* dEF, int B, int N, int C, int P,
                           cudaStream_t st) {
    if (P % 4 || dEF == nullptr) return -1;
    if (N <= 1024) {
        const long long tot = (long long)N * (P / 4) * 8;
        dsa_bwd_ef_small_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const bf16*)qkvv, ldq, dKV, dEF, B, N, C, P);
    } else {
        const long long tot = (long long)((N + 3) / 4) * (P / 4);
        dsa_bwd_ef_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const bf16*)qkvv, ldq, dKV, dEF, B, N, C, P);
    }
    FCD_LAUNCH_CHECK();
}

// Backward of fcd_dsa_fwd w.r.t. qkvv, EF, temperature(2), gamma.  (dt = dy passes straight to fcd_ln_bwd.)
// dtemp / dtemp2 [H], dgamma [C], dEF [N][P] are overwritten (H <= 256).
// work: dqh [B*N][C] fp32, dKV [B][2][C][P], dGhat [B][H][c][c], rqk [B][2][C], gpart 2*148*2*Cp floats.
FCD_API int fcd_dsa_bwd(const void* qkvv, long long ldq, const void* dy, long long lddy, const float* EF,
                        const float* temperature, const float* temperature2, const float* gamma,
                        const float* ca_scale, float sa_drop, long long seed, const long long* seed_dev,
                        const float* inv_n, const float* Ghat,
                        const float* A, const float* Ad, const float* KV, const float* xca, const float* tsa,
                        float* part, float* dqh, float* dKV, float* dGhat, float* rqk, float* gpart, void* dqkvv,
                        long long lddq, float* dEF, float* dtemp, float* dtemp2, float* dgamma, int B, int N, int C,
                        int Cp, int H, int P, cudaStream_t st) {
    if (C % H || P % 4 || Cp % 8 || H > 256 || C / H > 256) return -1;
    const int c = C / H;
    const int LPT = Cp / 8;
    if (LPT > 32 || (LPT & (LPT - 1))) return -1;
    const long long rows = (long long)B * N;
    const int gblk = 2 * fcd_num_sms();
    float ds; uint32_t dth;
    drop_params(sa_drop, ds, dth);
    dsa_dgamma_kernel<<<gblk, 256, 0, st>>>((const bf16*)dy, lddy, xca, tsa, gpart, rows, N, C, LPT, dtemp, dtemp2, H);
    pair_colsum_kernel<<<(C + 7) / 8, 256, 0, st>>>(gpart, gblk, Cp, C, dgamma, nullptr);
    {
        auto run = [&]() -> int {
            DSA_DISPATCH(launch_bwd_reduce, (const bf16*)qkvv, ldq, (const bf16*)dy, lddy, gamma, inv_n, KV, temperature2,
                         dqh, part, B, N, C, H, ds, dth, (unsigned long long)seed, seed_dev, st);
        };
        int rc = run();
        if (rc != 0) return rc;
    }
    const int ntiles = (N + 63) / 64;
    tile_group_sum(part, B * H, ntiles, 2LL * c * P + (long long)c * c + c + 1, st);
    dsa_bwd_finalize_kernel<<<dim3(H, B, finalize_zsplit(c, P)), 256, (c * c + c) * 4, st>>>(part, ntiles, tile_rows_after_sum(ntiles), temperature,
                                                                     Ghat, A, ca_scale, dKV, dGhat, rqk, dtemp, dtemp2, C, c,
                                                                     P);
    {
        auto run = [&]() -> int {
            DSA_DISPATCH(launch_bwd_apply, (const bf16*)qkvv, ldq, (const bf16*)dy, lddy, gamma, EF, inv_n, Ad, dGhat, rqk,
                         dKV, dqh, (bf16*)dqkvv, lddq, B, N, C, H, st);
        };
        int rc = run();
        if (rc != 0) return rc;
    }
    if (dEF == nullptr) FCD_LAUNCH_CHECK();        // the caller runs fcd_dsa_bwd_ef itself (off the critical path)
    return fcd_dsa_bwd_ef(qkvv, ldq, dKV, dEF, B, N, C, P, st);
}
