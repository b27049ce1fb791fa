// Generic implicit-GEMM convolution family on the legacy tensor path (mma.sync bf16, fp32 accumulate).
//
// This is the shape-generic fallback of the conv family: every Conv3d / ConvTranspose3d / Linear of the hot path
// (reference conv_blocks.py:393-437, 640-649; DSA qkvv conv_blocks.py:225) can run through it.  The tcgen05 kernel
// in conv_tc.cu takes over the FLOP-heavy 3x3x3 stride-1 layers; this one keeps odd shapes (stride 2, k2s2
// transposed conv scatter, 1x1, tiny deep levels) correct.
//
//   C[m, n] (+)= bias[n] + sum_t sum_k A[src(m, t), k] * W[t][n][k]
//
// m runs over the "M grid" (B, Dm, Hm, Wm); src(m,t) maps into the source grid (B, Ds, Hs, Ws):
//   mode 0 (forward conv):       s = m*stride + tap - pad
//   mode 1 (data gradient):      q = m + pad - tap, valid iff q % stride == 0, s = q / stride
// out_mode 1 scatters N = 8*Cq columns to the 2x2x2 sub-lattice of a (2Dm,2Hm,2Wm) grid (ConvTranspose3d k2 s2).
#include "common.cuh"
#include "last_block.cuh"

struct IGemmParams {
    const bf16* A;
    long long lda;
    const bf16* W;
    bf16* C;
    long long ldc;
    const float* bias;
    int Bn, Ds, Hs, Ws, Dm, Hm, Wm;
    int K, N;
    int kd, kh, kw, stride, pad, mode, out_mode, accumulate;
    int M, Cq;
    int ksplit;        // > 1: blockIdx.z takes a slice of the (tap, k-chunk) loop and writes fp32 partials to ws
    float* ws;         // [ksplit][M][N] fp32
    unsigned* tickets; // ksplit > 1: one slot per output tile; the CTA that finishes a tile last sums its partials
    // Parity-class mode (data gradient of a stride-2 conv): the GEMM rows are the voxels of ONE parity class of the
    // (Dm, Hm, Wm) grid, (z%2, y%2, x%2) = (par>>2&1, par>>1&1, par&1) for par_on, and only the ntaps taps that can reach
    // that class are multiplied (1, 2, 4 or 8 of the 27) -- walking all 27 taps over mixed-parity rows multiplies 7/8
    // zero rows (measured: 4.4 of SegResNet's 18.6 ms step).
    int par_on, par, ntaps, taps[8];
};

namespace {

constexpr int BM = 128;
constexpr int STAGES = 3;

template <int BK>
__device__ __forceinline__ int swz_off(int row, int chunk) {
    // byte offset of 16B chunk `chunk` of row `row` in a [rows][BK] bf16 tile, XOR-swizzled for ldmatrix
    if (BK == 32) return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
    return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4);
}

template <int BN, int WN, int BK>
__global__ void __launch_bounds__(128 * (BN / WN)) igemm_kernel(const IGemmParams p) {
    constexpr int WARPS_N = BN / WN;
    constexpr int NT = 128 * WARPS_N;
    constexpr int CPR = BK / 8;                 // 16B chunks per tile row
    constexpr int A_CHUNKS = BM * CPR;
    constexpr int A_PER_T = A_CHUNKS / NT;
    constexpr int B_CHUNKS = BN * CPR;
    constexpr int A_BYTES = BM * BK * 2;
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(A_CHUNKS % NT == 0, "A tile must divide evenly");

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % 4, wn = warp / 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // ---- per-thread gather rows (fixed across the K loop)
    int r_z[A_PER_T], r_y[A_PER_T], r_x[A_PER_T];
    long long r_base[A_PER_T];
    bool r_ok[A_PER_T];
    int r_row[A_PER_T], r_ch[A_PER_T];
#pragma unroll
    for (int i = 0; i < A_PER_T; ++i) {
        int idx = tid + i * NT;
        int row = idx / CPR;
        r_row[i] = row;
        r_ch[i] = idx % CPR;
        int m = m0 + row;
        r_ok[i] = m < p.M;
        int mm = r_ok[i] ? m : 0;
        int x, y, z;
        if (p.par_on) {
            const int W2 = p.Wm >> 1, H2 = p.Hm >> 1, D2 = p.Dm >> 1;
            x = 2 * (mm % W2) + (p.par & 1); mm /= W2;
            y = 2 * (mm % H2) + ((p.par >> 1) & 1); mm /= H2;
            z = 2 * (mm % D2) + ((p.par >> 2) & 1); mm /= D2;
        } else {
            x = mm % p.Wm; mm /= p.Wm;
            y = mm % p.Hm; mm /= p.Hm;
            z = mm % p.Dm; mm /= p.Dm;
        }
        r_x[i] = x; r_y[i] = y; r_z[i] = z;
        r_base[i] = (long long)mm * p.Ds * p.Hs * p.Ws;
    }

    const int kchunks = p.K / BK;
    const int T = p.par_on ? p.ntaps : p.kd * p.kh * p.kw;
    const int nk_all = T * kchunks;
    // split-K: this CTA runs iterations [it0, it0 + nk) of the flattened (tap, k-chunk) loop
    const int it0 = (int)((long long)nk_all * blockIdx.z / p.ksplit);
    const int nk = (int)((long long)nk_all * (blockIdx.z + 1) / p.ksplit) - it0;

    auto load_stage = [&](int it_local, int slot) {
        const int it = it0 + it_local;
        const int tl = it / kchunks, kc = it - tl * kchunks;
        const int t = p.par_on ? p.taps[tl] : tl;
        const int tx = t % p.kw, ty = (t / p.kw) % p.kh, tz = t / (p.kw * p.kh);
        const uint32_t sa = smem_base + slot * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            int sz, sy, sx;
            bool ok = r_ok[i];
            if (p.mode == 0) {
                sz = r_z[i] * p.stride + tz - p.pad;
                sy = r_y[i] * p.stride + ty - p.pad;
                sx = r_x[i] * p.stride + tx - p.pad;
            } else {
                int qz = r_z[i] + p.pad - tz, qy = r_y[i] + p.pad - ty, qx = r_x[i] + p.pad - tx;
                ok = ok && qz >= 0 && qy >= 0 && qx >= 0;
                if (p.stride > 1) {
                    ok = ok && (qz % p.stride == 0) && (qy % p.stride == 0) && (qx % p.stride == 0);
                    sz = qz / p.stride; sy = qy / p.stride; sx = qx / p.stride;
                } else {
                    sz = qz; sy = qy; sx = qx;
                }
            }
            ok = ok && sz >= 0 && sz < p.Ds && sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws;
            const bf16* src = p.A;
            if (ok) src += (r_base[i] + ((long long)sz * p.Hs + sy) * p.Ws + sx) * p.lda + kc * BK + r_ch[i] * 8;
            cp_async16(sa + swz_off<BK>(r_row[i], r_ch[i]), src, ok);
        }
        for (int idx = tid; idx < B_CHUNKS; idx += NT) {
            int n = idx / CPR, ch = idx % CPR;
            bool ok = (n0 + n) < p.N;
            const bf16* src = p.W;
            if (ok) src += ((long long)t * p.N + n0 + n) * p.K + kc * BK + ch * 8;
            cp_async16(sb + swz_off<BK>(n, ch), src, ok);
        }
    };

    float acc[2][WN / 8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < WN / 8; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = it + STAGES - 1;
            if (nx < nk) load_stage(nx, nx % STAGES);
            cp_async_commit();
        }
        const uint32_t sa = smem_base + (it % STAGES) * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
            uint32_t af[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                int row = wm * 32 + mi * 16 + (lane & 15);
                int ch = kk * 2 + (lane >> 4);
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], sa + swz_off<BK>(row, ch));
            }
#pragma unroll
            for (int nj = 0; nj < WN / 16; ++nj) {
                uint32_t b0, b1, b2, b3;
                int n = wn * WN + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(b0, b1, b2, b3, sb + swz_off<BK>(n, ch));
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_bf16_16816(acc[mi][nj * 2], af[mi], b0, b1);
                    mma_bf16_16816(acc[mi][nj * 2 + 1], af[mi], b2, b3);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    const int g = lane >> 2, tq = lane & 3;
    if (p.ksplit > 1) {
        // ---- split-K epilogue: raw fp32 partial tile -> ws[z][m][n]; igemm_splitk_reduce_kernel finishes
        float* w = p.ws + (long long)blockIdx.z * p.M * p.N;
#pragma unroll
        for (int nj = 0; nj < WN / 8; ++nj) {
            const int col = n0 + wn * WN + nj * 8 + tq * 2;
            if (col >= p.N) continue;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int row = m0 + wm * 32 + mi * 16 + g;
                if (row < p.M) *reinterpret_cast<float2*>(&w[(long long)row * p.N + col]) = make_float2(acc[mi][nj][0], acc[mi][nj][1]);
                if (row + 8 < p.M) *reinterpret_cast<float2*>(&w[(long long)(row + 8) * p.N + col]) = make_float2(acc[mi][nj][2], acc[mi][nj][3]);
            }
        }
        if (p.tickets == nullptr) return;          // partials only: fcd_splitk_reduce finishes
        // The CTA that writes the LAST partial of this output tile sums the ksplit partials (fixed order z = 0, 1, ...:
        // deterministic whatever the arrival order), adds the bias and writes the bf16 rows -- no reduce launch.
        if (!lastblk::arrive(p.tickets + (blockIdx.y * gridDim.x + blockIdx.x), (unsigned)p.ksplit)) return;
        constexpr int T_CHUNKS = BM * (BN / 8);
        for (int idx = tid; idx < T_CHUNKS; idx += NT) {
            const int row = idx / (BN / 8), c8 = idx % (BN / 8);
            const int m = m0 + row, n = n0 + c8 * 8;
            if (m >= p.M || n >= p.N) continue;
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = p.bias ? p.bias[n + k] : 0.f;
            for (int z = 0; z < p.ksplit; ++z) {
                const float4* src = reinterpret_cast<const float4*>(p.ws + ((long long)z * p.M + m) * p.N + n);
                const float4 u = __ldcg(src), v = __ldcg(src + 1);
                a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += v.x; a[5] += v.y; a[6] += v.z; a[7] += v.w;
            }
            bf16* dst = p.C + (long long)m * p.ldc + n;
            if (p.accumulate) {
                float b[8];
                unpack8(ld8(dst), b);
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] += b[k];
            }
            st8(dst, pack8(a));
        }
        return;
    }
    // ---- epilogue: (+bias) -> bf16 tile in smem -> 16B coalesced rows (plain / accumulate / k2s2 scatter)
    constexpr int CLD = BN + 8;   // padded row (elements)
    bf16* ctile = reinterpret_cast<bf16*>(smem);
#pragma unroll
    for (int nj = 0; nj < WN / 8; ++nj) {
        int col = wn * WN + nj * 8 + tq * 2;
        float bv0 = 0.f, bv1 = 0.f;
        if (p.bias != nullptr && (n0 + col) < p.N) {
            int nb = n0 + col;
            if (p.out_mode == 1) nb = nb % p.Cq;
            bv0 = p.bias[nb];
            bv1 = p.bias[nb + 1];
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            int row = wm * 32 + mi * 16 + g;
            *reinterpret_cast<__nv_bfloat162*>(&ctile[row * CLD + col]) =
                __floats2bfloat162_rn(acc[mi][nj][0] + bv0, acc[mi][nj][1] + bv1);
            *reinterpret_cast<__nv_bfloat162*>(&ctile[(row + 8) * CLD + col]) =
                __floats2bfloat162_rn(acc[mi][nj][2] + bv0, acc[mi][nj][3] + bv1);
        }
    }
    __syncthreads();
    constexpr int C_CHUNKS = BM * (BN / 8);
    for (int idx = tid; idx < C_CHUNKS; idx += NT) {
        int row = idx / (BN / 8), c8 = idx % (BN / 8);
        int m = m0 + row, n = n0 + c8 * 8;
        if (m >= p.M || n >= p.N) continue;
        bf16x8 v = *reinterpret_cast<const bf16x8*>(&ctile[row * CLD + c8 * 8]);
        bf16* dst;
        if (p.out_mode == 0 && p.par_on) {
            int mm = m;
            const int W2 = p.Wm >> 1, H2 = p.Hm >> 1, D2 = p.Dm >> 1;
            const int x = 2 * (mm % W2) + (p.par & 1); mm /= W2;
            const int y = 2 * (mm % H2) + ((p.par >> 1) & 1); mm /= H2;
            const int z = 2 * (mm % D2) + ((p.par >> 2) & 1); mm /= D2;
            dst = p.C + ((((long long)mm * p.Dm + z) * p.Hm + y) * p.Wm + x) * p.ldc + n;
        } else if (p.out_mode == 0) {
            dst = p.C + (long long)m * p.ldc + n;
        } else {
            int tap = n / p.Cq, co = n - tap * p.Cq;
            int mm = m;
            int x = mm % p.Wm; mm /= p.Wm;
            int y = mm % p.Hm; mm /= p.Hm;
            int z = mm % p.Dm; mm /= p.Dm;
            int oz = 2 * z + (tap >> 2), oy = 2 * y + ((tap >> 1) & 1), ox = 2 * x + (tap & 1);
            long long vox = (((long long)mm * (2 * p.Dm) + oz) * (2 * p.Hm) + oy) * (2 * p.Wm) + ox;
            dst = p.C + vox * p.ldc + co;
        }
        if (p.accumulate) {
            float a[8], b[8];
            unpack8(v, a);
            unpack8(ld8(dst), b);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += b[i];
            v = pack8(a);
        }
        st8(dst, v);
    }
}

template <int BN, int WN, int BK>
int launch_igemm(const IGemmParams& p, cudaStream_t stream) {
    constexpr int NT = 128 * (BN / WN);
    constexpr int pipe = STAGES * (BM * BK * 2 + BN * BK * 2);
    constexpr int epi = BM * (BN + 8) * 2;
    constexpr int smem = pipe > epi ? pipe : epi;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(igemm_kernel<BN, WN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.ksplit);
    igemm_kernel<BN, WN, BK><<<grid, NT, smem, stream>>>(p);
    return (int)cudaGetLastError();
}

// C[m][n] (+)= bias[n] + sum_z ws[z][m][n]  -> bf16 rows (fixed summation order: deterministic)
__global__ void igemm_splitk_reduce_kernel(const float* __restrict__ ws, bf16* __restrict__ C, long long ldc,
                                           const float* __restrict__ bias, int M, int N, int ksplit, int accumulate) {
    const int N8 = N / 8;
    const long long total = (long long)M * N8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / N8;
        const int n = (int)(i % N8) * 8;
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = bias ? bias[n + k] : 0.f;
        for (int z = 0; z < ksplit; ++z) {
            const float4* src = reinterpret_cast<const float4*>(ws + ((long long)z * M + m) * N + n);
            const float4 u = src[0], v = src[1];
            a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += v.x; a[5] += v.y; a[6] += v.z; a[7] += v.w;
        }
        bf16* dst = C + m * ldc + n;
        if (accumulate) {
            float b[8];
            unpack8(ld8(dst), b);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += b[k];
        }
        st8(dst, pack8(a));
    }
}

}  // namespace

// Replaces torch.nn.functional.conv3d / conv_transpose3d / linear as dispatched by the reference's nn.Conv3d,
// nn.ConvTranspose3d and nn.Linear modules (conv_blocks.py:393-437, 640-649, 225; ms_dsa_net.py:216,362).
// Returns a cudaError_t value (0 = ok), -1 for unsupported shapes.
FCD_API int fcd_igemm(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias,
                      int Bn, int Ds, int Hs, int Ws, int Dm, int Hm, int Wm, int K, int N, int kd, int kh, int kw,
                      int stride, int pad, int mode, int out_mode, int accumulate, int Cq, cudaStream_t stream) {
    if (K % 16 != 0 || N % 8 != 0 || lda % 8 != 0 || ldc % 8 != 0) return -1;
    if (out_mode == 1 && (Cq % 8 != 0 || N != 8 * Cq)) return -1;
    IGemmParams p;
    p.A = (const bf16*)A; p.lda = lda; p.W = (const bf16*)W; p.C = (bf16*)C; p.ldc = ldc; p.bias = bias;
    p.Bn = Bn; p.Ds = Ds; p.Hs = Hs; p.Ws = Ws; p.Dm = Dm; p.Hm = Hm; p.Wm = Wm;
    p.K = K; p.N = N; p.kd = kd; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad;
    p.mode = mode; p.out_mode = out_mode; p.accumulate = accumulate; p.Cq = Cq > 0 ? Cq : 1;
    p.ksplit = 1; p.ws = nullptr; p.tickets = nullptr; p.par_on = 0; p.par = 0; p.ntaps = 0;
    long long M = (long long)Bn * Dm * Hm * Wm;
    if (M <= 0 || M > 0x7fffffffLL) return -1;
    p.M = (int)M;
    const bool k32 = (K % 32 == 0);
    if (N <= 16) return k32 ? launch_igemm<16, 16, 32>(p, stream) : launch_igemm<16, 16, 16>(p, stream);
    if (N <= 32) return k32 ? launch_igemm<32, 32, 32>(p, stream) : launch_igemm<32, 32, 16>(p, stream);
    if (N <= 64 || (long long)((M + BM - 1) / BM) * ((N + 127) / 128) < 2 * fcd_num_sms())
        return k32 ? launch_igemm<64, 32, 32>(p, stream) : launch_igemm<64, 32, 16>(p, stream);
    return k32 ? launch_igemm<128, 64, 32>(p, stream) : launch_igemm<128, 64, 16>(p, stream);
}

// Split-K variant for the deep, spatially tiny levels (e.g. 512->512 on 4^3 x batch 2: M = 128 rows but K = 13824):
// the plain kernel would run the whole (tap, k-chunk) loop in a handful of CTAs.  fcd_igemm_ksplit proposes the
// split (1 = use fcd_igemm); ws holds ksplit*M*N floats.  out_mode 0 only.
FCD_API int fcd_igemm_ksplit(long long M, int N, int K, int T) {
    if (M <= 0 || N % 8 || K % 16) return 1;
    const int bk = (K % 32 == 0) ? 32 : 16;
    const int nk = T * (K / bk);
    const long long tiles = ((M + BM - 1) / BM) * ((N + 63) / 64);
    if (tiles >= fcd_num_sms() || nk < 16) return 1;
    long long ks = (2LL * fcd_num_sms() + tiles - 1) / tiles;
    if (ks > nk / 4) ks = nk / 4;
    if (ks > 64) ks = 64;
    while (ks > 1 && ks * M * N * 4 > (256LL << 20)) --ks;
    return ks < 2 ? 1 : (int)ks;
}

FCD_API int fcd_igemm_splitk(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias,
                             int Bn, int Ds, int Hs, int Ws, int Dm, int Hm, int Wm, int K, int N, int kd, int kh,
                             int kw, int stride, int pad, int mode, int accumulate, float* ws, int ksplit,
                             cudaStream_t stream) {
    if (K % 16 != 0 || N % 8 != 0 || lda % 8 != 0 || ldc % 8 != 0 || ksplit < 2 || ws == nullptr) return -1;
    IGemmParams p;
    p.A = (const bf16*)A; p.lda = lda; p.W = (const bf16*)W; p.C = (bf16*)C; p.ldc = ldc; p.bias = nullptr;
    p.Bn = Bn; p.Ds = Ds; p.Hs = Hs; p.Ws = Ws; p.Dm = Dm; p.Hm = Hm; p.Wm = Wm;
    p.K = K; p.N = N; p.kd = kd; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad;
    p.mode = mode; p.out_mode = 0; p.accumulate = accumulate; p.Cq = 1;
    p.ksplit = ksplit; p.ws = ws; p.par_on = 0; p.par = 0; p.ntaps = 0;
    long long M = (long long)Bn * Dm * Hm * Wm;
    if (M <= 0 || M > 0x7fffffffLL) return -1;
    p.M = (int)M;
    // in-kernel reduction: one ticket per 64 x 64 output tile (falls back to the reduce launch if the ring is too small)
    // (only for ksplit <= 2: one CTA summing many partial tiles is a serial tail -- measured 1.4x slower kernels at the
    // deep levels' ksplit of 8-27 -- while the reduce launch spreads the same reads over all SMs)
    const unsigned ntiles = (unsigned)(((M + BM - 1) / BM) * ((N + 63) / 64));
    p.tickets = ksplit <= 2 ? lastblk::next_tickets(ntiles) : nullptr;
    p.bias = p.tickets != nullptr ? bias : nullptr;
    int rc = (K % 32 == 0) ? launch_igemm<64, 32, 32>(p, stream) : launch_igemm<64, 32, 16>(p, stream);
    if (rc != 0 || p.tickets != nullptr) return rc;
    const long long total = M * (N / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 8 * fcd_num_sms()) blocks = 8 * fcd_num_sms();
    igemm_splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(ws, (bf16*)C, ldc, bias, (int)M, N, ksplit, accumulate);
    return (int)cudaGetLastError();
}

// C[m][n] = bias[n] + sum_z ws[z][m][n] -> bf16 rows: the second half of every split-K conv (fcd_igemm_splitk calls it
// itself; fcd_conv_gemm_tc leaves it to the caller).
FCD_API int fcd_splitk_reduce(const float* ws, void* C, long long ldc, const float* bias, long long M, int N, int ksplit,
                              int accumulate, cudaStream_t stream) {
    if (N % 8 || ldc % 8 || ksplit < 1 || M < 1 || M > 0x7fffffffLL) return -1;
    const long long total = M * (N / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 8 * fcd_num_sms()) blocks = 8 * fcd_num_sms();
    igemm_splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(ws, (bf16*)C, ldc, bias, (int)M, N, ksplit, accumulate);
    return (int)cudaGetLastError();
}

// Data gradient of a 3x3x3 stride-2 pad-1 conv (MONAI SegResNet down-sampling, segresnet_dsa.py:97): dX[Bn][Dm][Hm][Wm][N]
// from dY[Bn][Ds][Hs][Ws][K] (Ds = Dm / 2 ...) and the packed transposed weights W[27][N][K] (as fcd_igemm mode 1), as
// EIGHT launches, one per parity class of the dX grid, each multiplying only the taps that reach its class.
FCD_API int fcd_igemm_dgrad_s2(const void* A, long long lda, const void* W, void* C, long long ldc, int Bn, int Ds, int Hs,
                               int Ws, int Dm, int Hm, int Wm, int K, int N, cudaStream_t stream) {
    if (K % 16 != 0 || N % 8 != 0 || lda % 8 != 0 || ldc % 8 != 0) return -1;
    if ((Dm | Hm | Wm) & 1 || Ds * 2 != Dm || Hs * 2 != Hm || Ws * 2 != Wm) return -1;
    IGemmParams p;
    p.A = (const bf16*)A; p.lda = lda; p.W = (const bf16*)W; p.C = (bf16*)C; p.ldc = ldc; p.bias = nullptr;
    p.Bn = Bn; p.Ds = Ds; p.Hs = Hs; p.Ws = Ws; p.Dm = Dm; p.Hm = Hm; p.Wm = Wm;
    p.K = K; p.N = N; p.kd = 3; p.kh = 3; p.kw = 3; p.stride = 2; p.pad = 1;
    p.mode = 1; p.out_mode = 0; p.accumulate = 0; p.Cq = 1;
    p.ksplit = 1; p.ws = nullptr; p.tickets = nullptr; p.par_on = 1;
    const long long M = (long long)Bn * (Dm / 2) * (Hm / 2) * (Wm / 2);
    if (M <= 0 || M > 0x7fffffffLL) return -1;
    p.M = (int)M;
    for (int cls = 0; cls < 8; ++cls) {
        p.par = cls;
        // voxel index i, tap t, source q = (i + 1 - t) / 2: t must have the parity of i + 1 -> even i: t = 1; odd i: t in {0, 2}
        int tz[2], ty[2], tx[2];
        const int nz = ((cls >> 2) & 1) ? 2 : 1, ny = ((cls >> 1) & 1) ? 2 : 1, nx = (cls & 1) ? 2 : 1;
        tz[0] = nz == 2 ? 0 : 1; tz[1] = 2;
        ty[0] = ny == 2 ? 0 : 1; ty[1] = 2;
        tx[0] = nx == 2 ? 0 : 1; tx[1] = 2;
        p.ntaps = 0;
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b < ny; ++b)
                for (int c = 0; c < nx; ++c) p.taps[p.ntaps++] = (tz[a] * 3 + ty[b]) * 3 + tx[c];
        int rc;
        const bool k32 = (K % 32 == 0);
        if (N <= 16) rc = k32 ? launch_igemm<16, 16, 32>(p, stream) : launch_igemm<16, 16, 16>(p, stream);
        else if (N <= 32) rc = k32 ? launch_igemm<32, 32, 32>(p, stream) : launch_igemm<32, 32, 16>(p, stream);
        else rc = k32 ? launch_igemm<64, 32, 32>(p, stream) : launch_igemm<64, 32, 16>(p, stream);
        if (rc != 0) return rc;
    }
    return 0;
}
