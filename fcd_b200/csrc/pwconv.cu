// Pointwise (1x1x1, stride 1, no bias) convolution on large volumes with few channels (<= 32 in, <= 32 out): the
// residual conv3 of UnetResBlock on the two top levels (conv_blocks.py:420-424, 447-449) and its data gradient.
// These layers move 400 MB for 4 GFLOP: the implicit-GEMM kernel (one 128-row tile and ONE k-step per CTA, generic
// per-tap address math) ran them at 0.4 of the HBM roofline.  Here a thread owns whole rows: 16-byte loads of the
// row, fp32 FMAs against the weight matrix held in shared memory (float4 broadcast reads, two rows per thread share
// each read), 16-byte bf16 stores.
#include "common.cuh"

namespace {

struct PwParams {
    const bf16* A; long long lda;
    const float* Wf; long long sn, sk;
    int Nr, Kr, kseg, ksegpad, nsg, nsgpad;
    bf16* C; long long ldc;
    long long M;
};

template <int K, int N>
__global__ void __launch_bounds__(256) pw_conv_kernel(const PwParams p) {
    __shared__ __align__(16) float sW[K * N];                  // [k][n]
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int kp = i / N, np_ = i % N;
        const int kseg_i = kp / p.ksegpad, kw = kp % p.ksegpad;
        const int k = kseg_i * p.kseg + kw;
        const int nseg_i = np_ / p.nsgpad, nw = np_ % p.nsgpad;
        const int n = nseg_i * p.nsg + nw;
        float v = 0.f;
        if (kw < p.kseg && k < p.Kr && nw < p.nsg && n < p.Nr) v = p.Wf[n * p.sn + k * p.sk];
        sW[i] = v;
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two rows per thread per iteration: rows r and r + stride
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < p.M; r += 2 * stride) {
        const long long r2 = r + stride;
        const bool two = r2 < p.M;
        float x0[K], x1[K];
#pragma unroll
        for (int k = 0; k < K; k += 8) {
            unpack8(ld8_stream(p.A + r * p.lda + k), x0 + k);
            if (two) unpack8(ld8_stream(p.A + r2 * p.lda + k), x1 + k);
            else {
#pragma unroll
                for (int u = 0; u < 8; ++u) x1[k + u] = 0.f;
            }
        }
        float a0[N], a1[N];
#pragma unroll
        for (int n = 0; n < N; ++n) { a0[n] = 0.f; a1[n] = 0.f; }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int n = 0; n < N; n += 4) {
                const float4 w = *reinterpret_cast<const float4*>(&sW[k * N + n]);
                a0[n] = fmaf(x0[k], w.x, a0[n]); a0[n + 1] = fmaf(x0[k], w.y, a0[n + 1]);
                a0[n + 2] = fmaf(x0[k], w.z, a0[n + 2]); a0[n + 3] = fmaf(x0[k], w.w, a0[n + 3]);
                a1[n] = fmaf(x1[k], w.x, a1[n]); a1[n + 1] = fmaf(x1[k], w.y, a1[n + 1]);
                a1[n + 2] = fmaf(x1[k], w.z, a1[n + 2]); a1[n + 3] = fmaf(x1[k], w.w, a1[n + 3]);
            }
        }
#pragma unroll
        for (int n = 0; n < N; n += 8) {
            st8(p.C + r * p.ldc + n, pack8(a0 + n));
            if (two) st8(p.C + r2 * p.ldc + n, pack8(a1 + n));
        }
    }
}

template <int K, int N>
int launch_pw(const PwParams& p, cudaStream_t st) {
    long long blocks = (p.M + 511) / 512;                       // 2 rows per thread
    const long long cap = 8LL * fcd_num_sms();
    if (blocks > cap) blocks = cap;
    pw_conv_kernel<K, N><<<(unsigned)blocks, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace

// 1 if fcd_pw_conv takes a 1x1x1 conv with K (padded) input and N (padded) output channels on M voxels.
FCD_API int fcd_pw_conv_ok(long long M, int K, int N) {
    return (M >= 65536 && (K == 16 || K == 32) && (N == 16 || N == 32)) ? 1 : 0;
}

// C[v][n] = sum_k A[v][k] * W(n, k) for v < M.  A: bf16 rows of pitch lda (>= K), C: bf16 rows of pitch ldc (>= N);
// K, N: padded channel counts in {16, 32}.  W(n, k) = Wf[n*sn + k*sk] for real n < Nr, k < Kr with the concat-segment
// maps of fcd_pack_weight (padded k -> (k / ksegpad) * kseg + k % ksegpad, same for n); everything else is zero.
// Forward: sn = Cin, sk = 1.  Data gradient: A = dY, W transposed (sn = 1, sk = Cin), output channels = Cin.
FCD_API int fcd_pw_conv(const void* A, long long lda, const float* Wf, long long sn, long long sk, int Nr, int Kr,
                        int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, long long M, int K, int N,
                        cudaStream_t stream) {
    if (!fcd_pw_conv_ok(M, K, N) || lda % 8 || ldc % 8 || lda < K || ldc < N) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)C & 15) || kseg < 1 || ksegpad < 1 || nsg < 1 || nsgpad < 1) return -1;
    PwParams p;
    p.A = (const bf16*)A; p.lda = lda; p.Wf = Wf; p.sn = sn; p.sk = sk; p.Nr = Nr; p.Kr = Kr;
    p.kseg = kseg; p.ksegpad = ksegpad; p.nsg = nsg; p.nsgpad = nsgpad; p.C = (bf16*)C; p.ldc = ldc; p.M = M;
    if (K == 16 && N == 16) return launch_pw<16, 16>(p, stream);
    if (K == 32 && N == 16) return launch_pw<32, 16>(p, stream);
    if (K == 16 && N == 32) return launch_pw<16, 32>(p, stream);
    return launch_pw<32, 32>(p, stream);
}
