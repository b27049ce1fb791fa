// Weight gradient of the FIRST conv of every network (conv_blocks.py:393-416 at encoder1.conv1 / the 1x1x1 residual
// conv3; MONAI SegResNet convInit, segresnet_dsa.py:82): 2 real input channels (T1 + FLAIR, config.py:9) in a 16-channel
// row, 16 output channels, the full-resolution volume.
//
//   dW[n][ci][tap] = sum_v dY[v][n] * X[v + tap][ci]
//
// The generic tcgen05 kernel treats it as a 16 -> 16 conv: 27 taps x (3 kd x 16 k) rows of which 1/8 are real -- 280 us at
// the very END of the backward pass, where nothing is left to overlap it.  Here the 27 x Ci real
// (tap, ci) pairs are the N dimension of ONE mma.sync contraction over the voxels:
//
//   D[n][(tap, ci)] += A[n][voxel] * B[voxel][(tap, ci)],   16 voxels (consecutive along W) per mma.m16n8k16
//
// A comes from the dY tile ([voxel][16 ch] rows, ldmatrix.trans); a B fragment register is the pair (voxel 2t, voxel 2t+1)
// of one (tap, ci) column = two ADJACENT bf16 of a channel-separated shared-memory halo copy of X that keeps only the real
// channels.  Each channel plane is stored twice, the second copy shifted by one voxel, so that the pair is one aligned
// 32-bit word for even and for odd kw (the first version gathered four 2-byte loads per register from interleaved
// channels: 3-way bank conflicts, 175 us).  54 columns = 7 n-tiles, 28 accumulator registers per thread for the CTA's
// whole life; the kernel is bound by reading dY and X once (268 MB), not by the tensor or issue rate.
//
// CTA = (sample, 8-row y tile, d segment), marching over its planes with the X planes in a 4-slot ring and the dY tile
// double-buffered (cp.async, zero-fill outside the volume = the conv's padding); the CTA's partial goes to
// part[cta][tap][16][Kp] (only ci < Ci written -- fcd_wgrad_reduce reads nothing else) and fcd_wgrad_reduce sums the CTAs
// in a fixed order.
#include "common.cuh"

namespace {

constexpr int ROWS = 8;          // y rows per CTA
constexpr int NWARPS = 8;

struct SmallCParams {
    const bf16* X; long long ldx;
    const bf16* dY; long long ldy;
    float* part;
    int Bn, D, H, W, Ci, taps, halo, nyt, nseg, DL, Kp;
};

__device__ __forceinline__ int dyswz(int row, int chunk) { return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4); }

template <int NT>
__global__ void __launch_bounds__(32 * NWARPS, 2) wgrad_smallc_kernel(const SmallCParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = p.W, halo = p.halo, XW = W + 2 * halo, XR = ROWS + 2 * halo;
    const int XWp = XW + 2;                            // row pitch (elements) of a channel plane copy: even, room for the shift
    const int dy_bytes = ROWS * W * 32, cplane = XR * XWp * 2, xslot_bytes = 4 * cplane;
    unsigned char* sdy = smem;                         // 2 x [ROWS*W voxels][16 ch] bf16, chunk-swizzled
    unsigned char* sx = smem + 2 * dy_bytes;           // 4 slots x [ci 0..1][copy 0..1][XR][XWp] bf16; copy 1 shifted by +1
    const uint32_t sdy_u = smem_u32(sdy);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int item = blockIdx.x;
    const int seg = item % p.nseg; item /= p.nseg;
    const int yt = item % p.nyt;
    const int b = item / p.nyt;
    const int y0 = yt * ROWS, z0 = seg * p.DL, z1 = min(p.D, z0 + p.DL);
    const long long plane_vox = (long long)p.H * W;
    const bf16* Xb = p.X + (long long)b * p.D * plane_vox * p.ldx;
    const bf16* Yb = p.dY + (long long)b * p.D * plane_vox * p.ldy;

    // X plane z: the two real channels of every voxel of the halo'd rows as ONE 32-bit load each (zero outside the volume),
    // held in registers across the compute of the current plane, then scattered into the ring slot ((z - (z0 - halo)) & 3)
    constexpr int NPF = 11;                            // ceil(10 * 258 / 256): W <= 256
    const int nx = XR * XW;
    uint32_t pf[NPF];
    auto fetch_x = [&](int z) {
        const bool zok = z >= 0 && z < p.D;
#pragma unroll
        for (int u = 0; u < NPF; ++u) {
            const int i = tid + u * 32 * NWARPS;
            pf[u] = 0u;
            if (i < nx) {
                const int r = i / XW, c = i - r * XW;
                const int y = y0 + r - halo, x = c - halo;
                if (zok && y >= 0 && y < p.H && x >= 0 && x < W)
                    pf[u] = __ldg(reinterpret_cast<const uint32_t*>(Xb + ((long long)z * plane_vox + (long long)y * W + x) * p.ldx));
            }
        }
    };
    auto store_x = [&](int z) {
        unsigned char* slot = sx + (size_t)((z - (z0 - halo)) & 3) * xslot_bytes;
#pragma unroll
        for (int u = 0; u < NPF; ++u) {
            const int i = tid + u * 32 * NWARPS;
            if (i < nx) {
                const int r = i / XW, c = i - r * XW;
                const unsigned short c0 = (unsigned short)(pf[u] & 0xffffu), c1 = (unsigned short)(pf[u] >> 16);
                unsigned short* q = reinterpret_cast<unsigned short*>(slot) + r * XWp + c;
                q[0] = c0;                             // channel 0, copy 0
                q[XR * XWp + 1] = c0;                  // channel 0, copy 1 (shifted)
                q[2 * XR * XWp] = c1;                  // channel 1, copy 0
                q[3 * XR * XWp + 1] = c1;              // channel 1, copy 1
            }
        }
    };
    auto load_dy = [&](int z, int slot) {
        const uint32_t dst0 = sdy_u + (uint32_t)slot * dy_bytes;
        for (int i = tid; i < ROWS * W * 2; i += 32 * NWARPS) {
            const int v = i >> 1, ch = i & 1;
            const int r = v / W, x = v - r * W;
            const int y = y0 + r;
            const bool ok = y < p.H;
            const bf16* src = ok ? Yb + ((long long)z * plane_vox + (long long)y * W + x) * p.ldy + ch * 8 : Yb;
            cp_async16(dst0 + dyswz(v, ch), src, ok);
        }
    };

    // per-lane column constants: column c = 8 j + g  ->  (tap, ci) -> offsets inside an X plane
    const int g = lane >> 2, tq = lane & 3;
    const int ncol = p.taps * p.Ci;
    int coff[NT], cdz[NT];
    bool cok[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int c = 8 * j + g;
        cok[j] = c < ncol;
        const int tap = cok[j] ? c / p.Ci : 0, ci = cok[j] ? c % p.Ci : 0;
        const int dz = p.taps == 27 ? tap / 9 : 0, dy = p.taps == 27 ? (tap / 3) % 3 : 0, dx = p.taps == 27 ? tap % 3 : 0;
        cdz[j] = dz;
        const int copy = dx & 1;                       // odd kw: the shifted copy makes (x, x + 1) an aligned word
        coff[j] = (((ci * 2 + copy) * XR + dy) * XWp + dx + copy) * 2;
    }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;

    // prologue: X planes z0-halo .. z0+halo, dY plane z0
    load_dy(z0, 0);
    cp_async_commit();
    for (int z = z0 - halo; z <= z0 + halo; ++z) { fetch_x(z); store_x(z); }
    const int ngroups = ROWS * (W / 16), gpr = W / 16;
    for (int z = z0; z < z1; ++z) {
        const int slot = (z - z0) & 1;
        const bool more = z + 1 < z1;
        if (more) {
            fetch_x(z + 1 + halo);                     // lands while this plane is multiplied
            load_dy(z + 1, slot ^ 1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t dyb = sdy_u + (uint32_t)slot * dy_bytes;
        // ring slot of plane (z - halo + dz): ((z - z0) + dz) & 3
        const unsigned char* xp[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) xp[d] = sx + (size_t)(((z - z0) + d) & 3) * xslot_bytes;
        for (int gi = warp; gi < ngroups; gi += NWARPS) {
            const int r = gi / gpr, x0 = (gi - r * gpr) * 16;
            uint32_t a[4];
            {
                const int v = r * W + x0 + (lane & 7) + ((lane >> 4) << 3);
                const int ch = (lane >> 3) & 1;
                ldmatrix_x4_trans(a[0], a[1], a[2], a[3], dyb + dyswz(v, ch));
            }
            const int vbase = (r * XWp + x0 + 2 * tq) * 2;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                uint32_t b0 = 0u, b1 = 0u;
                if (cok[j]) {
                    const unsigned char* q = (cdz[j] == 0 ? xp[0] : (cdz[j] == 1 ? xp[1] : xp[2])) + vbase + coff[j];
                    b0 = *reinterpret_cast<const uint32_t*>(q);            // voxels 2t, 2t+1
                    b1 = *reinterpret_cast<const uint32_t*>(q + 16);       // voxels 2t+8, 2t+9
                }
                mma_bf16_16816(acc[j], a, b0, b1);
            }
        }
        if (more) store_x(z + 1 + halo);               // into the slot no plane of this iteration reads
        __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();

    // cross-warp reduction through shared memory: red[warp][16 n][8 NT cols]
    constexpr int NC = 8 * NT;
    float* red = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int col = 8 * j + 2 * tq;
        red[(warp * 16 + g) * NC + col] = acc[j][0];
        red[(warp * 16 + g) * NC + col + 1] = acc[j][1];
        red[(warp * 16 + g + 8) * NC + col] = acc[j][2];
        red[(warp * 16 + g + 8) * NC + col + 1] = acc[j][3];
    }
    __syncthreads();
    float* out = p.part + (long long)blockIdx.x * p.taps * 16 * p.Kp;
    for (int i = tid; i < 16 * NC; i += 32 * NWARPS) {
        const int n = i / NC, c = i - n * NC;
        if (c >= ncol) continue;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) s += red[(w * 16 + n) * NC + c];
        const int tap = c / p.Ci, ci = c - tap * p.Ci;
        out[((long long)tap * 16 + n) * p.Kp + ci] = s;
    }
}

int smem_bytes(int W, int halo, int NT) {
    const int pipe = 2 * ROWS * W * 32 + 4 * 4 * (ROWS + 2 * halo) * (W + 2 * halo + 2) * 2;
    const int red = NWARPS * 16 * 8 * NT * 4;
    return pipe > red ? pipe : red;
}

int pick_nseg(int Bn, int D, int H) {
    const int cols = Bn * ((H + ROWS - 1) / ROWS);
    int nseg = (2 * fcd_num_sms()) / cols;                   // two CTAs per SM, ONE wave (rounding up cost a second wave)
    if (nseg > D / 4) nseg = D / 4;                          // >= 4 planes per segment (two halo planes per segment)
    if (nseg < 1) nseg = 1;
    const int DL = (D + nseg - 1) / nseg;
    return (D + DL - 1) / DL;
}

template <int NT>
int launch(const SmallCParams& p, cudaStream_t st) {
    const int smem = smem_bytes(p.W, p.halo, NT);
    static int configured = 0;
    if (configured < smem) {
        if (cudaFuncSetAttribute(wgrad_smallc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return (int)cudaGetLastError();
        configured = smem;
    }
    wgrad_smallc_kernel<NT><<<p.Bn * p.nyt * p.nseg, 32 * NWARPS, smem, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace

// Number of partials (= CTAs) fcd_wgrad_smallc writes for a k x k x k stride-1 conv (k = 3 with pad 1, or k = 1) with Ci
// real input channels and Np padded output channels on (Bn, D, H, W); 0 = shape not taken.
FCD_API int fcd_wgrad_smallc_nsplit(int Bn, int D, int H, int W, int Ci, int Np, int k) {
    if (!(k == 3 || k == 1) || Ci < 1 || Ci > 2 || Np != 16 || W % 16 || W < 16 || W > 256) return 0;
    if ((long long)Bn * D * H * W < (1LL << 18)) return 0;
    if (smem_bytes(W, k == 3 ? 1 : 0, 7) > 112 * 1024) return 0;
    return Bn * ((H + ROWS - 1) / ROWS) * pick_nseg(Bn, D, H);
}

// X: NDHWC bf16 rows of pitch ldx (channels 0..Ci-1 real); dY: NDHWC bf16 rows of pitch ldy (16 channels);
// part: [nsplit][k^3][16][Kp] fp32 (entries with ci < Ci written), to be finished by fcd_wgrad_reduce.
FCD_API int fcd_wgrad_smallc(const void* X, long long ldx, const void* dY, long long ldy, float* part, int Bn, int D, int H,
                             int W, int Ci, int Kp, int k, int nsplit, cudaStream_t stream) {
    if (fcd_wgrad_smallc_nsplit(Bn, D, H, W, Ci, 16, k) != nsplit || nsplit < 1 || ldx % 2 || ldy % 8 || Kp < Ci) return -1;
    if (((uintptr_t)X & 3) || ((uintptr_t)dY & 15)) return -1;
    SmallCParams p;
    p.X = (const bf16*)X; p.ldx = ldx; p.dY = (const bf16*)dY; p.ldy = ldy; p.part = part;
    p.Bn = Bn; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.taps = k * k * k; p.halo = k == 3 ? 1 : 0;
    p.nyt = (H + ROWS - 1) / ROWS; p.nseg = pick_nseg(Bn, D, H); p.DL = (D + p.nseg - 1) / p.nseg; p.Kp = Kp;
    const int nt = (p.taps * Ci + 7) / 8;
    if (nt == 7) return launch<7>(p, stream);
    if (nt == 4) return launch<4>(p, stream);
    if (nt == 1) return launch<1>(p, stream);
    return -1;
}
