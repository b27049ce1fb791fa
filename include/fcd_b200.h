/* fcd_b200 -- C ABI of the B200-native volumetric-segmentation hot path (sm_100a).
 *
 * The reference (mehdirabiee/fcd) is pure Python and has no FFI of its own: its hot path is a sequence of
 * torch / MONAI library calls.  Each entry point below replaces one group of those calls; the reference call
 * site it replaces is cited as file:line (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator in the shipped host
 *     code); the library allocates nothing and never synchronises;
 *   - activations are channels-last bf16 rows: element (b,z,y,x,c) of a [B][D][H][W][C] tensor lives at
 *     ptr[(((b*D+z)*H+y)*W+x)*ld + c], ld >= C, ld % 8 == 0 (so a channel slice of a wider buffer is a valid
 *     operand -- this is how torch.cat is eliminated);
 *   - every function launches on `stream` and returns a cudaError_t value (0 = success) or -1 when the
 *     shape is outside what the kernel supports.  Nothing falls back to the CPU.
 */
#ifndef FCD_B200_H
#define FCD_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
#define FCD_API extern "C" __attribute__((visibility("default")))
#else
#define FCD_API
#endif

/* ---- status block: every pipeline wait of the tcgen05 kernels is bounded; the first wait that times out writes the
 * sticky error word (kernel id << 24 | wait site << 16 | CTA; kernel ids: 1 fcd_conv3_tc, 2 fcd_conv3_tcf, 6 fcd_rowgemm,
 * 3 fcd_conv_gemm_tc, 4 fcd_wgrad3_tc, 5 fcd_wgrad_gemm_tc) and a debug record into a device-resident block of
 * FCD_STATUS_INTS ints: [0] error word, [1] kernel id, [2] wait site, [3] CTA, [4] thread, [5] mbarrier shared
 * address, [6] parity waited for, [7] work item, [8] grid size, [9] blockIdx.y, [16..48) the CTA's per-role progress
 * counters (meaning documented at the top of each kernel).  fcd_loss_fwd and fcd_sw_finalize read word 0 on the device
 * and turn a non-zero value into NaN results (the reference has no such failure mode: train.py:374-382 would simply
 * hang or crash), so a fault cannot pass silently.  fcd_status copies the block to host_out[FCD_STATUS_INTS] (NULL:
 * skip), clears it when `clear` != 0, and returns word 0; it synchronises the device.  The per-kernel *_error
 * accessors below are aliases: value-and-clear of the same word. ---- */
#define FCD_STATUS_INTS 48
FCD_API int fcd_status(int* host_out, int clear);
FCD_API int fcd_status_device_ptr(void** out);

/* ---- layout (train.py:367-370 feeds fp32 NCDHW batches; the model returns NCDHW logits, train.py:374) ---- */
FCD_API int fcd_ncdhw_to_ndhwc(const float* src, void* dst, int B, int C, int Cp, long long S, cudaStream_t stream);
FCD_API int fcd_ndhwc_to_ncdhw(const void* src, float* dst, int B, int C, long long ld, long long S,
                               cudaStream_t stream);

/* ---- convolution family: nn.Conv3d / nn.ConvTranspose3d / nn.Linear
 *      (conv_blocks.py:393-437, 640-649, 225; ms_dsa_net.py:216, 362; segresnet_dsa.py:82, 97, 132) ---- */
FCD_API int fcd_pack_weight(const float* src, void* dst, int T, int N, int K, int Np, int Kp, long long sn,
                            long long sk, long long st, int kseg, int ksegpad, int nseg, int nsegpad,
                            cudaStream_t stream);
FCD_API int fcd_pack_weight_batched(const void* jobs, int njobs, int nblocks, cudaStream_t stream);
FCD_API int fcd_igemm(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias, int Bn,
                      int Ds, int Hs, int Ws, int Dm, int Hm, int Wm, int K, int N, int kd, int kh, int kw, int stride,
                      int pad, int mode, int out_mode, int accumulate, int Cq, cudaStream_t stream);
/* data gradient of a 3x3x3 stride-2 pad-1 conv (segresnet_dsa.py:97) as eight parity-class launches that multiply only
 * the taps reaching each class (W: packed transposed weights [27][N][K] as for fcd_igemm mode 1; even Dm, Hm, Wm) */
FCD_API int fcd_igemm_dgrad_s2(const void* A, long long lda, const void* W, void* C, long long ldc, int Bn, int Ds, int Hs,
                               int Ws, int Dm, int Hm, int Wm, int K, int N, cudaStream_t stream);
FCD_API int fcd_igemm_ksplit(long long M, int N, int K, int T);
FCD_API int fcd_igemm_splitk(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias,
                             int Bn, int Ds, int Hs, int Ws, int Dm, int Hm, int Wm, int K, int N, int kd, int kh,
                             int kw, int stride, int pad, int mode, int accumulate, float* ws, int ksplit,
                             cudaStream_t stream);
FCD_API int fcd_splitk_reduce(const float* ws, void* C, long long ldc, const float* bias, long long M, int N, int ksplit,
                              int accumulate, cudaStream_t stream);
/* Pointwise contractions on large volumes -- 1x1x1 Conv3d / nn.Linear rows (conv_blocks.py:420-437, 57, 225;
 * ms_dsa_net.py:215) and ConvTranspose3d k2 s2 (conv_blocks.py:640-649) forward / data gradient -- as ONE persistent
 * TMA + tcgen05 kernel (csrc/rowgemm_tma.cu): 128-row tiles streamed by cp.async.bulk.tensor, weights resident in shared
 * memory, two TMEM accumulators.  mode 0: C[m][n] = sum_k A[m][k] Wp[n][k] + bias[n]; mode 1: N = 8*Cq columns scattered
 * to the 2x2x2 fine voxels of the coarse voxel m (C = the fine, possibly concat, buffer); mode 2: data gradient of the
 * transposed conv (A = the fine grid, 8 strided taps).  K in {16,32,64}, N in {16,...,256}; _ok() = 1 if taken. */
FCD_API int fcd_rowgemm_ok(int mode, int Bn, int D, int H, int W, long long M, int K, int N);
FCD_API int fcd_rowgemm(int mode, const void* A, long long lda, const void* Wp, void* C, long long ldc, const float* bias,
                        int Bn, int D, int H, int W, long long M, int K, int N, int Cq, cudaStream_t stream);
/* tcgen05 split-K GEMM form of the 3x3x3 stride-1 pad-1 convs of the DEEP levels (K, N multiples of 64; conv_blocks.py:
 * 393-416 at encoder levels 3-6 / decoder / TransformerBlock.conv51): streamed packed weights, 128-voxel x up-to-256-
 * channel tiles, (tap, k-chunk) loop split over gridDim.z.  mode 0 forward, 1 data gradient. */
FCD_API int fcd_conv_gemm_tc_ksplit(long long M, int K, int N);
FCD_API int fcd_conv_gemm_tc_ksplit_vol(int Bn, int D, int H, int W, int K, int N);
FCD_API int fcd_conv_gemm_tc(const void* A, long long lda, const void* Wp, void* C, long long ldc, float* ws, int Bn,
                             int D, int H, int W, int K, int N, int mode, int ksplit, cudaStream_t stream);
/* feed selection: volumes whose 128-voxel row tiles are boxes (W, H, D powers of two) are fed by TMA halo tiles
 * (cp.async.bulk.tensor, out-of-volume coordinates = the zero padding); use_tma(0) forces the cp.async feed, (-1) only
 * queries; both return the previous setting. */
FCD_API int fcd_conv_gemm_tc_use_tma(int on);
FCD_API int fcd_conv_gemm_tc_tma_ok(int Bn, int D, int H, int W);
FCD_API int fcd_gemm_tc_error(void);
/* weight gradient of the same deep-level convs: voxels as the GEMM K dimension, 128 input channels x up to 256 output
 * channels per CTA, one CTA per (tap, tile, voxel slice); part[nsplit][27][Np][Kp] is finished by fcd_wgrad_reduce. */
FCD_API int fcd_wgrad_gemm_tc_nsplit(long long M, int Kp, int Np);
FCD_API int fcd_wgrad_gemm_tc(const void* X, long long ldx, const void* dY, long long ldy, float* part, int Bn, int D,
                              int H, int W, int Kp, int Np, cudaStream_t stream);
FCD_API int fcd_wgrad_gemm_tc_error(void);
/* weight gradient of the first conv of every network (k = 3 pad 1, or k = 1; stride 1; <= 2 real input channels in a
 * 16-channel row, 16 output channels, large volume): the (tap, ci) pairs as the N dimension of one mma.sync contraction
 * over the voxels (csrc/wgrad_smallc.cu).  _nsplit() = number of partials (0 = shape not taken);
 * part[nsplit][k^3][16][Kp] is finished by fcd_wgrad_reduce. */
FCD_API int fcd_wgrad_smallc_nsplit(int Bn, int D, int H, int W, int Ci, int Np, int k);
FCD_API int fcd_wgrad_smallc(const void* X, long long ldx, const void* dY, long long ldy, float* part, int Bn, int D, int H,
                             int W, int Ci, int Kp, int k, int nsplit, cudaStream_t stream);
FCD_API int fcd_wgrad_group(long long M, int T);
FCD_API int fcd_wgrad(const void* Q, long long ldq, const void* P, long long ldp, float* part, int Bn, int Ds, int Hs,
                      int Ws, int Dm, int Hm, int Wm, int Np, int Kp, int kd, int kh, int kw, int stride, int pad,
                      int nsplit, cudaStream_t stream);
FCD_API int fcd_wgrad_reduce(const float* part, float* out, int nsplit, int T, int N, int K, int Np, int Kp,
                             long long sn, long long sk, long long st, int kseg, int ksegpad, int accumulate,
                             cudaStream_t stream);

/* tcgen05/TMEM implicit-GEMM path of the 3x3x3 stride-1 pad-1 convs (conv_blocks.py:393-416 conv1/conv2 of
 * UnetResBlock; MONAI ResBlock segresnet_dsa.py:102) and, with flip=1, their data gradients.  TMA halo planes in,
 * NDHWC bf16 out, weights read straight from the fp32 parameter (no pack kernel), optional fp32 bias[N] added before
 * the bf16 rounding (MONAI SubpixelUpsample's conv, conv_blocks.py:727-735), optional fused InstanceNorm partial statistics part[Bn][nchunk][2][N] (sum, sum of squares of the
 * rounded outputs; nchunk = (H/16)*(W/8)*nseg), finished either by fcd_norm_finalize or -- when mean / rstd
 * ([Bn][N] fp32) are given -- by the LAST CTA of the conv itself (norm_mode / eps / running statistics as fcd_norm_stats:
 * the InstanceNorm / BatchNorm that follows the conv, conv_blocks.py:439-452, needs no launch of its own for them).  fcd_conv3_tc_nseg returns the
 * number of d-segments to use (0: shape unsupported -> use fcd_igemm).  fcd_tc_error: first timed-out pipeline wait
 * since the last call (0 = none; test / debug aid, synchronises the device). */
FCD_API int fcd_norm_fin_fold(int B, int nchunk, int L);   /* 1: B x nchunk partial rows of L floats fit the last-CTA finalize */
FCD_API int fcd_conv3_tc_nseg(int Bn, int D, int H, int W, int K, int N);
FCD_API int fcd_conv3_tc(const void* A, long long lda, const float* Wf, int Nr, int Kr, long long sn, long long sk,
                         long long st, int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, float* part,
                         const float* bias, int Bn, int D, int H, int W, int K, int N, int flip, int nseg, float* mean,
                         float* rstd,
                         int norm_mode, float eps, float* running_mean, float* running_var, int crun, float momentum,
                         cudaStream_t stream);
FCD_API int fcd_tc_error(void);
/* kd-folded variant for N in {16, 32}: one instruction of N = 3*Cout feeds three consecutive output planes from one A
 * tile (csrc/conv_tcf.cu).  Same arguments and results as fcd_conv3_tc; -1 when the shape is not taken.  accumulate = 1:
 * C += result (K-sliced data gradients of convs with more than 64 output channels: one launch per 64 channels of dY). */
FCD_API int fcd_conv3_tcf(const void* A, long long lda, const float* Wf, int Nr, int Kr, long long sn, long long sk,
                          long long st, int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, float* part,
                          const float* bias, int accumulate, int Bn, int D, int H, int W, int K, int N, int flip, int nseg,
                          float* mean, float* rstd,
                          int norm_mode, float eps, float* running_mean, float* running_var, int crun, float momentum,
                          cudaStream_t stream);
FCD_API int fcd_tcf_error(void);

/* Pointwise (1x1x1, stride 1, no bias) conv on >= 65536 voxels with 16/32 (padded) channels each side: UnetResBlock's
 * residual conv3 (conv_blocks.py:420-424) on the two top levels, forward and data gradient (A = dY, sn = 1, sk = Cin).
 * W(n, k) = Wf[n*sn + k*sk] for real n < Nr, k < Kr under the concat-segment maps of fcd_pack_weight. */
FCD_API int fcd_pw_conv_ok(long long M, int K, int N);
FCD_API int fcd_pw_conv(const void* A, long long lda, const float* Wf, long long sn, long long sk, int Nr, int Kr,
                        int kseg, int ksegpad, int nsg, int nsgpad, void* C, long long ldc, long long M, int K, int N,
                        cudaStream_t stream);

/* tcgen05/TMEM weight gradient of the same 3x3x3 stride-1 pad-1 convs (autograd of conv_blocks.py:393-416).  S: the
 * operand read shifted (conv input, CS channels from channel k_off), U: the unshifted one (output gradient, CU
 * channels from n_off); CS, CU in {16, 32}; wider layers are cut into nns x nks slices of CU x CS channels that run as
 * CTA rows of the same launch.  CTA column c of the fcd_wgrad3_tc_nsplit() columns writes its slice of the partial
 * part[c][27][ldn][ldk]; fcd_wgrad_reduce sums the partials. */
FCD_API int fcd_wgrad3_tc_nsplit(int Bn, int D, int H, int W);
FCD_API int fcd_wgrad3_tc(const void* S, long long lds, const void* U, long long ldu, float* part, int ldn, int ldk,
                          int n_off, int k_off, int nns, int nks, int Bn, int D, int H, int W, int CS, int CU,
                          cudaStream_t stream);
FCD_API int fcd_wgrad_tc_error(void);


/* ---- torch.max_pool3d(x, 2, 2) (ms_dsa_net.py:92, 378-382).  bwd: add (optional, rows of pitch ldadd) is the gradient
 *      x receives from its other consumer (the skip connection, ms_dsa_net.py:386-390): dx = pool_bwd(dy) + add. ---- */
FCD_API int fcd_maxpool2_fwd(const void* x, void* y, int B, int Do, int Ho, int Wo, int C, cudaStream_t stream);
FCD_API int fcd_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, const void* add, long long ldadd,
                             int B, int Do, int Ho, int Wo, int C, int accumulate, cudaStream_t stream);

/* ---- InstanceNorm3d / BatchNorm3d / GroupNorm(2 ch per group) fused with LeakyReLU/ReLU and the residual add
 *      (conv_blocks.py:439-452, 56; ms_dsa_net.py:217; MONAI ResBlock).  mode: 0 instance, 1 batch,
 *      2 group-of-2. ---- 
 * fcd_norm_bwd with x1 == NULL: xhat1 is reconstructed from the saved output, xhat1 = act^-1(y) - xhat2 (requires y, slope > 0, no
 * gamma1, no dres): the conv output x1 is then neither read by the backward nor kept by the forward. */
FCD_API int fcd_norm_stats(const void* x, long long ld, float* part, float* mean, float* rstd, int B, long long S,
                           int C, int nchunk, int mode, float eps, float* running_mean, float* running_var,
                           int crun, float momentum, cudaStream_t stream);
FCD_API int fcd_norm_finalize(const float* part, float* mean, float* rstd, int B, long long S, int C, int nchunk,
                              int mode, float eps, float* running_mean, float* running_var, int crun, float momentum,
                              cudaStream_t stream);
FCD_API int fcd_colsum(const void* x, long long ld, float* part, float* out, long long rows, int C, int nchunk,
                       cudaStream_t stream);
FCD_API int fcd_norm_apply(const void* x1, long long ld1, const float* mean1, const float* rstd1, const float* gamma1,
                           const float* beta1, const void* x2, long long ld2, const float* mean2, const float* rstd2,
                           const void* res, long long ldr, void* y, long long ldy, int B, long long S, int C,
                           float slope, cudaStream_t stream);
FCD_API int fcd_norm_bwd(const void* dy, long long lddy, const void* y, long long ldy, const void* x1, long long ld1,
                         const float* mean1, const float* rstd1, const float* gamma1, const void* x2, long long ld2,
                         const float* mean2, const float* rstd2, float* part, float* coef, float* dgamma,
                         float* dbeta, void* dx1, long long ldd1, void* dx2, long long ldd2, void* dres,
                         long long lddr, int acc_res, int B, long long S, int C, int nchunk, int mode, float slope,
                         cudaStream_t stream);
FCD_API int fcd_add(const void* a, long long lda, const void* b, long long ldb, void* o, long long ldo, long long rows,
                    int C, cudaStream_t stream);
FCD_API int fcd_copy_rows(const void* a, long long lda, void* o, long long ldo, long long rows, int C,
                          cudaStream_t stream);

/* ---- output head: 1x1x1 conv + bias -> fp32 NCDHW logits (ms_dsa_net.py:82, 362; segresnet_dsa.py:188-193) ---- */
FCD_API int fcd_outconv_blocks(void);
FCD_API int fcd_outconv_fwd(const void* x, long long ld, const float* w, const float* bias, float* out, int B,
                            long long S, int C, int Co, cudaStream_t stream);
FCD_API int fcd_outconv_bwd(const void* x, long long ld, const float* w, const float* dout, void* dx, long long lddx,
                            float* part, float* dw, float* db, int B, long long S, int C, int Co, cudaStream_t stream);

/* ---- CombinedLoss.forward: Dice (+CE | +focal) (+TV) (get_loss.py:24-39, 46-78, 100-165) ---- */
FCD_API int fcd_loss_blocks(void);
FCD_API int fcd_loss_fwd(const float* pred, const float* target, int B, int D, int H, int W, int kind,
                         float lambda_dice, float lambda_2, float w_bg, float w_fg, float gamma, int squared,
                         int jaccard, int w_type, float smooth_nr, float smooth_dr, float tv_w, int tv_norm,
                         int tv_exclude, unsigned char* keep, float* pbuf, float* part, float* tvpart, float* res,
                         cudaStream_t stream);
FCD_API int fcd_loss_bwd(const float* pred, const float* target, int B, int D, int H, int W, int kind,
                         float lambda_dice, float lambda_2, float w_bg, float w_fg, float gamma, int squared,
                         int jaccard, int w_type, float smooth_nr, float smooth_dr, float tv_w, int tv_norm,
                         int tv_exclude, const unsigned char* keep, const float* pbuf, const float* res, const float* gout,
                         float* dpred, cudaStream_t stream);

/* ---- on-device patch sampling + augmentation (get_transforms.py:45-89: RandCropByPosNegLabeld, RandFlipd x 3,
 * RandRotated(range_y), RandShiftIntensityd, RandGaussianNoised, RandCoarseDropoutd, GridMaskd = utils/gridmask.py:8-72)
 * on a volume resident in HBM; decisions = counter hash of (seed, sample), recorded in
 * meta[S][fcd_sampling_meta_floats()]: 0-2 z0, y0, x0; 3 flip bits; 4 shift; 5 noise std; 6 class; 7 rank; 8-10 cz, cy, cx;
 * 11 rotated; 12 cos; 13 sin; 14 angle; 15 holes applied; 16-39 hole corners (z, y, x) x 8; 40 grid mask on; 41 period d;
 * 42 stripe width; 43-45 phases; 46 inverted; 47 reserved.
 * pick_centers: rot_p / rot_range = probability and half-range (radians) of the rotation about spatial axis 1; cd_p,
 * holes (<= 8), (hz, hy, hx) = coarse dropout; grid_p, [d1, d2), grid_ratio, grid_invert = GridMask.
 * crop_augment: (hz, hy, hx) as above, hh = ceil(sqrt(rd^2 + rh^2 + rw^2)) (the mask cube of utils/gridmask.py:31);
 * D*H*W < 2^31, S <= 65535, ceil(rd*rh / rows per block) <= 65535 (else -1). ---- */
FCD_API int fcd_sampling_block_voxels(void);
FCD_API int fcd_sampling_meta_floats(void);
FCD_API int fcd_fg_block_counts(const float* label, long long V, int* counts, cudaStream_t stream);
FCD_API int fcd_pick_centers(const float* label, const int* counts, int D, int H, int W, int rd, int rh, int rw, int S,
                             unsigned long long seed, float pos_ratio, float flip_p, float shift_max, float shift_p,
                             float noise_std, float noise_p, float rot_p, float rot_range, float cd_p, int holes, int hz,
                             int hy, int hx, float grid_p, int d1, int d2, double grid_ratio, int grid_invert,
                             float* meta, cudaStream_t stream);
FCD_API int fcd_crop_augment(const float* img, const float* label, int C, int D, int H, int W, int rd, int rh, int rw,
                             int S, const float* meta, unsigned long long seed, int hz, int hy, int hx, int hh,
                             float* out_img, float* out_lab, cudaStream_t stream);

/* ---- TransformerBlock token path: pos_embed add + LayerNorm (conv_blocks.py:72-77) ---- */
FCD_API int fcd_ln_fwd(const void* x, long long ldx, const float* pos, const float* w, const float* b, void* t,
                       long long ldt, void* ln, long long ldl, float* mean, float* rstd, long long rows, int N, int C,
                       int Cp, float eps, cudaStream_t stream);
FCD_API int fcd_ln_bwd_blocks(int N, int Cp);
FCD_API int fcd_ln_bwd(const void* dln, long long lddl, const void* dtd, long long lddt, const void* t, long long ldt,
                       const float* mean, const float* rstd, const float* w, void* dx, long long lddx, float* dpos,
                       float* part, float* dw, float* db, int B, int N, int C, int Cp, cudaStream_t stream);

/* ---- nn.Dropout3d / nn.Dropout channel masks (conv_blocks.py:57, 347): out[i] = keep_i / (1 - p), keep_i a counter-based
 *      Bernoulli(1-p) bit of (seed, *seed_dev, i); seed_dev as for fcd_dsa_fwd. ---- */
FCD_API int fcd_keep_scale(float* out, int n, float p, long long seed, const long long* seed_dev, cudaStream_t stream);

/* ---- DSA.forward, sa_type='parallel' (conv_blocks.py:328-355) fused with `x + gamma * dsa` (line 77).
 *      ca_scale: optional [B][H][c][c] dropout scale (0 or 1/(1-p)) for attn_drop; sa_drop/seed: in-kernel
 *      counter-based dropout of the [N,P] spatial attention map (attn_drop_2); seed_dev: optional DEVICE step counter
 *      mixed into the seed, so that replays of a captured CUDA graph draw fresh masks. ---- */
FCD_API int fcd_dsa_fwd_part_floats(int B, int N, int C, int H, int P);
FCD_API int fcd_dsa_bwd_part_floats(int B, int N, int C, int H, int P);
FCD_API int fcd_dsa_fwd(const void* qkvv, long long ldq, const float* EF, const float* temperature,
                        const float* temperature2, const float* gamma, const void* t, long long ldt, void* y,
                        long long ldy, const float* ca_scale, float sa_drop, long long seed, const long long* seed_dev,
                        float* part, float* inv_n,
                        float* Ghat, float* A, float* Ad, float* KV, float* xca, float* tsa, int B, int N, int C,
                        int Cp, int H, int P, cudaStream_t stream);
FCD_API int fcd_dsa_bwd(const void* qkvv, long long ldq, const void* dy, long long lddy, const float* EF,
                        const float* temperature, const float* temperature2, const float* gamma,
                        const float* ca_scale, float sa_drop, long long seed, const long long* seed_dev,
                        const float* inv_n, const float* Ghat,
                        const float* A, const float* Ad, const float* KV, const float* xca, const float* tsa,
                        float* part, float* dqh, float* dKV, float* dGhat, float* rqk, float* gpart, void* dqkvv,
                        long long lddq, float* dEF, float* dtemp, float* dtemp2, float* dgamma, int B, int N, int C,
                        int Cp, int H, int P, cudaStream_t stream);
/* dEF alone (fcd_dsa_bwd with dEF == NULL skips it): a parameter gradient, launchable on a side stream once
 * fcd_dsa_bwd has produced dKV [B][2][C][P]. */
FCD_API int fcd_dsa_bwd_ef(const void* qkvv, long long ldq, const float* dKV, float* dEF, int B, int N, int C, int P,
                           cudaStream_t stream);

/* ---- MONAI SubpixelUpsample tail: pixelshuffle x2 + pad(1,0)x3 + AvgPool3d(2,1) (+ skip add / concat write)
 *      (conv_blocks.py:727-735, 771; segresnet_dsa.py:133-141, 217).  src channels are ordered tap*Cq + c. ---- */
FCD_API int fcd_ps_blur_fwd(const void* src, long long lds, const void* skip, long long ldk, void* out, long long ldo,
                            int B, int D, int H, int W, int Cq, cudaStream_t stream);
FCD_API int fcd_ps_blur_bwd(const void* dout, long long lddo, void* dsrc, long long ldds, int B, int D, int H, int W,
                            int Cq, cudaStream_t stream);
/* UpSample(mode="nontrainable") = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False) (same call sites
 * with upsample_mode="nontrainable"); optional skip add / concat-buffer output as above. */
FCD_API int fcd_trilinear_up_fwd(const void* src, long long lds, const void* skip, long long ldk, void* out,
                                 long long ldo, int B, int D, int H, int W, int C, cudaStream_t stream);
FCD_API int fcd_trilinear_up_bwd(const void* dout, long long lddo, void* dsrc, long long ldds, int B, int D, int H,
                                 int W, int C, cudaStream_t stream);

/* ---- F.mse_loss of the VAE reconstruction (segresnet_dsa.py:357); part: fcd_loss_blocks() floats ---- */
FCD_API int fcd_mse_fwd(const float* a, const float* b, long long n, float* part, float* out, cudaStream_t stream);
FCD_API int fcd_mse_bwd(const float* a, const float* b, long long n, const float* gout, float* da,
                        cudaStream_t stream);

/* ---- sliding_window_inference(mode="constant") support (train.py:148-165; seg_fcd_test.py:37-54; label map of
 *      train.py:185,209-211 / get_transforms.py:142-154).  starts_zyx is a HOST array of nwin*3 ints. ---- */
FCD_API int fcd_sw_gather(const float* vol, void* dst, int C, int Cp, int D, int H, int W, int r0, int r1, int r2,
                          int pz, int py, int px, const int* starts_zyx, int nwin, cudaStream_t stream);
/* blend: out[c*sc + (z0+z)*sz + (y0+y)*Wp + (x0+x)] += pred[c][z][y][x] (fp32, one launch per window: the reference's
 * window order).  The accumulation volume is channel-major [C][Dp][Hp][Wp] (sc = Dp*Hp*Wp, sz = Hp*Wp) or plane-major
 * [Dp][C][Hp][Wp] (sc = Hp*Wp, sz = C*Hp*Wp), the layout whose D-slabs are contiguous for the multi-GPU reduce-scatter.
 * finalize: unpadded planes z in [z_lo, z_hi): v = acc / (cz*cy*cx) (per-axis coverage counts, padded frame), written
 * to dst (optional, [C][out_planes][H][W], plane z - out_z0) and turned into the label map (mode 1: softmax >= 0.5 per
 * channel, float; mode 2: argmax, uint8; same plane addressing); acc plane 0 is padded plane acc_z0. */
FCD_API int fcd_sw_blend(const float* pred, float* out, int C, int r0, int r1, int r2, int Wp, long long sc,
                         long long sz, int z0, int y0, int x0, cudaStream_t stream);
FCD_API int fcd_sw_finalize(const float* acc, const int* cz, const int* cy, const int* cx, float* dst, float* label_f,
                            void* label_u8, int C, int H, int W, int Wp, long long sc, long long sz, int pz, int py,
                            int px, int z_lo, int z_hi, int acc_z0, int out_z0, int out_planes, int mode,
                            cudaStream_t stream);

/* ---- post-processing of the predicted mask: utils/utils_common.py:10-33 post_process_segment as called by
 *      ModelTrainer.post_process (train.py:167-182), on the device and bit-exact against the reference's scipy calls:
 *      binary_opening (6-connected, 1 iteration) -> binary_fill_holes(structure = ones 5^3) -> label(structure = ones
 *      3^3, numbered in raster order) -> keep components with >= l_min voxels (l_min = -1: the largest; the reference's
 *      quirks for empty / full masks and l_min <= 0 are kept).  Input: pred_f (fp32 [D][H][W], mask = pred_f >
 *      threshold, train.py:173) or pred_u8 (uint8, mask = != 0), exactly one non-NULL.  Outputs (either may be NULL):
 *      out_mask, out_lab fp32 [D][H][W] = output_msk / output_lab of the reference.  ws: fcd_post_process_ws_bytes()
 *      bytes of device scratch, 256-byte aligned.  No host synchronisation. ---- */
FCD_API long long fcd_post_process_ws_bytes(int D, int H, int W);
FCD_API int fcd_post_process(const float* pred_f, const void* pred_u8, float threshold, int l_min, float* out_mask,
                             float* out_lab, int D, int H, int W, void* ws, long long ws_bytes, cudaStream_t stream);

/* ---- voxel-level evaluation counts (metrics.py:74-126 `_compute_metrics`, called from train.py:220; seg_fcd_test.py:
 *      160-178 Dice / IoU; utils/utils_common.py:37-60 `evaluate_fp`), integer and bit-exact.
 *      fcd_confusion_counts: `items` volumes of n voxels each (one per (subject, channel)), prediction fp32 (pred_f) or
 *      uint8 (pred_u8), exactly one non-NULL; counts[items][4] = tp, fp, tn, fn of (pred > thr_pred) vs (label >
 *      thr_label) (MONAI get_confusion_matrix's column order); the call zeroes `counts` itself.
 *      fcd_component_overlap: cc = component ids (fp32 whole numbers in [0, max_id], 0 = background, e.g. out_lab of
 *      fcd_post_process), label = ground truth; out[0] = distinct ids present, out[1] = ids with a voxel where label != 0,
 *      out[2] = voxels with an id outside [0, max_id] (must be 0); evaluate_fp(cc, label) = out[0] - out[1].
 *      ws: fcd_component_overlap_ws_bytes(max_id) bytes of device scratch. ---- */
FCD_API int fcd_confusion_counts(const float* pred_f, const void* pred_u8, const float* label, float thr_pred,
                                 float thr_label, long long n, int items, long long* counts, cudaStream_t stream);
FCD_API long long fcd_component_overlap_ws_bytes(long long max_id);
FCD_API int fcd_component_overlap(const float* cc, const float* label, long long V, long long max_id, void* ws,
                                  long long ws_bytes, long long* out, cudaStream_t stream);

/* ---- optimizer step: torch.optim.AdamW as built by train_utils.py:63-71 and stepped at train.py:382, for ALL parameter
 *      tensors in one launch.  jobs: device array of njobs records {float* p; const float* g; float* m; float* v;
 *      long long n; long long blk0;} (48 bytes, blk0 = running sum of ceil(n / fcd_adamw_chunk()), ascending from 0);
 *      nblocks = the total; step: device fp32 scalar with the 1-based count of this update.  Decoupled weight decay,
 *      bias correction, eps after the square root -- the arithmetic of torch's implementation, fp32. ---- */
FCD_API int fcd_adamw_chunk(void);
FCD_API int fcd_adamw_multi(const void* jobs, int njobs, int nblocks, const float* step, double lr, double beta1,
                            double beta2, double eps, double weight_decay, cudaStream_t stream);

#endif /* FCD_B200_H */
