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

/* ---- layout (train.py:367-370 feeds fp32 NCDHW batches; the model returns NCDHW logits, train.py:374) ---- */
FCD_API int fcd_ncdhw_to_ndhwc(const float* src, void* dst, int B, int C, int Cp, long long S, cudaStream_t stream);
FCD_API int fcd_ndhwc_to_ncdhw(const void* src, float* dst, int B, int C, long long ld, long long S,
                               cudaStream_t stream);

/* ---- convolution family: nn.Conv3d / nn.ConvTranspose3d / nn.Linear
 *      (conv_blocks.py:393-437, 640-649, 225; ms_dsa_net.py:216, 362; segresnet_dsa.py:82, 97, 132) ---- */
FCD_API int fcd_pack_weight(const float* src, void* dst, int T, int N, int K, int Np, int Kp, long long sn,
                            long long sk, long long st, int kseg, int ksegpad, int nseg, int nsegpad,
                            cudaStream_t stream);
FCD_API int fcd_igemm(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias, int Bn,
                      int Ds, int Hs, int Ws, int Dm, int Hm, int Wm, int K, int N, int kd, int kh, int kw, int stride,
                      int pad, int mode, int out_mode, int accumulate, int Cq, cudaStream_t stream);
FCD_API int fcd_wgrad(const void* Q, long long ldq, const void* P, long long ldp, float* part, int Bn, int Ds, int Hs,
                      int Ws, int Dm, int Hm, int Wm, int Np, int Kp, int kd, int kh, int kw, int stride, int pad,
                      int nsplit, cudaStream_t stream);
FCD_API int fcd_wgrad_reduce(const float* part, float* out, int nsplit, int T, int N, int K, int Np, int Kp,
                             long long sn, long long sk, long long st, int kseg, int ksegpad, int accumulate,
                             cudaStream_t stream);

/* ---- torch.max_pool3d(x, 2, 2) (ms_dsa_net.py:92, 378-382) ---- */
FCD_API int fcd_maxpool2_fwd(const void* x, void* y, int B, int Do, int Ho, int Wo, int C, cudaStream_t stream);
FCD_API int fcd_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int B, int Do, int Ho, int Wo,
                             int C, int accumulate, cudaStream_t stream);

/* ---- InstanceNorm3d / BatchNorm3d / GroupNorm(2 ch per group) fused with LeakyReLU/ReLU and the residual add
 *      (conv_blocks.py:439-452, 56; ms_dsa_net.py:217; MONAI ResBlock).  mode: 0 instance, 1 batch,
 *      2 group-of-2. ---- */
FCD_API int fcd_norm_stats(const void* x, long long ld, float* part, float* mean, float* rstd, int B, long long S,
                           int C, int nchunk, int mode, float eps, float* running_mean, float* running_var,
                           int crun, float momentum, cudaStream_t stream);
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
                         int jaccard, float smooth_nr, float smooth_dr, float tv_w, int tv_norm, int tv_exclude,
                         unsigned char* keep, float* pbuf, float* part, float* tvpart, float* res,
                         cudaStream_t stream);
FCD_API int fcd_loss_bwd(const float* pred, const float* target, int B, int D, int H, int W, int kind,
                         float lambda_dice, float lambda_2, float w_bg, float w_fg, float gamma, int squared,
                         int jaccard, float smooth_nr, float smooth_dr, float tv_w, int tv_norm, int tv_exclude,
                         const unsigned char* keep, const float* pbuf, const float* res, const float* gout,
                         float* dpred, cudaStream_t stream);

#endif /* FCD_B200_H */
