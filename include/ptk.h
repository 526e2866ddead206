/*
 * ptk.h -- C ABI of the B200 (sm_100a) kernel library for the deformable-GAN training step.
 *
 * The reference (saurabhsharma1993/pose-transfer, src_deformable/) ships no native code: every
 * entry point below replaces a torch/cuDNN/cv2 LIBRARY CALL SITE of the reference, cited per
 * function as path:line relative to /root/reference/src_deformable/.
 *
 * Conventions
 *   - All activations are fp32 NHWC with an explicit pixel stride `ld` (floats): element (n,y,x,c) of a
 *     tensor lives at base[((n*H + y)*W + x)*ld + c].  `ld >= C` lets producers write straight into a
 *     channel slice of a wider concat buffer (this replaces the reference's torch.cat call sites,
 *     models/networks.py:241,245,284,286 and models/pose_gan.py:86,133-136).
 *   - Caller (PyTorch) owns every buffer, including scratch; the library never allocates or frees.
 *   - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately.
 *   - Return value: 0 = OK, otherwise an error code; ptk_last_error() gives the (thread-local) message.
 *   - Activation codes: 0 = identity, 1 = LeakyReLU(0.2), 2 = ReLU, 3 = tanh, 4 = sigmoid.
 */
#ifndef PTK_H_
#define PTK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTK_ACT_NONE 0
#define PTK_ACT_LEAKY 1
#define PTK_ACT_RELU 2
#define PTK_ACT_TANH 3
#define PTK_ACT_SIGMOID 4

#define PTK_IMPL_AUTO 0
#define PTK_IMPL_SIMT 1 /* fp32 CUDA-core implicit GEMM (exact-fp32 path) */
#define PTK_IMPL_TC 2   /* tcgen05 TF32 tensor-core implicit GEMM        */

int ptk_version(void);
const char* ptk_last_error(void);

/* ---------------------------------------------------------------- layout (replaces torch.cat / views) */
/* dst[n,y,x,c_dst0+c] = act(src[n,c_src0+c,y,x]); src NCHW with C_src channels, dst NHWC (ld_dst).
 * utils/pose_utils.py:227-233 (get_imgpose slices) + models/networks.py:271, models/pose_gan.py:86,133-136 */
int ptk_nchw_to_nhwc(const float* src, int C_src, int c_src0, float* dst, int ld_dst, int c_dst0,
                     int N, int C, int H, int W, int act, void* stream);
/* dst[n,c,y,x] = src[n,y,x,c_src0+c]  (NHWC slice -> dense NCHW, e.g. out_gen for the caller) */
int ptk_nhwc_to_nchw(const float* src, int ld_src, int c_src0, float* dst, int N, int C, int H, int W,
                     void* stream);
/* dst[n,y,x, c_dst0 + c] for c in [0, c_total) = srcs[q][n, c_src0[q] + (c - c_dst[q]), y, x] where segment q covers c, else 0:
 * assembles NHWC rows from up to four NCHW channel ranges in one pass and writes the whole range (float4 stores of full
 * sectors).  Replaces the per-slice copies of get_imgpose / torch.cat (utils/pose_utils.py:227-233, models/pose_gan.py:84-86,
 * 131-136) when a buffer is filled from several tensors.  c_total, c_dst0, ld_dst multiples of 4; c_total <= 96. */
int ptk_gather_nhwc(const float* const* srcs, const int* src_channels, const int* c_src0, const int* C, const int* c_dst,
                    int nseg, float* dst, int ld_dst, int c_dst0, int c_total, int N, int H, int W, void* stream);
/* Weight repack.  src is the torch layout [A][B][k][k] (Conv2d: A=Cout,B=Cin; ConvTranspose2d: A=Cin,
 * B=Cout; models/networks.py:154,156).  transpose==0: dst[tap][a][b_pad]; transpose==1: dst[tap][b][a_pad].
 * Padding entries are written as zero. */
int ptk_pack_weight(const float* src, float* dst, int A, int B, int taps, int A_pad, int B_pad,
                    int transpose, void* stream);
/* Both layouts in one pass: dst0[tap][a][b] with row count rows0 / row stride cols0, dst1[tap][b][a] with rows1 / cols1.
 * Padding entries are NOT written (the destination buffers must have been zero-initialised once). */
int ptk_pack_weight_dual(const float* src, float* dst0, float* dst1, int A, int B, int taps, int rows0, int cols0,
                         int rows1, int cols1, void* stream);
/* grad[a][b][tap] (+)= src[tap][a][b_pad]  (inverse of transpose==0 packing, drops padding) */
int ptk_unpack_weight_grad(const float* src, float* grad, int A, int B, int taps, int B_pad,
                           int accumulate, void* stream);
/* grad[a][b][tap] (+)= sum_{q < nparts} src[q * part_stride + (tap*A + a)*B_pad + b] */
int ptk_unpack_weight_grad_parts(const float* src, int nparts, int64_t part_stride, float* grad, int A, int B,
                                 int taps, int B_pad, int accumulate, void* stream);
/* dst[tap][b][a] = src[tap][a][b]; A and Bp multiples of 32.  With the weights kept in GEMM layout ([tap][A][B_pad] =
 * the K-major tensor-core operand of one direction = the layout the weight gradient is produced in) this is the only
 * repacking left per optimiser step. */
int ptk_transpose_weight(const float* src, float* dst, int taps, int A, int Bp, void* stream);
/* dst[i] (+)= sum_{q < nparts} src[q * part_stride + i]  (fixed summation order; n, part_stride multiples of 4) */
int ptk_sum_parts(const float* src, int nparts, int64_t part_stride, float* dst, int64_t n, int accumulate, void* stream);
int ptk_fill(float* dst, int64_t n, float value, void* stream);

/* ---------------------------------------------------------------- convolutions
 * One geometry struct describes Conv2d and ConvTranspose2d(+Cropping2D) fprop/dgrad/wgrad
 * (models/networks.py:154 Conv k4 s2 p1, :156-157 ConvT k4 s2 p0 + crop 1 == p1, :186 stem k3 s1 p1,
 *  :232 final k3, :341 D stem k4 s2 p0; torchvision vgg19.features[0] is fused in ptk_nnloss_*). */
typedef struct {
  int N;            /* batch                                                          */
  int H, W;         /* spatial extent of the SMALL-stride side input (see transposed) */
  int Cin, ldx;     /* channels consumed from x, pixel stride of x                    */
  int OH, OW;       /* spatial extent of y                                            */
  int Cout, ldy;    /* channels produced, pixel stride of y                           */
  int k, stride, pad;
  int transposed;   /* 0: y = conv(x): iy = oy*stride - pad + kh.  1: y = conv_transpose(x): oy = iy*stride - pad + kh
                       (pad = torch padding; ConvT(p=0)+Cropping2D(1) is pad = 1).  dgrad of one kind is the
                       forward of the other kind with the weight roles swapped. */
  int impl;         /* PTK_IMPL_*                                                     */
} ptk_conv_geom;

/* y = epilogue(conv(x, w) + bias).  w is packed [tap][Cin_pad][Cout_pad] ("WT": transpose==1 packing of
 * a Conv2d weight, transpose==0 packing of a ConvTranspose2d weight) for the SIMT path, and
 * [tap][Cout][Cin] ("WK") for the tensor-core path; pass both (either may be NULL if that path is
 * never selected).  `stats` (may be NULL) receives per-sample {sum, sum of squares} of the conv output
 * (double[N][2], must be zeroed by the caller) -- the first half of the reference's
 * InstanceNorm3d (models/networks.py:159).  If y2 != NULL the NCHW copy of the result is also written
 * (only supported for Cout <= 4). */
/* 1 if the tcgen05 path can serve this geometry */
int ptk_conv_tc_supported(const ptk_conv_geom* g);
int ptk_conv_forward(const ptk_conv_geom* g, const float* x, const float* w_t, const float* w_k,
                     const float* bias, int act, float* y, float* y_nchw, double* stats, void* stream);
/* Same, with a caller-owned scratch of scratch_floats floats (may be NULL / 0).  Layers with too few output tiles to fill
 * the machine split their K loop over several CTAs; with a scratch of at least 2 * N*OH*OW*Cout floats every split stores
 * its partial tile there and one reduction kernel sums them in a fixed order (deterministic, statistics fused); without it
 * the splits accumulate into y with fp32 atomics. */
int ptk_conv_forward_ws(const ptk_conv_geom* g, const float* x, const float* w_t, const float* w_k,
                        const float* bias, int act, float* y, float* y_nchw, double* stats, float* scratch,
                        int64_t scratch_floats, void* stream);
/* dw[tap][A][B] = sum_pixels small[m][a] * big[m*stride+off(tap)][b]; for Conv2d small=dy,big=x
 * (result [tap][Cout][Cin]); for ConvTranspose2d small=x,big=dy (result [tap][Cin][Cout]), with Cin/Cout as given in
 * the geometry (i.e. including channel padding).  dw is scratch owned by the caller and is OVERWRITTEN (the library
 * zero-fills it itself when it needs split-K accumulation).  tcgen05 TF32 path for k4 s2 layers with wide channels,
 * fp32 CUDA-core path otherwise. */
int ptk_conv_wgrad(const ptk_conv_geom* g, const float* x, const float* dy, float* dw, void* stream);
/* Same gradient, split-K without atomics: dw has room for dw_capacity floats (>= one gradient = k*k*Cin*Cout); the
 * kernel may write *nparts partial gradients (each k*k*Cin*Cout floats, back to back) that the caller sums with
 * ptk_unpack_weight_grad_parts -- a fixed summation order, i.e. a deterministic weight gradient. */
int ptk_conv_wgrad_parts(const ptk_conv_geom* g, const float* x, const float* dy, float* dw, int64_t dw_capacity,
                         int* nparts, void* stream);
/* How many partial gradients ptk_conv_wgrad_parts would write for this geometry and scratch capacity (a pure function of
 * its arguments; with dw_capacity == one gradient the answer is always 1, i.e. the kernel writes the final result). */
int ptk_conv_wgrad_plan(const ptk_conv_geom* g, int64_t dw_capacity, int* nparts);
/* dbias[c] += sum_pixels dy[pixel][c] */
/* ---- narrow 3x3 output head (ReLU -> Conv2d(C -> Co<=3, k3, p1) -> Tanh, models/networks.py:227-232) as three 1x1 GEMMs
 * on a 32-column "tap x channel" tensor, so that the C-channel operand is read / written once per pass (csrc/head.cu).
 * wk [32][Cin] / wd [Cin][32]: the two GEMM layouts of the torch weight [Co][Cin][3][3]. */
int ptk_head_pack_weights(const float* w, int Co, int Cin, float* wk, float* wd, void* stream);
/* y = act(bias + sum_tap z[p + off(tap)][tap*Co + co]); z [N,H,W,32]; y written NCHW and/or into an NHWC slice (ldy) */
int ptk_head_shift_add(const float* z, const float* bias, int N, int Co, int H, int W, int act, float* y_nchw,
                       float* y_nhwc, int ldy, void* stream);
/* dzs[q][tap*Co + co] = dz[q - off(tap)][co]; dz [N,H,W,ldz], dzs [N,H,W,32] (columns >= 9*Co zero) */
int ptk_head_shift_gather(const float* dz, int ldz, int N, int Co, int H, int W, float* dzs, void* stream);
/* grad[co][c][tap] (+)= sum_{q<nparts} dwT[q*part_stride + c*32 + tap*Co + co]   (torch layout [Co][Cin][3][3]) */
int ptk_head_wgrad_scatter(const float* dwT, int nparts, int64_t part_stride, int Co, int Cin, float* grad,
                           int accumulate, void* stream);

int ptk_bias_grad(const float* dy, int ld, int64_t pixels, int C, float* dbias, void* stream);

/* ---------------------------------------------------------------- norm (models/networks.py:159,164-172) */
/* stats[n] = {sum, sumsq} over HW*C elements of sample n (double, caller zeroes). */
int ptk_gn_stats(const float* z, int ld, int N, int64_t HW, int C, double* stats, void* stream);
/* y = ((z-mean)*rstd*gamma+beta) * drop[n,c];  out1 = act1(y) [, out2 = act2(y)].  gamma/beta are device
 * scalars (NULL => identity affine AND no normalisation: y = z, used by the bn=False blocks).
 * drop is [N][C] (values 0 or 2, models/networks.py:161) or NULL. eps = 1e-3, biased variance. */
int ptk_gn_apply(const float* z, int ldz, const double* stats, const float* gamma, const float* beta,
                 const float* drop, int N, int64_t HW, int C, float* out1, int ld1, int act1,
                 float* out2, int ld2, int act2, void* stream);
/* Backward.  dy = sum_i g_i * act_i'(a_i) * drop  (i = 1,2; a_i = the stored activated output, NULL =>
 * identity).  Writes dy dense (ld = C) and, when gamma != NULL, sums[n] += {sum dy, sum dy*xhat}. */
int ptk_gn_bwd_reduce(const float* g1, int ldg1, const float* a1, int lda1, int act1,
                      const float* g2, int ldg2, const float* a2, int lda2, int act2,
                      const float* drop, const float* z, int ldz, const double* stats,
                      int N, int64_t HW, int C, float* dy, double* sums, void* stream);
/* dz = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)) in place over dy; dgamma += sum_n S2_n,
 * dbeta += sum_n S1_n. */
int ptk_gn_bwd_apply(float* dy, const float* z, int ldz, const double* stats, const double* sums,
                     const float* gamma, int N, int64_t HW, int C, float* dgamma, float* dbeta,
                     void* stream);

/* ---------------------------------------------------------------- affine warp (utils/pose_transform.py:16-92) */
/* cv2.resize(INTER_LINEAR) of the f64 masks [N,K,H0,W0] to f32 [N,h,w,K] (pose_transform.py:82-87). */
int ptk_mask_pyramid(const double* masks, int N, int K, int H0, int W0, float* out, int h, int w,
                     void* stream);
/* the same for up to four levels in ONE launch (outs[q] = f32 [N,hs[q],ws[q],K]): what one generator pass needs
 * (models/networks.py:205-213 call AffineTransformLayer once per warped skip level, each resizing the masks again) */
int ptk_mask_pyramid_levels(const double* masks, int N, int K, int H0, int W0, float* const* outs, const int* hs,
                            const int* ws, int nlevels, void* stream);
/* y[n,y,x,c] = act(max_k m[n,y,x,k] * bilinear(x[n,:,:,c]; theta_k(y,x))).  warps = raw [N,K,8] rows
 * (first 6 used, :28).  argk = opaque record of the winning part per element (capacity N*h*w*C bytes; the layout is
 * private to the library: bytes, or 4-bit codes on the fast path).  H0,W0 = init_image_size. */
int ptk_warp_forward(const float* x, int ldx, const float* warps, const float* mask_lvl, float* y, int ldy,
                     uint8_t* argk, int N, int C, int h, int w, int K, int H0, int W0, int align_corners,
                     int act, void* stream);
/* dx (dense, ld = C, caller zeroes) += scatter of dy*act'(y)*m*bilinear weights through argk. */
int ptk_warp_backward(const float* dy, int lddy, const float* y, int ldy, int act, const float* warps,
                      const float* mask_lvl, const uint8_t* argk, float* dx, int N, int C, int h, int w,
                      int K, int H0, int W0, int align_corners, void* stream);

/* All warped skip levels of one generator pass in ONE launch (models/networks.py:279-288 calls the layer once per level).
 * Forward uses x, ldx, mask, y, ldy, argk; backward uses dy, lddy, mask, argk, dx (dense, ld = C) and, for act = LeakyReLU,
 * y / ldy.  argk is an opaque winner record owned by the caller (N*h*w*C bytes of capacity per level) that must be passed
 * unchanged from the forward to the backward call.  zero_dx != 0: the library zero-fills every dx first. */
typedef struct {
  const float* x; int ldx;
  const float* mask;
  float* y; int ldy;
  uint8_t* argk;
  const float* dy; int lddy;
  float* dx;
  int C, h, w;
} ptk_warp_level;
int ptk_warp_forward_levels(const ptk_warp_level* levels, int nlevels, const float* warps, int N, int K, int H0, int W0,
                            int act, void* stream);
int ptk_warp_backward_levels(const ptk_warp_level* levels, int nlevels, const float* warps, int N, int K, int H0, int W0,
                             int act, int zero_dx, void* stream);

/* ---------------------------------------------------------------- losses (models/pose_gan.py:90-98,140-199) */
/* logits [rows][J] (pre-sigmoid).  rows < n_true: -mean_j log(sig+1e-7); other rows: -mean_j log(1-sig+1e-7);
 * both scaled by `scale`.  loss[0] += true part, loss[1] += fake part; if dlogits != NULL the gradient of
 * element i is written to dlogits[i*ldd] (ldd = 4 gives the zero-padded 4-channel NHWC layout the dgrad reads). */
int ptk_adv_loss(const float* logits, int rows, int J, int n_true, float scale, float* loss,
                 float* dlogits, int ldd, void* stream);
/* loss[0] += scale*mean|a-b| ; grad = scale*sign(a-b)/n (if non-NULL) */
int ptk_l1_loss(const float* a, const float* b, int64_t n, float scale, float* loss, float* grad,
                void* stream);
/* Fused Feature_Extractor('block1_conv2') (utils/pose_utils.py:320-338, vgg features[0..1] incl. the
 * view-based preprocessing) + nn_loss (models/pose_gan.py:173-199).  pred/target NCHW [N,3,H,W];
 * vgg_w [64][3][3][3], vgg_b [64].  loss[0] += scale * mean_{n,y,x} min_shift sum_c |gt_shift - pred|;
 * argmin[n,y,x] = winning shift (row-major in the area x area window). */
int ptk_nnloss_forward(const float* pred, const float* target, const float* vgg_w, const float* vgg_b,
                       int N, int H, int W, int area, float scale, float* loss, uint8_t* argmin,
                       void* stream);
/* dpred NCHW [N,3,H,W] = d(scale*nn_loss)/dpred (overwrites). */
int ptk_nnloss_backward(const float* pred, const float* target, const float* vgg_w, const float* vgg_b,
                        const uint8_t* argmin, int N, int H, int W, int area, float scale, float* dpred,
                        void* stream);
/* nn_loss on MATERIALISED features (the public method DeformablePose_GAN.nn_loss, models/pose_gan.py:173-199):
 * pred / gt NCHW [N,C,H,W]; loss[0] += scale * mean_{n,y,x} min_shift sum_c |gt_shift - pred| with -10000 padding;
 * argmin[n,y,x] = winning shift.  Backward: dpred = scale * (-sign(gt_shift* - pred)) / (N*H*W) (overwrites). */
int ptk_nnloss_features_forward(const float* pred, const float* gt, int N, int C, int H, int W, int area, float scale,
                                float* loss, uint8_t* argmin, void* stream);
int ptk_nnloss_features_backward(const float* pred, const float* gt, const uint8_t* argmin, int N, int C, int H, int W,
                                 int area, float scale, float* dpred, void* stream);
/* dz[n,y,x,c] (NHWC, ld) = (g_nchw[n,c,y,x] + g_nhwc[(n,y,x)*ldg + c]) * (1 - out[n,c,y,x]^2); either g may be NULL */
int ptk_tanh_bwd_combine(const float* g_nchw, const float* g_nhwc, int ldg, const float* out_nchw,
                         float* dz, int ld, int N, int C, int H, int W, void* stream);

/* ---------------------------------------------------------------- VGG-19 prefix deeper than block1_conv2
 * (utils/pose_utils.py:320-338 Feature_Extractor with content_loss_layer = blockB_convC: the convolutions run through
 * ptk_conv_forward; these are the remaining torchvision ops of the prefix.)
 * ptk_vgg_preprocess: backward == 0: out NHWC [N,H,W,ld] channels 0..2 = preprocess_for_vgg(x NCHW [N,3,H,W]) -- the
 *   reference re-views the NCHW buffer as NHWC, so flat per-sample element i uses mean[i % 3], std[i % 3];
 *   backward != 0: x is the NHWC gradient [N,H,W,ld], out the NCHW gradient [N,3,H,W] (divided by the same std).
 * ptk_maxpool2_*: nn.MaxPool2d(kernel_size=2, stride=2) on NHWC; the backward routes to the first maximum of the window
 *   and writes every element of dx (even H, W required).
 * ptk_relu_backward: dy *= (y > 0) in place from the saved output. */
int ptk_vgg_preprocess(const float* x, float* out, int ld, int N, int H, int W, int backward, void* stream);
int ptk_maxpool2_forward(const float* x, int ldx, float* y, int ldy, int N, int H, int W, int C, void* stream);
int ptk_maxpool2_backward(const float* dy, int lddy, const float* x, int ldx, float* dx, int lddx, int N, int H, int W,
                          int C, void* stream);
int ptk_relu_backward(const float* y, int ldy, float* dy, int lddy, int64_t pixels, int C, void* stream);

/* ---------------------------------------------------------------- device-side data path
 * Replaces the per-sample numpy / skimage work of datasets/PoseTransfer_Dataset.py:78-109 on the host.
 * kp: int32 [N,P,2] key-points as (y, x), -1 = missing (utils/pose_utils.py:42). */
/* out[n, c0+p, y, x] = exp(-((y-ky)^2 + (x-kx)^2) / (2 sigma^2)) (zero plane if missing); out is NCHW with C_total
 * channels (utils/pose_utils.py:79-86 cords_to_map, sigma = 6). */
int ptk_pose_heatmaps(const int* kp, int N, int P, int H, int W, float sigma, float* out, int C_total, int c0, void* stream);
/* masks[N,10,H,W] float64: body (all ones), head rectangle, eight limb quadrilaterals (utils/pose_transform.py:143-214
 * pose_masks / estimate_polygon / mask_from_kp_array; point-in-polygon = skimage.measure.grid_points_in_poly).  P = 16 or 18.
 * The caller must have checked that Rhip, Lhip, Rsho, Lsho are present (the reference raises KeyError otherwise). */
int ptk_pose_masks(const int* kp, int N, int P, int H, int W, double* masks, void* stream);

/* ---------------------------------------------------------------- optimiser (models/pose_gan.py:49-51) */
/* torch.optim.Adam (no weight decay, no amsgrad) on a flat fp32 arena. step >= 1. grad_scale multiplies g
 * first (1/world for data-parallel averaging). */
int ptk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, void* stream);

/* ---------------------------------------------------------------- diagnostics */
/* number of kernels this process has launched through the library (for bench.py's gpu_launches) */
int64_t ptk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PTK_H_ */
