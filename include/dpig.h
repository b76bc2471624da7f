/*
 * dpig.h -- C ABI of libdpig.so: the B200-native hot path of
 * charliememory/Disentangled-Person-Image-Generation (Fg/Bg/Pose encoder-decoder generator and
 * GAN discriminator conv forward/backward, losses, optimiser).
 *
 * The reference has no FFI layer: its hot path is Python calling TensorFlow-1.4 kernels.  Each
 * entry point below therefore names the TensorFlow call site in the reference that it replaces
 * (file:line under the reference tree); INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every entry returns 0 (DPIG_OK) or a negative DPIG_E* code; dpig_last_error() has the text;
 *     nothing throws, nothing allocates user-visible memory, nothing syncs the device;
 *   - all pointers are caller-owned DEVICE pointers (torch.Tensor.data_ptr()) unless a parameter
 *     is documented as host memory;
 *   - activations are NHWC "split-bf16" tensors (struct dpig_tensor): two bf16 planes hi/lo with
 *     x ~= hi + lo (|err| <= 2^-18 |x|); the tensor cores consume the planes directly
 *     (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) which reproduces fp32 convolution to
 *     ~1e-5 relative.  Parameters, parameter gradients, losses, logits, images are plain fp32;
 *   - weights keep TensorFlow's HWIO order ([kh][kw][cin][cout], tflib/ops/conv2d.py:76-80,
 *     slim.conv2d) in their fp32 master copy; dpig_weight_pack() derives the bf16 operand copies;
 *   - one dpig_ctx per (device, host thread); all work is enqueued on the caller's stream.
 */
#ifndef DPIG_H_
#define DPIG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpig_ctx dpig_ctx;
typedef void* dpig_stream; /* cudaStream_t */

enum {
  DPIG_OK = 0,
  DPIG_EINVAL = -1,       /* bad argument / shape */
  DPIG_ECUDA = -2,        /* CUDA runtime or driver error */
  DPIG_EUNSUPPORTED = -3, /* shape outside what the kernels cover */
  DPIG_ENODEVICE = -4     /* no sm_100 device */
};

enum { DPIG_ACT_NONE = 0, DPIG_ACT_RELU = 1, DPIG_ACT_LRELU = 2 };
enum { DPIG_NORM_LAYER = 0, DPIG_NORM_BATCH = 1, DPIG_NORM_INSTANCE = 2 };
enum { DPIG_GAN_DCGAN = 0, DPIG_GAN_WGAN = 1, DPIG_GAN_WGAN_GP = 2, DPIG_GAN_LSGAN = 3 };

/* NHWC split-bf16 activation view.  Pixel (n,y,x) starts at ((n*h + y)*w + x)*pix_stride elements
 * from hi / lo; pix_stride >= c lets a tensor live inside a wider (concat) buffer.
 * lo may be NULL for exactly-representable data (masks, {-1,+1} pose maps). */
typedef struct dpig_tensor {
  void* hi;
  void* lo;
  int32_t n, h, w, c;
  int64_t pix_stride;
} dpig_tensor;

/* Fused epilogue of the tensor-core convolutions:
 *   pre = acc + bias;   v = act(pre) + addend;
 *   mask_out bit(pixel, ch) = pre > 0            (saved for the backward pass)
 *   out        <- v                               (split-bf16, optional)
 *   out_masked <- v * (mask_in bit ? 1 : mask_neg)  (split-bf16, optional; backward of ReLU/LReLU)
 *   out_f32    <- v                               (fp32, optional)
 * upsample=2 replicates every produced pixel into a 2x2 block of out (nearest-neighbour
 * resize fused into the 1x1 conv that follows it in the reference, models.py:569-570). */
typedef struct dpig_conv_epilogue {
  const float* bias;
  int32_t act;
  float alpha;
  const dpig_tensor* addend;
  const uint32_t* mask_in;
  float mask_neg;
  uint32_t* mask_out;
  const dpig_tensor* out;
  const dpig_tensor* out_masked;
  float* out_f32;
  int64_t out_f32_pix_stride;
  int32_t upsample;
  /* optional fp32 [n][9][cout]: extra bias selected by the output pixel's border class
   * (row class 0/1/2 = first/interior/last row) * 3 + column class -- the exact contribution of input channels
   * that are constant over space (the broadcast embedding, trainer.py:588-590), see dpig_stem_class_bias. */
  const float* class_bias;
  /* optional fp32 [cout]: += sum over pixels of out_masked -- the bias gradient of the layer that out_masked is the
   * output-gradient of, fused here so that the gradient tensor is not re-read by dpig_bias_grad. */
  float* colsum_masked;
  /* optional fp64 [2][groups]: raw sums (sum x, sum x^2) of pre = acc + bias, accumulated by the epilogue so that the
   * normalisation that follows the conv (wgan_gp.py:417-431: Conv2D -> Batchnorm / Layernorm -> LeakyReLU) needs no
   * separate statistics pass over the conv output.  stat_mode DPIG_NORM_BATCH: groups = cout (per channel over
   * N,H,W; tflib/ops/batchnorm.py:29-30); DPIG_NORM_LAYER: groups = n (per sample over C,H,W; layernorm.py:6-20).
   * The buffer is zeroed by the call.  Needs act = DPIG_ACT_NONE and no addend / out_masked.  These raw sums are what
   * data-parallel ranks all-reduce for sync-BN before dpig_norm_act_fwd. */
  double* stat_sums;
  int32_t stat_mode;
} dpig_conv_epilogue;

/* ---- context ----------------------------------------------------------------------------- */
int dpig_ctx_create(int device, dpig_ctx** out);
void dpig_ctx_destroy(dpig_ctx* ctx);
const char* dpig_last_error(const dpig_ctx* ctx);
/* fast=1: single bf16 pass on the hi planes only (NOT the parity mode; ~1e-2 relative). */
int dpig_ctx_set_fast_mode(dpig_ctx* ctx, int fast);
/* conv kernels as 2-CTA clusters (tcgen05 cta_group::2, M=256 MMAs over two adjacent pixel tiles, each CTA staging
 * half of the weight rows): 0 never, 1 where it measured faster (default; wide channel blocks and single-K-chunk
 * layers), 2 wherever the shape allows.  Results are identical in every mode (same products, same fp32 accumulation
 * order per output element). */
int dpig_ctx_set_pair_mode(dpig_ctx* ctx, int mode);
/* Tuning switches by name ("epi_specialise", "wgrad_split", "wide_b", ... = the DPIG_* environment variables read at
 * context creation); results are identical under every setting.  For A/B measurements and tests. */
int dpig_ctx_set_option(dpig_ctx* ctx, const char* name, int value);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches). */
unsigned long long dpig_launch_count(const dpig_ctx* ctx);
const char* dpig_version(void);

/* ---- convolution (replaces slim.conv2d: models.py:396-399,425-429,458-462,528-539,564-573 and
 *      tflib/ops/conv2d.py:106-120 tf.nn.conv2d(..., padding='SAME') + bias_add) --------------- */

/* fp32 HWIO master weights -> bf16 operand copies.
 *   fwd_*: [kh*kw][cout][cin_pad]   (K = cin contiguous; forward B operand)
 *   bwd_*: [kh*kw][cin][cout_pad]   (K = cout contiguous; data-gradient B operand)
 * cin_pad / cout_pad are the channel counts rounded up to a multiple of 8 (zero filled). */
int dpig_weight_pack(dpig_ctx* ctx, const float* w_hwio, int32_t taps, int32_t cin, int32_t cout,
                     int32_t cin_pad, int32_t cout_pad, void* fwd_hi, void* fwd_lo, void* bwd_hi,
                     void* bwd_lo, dpig_stream stream);

/* y = epilogue(conv2d_SAME(x, w, stride)).  x->c must equal cin_pad. k in {1,3,5}, stride in {1,2}. */
int dpig_conv2d_fwd(dpig_ctx* ctx, const dpig_tensor* x, const void* wf_hi, const void* wf_lo,
                    int32_t kh, int32_t kw, int32_t stride, int32_t cout,
                    const dpig_conv_epilogue* ep, dpig_stream stream);

/* dx = epilogue(conv2d_backprop_input(dy, w)) (tf.gradients through tf.nn.conv2d; trainer.py:622-625).
 * dy->c must equal cout_pad; dx is [dy->n, in_h, in_w, cin]. */
int dpig_conv2d_bwd_data(dpig_ctx* ctx, const dpig_tensor* dy, const void* wb_hi,
                         const void* wb_lo, int32_t kh, int32_t kw, int32_t stride, int32_t in_h,
                         int32_t in_w, int32_t cin, const dpig_conv_epilogue* ep,
                         dpig_stream stream);

/* dw[kh][kw][ci][co] += sum_pixels x[...]*dy[...]  (conv2d_backprop_filter).  dw is fp32 HWIO with
 * logical sizes cin x cout (rows/cols beyond are padding and are not touched). */
int dpig_conv2d_bwd_filter(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh,
                           int32_t kw, int32_t stride, int32_t cin, int32_t cout, float* dw,
                           dpig_stream stream);

/* Same as dpig_conv2d_bwd_filter for a ROW SLICE of a wider filter: dw_rows points at row ci0 of tap 0 of an
 * HWIO gradient whose taps are cin_total rows apart; rows [0,cin) of the slice are accumulated. */
int dpig_conv2d_bwd_filter_rows(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh,
                                int32_t kw, int32_t stride, int32_t cin, int32_t cout, float* dw_rows,
                                int32_t cin_total, dpig_stream stream);
/* dpig_weight_pack for input-channel rows [ci0, ci0+cin) of a filter with cin_total rows per tap. */
int dpig_weight_pack_rows(dpig_ctx* ctx, const float* w_hwio, int32_t taps, int32_t cin_total, int32_t ci0,
                          int32_t cin, int32_t cout, int32_t cin_pad, int32_t cout_pad, void* fwd_hi,
                          void* fwd_lo, void* bwd_hi, void* bwd_lo, dpig_stream stream);

/* Stem shortcut for spatially constant input channels (the tiled embedding of trainer.py:588-590 entering the
 * 3x3 SAME stem conv models.py:528): with E[tap][n][co] = sum_ci emb[n,ci] * W[tap,ci,co] (9 small GEMMs),
 *   class_bias[n][class][co] = sum over the taps that fall inside the image for that border class of E,
 * which dpig_conv2d_fwd adds per pixel (dpig_conv_epilogue.class_bias): exact, 3.5 GMAC/image cheaper.
 * Backward: class_sums[n][class][co] = sum of g over the pixels of each border class;
 *   tap_sums[tap][n][co] = sum over classes where the tap is inside = sum_pixels valid(tap) g  -> dW, d_emb. */
int dpig_stem_class_bias(dpig_ctx* ctx, const float* e_taps /* [9][n][cout] */, int32_t n, int32_t cout,
                         int32_t h, int32_t w_, float* class_bias /* [n][9][cout] */, dpig_stream stream);
int dpig_stem_tap_sums(dpig_ctx* ctx, const dpig_tensor* g, float* class_sums /* [n][9][c] workspace */,
                       float* tap_sums /* [9][n][c] */, dpig_stream stream);

/* CUDA-core fp32 convolutions for the 3-channel ends of the networks (image in, image out):
 * x, y, dy, dx are plain fp32 NHWC here; w is the fp32 HWIO master. */
int dpig_conv2d_small_fwd(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_,
                          int32_t cin, const float* w, const float* bias, int32_t kh, int32_t kw,
                          int32_t stride, int32_t cout, int32_t act, float alpha,
                          const dpig_tensor* out, float* out_f32, uint32_t* mask_out,
                          dpig_stream stream);
int dpig_conv2d_small_bwd_data(dpig_ctx* ctx, const float* dy, int32_t n, int32_t oh, int32_t ow,
                               int32_t cout, const float* w, int32_t kh, int32_t kw,
                               int32_t stride, int32_t in_h, int32_t in_w, int32_t cin, float* dx,
                               dpig_stream stream);
int dpig_conv2d_small_bwd_filter(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_,
                                 int32_t cin, const float* dy, int32_t kh, int32_t kw,
                                 int32_t stride, int32_t cout, float* dw, dpig_stream stream);

/* The 3-channel ends as dense 1x1 contractions (patch.cu).  (tap, channel) pairs of the 3-channel side become channels:
 *   dpig_im2col_small:  out[n,oy,ox,(i*kw+j)*c_src + c] = src[n, oy*s+i-pt, ox*s+j-pl, c]  (0 outside; SAME padding);
 *                       transposed=1 (stride 1): src[n, oy-(i-pt), ox-(j-pl), c] -- the patches of an output gradient.
 *                       c_src leading channels of src are used; out->c >= kh*kw*c_src, remaining channels are zero.
 *     cin = 3 layers (models.py:396; wgan_gp.py:419):  conv = dpig_conv2d_fwd(patches, k=1) on the same HWIO filter read
 *     as [kh*kw*3][cout];  filter gradient = dpig_conv2d_bwd_filter(patches, dy, k=1).
 *   dpig_col2im_small:  out[n,y,x,co] = bias[co] + sum_taps Y[n, y+i-pt, x+j-pl, (i*kw+j)*cout + co] for the cout = 3
 *     output conv (models.py:573) computed as Y = dpig_conv2d_fwd(x, k=1) with the filter re-laid by
 *   dpig_permute_taps:  mode 0 dst[ci][tap*cout+co] = src[tap][ci][co]; mode 1 dst[tap*cout+co][ci] = src[tap][ci][co];
 *                       mode 2 dst[tap][ci][co] += src[ci][tap*cout+co]  (fp32 filters / filter gradients). */
int dpig_im2col_small(dpig_ctx* ctx, const dpig_tensor* src, int32_t c_src, int32_t kh, int32_t kw, int32_t stride,
                      int32_t transposed, const dpig_tensor* out, dpig_stream stream);
int dpig_col2im_small(dpig_ctx* ctx, const float* y, int64_t y_pix_stride, int32_t n, int32_t h, int32_t w_, int32_t kh,
                      int32_t kw, int32_t cout, const float* bias, float* out_f32, int64_t out_f32_pix_stride,
                      const dpig_tensor* out, dpig_stream stream);
int dpig_permute_taps(dpig_ctx* ctx, const float* src, float* dst, int32_t taps, int32_t cin, int32_t cout, int32_t mode,
                      dpig_stream stream);

/* db[c] += sum over pixels of dy (bias gradient of bias_add). */
int dpig_bias_grad(dpig_ctx* ctx, const dpig_tensor* dy, float* db, dpig_stream stream);
int dpig_bias_grad_f32(dpig_ctx* ctx, const float* dy, int64_t pixels, int32_t c, float* db,
                       dpig_stream stream);

/* ---- element-wise glue on split tensors ----------------------------------------------------- */
/* out = (a + b + c + f32) [2x2 sum-pooled if pool2] * (mask bit ? 1 : mask_neg); any input may be
 * NULL.  Inputs are addressed at out's pixel grid (or the 2x finer grid when pool2). */
int dpig_ew_combine(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a,
                    const dpig_tensor* b, const dpig_tensor* c, const float* f32,
                    int64_t f32_pix_stride, const uint32_t* mask, float mask_neg, int32_t pool2,
                    dpig_stream stream);
/* dpig_ew_combine that also accumulates colsum[c] += sum over pixels of out[., c] (fp32 [out->c]): the bias gradient of
 * the conv whose output gradient `out` is (tf.gradients of bias_add), without re-reading the tensor (dpig_bias_grad). */
int dpig_ew_combine_colsum(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a,
                           const dpig_tensor* b, const dpig_tensor* c, const float* f32,
                           int64_t f32_pix_stride, const uint32_t* mask, float mask_neg,
                           int32_t pool2, float* colsum, dpig_stream stream);
/* fp32 NHWC -> split (optionally into a channel slice of a wider buffer); pads channels >= c_src
 * with zeros up to out->c. */
int dpig_pack_f32(dpig_ctx* ctx, const float* src, int64_t src_pix_stride, int32_t c_src,
                  const dpig_tensor* out, dpig_stream stream);
/* split -> fp32 NHWC */
int dpig_unpack_f32(dpig_ctx* ctx, const dpig_tensor* src, float* dst, int64_t dst_pix_stride,
                    dpig_stream stream);
/* x_fg = x*m, x_bg = x*(1-m) (models.py:402-403); m is fp32 [n,h,w]. Either output may be NULL. */
int dpig_mask_split(dpig_ctx* ctx, const dpig_tensor* x, const float* m, const dpig_tensor* fg,
                    const dpig_tensor* bg, dpig_stream stream);
/* out[n,y,x,c0:c0+ce] = emb[n,:] broadcast over space (trainer.py:588-590), written into a channel
 * slice of the generator's input buffer. */
int dpig_broadcast_embedding(dpig_ctx* ctx, const float* emb, int32_t ce, const dpig_tensor* out,
                             dpig_stream stream);
/* g_emb[n,c] = sum over pixels of g[n,y,x,c] (gradient of the broadcast). */
int dpig_spatial_sum(dpig_ctx* ctx, const dpig_tensor* g, float* out, dpig_stream stream);

/* emb[b, i*part_z + j] = fea[i*batch + b, j] * vis[b, i]  (i < parts), emb[b, parts*part_z + j] = bg[b, j]
 * (models.py:433-442, 467-468).  backward=1: fea/bg receive the gradient held in emb. */
int dpig_embedding_assemble(dpig_ctx* ctx, float* fea, float* bg, const float* vis, int32_t batch,
                            int32_t parts, int32_t part_z, int32_t bg_z, float* emb, int32_t backward,
                            dpig_stream stream);

/* ---- tf.image.crop_and_resize (models.py:350,415), bilinear, extrapolation 0 ------------------ */
/* boxes: fp32 [nbox][4] = (y1,x1,y2,x2) normalised as in the reference (pixel / H, pixel / W);
 * box_ind: int32 [nbox]; mask (optional fp32 [n,h,w]) multiplies the image on the fly (x_fg). */
int dpig_crop_and_resize_fwd(dpig_ctx* ctx, const dpig_tensor* image, const float* mask,
                             const float* boxes, const int32_t* box_ind, int32_t nbox,
                             const dpig_tensor* out, dpig_stream stream);
/* grad_image (fp32 NHWC, dense) = CropAndResizeGradImage(grad) (* mask).  Gather form: every image pixel sums the crop
 * samples whose bilinear footprint covers it, in a fixed order -- no atomics, run-to-run identical, grad_image is
 * overwritten.  (DPIG_CROP_GATHER=0 selects the older atomic scatter form, which ACCUMULATES into a caller-zeroed
 * grad_image.) */
int dpig_crop_and_resize_bwd(dpig_ctx* ctx, const dpig_tensor* grad, const float* mask,
                             const float* boxes, const int32_t* box_ind, int32_t nbox,
                             float* grad_image, int32_t n, int32_t h, int32_t w_, int32_t c,
                             dpig_stream stream);

/* ---- fully connected (slim.fully_connected models.py:431,464,478-484,545,554;
 *      tflib/ops/linear.py:133-147) -- fp32 CUDA-core GEMMs, y[m,n] = x[m,k] w[k,n] + b -------- */
int dpig_linear_fwd(dpig_ctx* ctx, const float* x, const float* w, const float* b, float* y,
                    int32_t m, int32_t k, int32_t n, int32_t act, float alpha, dpig_stream stream);
/* dx[m,k] = dy[m,n] w[k,n]^T ; dw[k,n] += x^T dy ; db[n] += sum_m dy */
int dpig_linear_bwd(dpig_ctx* ctx, const float* x, const float* w, const float* dy, float* dx,
                    float* dw, float* db, int32_t m, int32_t k, int32_t n, dpig_stream stream);
/* out = sa*a + sb*b (b may be NULL): residual adds of the FC nets (models.py:483, 496, 509) */
int dpig_add_f32(dpig_ctx* ctx, float* out, const float* a, const float* b, int64_t count, float sa, float sb,
                 dpig_stream stream);
/* dy *= (y > 0 ? 1 : neg) in place, for activated linear layers */
int dpig_act_bwd_f32(dpig_ctx* ctx, const float* y, float* dy, int64_t count, float neg,
                     dpig_stream stream);

/* ---- discriminator normalisation + LeakyReLU (wgan_gp.py:34-40,407-440;
 *      tflib/ops/layernorm.py:6-20; tflib/ops/batchnorm.py:29-30) ----------------------------- */
/* x: fp32 NHWC [n,h,w,c] pre-norm conv output.
 * sums : fp64 [2][groups] raw (sum x, sum x^2); groups = n (LAYER), c (BATCH), n*c (INSTANCE).
 *   Data-parallel BATCH mode all-reduces `sums` between dpig_norm_stats and dpig_norm_act_fwd and
 *   passes the global element count -- the sync-BN hook.
 * stats: fp32 [2][groups] = (mean, rstd), written by dpig_norm_act_fwd, kept for the backward pass.
 * y = lrelu((x-mean)*rstd*scale[c] + offset[c]); emitted as split tensor + sign bitmask. */
int dpig_norm_stats(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_, int32_t c,
                    int32_t mode, double* sums /* fp64 [2][groups] raw (sum x, sum x^2) */,
                    dpig_stream stream);
int dpig_norm_act_fwd(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_, int32_t c,
                      int32_t mode, float eps, const double* sums, double count,
                      const float* scale, const float* offset, int32_t act, float alpha,
                      float* stats, const dpig_tensor* out, uint32_t* mask_out,
                      dpig_stream stream);
/* backward: given dy (split, grad wrt activated output), x, stats, mask: computes
 *   dscale[c] +=, doffset[c] +=, and dx (fp32) wrt the pre-norm conv output.
 * red: fp64 [2][groups] workspace for (sum dyhat, sum dyhat*xhat); two-phase so that BATCH mode can
 * allreduce `red` between the phases. */
int dpig_norm_act_bwd_reduce(dpig_ctx* ctx, const dpig_tensor* dy, const float* x,
                             const float* stats, const uint32_t* mask, float alpha, int32_t mode,
                             const float* scale, double* red, float* dscale, float* doffset,
                             dpig_stream stream);
int dpig_norm_act_bwd_apply(dpig_ctx* ctx, const dpig_tensor* dy, const float* x,
                            const float* stats, const uint32_t* mask, float alpha, int32_t mode,
                            const float* scale, const double* red, double count,
                            const dpig_tensor* dx, dpig_stream stream);

/* Second-order LayerNorm pieces of the WGAN-GP penalty (trainer.py:226-236): the parameter gradient of the
 * penalty is the gradient of the directional derivative (JVP) of D along v = d(lambda*gp)/d(grad).  Conv,
 * LeakyReLU and Linear are piecewise linear -- their JVP / adjoint reuse the entries above -- LayerNorm is not.
 *   jvp_fwd: hdot = lrelu'(mask) * gamma*rstd*(pdot - mean(pdot) - xhat*mean(xhat*pdot))
 *            p, pdot: fp32 NHWC primal / tangent pre-norm conv outputs; stats from dpig_norm_act_fwd (LAYER);
 *            tsums: fp64 [2][n] workspace, kept for jvp_bwd.
 *   jvp_bwd: given hdot_bar, returns pdot_bar (split; adjoint of the tangent input), p_bar (fp32; adjoint of the
 *            primal pre-norm input through xhat and sigma) and accumulates dscale.  asums: fp64 [3][n] workspace. */
int dpig_layernorm_jvp_fwd(dpig_ctx* ctx, const float* p, const float* pdot, int32_t n, int32_t h, int32_t w_,
                           int32_t c, const float* stats, const float* scale, const uint32_t* mask, float alpha,
                           double* tsums, const dpig_tensor* hdot, dpig_stream stream);
int dpig_layernorm_jvp_bwd(dpig_ctx* ctx, const dpig_tensor* hdot_bar, const uint32_t* mask, float alpha,
                           const float* p, const float* pdot, const float* stats, const float* scale,
                           const double* tsums, double* asums, float* dscale, const dpig_tensor* pdot_bar,
                           float* p_bar, dpig_stream stream);

/* ---- losses (trainer.py:217-252, 606-607) ---------------------------------------------------- */
/* out[0] = mean|g-x|;  if dg != NULL: dg += weight * sign(g-x)/count  (fp32 images) */
int dpig_loss_l1(dpig_ctx* ctx, const float* g, const float* x, int64_t count, float weight,
                 float* out, float* dg, dpig_stream stream);
/* GAN losses on logits.  out[0]=g_loss, out[1]=d_loss.  d_fake_g: dL_g/d(fake logits);
 * d_real_d, d_fake_d: dL_d/d(real/fake logits).  Any gradient pointer may be NULL. */
int dpig_loss_gan(dpig_ctx* ctx, int32_t mode, const float* d_real, const float* d_fake,
                  int32_t count, float* out, float* d_fake_g, float* d_real_d, float* d_fake_d,
                  dpig_stream stream);
/* Pose auto-encoder loss of --model=2 (trainer.py:638-660): vis = binaryRound(sigmoid(vis_logit)) (models.py:97-108,
 * 512-513; straight-through gradient), G_rcv[b,k] = (coord[b,2k], coord[b,2k+1], vis[b,k]) (written to g_rcv, optional),
 * out[0] = mean((target - G_rcv)^2) over batch*keypoints*3; d_coord / d_vis_logit (optional) receive
 * weight * d(out)/d(.) (the trainer minimises reconstruct_loss * 20, trainer.py:663). */
int dpig_pose_ae_loss(dpig_ctx* ctx, const float* target, const float* coord, const float* vis_logit, int32_t batch,
                      int32_t keypoints, float weight, float* out, float* d_coord, float* d_vis_logit, float* g_rcv,
                      dpig_stream stream);
/* WGAN-GP pieces (trainer.py:226-236).  xhat = x + alpha[n]*(g - x) */
int dpig_gp_interpolate(dpig_ctx* ctx, const float* x, const float* g, const float* alpha,
                        int32_t n, int64_t per_sample, float* xhat, dpig_stream stream);
/* slopes[n] = sqrt(sum grad^2); out[0] = mean((slopes-1)^2);
 * dgrad[n,:] = lambda * 2*(slope-1)/(n*slope) * grad  (the seed of the second backward pass) */
int dpig_gp_penalty(dpig_ctx* ctx, const float* grad, int32_t n, int64_t per_sample, float lambda,
                    float* slopes, float* out, float* dgrad, dpig_stream stream);

/* ---- optimisers (trainer.py:116-149; TF formulas) --------------------------------------------- */
/* TF AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v update; p -= lr_t*m/(sqrt(v)+eps).
 * grad_scale multiplies g first (1/world_size after a sum-allreduce). */
int dpig_adam_step(dpig_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count,
                   float lr, float beta1, float beta2, float eps, int32_t t, float grad_scale,
                   dpig_stream stream);
/* TF RMSPropOptimizer (decay .9, momentum 0, eps 1e-10, ms initialised to ones) + optional clip */
int dpig_rmsprop_step(dpig_ctx* ctx, float* p, const float* g, float* ms, int64_t count, float lr,
                      float decay, float eps, float grad_scale, float clip, dpig_stream stream);
/* Graph-replayable forms: the step size is read from device memory at run time (lr_t_dev[0] = the bias-corrected
 * Adam step lr*sqrt(1-b2^t)/(1-b1^t) of trainer.py:119-123 computed by the host; lr_dev[0] = the RMSProp lr), so one
 * captured CUDA graph of "forward + backward + update" serves every step while t and the halved lr change. */
int dpig_adam_step_dev(dpig_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count,
                       const float* lr_t_dev, float beta1, float beta2, float eps, float grad_scale,
                       dpig_stream stream);
int dpig_rmsprop_step_dev(dpig_ctx* ctx, float* p, const float* g, float* ms, int64_t count,
                          const float* lr_dev, float decay, float eps, float grad_scale, float clip,
                          dpig_stream stream);
int dpig_clip(dpig_ctx* ctx, float* p, int64_t count, float lo, float hi, dpig_stream stream);

/* ---- image / pose ends (utils.py:88-89, 259-318) ---------------------------------------------- */
/* u8 = clip((g+1)*127.5, 0, 255) */
int dpig_denorm_u8(dpig_ctx* ctx, const float* g, int64_t count, uint8_t* out,
                   dpig_stream stream);
/* SSIM of generate() (trainer.py:514-526, tester.py:236-241) on the device: a, b uint8 [n,h,w,3] (a = the generated
 * image, b = the input image); out[i] = skimage.measure.compare_ssim(rgb2gray(a[i]), rgb2gray(b[i]),
 * data_range = gray(b[i]).max() - gray(b[i]).min()) with scikit-image's defaults (7x7 uniform window, K1 .01, K2 .03,
 * sample covariance, float64). */
int dpig_ssim_gray_u8(dpig_ctx* ctx, const uint8_t* a, const uint8_t* b, int32_t n, int32_t h, int32_t w,
                      float* out, dpig_stream stream);
/* rcv: fp32 [n][k][3] (row, col, visible) -> {-1,+1} maps [n,h,w,k] dilated by the radius-4 disc
 * of tf_poseInflate; written as a split tensor slice (hi plane only is needed). */
int dpig_pose_rasterize(dpig_ctx* ctx, const float* rcv, int32_t n, int32_t k, int32_t h,
                        int32_t w_, int32_t radius, const dpig_tensor* out, float* out_f32,
                        dpig_stream stream);
/* The kh x kw SAME-padded (stride 1) patches of those maps, built straight from the keypoints:
 * out[n,y,x,(i*kw + j)*k + c] = map_c(y + i - pad, x + j - pad), 0 outside the image and in the pad channels
 * (out->c >= kh*kw*k).  Input of the U-Net stem's pose rows in patch form (models.py:520-528: the 18 pose channels of
 * concat([emb, pose]) under the 3x3 stem conv, as ONE 1x1 contraction over kh*kw*k channels). */
int dpig_pose_patch(dpig_ctx* ctx, const float* rcv, int32_t n, int32_t k, int32_t h, int32_t w_,
                    int32_t radius, int32_t kh, int32_t kw, const dpig_tensor* out, dpig_stream stream);

/* ---- host utility ------------------------------------------------------------------------------ */
/* CRC-32C (Castagnoli) of a HOST buffer, chained through `crc` (start with 0): the checksum TensorFlow's checkpoint
 * format stores per tensor and per index block (tf.train.Saver, reference trainer.py:180-213, 365-366); used by the
 * TensorFlow-free checkpoint reader / writer tf_checkpoint.py.  No context, no device work. */
uint32_t dpig_crc32c(uint32_t crc, const void* data, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* DPIG_H_ */
