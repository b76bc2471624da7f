// The 3-channel ends of the networks (image in: slim.conv2d models.py:396 / tflib Conv2D wgan_gp.py:419; image out:
// models.py:573) as dense 1x1 contractions.
//
// The implicit-GEMM conv kernel spends a whole 64-channel K chunk per filter tap on 3 (padded to 8) valid channels
// and, for the 256 -> 3 output conv, re-reads the 256-channel operand from L2 once per tap for a 32-column MMA:
// 4-10 TFLOP/s, 6 % of a Stage-I iteration.  Moving the (tap, channel) pairs of the 3-channel side into the channel
// dimension removes both:
//   cin = 3  : P = im2col(x)  [pixels, k*k*3 -> pad 32 / 80], then y = P * W as a 1x1 conv with ONE K chunk (the HWIO
//              filter [k][k][3][cout] already is the [k*k*3][cout] matrix); the filter gradient is the 1x1 wgrad P^T dy.
//   cout = 3 : Y = x * W' as a 1x1 conv to k*k*3 = 27 "tap channels" (x is read once), then out[p] = bias +
//              sum_tap Y[p + shift(tap)][tap] (col2im gather); data / filter gradients use DP = im2col^T(dy):
//              dx = DP * W'' (1x1, one K chunk) and dW' = x^T DP (1x1 wgrad).
// These kernels are the HBM-bound gather / scatter halves; the contractions run on the tensor-core conv kernels.
#include "common.cuh"

namespace dpig {

// out[n, oy, ox, (i*kw + j)*cs + c] = src[n, oy*s + i - pt, ox*s + j - pl, c]   (transposed: y - (i - pt), x - (j - pl))
// One thread per (pixel, group of 8 output channels): 16-byte stores per plane.
// CS / KW > 0: compile-time source channel count / filter width (k / cs, tap / kw become multiply-shifts; with run-time
// divisors the kernel is bound by its ~16 integer divisions per thread, not by memory).
template <int CS, int KW>
__global__ void im2col_small_kernel(const __nv_bfloat16* shi, const __nv_bfloat16* slo, long long sps, int N, int H,
                                    int W, int cs_rt, int kh, int kw_rt, int stride, int pt, int pl, int transposed, int OH,
                                    int OW, __nv_bfloat16* ohi, __nv_bfloat16* olo, long long ops, int Kp) {
  const int cs = CS > 0 ? CS : cs_rt;
  const int kw = KW > 0 ? KW : kw_rt;
  const int K8 = Kp / 8;
  const int kvalid = kh * kw * cs;
  const long long total = static_cast<long long>(N) * OH * OW * K8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % K8);
    const long long pix = idx / K8;
    const int ox = static_cast<int>(pix % OW);
    const int oy = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      uint16_t hv = 0, lv = 0;
      if (k < kvalid) {
        const int tap = k / cs, c = k % cs;
        const int i = tap / kw, j = tap % kw;
        const int iy = transposed ? oy - (i - pt) : oy * stride + i - pt;
        const int ix = transposed ? ox - (j - pl) : ox * stride + j - pl;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const long long off = ((static_cast<long long>(n) * H + iy) * W + ix) * sps + c;
          hv = __bfloat16_as_ushort(shi[off]);
          if (slo) lv = __bfloat16_as_ushort(slo[off]);
        }
      }
      h[e >> 1] |= static_cast<uint32_t>(hv) << ((e & 1) * 16);
      l[e >> 1] |= static_cast<uint32_t>(lv) << ((e & 1) * 16);
    }
    const long long o = pix * ops + g * 8;
    *reinterpret_cast<uint4*>(ohi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (olo) *reinterpret_cast<uint4*>(olo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// out[n, y, x, co] = bias[co] + sum_{i,j} Y[n, y + i - pt, x + j - pl, (i*kw + j)*cout + co]   (zero outside the image)
__global__ void col2im_small_kernel(const float* y, long long yps, int N, int H, int W, int kh, int kw, int pt, int pl,
                                    int cout, const float* bias, float* out_f32, long long fps, __nv_bfloat16* ohi,
                                    __nv_bfloat16* olo, long long ops, int opad) {
  const long long total = static_cast<long long>(N) * H * W;
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const long long nbase = pix - static_cast<long long>(yy) * W - x;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (bias && c < cout) ? __ldg(bias + c) : 0.f;
    for (int i = 0; i < kh; ++i) {
      const int iy = yy + i - pt;
      if (iy < 0 || iy >= H) continue;
      for (int j = 0; j < kw; ++j) {
        const int ix = x + j - pl;
        if (ix < 0 || ix >= W) continue;
        const float* src = y + (nbase + static_cast<long long>(iy) * W + ix) * yps + (i * kw + j) * cout;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < cout) acc[c] += __ldg(src + c);
      }
    }
    if (out_f32)
      for (int c = 0; c < cout; ++c) out_f32[pix * fps + c] = acc[c];
    if (ohi) {
      for (int c = 0; c < opad; ++c) {
        __nv_bfloat16 hv, lv;
        split_bf16(c < cout ? acc[c] : 0.f, hv, lv);
        ohi[pix * ops + c] = hv;
        if (olo) olo[pix * ops + c] = lv;
      }
    }
  }
}

// Filter re-layouts between HWIO [tap][ci][co] and the two 1x1 forms of the cout = 3 convolution:
//   mode 0: dst[ci][tap*cout + co]  = src[tap][ci][co]      (forward weights, 1x1 conv cin -> taps*cout)
//   mode 1: dst[tap*cout + co][ci]  = src[tap][ci][co]      (data-gradient weights, 1x1 conv taps*cout -> cin)
//   mode 2: dst[tap][ci][co]       += src[ci][tap*cout + co] (filter gradient back to HWIO)
__global__ void permute_taps_kernel(const float* src, float* dst, int taps, int cin, int cout, int mode) {
  const int total = taps * cin * cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % cout, ci = (idx / cout) % cin, tap = idx / (cout * cin);   // idx walks HWIO
    const int j = tap * cout + co;
    if (mode == 0) dst[static_cast<long long>(ci) * taps * cout + j] = src[idx];
    else if (mode == 1) dst[static_cast<long long>(j) * cin + ci] = src[idx];
    else dst[idx] += src[static_cast<long long>(ci) * taps * cout + j];
  }
}

static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  return static_cast<int>(g > cap ? cap : g);
}

}  // namespace dpig

using namespace dpig;

extern "C" int dpig_im2col_small(dpig_ctx* ctx, const dpig_tensor* src, int32_t c_src, int32_t kh, int32_t kw,
                                 int32_t stride, int32_t transposed, const dpig_tensor* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!src || !out || !src->hi || !out->hi) return set_error(ctx, DPIG_EINVAL, "im2col_small: null argument");
  if (c_src < 1 || c_src > src->c || kh * kw * c_src > out->c || out->c % 8 || out->pix_stride % 8)
    return set_error(ctx, DPIG_EINVAL, "im2col_small: %d taps x %d channels do not fit %d patch channels", kh * kw, c_src,
                     out->c);
  if (transposed && stride != 1) return set_error(ctx, DPIG_EUNSUPPORTED, "im2col_small: transposed needs stride 1");
  const int OH = same_out(src->h, stride), OW = same_out(src->w, stride);
  if (out->n != src->n || out->h != OH || out->w != OW)
    return set_error(ctx, DPIG_EINVAL, "im2col_small: output is not [%d,%d,%d,.]", src->n, OH, OW);
  const int pt = same_pad_before(src->h, kh, stride), pl = same_pad_before(src->w, kw, stride);
  const long long total = static_cast<long long>(out->n) * OH * OW * (out->c / 8);
  auto launch = [&](auto kern) {
    kern<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src->hi), static_cast<const __nv_bfloat16*>(src->lo), src->pix_stride, src->n,
        src->h, src->w, c_src, kh, kw, stride, pt, pl, transposed, OH, OW, static_cast<__nv_bfloat16*>(out->hi),
        static_cast<__nv_bfloat16*>(out->lo), out->pix_stride, out->c);
  };
  if (c_src == 3 && kw == 3) launch(im2col_small_kernel<3, 3>);        // the image under the 3x3 stems / output conv
  else if (c_src == 3 && kw == 5) launch(im2col_small_kernel<3, 5>);   // the image under the critic's 5x5 first layer
  else launch(im2col_small_kernel<0, 0>);
  ctx->launches++;
  return check_launch(ctx, "im2col_small");
}

extern "C" int dpig_col2im_small(dpig_ctx* ctx, const float* y, int64_t y_pix_stride, int32_t n, int32_t h, int32_t w_,
                                 int32_t kh, int32_t kw, int32_t cout, const float* bias, float* out_f32,
                                 int64_t out_f32_pix_stride, const dpig_tensor* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!y || (!out_f32 && !out)) return set_error(ctx, DPIG_EINVAL, "col2im_small: null argument");
  if (cout < 1 || cout > 8 || kh * kw * cout > y_pix_stride)
    return set_error(ctx, DPIG_EUNSUPPORTED, "col2im_small: cout=%d (max 8), %d tap channels in stride %lld", cout,
                     kh * kw * cout, static_cast<long long>(y_pix_stride));
  if (out && (out->n != n || out->h != h || out->w != w_ || out->c < cout || out->c > 8))
    return set_error(ctx, DPIG_EINVAL, "col2im_small: split output shape mismatch");
  const int pt = same_pad_before(h, kh, 1), pl = same_pad_before(w_, kw, 1);
  const long long total = static_cast<long long>(n) * h * w_;
  col2im_small_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, y_pix_stride, n, h, w_, kh, kw, pt, pl, cout, bias, out_f32, out_f32_pix_stride,
      out ? static_cast<__nv_bfloat16*>(out->hi) : nullptr, out ? static_cast<__nv_bfloat16*>(out->lo) : nullptr,
      out ? out->pix_stride : 0, out ? out->c : 0);
  ctx->launches++;
  return check_launch(ctx, "col2im_small");
}

extern "C" int dpig_permute_taps(dpig_ctx* ctx, const float* src, float* dst, int32_t taps, int32_t cin, int32_t cout,
                                 int32_t mode, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!src || !dst || mode < 0 || mode > 2) return set_error(ctx, DPIG_EINVAL, "permute_taps: bad argument");
  permute_taps_kernel<<<grid_for(static_cast<long long>(taps) * cin * cout), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, taps, cin, cout, mode);
  ctx->launches++;
  return check_launch(ctx, "permute_taps");
}
