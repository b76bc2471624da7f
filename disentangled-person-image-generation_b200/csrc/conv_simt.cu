// CUDA-core fp32 convolutions for the 3-channel ends of the networks: the encoder stem
// (3->128, models.py:396), the discriminator's first layer (3->64 5x5/s2, wgan_gp.py:415) and
// their gradients.  K = kh*kw*3 is far too small for the tensor-core path; these are HBM-bound.
// Weights are the fp32 HWIO masters, images are fp32 NHWC.
#include "common.cuh"

namespace dpig {

// thread = (pixel, co); co fastest so weight reads coalesce and image reads broadcast.
__global__ void conv_small_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int Cin,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      int KH, int KW, int stride, int pt, int pl, int OH, int OW,
                                      int Cout, int act, float alpha, __nv_bfloat16* ohi,
                                      __nv_bfloat16* olo, long long ops, float* of32, uint32_t* mask_out) {
  const long long total = static_cast<long long>(N) * OH * OW * Cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    const long long pix = i / Cout;
    const int ox = static_cast<int>(pix % OW);
    const int oy = static_cast<int>((pix / OW) % OH);
    const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
    float acc = bias ? __ldg(bias + co) : 0.f;
    for (int ky = 0; ky < KH; ++ky) {
      const int iy = oy * stride + ky - pt;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int ix = ox * stride + kx - pl;
        if (ix < 0 || ix >= W) continue;
        const float* xp = x + ((static_cast<long long>(n) * H + iy) * W + ix) * Cin;
        const float* wp = w + (static_cast<long long>(ky * KW + kx) * Cin) * Cout + co;
        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(__ldg(xp + ci), __ldg(wp + static_cast<long long>(ci) * Cout), acc);
      }
    }
    const bool pos = acc > 0.f;
    if (mask_out) {
      // Cout % 32 == 0 and blockDim % 32 == 0: a warp covers 32 consecutive channels of one pixel
      const uint32_t bits = __ballot_sync(0xffffffffu, pos);
      if ((threadIdx.x & 31) == 0) mask_out[pix * (Cout / 32) + (co >> 5)] = bits;
    }
    if (act == DPIG_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == DPIG_ACT_LRELU) acc = pos ? acc : alpha * acc;
    if (ohi) {
      __nv_bfloat16 h, l;
      split_bf16(acc, h, l);
      ohi[pix * ops + co] = h;
      if (olo) olo[pix * ops + co] = l;
    }
    if (of32) of32[pix * Cout + co] = acc;
  }
}

// warp = one input pixel; lanes stride over co; Cin <= 4 partial sums reduced by shuffles.
__global__ void conv_small_bwd_data_kernel(const float* __restrict__ dy, int N, int OH, int OW, int Cout,
                                           const float* __restrict__ w, int KH, int KW, int stride,
                                           int pt, int pl, int H, int W, int Cin, float* dx) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long total = static_cast<long long>(N) * H * W;
  for (long long pix = warp_id; pix < total; pix += nwarps) {
    const int ix = static_cast<int>(pix % W);
    const int iy = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < KH; ++ky) {
      const int ey = iy + pt - ky;
      if (ey < 0 || ey % stride) continue;
      const int oy = ey / stride;
      if (oy >= OH) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int ex = ix + pl - kx;
        if (ex < 0 || ex % stride) continue;
        const int ox = ex / stride;
        if (ox >= OW) continue;
        const float* dp = dy + ((static_cast<long long>(n) * OH + oy) * OW + ox) * Cout;
        const float* wp = w + static_cast<long long>(ky * KW + kx) * Cin * Cout;
        for (int co = lane; co < Cout; co += 32) {
          const float g = __ldg(dp + co);
          for (int ci = 0; ci < Cin; ++ci) acc[ci] = fmaf(g, __ldg(wp + static_cast<long long>(ci) * Cout + co), acc[ci]);
        }
      }
    }
    for (int ci = 0; ci < Cin; ++ci) {
      float v = acc[ci];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) dx[pix * Cin + ci] = v;
    }
  }
}

// block = chunk of output pixels; thread = co (strided); TAPS*CIN register accumulators.
template <int KK>  // KK = KH*KW*Cin
__global__ void conv_small_bwd_filter_kernel(const float* __restrict__ x, int N, int H, int W, int Cin,
                                             const float* __restrict__ dy, int KH, int KW, int stride,
                                             int pt, int pl, int OH, int OW, int Cout, float* dw,
                                             int chunk) {
  const long long total = static_cast<long long>(N) * OH * OW;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, total);
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k] = 0.f;
    for (long long pix = p0; pix < p1; ++pix) {
      const int ox = static_cast<int>(pix % OW);
      const int oy = static_cast<int>((pix / OW) % OH);
      const int n = static_cast<int>(pix / (static_cast<long long>(OW) * OH));
      const float g = __ldg(dy + pix * Cout + co);
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const int ci = k % Cin;
        const int t = k / Cin;
        const int ky = t / KW, kx = t % KW;
        const int iy = oy * stride + ky - pt, ix = ox * stride + kx - pl;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          acc[k] = fmaf(__ldg(x + ((static_cast<long long>(n) * H + iy) * W + ix) * Cin + ci), g, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < KK; ++k) atomicAdd(dw + static_cast<long long>(k) * Cout + co, acc[k]);
  }
}

}  // namespace dpig
using namespace dpig;

static inline int grid_cap(long long total, int block, int cap) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" int dpig_conv2d_small_fwd(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_,
                                     int32_t cin, const float* w, const float* bias, int32_t kh, int32_t kw,
                                     int32_t stride, int32_t cout, int32_t act, float alpha,
                                     const dpig_tensor* out, float* out_f32, uint32_t* mask_out,
                                     dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !w) return set_error(ctx, DPIG_EINVAL, "conv2d_small_fwd: null argument");
  if (mask_out && cout % 32) return set_error(ctx, DPIG_EINVAL, "conv2d_small_fwd: mask needs cout %% 32 == 0");
  const int OH = same_out(h, stride), OW = same_out(w_, stride);
  const int pt = same_pad_before(h, kh, stride), pl = same_pad_before(w_, kw, stride);
  if (out && (out->n != n || out->h != OH || out->w != OW || out->c < cout))
    return set_error(ctx, DPIG_EINVAL, "conv2d_small_fwd: output shape mismatch");
  const long long total = static_cast<long long>(n) * OH * OW * cout;
  // grid-stride with blockDim 256: total is a multiple of 32 whenever cout is, so ballots are full
  conv_small_fwd_kernel<<<grid_cap(total, 256, 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, h, w_, cin, w, bias, kh, kw, stride, pt, pl, OH, OW, cout, act, alpha,
      out ? static_cast<__nv_bfloat16*>(out->hi) : nullptr, out ? static_cast<__nv_bfloat16*>(out->lo) : nullptr,
      out ? out->pix_stride : 0, out_f32, mask_out);
  ctx->launches++;
  return check_launch(ctx, "conv_small_fwd");
}

extern "C" int dpig_conv2d_small_bwd_data(dpig_ctx* ctx, const float* dy, int32_t n, int32_t oh, int32_t ow,
                                          int32_t cout, const float* w, int32_t kh, int32_t kw, int32_t stride,
                                          int32_t in_h, int32_t in_w, int32_t cin, float* dx, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !w || !dx || cin > 4) return set_error(ctx, DPIG_EINVAL, "conv2d_small_bwd_data: bad argument");
  const int pt = same_pad_before(in_h, kh, stride), pl = same_pad_before(in_w, kw, stride);
  const long long total = static_cast<long long>(n) * in_h * in_w;
  conv_small_bwd_data_kernel<<<grid_cap(total * 32, 256, 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, n, oh, ow, cout, w, kh, kw, stride, pt, pl, in_h, in_w, cin, dx);
  ctx->launches++;
  return check_launch(ctx, "conv_small_bwd_data");
}

extern "C" int dpig_conv2d_small_bwd_filter(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_,
                                            int32_t cin, const float* dy, int32_t kh, int32_t kw, int32_t stride,
                                            int32_t cout, float* dw, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !dy || !dw) return set_error(ctx, DPIG_EINVAL, "conv2d_small_bwd_filter: null argument");
  const int OH = same_out(h, stride), OW = same_out(w_, stride);
  const int pt = same_pad_before(h, kh, stride), pl = same_pad_before(w_, kw, stride);
  const long long total = static_cast<long long>(n) * OH * OW;
  int chunk = static_cast<int>((total + 148 * 8 - 1) / (148 * 8));
  if (chunk < 8) chunk = 8;
  const int blocks = static_cast<int>((total + chunk - 1) / chunk);
  const int threads = cout >= 128 ? 128 : 64;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kk = kh * kw * cin;
  if (kk == 27)
    conv_small_bwd_filter_kernel<27><<<blocks, threads, 0, s>>>(x, n, h, w_, cin, dy, kh, kw, stride, pt, pl, OH, OW, cout, dw, chunk);
  else if (kk == 75)
    conv_small_bwd_filter_kernel<75><<<blocks, threads, 0, s>>>(x, n, h, w_, cin, dy, kh, kw, stride, pt, pl, OH, OW, cout, dw, chunk);
  else
    return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_small_bwd_filter: only 3x3x3 and 5x5x3 filters");
  ctx->launches++;
  return check_launch(ctx, "conv_small_bwd_filter");
}
