// Losses and optimisers of the Stage-I/II trainers, restating the TensorFlow-1.4 formulas the reference
// calls:  L1 reconstruction (trainer.py:606-607, 622), GAN losses incl. the WGAN-GP interpolation and
// penalty (trainer.py:217-252), Adam / RMSProp(+clip) (trainer.py:116-149).  All HBM-bound fp32.
#include <algorithm>
#include "common.cuh"

namespace dpig {

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

__global__ void l1_kernel(const float* g, const float* x, long long count, float weight, float* out, float* dg) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float inv = 1.0f / static_cast<float>(count);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = g[i] - x[i];
    acc += fabsf(d);
    if (dg) dg[i] += weight * inv * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  }
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, s * inv);
}

__device__ __forceinline__ float sce(float z, float l) { return fmaxf(z, 0.f) - z * l + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigm(float z) { return 1.f / (1.f + expf(-z)); }

// single block; count is the number of logits per side
__global__ void gan_loss_kernel(int mode, const float* dr, const float* df, int count, float* out, float* dfg,
                                float* drd, float* dfd) {
  __shared__ float sh[32];
  float a = 0.f, b = 0.f, c = 0.f;  // a: generator term on fake, b: disc term on fake, c: disc term on real
  const float inv = 1.0f / count;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float f = df[i];
    const float r = dr ? dr[i] : 0.f;
    if (mode == DPIG_GAN_DCGAN) {
      a += sce(f, 1.f);
      b += sce(f, 0.f);
      c += sce(r, 1.f);
      if (dfg) dfg[i] = (sigm(f) - 1.f) * inv;
      if (dfd) dfd[i] = sigm(f) * inv * 0.5f;
      if (drd) drd[i] = (sigm(r) - 1.f) * inv * 0.5f;
    } else if (mode == DPIG_GAN_LSGAN) {
      a += (f - 1.f) * (f - 1.f);
      b += f * f;
      c += (r - 1.f) * (r - 1.f);
      if (dfg) dfg[i] = 2.f * (f - 1.f) * inv;
      if (dfd) dfd[i] = f * inv;
      if (drd) drd[i] = (r - 1.f) * inv;
    } else {  // wgan, wgan-gp (the penalty is added by dpig_gp_penalty)
      a += -f;
      b += f;
      c += -r;
      if (dfg) dfg[i] = -inv;
      if (dfd) dfd[i] = inv;
      if (drd) drd[i] = -inv;
    }
  }
  const float sa = block_sum(a, sh);
  const float sb = block_sum(b, sh);
  const float sc = block_sum(c, sh);
  if (threadIdx.x == 0) {
    out[0] = sa * inv;
    const float d = sb * inv + sc * inv;
    out[1] = (mode == DPIG_GAN_DCGAN || mode == DPIG_GAN_LSGAN) ? 0.5f * d : d;
  }
}

// Pose auto-encoder reconstruction loss (trainer.py:638-660, models.py:97-108, 501-515): single block.
//   vis = round(sigmoid(logit)) with the straight-through gradient of binaryRound,
//   G_rcv[b,k,:] = (coord[b,2k], coord[b,2k+1], vis[b,k]),  loss = mean((target - G_rcv)^2) over B*K*3.
__global__ void pose_ae_loss_kernel(const float* target, const float* coord, const float* logit, int B, int K,
                                    float weight, float* out, float* dcoord, float* dlogit, float* g_rcv) {
  __shared__ float sh[32];
  const int total = B * K;
  const float inv = 1.0f / static_cast<float>(total * 3);
  float acc = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const float* t = target + 3 * i;
    const float r = coord[2 * i], c = coord[2 * i + 1];
    const float sg = sigm(logit[i]);
    const float v = rintf(sg);  // tf.round: half to even
    const float dr = r - t[0], dc = c - t[1], dv = v - t[2];
    acc += dr * dr + dc * dc + dv * dv;
    if (dcoord) {
      dcoord[2 * i] = weight * 2.f * dr * inv;
      dcoord[2 * i + 1] = weight * 2.f * dc * inv;
    }
    if (dlogit) dlogit[i] = weight * 2.f * dv * inv * sg * (1.f - sg);
    if (g_rcv) {
      g_rcv[3 * i] = r;
      g_rcv[3 * i + 1] = c;
      g_rcv[3 * i + 2] = v;
    }
  }
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = s * inv;
}

__global__ void gp_interp_kernel(const float* x, const float* g, const float* alpha, long long per, long long total,
                                 float* xhat) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float a = alpha[i / per];
    xhat[i] = x[i] + a * (g[i] - x[i]);
  }
}

// block = one sample
__global__ void gp_slopes_kernel(const float* grad, long long per, float* slopes) {
  __shared__ float sh[32];
  const float* gptr = grad + blockIdx.x * per;
  float acc = 0.f;
  for (long long j = threadIdx.x; j < per; j += blockDim.x) acc = fmaf(gptr[j], gptr[j], acc);
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) slopes[blockIdx.x] = sqrtf(s);
}
__global__ void gp_finish_kernel(const float* grad, const float* slopes, int n, long long per, float lambda,
                                 float* out, float* dgrad) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += (slopes[i] - 1.f) * (slopes[i] - 1.f);
    out[0] = acc / n;
  }
  if (!dgrad) return;
  const long long total = per * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = slopes[i / per];
    dgrad[i] = lambda * 2.f * (s - 1.f) / (static_cast<float>(n) * s) * grad[i];
  }
}

// lr_dev (optional): the step size is read from device memory instead of the launch argument, so that a captured
// CUDA graph of the whole optimiser step can be replayed while Adam's bias-corrected lr_t changes every step.
// 16-byte accesses on the 4-aligned body (28 B/param of HBM traffic is the whole cost of this kernel), scalar tail.
__global__ void adam_kernel(float* p, const float* g, float* m, float* v, long long count, float lr_arg,
                            const float* lr_dev, float b1, float b2, float eps, float gs, int vec) {
  const float lr_t = lr_dev ? __ldg(lr_dev) : lr_arg;
  const long long n4 = vec ? count >> 2 : 0;
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = tid; i < n4; i += nth) {
    const float4 gv = g4[i];
    float4 mv = m4[i], vv = v4[i], pv = p4[i];
    const float ga[4] = {gv.x * gs, gv.y * gs, gv.z * gs, gv.w * gs};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ma[j] = b1 * ma[j] + (1.f - b1) * ga[j];
      va[j] = b2 * va[j] + (1.f - b2) * ga[j] * ga[j];
      pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + eps);
    }
    m4[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v4[i] = make_float4(va[0], va[1], va[2], va[3]);
    p4[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
  }
  for (long long i = (n4 << 2) + tid; i < count; i += nth) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void rmsprop_kernel(float* p, const float* g, float* ms, long long count, float lr_arg, const float* lr_dev,
                               float decay, float eps, float gs, float clip) {
  const float lr = lr_dev ? __ldg(lr_dev) : lr_arg;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gs;
    const float msi = decay * ms[i] + (1.f - decay) * gi * gi;
    ms[i] = msi;
    float pi = p[i] - lr * gi / sqrtf(msi + eps);
    if (clip > 0.f) pi = fminf(fmaxf(pi, -clip), clip);
    p[i] = pi;
  }
}

__global__ void clip_kernel(float* p, long long count, float lo, float hi) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] = fminf(fmaxf(p[i], lo), hi);
}

// tf coord2channel_simple_rcv + tf_poseInflate (utils.py:259-318): +1 inside the radius-4 disc around
// each visible keypoint (int-truncated row/col), -1 elsewhere.
__global__ void pose_raster_kernel(const float* rcv, int N, int K, int H, int W, int radius,
                                   __nv_bfloat16* ohi, __nv_bfloat16* olo, long long ops, float* of32) {
  const long long total = static_cast<long long>(N) * H * W * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const long long pix = i / K;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    const float* q = rcv + (static_cast<long long>(n) * K + k) * 3;
    const int r0 = static_cast<int>(q[0]), c0 = static_cast<int>(q[1]);  // tf.to_int32 truncates
    const float vis = q[2];
    const int dr = y - r0, dc = x - c0;
    const bool inside = (dr * dr + dc * dc <= radius * radius);
    float v = (inside ? fminf(vis, 1.f) : 0.f) * 2.f - 1.f;
    if (ohi) {
      ohi[pix * ops + k] = __float2bfloat16_rn(v);
      if (olo) olo[pix * ops + k] = __float2bfloat16_rn(v - __bfloat162float(__float2bfloat16_rn(v)));
    }
    if (of32) of32[pix * K + k] = v;
  }
}

// The 3x3 (kh x kw) SAME-padded patches of the inflated pose maps straight from the keypoints: out[n,y,x,(i*kw+j)*K + k] =
// map_k(y + i - pt, x + j - pl), 0 outside the image (the conv's zero padding) and in the pad channels.  The U-Net stem
// reads the pose channels in this patch form (one 1x1 contraction over K = 9*18 instead of nine 64-deep K chunks of
// 18 channels); building the patches from the 18 keypoints costs no read of the maps at all.
// thread = (pixel, 8 patch channels): one 16-byte store per plane, consecutive lanes write consecutive 16-byte pieces
// (a pixel's patch row is 24 pieces).  K and the tap grid are compile-time constants so that c / K, c % K, tap / KW cost a
// multiply-shift: with run-time divisors this kernel was instruction-bound (0.5 ms per call), and a (pixel, tap) thread
// layout with 4-byte stores at a 36-byte stride was bound by its 8x store-sector amplification (0.57 ms).
// The (row, col, visible) triples of the image's keypoints are decoded once per block into shared memory.
template <int K, int KH, int KW>
__global__ void __launch_bounds__(256)
pose_patch_kernel(const float* rcv, int N, int H, int W, int radius, int pt, int pl, __nv_bfloat16* ohi,
                  __nv_bfloat16* olo, long long ops, int Cp, int blocks_per_image) {
  __shared__ int s_r[K], s_c[K];
  __shared__ float s_v[K];
  const int n = blockIdx.x / blocks_per_image;
  if (threadIdx.x < K) {
    const float* q = rcv + (static_cast<long long>(n) * K + threadIdx.x) * 3;
    s_r[threadIdx.x] = static_cast<int>(q[0]);     // tf.to_int32 truncates
    s_c[threadIdx.x] = static_cast<int>(q[1]);
    s_v[threadIdx.x] = fminf(q[2], 1.f) * 2.f - 1.f;
  }
  __syncthreads();
  const int G8 = Cp / 8;
  const int r2 = radius * radius;
  const int items = H * W * G8;
  for (int it = (blockIdx.x % blocks_per_image) * blockDim.x + threadIdx.x; it < items;
       it += blocks_per_image * blockDim.x) {
    const int g = it % G8;
    const int pin = it / G8;
    const int x = pin % W, y = pin / W;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      uint16_t hv = 0, lv = 0;
      if (c < KH * KW * K) {
        const int tap = c / K, k = c % K;
        const int yy = y + tap / KW - pt, xx = x + tap % KW - pl;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const int dr = yy - s_r[k], dc = xx - s_c[k];
          const float v = (dr * dr + dc * dc <= r2) ? s_v[k] : -1.f;           // as pose_raster_kernel
          const __nv_bfloat16 hb = __float2bfloat16_rn(v);
          hv = __bfloat16_as_ushort(hb);
          lv = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hb)));
        }
      }
      h[e >> 1] |= static_cast<uint32_t>(hv) << ((e & 1) * 16);
      l[e >> 1] |= static_cast<uint32_t>(lv) << ((e & 1) * 16);
    }
    const long long o = (static_cast<long long>(n) * H * W + pin) * ops + g * 8;
    *reinterpret_cast<uint4*>(ohi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (olo) *reinterpret_cast<uint4*>(olo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

static inline int gridn(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace dpig
using namespace dpig;

extern "C" int dpig_loss_l1(dpig_ctx* ctx, const float* g, const float* x, int64_t count, float weight,
                            float* out, float* dg, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!g || !x || !out) return set_error(ctx, DPIG_EINVAL, "loss_l1: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(out, 0, sizeof(float), s);
  l1_kernel<<<gridn(count, 256, 148 * 4), 256, 0, s>>>(g, x, count, weight, out, dg);
  ctx->launches++;
  return check_launch(ctx, "loss_l1");
}

extern "C" int dpig_loss_gan(dpig_ctx* ctx, int32_t mode, const float* d_real, const float* d_fake, int32_t count,
                             float* out, float* d_fake_g, float* d_real_d, float* d_fake_d, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!d_fake || !out) return set_error(ctx, DPIG_EINVAL, "loss_gan: null argument");
  gan_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(mode, d_real, d_fake, count, out, d_fake_g,
                                                                     d_real_d, d_fake_d);
  ctx->launches++;
  return check_launch(ctx, "loss_gan");
}

extern "C" int dpig_pose_ae_loss(dpig_ctx* ctx, const float* target, const float* coord, const float* vis_logit,
                                 int32_t batch, int32_t keypoints, float weight, float* out, float* d_coord,
                                 float* d_vis_logit, float* g_rcv, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!target || !coord || !vis_logit || !out) return set_error(ctx, DPIG_EINVAL, "pose_ae_loss: null argument");
  pose_ae_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(target, coord, vis_logit, batch, keypoints,
                                                                     weight, out, d_coord, d_vis_logit, g_rcv);
  ctx->launches++;
  return check_launch(ctx, "pose_ae_loss");
}

extern "C" int dpig_gp_interpolate(dpig_ctx* ctx, const float* x, const float* g, const float* alpha, int32_t n,
                                   int64_t per_sample, float* xhat, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !g || !alpha || !xhat) return set_error(ctx, DPIG_EINVAL, "gp_interpolate: null argument");
  const long long total = static_cast<long long>(n) * per_sample;
  gp_interp_kernel<<<gridn(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, g, alpha, per_sample, total, xhat);
  ctx->launches++;
  return check_launch(ctx, "gp_interpolate");
}

extern "C" int dpig_gp_penalty(dpig_ctx* ctx, const float* grad, int32_t n, int64_t per_sample, float lambda,
                               float* slopes, float* out, float* dgrad, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!grad || !slopes || !out) return set_error(ctx, DPIG_EINVAL, "gp_penalty: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gp_slopes_kernel<<<n, 256, 0, s>>>(grad, per_sample, slopes);
  gp_finish_kernel<<<gridn(static_cast<long long>(n) * per_sample), 256, 0, s>>>(grad, slopes, n, per_sample, lambda,
                                                                                 out, dgrad);
  ctx->launches += 2;
  return check_launch(ctx, "gp_penalty");
}

static bool aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

extern "C" int dpig_adam_step(dpig_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count, float lr,
                              float beta1, float beta2, float eps, int32_t t, float grad_scale, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!p || !g || !m || !v || t < 1) return set_error(ctx, DPIG_EINVAL, "adam_step: bad argument");
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(beta2), t)) /
                      (1.0 - pow(static_cast<double>(beta1), t));
  // the vector body needs 16-byte aligned arrays; otherwise everything goes through the scalar loop
  adam_kernel<<<gridn(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, count, static_cast<float>(lr_t),
                                                                           nullptr, beta1, beta2, eps, grad_scale,
                                                                           aligned16(p, g, m, v) ? 1 : 0);
  ctx->launches++;
  return check_launch(ctx, "adam_step");
}

extern "C" int dpig_adam_step_dev(dpig_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count,
                                  const float* lr_t_dev, float beta1, float beta2, float eps, float grad_scale,
                                  dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!p || !g || !m || !v || !lr_t_dev) return set_error(ctx, DPIG_EINVAL, "adam_step_dev: null argument");
  adam_kernel<<<gridn(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, count, 0.f, lr_t_dev, beta1,
                                                                           beta2, eps, grad_scale,
                                                                           aligned16(p, g, m, v) ? 1 : 0);
  ctx->launches++;
  return check_launch(ctx, "adam_step_dev");
}

extern "C" int dpig_rmsprop_step(dpig_ctx* ctx, float* p, const float* g, float* ms, int64_t count, float lr,
                                 float decay, float eps, float grad_scale, float clip, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!p || !g || !ms) return set_error(ctx, DPIG_EINVAL, "rmsprop_step: null argument");
  rmsprop_kernel<<<gridn(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, ms, count, lr, nullptr, decay, eps,
                                                                              grad_scale, clip);
  ctx->launches++;
  return check_launch(ctx, "rmsprop_step");
}

extern "C" int dpig_rmsprop_step_dev(dpig_ctx* ctx, float* p, const float* g, float* ms, int64_t count,
                                     const float* lr_dev, float decay, float eps, float grad_scale, float clip,
                                     dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!p || !g || !ms || !lr_dev) return set_error(ctx, DPIG_EINVAL, "rmsprop_step_dev: null argument");
  rmsprop_kernel<<<gridn(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, ms, count, 0.f, lr_dev, decay, eps,
                                                                              grad_scale, clip);
  ctx->launches++;
  return check_launch(ctx, "rmsprop_step_dev");
}

extern "C" int dpig_clip(dpig_ctx* ctx, float* p, int64_t count, float lo, float hi, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  clip_kernel<<<gridn(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, count, lo, hi);
  ctx->launches++;
  return check_launch(ctx, "clip");
}

extern "C" int dpig_pose_rasterize(dpig_ctx* ctx, const float* rcv, int32_t n, int32_t k, int32_t h, int32_t w_,
                                   int32_t radius, const dpig_tensor* out, float* out_f32, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!rcv || (!out && !out_f32)) return set_error(ctx, DPIG_EINVAL, "pose_rasterize: null argument");
  if (out && (out->n != n || out->h != h || out->w != w_ || out->c < k))
    return set_error(ctx, DPIG_EINVAL, "pose_rasterize: output shape mismatch");
  const long long total = static_cast<long long>(n) * h * w_ * k;
  pose_raster_kernel<<<gridn(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rcv, n, k, h, w_, radius, out ? static_cast<__nv_bfloat16*>(out->hi) : nullptr,
      out ? static_cast<__nv_bfloat16*>(out->lo) : nullptr, out ? out->pix_stride : 0, out_f32);
  ctx->launches++;
  return check_launch(ctx, "pose_rasterize");
}

extern "C" int dpig_pose_patch(dpig_ctx* ctx, const float* rcv, int32_t n, int32_t k, int32_t h, int32_t w_,
                               int32_t radius, int32_t kh, int32_t kw, const dpig_tensor* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!rcv || !out || !out->hi) return set_error(ctx, DPIG_EINVAL, "pose_patch: null argument");
  if (k != 18 || kh != 3 || kw != 3)
    return set_error(ctx, DPIG_EUNSUPPORTED, "pose_patch: built for the reference's 18 keypoints under a 3x3 stem (got %d, %dx%d)",
                     k, kh, kw);
  if (out->n != n || out->h != h || out->w != w_ || out->c < kh * kw * k || out->c % 8 || out->pix_stride % 8 ||
      reinterpret_cast<uintptr_t>(out->hi) % 16 || (out->lo && reinterpret_cast<uintptr_t>(out->lo) % 16))
    return set_error(ctx, DPIG_EINVAL, "pose_patch: output must be [%d,%d,%d,>=%d] with channels / stride multiples of 8",
                     n, h, w_, kh * kw * k);
  const long long items = static_cast<long long>(h) * w_ * (out->c / 8);
  int bpi = static_cast<int>((items + 255) / 256);
  const int cap = std::max(1, 148 * 16 / std::max(1, n));    // grid ~ a few waves of the 148 SMs
  if (bpi > cap) bpi = cap;
  pose_patch_kernel<18, 3, 3><<<n * bpi, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rcv, n, h, w_, radius, same_pad_before(h, kh, 1), same_pad_before(w_, kw, 1),
      static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->pix_stride, out->c, bpi);
  ctx->launches++;
  return check_launch(ctx, "pose_patch");
}
