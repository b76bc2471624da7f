// Normalisation + LeakyReLU blocks of the discriminator (reference wgan_gp.py:34-40, 418-430):
//   MODE 'wgan-gp' -> per-sample LayerNorm over (C,H,W)      tflib/ops/layernorm.py:6-20
//   otherwise      -> training-mode BatchNorm over (N,H,W)   tflib/ops/batchnorm.py:29-30
//   (+ InstanceNorm, models.py:154-166, defined in the reference but never called).
// All use the biased variance and rsqrt(var + eps).  x is the fp32 NHWC conv output (bias added).
// Raw sums are accumulated in fp64 so var = E[x^2] - mean^2 is safe; the raw-sum interface is also the
// sync-BN hook: data-parallel ranks all-reduce `sums` (forward) and `red` (backward) between phases.
#include "common.cuh"

namespace dpig {

__device__ __forceinline__ int group_of(int mode, int n, int c, int C) {
  return mode == DPIG_NORM_BATCH ? c : (mode == DPIG_NORM_LAYER ? n : n * C + c);
}

// grid = (pixel chunks, N); threads stride channels.
__global__ void norm_stats_kernel(const float* x, int C, long long ppi, int chunk, int mode, int groups,
                                  double* sums) {
  __shared__ double red[2][32];
  const int n = blockIdx.y;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, ppi);
  double ls = 0.0, ls2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, s2 = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const float v = x[(n * ppi + p) * C + c];
      s += v;
      s2 = fmaf(v, v, s2);
    }
    if (mode == DPIG_NORM_LAYER) {
      ls += s;
      ls2 += s2;
    } else {
      const int g = group_of(mode, n, c, C);
      atomicAdd(sums + g, static_cast<double>(s));
      atomicAdd(sums + groups + g, static_cast<double>(s2));
    }
  }
  if (mode == DPIG_NORM_LAYER) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ls += __shfl_xor_sync(0xffffffffu, ls, o);
      ls2 += __shfl_xor_sync(0xffffffffu, ls2, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
      red[0][w] = ls;
      red[1][w] = ls2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) {
        a += red[0][i];
        b += red[1][i];
      }
      atomicAdd(sums + n, a);
      atomicAdd(sums + groups + n, b);
    }
  }
}

__global__ void norm_finalize_kernel(const double* sums, int groups, double count, float eps, float* stats) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const double mean = sums[g] / count;
  double var = sums[groups + g] / count - mean * mean;
  if (var < 0) var = 0;
  stats[g] = static_cast<float>(mean);
  stats[groups + g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// thread = (pixel, 32-channel word).  kFused: (mean, rstd) of every group are derived from the raw fp64 sums by each
// block into shared memory (<= kMaxFusedGroups groups: 2 fp64 divisions + a square root per group and block), block 0
// also writes them to `stats` for the backward pass -- one launch instead of finalize + apply.
constexpr int kMaxFusedGroups = 2048;
template <bool kFused>
__global__ void norm_act_fwd_kernel(const float* x, int N, long long ppi, int C, int mode, int groups,
                                    const double* sums, double count, float eps, float* stats, const float* scale,
                                    const float* offset, int act, float alpha, __nv_bfloat16* ohi,
                                    __nv_bfloat16* olo, long long ops, uint32_t* mask_out) {
  extern __shared__ float s_stats[];   // kFused: [2][groups]
  const float* st = stats;
  if (kFused) {
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      const double mean = sums[g] / count;
      double var = sums[groups + g] / count - mean * mean;
      if (var < 0) var = 0;
      const float m = static_cast<float>(mean), r = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      s_stats[g] = m;
      s_stats[groups + g] = r;
      if (blockIdx.x == 0) {
        stats[g] = m;
        stats[groups + g] = r;
      }
    }
    __syncthreads();
    st = s_stats;
  }
  const int words = C / 32;
  const long long total = static_cast<long long>(N) * ppi * words;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wd = static_cast<int>(i % words);
    const long long pix = i / words;
    const int n = static_cast<int>(pix / ppi);
    uint32_t bits = 0;
    const float4* x4 = reinterpret_cast<const float4*>(x + pix * C + wd * 32);
    uint4* oh4 = reinterpret_cast<uint4*>(ohi + pix * ops + wd * 32);
    uint4* ol4 = olo ? reinterpret_cast<uint4*>(olo + pix * ops + wd * 32) : nullptr;
    const bool vec = (ops % 8 == 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {     // 8 channels per round: two 16-byte loads, one 16-byte store per plane
      const float4 a = __ldg(x4 + 2 * q), b = __ldg(x4 + 2 * q + 1);
      const float in[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = q * 8 + e;
        const int c = wd * 32 + j;
        const int g = group_of(mode, n, c, C);
        float v = (in[e] - st[g]) * st[groups + g] * __ldg(scale + c) + __ldg(offset + c);
        if (v > 0.f) bits |= 1u << j;
        if (act == DPIG_ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == DPIG_ACT_LRELU) v = v > 0.f ? v : alpha * v;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        if (vec) {
          if (e & 1) {
            hw[e >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(h)) << 16;
            lw[e >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(l)) << 16;
          } else {
            hw[e >> 1] = __bfloat16_as_ushort(h);
            lw[e >> 1] = __bfloat16_as_ushort(l);
          }
        } else {
          ohi[pix * ops + c] = h;
          if (olo) olo[pix * ops + c] = l;
        }
      }
      if (vec) {
        oh4[q] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (ol4) ol4[q] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
    if (mask_out) mask_out[pix * words + wd] = bits;
  }
}

__device__ __forceinline__ float ld_dz(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long off,
                                       const uint32_t* mask, long long pix, int words, int c, float alpha) {
  float v = __bfloat162float(hi[off]);
  if (lo) v += __bfloat162float(lo[off]);
  if (mask) {
    const uint32_t m = mask[pix * words + (c >> 5)];
    if (!((m >> (c & 31)) & 1u)) v *= alpha;
  }
  return v;
}

// grid = (pixel chunks, N); threads stride channels.
__global__ void norm_bwd_reduce_kernel(const __nv_bfloat16* ghi, const __nv_bfloat16* glo, long long gps,
                                       const float* x, int C, long long ppi, int chunk, int mode, int groups,
                                       const float* stats, const uint32_t* mask, float alpha,
                                       const float* scale, double* red, float* dscale, float* doffset) {
  __shared__ double sred[2][32];
  const int n = blockIdx.y;
  const int words = C / 32;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, ppi);
  double l1 = 0.0, l2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = group_of(mode, n, c, C);
    const float mean = stats[g], rstd = stats[groups + g];
    float db = 0.f, dg = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const long long pix = n * ppi + p;
      const float dz = ld_dz(ghi, glo, pix * gps + c, mask, pix, words, c, alpha);
      const float xh = (x[pix * C + c] - mean) * rstd;
      db += dz;
      dg = fmaf(dz, xh, dg);
    }
    if (doffset) atomicAdd(doffset + c, db);
    if (dscale) atomicAdd(dscale + c, dg);
    const float gam = scale[c];
    if (mode == DPIG_NORM_LAYER) {
      l1 += static_cast<double>(gam) * db;
      l2 += static_cast<double>(gam) * dg;
    } else {
      atomicAdd(red + g, static_cast<double>(gam) * db);
      atomicAdd(red + groups + g, static_cast<double>(gam) * dg);
    }
  }
  if (mode == DPIG_NORM_LAYER) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l1 += __shfl_xor_sync(0xffffffffu, l1, o);
      l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
      sred[0][w] = l1;
      sred[1][w] = l2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) {
        a += sred[0][i];
        b += sred[1][i];
      }
      atomicAdd(red + n, a);
      atomicAdd(red + groups + n, b);
    }
  }
}

__global__ void norm_bwd_apply_kernel(const __nv_bfloat16* ghi, const __nv_bfloat16* glo, long long gps,
                                      const float* x, int N, long long ppi, int C, int mode, int groups,
                                      const float* stats, const uint32_t* mask, float alpha, const float* scale,
                                      const double* red, double count, __nv_bfloat16* ohi, __nv_bfloat16* olo,
                                      long long ops) {
  const int words = C / 32;
  const long long total = static_cast<long long>(N) * ppi * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int n = static_cast<int>(pix / ppi);
    const int g = group_of(mode, n, c, C);
    const float mean = stats[g], rstd = stats[groups + g];
    const float dz = ld_dz(ghi, glo, pix * gps + c, mask, pix, words, c, alpha);
    const float xh = (x[pix * C + c] - mean) * rstd;
    const float m1 = static_cast<float>(red[g] / count), m2 = static_cast<float>(red[groups + g] / count);
    const float dx = rstd * (dz * scale[c] - m1 - xh * m2);
    __nv_bfloat16 h, l;
    split_bf16(dx, h, l);
    ohi[pix * ops + c] = h;
    if (olo) olo[pix * ops + c] = l;
  }
}

static inline int groups_of(int mode, int n, int c) {
  return mode == DPIG_NORM_BATCH ? c : (mode == DPIG_NORM_LAYER ? n : n * c);
}

}  // namespace dpig
using namespace dpig;

extern "C" int dpig_norm_stats(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_, int32_t c,
                               int32_t mode, double* sums, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !sums) return set_error(ctx, DPIG_EINVAL, "norm_stats: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = groups_of(mode, n, c);
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups, s);
  const long long ppi = static_cast<long long>(h) * w_;
  const int chunk = 32;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), n);
  norm_stats_kernel<<<grid, 256, 0, s>>>(x, c, ppi, chunk, mode, groups, sums);
  ctx->launches++;
  return check_launch(ctx, "norm_stats");
}

extern "C" int dpig_norm_act_fwd(dpig_ctx* ctx, const float* x, int32_t n, int32_t h, int32_t w_, int32_t c,
                                 int32_t mode, float eps, const double* sums, double count, const float* scale,
                                 const float* offset, int32_t act, float alpha, float* stats,
                                 const dpig_tensor* out, uint32_t* mask_out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !sums || !scale || !offset || !stats || !out || c % 32)
    return set_error(ctx, DPIG_EINVAL, "norm_act_fwd: bad argument (c must be a multiple of 32)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = groups_of(mode, n, c);
  const long long ppi = static_cast<long long>(h) * w_;
  const long long total = static_cast<long long>(n) * ppi * (c / 32);
  long long grid = (total + 127) / 128;
  if (grid > 148 * 16) grid = 148 * 16;
  if (reinterpret_cast<uintptr_t>(out->hi) % 16 || (out->lo && reinterpret_cast<uintptr_t>(out->lo) % 16))
    return set_error(ctx, DPIG_EINVAL, "norm_act_fwd: output planes must be 16-byte aligned");
  if (groups <= kMaxFusedGroups) {
    norm_act_fwd_kernel<true><<<static_cast<int>(grid), 128, sizeof(float) * 2 * groups, s>>>(
        x, n, ppi, c, mode, groups, sums, count, eps, stats, scale, offset, act, alpha,
        static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->pix_stride, mask_out);
    ctx->launches += 1;
    return check_launch(ctx, "norm_act_fwd");
  }
  norm_finalize_kernel<<<(groups + 127) / 128, 128, 0, s>>>(sums, groups, count, eps, stats);
  norm_act_fwd_kernel<false><<<static_cast<int>(grid), 128, 0, s>>>(
      x, n, ppi, c, mode, groups, sums, count, eps, stats, scale, offset, act, alpha,
      static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->pix_stride, mask_out);
  ctx->launches += 2;
  return check_launch(ctx, "norm_act_fwd");
}

extern "C" int dpig_norm_act_bwd_reduce(dpig_ctx* ctx, const dpig_tensor* dy, const float* x, const float* stats,
                                        const uint32_t* mask, float alpha, int32_t mode, const float* scale,
                                        double* red, float* dscale, float* doffset, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !x || !stats || !scale || !red) return set_error(ctx, DPIG_EINVAL, "norm_act_bwd_reduce: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = groups_of(mode, dy->n, dy->c);
  cudaMemsetAsync(red, 0, sizeof(double) * 2 * groups, s);
  const long long ppi = static_cast<long long>(dy->h) * dy->w;
  const int chunk = 32;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), dy->n);
  norm_bwd_reduce_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dy->hi),
                                              static_cast<const __nv_bfloat16*>(dy->lo), dy->pix_stride, x, dy->c,
                                              ppi, chunk, mode, groups, stats, mask, alpha, scale, red, dscale,
                                              doffset);
  ctx->launches++;
  return check_launch(ctx, "norm_act_bwd_reduce");
}

extern "C" int dpig_norm_act_bwd_apply(dpig_ctx* ctx, const dpig_tensor* dy, const float* x, const float* stats,
                                       const uint32_t* mask, float alpha, int32_t mode, const float* scale,
                                       const double* red, double count, const dpig_tensor* dx, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !x || !stats || !scale || !red || !dx) return set_error(ctx, DPIG_EINVAL, "norm_act_bwd_apply: null argument");
  const int groups = groups_of(mode, dy->n, dy->c);
  const long long ppi = static_cast<long long>(dy->h) * dy->w;
  const long long total = static_cast<long long>(dy->n) * ppi * dy->c;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  norm_bwd_apply_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy->hi), static_cast<const __nv_bfloat16*>(dy->lo), dy->pix_stride, x,
      dy->n, ppi, dy->c, mode, groups, stats, mask, alpha, scale, red, count,
      static_cast<__nv_bfloat16*>(dx->hi), static_cast<__nv_bfloat16*>(dx->lo), dx->pix_stride);
  ctx->launches++;
  return check_launch(ctx, "norm_act_bwd_apply");
}

// ------------------------------------------------------------------------------------------------
// Second-order pieces for the WGAN-GP gradient penalty (reference trainer.py:226-236, wgan_gp.py:605-619).
// d(lambda*gp)/d(theta) = d/d(theta) sum_n <v_n, dD(xhat_n)/dxhat_n> with v = d(lambda*gp)/d(grad) held
// constant, i.e. the parameter gradient of the directional derivative (JVP) of D along v.  Convolutions,
// LeakyReLU and the linear layer are (piecewise) linear, so their JVP / adjoint reuse the conv kernels; the
// per-sample LayerNorm (tflib/ops/layernorm.py:6-20) is the only layer with a genuinely second-order term:
//   forward tangent:  zdot = gamma * u,  u = rstd*(pdot - m1 - xhat*m2),  m1 = mean(pdot), m2 = mean(xhat*pdot)
//   adjoint (given zbar = dS/dzdot):  ubar = gamma*zbar, wbar = rstd*ubar,
//        A = mean(wbar), Bq = mean(xhat*wbar), Cq = mean(ubar*u)
//        pdot_bar = wbar - A - xhat*Bq
//        p_bar    = rstd*( -m2*(wbar - A) - Bq*(pdot - m1) + xhat*(2*m2*Bq - Cq) )     (via xhat and sigma)
//        dgamma  += sum zbar*u
// Means are over the sample's C*H*W elements.  Reductions run in fp64.
namespace dpig {

// grid = (pixel chunks, N): per-sample sums of (a, xhat*a) where a = pdot (mode 0) or
// (wbar, xhat*wbar, ubar*u) (mode 1, three sums).
__global__ void ln_jvp_reduce_kernel(const float* p, const float* pdot, const float* stats, int N, int C,
                                     long long ppi, int chunk, const float* scale, const __nv_bfloat16* zhi,
                                     const __nv_bfloat16* zlo, long long zps, const uint32_t* mask, float alpha,
                                     const double* tsums, double* out, float* dscale, int mode) {
  __shared__ double red[3][32];
  const int n = blockIdx.y;
  const float mean = stats[n], rstd = stats[N + n];
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, ppi);
  const double cnt = static_cast<double>(ppi) * C;
  float m1 = 0.f, m2 = 0.f;
  if (mode == 1) {
    m1 = static_cast<float>(tsums[n] / cnt);
    m2 = static_cast<float>(tsums[N + n] / cnt);
  }
  const int words = C / 32;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, dg = 0.f;
    const float gam = mode == 1 ? scale[c] : 1.f;
    for (long long q = p0; q < p1; ++q) {
      const long long pix = n * ppi + q;
      const float xh = (p[pix * C + c] - mean) * rstd;
      const float pd = pdot[pix * C + c];
      if (mode == 0) {
        a0 += pd;
        a1 = fmaf(xh, pd, a1);
      } else {
        const float zb = ld_dz(zhi, zlo, pix * zps + c, mask, pix, words, c, alpha);
        const float u = rstd * (pd - m1 - xh * m2);
        const float ub = gam * zb;
        const float wb = rstd * ub;
        a0 += wb;
        a1 = fmaf(xh, wb, a1);
        a2 = fmaf(ub, u, a2);
        dg = fmaf(zb, u, dg);
      }
    }
    s0 += a0;
    s1 += a1;
    s2 += a2;
    if (mode == 1 && dscale) atomicAdd(dscale + c, dg);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    red[0][w] = s0;
    red[1][w] = s1;
    red[2][w] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c2 = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) {
      a += red[0][i];
      b += red[1][i];
      c2 += red[2][i];
    }
    atomicAdd(out + n, a);
    atomicAdd(out + N + n, b);
    if (mode == 1) atomicAdd(out + 2 * N + n, c2);
  }
}

// hdot = lrelu'(z) * gamma * rstd * (pdot - m1 - xhat*m2)   (split output)
__global__ void ln_jvp_fwd_apply_kernel(const float* p, const float* pdot, const float* stats, int N, int C,
                                        long long ppi, const float* scale, const uint32_t* mask, float alpha,
                                        const double* tsums, __nv_bfloat16* ohi, __nv_bfloat16* olo, long long ops) {
  const long long total = static_cast<long long>(N) * ppi * C;
  const int words = C / 32;
  const double cnt = static_cast<double>(ppi) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int n = static_cast<int>(pix / ppi);
    const float mean = stats[n], rstd = stats[N + n];
    const float m1 = static_cast<float>(tsums[n] / cnt), m2 = static_cast<float>(tsums[N + n] / cnt);
    const float xh = (p[i] - mean) * rstd;
    float v = scale[c] * rstd * (pdot[i] - m1 - xh * m2);
    if (mask) {
      const uint32_t m = mask[pix * words + (c >> 5)];
      if (!((m >> (c & 31)) & 1u)) v *= alpha;
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    ohi[pix * ops + c] = h;
    if (olo) olo[pix * ops + c] = l;
  }
}

__global__ void ln_jvp_bwd_apply_kernel(const float* p, const float* pdot, const float* stats, int N, int C,
                                        long long ppi, const float* scale, const __nv_bfloat16* zhi,
                                        const __nv_bfloat16* zlo, long long zps, const uint32_t* mask, float alpha,
                                        const double* tsums, const double* asums, __nv_bfloat16* ohi,
                                        __nv_bfloat16* olo, long long ops, float* pbar) {
  const long long total = static_cast<long long>(N) * ppi * C;
  const int words = C / 32;
  const double cnt = static_cast<double>(ppi) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int n = static_cast<int>(pix / ppi);
    const float mean = stats[n], rstd = stats[N + n];
    const float m1 = static_cast<float>(tsums[n] / cnt), m2 = static_cast<float>(tsums[N + n] / cnt);
    const float A = static_cast<float>(asums[n] / cnt), Bq = static_cast<float>(asums[N + n] / cnt),
                Cq = static_cast<float>(asums[2 * N + n] / cnt);
    const float xh = (p[i] - mean) * rstd;
    const float zb = ld_dz(zhi, zlo, pix * zps + c, mask, pix, words, c, alpha);
    const float wb = rstd * scale[c] * zb;
    const float pdb = wb - A - xh * Bq;
    const float pb = rstd * (-m2 * (wb - A) - Bq * (pdot[i] - m1) + xh * (2.f * m2 * Bq - Cq));
    __nv_bfloat16 h, l;
    split_bf16(pdb, h, l);
    ohi[pix * ops + c] = h;
    if (olo) olo[pix * ops + c] = l;
    pbar[i] = pb;
  }
}

}  // namespace dpig

extern "C" int dpig_layernorm_jvp_fwd(dpig_ctx* ctx, const float* p, const float* pdot, int32_t n, int32_t h,
                                      int32_t w_, int32_t c, const float* stats, const float* scale,
                                      const uint32_t* mask, float alpha, double* tsums, const dpig_tensor* hdot,
                                      dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!p || !pdot || !stats || !scale || !tsums || !hdot || c % 32)
    return set_error(ctx, DPIG_EINVAL, "layernorm_jvp_fwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(tsums, 0, sizeof(double) * 2 * n, s);
  const long long ppi = static_cast<long long>(h) * w_;
  const int chunk = 32;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), n);
  ln_jvp_reduce_kernel<<<grid, 256, 0, s>>>(p, pdot, stats, n, c, ppi, chunk, scale, nullptr, nullptr, 0, nullptr, alpha,
                                            nullptr, tsums, nullptr, 0);
  const long long total = static_cast<long long>(n) * ppi * c;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  ln_jvp_fwd_apply_kernel<<<static_cast<int>(g), 256, 0, s>>>(p, pdot, stats, n, c, ppi, scale, mask, alpha, tsums,
                                                               static_cast<__nv_bfloat16*>(hdot->hi),
                                                               static_cast<__nv_bfloat16*>(hdot->lo), hdot->pix_stride);
  ctx->launches += 2;
  return check_launch(ctx, "layernorm_jvp_fwd");
}

extern "C" int dpig_layernorm_jvp_bwd(dpig_ctx* ctx, const dpig_tensor* hdot_bar, const uint32_t* mask, float alpha,
                                      const float* p, const float* pdot, const float* stats, const float* scale,
                                      const double* tsums, double* asums, float* dscale, const dpig_tensor* pdot_bar,
                                      float* p_bar, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!hdot_bar || !p || !pdot || !stats || !scale || !tsums || !asums || !pdot_bar || !p_bar || hdot_bar->c % 32)
    return set_error(ctx, DPIG_EINVAL, "layernorm_jvp_bwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = hdot_bar->n, c = hdot_bar->c;
  cudaMemsetAsync(asums, 0, sizeof(double) * 3 * n, s);
  const long long ppi = static_cast<long long>(hdot_bar->h) * hdot_bar->w;
  const int chunk = 32;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), n);
  const __nv_bfloat16* zhi = static_cast<const __nv_bfloat16*>(hdot_bar->hi);
  const __nv_bfloat16* zlo = static_cast<const __nv_bfloat16*>(hdot_bar->lo);
  ln_jvp_reduce_kernel<<<grid, 256, 0, s>>>(p, pdot, stats, n, c, ppi, chunk, scale, zhi, zlo, hdot_bar->pix_stride, mask,
                                            alpha, tsums, asums, dscale, 1);
  const long long total = static_cast<long long>(n) * ppi * c;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  ln_jvp_bwd_apply_kernel<<<static_cast<int>(g), 256, 0, s>>>(p, pdot, stats, n, c, ppi, scale, zhi, zlo,
                                                               hdot_bar->pix_stride, mask, alpha, tsums, asums,
                                                               static_cast<__nv_bfloat16*>(pdot_bar->hi),
                                                               static_cast<__nv_bfloat16*>(pdot_bar->lo),
                                                               pdot_bar->pix_stride, p_bar);
  ctx->launches += 2;
  return check_launch(ctx, "layernorm_jvp_bwd");
}
