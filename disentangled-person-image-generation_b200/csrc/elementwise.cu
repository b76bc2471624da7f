// HBM-bound glue kernels on split-bf16 NHWC tensors: gradient combination / ReLU masks / 2x2 pooling,
// fp32 <-> split conversion, foreground/background masking (models.py:402-403), embedding
// broadcast (trainer.py:588-590) and its gradient, bias gradients, weight packing.
// All kernels move 16 bytes per thread per plane (8 bf16 channels) with grid-stride loops.
#include <algorithm>
#include "common.cuh"

namespace dpig {

__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long off,
                                      float (&f)[8], bool accumulate) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  float t[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    t[2 * j] = bf16_bits_to_float(aw[j] & 0xFFFF);
    t[2 * j + 1] = bf16_bits_to_float(aw[j] >> 16);
  }
  if (lo) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + off));
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[2 * j] += bf16_bits_to_float(bw[j] & 0xFFFF);
      t[2 * j + 1] += bf16_bits_to_float(bw[j] >> 16);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = accumulate ? f[j] + t[j] : t[j];
}

__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off,
                                       const float (&f)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(f[2 * j], h0, l0);
    split_bf16(f[2 * j + 1], h1, l1);
    h[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    l[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

struct TView {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
  long long ps;
};

// colsum (optional, [C]): += the column sums over pixels of the values written -- the bias gradient of the layer whose
// output gradient this is, taken from the registers that hold them instead of a second read by dpig_bias_grad.  Needs
// gridDim.x * blockDim.x to be a multiple of C / 8 (the host rounds the grid), so that a thread keeps its channel group
// over the grid-stride loop; block-level combine in shared memory, one fp32 atomic per channel and block.
__global__ void __launch_bounds__(256)
ew_combine_kernel(__nv_bfloat16* ohi, __nv_bfloat16* olo, long long ops, int N, int H,
                                  int W, int C, TView a, TView b, TView c, const float* f32,
                                  long long f32_ps, const uint32_t* mask, float mask_neg, int pool2, float* colsum) {
  __shared__ float s_cs[256][8];
  const int C8 = C / 8;
  const long long total = static_cast<long long>(N) * H * W * C8;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const long long pix = i / C8;
    const int ch = c8 * 8;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int reps = pool2 ? 2 : 1;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    for (int dy = 0; dy < reps; ++dy)
      for (int dx = 0; dx < reps; ++dx) {
        const long long ip = pool2 ? ((static_cast<long long>(n) * (2 * H) + 2 * y + dy) * (2 * W) + 2 * x + dx) : pix;
        if (a.hi) load8(a.hi, a.lo, ip * a.ps + ch, f, true);
        if (b.hi) load8(b.hi, b.lo, ip * b.ps + ch, f, true);
        if (c.hi) load8(c.hi, c.lo, ip * c.ps + ch, f, true);
        if (f32) {
          const float* s = f32 + ip * f32_ps + ch;
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += __ldg(s + j);
        }
      }
    if (mask) {
      const uint32_t m = mask[pix * ((C + 31) / 32) + (ch >> 5)] >> (ch & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= ((m >> j) & 1u) ? 1.f : mask_neg;
    }
    store8(ohi, olo, pix * ops + ch, f);
    if (colsum) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] += f[j];
    }
  }
  if (colsum) {     // uniform over the block
#pragma unroll
    for (int j = 0; j < 8; ++j) s_cs[threadIdx.x][j] = cs[j];
    __syncthreads();
    // threads t, t + C8, t + 2*C8, ... of the block hold the same channel group (blockDim.x % C8 == 0 or C8 > blockDim.x)
    const int groups = C8 < 256 ? C8 : 256;
    if (static_cast<int>(threadIdx.x) < groups) {
      float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = threadIdx.x; t < 256; t += groups)
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[j] += s_cs[t][j];
      // this thread's channel group: that of its first element
      const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
      const int ch = static_cast<int>(i0 % C8) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(colsum + ch + j, tot[j]);
    }
  }
}

__global__ void pack_f32_kernel(const float* src, long long sps, int csrc, __nv_bfloat16* ohi,
                                __nv_bfloat16* olo, long long ops, long long pixels, int C) {
  const int C8 = C / 8;
  const long long total = pixels * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C8) * 8;
    const long long pix = i / C8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (ch + j < csrc) ? __ldg(src + pix * sps + ch + j) : 0.f;
    store8(ohi, olo, pix * ops + ch, f);
  }
}

__global__ void unpack_f32_kernel(TView s, float* dst, long long dps, long long pixels, int C) {
  const int C8 = C / 8;
  const long long total = pixels * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C8) * 8;
    const long long pix = i / C8;
    float f[8];
    load8(s.hi, s.lo, pix * s.ps + ch, f, false);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[pix * dps + ch + j] = f[j];
  }
}

__global__ void mask_split_kernel(TView x, const float* m, __nv_bfloat16* fhi, __nv_bfloat16* flo,
                                  long long fps, __nv_bfloat16* bhi, __nv_bfloat16* blo, long long bps,
                                  long long pixels, int C) {
  const int C8 = C / 8;
  const long long total = pixels * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C8) * 8;
    const long long pix = i / C8;
    float f[8], g[8];
    load8(x.hi, x.lo, pix * x.ps + ch, f, false);
    const float mv = __ldg(m + pix);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g[j] = f[j] * (1.0f - mv);
      f[j] = f[j] * mv;
    }
    if (fhi) store8(fhi, flo, pix * fps + ch, f);
    if (bhi) store8(bhi, blo, pix * bps + ch, g);
  }
}

__global__ void broadcast_emb_kernel(const float* emb, int ce, __nv_bfloat16* ohi, __nv_bfloat16* olo,
                                     long long ops, long long pix_per_img, long long pixels) {
  const int C8 = ce / 8;
  const long long total = pixels * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C8) * 8;
    const long long pix = i / C8;
    const long long n = pix / pix_per_img;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldg(emb + n * ce + ch + j);
    store8(ohi, olo, pix * ops + ch, f);
  }
}

// out[n][c] = sum over the image's pixels.  grid = (chunks, n); block = 256 threads.
__global__ void spatial_sum_kernel(TView g, float* out, int C, long long pix_per_img, int chunk) {
  const int n = blockIdx.y;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, pix_per_img);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const long long off = (n * pix_per_img + p) * g.ps + c;
      float v = __bfloat162float(g.hi[off]);
      if (g.lo) v += __bfloat162float(g.lo[off]);
      acc += v;
    }
    atomicAdd(out + static_cast<long long>(n) * C + c, acc);
  }
}

// db[c] += sum over pixels.  256 threads = (C/8 channel groups) x (rows of pixels); each thread streams
// 16-byte (8-channel) loads down its pixel rows, rows are combined through shared memory, then one fp32
// atomic per channel per block.  HBM-bound: reads every gradient element exactly once.
__global__ void __launch_bounds__(256)
bias_grad_kernel(TView g, const float* f32, long long f32_ps, float* db, int C, long long pixels, int chunk) {
  __shared__ float sh[256][8];
  const int C8 = (C + 7) / 8;
  const int groups = C8 < 256 ? C8 : 256;     // channel groups handled per pass
  const int rows = 256 / groups;
  const int cg = threadIdx.x % groups, prow = threadIdx.x / groups;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, pixels);
  for (int cg0 = 0; cg0 < C8; cg0 += groups) {
    const int ch = (cg0 + cg) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (prow < rows && ch < C) {
#pragma unroll 4
      for (long long p = p0 + prow; p < p1; p += rows) {
        if (f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (ch + j < C) acc[j] += __ldg(f32 + p * f32_ps + ch + j);
        } else if (ch + 8 <= C && (g.ps % 8 == 0)) {
          load8(g.hi, g.lo, p * g.ps + ch, acc, true);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (ch + j < C) {
              float v = __bfloat162float(g.hi[p * g.ps + ch + j]);
              if (g.lo) v += __bfloat162float(g.lo[p * g.ps + ch + j]);
              acc[j] += v;
            }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = acc[j];
    __syncthreads();
    if (prow == 0 && ch < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += sh[r * groups + cg][j];
        if (ch + j < C) atomicAdd(db + ch + j, s);
      }
    }
    __syncthreads();
  }
}

__global__ void act_bwd_f32_kernel(const float* y, float* dy, long long count, float neg) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dy[i] *= (y[i] > 0.f) ? 1.f : neg;
}

__global__ void denorm_u8_kernel(const float* g, long long count, uint8_t* out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = (g[i] + 1.f) * 127.5f;
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = static_cast<uint8_t>(v);  // tf.cast(float->uint8) truncates; save_image path
  }
}

// One tap at a time: w[tap][ci][co] -> fwd[tap][co][ci_pad] (transposed) and bwd[tap][ci][co_pad].
__global__ void weight_pack_kernel(const float* w, int cin, int cout, int cin_pad, int cout_pad,
                                   __nv_bfloat16* fhi, __nv_bfloat16* flo, __nv_bfloat16* bhi,
                                   __nv_bfloat16* blo, int cin_total) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const float* wt = w + static_cast<long long>(tap) * cin_total * cout;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    const float v = (ci < cin && co < cout) ? wt[static_cast<long long>(ci) * cout + co] : 0.f;
    tile[r][threadIdx.x] = v;
    if (bhi && ci < cin && co < cout_pad) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      const long long o = (static_cast<long long>(tap) * cin + ci) * cout_pad + co;
      bhi[o] = h;
      if (blo) blo[o] = l;
    }
  }
  __syncthreads();
  if (fhi) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int co = co0 + r, ci = ci0 + threadIdx.x;
      if (co < cout && ci < cin_pad) {
        __nv_bfloat16 h, l;
        split_bf16(tile[threadIdx.x][r], h, l);
        const long long o = (static_cast<long long>(tap) * cout + co) * cin_pad + ci;
        fhi[o] = h;
        if (flo) flo[o] = l;
      }
    }
  }
}

static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
static inline TView view(const dpig_tensor* t) {
  TView v{nullptr, nullptr, 0};
  if (t) {
    v.hi = static_cast<const __nv_bfloat16*>(t->hi);
    v.lo = static_cast<const __nv_bfloat16*>(t->lo);
    v.ps = t->pix_stride;
  }
  return v;
}
static inline bool aligned8(const dpig_tensor* t) {
  return !t || (t->c % 8 == 0 && t->pix_stride % 8 == 0 && reinterpret_cast<uintptr_t>(t->hi) % 16 == 0 &&
                reinterpret_cast<uintptr_t>(t->lo) % 16 == 0);
}

}  // namespace dpig
using namespace dpig;

static int ew_combine_impl(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a, const dpig_tensor* b,
                           const dpig_tensor* c, const float* f32, int64_t f32_pix_stride, const uint32_t* mask,
                           float mask_neg, int32_t pool2, float* colsum, dpig_stream stream);

extern "C" int dpig_ew_combine(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a,
                               const dpig_tensor* b, const dpig_tensor* c, const float* f32,
                               int64_t f32_pix_stride, const uint32_t* mask, float mask_neg,
                               int32_t pool2, dpig_stream stream) {
  return ew_combine_impl(ctx, out, a, b, c, f32, f32_pix_stride, mask, mask_neg, pool2, nullptr, stream);
}

extern "C" int dpig_ew_combine_colsum(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a,
                                      const dpig_tensor* b, const dpig_tensor* c, const float* f32,
                                      int64_t f32_pix_stride, const uint32_t* mask, float mask_neg,
                                      int32_t pool2, float* colsum, dpig_stream stream) {
  return ew_combine_impl(ctx, out, a, b, c, f32, f32_pix_stride, mask, mask_neg, pool2, colsum, stream);
}

static int ew_combine_impl(dpig_ctx* ctx, const dpig_tensor* out, const dpig_tensor* a, const dpig_tensor* b,
                           const dpig_tensor* c, const float* f32, int64_t f32_pix_stride, const uint32_t* mask,
                           float mask_neg, int32_t pool2, float* colsum, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!out || !out->hi) return set_error(ctx, DPIG_EINVAL, "ew_combine: null output");
  if (!aligned8(out) || !aligned8(a) || !aligned8(b) || !aligned8(c))
    return set_error(ctx, DPIG_EINVAL, "ew_combine: tensors must be 8-channel / 16-byte aligned");
  const int k = pool2 ? 2 : 1;
  const dpig_tensor* ins[3] = {a, b, c};
  for (auto t : ins)
    if (t && (t->n != out->n || t->h != out->h * k || t->w != out->w * k || t->c < out->c))
      return set_error(ctx, DPIG_EINVAL, "ew_combine: input shape mismatch");
  const long long total = static_cast<long long>(out->n) * out->h * out->w * (out->c / 8);
  int grid = grid_for(total);
  if (colsum) {
    // every block ends with one atomic per channel: 148 x 16 blocks put 2368 same-address atomics on each of a few
    // hundred addresses and cost more than the read pass this fusion removes (measured 1.19 ms against 1.03 ms for the
    // eight gradient tensors of an iteration); four blocks per SM keep the loads in flight at a quarter of the atomics
    grid = grid_for(total, 256, 148 * 4);
    // a thread must keep its channel group over the grid-stride loop: grid * 256 a multiple of C / 8
    const int c8 = out->c / 8;
    if (c8 > 256 || 256 % c8 != 0) {
      int g = 1;
      while ((static_cast<long long>(g) * 256) % c8) ++g;
      grid = std::max(g, grid / g * g);
    }
    if (c8 > 256)
      return set_error(ctx, DPIG_EUNSUPPORTED, "ew_combine_colsum: at most 2048 channels");
  }
  ew_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->pix_stride, out->n,
      out->h, out->w, out->c, view(a), view(b), view(c), f32, f32_pix_stride, mask, mask_neg, pool2, colsum);
  ctx->launches++;
  return check_launch(ctx, "ew_combine");
}

extern "C" int dpig_pack_f32(dpig_ctx* ctx, const float* src, int64_t src_pix_stride, int32_t c_src,
                             const dpig_tensor* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!src || !out || !aligned8(out)) return set_error(ctx, DPIG_EINVAL, "pack_f32: bad argument");
  const long long pixels = static_cast<long long>(out->n) * out->h * out->w;
  pack_f32_kernel<<<grid_for(pixels * (out->c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_pix_stride, c_src, static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo),
      out->pix_stride, pixels, out->c);
  ctx->launches++;
  return check_launch(ctx, "pack_f32");
}

extern "C" int dpig_unpack_f32(dpig_ctx* ctx, const dpig_tensor* src, float* dst, int64_t dst_pix_stride,
                               dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!src || !dst || !aligned8(src)) return set_error(ctx, DPIG_EINVAL, "unpack_f32: bad argument");
  const long long pixels = static_cast<long long>(src->n) * src->h * src->w;
  unpack_f32_kernel<<<grid_for(pixels * (src->c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      view(src), dst, dst_pix_stride, pixels, src->c);
  ctx->launches++;
  return check_launch(ctx, "unpack_f32");
}

extern "C" int dpig_mask_split(dpig_ctx* ctx, const dpig_tensor* x, const float* m, const dpig_tensor* fg,
                               const dpig_tensor* bg, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !m || !aligned8(x) || !aligned8(fg) || !aligned8(bg))
    return set_error(ctx, DPIG_EINVAL, "mask_split: bad argument");
  const long long pixels = static_cast<long long>(x->n) * x->h * x->w;
  mask_split_kernel<<<grid_for(pixels * (x->c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      view(x), m, fg ? static_cast<__nv_bfloat16*>(fg->hi) : nullptr,
      fg ? static_cast<__nv_bfloat16*>(fg->lo) : nullptr, fg ? fg->pix_stride : 0,
      bg ? static_cast<__nv_bfloat16*>(bg->hi) : nullptr, bg ? static_cast<__nv_bfloat16*>(bg->lo) : nullptr,
      bg ? bg->pix_stride : 0, pixels, x->c);
  ctx->launches++;
  return check_launch(ctx, "mask_split");
}

extern "C" int dpig_broadcast_embedding(dpig_ctx* ctx, const float* emb, int32_t ce, const dpig_tensor* out,
                                        dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!emb || !out || ce % 8 || !aligned8(out) || out->c < ce)
    return set_error(ctx, DPIG_EINVAL, "broadcast_embedding: bad argument");
  const long long ppi = static_cast<long long>(out->h) * out->w;
  const long long pixels = ppi * out->n;
  broadcast_emb_kernel<<<grid_for(pixels * (ce / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      emb, ce, static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->pix_stride,
      ppi, pixels);
  ctx->launches++;
  return check_launch(ctx, "broadcast_embedding");
}

extern "C" int dpig_spatial_sum(dpig_ctx* ctx, const dpig_tensor* g, float* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!g || !out) return set_error(ctx, DPIG_EINVAL, "spatial_sum: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(out, 0, sizeof(float) * g->n * g->c, s);
  const long long ppi = static_cast<long long>(g->h) * g->w;
  const int chunk = 64;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), g->n);
  spatial_sum_kernel<<<grid, 256, 0, s>>>(view(g), out, g->c, ppi, chunk);
  ctx->launches++;
  return check_launch(ctx, "spatial_sum");
}

extern "C" int dpig_bias_grad(dpig_ctx* ctx, const dpig_tensor* dy, float* db, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !db) return set_error(ctx, DPIG_EINVAL, "bias_grad: bad argument");
  const long long pixels = static_cast<long long>(dy->n) * dy->h * dy->w;
  int chunk = static_cast<int>((pixels + 148 * 4 - 1) / (148 * 4));
  if (chunk < 16) chunk = 16;
  bias_grad_kernel<<<static_cast<unsigned>((pixels + chunk - 1) / chunk), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      view(dy), nullptr, 0, db, dy->c, pixels, chunk);
  ctx->launches++;
  return check_launch(ctx, "bias_grad");
}

extern "C" int dpig_bias_grad_f32(dpig_ctx* ctx, const float* dy, int64_t pixels, int32_t c, float* db,
                                  dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !db) return set_error(ctx, DPIG_EINVAL, "bias_grad_f32: bad argument");
  int chunk = static_cast<int>((pixels + 148 * 4 - 1) / (148 * 4));
  if (chunk < 16) chunk = 16;
  TView none{nullptr, nullptr, 0};
  bias_grad_kernel<<<static_cast<unsigned>((pixels + chunk - 1) / chunk), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      none, dy, c, db, c, pixels, chunk);
  ctx->launches++;
  return check_launch(ctx, "bias_grad_f32");
}

extern "C" int dpig_act_bwd_f32(dpig_ctx* ctx, const float* y, float* dy, int64_t count, float neg,
                                dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  act_bwd_f32_kernel<<<grid_for(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, count, neg);
  ctx->launches++;
  return check_launch(ctx, "act_bwd_f32");
}

extern "C" int dpig_denorm_u8(dpig_ctx* ctx, const float* g, int64_t count, uint8_t* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  denorm_u8_kernel<<<grid_for(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, count, out);
  ctx->launches++;
  return check_launch(ctx, "denorm_u8");
}

extern "C" int dpig_weight_pack(dpig_ctx* ctx, const float* w_hwio, int32_t taps, int32_t cin, int32_t cout,
                                int32_t cin_pad, int32_t cout_pad, void* fwd_hi, void* fwd_lo, void* bwd_hi,
                                void* bwd_lo, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!w_hwio || cin_pad < cin || cout_pad < cout)
    return set_error(ctx, DPIG_EINVAL, "weight_pack: bad argument");
  dim3 grid((std::max(cout, cout_pad) + 31) / 32, (std::max(cin, cin_pad) + 31) / 32, taps);
  dim3 block(32, 8);
  weight_pack_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      w_hwio, cin, cout, cin_pad, cout_pad, static_cast<__nv_bfloat16*>(fwd_hi),
      static_cast<__nv_bfloat16*>(fwd_lo), static_cast<__nv_bfloat16*>(bwd_hi),
      static_cast<__nv_bfloat16*>(bwd_lo), cin);
  ctx->launches++;
  return check_launch(ctx, "weight_pack");
}

extern "C" int dpig_weight_pack_rows(dpig_ctx* ctx, const float* w_hwio, int32_t taps, int32_t cin_total, int32_t ci0,
                                     int32_t cin, int32_t cout, int32_t cin_pad, int32_t cout_pad, void* fwd_hi,
                                     void* fwd_lo, void* bwd_hi, void* bwd_lo, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!w_hwio || cin_pad < cin || cout_pad < cout || ci0 + cin > cin_total)
    return set_error(ctx, DPIG_EINVAL, "weight_pack_rows: bad argument");
  dim3 grid((std::max(cout, cout_pad) + 31) / 32, (std::max(cin, cin_pad) + 31) / 32, taps);
  dim3 block(32, 8);
  weight_pack_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      w_hwio + static_cast<long long>(ci0) * cout, cin, cout, cin_pad, cout_pad, static_cast<__nv_bfloat16*>(fwd_hi),
      static_cast<__nv_bfloat16*>(fwd_lo), static_cast<__nv_bfloat16*>(bwd_hi), static_cast<__nv_bfloat16*>(bwd_lo),
      cin_total);
  ctx->launches++;
  return check_launch(ctx, "weight_pack_rows");
}

namespace dpig {
// border class c = rh*3 + rw (r = 0 first, 1 interior, 2 last); tap (ky,kx) of a 3x3 SAME conv reads inside the
// image for row class rh iff ky != 0 when rh == 0 and ky != 2 when rh == 2 (same for columns).
__device__ __forceinline__ bool tap_valid(int cls, int tap, int H, int W) {
  const int rh = cls / 3, rw = cls % 3, ky = tap / 3, kx = tap % 3;
  (void)H;
  (void)W;  // H, W >= 2 (checked by the callers): every pixel has exactly one row class and one column class
  const bool vy = !((rh == 0 && ky == 0) || (rh == 2 && ky == 2));
  const bool vx = !((rw == 0 && kx == 0) || (rw == 2 && kx == 2));
  return vy && vx;
}
__global__ void stem_class_bias_kernel(const float* e, int N, int C, int H, int W, float* out) {
  const int total = N * 9 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, cls = (i / C) % 9, n = i / (9 * C);
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap)
      if (tap_valid(cls, tap, H, W)) acc += e[(static_cast<long long>(tap) * N + n) * C + c];
    out[i] = acc;
  }
}
// grid = (pixel chunks, N); 9 class bins per thread-channel kept in registers.
__global__ void stem_class_sum_kernel(TView g, int C, int H, int W, int chunk, float* out) {
  const int n = blockIdx.y;
  const long long ppi = static_cast<long long>(H) * W;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long p1 = min(p0 + chunk, ppi);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long p = p0; p < p1; ++p) {
      const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
      const int cls = (h == 0 ? 0 : (h == H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
      const long long off = (n * ppi + p) * g.ps + c;
      float v = __bfloat162float(g.hi[off]);
      if (g.lo) v += __bfloat162float(g.lo[off]);
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[k] += (k == cls) ? v : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (acc[k] != 0.f) atomicAdd(out + (static_cast<long long>(n) * 9 + k) * C + c, acc[k]);
  }
}
__global__ void stem_tap_sums_kernel(const float* cls_sums, int N, int C, int H, int W, float* taps) {
  const int total = 9 * N * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, n = (i / C) % N, tap = i / (N * C);
    float acc = 0.f;
    for (int cls = 0; cls < 9; ++cls)
      if (tap_valid(cls, tap, H, W)) acc += cls_sums[(static_cast<long long>(n) * 9 + cls) * C + c];
    taps[i] = acc;
  }
}
}  // namespace dpig

extern "C" int dpig_stem_class_bias(dpig_ctx* ctx, const float* e_taps, int32_t n, int32_t cout, int32_t h, int32_t w_,
                                    float* class_bias, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!e_taps || !class_bias || h < 2 || w_ < 2) return set_error(ctx, DPIG_EINVAL, "stem_class_bias: bad argument");
  const int total = n * 9 * cout;
  stem_class_bias_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(e_taps, n, cout, h, w_, class_bias);
  ctx->launches++;
  return check_launch(ctx, "stem_class_bias");
}

extern "C" int dpig_stem_tap_sums(dpig_ctx* ctx, const dpig_tensor* g, float* class_sums, float* tap_sums,
                                  dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!g || !class_sums || !tap_sums || g->h < 2 || g->w < 2) return set_error(ctx, DPIG_EINVAL, "stem_tap_sums: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(class_sums, 0, sizeof(float) * g->n * 9 * g->c, s);
  const long long ppi = static_cast<long long>(g->h) * g->w;
  const int chunk = 64;
  dim3 grid(static_cast<unsigned>((ppi + chunk - 1) / chunk), g->n);
  stem_class_sum_kernel<<<grid, 128, 0, s>>>(view(g), g->c, g->h, g->w, chunk, class_sums);
  const int total = 9 * g->n * g->c;
  stem_tap_sums_kernel<<<(total + 255) / 256, 256, 0, s>>>(class_sums, g->n, g->c, g->h, g->w, tap_sums);
  ctx->launches += 2;
  return check_launch(ctx, "stem_tap_sums");
}

namespace dpig {
// Body-part features -> embedding (reference models.py:433-442, 467-468): the 7 ROI feature blocks are the
// rows i*B..(i+1)*B of `fea` (ROIs are concatenated on the batch axis, models.py:420), each scaled by the
// part's visibility, followed by the background feature.  backward=1 routes the gradient the other way.
__global__ void emb_assemble_kernel(float* fea, float* bg, const float* vis, int B, int parts, int pz, int bgz,
                                    float* emb, int backward) {
  const int E = parts * pz + bgz;
  const int total = B * E;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / E, e = i % E;
    if (e < parts * pz) {
      const int part = e / pz, j = e % pz;
      const float v = vis[b * parts + part];
      float* f = fea + (static_cast<long long>(part) * B + b) * pz + j;
      if (backward) *f = emb[i] * v;
      else emb[i] = *f * v;
    } else {
      float* f = bg + static_cast<long long>(b) * bgz + (e - parts * pz);
      if (backward) *f = emb[i];
      else emb[i] = *f;
    }
  }
}
}  // namespace dpig

extern "C" int dpig_embedding_assemble(dpig_ctx* ctx, float* fea, float* bg, const float* vis, int32_t batch,
                                       int32_t parts, int32_t part_z, int32_t bg_z, float* emb, int32_t backward,
                                       dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!fea || (!bg && bg_z > 0) || !vis || !emb) return set_error(ctx, DPIG_EINVAL, "embedding_assemble: null argument");
  const int total = batch * (parts * part_z + bg_z);
  emb_assemble_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fea, bg, vis, batch, parts, part_z, bg_z, emb, backward);
  ctx->launches++;
  return check_launch(ctx, "embedding_assemble");
}

namespace dpig {
__global__ void add_f32_kernel(float* out, const float* a, const float* b, long long count, float sa, float sb) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = sa * a[i] + (b ? sb * b[i] : 0.f);
}
}  // namespace dpig

extern "C" int dpig_add_f32(dpig_ctx* ctx, float* out, const float* a, const float* b, int64_t count, float sa, float sb,
                            dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!out || !a) return set_error(ctx, DPIG_EINVAL, "add_f32: null argument");
  add_f32_kernel<<<grid_for(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, a, b, count, sa, sb);
  ctx->launches++;
  return check_launch(ctx, "add_f32");
}
