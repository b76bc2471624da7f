// tf.image.crop_and_resize (bilinear, extrapolation_value = 0) and its image gradient, as used for the
// 7 body-part ROIs of the appearance encoder (reference models.py:406-415; DF: models.py:342-350).
// Sampling rule restated from TensorFlow 1.4 core/kernels/crop_and_resize_op.cc:
//   in_y = y1*(H-1) + y*(y2-y1)*(H-1)/(crop_h-1)   (crop_h > 1), 0.5*(y1+y2)*(H-1) otherwise;
//   rows / columns with in_y outside [0, H-1] (resp. in_x, W-1) produce the extrapolation value;
//   value = top + (bottom-top)*y_lerp with top = tl + (tr-tl)*x_lerp, bottom = bl + (br-bl)*x_lerp.
// The optional mask multiplies the image on the fly, so x_fg = x*m (models.py:402) is never stored.
#include "common.cuh"

namespace dpig {

struct CropGeom {
  int N, H, W, C, CH, CW;
};

__device__ __forceinline__ bool sample_coords(const float* box, int y, int x, const CropGeom& g, int& top,
                                              int& bottom, int& left, int& right, float& ylerp, float& xlerp) {
  const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
  const float hs = (g.CH > 1) ? (y2 - y1) * (g.H - 1) / (g.CH - 1) : 0.f;
  const float ws = (g.CW > 1) ? (x2 - x1) * (g.W - 1) / (g.CW - 1) : 0.f;
  const float in_y = (g.CH > 1) ? y1 * (g.H - 1) + y * hs : 0.5f * (y1 + y2) * (g.H - 1);
  if (in_y < 0.f || in_y > g.H - 1) return false;
  const float in_x = (g.CW > 1) ? x1 * (g.W - 1) + x * ws : 0.5f * (x1 + x2) * (g.W - 1);
  if (in_x < 0.f || in_x > g.W - 1) return false;
  top = static_cast<int>(floorf(in_y));
  bottom = static_cast<int>(ceilf(in_y));
  left = static_cast<int>(floorf(in_x));
  right = static_cast<int>(ceilf(in_x));
  ylerp = in_y - top;
  xlerp = in_x - left;
  return true;
}

__device__ __forceinline__ float ld_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long off) {
  float v = __bfloat162float(hi[off]);
  if (lo) v += __bfloat162float(lo[off]);
  return v;
}

__device__ __forceinline__ void ld8_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long off, float scale,
                                          float w, float (&acc)[8]) {
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
  for (int j = 0; j < 8; ++j) acc[j] = fmaf(t[j] * scale, w, acc[j]);
}

// thread = (box, y, x, 8 channels): 16-byte gathers of the 4 bilinear corners, 16-byte stores.
// The lerp is evaluated as a weighted sum; with TF's form top + (bottom-top)*yl the results agree to fp32 rounding.
__global__ void crop_resize_fwd_kernel(const __nv_bfloat16* ihi, const __nv_bfloat16* ilo, long long ips,
                                       const float* mask, const float* boxes, const int* box_ind, int nbox,
                                       CropGeom g, __nv_bfloat16* ohi, __nv_bfloat16* olo, long long ops) {
  const int C8 = g.C / 8;
  const long long total = static_cast<long long>(nbox) * g.CH * g.CW * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C8) * 8;
    const long long opix = i / C8;
    const int x = static_cast<int>(opix % g.CW);
    const int y = static_cast<int>((opix / g.CW) % g.CH);
    const int b = static_cast<int>(opix / (static_cast<long long>(g.CW) * g.CH));
    const int n = box_ind[b];
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int t, bo, l, r;
    float yl, xl;
    if (n >= 0 && n < g.N && sample_coords(boxes + 4 * b, y, x, g, t, bo, l, r, yl, xl)) {
      const long long base = static_cast<long long>(n) * g.H * g.W;
      const long long ptl = base + static_cast<long long>(t) * g.W + l, ptr = base + static_cast<long long>(t) * g.W + r;
      const long long pbl = base + static_cast<long long>(bo) * g.W + l, pbr = base + static_cast<long long>(bo) * g.W + r;
      const float mtl = mask ? mask[ptl] : 1.f, mtr = mask ? mask[ptr] : 1.f;
      const float mbl = mask ? mask[pbl] : 1.f, mbr = mask ? mask[pbr] : 1.f;
      ld8_split(ihi, ilo, ptl * ips + c, mtl, (1.f - xl) * (1.f - yl), v);
      ld8_split(ihi, ilo, ptr * ips + c, mtr, xl * (1.f - yl), v);
      ld8_split(ihi, ilo, pbl * ips + c, mbl, (1.f - xl) * yl, v);
      ld8_split(ihi, ilo, pbr * ips + c, mbr, xl * yl, v);
    }
    uint32_t h[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * j], h0, l0);
      split_bf16(v[2 * j + 1], h1, l1);
      h[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      lw[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    *reinterpret_cast<uint4*>(ohi + opix * ops + c) = make_uint4(h[0], h[1], h[2], h[3]);
    if (olo) *reinterpret_cast<uint4*>(olo + opix * ops + c) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

__global__ void crop_resize_bwd_kernel(const __nv_bfloat16* ghi, const __nv_bfloat16* glo, long long gps,
                                       const float* mask, const float* boxes, const int* box_ind, int nbox,
                                       CropGeom g, float* gimg) {
  const long long total = static_cast<long long>(nbox) * g.CH * g.CW * g.C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % g.C);
    const long long opix = i / g.C;
    const int x = static_cast<int>(opix % g.CW);
    const int y = static_cast<int>((opix / g.CW) % g.CH);
    const int b = static_cast<int>(opix / (static_cast<long long>(g.CW) * g.CH));
    const int n = box_ind[b];
    int t, bo, l, r;
    float yl, xl;
    if (n < 0 || n >= g.N || !sample_coords(boxes + 4 * b, y, x, g, t, bo, l, r, yl, xl)) continue;
    const float gv = ld_split(ghi, glo, opix * gps + c);
    const long long base = static_cast<long long>(n) * g.H * g.W;
    const long long ptl = base + static_cast<long long>(t) * g.W + l, ptr = base + static_cast<long long>(t) * g.W + r;
    const long long pbl = base + static_cast<long long>(bo) * g.W + l, pbr = base + static_cast<long long>(bo) * g.W + r;
    const float dtop = (1.f - yl) * gv, dbot = yl * gv;
    float wtl = (1.f - xl) * dtop, wtr = xl * dtop, wbl = (1.f - xl) * dbot, wbr = xl * dbot;
    if (mask) {
      wtl *= mask[ptl];
      wtr *= mask[ptr];
      wbl *= mask[pbl];
      wbr *= mask[pbr];
    }
    atomicAdd(gimg + ptl * g.C + c, wtl);
    atomicAdd(gimg + ptr * g.C + c, wtr);
    atomicAdd(gimg + pbl * g.C + c, wbl);
    atomicAdd(gimg + pbr * g.C + c, wbr);
  }
}

// Gather form of the image gradient: thread = (image pixel, 8 channels) sums every crop sample whose bilinear footprint
// covers the pixel -- no atomics (the scatter form above issues 4 fp32 atomics per gradient element and ran at 10 % of
// the HBM roofline), no zero fill of the 0.27 GB gradient image, a fixed summation order (boxes ascending, crop rows,
// crop columns), one 32-byte store per thread.
//   The sample position of crop row y is in_y(y) = a + y*hs (sample_coords); it touches image rows floor(in_y) with weight
//   1 - frac and ceil(in_y) with weight frac.  The candidate rows for image row Y are those with in_y in (Y-1, Y+1): the
//   range is bracketed from the affine map (one row of slack) and every candidate re-evaluates in_y with the forward
//   kernel's expression, so both directions agree on floor / ceil bit for bit.
// A block serves one image; its boxes (box_ind[b] == n) are compacted in ascending order into shared memory, kCropListCap
// at a time.
constexpr int kCropListCap = 256;

struct CropBoxAffine {   // per box of the block's image: sample position = a + index * s, and the pixels it can touch
  float ay, hs, ax, ws, rhs, rws;
  int b;
  short ymin, ymax, xmin, xmax;
};

// candidate crop indices whose sample position lies within one pixel of Y (one index of slack on both sides; every
// candidate is re-checked with the forward kernel's expression)
__device__ __forceinline__ void cand_range(float a, float s, float rs, int crop, int Y, int& lo, int& hi) {
  if (crop <= 1) {
    lo = hi = 0;
    return;
  }
  if (s == 0.f) {
    lo = 0;
    hi = crop - 1;
    return;
  }
  float t0 = (static_cast<float>(Y) - 1.f - a) * rs, t1 = (static_cast<float>(Y) + 1.f - a) * rs;
  if (t0 > t1) {
    const float t = t0;
    t0 = t1;
    t1 = t;
  }
  t0 = fminf(fmaxf(floorf(t0) - 1.f, 0.f), static_cast<float>(crop));       // clamp before the int conversion
  t1 = fminf(fmaxf(ceilf(t1) + 1.f, -1.f), static_cast<float>(crop - 1));
  lo = static_cast<int>(t0);
  hi = static_cast<int>(t1);
}

__global__ void __launch_bounds__(256)
crop_resize_bwd_gather_kernel(const __nv_bfloat16* ghi, const __nv_bfloat16* glo, long long gps, const float* mask,
                              const float* boxes, const int* box_ind, int nbox, CropGeom g, float* gimg) {
  __shared__ int s_list[kCropListCap];
  __shared__ CropBoxAffine s_box[kCropListCap];
  __shared__ int s_cnt, s_next;
  const int n = blockIdx.y;
  const int C8 = g.C / 8;
  const long long item = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;   // (pixel of image n, channel group)
  const bool live = item < static_cast<long long>(g.H) * g.W * C8;
  const int c = static_cast<int>(item % C8) * 8;
  const int pin = static_cast<int>(item / C8);
  const int X = pin % g.W, Y = pin / g.W;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int scanned = 0;
  while (scanned < nbox) {
    if (threadIdx.x < 32) {   // ordered compaction of this image's boxes, 32 candidates per round
      int cnt = 0, b0 = scanned;
      for (; b0 < nbox; b0 += 32) {
        const int b = b0 + threadIdx.x;
        const bool hit = b < nbox && box_ind[b] == n;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int k = __popc(m);
        if (cnt + k > kCropListCap) break;
        if (hit) s_list[cnt + __popc(m & ((1u << threadIdx.x) - 1u))] = b;
        cnt += k;
      }
      if (threadIdx.x == 0) {
        s_cnt = cnt;
        s_next = b0 < nbox ? b0 : nbox;
      }
    }
    __syncthreads();
    const int cnt = s_cnt;
    scanned = s_next;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {   // the affine sample maps of the boxes, once per block
      const int b = s_list[k];
      const float* box = boxes + 4 * b;
      const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
      CropBoxAffine q;
      q.b = b;
      q.hs = (g.CH > 1) ? (y2 - y1) * (g.H - 1) / (g.CH - 1) : 0.f;
      q.ws = (g.CW > 1) ? (x2 - x1) * (g.W - 1) / (g.CW - 1) : 0.f;
      q.ay = (g.CH > 1) ? y1 * (g.H - 1) : 0.5f * (y1 + y2) * (g.H - 1);
      q.ax = (g.CW > 1) ? x1 * (g.W - 1) : 0.5f * (x1 + x2) * (g.W - 1);
      q.rhs = q.hs != 0.f ? 1.f / q.hs : 0.f;
      q.rws = q.ws != 0.f ? 1.f / q.ws : 0.f;
      const float ye = q.ay + (g.CH - 1) * q.hs, xe = q.ax + (g.CW - 1) * q.ws;   // last sample (first: ay / ax)
      q.ymin = static_cast<short>(fminf(fmaxf(floorf(fminf(q.ay, ye)) - 1.f, 0.f), 32767.f));
      q.ymax = static_cast<short>(fminf(fmaxf(ceilf(fmaxf(q.ay, ye)) + 1.f, -1.f), static_cast<float>(g.H - 1)));
      q.xmin = static_cast<short>(fminf(fmaxf(floorf(fminf(q.ax, xe)) - 1.f, 0.f), 32767.f));
      q.xmax = static_cast<short>(fminf(fmaxf(ceilf(fmaxf(q.ax, xe)) + 1.f, -1.f), static_cast<float>(g.W - 1)));
      s_box[k] = q;
    }
    __syncthreads();
    if (live) {
      for (int k = 0; k < cnt; ++k) {
        const CropBoxAffine& q = s_box[k];
        if (Y < q.ymin || Y > q.ymax || X < q.xmin || X > q.xmax) continue;   // most boxes miss most pixels
        const float* box = boxes + 4 * q.b;
        const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
        const float hs = q.hs, ws = q.ws;
        int ylo, yhi, xlo, xhi;
        cand_range(q.ay, hs, q.rhs, g.CH, Y, ylo, yhi);
        cand_range(q.ax, ws, q.rws, g.CW, X, xlo, xhi);
        for (int y = ylo; y <= yhi; ++y) {
          const float in_y = (g.CH > 1) ? y1 * (g.H - 1) + y * hs : 0.5f * (y1 + y2) * (g.H - 1);
          if (in_y < 0.f || in_y > g.H - 1) continue;
          const int t = static_cast<int>(floorf(in_y)), bo = static_cast<int>(ceilf(in_y));
          if (t != Y && bo != Y) continue;
          const float yl = in_y - t;
          const float wy = (t == Y ? 1.f - yl : 0.f) + (bo == Y ? yl : 0.f);
          for (int x = xlo; x <= xhi; ++x) {
            const float in_x = (g.CW > 1) ? x1 * (g.W - 1) + x * ws : 0.5f * (x1 + x2) * (g.W - 1);
            if (in_x < 0.f || in_x > g.W - 1) continue;
            const int l = static_cast<int>(floorf(in_x)), r = static_cast<int>(ceilf(in_x));
            if (l != X && r != X) continue;
            const float xl = in_x - l;
            const float wx = (l == X ? 1.f - xl : 0.f) + (r == X ? xl : 0.f);
            const long long opix = (static_cast<long long>(q.b) * g.CH + y) * g.CW + x;
            ld8_split(ghi, glo, opix * gps + c, 1.f, wy * wx, acc);
          }
        }
      }
    }
    __syncthreads();
  }
  if (live) {
    const long long pix = static_cast<long long>(n) * g.H * g.W + pin;
    const float m = mask ? mask[pix] : 1.f;
    float4* o = reinterpret_cast<float4*>(gimg + pix * g.C + c);
    o[0] = make_float4(acc[0] * m, acc[1] * m, acc[2] * m, acc[3] * m);
    o[1] = make_float4(acc[4] * m, acc[5] * m, acc[6] * m, acc[7] * m);
  }
}

}  // namespace dpig
using namespace dpig;

extern "C" int dpig_crop_and_resize_fwd(dpig_ctx* ctx, const dpig_tensor* image, const float* mask,
                                        const float* boxes, const int32_t* box_ind, int32_t nbox,
                                        const dpig_tensor* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!image || !boxes || !box_ind || !out || out->n != nbox || out->c != image->c)
    return set_error(ctx, DPIG_EINVAL, "crop_and_resize_fwd: bad argument");
  if (image->c % 8 || image->pix_stride % 8 || out->pix_stride % 8)
    return set_error(ctx, DPIG_EINVAL, "crop_and_resize_fwd: channels / pixel strides must be multiples of 8");
  CropGeom g{image->n, image->h, image->w, image->c, out->h, out->w};
  const long long total = static_cast<long long>(nbox) * g.CH * g.CW * (g.C / 8);
  long long grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  crop_resize_fwd_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(image->hi), static_cast<const __nv_bfloat16*>(image->lo),
      image->pix_stride, mask, boxes, box_ind, nbox, g, static_cast<__nv_bfloat16*>(out->hi),
      static_cast<__nv_bfloat16*>(out->lo), out->pix_stride);
  ctx->launches++;
  return check_launch(ctx, "crop_resize_fwd");
}

extern "C" int dpig_crop_and_resize_bwd(dpig_ctx* ctx, const dpig_tensor* grad, const float* mask,
                                        const float* boxes, const int32_t* box_ind, int32_t nbox,
                                        float* grad_image, int32_t n, int32_t h, int32_t w_, int32_t c,
                                        dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!grad || !boxes || !box_ind || !grad_image || grad->n != nbox || grad->c != c)
    return set_error(ctx, DPIG_EINVAL, "crop_and_resize_bwd: bad argument");
  CropGeom g{n, h, w_, c, grad->h, grad->w};
  if (ctx->crop_gather && c % 8 == 0 && grad->pix_stride % 8 == 0 && reinterpret_cast<uintptr_t>(grad_image) % 16 == 0 &&
      n <= 65535) {
    // gather form: grad_image is overwritten (the zero fill the scatter form needs is harmless but not required)
    const long long items = static_cast<long long>(h) * w_ * (c / 8);
    dim3 grid2(static_cast<unsigned>((items + 255) / 256), n);
    crop_resize_bwd_gather_kernel<<<grid2, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(grad->hi), static_cast<const __nv_bfloat16*>(grad->lo), grad->pix_stride, mask,
        boxes, box_ind, nbox, g, grad_image);
    ctx->launches++;
    return check_launch(ctx, "crop_resize_bwd_gather");
  }
  const long long total = static_cast<long long>(nbox) * g.CH * g.CW * g.C;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  crop_resize_bwd_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(grad->hi), static_cast<const __nv_bfloat16*>(grad->lo),
      grad->pix_stride, mask, boxes, box_ind, nbox, g, grad_image);
  ctx->launches++;
  return check_launch(ctx, "crop_resize_bwd");
}
