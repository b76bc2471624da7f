// Shared declarations for the DPIG hot-path kernels (internal; the public surface is include/dpig.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include "../../include/dpig.h"

struct dpig_ctx {
  int device = 0;
  int num_sms = 148;
  int max_smem_optin = 0;
  std::string last_error;
  // driver entry point, resolved at ctx creation (no link-time libcuda dependency)
  CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill) = nullptr;
  bool fast_mode = false;  // hi-plane only (single bf16 MMA pass); NOT the parity mode
  int pair_mode = 1;       // conv kernels as 2-CTA clusters (cta_group::2): 0 never, 1 where it wins, 2 wherever legal
  bool merge_planes = true;  // fetch hi+lo with one TMA instruction where the layout allows (DPIG_CONV_MERGE=0 disables)
  bool epi_tma = true;       // split-bf16 conv outputs leave the SM as TMA tensor stores (DPIG_EPI_TMA=0: per-lane stores)
  bool dgrad_merge = true;   // stride-2 data gradients: the four parity classes in one launch (DPIG_DGRAD_MERGE=0: four)
  bool add_prefetch = true;  // L2 prefetch of a tile's residual rows ahead of the epilogue (DPIG_ADD_PREFETCH=0: off)
  int tune_small = 1;        // channel block of few-pixel layers: 1 cost model, 2 the old halving rule, 0 always the widest (DPIG_TUNE_SMALL)
  int epi_bufs = 0;          // epilogue staging buffers per warp set: 0 auto, 1 always one, 2 two wherever two stages still fit (DPIG_EPI_BUFS)
  bool wide_b = true;        // hi|lo weight rows as one N = 2*block_n MMA operand for block_n <= 128 (DPIG_WIDE_B=0: three N = block_n MMAs)
  bool wgrad_pair = true;    // filter-gradient kernel as 2-CTA clusters (cta_group::2) where the shape allows (DPIG_WGRAD_PAIR=0: never)
  bool wgrad_split = true;   // filter gradients of 384- / 640- / 896-wide layers as 256-wide column segments + remainder (DPIG_WGRAD_SPLIT=0: one launch)
  bool wgrad_vec_red = true; // filter-gradient partials leave as 16-byte vector reductions (red.global.add.v4.f32) (DPIG_WGRAD_VEC_RED=0: scalar atomics)
  int wgrad_group = 0;     // filter taps per wgrad CTA: 0 = default (1); DPIG_WGRAD_GROUP
  int wgrad_px = 0;        // pixels per wgrad pipeline step: 0 = auto, 32 / 64 forced; DPIG_WGRAD_PX
  bool epi_specialise = true;  // conv launches run on the smallest epilogue instantiation covering them (DPIG_EPI_SPECIALISE=0: generic)
  bool crop_gather = true;  // crop_and_resize image gradient in gather form (DPIG_CROP_GATHER=0: atomic scatter)
  int max_stages = 0;      // >= 2: cap on the conv smem pipeline depth (DPIG_CONV_STAGES, tuning experiments)
  unsigned long long launches = 0;
};

namespace dpig {

int set_error(dpig_ctx* ctx, int code, const char* fmt, ...);
int check_launch(dpig_ctx* ctx, const char* what);

// Split-bf16 storage: x ~= float(hi) + float(lo), |x - hi - lo| <= 2^-18 |x|.
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
  return __uint_as_float(static_cast<uint32_t>(b) << 16);
}

// TF 'SAME' padding (tensorflow/core/framework/common_shape_fns.cc semantics):
// out = ceil(in / s), pad_total = max((out-1)*s + k - in, 0), pad_before = pad_total / 2.
inline int same_out(int in, int s) { return (in + s - 1) / s; }
inline int same_pad_before(int in, int k, int s) {
  int out = same_out(in, s);
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace dpig

#define DPIG_CHECK_CTX(ctx) \
  if (!(ctx)) return DPIG_EINVAL;
