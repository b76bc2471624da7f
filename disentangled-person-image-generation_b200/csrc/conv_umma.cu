// Tensor-core convolutions for sm_100a: TMA-staged implicit GEMM on tcgen05 with TMEM accumulators.
//
// Replaces the TensorFlow kernels behind slim.conv2d (reference models.py:396-399, 425-429,
// 458-462, 528-539, 564-573) and tflib/ops/conv2d.py:106-120, forward and both gradients.
//
// Numerics: operands are split-bf16 planes (x = hi + lo).  Each 64-deep K step issues
// hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator, i.e. a ~2^-17-accurate product, so the
// result matches an fp32 convolution to ~1e-5 relative (see DESIGN.md "precision").
//
// conv kernel (forward and data-gradient):
//   GEMM M = a box of <=128 output pixels (BW x BH x BN), N = block_n output channels,
//   K = (filter taps) x (input channels in chunks of 64).
//   A tile  : one TMA box of the NHWC activation per (tap, chunk), hi and lo planes in one instruction; the tap shift is
//             applied to the box coordinates and TMA's out-of-bounds zero fill *is* the SAME padding.
//             Stride-2 convs read one of four parity views of the input (src index), so every
//             load is a plain dense box.
//   B tile  : TMA box of the packed weights [tap][n][k] (K-major), hi rows then lo rows.
//   both land in 128B-swizzled shared memory and are consumed by tcgen05.mma via descriptors.
//   Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2..9 = epilogue
//   (tcgen05.ld -> bias/act/residual/mask -> split-bf16 rows staged in smem -> TMA tensor stores).
//   Variants: <pair> 2-CTA clusters with cta_group::2 M=256 MMAs; <wide> hi|lo weight rows as one N = 2*block_n operand;
//   stride-2 data gradients run their four parity classes in one launch (ConvUmmaParams::nclass).
//
// wgrad kernel (filter gradient):
//   GEMM M = 128 input channels, N = block_n output channels, K = pixels (64 per step).
//   Both operands are MN-major views of the same kind of TMA boxes (rows = pixels).
//   One filter tap per CTA, split-K over pixel tiles (blockIdx.z), fp32 atomics into the HWIO gradient;
//   <pair>: two (tap, channel-tile) units per 2-CTA cluster share one dy tile (M=256 cta_group::2 MMAs).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "ptx.cuh"

namespace dpig {

struct ConvTap {
  int16_t src, dh, dw, wtap;
};
constexpr int kMaxTaps = 25;
constexpr uint32_t kABytes = 16384;  // 128 rows x 128 B, per plane per stage

struct ConvUmmaParams {
  CUtensorMap a_map[4][2];
  CUtensorMap b_map[2];
  ConvTap taps[kMaxTaps];
  int num_taps, planes, kchunks;
  // Parity classes of a stride-2 data gradient merged into ONE launch: a work unit runs its pixel tile once per class,
  // back to back (taps [cls_tap0[c], cls_tap0[c+1]), output offset (cls_oh[c], cls_ow[c])).  The same dy region then
  // feeds all four classes from L2 while the four interleaved quarters of the dx region are written close in time;
  // every unit carries all kh*kw taps, so the static tile schedule stays balanced.  nclass = 1 otherwise.
  int nclass;
  int16_t cls_tap0[5], cls_oh[4], cls_ow[4];
  int BW, BH, BN;
  int tiles_w, tiles_h, tiles_n;
  int Wo, Ho, No;
  int block_n, cout, stages;
  // wide-B form of the three split-bf16 passes (single-CTA kernel, block_n <= 128): the hi and lo weight rows sit back
  // to back in a stage, so ONE N = 2*block_n MMA computes A_hi*B_hi | A_hi*B_lo into two column ranges of the
  // accumulator and a second N = block_n MMA adds A_lo*B_hi; the epilogue sums the two ranges.  Same tensor time, but
  // the A_hi tile is read from shared memory once instead of twice: 20 KB instead of 24 KB per 16-deep K slice of a
  // 128-wide tile, whose MMA rate is capped by the 128 B/clk shared-memory port.
  int wide_b;
  int pair;  // 1: launched as 2-CTA clusters running cta_group::2 MMAs (b_bytes = this CTA's half of the B rows)
  // hi and lo planes fetched by ONE TMA instruction (map slot [..][0] then has a trailing "plane" dimension)
  int a_merged, b_merged;
  uint32_t a_tx_bytes, b_bytes, tmem_cols;
  const float* bias;
  const float* class_bias;  // [N][9][cout]: per-image bias indexed by the pixel's border class (stem shortcut)
  int act;
  float alpha;
  int sh, sw, oh, ow, rep, out_H, out_W;
  __nv_bfloat16 *out_hi, *out_lo;
  long long out_ps;
  __nv_bfloat16 *out2_hi, *out2_lo;
  long long out2_ps;
  float* out_f32;
  long long out_f32_ps;
  const __nv_bfloat16 *add_hi, *add_lo;
  long long add_ps;
  uint32_t* mask_out;
  int mask_out_words;
  const uint32_t* mask_in;
  int mask_in_words;
  float mask_neg;
  float* colsum;  // [cout] += column sums (over pixels) of the masked output: the bias gradient of the layer below
  // raw normalisation sums of pre = acc + bias (dpig_conv_epilogue::stat_sums): [2][stat_groups] fp64, per output channel
  // (stat_mode DPIG_NORM_BATCH) or per image (DPIG_NORM_LAYER)
  double* stat_sums;
  int stat_mode, stat_groups;
  // TMA tensor-store maps of the split-bf16 outputs, [index][plane hi / lo]; index = parity class of a merged stride-2
  // data gradient, or 2x2 replica of the fused upsample, else 0.  Box {32 channels, BW, BH, BN}, SWIZZLE_64B.
  CUtensorMap out_map[4][2];
  CUtensorMap out2_map[4][2];
  int out_tma, out2_tma;
  int epi_bufs;  // staging buffers per warp set (1 or 2), used alternately by successive TMA stores
  int add_prefetch;  // pull the residual rows of a tile into L2 before its accumulator is awaited
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) |
         (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

// Store 32 consecutive channels of one pixel as split bf16 (hi / lo planes).
__device__ __forceinline__ void store_split32(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off,
                                              const float (&f)[32], int nvalid, bool vec) {
  if (vec) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(f[2 * i], h0, l0);
      split_bf16(f[2 * i + 1], h1, l1);
      h[i] = pack2(h0, h1);
      l[i] = pack2(l0, l1);
    }
    uint4* ph = reinterpret_cast<uint4*>(hi + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) ph[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    if (lo) {
      uint4* pl = reinterpret_cast<uint4*>(lo + off);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pl[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < nvalid) {
        __nv_bfloat16 h0, l0;
        split_bf16(f[i], h0, l0);
        hi[off + i] = h0;
        if (lo) lo[off + i] = l0;
      }
    }
  }
}

// Two fp32 values -> packed bf16x2 hi word and lo word of their split representation.  cvt.rn.bf16x2.f32 (F2FP, the
// half-rate ALU pipe) converts a pair per instruction; the per-element F2F.BF16.F32 the scalar form compiles to sits on
// the quarter-rate conversion pipe and was 21 % of all stall samples of an epilogue-bound launch
// (profiles/r02_epilogue_ncu.txt).  Same round-to-nearest-even values as split_bf16().
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& h, uint32_t& l) {
  const __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);     // .x (low half) = x0
  h = *reinterpret_cast<const uint32_t*>(&hb);
  const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xFFFF0000u);
  const __nv_bfloat162 lb = __floats2bfloat162_rn(r0, r1);
  l = *reinterpret_cast<const uint32_t*>(&lb);
}

// Epilogue feature groups a kernel instantiation carries (template parameter kF of the conv kernel).  A launch runs on
// the smallest instantiation that covers what its epilogue asks for; features outside kF are compiled out -- their
// code, their branches and their registers -- instead of being skipped at run time by every chunk of every tile
// (~130 of the 686 instructions per 32-channel chunk were control flow around absent features).
enum : int {
  EF_FWD = 1,     // bias, per-pixel class bias, activation, sign-mask output          (forward-type)
  EF_GRAD = 2,    // second output x incoming sign mask, bias-gradient column sums     (data-gradient-type)
  EF_F32 = 4,     // fp32 output
  EF_STATS = 8,   // normalisation sums
  EF_ADD = 16,    // residual / addend
  EF_ALL = 31
};
constexpr int kEpiFwd = EF_FWD | EF_ADD, kEpiGrad = EF_GRAD | EF_ADD;

constexpr int kEpiWarps = 8;          // two epilogue warps per TMEM lane group: they take alternate 32-column chunks
constexpr int kConvThreads = 64 + 32 * kEpiWarps;
// Epilogue staging in shared memory.  The eight epilogue warps form two SETS (the warps taking the even / the odd
// 32-column chunks; each set holds one warp per TMEM lane group = all 128 pixel rows of the tile).  One 16 KB buffer per
// set: [hi plane: 128 rows x 64 B][lo plane: 128 rows x 64 B]; warp lg owns rows lg*32 .. lg*32+31 of both planes.
// 16-byte piece q of row r sits at r*64 + ((q ^ ((r >> 1) & 3)) << 4): that is TMA's SWIZZLE_64B pattern (buffers are
// 1024-byte aligned), so the same buffer is bank-conflict free for "thread = row" accesses AND a valid source box
// {32 channels, BW, BH, BN} of a TMA tensor store -- the split-bf16 outputs leave the SM as two bulk tensor stores per
// chunk issued by one thread (TMA clips at the tensor edge and walks the parity-strided outputs of stride-2 gradients /
// the 2x2 replicas of the fused upsample) instead of 8 shuffles + 8 predicated 16-byte stores per lane.  Measured before
// the change (tests/bench_dgrad_micro.py, profiles/r01_dgrad_epilogue_micro.txt): every split output cost ~2.6k clk per
// chunk and warp -- short-K layers (stride-2 parity classes, 1x1) ran at the epilogue's speed, not the tensor pipe's.
// With shared memory to spare a set gets TWO such buffers and alternates between them: a TMA store queues behind the
// producer's in-flight loads and needs ~2k clk until its source is free again, which a single buffer exposes on every
// chunk of a layer with two outputs or a short K loop.
constexpr uint32_t kEpiSetBytes = 16384, kEpiLoOff = 8192;
constexpr uint32_t kEpiBytes = 2 * kEpiSetBytes;   // both sets, one buffer each

// Byte offset of 16-byte piece `piece` of row `row` (relative to the warp's first row; warps start 2048 B apart).
__device__ __forceinline__ uint32_t swz64(int row, int piece) {   // 64 B rows, 4 pieces (split-bf16 planes)
  return row * 64 + ((piece ^ ((row >> 1) & 3)) << 4);
}
// fp32 staging (128 B rows, warp-private, never a TMA source): rows 0-15 in the warp's hi slot, 16-31 in its lo slot
__device__ __forceinline__ uint32_t swz128(int row, int piece) {
  return ((row & 16) ? kEpiLoOff : 0u) + (row & 15) * 128 + ((piece ^ (row & 7)) << 4);
}

// thread's 32 channels -> its staging row (hi plane at +0, lo plane at +kEpiLoOff)
__device__ __forceinline__ void epi_stage_split(uint32_t st, const float (&f)[32], int lane, bool with_lo) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) split_pack2(f[q * 8 + 2 * jj], f[q * 8 + 2 * jj + 1], h[jj], l[jj]);
    ptx::sts128(st + swz64(lane, q), make_uint4(h[0], h[1], h[2], h[3]));
    if (with_lo) ptx::sts128(st + kEpiLoOff + swz64(lane, q), make_uint4(l[0], l[1], l[2], l[3]));
  }
}

// staging rows -> global without TMA (outputs whose layout TMA cannot address): each instruction writes 8 pixel rows
// x 64 contiguous bytes per plane
__device__ __forceinline__ void epi_scatter_rows(uint32_t st, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                                 long long ps, int cbase, int pix, bool valid, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = 8 * i + (lane >> 2), pc = lane & 3;
    const int pq = __shfl_sync(0xffffffffu, pix, q);
    const int vq = __shfl_sync(0xffffffffu, static_cast<int>(valid), q);
    if (vq) {
      const long long off = static_cast<long long>(pq) * ps + cbase + pc * 8;
      *reinterpret_cast<uint4*>(hi + off) = ptx::lds128(st + swz64(q, pc));
      if (lo) *reinterpret_cast<uint4*>(lo + off) = ptx::lds128(st + kEpiLoOff + swz64(q, pc));
    }
  }
}

// global -> staging rows (residual / addend), same access shape as epi_scatter_rows.  All eight 16-byte loads are
// issued before the first staging store: in program order (load, load, store) x 4 the compiler kept each store behind
// its loads and the next loads behind the store, i.e. four exposed global latencies per chunk (30 % of all stall
// samples of a 128-channel residual layer, profiles/r01_epilogue_ncu.txt).
__device__ __forceinline__ void epi_gather_rows(uint32_t st, const __nv_bfloat16* hi, const __nv_bfloat16* lo,
                                                long long ps, int cbase, int pix, bool valid, int lane) {
  uint4 a[4], b[4];
  const int pc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = 8 * i + (lane >> 2);
    const int pq = __shfl_sync(0xffffffffu, pix, q);
    const int vq = __shfl_sync(0xffffffffu, static_cast<int>(valid), q);
    a[i] = make_uint4(0, 0, 0, 0);
    b[i] = make_uint4(0, 0, 0, 0);
    if (vq) {
      const long long off = static_cast<long long>(pq) * ps + cbase + pc * 8;
      a[i] = __ldg(reinterpret_cast<const uint4*>(hi + off));
      if (lo) b[i] = __ldg(reinterpret_cast<const uint4*>(lo + off));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = 8 * i + (lane >> 2);
    ptx::sts128(st + swz64(q, pc), a[i]);
    if (lo) ptx::sts128(st + kEpiLoOff + swz64(q, pc), b[i]);
  }
}

// Column sums over the warp's 32 rows (lane = row, f = the row's 32 channels), transposed: lane l returns the total of
// channel l.  31 shuffles: after step `off` every lane keeps the half of its values whose channel index has bit `off`
// equal to its lane-id bit.  Destroys f.
__device__ __forceinline__ float warp_colsum32(float (&f)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? f[i] : f[i + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      f[i] = (upper ? f[i + off] : f[i]) + recv;
    }
  }
  return f[0];
}

// What an epilogue warp knows about the tile it is draining.
struct EpiTile {
  int w0, h0, n0;     // origin of the pixel box in the (class) output grid = TMA store coordinates
  int cls;            // parity class (merged stride-2 data gradient), else 0
  uint32_t set_base;  // shared address of the warp set's staging buffer 0 (buffer 1 is kEpiBytes further)
  uint32_t bar_id;    // named barrier of the set (4 warps)
  bool leader;        // the one thread of the set that issues / tracks the TMA stores
  uint32_t phase;     // TMA-store phases issued so far by the set: the next staging rows go to buffer phase % bufs
  bool pending;       // stores issued since the last acquire may still be reading the buffer about to be written
};

// One 32-channel chunk of one accumulator row (pixel): bias / activation / residual / sign mask / outputs.
// v[i] = raw fp32 accumulator of channel cbase+i; all 32 lanes of the warp, and all four warps of the set, call this
// together.  colsum: lane l's running total of channel cbase + l of the masked output (see the flush in the kernel).
template <int kF>
__device__ __forceinline__ void epi_chunk(const ConvUmmaParams& p, EpiTile& t, uint32_t st0,
                                          const uint32_t (&v)[32], int cbase, bool valid, int lpix, int ppix,
                                          const float* cbias, int lane, float& colsum, float& sqsum) {
  constexpr bool kFwd = (kF & EF_FWD) != 0, kGrad = (kF & EF_GRAD) != 0, kF32 = (kF & EF_F32) != 0;
  constexpr bool kStats = (kF & EF_STATS) != 0, kAdd = (kF & EF_ADD) != 0;
  const int nvalid = min(32, p.cout - cbase);
  if (nvalid <= 0) return;  // uniform over the warp set
  const bool full32 = (nvalid == 32);
  // The buffer about to be written may still be read by an earlier TMA store: whoever writes staging rows next first
  // lets the leader wait for those reads and tells the other warps (every condition below is uniform over the set).
  uint32_t st = st0, set_buf = t.set_base;
  auto acquire = [&]() {
    const uint32_t boff = (p.epi_bufs == 2 && (t.phase & 1)) ? kEpiBytes : 0u;
    st = st0 + boff;
    set_buf = t.set_base + boff;
    if (t.pending) {
      if (t.leader) {
        if (p.epi_bufs == 2) ptx::bulk_wait_group_read1();
        else ptx::bulk_wait_group_read0();
      }
      ptx::named_bar_sync(t.bar_id, 128);
      t.pending = false;
    }
  };
  float f[32];
  uint32_t mbits = 0;
  // sign mask of the consuming activation (second output): loaded first, used last
  uint32_t mi = 0xFFFFFFFFu;
  if (kGrad && p.out2_hi && p.mask_in && valid)
    mi = __ldg(p.mask_in + static_cast<long long>(ppix) * p.mask_in_words + (cbase >> 5));
  // bias of the chunk's 32 channels: ONE coalesced load per warp, broadcast by shuffles.  (32 predicated scalar loads,
  // each followed by its dependent add, cost ~10k clk per chunk on the long scoreboard -- half of the whole tile time
  // of every short-K layer, profiles/r01_epilogue_bias_ncu.txt.)
  if (kFwd && p.bias) {
    float bl = 0.f;
    if (lane < nvalid) bl = __ldg(p.bias + cbase + lane);
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + __shfl_sync(0xffffffffu, bl, i);
  } else {   // data gradients carry no bias: no shuffles
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  }
  if (kFwd && cbias) {  // per-pixel (border class) bias row: vector loads, issued back to back
    if (full32) {
      const float4* cb4 = reinterpret_cast<const float4*>(cbias + cbase);
      float4 tt[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) tt[q] = __ldg(cb4 + q);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        f[4 * q] += tt[q].x;
        f[4 * q + 1] += tt[q].y;
        f[4 * q + 2] += tt[q].z;
        f[4 * q + 3] += tt[q].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) f[i] += __ldg(cbias + cbase + i);
    }
  }
  if (kFwd && (p.mask_out || p.act != DPIG_ACT_NONE)) {
    if (p.act == DPIG_ACT_RELU) {        // (the activation switch is uniform: one loop per kind, no select per element)
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mbits |= (f[i] > 0.f ? 1u : 0u) << i;
        f[i] = fmaxf(f[i], 0.f);
      }
    } else if (p.act == DPIG_ACT_LRELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mbits |= (f[i] > 0.f ? 1u : 0u) << i;
        f[i] = f[i] > 0.f ? f[i] : p.alpha * f[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) mbits |= (f[i] > 0.f ? 1u : 0u) << i;
    }
  }
  // ---- residual / addend
  if (kAdd && p.add_hi) {
    if (full32 && (p.add_ps % 8 == 0)) {
      acquire();
      epi_gather_rows(st, p.add_hi, p.add_lo, p.add_ps, cbase, ppix, valid, lane);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 a = ptx::lds128(st + swz64(lane, q));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          f[q * 8 + 2 * jj] += bf16_bits_to_float(aw[jj] & 0xFFFF);
          f[q * 8 + 2 * jj + 1] += bf16_bits_to_float(aw[jj] >> 16);
        }
        if (p.add_lo) {
          const uint4 b2 = ptx::lds128(st + kEpiLoOff + swz64(lane, q));
          const uint32_t bw[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            f[q * 8 + 2 * jj] += bf16_bits_to_float(bw[jj] & 0xFFFF);
            f[q * 8 + 2 * jj + 1] += bf16_bits_to_float(bw[jj] >> 16);
          }
        }
      }
      __syncwarp();
    } else if (valid) {
      const long long off = static_cast<long long>(ppix) * p.add_ps + cbase;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) {
          float a = __bfloat162float(p.add_hi[off + i]);
          if (p.add_lo) a += __bfloat162float(p.add_lo[off + i]);
          f[i] += a;
        }
    }
  }
  if (kFwd && p.mask_out && valid) p.mask_out[static_cast<long long>(lpix) * p.mask_out_words + (cbase >> 5)] = mbits;
  // ---- split-bf16 output
  if (p.out_hi) {
    if (full32 && p.out_tma) {
      acquire();
      epi_stage_split(st, f, lane, p.out_lo != nullptr);
      ptx::fence_proxy_async();   // generic-proxy writes of the rows -> visible to the async proxy
      ptx::named_bar_sync(t.bar_id, 128);
      if (t.leader) {
        for (int r = 0; r < p.rep * p.rep; ++r) {   // rep = 2: the four replicas of the fused nearest-neighbour upsample
          ptx::tma_store_4d(&p.out_map[t.cls + r][0], set_buf, cbase, t.w0, t.h0, t.n0);
          if (p.out_lo) ptx::tma_store_4d(&p.out_map[t.cls + r][1], set_buf + kEpiLoOff, cbase, t.w0, t.h0, t.n0);
        }
        ptx::bulk_commit_group();
      }
      t.phase++;
      t.pending = true;
    } else if (full32 && (p.out_ps % 8 == 0)) {
      acquire();
      epi_stage_split(st, f, lane, p.out_lo != nullptr);
      __syncwarp();
      for (int dy = 0; dy < p.rep; ++dy)
        for (int dx = 0; dx < p.rep; ++dx)
          epi_scatter_rows(st, p.out_hi, p.out_lo, p.out_ps, cbase, ppix + dy * p.out_W + dx, valid, lane);
      __syncwarp();
    } else if (valid) {
      for (int dy = 0; dy < p.rep; ++dy)
        for (int dx = 0; dx < p.rep; ++dx)
          store_split32(p.out_hi, p.out_lo,
                        static_cast<long long>(ppix + dy * p.out_W + dx) * p.out_ps + cbase, f, nvalid, false);
    }
  }
  // ---- fp32 output
  if (kF32 && p.out_f32) {
    if (full32 && (p.out_f32_ps % 4 == 0)) {
      acquire();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        ptx::sts128(st + swz128(lane, q), make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                                                     __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3])));
      __syncwarp();
      for (int dy = 0; dy < p.rep; ++dy)
        for (int dx = 0; dx < p.rep; ++dx) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int q = 4 * i + (lane >> 3), pc = lane & 7;
            const int pq = __shfl_sync(0xffffffffu, ppix, q);
            const int vq = __shfl_sync(0xffffffffu, static_cast<int>(valid), q);
            if (vq)
              *reinterpret_cast<uint4*>(p.out_f32 + static_cast<long long>(pq + dy * p.out_W + dx) * p.out_f32_ps + cbase + pc * 4) =
                  ptx::lds128(st + swz128(q, pc));
          }
        }
      __syncwarp();
    } else if (valid) {
      for (int dy = 0; dy < p.rep; ++dy)
        for (int dx = 0; dx < p.rep; ++dx) {
          float* o = p.out_f32 + static_cast<long long>(ppix + dy * p.out_W + dx) * p.out_f32_ps + cbase;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) o[i] = f[i];
        }
    }
  }
  // ---- second output multiplied by the incoming sign mask (backward of ReLU / LeakyReLU)
  if (kGrad && p.out2_hi) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] *= ((mi >> i) & 1u) ? 1.f : p.mask_neg;
    if (full32 && p.out2_tma) {
      acquire();
      epi_stage_split(st, f, lane, p.out2_lo != nullptr);
      ptx::fence_proxy_async();
      ptx::named_bar_sync(t.bar_id, 128);
      if (t.leader) {
        ptx::tma_store_4d(&p.out2_map[t.cls][0], set_buf, cbase, t.w0, t.h0, t.n0);
        if (p.out2_lo) ptx::tma_store_4d(&p.out2_map[t.cls][1], set_buf + kEpiLoOff, cbase, t.w0, t.h0, t.n0);
        ptx::bulk_commit_group();
      }
      t.phase++;
      t.pending = true;
    } else if (full32 && (p.out2_ps % 8 == 0)) {
      acquire();
      epi_stage_split(st, f, lane, p.out2_lo != nullptr);
      __syncwarp();
      epi_scatter_rows(st, p.out2_hi, p.out2_lo, p.out2_ps, cbase, ppix, valid, lane);
      __syncwarp();
    } else if (valid) {
      store_split32(p.out2_hi, p.out2_lo, static_cast<long long>(ppix) * p.out2_ps + cbase, f, nvalid, false);
    }
    if (p.colsum) {
      // bias gradient of the consuming layer = column sums of this masked gradient.  Transpose-reduce over the
      // warp's 32 pixel rows in 31 shuffles: after step `off` every lane keeps the half of its values whose channel
      // index has bit `off` equal to its lane-id bit, so lane l ends with the total of channel cbase + l.
      // (Runs while the TMA store of the rows above drains.)
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = 0.f;
      }
      colsum += warp_colsum32(f, lane);
    }
  }
  // ---- raw normalisation sums of pre = acc + bias (act = none, no addend: f still holds pre).  Runs last, while the
  // stores above drain.  Batch mode: per-channel totals of the warp's 32 rows (lane l = channel cbase + l);
  // layer mode: this row's totals over the chunk's channels (the caller reduces rows of one image).
  if (kStats && p.stat_sums) {
    if (!valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = 0.f;
    }
    if (p.stat_mode == DPIG_NORM_BATCH) {
      float q[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) q[i] = f[i] * f[i];
      sqsum += warp_colsum32(q, lane);
      colsum += warp_colsum32(f, lane);
    } else {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) {
          a += f[i];
          b = fmaf(f[i], f[i], b);
        }
      colsum += a;
      sqsum += b;
    }
  }
}

// Persistent: grid = min(#tiles, #SMs); CTA b processes tiles b, b+grid, ...  The smem pipeline runs
// continuously across tiles and the fp32 accumulator is double-buffered in TMEM (2 x block_n columns), so
// the epilogue of tile j overlaps the MMAs of tile j+1.
//
// kPair: the grid is a list of 2-CTA clusters (one TPC each).  The pair computes two pixel tiles of the same output
// channel block with ONE M=256 `cta_group::2` MMA stream issued by the leader (cluster rank 0): each CTA stages its
// own 128 pixel rows of A and HALF of the weight rows (p.b_bytes is the per-CTA share), so the L2 -> SM operand
// traffic per FLOP drops by a third (N=256) -- the single-CTA kernel is bound by that feed, not by the tensor pipe.
//   full[s]        leader only; count 1 (leader's expect_tx covers both CTAs' bytes, peer TMAs signal it remotely)
//   empty[s]       one per CTA; arrived by the leader's multicast tcgen05.commit
//   tmem_full[a]   one per CTA; multicast commit after the last MMA of a tile
//   tmem_empty[a]  leader only; count 2*kEpiWarps (peer epilogue warps arrive remotely)
// kWide: the wide-B form of the three passes (ConvUmmaParams::wide_b), single-CTA kernel only; a template parameter so
// that the second TMEM read of its epilogue costs the other variants no registers.
// kF: epilogue feature groups compiled in (EF_*); see launch_conv for the choice.
template <bool kPair, bool kWide, int kF>
// 10 warps = 3+3+2+2 per SM sub-partition (16K registers each) caps the kernel at 168 registers per thread.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ ConvUmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const uint32_t b_off = p.planes * kABytes;
  const uint32_t stage_bytes = p.planes * (kABytes + p.b_bytes);
  uint8_t* epi_smem = smem + p.stages * stage_bytes;        // epi_bufs x kEpiBytes, 1024-byte aligned (TMA store source)
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + p.epi_bufs * kEpiBytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int n_tiles = (p.cout + p.block_n - 1) / p.block_n;
  // work units: single CTA -> (pixel tile, channel block); pair -> (two adjacent pixel tiles, channel block)
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const int unit0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int unit_pix_tiles = kPair ? (pix_tiles + 1) / 2 : pix_tiles;
  const int total_tiles = unit_pix_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], kPair ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_alloc_pair<64>(tmem_slot); break;
        case 128: ptx::tmem_alloc_pair<128>(tmem_slot); break;
        case 256: ptx::tmem_alloc_pair<256>(tmem_slot); break;
        default: ptx::tmem_alloc_pair<512>(tmem_slot); break;
      }
    } else {
      switch (p.tmem_cols) {  // power of two >= 2 * block_n
        case 64: ptx::tmem_alloc<64>(tmem_slot); break;
        case 128: ptx::tmem_alloc<128>(tmem_slot); break;
        case 256: ptx::tmem_alloc<256>(tmem_slot); break;
        default: ptx::tmem_alloc<512>(tmem_slot); break;
      }
    }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();  // peer barriers must be initialised before any remote arrive / multicast commit
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < 4; ++s)
        for (int pl = 0; pl < p.planes; ++pl) ptx::prefetch_tmap(&p.a_map[s][pl]);
      for (int pl = 0; pl < p.planes; ++pl) ptx::prefetch_tmap(&p.b_map[pl]);
      int s = 0;
      uint32_t ph = 0;
      const int b_row0 = kPair ? static_cast<int>(rank) * (p.block_n >> 1) : 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int upt = tile % unit_pix_tiles, nt = tile / unit_pix_tiles;
        // pt == pix_tiles for the odd tail of a pair: an all-out-of-bounds (zero-filled) tile
        const int pt = kPair ? 2 * upt + static_cast<int>(rank) : upt;
        const int tw = pt % p.tiles_w;
        const int th = (pt / p.tiles_w) % p.tiles_h;
        const int tn = pt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;
        for (int t = 0; t < p.num_taps; ++t) {   // all classes of the unit, in class order (taps are sorted by class)
          const ConvTap tap = p.taps[t];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * stage_bytes;
            if (kPair) {
              const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full[s]), 0);
              if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * p.planes * (p.a_tx_bytes + p.b_bytes));
              if (p.a_merged)
                ptx::tma_load_5d_pair(st, &p.a_map[tap.src][0], lead_full, kc * 64, w0 + tap.dw, h0 + tap.dh, n0, 0);
              else
                for (int pl = 0; pl < p.planes; ++pl)
                  ptx::tma_load_4d_pair(st + pl * kABytes, &p.a_map[tap.src][pl], lead_full, kc * 64, w0 + tap.dw,
                                        h0 + tap.dh, n0);
              if (p.b_merged)
                ptx::tma_load_4d_pair(st + b_off, &p.b_map[0], lead_full, kc * 64, nt * p.block_n + b_row0, tap.wtap, 0);
              else
                for (int pl = 0; pl < p.planes; ++pl)
                  ptx::tma_load_3d_pair(st + b_off + pl * p.b_bytes, &p.b_map[pl], lead_full, kc * 64,
                                        nt * p.block_n + b_row0, tap.wtap);
            } else {
            ptx::mbar_expect_tx(&full[s], p.planes * (p.a_tx_bytes + p.b_bytes));
            if (p.a_merged)
              ptx::tma_load_5d(st, &p.a_map[tap.src][0], &full[s], kc * 64, w0 + tap.dw, h0 + tap.dh, n0, 0);
            else
              for (int pl = 0; pl < p.planes; ++pl)
                ptx::tma_load_4d(st + pl * kABytes, &p.a_map[tap.src][pl], &full[s], kc * 64, w0 + tap.dw,
                                 h0 + tap.dh, n0);
            if (p.b_merged)
              ptx::tma_load_4d(st + b_off, &p.b_map[0], &full[s], kc * 64, nt * p.block_n, tap.wtap, 0);
            else
              for (int pl = 0; pl < p.planes; ++pl)
                ptx::tma_load_3d(st + b_off + pl * p.b_bytes, &p.b_map[pl], &full[s], kc * 64, nt * p.block_n,
                                 tap.wtap);
            }
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kPair ? 256 : 128, p.block_n, 0, 0);
      const uint32_t idesc_wide = ptx::make_idesc_bf16(128, 2 * p.block_n, 0, 0);
      const uint32_t acc_cols = kWide ? 2 * p.block_n : p.block_n;
      int s = 0;
      uint32_t ph = 0;
      int j = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step)
      for (int cls = 0; cls < p.nclass; ++cls, ++j) {
        const int acc = j & 1;
        ptx::mbar_wait(&tmem_empty[acc], ((j >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * acc_cols;
        const int steps_per_tile = (p.cls_tap0[cls + 1] - p.cls_tap0[cls]) * p.kchunks;
        for (int step = 0; step < steps_per_tile; ++step) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_hi = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t a_lo = a_hi + kABytes;
          const uint32_t b_hi = a_hi + b_off;
          const uint32_t b_lo = b_hi + p.b_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da_hi = ptx::make_desc_sw128(a_hi + k * 32, 16, 1024);
            const uint64_t db_hi = ptx::make_desc_sw128(b_hi + k * 32, 16, 1024);
            if (kPair) {
              ptx::umma_bf16_pair(tmem_acc, da_hi, db_hi, idesc, (step | k) ? 1u : 0u);
              if (p.planes == 2) {
                const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 32, 16, 1024);
                const uint64_t db_lo = ptx::make_desc_sw128(b_lo + k * 32, 16, 1024);
                ptx::umma_bf16_pair(tmem_acc, da_hi, db_lo, idesc, 1u);
                ptx::umma_bf16_pair(tmem_acc, da_lo, db_hi, idesc, 1u);
              }
            } else if (kWide) {
              // rows [0, block_n) of the B slot are the hi weights, rows [block_n, 2*block_n) the lo weights
              const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 32, 16, 1024);
              ptx::umma_bf16(tmem_acc, da_hi, db_hi, idesc_wide, (step | k) ? 1u : 0u);
              ptx::umma_bf16(tmem_acc, da_lo, db_hi, idesc, 1u);
            } else {
              ptx::umma_bf16(tmem_acc, da_hi, db_hi, idesc, (step | k) ? 1u : 0u);
              if (p.planes == 2) {
                const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 32, 16, 1024);
                const uint64_t db_lo = ptx::make_desc_sw128(b_lo + k * 32, 16, 1024);
                ptx::umma_bf16(tmem_acc, da_hi, db_lo, idesc, 1u);
                ptx::umma_bf16(tmem_acc, da_lo, db_hi, idesc, 1u);
              }
            }
          }
          if (kPair) ptx::umma_commit_pair(&empty[s], 3);
          else ptx::umma_commit(&empty[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (kPair) ptx::umma_commit_pair(&tmem_full[acc], 3);
        else ptx::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue: 8 warps in two sets, thread = one accumulator row (= one output pixel).
    // Output rows are staged in the set's shared-memory buffer(s) and leave as TMA tensor stores (see kEpiSetBytes).
    const int lg = warp & 3;            // TMEM lane group this warp may read (hardware: warp id % 4)
    const int ew = warp - 2;            // epilogue warp index 0..kEpiWarps-1
    const int cpar = ew >> 2;           // which alternate 32-column chunks this warp takes
    const int r = lg * 32 + lane;
    const int wl = r % p.BW;
    const int tq = r / p.BW;
    const int hl = tq % p.BH;
    const int nl = tq / p.BH;
    EpiTile et;
    et.set_base = ptx::smem_u32(epi_smem) + cpar * kEpiSetBytes;
    et.bar_id = 1 + cpar;
    et.leader = (lg == 0 && lane == 0);
    et.phase = 0;
    et.pending = false;
    const uint32_t st = et.set_base + lg * 2048;   // this warp's 32 rows of the hi plane
    const uint32_t lead_tmem_empty[2] = {kPair ? ptx::mapa_u32(ptx::smem_u32(&tmem_empty[0]), 0) : 0u,
                                         kPair ? ptx::mapa_u32(ptx::smem_u32(&tmem_empty[1]), 0) : 0u};
    // Bias-gradient column sums stay in registers across all tiles of one channel block (a warp owns at most four
    // 32-channel chunks of it; lane l holds channel chunk*32 + l) and reach global memory with one atomic per channel
    // when the block changes / at the end.  One atomic per chunk and 32 pixel rows (the first version) put
    // pixels/32 same-address fp32 atomics on every channel: ~1.3 clk each at the L2, 0.6 ms on a 1 M-pixel layer --
    // it made every short-K data gradient (stride-2 parity classes, 1x1) atomic-bound instead of tensor-bound.
    // The per-channel normalisation sums of a batch-norm conv (stat_sums, never together with colsum) use the same
    // registers plus four for the squares; they reach memory as fp64 atomics, one per channel, warp and channel block.
    float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
    float qs0 = 0.f, qs1 = 0.f, qs2 = 0.f, qs3 = 0.f;
    int cs_nt = -1;
    constexpr bool kGrad = (kF & EF_GRAD) != 0, kStats = (kF & EF_STATS) != 0, kAdd = (kF & EF_ADD) != 0;
    const bool stat_batch = kStats && p.stat_sums != nullptr && p.stat_mode == DPIG_NORM_BATCH;
    const bool stat_layer = kStats && p.stat_sums != nullptr && p.stat_mode != DPIG_NORM_BATCH;
    auto flush_colsum = [&]() {
      if (!kGrad && !kStats) return;
      if (((!kGrad || p.colsum == nullptr) && !stat_batch) || cs_nt < 0) return;
      const float cs[4] = {cs0, cs1, cs2, cs3};
      const float qs[4] = {qs0, qs1, qs2, qs3};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = cs_nt * p.block_n + ((cpar + k * (kEpiWarps / 4)) << 5) + lane;
        if (cpar + k * (kEpiWarps / 4) < (p.block_n >> 5) && c < p.cout) {
          if (stat_batch) {
            atomicAdd(p.stat_sums + c, static_cast<double>(cs[k]));
            atomicAdd(p.stat_sums + p.stat_groups + c, static_cast<double>(qs[k]));
          } else {
            atomicAdd(p.colsum + c, cs[k]);
          }
        }
      }
      cs0 = cs1 = cs2 = cs3 = 0.f;
      qs0 = qs1 = qs2 = qs3 = 0.f;
    };
    int j = 0;
    for (int tile = unit0; tile < total_tiles; tile += unit_step)
    for (int cls = 0; cls < p.nclass; ++cls, ++j) {
      const int acc = j & 1;
      const int upt = tile % unit_pix_tiles, nt = tile / unit_pix_tiles;
      if (nt != cs_nt) {
        flush_colsum();
        cs_nt = nt;
      }
      const int pt = kPair ? 2 * upt + static_cast<int>(rank) : upt;
      const int tw = pt % p.tiles_w;
      const int th = (pt / p.tiles_w) % p.tiles_h;
      const int tn = pt / (p.tiles_w * p.tiles_h);
      const int w = tw * p.BW + wl, h = th * p.BH + hl, n = tn * p.BN + nl;
      et.w0 = tw * p.BW;
      et.h0 = th * p.BH;
      et.n0 = tn * p.BN;
      et.cls = cls;
      const bool valid = (nl < p.BN) && (w < p.Wo) && (h < p.Ho) && (n < p.No);
      const int lpix = (n * p.Ho + h) * p.Wo + w;  // logical pixel (< 2^31 for every supported shape)
      const int py = h * p.sh + p.cls_oh[cls], px = w * p.sw + p.cls_ow[cls];
      const int ppix = (n * p.out_H + py) * p.out_W + px;
      const float* cbias = nullptr;
      if ((kF & EF_FWD) && p.class_bias && valid) {
        const int ch = h == 0 ? 0 : (h == p.Ho - 1 ? 2 : 1), cw = w == 0 ? 0 : (w == p.Wo - 1 ? 2 : 1);
        cbias = p.class_bias + (static_cast<long long>(n) * 9 + ch * 3 + cw) * p.cout;
      }

      // The residual rows of this tile are gathered chunk by chunk further down, each gather an exposed global-load
      // latency (12 % of all stall samples of a two-output data gradient, profiles/r01_dgrad_epilogue_micro.txt):
      // pull them into L2 now, while the tile's MMAs are still running.
      if (kAdd && p.add_prefetch && p.add_hi && valid && cpar == 0) {
        const long long e0 = static_cast<long long>(ppix) * p.add_ps + nt * p.block_n;
        const int ne = min(p.block_n, p.cout - nt * p.block_n);
        for (int e = 0; e < ne; e += 64) {
          ptx::prefetch_l2(p.add_hi + e0 + e);
          if (p.add_lo) ptx::prefetch_l2(p.add_lo + e0 + e);
        }
      }
      ptx::mbar_wait(&tmem_full[acc], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * (kWide ? 2 * p.block_n : p.block_n) + (static_cast<uint32_t>(lg * 32) << 16);

      const int nchunks = p.block_n >> 5;
      bool released = false;
      for (int ci = cpar; ci < nchunks; ci += kEpiWarps / 4) {
        const int c0 = ci << 5;
        uint32_t v[32];
        ptx::tmem_ld32(tmem_acc + c0, v);
        if (kWide) {   // second column range: the A_hi * B_lo products
          uint32_t v2[32];
          ptx::tmem_ld32(tmem_acc + p.block_n + c0, v2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        } else {
          ptx::tmem_ld_wait();
        }
        if (ci + kEpiWarps / 4 >= nchunks) {
          // this warp's last TMEM read of the tile: hand the accumulator stage back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair) ptx::mbar_arrive_cluster(lead_tmem_empty[acc]);
            else ptx::mbar_arrive(&tmem_empty[acc]);
          }
          released = true;
        }
        float csum = 0.f, qsum = 0.f;
        epi_chunk<kF>(p, et, st, v, nt * p.block_n + c0, valid, lpix, ppix, cbias, lane, csum, qsum);
        if (!kGrad && !kStats) {
          // no column sums in this instantiation
        } else if (stat_layer) {   // this row's sums over the chunk's channels: collected per tile (cs0 / qs0), flushed below
          cs0 += csum;
          qs0 += qsum;
        } else {
          const int k = (ci - cpar) / (kEpiWarps / 4);   // this warp's k-th chunk of the block
          cs0 += k == 0 ? csum : 0.f;
          cs1 += k == 1 ? csum : 0.f;
          cs2 += k == 2 ? csum : 0.f;
          cs3 += k == 3 ? csum : 0.f;
          qs0 += k == 0 ? qsum : 0.f;
          qs1 += k == 1 ? qsum : 0.f;
          qs2 += k == 2 ? qsum : 0.f;
          qs3 += k == 3 ? qsum : 0.f;
        }
      }
      if (stat_layer) {
        // per-sample sums: rows of one image reduce inside the warp when the whole warp sits in one image (the usual
        // case: >= 32 pixels per image and tile), else every row adds its own share
        const int n_first = __shfl_sync(0xffffffffu, n, 0);
        const bool uniform = __all_sync(0xffffffffu, !valid || n == n_first) && __shfl_sync(0xffffffffu, static_cast<int>(valid), 0);
        if (uniform) {
          float a = cs0, b = qs0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (lane == 0) {
            atomicAdd(p.stat_sums + n_first, static_cast<double>(a));
            atomicAdd(p.stat_sums + p.stat_groups + n_first, static_cast<double>(b));
          }
        } else if (valid) {
          atomicAdd(p.stat_sums + n, static_cast<double>(cs0));
          atomicAdd(p.stat_sums + p.stat_groups + n, static_cast<double>(qs0));
        }
        cs0 = qs0 = 0.f;
      }
      if (!released) {  // fewer chunks than epilogue warps per lane group (block_n == 32)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) ptx::mbar_arrive_cluster(lead_tmem_empty[acc]);
          else ptx::mbar_arrive(&tmem_empty[acc]);
        }
      }
    }
    flush_colsum();
    if (et.leader) ptx::bulk_wait_group0();   // the last TMA stores read this CTA's shared memory
  }

  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();  // the leader's MMAs read the peer's smem; nobody leaves before both are done
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_dealloc_pair<64>(tmem_base); break;
        case 128: ptx::tmem_dealloc_pair<128>(tmem_base); break;
        case 256: ptx::tmem_dealloc_pair<256>(tmem_base); break;
        default: ptx::tmem_dealloc_pair<512>(tmem_base); break;
      }
    } else {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_dealloc<64>(tmem_base); break;
        case 128: ptx::tmem_dealloc<128>(tmem_base); break;
        case 256: ptx::tmem_dealloc<256>(tmem_base); break;
        default: ptx::tmem_dealloc<512>(tmem_base); break;
      }
    }
  }
}

// ------------------------------------------------------------------------------------- wgrad
// A CTA owns a GROUP of filter taps (p.group taps, one fp32 accumulator each in TMEM): the dy tile of a pixel step
// is loaded once and multiplied with the tap-shifted x tiles of every tap of the group, which cuts the L2 -> SM bytes
// per FLOP (2 taps x 256 columns: 43 B per MMA clock instead of 64).  Measured on B200 (profiles/r01_wgrad_ab.txt) the
// kernel was NOT bound by those bytes but by the number of TMA instructions per step (12 x ~200 clk on the one issuing
// thread): fetching all 64-channel boxes of an operand plane with ONE 5-D TMA (4 per step) took the 256-wide layers
// from 440 to 477 TFLOP/s algorithmic (= the tensor peak at 3 passes) and the 128-wide ones from 280 to 349; with
// that, groups of 2 / 3 taps measured 0-8 % slower than one tap per CTA (32-pixel steps are needed to fit two
// stages), so the default group is 1 and the grouping stays available as a tuning knob (DPIG_WGRAD_GROUP).
struct WgradParams {
  CUtensorMap x_map[4][2];
  CUtensorMap dy_map[2];
  ConvTap taps[kMaxTaps];
  int planes;
  int PW, PH, PN, tiles_w, tiles_h;
  int total_tiles, tiles_per_cta;
  int block_n, n_tiles, m_tiles, cin, cout, stages;
  int cin_pitch;  // rows per tap in the HWIO gradient (>= cin when dw is a row-slice of a wider filter)
  int cout_pitch; // columns per row of the HWIO gradient (>= cout when this launch covers a column segment of it)
  int num_taps, group;      // taps per CTA (accumulators), blockIdx.y = tap group
  int px;                   // pixels (GEMM K) per pipeline step: 32 or 64
  int x_grouped, dy_grouped;  // operand fetched by ONE 5-D TMA per plane (64-channel groups as the 5th box dimension)
  int wide_b;               // single-CTA, one tap, block_n <= 128: dy hi|lo as one N = 2*block_n operand (see ConvUmmaParams)
  int vec_red;              // rows of dw are 16-byte aligned: partial sums leave as red.global.add.v4.f32
  uint32_t a_bytes;         // one x tile of one plane: 2 boxes of px rows x 128 B
  uint32_t b_bytes, tmem_cols;
  float* dw;
};

// kPair: 2-CTA clusters running ONE M=256 `cta_group::2` MMA stream (issued by cluster rank 0).  The two CTAs take two
// work units (filter tap x 128-input-channel tile) that share the dy tile: each stages its own tap-shifted x tile (its
// 128 rows of D) and HALF of the dy channels, so the L2 -> SM operand bytes per MMA clock drop from 62 to 42 B (block_n =
// 256) -- the single-CTA kernel sits at the ~49 B/clk/SM feed limit -- and three 64 KB stages fit where two 96 KB ones
// did.  Units are paired over (tap, channel tile), so 128-input-channel layers pair two taps.  Barriers as in the conv
// pair kernel: full[s] on the leader (its expect_tx covers both CTAs' bytes, the peer's TMAs signal it remotely),
// empty[s] / tmem_full in each CTA, arrived by the leader's multicast commits.
template <bool kPair>
__global__ void __launch_bounds__(192, 1)
wgrad_umma_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  // stage layout: [tap 0: x hi | x lo] ... [tap group-1: x hi | x lo] [dy hi | dy lo]   (pair: this CTA's half of dy)
  const uint32_t box_bytes = p.px * 128;
  const uint32_t b_bytes = kPair ? p.b_bytes / 2 : p.b_bytes;
  const uint32_t b_off = p.group * p.planes * p.a_bytes;
  const uint32_t stage_bytes = p.planes * (p.group * p.a_bytes + b_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  int nt, mt, tap0, ntap;
  bool active = true;
  if (kPair) {
    const int pi = static_cast<int>(blockIdx.x >> 1);
    nt = pi % p.n_tiles;
    const int units = p.num_taps * p.m_tiles;
    int unit = 2 * (pi / p.n_tiles) + static_cast<int>(rank);
    if (unit >= units) {   // odd unit count: the last cluster's second CTA mirrors its partner and writes nothing
      unit = units - 1;
      active = false;
    }
    tap0 = unit / p.m_tiles;
    mt = unit % p.m_tiles;
    ntap = 1;
  } else {
    nt = blockIdx.x % p.n_tiles;  // output-channel tile
    mt = blockIdx.x / p.n_tiles;  // input-channel tile (128 wide)
    tap0 = blockIdx.y * p.group;
    ntap = min(p.group, p.num_taps - tap0);
  }
  const int tile_begin = blockIdx.z * p.tiles_per_cta;
  const int tile_end = min(tile_begin + p.tiles_per_cta, p.total_tiles);
  const int nsteps = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_alloc_pair<64>(tmem_slot); break;
        case 128: ptx::tmem_alloc_pair<128>(tmem_slot); break;
        case 256: ptx::tmem_alloc_pair<256>(tmem_slot); break;
        default: ptx::tmem_alloc_pair<512>(tmem_slot); break;
      }
    } else {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_alloc<64>(tmem_slot); break;
        case 128: ptx::tmem_alloc<128>(tmem_slot); break;
        case 256: ptx::tmem_alloc<256>(tmem_slot); break;
        default: ptx::tmem_alloc<512>(tmem_slot); break;
      }
    }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();  // peer barriers must be initialised before any remote arrive / multicast commit
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nboxes_b = p.block_n / 64;
  const int nboxes_cta = kPair ? nboxes_b / 2 : nboxes_b;   // dy boxes staged by this CTA
  const int box0_b = nt * nboxes_b + (kPair ? static_cast<int>(rank) * nboxes_cta : 0);

  if (nsteps > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int t = tile_begin; t < tile_end; ++t) {
          const int tw = t % p.tiles_w;
          const int th = (t / p.tiles_w) % p.tiles_h;
          const int tn = t / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.PW, h0 = th * p.PH, n0 = tn * p.PN;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * stage_bytes;
          if (kPair) {
            const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full[s]), 0);
            if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * p.planes * (p.a_bytes + b_bytes));
            const ConvTap tap = p.taps[tap0];
            for (int pl = 0; pl < p.planes; ++pl) {
              ptx::tma_load_5d_pair(st + b_off + pl * b_bytes, &p.dy_map[pl], lead_full, 0, w0, h0, n0, box0_b);
              ptx::tma_load_5d_pair(st + pl * p.a_bytes, &p.x_map[tap.src][pl], lead_full, 0, w0 + tap.dw, h0 + tap.dh,
                                    n0, mt * 2);
            }
          } else {
          ptx::mbar_expect_tx(&full[s], p.planes * (ntap * p.a_bytes + p.b_bytes));
          // (the kernel is sensitive to the NUMBER of TMA instructions per step, ~200 clk each on one issuing thread:
          //  a 64-channel box per instruction made 12 per step and capped it at ~55 % of the MMA rate)
          for (int pl = 0; pl < p.planes; ++pl) {
            if (p.dy_grouped) {
              ptx::tma_load_5d(st + b_off + pl * p.b_bytes, &p.dy_map[pl], &full[s], 0, w0, h0, n0,
                               nt * nboxes_b);
            } else {
              for (int j = 0; j < nboxes_b; ++j)
                ptx::tma_load_4d(st + b_off + pl * p.b_bytes + j * box_bytes, &p.dy_map[pl], &full[s],
                                 nt * p.block_n + j * 64, w0, h0, n0);
            }
          }
          for (int i = 0; i < ntap; ++i) {
            const ConvTap tap = p.taps[tap0 + i];
            for (int pl = 0; pl < p.planes; ++pl) {
              uint8_t* dst = st + (i * p.planes + pl) * p.a_bytes;
              if (p.x_grouped) {
                ptx::tma_load_5d(dst, &p.x_map[tap.src][pl], &full[s], 0, w0 + tap.dw, h0 + tap.dh, n0, mt * 2);
              } else {
                for (int j = 0; j < 2; ++j)
                  ptx::tma_load_4d(dst + j * box_bytes, &p.x_map[tap.src][pl], &full[s], mt * 128 + j * 64,
                                   w0 + tap.dw, h0 + tap.dh, n0);
              }
            }
          }
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        const uint32_t idesc = ptx::make_idesc_bf16(kPair ? 256 : 128, p.block_n, 1, 1);
        const uint32_t idesc_wide = ptx::make_idesc_bf16(128, 2 * p.block_n, 1, 1);
        const int ksteps = p.px / 16;  // 16 pixels (K) per MMA = 2 KB of rows
        int s = 0;
        uint32_t ph = 0;
        for (int step = 0; step < nsteps; ++step) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t st = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t b_hi = st + b_off;
          const uint32_t b_lo = b_hi + b_bytes;
          for (int i = 0; i < ntap; ++i) {
            const uint32_t a_hi = st + i * p.planes * p.a_bytes;
            const uint32_t a_lo = a_hi + p.a_bytes;
            const uint32_t acc = tmem_base + i * p.block_n;
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t da_hi = ptx::make_desc_sw128(a_hi + k * 2048, box_bytes, 1024);
              const uint64_t db_hi = ptx::make_desc_sw128(b_hi + k * 2048, box_bytes, 1024);
              if (kPair) {
                ptx::umma_bf16_pair(acc, da_hi, db_hi, idesc, (step | k) ? 1u : 0u);
                if (p.planes == 2) {
                  const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 2048, box_bytes, 1024);
                  const uint64_t db_lo = ptx::make_desc_sw128(b_lo + k * 2048, box_bytes, 1024);
                  ptx::umma_bf16_pair(acc, da_hi, db_lo, idesc, 1u);
                  ptx::umma_bf16_pair(acc, da_lo, db_hi, idesc, 1u);
                }
              } else if (p.wide_b) {
                // the lo boxes of dy follow the hi boxes in the stage: atoms 0.. = hi channels, then the lo channels
                const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 2048, box_bytes, 1024);
                ptx::umma_bf16(acc, da_hi, db_hi, idesc_wide, (step | k) ? 1u : 0u);
                ptx::umma_bf16(acc, da_lo, db_hi, idesc, 1u);
              } else {
                ptx::umma_bf16(acc, da_hi, db_hi, idesc, (step | k) ? 1u : 0u);
                if (p.planes == 2) {
                  const uint64_t da_lo = ptx::make_desc_sw128(a_lo + k * 2048, box_bytes, 1024);
                  const uint64_t db_lo = ptx::make_desc_sw128(b_lo + k * 2048, box_bytes, 1024);
                  ptx::umma_bf16(acc, da_hi, db_lo, idesc, 1u);
                  ptx::umma_bf16(acc, da_lo, db_hi, idesc, 1u);
                }
              }
            }
          }
          if (kPair) ptx::umma_commit_pair(&empty[s], 3);
          else ptx::umma_commit(&empty[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (kPair) ptx::umma_commit_pair(tmem_full, 3);
        else ptx::umma_commit(tmem_full);
      }
    } else {
      const int lg = warp & 3;
      const int ci = mt * 128 + lg * 32 + lane;
      ptx::mbar_wait(tmem_full, 0);
      ptx::tc_fence_after();
      for (int i = 0; i < ntap; ++i) {
        const int wtap = p.taps[tap0 + i].wtap;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + i * p.block_n + c0, v);
          if (p.wide_b) {
            uint32_t v2[32];
            ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + p.block_n + c0, v2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          } else {
            ptx::tmem_ld_wait();
          }
          const int cbase = nt * p.block_n + c0;
          if (active && ci < p.cin) {
            float* o = p.dw + (static_cast<long long>(wtap) * p.cin_pitch + ci) * p.cout_pitch + cbase;
            if (p.vec_red && cbase + 32 <= p.cout) {
              // 8 vector reductions per row instead of 32 scalar atomics: the partial sums of a few-pixel layer (a short
              // K loop per CTA) spend longer in this epilogue than in their MMAs
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                ptx::red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < p.cout) atomicAdd(o + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();  // the leader's MMAs read the peer's smem; nobody leaves before both are done
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_dealloc_pair<64>(tmem_base); break;
        case 128: ptx::tmem_dealloc_pair<128>(tmem_base); break;
        case 256: ptx::tmem_dealloc_pair<256>(tmem_base); break;
        default: ptx::tmem_dealloc_pair<512>(tmem_base); break;
      }
    } else {
      switch (p.tmem_cols) {
        case 64: ptx::tmem_dealloc<64>(tmem_base); break;
        case 128: ptx::tmem_dealloc<128>(tmem_base); break;
        case 256: ptx::tmem_dealloc<256>(tmem_base); break;
        default: ptx::tmem_dealloc<512>(tmem_base); break;
      }
    }
  }
}

// ------------------------------------------------------------------------------------- host
static int encode_map(dpig_ctx* ctx, CUtensorMap* map, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (dims[i] == 0) return set_error(ctx, DPIG_EUNSUPPORTED, "tensor map: empty dimension %d", i);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16) return set_error(ctx, DPIG_EINVAL, "tensor map: stride %d not 16B aligned", i);
  }
  if (reinterpret_cast<uintptr_t>(base) % 16) return set_error(ctx, DPIG_EINVAL, "tensor map: base not 16B aligned");
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base),
                                 gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ctx, DPIG_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DPIG_OK;
}

// Activation view -> 4-D map over (C, W', H', N) where W'/H' are either the full grid or one
// parity class of it (step = 2, offset = parity).
// plane_stride_bytes > 0 appends a 5th "plane" dimension of extent 2 (hi, lo) fetched by the same box.
static int act_map(dpig_ctx* ctx, CUtensorMap* map, const void* plane, const dpig_tensor* t,
                   int step, int py, int px, int box_w, int box_h, int box_n, long long plane_stride_bytes = 0,
                   int cgroups = 0) {
  const uint64_t ps = static_cast<uint64_t>(t->pix_stride);
  const int W2 = (t->w - px + step - 1) / step;
  const int H2 = (t->h - py + step - 1) / step;
  uint64_t dims[4] = {static_cast<uint64_t>(t->c), static_cast<uint64_t>(W2), static_cast<uint64_t>(H2),
                      static_cast<uint64_t>(t->n)};
  uint64_t strides[3] = {ps * 2 * step, ps * 2 * t->w * step, ps * 2 * t->w * t->h};
  uint32_t box[4] = {64, static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h),
                     static_cast<uint32_t>(box_n)};
  const char* base = static_cast<const char*>(plane) + (static_cast<uint64_t>(py) * t->w + px) * ps * 2;
  if (plane_stride_bytes > 0) {
    uint64_t dims5[5] = {dims[0], dims[1], dims[2], dims[3], 2};
    uint64_t strides4[4] = {strides[0], strides[1], strides[2], static_cast<uint64_t>(plane_stride_bytes)};
    uint32_t box5[5] = {box[0], box[1], box[2], box[3], 2};
    return encode_map(ctx, map, base, 5, dims5, strides4, box5);
  }
  if (cgroups > 0) {
    // 5th dimension = 64-channel group (128 B apart): one instruction fetches `cgroups` swizzled 64-channel boxes,
    // laid out one after the other in shared memory (t->c must be a multiple of 64)
    uint64_t dims5[5] = {64, dims[1], dims[2], dims[3], static_cast<uint64_t>(t->c / 64)};
    uint64_t strides4[4] = {strides[0], strides[1], strides[2], 128};
    uint32_t box5[5] = {64, box[1], box[2], box[3], static_cast<uint32_t>(cgroups)};
    return encode_map(ctx, map, base, 5, dims5, strides4, box5);
  }
  return encode_map(ctx, map, base, 4, dims, strides, box);
}

struct Box {
  int bw, bh, bn;
};

// Byte distance lo - hi when both planes can be addressed by one tensor map (same allocation, 16 B granular), else 0.
static long long plane_stride(const void* hi, const void* lo) {
  if (!hi || !lo) return 0;
  const long long d = static_cast<const char*>(lo) - static_cast<const char*>(hi);
  return (d > 0 && d % 16 == 0 && d < (1ll << 40)) ? d : 0;
}
// Pick the pixel box (<= max_rows rows) with the best tile utilisation.
static Box choose_box(int W, int H, int N, int max_rows, bool pow2_exact) {
  Box best{1, 1, 1};
  double best_u = -1;
  for (int bw = 1; bw <= std::min(W, max_rows); ++bw) {
    if (pow2_exact && (bw & (bw - 1))) continue;
    for (int bh = 1; bh <= std::min(H, max_rows / bw); ++bh) {
      if (pow2_exact && (bh & (bh - 1))) continue;
      int bn = std::min(N, max_rows / (bw * bh));
      if (pow2_exact) {
        bn = max_rows / (bw * bh);  // exact product; TMA zero-fills images beyond N
        if (bn > 256) continue;
      }
      if (bn < 1) continue;
      const double tiles = double((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((N + bn - 1) / bn);
      const double u = double(W) * H * N / (tiles * max_rows) + 1e-6 * bw + 1e-9 * bh;
      if (u > best_u) {
        best_u = u;
        best = {bw, bh, bn};
      }
    }
  }
  if (pow2_exact && best_u < 0) {
    // tensors smaller than the box in every dimension: let TMA zero-fill
    int bw = 1;
    while (bw < W && bw < max_rows) bw <<= 1;
    int bh = 1;
    while (bh < H && bw * bh < max_rows) bh <<= 1;
    best = {bw, bh, max_rows / (bw * bh)};
  }
  return best;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (static_cast<int>(c) < n) c <<= 1;
  return c;
}

static int pick_block_n(int cout) {
  // Largest multiple of 32 that is <= 256 and splits cout (rounded up to 32) evenly.
  const int c32 = (cout + 31) / 32 * 32;
  if (c32 <= 256) return c32;
  int best = 32;
  for (int bn = 32; bn <= 256; bn += 32)
    if (c32 % bn == 0) best = bn;
  return best;
}

// Few-pixel layers (8x4, 3x3 maps; fewer work units than SMs at the widest channel block): pick the channel block by
// a cost model instead of the widest one.  Per K step (64 channels of one tap) a CTA needs 32 KB of activations plus
// block_n x 256 B of weights through the ~49 B/clk/SM L2 -> SM feed and 12 MMAs of block_n/2 clk (1.45x that below
// N = 256, where the shared-memory port caps the MMA rate); a unit costs steps x max(feed, MMA), a launch
// ceil(units / SMs) unit times.  Narrow blocks buy parallelism but re-fetch the activation tile once per block: the
// old rule (halve until the grid fills the SMs) went down to N = 64 and ran two half-empty rounds at 3x the bytes.
static int tune_block_n(const dpig_ctx* ctx, int cout, int pix_tiles, int steps) {
  const int widest = pick_block_n(cout);
  if (ctx->tune_small == 0 || pix_tiles * ((cout + widest - 1) / widest) >= ctx->num_sms) return widest;
  if (ctx->tune_small == 2) {   // the old rule, kept for A/B runs
    int bn = widest;
    while (bn >= 128 && (bn / 2) % 32 == 0 && pix_tiles * ((cout + bn - 1) / bn) < ctx->num_sms) bn /= 2;
    return bn;
  }
  const int c32 = (cout + 31) / 32 * 32;
  int best = widest;
  double best_cost = -1.0;
  for (int bn = 32; bn <= 256; bn += 32) {
    if (c32 % bn) continue;
    const int n_tiles = c32 / bn;
    const bool pair = bn >= 64 && pix_tiles >= 2 && ctx->pair_mode >= 1 && (bn > 128 || (steps > 0 && ctx->pair_mode == 2));
    const double feed = (32768.0 + (pair ? bn / 2 : bn) * 256.0) / 49.0;
    const double mma = 12.0 * (bn / 2) * (bn < 256 ? 1.45 : 1.0);
    const double unit = steps * std::max(feed, mma) + 2500.0 + (bn / 32) * 600.0;
    const int units = pair ? ((pix_tiles + 1) / 2) * n_tiles : pix_tiles * n_tiles;
    const int slots = pair ? ctx->num_sms / 2 : ctx->num_sms;
    const double cost = ((units + slots - 1) / slots) * unit;
    if (best_cost < 0 || cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && bn > best)) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// CTA pairs: a 2-CTA cluster runs cta_group::2 M=256 MMAs over two adjacent pixel tiles, each CTA staging half of the
// weight rows (fewer operand bytes per FLOP; the single-CTA kernel is bound by the ~49 B/clk/SM operand feed, not by
// the tensor pipe).  Measured on B200 (profiles/r01_pair_vs_single.txt): channel blocks wider than 128 gain 7-30 %
// and so do single-K-chunk layers; N = 128 blocks with a long K loop lose 4-10 % to the cross-CTA handshake.
// A single-CTA variant with two accumulators per CTA ("dual", two activation tiles next to one weight tile) was also
// built and measured: +11 % on the bare main loop at N = 128 but slower with the real epilogue, and its extra
// registers cost the plain kernel 6-20 %; it was removed again (measurement history).
// mode 0 never, 1 where it measured faster, 2 wherever legal (A/B runs, tests).
static void choose_pair(const dpig_ctx* ctx, ConvUmmaParams& P) {
  const int pix_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  const bool legal = P.block_n >= 64 && P.block_n % 32 == 0 && pix_tiles >= 2 && ctx->num_sms >= 2;
  const bool wins = P.block_n > 128 || P.kchunks == 1;
  P.pair = (legal && (ctx->pair_mode == 2 || (ctx->pair_mode == 1 && wins))) ? 1 : 0;
  P.b_bytes = (P.pair ? P.block_n / 2 : P.block_n) * 128;
  P.wide_b = (!P.pair && ctx->wide_b && !ctx->fast_mode && P.planes == 2 && P.block_n <= 128 && P.block_n % 8 == 0) ? 1 : 0;
}

struct EpilogueGeom {
  int sh, sw, oh, ow, rep, out_H, out_W;
};

// TMA tensor-store maps of one split-bf16 output: per index (parity class / upsample replica) a 4-D view (C, W', H', N)
// of the pixels {oy + step*h', ox + step*w'}, box {32 channels, BW, BH, BN}, SWIZZLE_64B (the staging layout of the
// epilogue).  *use = 0 when the layout cannot be addressed by TMA (the epilogue then scatters rows itself).
static int store_maps(dpig_ctx* ctx, CUtensorMap (*maps)[2], int* use, const dpig_tensor* t, const ConvUmmaParams& P,
                      const EpilogueGeom& g, bool replicas) {
  *use = 0;
  if (!ctx->epi_tma || t->pix_stride % 8 || t->c < 32 || P.cout < 32) return DPIG_OK;
  if (reinterpret_cast<uintptr_t>(t->hi) % 16 || reinterpret_cast<uintptr_t>(t->lo) % 16) return DPIG_OK;
  const int step = g.sh;
  int nidx = 1;
  int oy[4] = {g.oh, 0, 0, 0}, ox[4] = {g.ow, 0, 0, 0};
  if (P.nclass > 1) {
    nidx = P.nclass;
    for (int c = 0; c < nidx; ++c) {
      oy[c] = P.cls_oh[c];
      ox[c] = P.cls_ow[c];
    }
  } else if (g.rep > 1) {
    if (!replicas || g.rep != 2) return DPIG_OK;
    nidx = 4;
    for (int r = 0; r < 4; ++r) {
      oy[r] = r / 2;
      ox[r] = r % 2;
    }
  }
  const uint64_t ps = static_cast<uint64_t>(t->pix_stride);
  for (int i = 0; i < nidx; ++i) {
    const int W2 = (t->w - ox[i] + step - 1) / step, H2 = (t->h - oy[i] + step - 1) / step;
    if (W2 <= 0 || H2 <= 0) return DPIG_OK;
    uint64_t dims[4] = {static_cast<uint64_t>(t->c), static_cast<uint64_t>(W2), static_cast<uint64_t>(H2),
                        static_cast<uint64_t>(t->n)};
    uint64_t strides[3] = {ps * 2 * step, ps * 2 * t->w * step, ps * 2 * t->w * t->h};
    uint32_t box[4] = {32, static_cast<uint32_t>(P.BW), static_cast<uint32_t>(P.BH), static_cast<uint32_t>(P.BN)};
    for (int pln = 0; pln < 2; ++pln) {
      const void* plane = pln ? t->lo : t->hi;
      if (!plane) {
        maps[i][1] = maps[i][0];
        continue;
      }
      const char* base = static_cast<const char*>(plane) + (static_cast<uint64_t>(oy[i]) * t->w + ox[i]) * ps * 2;
      int rc = encode_map(ctx, &maps[i][pln], base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  }
  for (int i = nidx; i < 4; ++i)
    for (int pln = 0; pln < 2; ++pln) maps[i][pln] = maps[0][pln];
  *use = 1;
  return DPIG_OK;
}

static int fill_epilogue(dpig_ctx* ctx, ConvUmmaParams& P, const dpig_conv_epilogue* ep,
                         const EpilogueGeom& g, int n, int cout) {
  P.bias = ep->bias;
  P.class_bias = ep->class_bias;
  P.act = ep->act;
  P.alpha = ep->alpha;
  P.sh = g.sh;
  P.sw = g.sw;
  P.oh = g.oh;
  P.ow = g.ow;
  P.rep = g.rep;
  P.out_H = g.out_H;
  P.out_W = g.out_W;
  auto chk = [&](const dpig_tensor* t, const char* name) -> int {
    if (t->n != n || t->h != g.out_H || t->w != g.out_W || t->c < cout)
      return set_error(ctx, DPIG_EINVAL, "%s shape [%d,%d,%d,%d] does not match conv output [%d,%d,%d,%d]",
                       name, t->n, t->h, t->w, t->c, n, g.out_H, g.out_W, cout);
    return DPIG_OK;
  };
  int rc;
  if (ep->out) {
    if ((rc = chk(ep->out, "out"))) return rc;
    P.out_hi = static_cast<__nv_bfloat16*>(ep->out->hi);
    P.out_lo = static_cast<__nv_bfloat16*>(ep->out->lo);
    P.out_ps = ep->out->pix_stride;
    if ((rc = store_maps(ctx, P.out_map, &P.out_tma, ep->out, P, g, true))) return rc;
  }
  if (ep->out_masked) {
    if ((rc = chk(ep->out_masked, "out_masked"))) return rc;
    P.out2_hi = static_cast<__nv_bfloat16*>(ep->out_masked->hi);
    P.out2_lo = static_cast<__nv_bfloat16*>(ep->out_masked->lo);
    P.out2_ps = ep->out_masked->pix_stride;
    if ((rc = store_maps(ctx, P.out2_map, &P.out2_tma, ep->out_masked, P, g, false))) return rc;
  }
  if (ep->addend) {
    if ((rc = chk(ep->addend, "addend"))) return rc;
    P.add_hi = static_cast<const __nv_bfloat16*>(ep->addend->hi);
    P.add_lo = static_cast<const __nv_bfloat16*>(ep->addend->lo);
    P.add_ps = ep->addend->pix_stride;
    P.add_prefetch = ctx->add_prefetch ? 1 : 0;
  }
  P.out_f32 = ep->out_f32;
  P.out_f32_ps = ep->out_f32_pix_stride;
  P.mask_out = ep->mask_out;
  P.mask_out_words = (cout + 31) / 32;
  P.mask_in = ep->mask_in;
  P.mask_in_words = (cout + 31) / 32;
  P.mask_neg = ep->mask_neg;
  P.colsum = ep->colsum_masked;
  if (ep->colsum_masked && !ep->out_masked)
    return set_error(ctx, DPIG_EINVAL, "colsum_masked needs out_masked");
  P.stat_sums = ep->stat_sums;
  P.stat_mode = ep->stat_mode;
  P.stat_groups = ep->stat_mode == DPIG_NORM_BATCH ? cout : n;
  if (ep->stat_sums) {
    if (ep->stat_mode != DPIG_NORM_BATCH && ep->stat_mode != DPIG_NORM_LAYER)
      return set_error(ctx, DPIG_EUNSUPPORTED, "stat_sums: batch (per channel) or layer (per sample) statistics only");
    if (ep->act != DPIG_ACT_NONE || ep->addend || ep->out_masked || ep->colsum_masked || g.rep != 1 || P.nclass != 1)
      return set_error(ctx, DPIG_EINVAL, "stat_sums needs a plain conv epilogue (act none, no addend / out_masked / upsample)");
  }
  if (g.rep != 1 && (ep->addend || ep->out_masked))
    return set_error(ctx, DPIG_EUNSUPPORTED, "upsampling epilogue cannot take addend/out_masked");
  return DPIG_OK;
}

static int launch_conv(dpig_ctx* ctx, ConvUmmaParams& P, cudaStream_t stream) {
  P.planes = ctx->fast_mode ? 1 : P.planes;
  const uint32_t stage_bytes = P.planes * (kABytes + P.b_bytes);
  uint32_t extra = 1024 + 256 + kEpiBytes;  // alignment slack + barriers + epilogue staging
  int stages = (ctx->max_smem_optin - static_cast<int>(extra)) / static_cast<int>(stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) return set_error(ctx, DPIG_EUNSUPPORTED, "conv tile does not fit shared memory");
  if (ctx->max_stages >= 2 && stages > ctx->max_stages) stages = ctx->max_stages;
  // Second staging buffer per warp set: free when it costs no pipeline stage; at the price of the third stage where
  // the epilogue, not the K loop, sets the pace (two outputs per chunk, or a K loop of a few steps per tile).
  P.epi_bufs = 1;
  if (ctx->epi_bufs != 1 && (P.out_tma || P.out2_tma)) {
    const int stages2 = std::min(6, (ctx->max_smem_optin - static_cast<int>(extra + kEpiBytes)) / static_cast<int>(stage_bytes));
    const int steps = P.num_taps * P.kchunks / std::max(1, P.nclass);
    // (round 2, same-process A/B profiles/r02_ab_epi_bufs.txt: two-output data gradients with a K loop of more than four
    //  steps and no residual gather run 9-22 % FASTER with one buffer and the third pipeline stage; with the residual
    //  gather or a short K loop the second buffer wins by 3-8 %)
    const bool heavy = (P.out_tma && P.out2_tma && P.add_hi != nullptr) || steps <= 4;
    if (stages2 >= stages || (stages2 >= 2 && (heavy || ctx->epi_bufs == 2))) {
      P.epi_bufs = 2;
      stages = std::min(stages, stages2);
      extra += kEpiBytes;
    }
  }
  P.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + extra;
  // the smallest epilogue instantiation that covers the launch (EF_*; DPIG_EPI_SPECIALISE=0: always the generic one)
  int need = 0;
  if (P.bias || P.class_bias || P.act != DPIG_ACT_NONE || P.mask_out) need |= EF_FWD;
  if (P.out2_hi || P.colsum) need |= EF_GRAD;
  if (P.out_f32) need |= EF_F32;
  if (P.stat_sums) need |= EF_STATS;
  if (P.add_hi) need |= EF_ADD;
  const int kind = !ctx->epi_specialise ? 2 : ((need & ~kEpiFwd) == 0 ? 0 : ((need & ~kEpiGrad) == 0 ? 1 : 2));
  using Kern = void (*)(const ConvUmmaParams);
  static const Kern kernels[3][3] = {
      {conv_umma_kernel<false, false, kEpiFwd>, conv_umma_kernel<false, false, kEpiGrad>, conv_umma_kernel<false, false, EF_ALL>},
      {conv_umma_kernel<false, true, kEpiFwd>, conv_umma_kernel<false, true, kEpiGrad>, conv_umma_kernel<false, true, EF_ALL>},
      {conv_umma_kernel<true, false, kEpiFwd>, conv_umma_kernel<true, false, kEpiGrad>, conv_umma_kernel<true, false, EF_ALL>}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        cudaFuncSetAttribute(kernels[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
    attr_set = true;
  }
  const int pix_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  const int n_tiles = (P.cout + P.block_n - 1) / P.block_n;
  P.tmem_cols = 64;
  while (static_cast<int>(P.tmem_cols) < (P.wide_b ? 4 : 2) * P.block_n) P.tmem_cols <<= 1;
  if (P.stat_sums) cudaMemsetAsync(P.stat_sums, 0, sizeof(double) * 2 * P.stat_groups, stream);
  ctx->launches++;
  if (P.pair) {
    const int units = ((pix_tiles + 1) / 2) * n_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * std::min(units, ctx->num_sms / 2));
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[2][kind], P);
    if (e != cudaSuccess) return set_error(ctx, DPIG_ECUDA, "conv_umma_kernel<pair> launch: %s", cudaGetErrorString(e));
    return check_launch(ctx, "conv_umma_kernel<pair>");
  }
  dim3 grid(std::min(pix_tiles * n_tiles, ctx->num_sms));
  kernels[P.wide_b ? 1 : 0][kind]<<<grid, kConvThreads, smem, stream>>>(P);
  return check_launch(ctx, "conv_umma_kernel");
}

}  // namespace dpig

using namespace dpig;

extern "C" int dpig_conv2d_fwd(dpig_ctx* ctx, const dpig_tensor* x, const void* wf_hi,
                               const void* wf_lo, int32_t kh, int32_t kw, int32_t stride,
                               int32_t cout, const dpig_conv_epilogue* ep, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !wf_hi || !ep) return set_error(ctx, DPIG_EINVAL, "conv2d_fwd: null argument");
  if (kh * kw > kMaxTaps || (stride != 1 && stride != 2))
    return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_fwd: k=%dx%d stride=%d unsupported", kh, kw, stride);
  if (x->c % 8 || x->pix_stride % 8)
    return set_error(ctx, DPIG_EINVAL, "conv2d_fwd: channels / pixel stride must be multiples of 8");
  const int OH = same_out(x->h, stride), OW = same_out(x->w, stride);
  const int pt = same_pad_before(x->h, kh, stride), pl = same_pad_before(x->w, kw, stride);
  const int up = ep->upsample > 1 ? ep->upsample : 1;

  ConvUmmaParams P;
  memset(&P, 0, sizeof(P));
  P.planes = (x->lo && wf_lo) ? 2 : 1;
  const int planes = ctx->fast_mode ? 1 : P.planes;
  P.kchunks = (x->c + 63) / 64;
  P.Wo = OW;
  P.Ho = OH;
  P.No = x->n;
  P.cout = cout;
  Box b = choose_box(OW, OH, x->n, 128, false);
  P.BW = b.bw;
  P.BH = b.bh;
  P.BN = b.bn;
  P.tiles_w = (OW + b.bw - 1) / b.bw;
  P.tiles_h = (OH + b.bh - 1) / b.bh;
  P.tiles_n = (x->n + b.bn - 1) / b.bn;
  P.a_tx_bytes = b.bw * b.bh * b.bn * 128;
  P.block_n = tune_block_n(ctx, cout, P.tiles_w * P.tiles_h * P.tiles_n, kh * kw * P.kchunks);
  choose_pair(ctx, P);

  int rc;
  P.num_taps = 0;
  bool used[4] = {false, false, false, false};
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      const int dy = i - pt, dx = j - pl;
      ConvTap t;
      if (stride == 1) {
        t.src = 0;
        t.dh = dy;
        t.dw = dx;
      } else {
        const int py = ((dy % 2) + 2) % 2, px = ((dx % 2) + 2) % 2;
        t.src = py * 2 + px;
        t.dh = floordiv(dy, 2);
        t.dw = floordiv(dx, 2);
      }
      t.wtap = i * kw + j;
      used[t.src] = true;
      P.taps[P.num_taps++] = t;
    }
  // one TMA instruction per operand and stage when hi / lo sit in one allocation and the box fills the 16 KB slot
  const long long a_ps = (planes == 2 && ctx->merge_planes && P.a_tx_bytes == kABytes) ? plane_stride(x->hi, x->lo) : 0;
  const long long b_ps = (planes == 2 && ctx->merge_planes) ? plane_stride(wf_hi, wf_lo) : 0;
  P.a_merged = a_ps > 0;
  P.b_merged = b_ps > 0;
  int first_used = -1;
  for (int s = 0; s < 4; ++s) {
    if (!used[s]) continue;
    if (first_used < 0) first_used = s;
    for (int pln = 0; pln < (P.a_merged ? 1 : planes); ++pln) {
      const void* plane = pln ? x->lo : x->hi;
      if ((rc = act_map(ctx, &P.a_map[s][pln], plane, x, stride, stride == 1 ? 0 : s / 2,
                        stride == 1 ? 0 : s % 2, b.bw, b.bh, b.bn, a_ps)))
        return rc;
    }
    if (P.a_merged) P.a_map[s][1] = P.a_map[s][0];  // slot [1] is only prefetched
  }
  for (int s = 0; s < 4; ++s)  // unused slots alias a valid map (they are only prefetched)
    if (!used[s])
      for (int pln = 0; pln < planes; ++pln) P.a_map[s][pln] = P.a_map[first_used][pln];
  {
    const int cin_pad = x->c;
    uint64_t dims[3] = {static_cast<uint64_t>(cin_pad), static_cast<uint64_t>(cout),
                        static_cast<uint64_t>(kh * kw)};
    uint64_t strides[2] = {static_cast<uint64_t>(cin_pad) * 2, static_cast<uint64_t>(cin_pad) * cout * 2};
    uint32_t box[3] = {64, P.b_bytes / 128, 1};
    if (P.b_merged) {
      uint64_t dims4[4] = {dims[0], dims[1], dims[2], 2};
      uint64_t strides3[3] = {strides[0], strides[1], static_cast<uint64_t>(b_ps)};
      uint32_t box4[4] = {box[0], box[1], 1, 2};
      if ((rc = encode_map(ctx, &P.b_map[0], wf_hi, 4, dims4, strides3, box4))) return rc;
      P.b_map[1] = P.b_map[0];
    } else {
      for (int pln = 0; pln < planes; ++pln)
        if ((rc = encode_map(ctx, &P.b_map[pln], pln ? wf_lo : wf_hi, 3, dims, strides, box))) return rc;
    }
  }
  P.nclass = 1;
  P.cls_tap0[0] = 0;
  P.cls_tap0[1] = static_cast<int16_t>(P.num_taps);
  EpilogueGeom g{up, up, 0, 0, up, OH * up, OW * up};
  if ((rc = fill_epilogue(ctx, P, ep, g, x->n, cout))) return rc;
  return launch_conv(ctx, P, static_cast<cudaStream_t>(stream));
}

extern "C" int dpig_conv2d_bwd_data(dpig_ctx* ctx, const dpig_tensor* dy, const void* wb_hi,
                                    const void* wb_lo, int32_t kh, int32_t kw, int32_t stride,
                                    int32_t in_h, int32_t in_w, int32_t cin,
                                    const dpig_conv_epilogue* ep, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy || !wb_hi || !ep) return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_data: null argument");
  if (kh * kw > kMaxTaps || (stride != 1 && stride != 2))
    return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_bwd_data: k=%dx%d stride=%d unsupported", kh, kw, stride);
  if (dy->c % 8 || dy->pix_stride % 8)
    return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_data: channels / pixel stride must be multiples of 8");
  if (same_out(in_h, stride) != dy->h || same_out(in_w, stride) != dy->w)
    return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_data: dy %dx%d is not the SAME output of %dx%d / %d",
                     dy->h, dy->w, in_h, in_w, stride);
  if (ep->upsample > 1) return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_bwd_data: no upsample");
  const int pt = same_pad_before(in_h, kh, stride), pl = same_pad_before(in_w, kw, stride);
  int rc;
  // Stride 2: the four parity classes of dx share one launch when they have the same grid (even in_h, in_w), see
  // ConvUmmaParams::nclass; otherwise (and with DPIG_DGRAD_MERGE=0) one launch per class.
  const bool merged = stride == 2 && ctx->dgrad_merge && in_h % 2 == 0 && in_w % 2 == 0;
  const int nlaunch = merged ? 1 : stride * stride;
  for (int li = 0; li < nlaunch; ++li) {
    const int py0 = merged ? 0 : li / stride, px0 = merged ? 0 : li % stride;
    const int GH = (in_h - py0 + stride - 1) / stride, GW = (in_w - px0 + stride - 1) / stride;
    if (GH <= 0 || GW <= 0) continue;
    ConvUmmaParams P;
    memset(&P, 0, sizeof(P));
    P.planes = (dy->lo && wb_lo) ? 2 : 1;
    const int planes = ctx->fast_mode ? 1 : P.planes;
    P.kchunks = (dy->c + 63) / 64;
    P.Wo = GW;
    P.Ho = GH;
    P.No = dy->n;
    P.cout = cin;
    Box b = choose_box(GW, GH, dy->n, 128, false);
    P.BW = b.bw;
    P.BH = b.bh;
    P.BN = b.bn;
    P.tiles_w = (GW + b.bw - 1) / b.bw;
    P.tiles_h = (GH + b.bh - 1) / b.bh;
    P.tiles_n = (dy->n + b.bn - 1) / b.bn;
    P.a_tx_bytes = b.bw * b.bh * b.bn * 128;
    P.block_n = tune_block_n(ctx, cin, P.tiles_w * P.tiles_h * P.tiles_n, kh * kw * P.kchunks / (merged ? 1 : stride * stride));
    choose_pair(ctx, P);
    P.num_taps = 0;
    P.nclass = merged ? stride * stride : 1;
    for (int c = 0; c < P.nclass; ++c) {
      const int py = merged ? c / stride : py0, px = merged ? c % stride : px0;
      P.cls_tap0[c] = static_cast<int16_t>(P.num_taps);
      P.cls_oh[c] = static_cast<int16_t>(py);
      P.cls_ow[c] = static_cast<int16_t>(px);
      for (int i = 0; i < kh; ++i)
        for (int j = 0; j < kw; ++j) {
          // dx[s*a+py] gets dy[(s*a + py + pt - i)/s] when divisible
          const int ey = py + pt - i, ex = px + pl - j;
          if (((ey % stride) + stride) % stride || ((ex % stride) + stride) % stride) continue;
          ConvTap t;
          t.src = 0;
          t.dh = floordiv(ey, stride);
          t.dw = floordiv(ex, stride);
          t.wtap = i * kw + j;
          P.taps[P.num_taps++] = t;
        }
      if (P.num_taps == P.cls_tap0[c])
        return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_bwd_data: parity class without taps");
    }
    P.cls_tap0[P.nclass] = static_cast<int16_t>(P.num_taps);
    const long long a_ps =
        (planes == 2 && ctx->merge_planes && P.a_tx_bytes == kABytes) ? plane_stride(dy->hi, dy->lo) : 0;
    const long long b_ps = (planes == 2 && ctx->merge_planes) ? plane_stride(wb_hi, wb_lo) : 0;
    P.a_merged = a_ps > 0;
    P.b_merged = b_ps > 0;
    for (int pln = 0; pln < planes; ++pln) {
      if (pln == 0 || !P.a_merged) {
        if ((rc = act_map(ctx, &P.a_map[0][pln], pln ? dy->lo : dy->hi, dy, 1, 0, 0, b.bw, b.bh, b.bn, a_ps)))
          return rc;
      } else {
        P.a_map[0][1] = P.a_map[0][0];
      }
      for (int s = 1; s < 4; ++s) P.a_map[s][pln] = P.a_map[0][pln];
    }
    {
      const int cout_pad = dy->c;
      uint64_t dims[3] = {static_cast<uint64_t>(cout_pad), static_cast<uint64_t>(cin),
                          static_cast<uint64_t>(kh * kw)};
      uint64_t strides[2] = {static_cast<uint64_t>(cout_pad) * 2,
                             static_cast<uint64_t>(cout_pad) * cin * 2};
      uint32_t box[3] = {64, P.b_bytes / 128, 1};
      if (P.b_merged) {
        uint64_t dims4[4] = {dims[0], dims[1], dims[2], 2};
        uint64_t strides3[3] = {strides[0], strides[1], static_cast<uint64_t>(b_ps)};
        uint32_t box4[4] = {box[0], box[1], 1, 2};
        if ((rc = encode_map(ctx, &P.b_map[0], wb_hi, 4, dims4, strides3, box4))) return rc;
        P.b_map[1] = P.b_map[0];
      } else {
        for (int pln = 0; pln < planes; ++pln)
          if ((rc = encode_map(ctx, &P.b_map[pln], pln ? wb_lo : wb_hi, 3, dims, strides, box))) return rc;
      }
    }
    EpilogueGeom g{stride, stride, py0, px0, 1, in_h, in_w};
    if ((rc = fill_epilogue(ctx, P, ep, g, dy->n, cin))) return rc;
    if ((rc = launch_conv(ctx, P, static_cast<cudaStream_t>(stream)))) return rc;
  }
  return DPIG_OK;
}

static int wgrad_impl(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh, int32_t kw,
                      int32_t stride, int32_t cin, int32_t cout, float* dw, int32_t cin_pitch, dpig_stream stream);
static int wgrad_segment(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh, int32_t kw,
                         int32_t stride, int32_t cin, int32_t cout, float* dw, int32_t cin_pitch, int32_t cout_pitch,
                         dpig_stream stream);

extern "C" int dpig_conv2d_bwd_filter(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy,
                                      int32_t kh, int32_t kw, int32_t stride, int32_t cin,
                                      int32_t cout, float* dw, dpig_stream stream) {
  return wgrad_impl(ctx, x, dy, kh, kw, stride, cin, cout, dw, cin, stream);
}

extern "C" int dpig_conv2d_bwd_filter_rows(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy,
                                           int32_t kh, int32_t kw, int32_t stride, int32_t cin, int32_t cout,
                                           float* dw_rows, int32_t cin_total, dpig_stream stream) {
  return wgrad_impl(ctx, x, dy, kh, kw, stride, cin, cout, dw_rows, cin_total, stream);
}

// Output-channel counts that no 256-wide block divides (384, 640, 896: every other level of the pyramids) used to run
// as 192- / 128-wide blocks on the single-CTA kernel at 290-340 TFLOP/s; they are cut into 256-wide column segments
// plus the remainder, one launch each, so that all but the remainder run on the 2-CTA pair kernel (520-578 TFLOP/s).
// The segments write disjoint columns of dw.  (DPIG_WGRAD_SPLIT=0: one launch.)
static int wgrad_impl(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh, int32_t kw,
                      int32_t stride, int32_t cin, int32_t cout, float* dw, int32_t cin_pitch, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  // Measured (profiles/r02_wgrad_split_ab.txt): the segments win where there are many pixel steps to share out and
  // several input-channel tiles to pair (384 -> 384 on 32x16 / 12x12 maps: -10 %); few-pixel layers lose more to the two
  // extra launches and their partial-sum epilogues than the pair kernel gains (640 -> 640 on 8x4: +60 %).
  const long long out_pixels = (x && dy) ? static_cast<long long>(dy->n) * dy->h * dy->w : 0;
  if (x && dy && dw && ctx->wgrad_split && cout > 256 && cout % 256 != 0 && cout % 64 == 0 && dy->c >= cout &&
      out_pixels >= 32768 && cin >= 384) {
    for (int co0 = 0; co0 < cout; co0 += 256) {
      const int seg = std::min(256, cout - co0);
      dpig_tensor dyv = *dy;
      dyv.hi = static_cast<char*>(dy->hi) + static_cast<size_t>(co0) * 2;
      dyv.lo = dy->lo ? static_cast<char*>(dy->lo) + static_cast<size_t>(co0) * 2 : nullptr;
      dyv.c = seg;
      int rc = wgrad_segment(ctx, x, &dyv, kh, kw, stride, cin, seg, dw + co0, cin_pitch, cout, stream);
      if (rc) return rc;
    }
    return DPIG_OK;
  }
  return wgrad_segment(ctx, x, dy, kh, kw, stride, cin, cout, dw, cin_pitch, cout, stream);
}

static int wgrad_segment(dpig_ctx* ctx, const dpig_tensor* x, const dpig_tensor* dy, int32_t kh, int32_t kw,
                         int32_t stride, int32_t cin, int32_t cout, float* dw, int32_t cin_pitch, int32_t cout_pitch,
                         dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !dy || !dw) return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_filter: null argument");
  if (kh * kw > kMaxTaps || (stride != 1 && stride != 2))
    return set_error(ctx, DPIG_EUNSUPPORTED, "conv2d_bwd_filter: k=%dx%d stride=%d unsupported", kh, kw, stride);
  if (x->c % 8 || x->pix_stride % 8 || dy->c % 8 || dy->pix_stride % 8)
    return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_filter: channels / pixel stride must be multiples of 8");
  const int OH = same_out(x->h, stride), OW = same_out(x->w, stride);
  if (OH != dy->h || OW != dy->w || x->n != dy->n)
    return set_error(ctx, DPIG_EINVAL, "conv2d_bwd_filter: x/dy shape mismatch");
  const int pt = same_pad_before(x->h, kh, stride), pl = same_pad_before(x->w, kw, stride);

  WgradParams P;
  memset(&P, 0, sizeof(P));
  P.planes = (x->lo && dy->lo && !ctx->fast_mode) ? 2 : 1;
  P.cin = cin;
  P.cin_pitch = cin_pitch;
  P.cout_pitch = cout_pitch;
  P.cout = cout;
  P.vec_red = (ctx->wgrad_vec_red && cout_pitch % 4 == 0 && reinterpret_cast<uintptr_t>(dw) % 16 == 0) ? 1 : 0;
  const int c64 = (cout + 63) / 64 * 64;
  P.block_n = c64 <= 256 ? c64 : (c64 % 256 == 0 ? 256 : (c64 % 192 == 0 ? 192 : (c64 % 128 == 0 ? 128 : 64)));
  P.n_tiles = (cout + P.block_n - 1) / P.block_n;
  // taps per CTA (one fp32 accumulator each; at most 512 TMEM columns): 1 by default, see the kernel's header
  const int ntaps_total = kh * kw;
  int group = 1;
  if (ctx->wgrad_group > 0) group = std::min(ctx->wgrad_group, 512 / P.block_n);
  group = std::max(1, std::min(group, ntaps_total));
  P.group = group;
  P.num_taps = ntaps_total;
  const int tap_groups = (ntaps_total + group - 1) / group;
  uint32_t cols = pow2_cols(group * P.block_n);
  P.tmem_cols = cols < 64 ? 64 : cols;
  // pixels per pipeline step: 64 when >= 2 stages of (group x tiles + dy tile) fit, else 32
  const int planes_w = (x->lo && dy->lo && !ctx->fast_mode) ? 2 : 1;
  P.px = 64;
  if ((ctx->max_smem_optin - 1024 - 256) / (planes_w * (group * 2 * 64 * 128 + P.block_n * 64 * 2)) < 2) P.px = 32;
  if (ctx->wgrad_px == 32 || ctx->wgrad_px == 64) P.px = ctx->wgrad_px;
  P.x_grouped = (x->c % 64 == 0 && ctx->merge_planes) ? 1 : 0;
  P.dy_grouped = (dy->c % 64 == 0 && ctx->merge_planes) ? 1 : 0;
  P.a_bytes = 2 * P.px * 128;
  P.b_bytes = (P.block_n / 64) * P.px * 128;
  Box b = choose_box(OW, OH, x->n, P.px, true);
  P.PW = b.bw;
  P.PH = b.bh;
  P.PN = b.bn;
  P.tiles_w = (OW + b.bw - 1) / b.bw;
  P.tiles_h = (OH + b.bh - 1) / b.bh;
  const int tiles_n = (x->n + b.bn - 1) / b.bn;
  P.total_tiles = P.tiles_w * P.tiles_h * tiles_n;
  const int m_tiles = (cin + 127) / 128;
  P.m_tiles = m_tiles;
  // CTA pairs (see the kernel): two (tap, channel-tile) units per cluster sharing one dy tile, half of it staged by each
  // (measured, profiles/r01_wgrad_pair_ab.txt: 256-wide blocks gain 17-21 %, 128 -> 128 layers lose 2-4 %)
  const bool pair = ctx->wgrad_pair && group == 1 && P.px == 64 && P.x_grouped && P.dy_grouped && P.block_n % 128 == 0 &&
                    !(P.block_n == 128 && cin <= 128) && ntaps_total * m_tiles >= 2 && ctx->num_sms >= 2;
  P.wide_b = (!pair && ctx->wide_b && group == 1 && P.planes == 2 && P.block_n <= 128) ? 1 : 0;
  if (P.wide_b) P.tmem_cols = std::max<uint32_t>(64, pow2_cols(2 * P.block_n));
  const int base = pair ? 2 * ((ntaps_total * m_tiles + 1) / 2) * P.n_tiles : m_tiles * P.n_tiles * tap_groups;
  // split-K so that the grid is as close as possible to (but not above) a whole number of waves of the
  // 148 one-CTA-per-SM slots: a grid of 450 CTAs would run 4 rounds with the last one 4 % full.
  int ksplit = 1;
  {
    double best = -1.0;
    const int max_split = std::min(P.total_tiles, std::max(1, 4 * ctx->num_sms / base));
    const int fixed = 6 * 64 / P.px;   // per-CTA prologue / atomics epilogue, in pipeline steps
    for (int ks = 1; ks <= max_split; ++ks) {
      const int tpc = (P.total_tiles + ks - 1) / ks;
      const int eff_ks = (P.total_tiles + tpc - 1) / tpc;
      const int ctas = base * eff_ks;
      const int rounds = (ctas + ctx->num_sms - 1) / ctx->num_sms;
      const double cost = static_cast<double>(rounds) * (tpc + fixed);
      if (best < 0 || cost < best) {
        best = cost;
        ksplit = eff_ks;
      }
    }
  }
  P.tiles_per_cta = (P.total_tiles + ksplit - 1) / ksplit;
  ksplit = (P.total_tiles + P.tiles_per_cta - 1) / P.tiles_per_cta;
  P.dw = dw;

  int rc;
  int ntaps = 0;
  bool used[4] = {false, false, false, false};
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      const int dyy = i - pt, dxx = j - pl;
      ConvTap t;
      if (stride == 1) {
        t.src = 0;
        t.dh = dyy;
        t.dw = dxx;
      } else {
        const int py = ((dyy % 2) + 2) % 2, px = ((dxx % 2) + 2) % 2;
        t.src = py * 2 + px;
        t.dh = floordiv(dyy, 2);
        t.dw = floordiv(dxx, 2);
      }
      t.wtap = i * kw + j;
      used[t.src] = true;
      P.taps[ntaps++] = t;
    }
  for (int s = 0; s < 4; ++s) {
    if (!used[s]) continue;
    for (int pln = 0; pln < P.planes; ++pln)
      if ((rc = act_map(ctx, &P.x_map[s][pln], pln ? x->lo : x->hi, x, stride, stride == 1 ? 0 : s / 2,
                        stride == 1 ? 0 : s % 2, b.bw, b.bh, b.bn, 0, P.x_grouped ? 2 : 0)))
        return rc;
  }
  for (int pln = 0; pln < P.planes; ++pln)
    if ((rc = act_map(ctx, &P.dy_map[pln], pln ? dy->lo : dy->hi, dy, 1, 0, 0, b.bw, b.bh, b.bn, 0,
                      P.dy_grouped ? P.block_n / (pair ? 128 : 64) : 0)))
      return rc;

  const uint32_t stage_bytes = P.planes * (P.group * P.a_bytes + (pair ? P.b_bytes / 2 : P.b_bytes));
  int stages = (ctx->max_smem_optin - 1024 - 256) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return set_error(ctx, DPIG_EUNSUPPORTED, "wgrad tile does not fit shared memory");
  P.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
    cudaFuncSetAttribute(wgrad_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
    attr_set = true;
  }
  ctx->launches++;
  if (pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(base, 1, ksplit);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_umma_kernel<true>, P);
    if (e != cudaSuccess) return set_error(ctx, DPIG_ECUDA, "wgrad_umma_kernel<pair> launch: %s", cudaGetErrorString(e));
    return check_launch(ctx, "wgrad_umma_kernel<pair>");
  }
  dim3 grid(m_tiles * P.n_tiles, tap_groups, ksplit);
  wgrad_umma_kernel<false><<<grid, 192, smem, static_cast<cudaStream_t>(stream)>>>(P);
  return check_launch(ctx, "wgrad_umma_kernel");
}
