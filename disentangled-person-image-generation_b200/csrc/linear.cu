// Fully connected layers (slim.fully_connected, reference models.py:431, 464, 478-484, 491-512, 545,
// 554; tflib/ops/linear.py:133-147) as fp32 CUDA-core GEMMs.  The shapes are skinny (M = batch or
// 7*batch rows, K up to 20480, N = 32..4096): weight-bandwidth bound, so the kernel is a plain
// 64x64x16 shared-memory tiled SGEMM with split-K over grid.z (fp32 atomics) to fill the 148 SMs.
#include "common.cuh"

namespace dpig {

// C[M,N] (+)= op(A)[M,K] * op(B)[K,N];  TA: A stored [K,M];  TB: B stored [N,K].
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* C, int M, int N, int K,
             int kchunk, const float* bias, int act, float alpha, bool atomic) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int k_begin = blockIdx.z * kchunk;
  const int k_end = min(K, k_begin + kchunk);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    // 64x16 tile of A and 16x64 tile of B, 4 elements per thread each
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 256;
      {
        int m, k;
        if (TA) { m = idx % 64; k = idx / 64; } else { k = idx % 16; m = idx / 16; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.f;
        if (gm < M && gk < k_end) v = TA ? A[static_cast<long long>(gk) * M + gm] : A[static_cast<long long>(gm) * K + gk];
        As[k][m] = v;
      }
      {
        int n, k;
        if (TB) { k = idx % 16; n = idx / 16; } else { n = idx % 64; k = idx / 64; }
        const int gn = n0 + n, gk = k0 + k;
        float v = 0.f;
        if (gn < N && gk < k_end) v = TB ? B[static_cast<long long>(gn) * K + gk] : B[static_cast<long long>(gk) * N + gn];
        Bs[k][n] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += bias[gn];
      float* c = C + static_cast<long long>(gm) * N + gn;
      if (atomic) {
        atomicAdd(c, v);
      } else {
        if (act == DPIG_ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == DPIG_ACT_LRELU) v = v > 0.f ? v : alpha * v;
        *c = v;
      }
    }
  }
}

__global__ void colsum_kernel(const float* dy, int M, int N, float* db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int m = 0; m < M; ++m) acc += dy[static_cast<long long>(m) * N + n];
  atomicAdd(db + n, acc);
}

// C must be zero-initialised (or hold the value to accumulate onto) when accumulate = true.
template <bool TA, bool TB>
static int run_gemm(dpig_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K,
                    const float* bias, int act, float alpha, bool accumulate, cudaStream_t s) {
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  int splitk = 1;
  if (act == DPIG_ACT_NONE) {
    splitk = (2 * ctx->num_sms + tiles - 1) / tiles;
    const int maxsplit = (K + 63) / 64;
    if (splitk > maxsplit) splitk = maxsplit;
    if (splitk < 1) splitk = 1;
  }
  int kchunk = ((K + splitk - 1) / splitk + 15) / 16 * 16;
  splitk = (K + kchunk - 1) / kchunk;
  const bool atomic = accumulate || splitk > 1;
  if (atomic && !accumulate) cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, s);
  dim3 grid((N + 63) / 64, (M + 63) / 64, splitk);
  sgemm_kernel<TA, TB><<<grid, 256, 0, s>>>(A, B, C, M, N, K, kchunk, bias, act, alpha, atomic);
  ctx->launches++;
  return check_launch(ctx, "sgemm");
}

}  // namespace dpig
using namespace dpig;

extern "C" int dpig_linear_fwd(dpig_ctx* ctx, const float* x, const float* w, const float* b, float* y,
                               int32_t m, int32_t k, int32_t n, int32_t act, float alpha, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!x || !w || !y) return set_error(ctx, DPIG_EINVAL, "linear_fwd: null argument");
  return run_gemm<false, false>(ctx, x, w, y, m, n, k, b, act, alpha, false, static_cast<cudaStream_t>(stream));
}

extern "C" int dpig_linear_bwd(dpig_ctx* ctx, const float* x, const float* w, const float* dy, float* dx,
                               float* dw, float* db, int32_t m, int32_t k, int32_t n, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!dy) return set_error(ctx, DPIG_EINVAL, "linear_bwd: null dy");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  // dx[m,k] = dy[m,n] * w[k,n]^T
  if (dx && (rc = run_gemm<false, true>(ctx, dy, w, dx, m, k, n, nullptr, DPIG_ACT_NONE, 0.f, false, s))) return rc;
  // dw[k,n] += x[m,k]^T * dy[m,n]
  if (dw && (rc = run_gemm<true, false>(ctx, x, dy, dw, k, n, m, nullptr, DPIG_ACT_NONE, 0.f, true, s))) return rc;
  if (db) {
    colsum_kernel<<<(n + 127) / 128, 128, 0, s>>>(dy, m, n, db);
    ctx->launches++;
    if ((rc = check_launch(ctx, "colsum"))) return rc;
  }
  return DPIG_OK;
}
