// Image-side ends of the sampling path on the GPU: the SSIM that the reference's generate() computes per sample on
// the host with scikit-image (trainer.py:514-526, tester.py:236-241: rgb2gray of the uint8 images, then
// skimage.measure.compare_ssim(G_gray, x_gray, data_range = x_gray.max() - x_gray.min(), multichannel=False)).
//
// scikit-image is an un-vendored dependency of the reference (imported at trainer.py:15-16, no pinned version); the
// algorithm restated here is its published default path (Wang et al. 2004 as implemented by compare_ssim /
// structural_similarity): 7x7 uniform window, K1 = 0.01, K2 = 0.03, SAMPLE covariance (normalised by N/(N-1), N = 49),
// float64 arithmetic, mean of the SSIM map over the windows that lie fully inside the image (crop of (7-1)/2 pixels);
// rgb2gray = 0.2125 R + 0.7154 G + 0.0721 B on the [0,1]-scaled image.  oracle/image_metrics.py is the CPU restatement.
#include "common.cuh"

namespace dpig {

__device__ __forceinline__ double gray_u8(const uint8_t* p) {
  return (0.2125 * p[0] + 0.7154 * p[1] + 0.0721 * p[2]) / 255.0;
}

__device__ __forceinline__ double block_reduce(double v, double* sh, int op) {  // op 0 sum, 1 min, 2 max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 16; off >= 1; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, v, off);
    v = op == 0 ? v + o : (op == 1 ? fmin(v, o) : fmax(v, o));
  }
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = sh[0];
  for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) r = op == 0 ? r + sh[i] : (op == 1 ? fmin(r, sh[i]) : fmax(r, sh[i]));
  return r;
}

// one block per image pair; a, b: uint8 [n][h][w][3]
__global__ void ssim_gray_u8_kernel(const uint8_t* a, const uint8_t* b, int h, int w, float* out) {
  __shared__ double sh[32];
  const long long img = static_cast<long long>(blockIdx.x) * h * w * 3;
  const uint8_t* pa = a + img;
  const uint8_t* pb = b + img;
  // data_range = max - min of the SECOND image's gray values
  double mn = 1e300, mx = -1e300;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const double g = gray_u8(pb + 3 * i);
    mn = fmin(mn, g);
    mx = fmax(mx, g);
  }
  mn = block_reduce(mn, sh, 1);
  mx = block_reduce(mx, sh, 2);
  const double R = mx - mn;
  const double C1 = (0.01 * R) * (0.01 * R), C2 = (0.03 * R) * (0.03 * R);
  const int oh = h - 6, ow = w - 6;
  double acc = 0.0;
  for (int i = threadIdx.x; i < oh * ow; i += blockDim.x) {
    const int y0 = i / ow, x0 = i % ow;
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int dy = 0; dy < 7; ++dy)
      for (int dx = 0; dx < 7; ++dx) {
        const int o = 3 * ((y0 + dy) * w + x0 + dx);
        const double x = gray_u8(pa + o), y = gray_u8(pb + o);
        sx += x;
        sy += y;
        sxx += x * x;
        syy += y * y;
        sxy += x * y;
      }
    const double ux = sx / 49.0, uy = sy / 49.0;
    const double cn = 49.0 / 48.0;
    const double vx = cn * (sxx / 49.0 - ux * ux), vy = cn * (syy / 49.0 - uy * uy), vxy = cn * (sxy / 49.0 - ux * uy);
    acc += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
  }
  acc = block_reduce(acc, sh, 0);
  if (threadIdx.x == 0) out[blockIdx.x] = static_cast<float>(acc / (static_cast<double>(oh) * ow));
}

}  // namespace dpig

using namespace dpig;

extern "C" int dpig_ssim_gray_u8(dpig_ctx* ctx, const uint8_t* a, const uint8_t* b, int32_t n, int32_t h, int32_t w,
                                 float* out, dpig_stream stream) {
  DPIG_CHECK_CTX(ctx);
  if (!a || !b || !out) return set_error(ctx, DPIG_EINVAL, "ssim_gray_u8: null argument");
  if (h < 7 || w < 7) return set_error(ctx, DPIG_EINVAL, "ssim_gray_u8: image smaller than the 7x7 window");
  if (n <= 0) return DPIG_OK;
  ssim_gray_u8_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, h, w, out);
  ctx->launches++;
  return check_launch(ctx, "ssim_gray_u8");
}
