// Context management and error plumbing of libdpig.so (C ABI in include/dpig.h).
#include <cstdarg>
#include <cstdio>
#include <stdlib.h>
#include "common.cuh"
#include <string.h>

namespace dpig {

int set_error(dpig_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  return code;
}

int check_launch(dpig_ctx* ctx, const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(ctx, DPIG_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return DPIG_OK;
}

}  // namespace dpig

using namespace dpig;

extern "C" int dpig_ctx_create(int device, dpig_ctx** out) {
  if (!out) return DPIG_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return DPIG_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DPIG_ECUDA;
  if (prop.major != 10) return DPIG_ENODEVICE;  // kernels are sm_100a only; there is no fallback
  if (cudaSetDevice(device) != cudaSuccess) return DPIG_ECUDA;
  dpig_ctx* ctx = new dpig_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      !fn) {
    delete ctx;
    return DPIG_ECUDA;
  }
  ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
  if (const char* e = getenv("DPIG_CONV_PAIR")) ctx->pair_mode = atoi(e);  // A/B switches for the conv tilings
  if (const char* e = getenv("DPIG_WGRAD_GROUP")) ctx->wgrad_group = atoi(e);
  if (const char* e = getenv("DPIG_WGRAD_PX")) ctx->wgrad_px = atoi(e);
  if (const char* e = getenv("DPIG_CONV_STAGES")) ctx->max_stages = atoi(e);
  if (const char* e = getenv("DPIG_CONV_MERGE")) ctx->merge_planes = atoi(e) != 0;
  if (const char* e = getenv("DPIG_DGRAD_MERGE")) ctx->dgrad_merge = atoi(e) != 0;
  if (const char* e = getenv("DPIG_EPI_TMA")) ctx->epi_tma = atoi(e) != 0;
  if (const char* e = getenv("DPIG_WGRAD_PAIR")) ctx->wgrad_pair = atoi(e) != 0;
  if (const char* e = getenv("DPIG_WGRAD_SPLIT")) ctx->wgrad_split = atoi(e) != 0;
  if (const char* e = getenv("DPIG_WGRAD_VEC_RED")) ctx->wgrad_vec_red = atoi(e) != 0;
  if (const char* e = getenv("DPIG_WIDE_B")) ctx->wide_b = atoi(e) != 0;
  if (const char* e = getenv("DPIG_EPI_BUFS")) ctx->epi_bufs = atoi(e);
  if (const char* e = getenv("DPIG_TUNE_SMALL")) ctx->tune_small = atoi(e);
  if (const char* e = getenv("DPIG_ADD_PREFETCH")) ctx->add_prefetch = atoi(e) != 0;
  if (const char* e = getenv("DPIG_CROP_GATHER")) ctx->crop_gather = atoi(e) != 0;
  if (const char* e = getenv("DPIG_EPI_SPECIALISE")) ctx->epi_specialise = atoi(e) != 0;
  *out = ctx;
  return DPIG_OK;
}

extern "C" void dpig_ctx_destroy(dpig_ctx* ctx) { delete ctx; }

extern "C" const char* dpig_last_error(const dpig_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : "null context";
}

extern "C" int dpig_ctx_set_fast_mode(dpig_ctx* ctx, int fast) {
  DPIG_CHECK_CTX(ctx);
  ctx->fast_mode = fast != 0;
  return DPIG_OK;
}

extern "C" int dpig_ctx_set_pair_mode(dpig_ctx* ctx, int mode) {
  DPIG_CHECK_CTX(ctx);
  if (mode < 0 || mode > 2) return set_error(ctx, DPIG_EINVAL, "pair mode must be 0, 1 or 2");
  ctx->pair_mode = mode;
  return DPIG_OK;
}

// Tuning switches by name (the DPIG_* environment variables read at context creation, settable at run time: A/B
// measurements inside one process, tests).  Results are identical under every setting.
extern "C" int dpig_ctx_set_option(dpig_ctx* ctx, const char* name, int value) {
  DPIG_CHECK_CTX(ctx);
  if (!name) return set_error(ctx, DPIG_EINVAL, "set_option: null name");
  const std::string n(name);
  if (n == "epi_specialise") ctx->epi_specialise = value != 0;
  else if (n == "wgrad_split") ctx->wgrad_split = value != 0;
  else if (n == "wgrad_pair") ctx->wgrad_pair = value != 0;
  else if (n == "wgrad_group") ctx->wgrad_group = value;
  else if (n == "wgrad_vec_red") ctx->wgrad_vec_red = value != 0;
  else if (n == "wide_b") ctx->wide_b = value != 0;
  else if (n == "epi_bufs") ctx->epi_bufs = value;
  else if (n == "epi_tma") ctx->epi_tma = value != 0;
  else if (n == "dgrad_merge") ctx->dgrad_merge = value != 0;
  else if (n == "add_prefetch") ctx->add_prefetch = value != 0;
  else if (n == "tune_small") ctx->tune_small = value;
  else if (n == "crop_gather") ctx->crop_gather = value != 0;
  else if (n == "max_stages") ctx->max_stages = value;
  else if (n == "merge_planes") ctx->merge_planes = value != 0;
  else return set_error(ctx, DPIG_EINVAL, "set_option: unknown option '%s'", name);
  return DPIG_OK;
}

extern "C" unsigned long long dpig_launch_count(const dpig_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" const char* dpig_version(void) { return "dpig-b200 0.1 (sm_100a)"; }

// ---- host utility: CRC-32C (Castagnoli) for the TensorFlow checkpoint reader / writer (tf_checkpoint.py).
// Slicing-by-8 table version; ~1 GB/s per core, enough for the 0.5 GB of Stage-I parameters.
namespace {
struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int j = 1; j < 8; ++j) t[j][i] = (t[j - 1][i] >> 8) ^ t[0][t[j - 1][i] & 0xFF];
  }
};
}  // namespace

extern "C" uint32_t dpig_crc32c(uint32_t crc, const void* data, size_t n) {
  static const Crc32cTables tab;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    c = tab.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = tab.t[7][w & 0xFF] ^ tab.t[6][(w >> 8) & 0xFF] ^ tab.t[5][(w >> 16) & 0xFF] ^ tab.t[4][(w >> 24) & 0xFF] ^
        tab.t[3][(w >> 32) & 0xFF] ^ tab.t[2][(w >> 40) & 0xFF] ^ tab.t[1][(w >> 48) & 0xFF] ^ tab.t[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
