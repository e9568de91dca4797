// C-ABI of libmmduet_b200.so (declared in include/mmduet_b200.h).
#include "../../include/mmduet_b200.h"
#include "gemm.cuh"

#include <string>

namespace {
thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; }
}  // namespace

struct mmd_ctx {
  int device;
  mmd::GemmContext* gemm;
};

extern "C" {

const char* mmd_version(void) { return "mmduet_b200 0.1 (sm_100a)"; }
const char* mmd_last_error(void) { return g_err.c_str(); }

mmd_ctx* mmd_create(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_err("cudaGetDeviceProperties failed"); return nullptr; }
  if (prop.major != 10) {
    set_err("mmduet_b200 needs an sm_100a (B200) device; found sm_" + std::to_string(prop.major * 10 + prop.minor));
    return nullptr;
  }
  mmd::GemmContext* g = mmd::gemm_context_create(device);
  if (g == nullptr) { set_err(mmd::gemm_last_error()); return nullptr; }
  mmd_ctx* c = new mmd_ctx();
  c->device = device;
  c->gemm = g;
  return c;
}

void mmd_destroy(mmd_ctx* c) {
  if (c == nullptr) return;
  mmd::gemm_context_destroy(c->gemm);
  delete c;
}

int mmd_gemm_bf16(mmd_ctx* c, int epi, int act, const void* X, const void* X2, int64_t x_rows, int64_t ldx,
                  const void* Y, int64_t y_rows, int64_t ldy, int64_t K, const float* bias, void* out, int64_t ldo,
                  int k_splits, int64_t split_stride, void* stream) {
  if (c == nullptr) { set_err("null context"); return MMD_ERR_ARG; }
  mmd::GemmArgs a;
  a.X = static_cast<const __nv_bfloat16*>(X);
  a.X2 = static_cast<const __nv_bfloat16*>(X2);
  a.Y = static_cast<const __nv_bfloat16*>(Y);
  a.x_rows = (int)x_rows; a.y_rows = (int)y_rows; a.K = (int)K;
  a.ldx = ldx; a.ldy = ldy; a.epi = epi; a.act = act; a.bias = bias; a.out = out; a.ldo = ldo;
  a.k_splits = k_splits; a.split_stride = split_stride;
  int rc = mmd::gemm_launch(c->gemm, a, static_cast<cudaStream_t>(stream));
  if (rc != 0) set_err(mmd::gemm_last_error());
  return rc;
}

int mmd_gemm_splits(int64_t K, int k_splits) { return mmd::gemm_effective_splits((int)K, k_splits); }

}  // extern "C"
