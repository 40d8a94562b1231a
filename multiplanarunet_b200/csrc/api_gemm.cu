// C-ABI entry points for the raw multi-tap GEMM kernels (kernel-level parity tests and bring-up).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mpunet_b200.h"
#include "common.h"
#include "mtgemm.cuh"

namespace mpu {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
}  // namespace mpu

using namespace mpu;

extern "C" {

const char* mpu_last_error(void) { return mpu::last_error(); }

int mpu_version(void) { return 100; }

int mpu_mtgemm_fwd(const void* A0, long long rowsA0, int C0, int ldA0, const void* A1,
                   long long rowsA1, int C1, int ldA1, const void* W, int w_taps, int n_phys,
                   int k_total, int ntaps, const int* tap_a_off, const int* tap_w, int M_rows, int BN,
                   int Hp, int Wp, int oHp, int oWp, int s, int py, int px, void* out, int ldo,
                   const float* bias, const void* mask, int ldm, int relu, void* stream) {
  if (!A0 || !W || !out || ntaps < 1 || ntaps > kMaxTaps) {
    set_error("mpu_mtgemm_fwd: bad arguments");
    return MPU_ERR_ARG;
  }
  FwdParams p;
  memset(&p, 0, sizeof(p));
  MPU_TRY(make_tmap_2d(&p.tmA0, A0, (uint64_t)rowsA0, (uint64_t)C0, (uint64_t)ldA0, 64, 128));
  p.chunks0 = (C0 + 63) / 64;
  if (A1) {
    MPU_TRY(make_tmap_2d(&p.tmA1, A1, (uint64_t)rowsA1, (uint64_t)C1, (uint64_t)ldA1, 64, 128));
    p.chunks1 = (C1 + 63) / 64;
    p.kofs1 = C0;
  }
  MPU_TRY(make_tmap_2d(&p.tmB, W, (uint64_t)w_taps * n_phys, (uint64_t)k_total, (uint64_t)k_total, 64,
                       (uint32_t)BN));
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    p.tap_a_off[t] = tap_a_off[t];
    p.tap_w[t] = tap_w[t];
  }
  p.w_rows_per_tap = n_phys;
  p.M_rows = M_rows;
  p.n_valid = n_phys;
  p.BN = BN;
  p.map = RowMap{Hp, Wp, oHp, oWp, s, py, px};
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.bias = bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask);
  p.ldm = ldm;
  p.relu = relu;
  return launch_fwd(p, reinterpret_cast<cudaStream_t>(stream));
}

int mpu_mtgemm_wgrad(const void* X, long long rowsX, int Cx, int ldX, const void* dY,
                     long long rowsDY, int Cy, int ldDY, int ntaps, const int* tap_x_off,
                     const int* tap_w, int ngroups, const int* group_first, const int* group_count,
                     const int* group_dy_off, int rows_total, int BN, int splits, float* dW, int ldw,
                     int w_rows_per_tap, int dw_col0, int ci_valid, int co_valid, int a_lbo, int a_sbo,
                     int b_lbo, int b_sbo, int kstep_bytes, void* stream) {
  if (!X || !dY || !dW || ntaps < 1 || ntaps > kMaxTaps || ngroups < 1 || ngroups > kMaxTaps) {
    set_error("mpu_mtgemm_wgrad: bad arguments");
    return MPU_ERR_ARG;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  MPU_TRY(make_tmap_2d(&p.tmX, X, (uint64_t)rowsX, (uint64_t)Cx, (uint64_t)ldX, 64, 64));
  MPU_TRY(make_tmap_2d(&p.tmDY, dY, (uint64_t)rowsDY, (uint64_t)Cy, (uint64_t)ldDY, 64, 64));
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    p.tap_x_off[t] = tap_x_off[t];
    p.tap_w[t] = tap_w[t];
  }
  p.ngroups = ngroups;
  for (int g = 0; g < ngroups; ++g) p.groups[g] = WgradGroup{group_first[g], group_count[g], group_dy_off[g]};
  p.BN = BN;
  p.ci_tiles = (ci_valid + 127) / 128;
  p.co_tiles = (co_valid + BN - 1) / BN;
  p.kblocks = (rows_total + 63) / 64;
  if (splits < 1) splits = 1;
  if (splits > p.kblocks) splits = p.kblocks;
  p.kblocks_per_split = (p.kblocks + splits - 1) / splits;
  p.splits = (p.kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.dW = dW;
  p.ldw = ldw;
  p.w_rows_per_tap = w_rows_per_tap;
  p.dw_col0 = dw_col0;
  p.ci_valid = ci_valid;
  p.co_valid = co_valid;
  p.a_lbo = a_lbo > 0 ? a_lbo : 8192;
  p.a_sbo = a_sbo > 0 ? a_sbo : 1024;
  p.b_lbo = b_lbo > 0 ? b_lbo : 8192;
  p.b_sbo = b_sbo > 0 ? b_sbo : 1024;
  p.kstep_bytes = kstep_bytes > 0 ? kstep_bytes : 2048;
  return launch_wgrad(p, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
