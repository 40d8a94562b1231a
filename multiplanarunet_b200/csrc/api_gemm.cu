// C-ABI entry points for the raw multi-tap GEMM kernels (kernel-level parity tests and bring-up).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

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

long long g_launch_count = 0;

int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) return 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// GEMM event timer: when enabled, every tensor-core GEMM launch is bracketed by a CUDA event pair on
// its own stream; mpu_profile_gemm_read() synchronises and sums the elapsed times.
static bool g_timer_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
static size_t g_events_used = 0;
bool gemm_timer_on() { return g_timer_on; }
void gemm_timer_begin(cudaStream_t st) {
  if (!g_timer_on) return;
  if (g_events_used == g_events.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    g_events.push_back({a, b});
  }
  cudaEventRecord(g_events[g_events_used].first, st);
}
void gemm_timer_end(cudaStream_t st) {
  if (!g_timer_on) return;
  cudaEventRecord(g_events[g_events_used].second, st);
  ++g_events_used;
}
}  // namespace mpu

using namespace mpu;

extern "C" {

const char* mpu_last_error(void) { return mpu::last_error(); }

int mpu_version(void) { return 100; }

long long mpu_launch_count(void) { return mpu::g_launch_count; }

int mpu_profile_gemm(int enable) {
  mpu::g_timer_on = enable != 0;
  mpu::g_events_used = 0;
  return MPU_OK;
}

int mpu_profile_gemm_read(double* total_ms, int* launches) {
  double tot = 0;
  for (size_t i = 0; i < mpu::g_events_used; ++i) {
    cudaError_t e = cudaEventSynchronize(mpu::g_events[i].second);
    if (e != cudaSuccess) {
      set_error("mpu_profile_gemm_read: %s", cudaGetErrorString(e));
      return MPU_ERR_CUDA;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, mpu::g_events[i].first, mpu::g_events[i].second);
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)mpu::g_events_used;
  mpu::g_events_used = 0;
  return MPU_OK;
}

int mpu_mtgemm_fwd(const void* A0, long long rowsA0, int C0, int ldA0, const void* A1,
                   long long rowsA1, int C1, int ldA1, const void* W, int w_taps, int n_phys,
                   int k_total, int ntaps, const int* tap_a_off, const int* tap_w, int M_rows,
                   int Hp, int Wp, int oHp, int oWp, int s, int py, int px, void* out, int ldo,
                   const float* bias, const void* mask, int ldm, int relu, void* stream) {
  if (!A0 || !W || !out || ntaps < 1 || ntaps > kMaxTaps) {
    set_error("mpu_mtgemm_fwd: bad arguments");
    return MPU_ERR_ARG;
  }
  FwdDesc d;
  memset(&d, 0, sizeof(d));
  d.A0 = A0; d.rowsA0 = rowsA0; d.C0 = C0; d.ldA0 = ldA0;
  d.A1 = A1; d.rowsA1 = rowsA1; d.C1 = C1; d.ldA1 = ldA1;
  d.W = W; d.w_taps = w_taps; d.n_phys = n_phys; d.k_total = k_total;
  d.ntaps = ntaps; d.tap_a_off = tap_a_off; d.tap_w = tap_w;
  d.M_rows = M_rows;
  d.map = RowMap{Hp, Wp, oHp, oWp, s, py, px};
  d.out = out; d.ldo = ldo; d.bias = bias; d.mask = mask; d.ldm = ldm; d.relu = relu;
  d.stats = nullptr;
  FwdParams p;
  MPU_TRY(fwd_setup(p, d));
  return launch_fwd(p, reinterpret_cast<cudaStream_t>(stream));
}

int mpu_mtgemm_wgrad(const void* X, long long rowsX, int Cx, int ldX, const void* dY,
                     long long rowsDY, int Cy, int ldDY, int ntaps, const int* tap_x_off,
                     const int* tap_dy_off, const int* tap_w, long long rows_total, int splits, float* dW,
                     int ldw, int w_rows_per_tap, int dw_col0, void* stream) {
  if (!X || !dY || !dW || !tap_x_off || !tap_w || ntaps < 1 || ntaps > kMaxTaps) {
    set_error("mpu_mtgemm_wgrad: bad arguments");
    return MPU_ERR_ARG;
  }
  WgradDesc d;
  memset(&d, 0, sizeof(d));
  d.X = X; d.rowsX = rowsX; d.Cx = Cx; d.ldX = ldX;
  d.dY = dY; d.rowsDY = rowsDY; d.Cy = Cy; d.ldDY = ldDY;
  d.ntaps = ntaps; d.tap_x_off = tap_x_off; d.tap_dy_off = tap_dy_off; d.tap_w = tap_w;
  d.rows_total = rows_total;
  d.splits = splits;
  d.dW = dW; d.ldw = ldw; d.w_rows_per_tap = w_rows_per_tap; d.dw_col0 = dw_col0;
  WgradParams p;
  MPU_TRY(wgrad_setup(p, d));
  return launch_wgrad(p, reinterpret_cast<cudaStream_t>(stream));
}

// Bring-up / test entry (not part of include/mpunet_b200.h): the work decomposition wgrad_setup would choose for a
// weight-gradient GEMM of these shapes, without touching a device.  out[0..11] = grid, splits, splits_part,
// kblocks_per_split, kblocks_per_split_part, kblocks, ci_tiles_full, ci_tiles, co_tiles, ngroups, n_entries, CA;
// out[12 .. 12 + n_entries) = the launch-order table.  `out` holds at least 12 + 512 ints.
int mpu_debug_wgrad_plan(int Cx, int Cy, long long rows_total, int ntaps, const int* tap_x_off, const int* tap_dy_off,
                         int* out) {
  if (!out || !tap_x_off || ntaps < 1 || ntaps > kMaxTaps) {
    set_error("mpu_debug_wgrad_plan: bad arguments");
    return MPU_ERR_ARG;
  }
  int widx[kMaxTaps];
  for (int t = 0; t < ntaps; ++t) widx[t] = t;
  WgradDesc d;
  memset(&d, 0, sizeof(d));
  d.Cx = Cx; d.Cy = Cy; d.ntaps = ntaps;
  d.tap_x_off = tap_x_off; d.tap_dy_off = tap_dy_off; d.tap_w = widx;
  d.rows_total = rows_total;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  MPU_TRY(wgrad_plan(p, d));
  out[0] = p.co_tiles * p.ngroups * (p.ci_tiles_full * p.splits + (p.ci_tiles - p.ci_tiles_full) * p.splits_part);
  out[1] = p.splits; out[2] = p.splits_part; out[3] = p.kblocks_per_split; out[4] = p.kblocks_per_split_part;
  out[5] = p.kblocks; out[6] = p.ci_tiles_full; out[7] = p.ci_tiles; out[8] = p.co_tiles; out[9] = p.ngroups;
  out[10] = p.n_entries; out[11] = p.CA;
  for (int i = 0; i < p.n_entries; ++i) out[12 + i] = (int)p.order[i];
  return MPU_OK;
}

}  // extern "C"
