// Multi-tap GEMM on tcgen05 tensor cores — the one contraction every conv of the mpunet 2D U-Net
// reduces to once activations live in zero-bordered ("padded-linear") NHWC bf16 layout:
//
//   forward / dgrad :  D[m, n]   = sum_t sum_c  A[m + a_off[t], c] * W[t][n][c]        (K-major operands)
//   wgrad           :  dW[t][n][c] = sum_m      X[m + x_off[t], c] * dY[m + dy_off, n]  (MN-major operands)
//
// A / X / dY are 2-D row-major matrices [rows][channels]; a tap shift is a pure row offset because
// every image is stored with a 1-pixel zero border, so SAME padding needs no per-tap predicates.
// Replaces (reference): tf.keras Conv2D fwd/bwd inside mpunet/models/unet.py:120-179 (cuDNN in TF).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mpu {

constexpr int kMaxTaps = 16;

// Geometry of an "anchor" row index m -> (image, y, x) in padded space and the output row it maps to.
struct RowMap {
  int Hp, Wp;        // padded dims of the anchor space (H+2, W+2)
  int oHp, oWp;      // padded dims of the output space
  int s;             // output stride (1: same grid, 2: nearest-upsampled grid)
  int py, px;        // output phase: out (y,x) = (s*(ya-1)+py+1, s*(xa-1)+px+1)
};

struct FwdParams {
  CUtensorMap tmA0, tmA1, tmB;
  int chunks0, chunks1;          // 64-channel K chunks per A source (source 1 = concat partner)
  int kofs1;                     // weight-matrix K offset of source 1
  int ntaps;
  int tap_a_off[kMaxTaps];       // row offset added to the anchor row for this tap
  int tap_w[kMaxTaps];           // tap index inside the weight matrix
  int w_rows_per_tap;            // rows (output channels, physical) per tap in the weight matrix
  int M_rows;                    // total anchor rows
  int n_valid;                   // physical output channels to store
  int BN;                        // N tile (multiple of 16, <= 256)
  int m_tiles, n_tiles;
  RowMap map;
  __nv_bfloat16* out;            // [out rows][ldo]
  int ldo;
  const float* bias;             // [n_valid] or null
  const __nv_bfloat16* mask;     // same indexing as out; keep value only where mask > 0 (ReLU bwd)
  int ldm;
  int relu;
  int stages;
};

struct WgradGroup {
  int first, count;              // range in the tap arrays
  int dy_off;                    // row offset applied to dY for this group
};

struct WgradParams {
  CUtensorMap tmX, tmDY;
  int ntaps;
  int tap_x_off[kMaxTaps];
  int tap_w[kMaxTaps];
  int ngroups;
  WgradGroup groups[kMaxTaps];
  int BN;                        // co tile (multiple of 16, <= 256); group count * BN <= 512
  int ci_tiles, co_tiles, splits;
  int kblocks;                   // total 64-row K blocks
  int kblocks_per_split;
  float* dW;                     // [tap][co][ldw] fp32, accumulated with atomics
  int ldw;
  int w_rows_per_tap;
  int dw_col0;                   // column offset (concat source 1)
  int ci_valid, co_valid;
  int stages;
  // descriptor strides exposed for bring-up sweeps
  int a_lbo, a_sbo, b_lbo, b_sbo, kstep_bytes;
};

// Host helpers -----------------------------------------------------------------------------------
// 2-D bf16 row-major tensor map with 128B swizzle; box = {box_cols (<=64), box_rows}.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows);

int launch_fwd(FwdParams& p, cudaStream_t stream);
int launch_wgrad(WgradParams& p, cudaStream_t stream);
int num_sms();

}  // namespace mpu
