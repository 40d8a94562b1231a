// Multi-tap GEMM on tcgen05 tensor cores — the one contraction every conv of the mpunet 2D U-Net
// reduces to once activations live in zero-bordered ("padded-linear") NHWC bf16 layout:
//
//   forward / dgrad :  D[n, m]   = sum_t sum_c  W[t][n][c] * A[m + a_off[t], c]        (K-major operands;
//                      output channels n on the 128 MMA rows, 256 pixels m on the MMA columns)
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

// Taps whose row offsets differ by < 8 share one shared-memory "slab" of A rows: the slab is loaded once
// and each tap's MMA reads it through a descriptor whose start address is shifted by whole 128 B rows.
constexpr int kMaxGroupTaps = 9;
struct TapGroup {
  int a_off;                     // row offset of the slab relative to the anchor row
  int ntaps;                     // 1..9
  int shift[kMaxGroupTaps];      // extra row shift of each tap inside the slab (0 .. slab_rows - 257)
  int w_idx[kMaxGroupTaps];      // tap index inside the weight matrix
};

struct FwdParams {
  CUtensorMap tmA0_hi, tmA0_lo, tmA1_hi, tmA1_lo;  // activation slabs: boxes of 136 + 128 rows x 64 ch
  CUtensorMap tmA0_ext, tmA1_ext;                  // + ext_rows more rows when the slab is taller than 264 rows
  int slab_rows, ext_rows;       // slab = 256 pixels + the largest tap shift of a group, rounded up to 8 rows:
                                 // 264 rows serve the 3 kx taps of one kernel row; where the padded image rows are
                                 // short (levels >= 2) ONE slab of 264 + 2*(W+2) rows serves all 9 taps of a 3x3 conv
  CUtensorMap tmW;               // weights: box 128 rows (output channels) x 64 K; or, w_mn, 64 K rows x 64 channels
  int w_mn;                      // 1: weights read MN-major from the FORWARD layout [tap][K rows][channels]
                                 //    (dgrad reuses the forward weights: no transposed copy)
  int chunks0, chunks1;          // 64-channel K chunks per activation source (source 1 = concat partner)
  int c0_valid, c1_valid;        // physical channels per source (the last chunk issues fewer K=16 steps)
  int kofs1;                     // weight-matrix K offset of source 1
  int ngroups;
  TapGroup groups[kMaxTaps];
  int w_rows_per_tap;            // rows (output channels, physical) per tap in the weight matrix
  int M_rows;                    // total anchor rows (pixels)
  int n_valid;                   // physical output channels to store
  int m_tiles, n_tiles;          // 256-pixel tiles x 128-channel tiles
  RowMap map;
  __nv_bfloat16* out;            // [out rows][ldo]
  int ldo;
  const float* bias;             // [n_valid] or null
  const __nv_bfloat16* mask;     // same indexing as out; keep value only where mask > 0 (ReLU bwd)
  int ldm;
  int relu;
  double* stats;                 // optional [2][n_valid]: per-channel sum / sum of squares of the stored
                                 // (bf16-rounded) outputs over valid pixels (BatchNorm statistics, bias grads)
  // Column reductions of the STORED (rounded, masked) outputs, taken in the row-wise copy-out phase - what the
  // backward pass otherwise computes with separate passes over the tensor just written:
  float* csum_f;                 // optional [n_valid] fp32 += sum over pixels (a conv's bias gradient = column sum of dz)
  double* red_d;                 // optional [2][red_C] += [sum out | sum out * y] for output channels red_col0 ..
  const __nv_bfloat16* red_y;    //   red_col0 + red_C - 1, y = red_y[out row][ch - red_col0] (BatchNorm backward's
  int red_ldy, red_col0, red_C;  //   sum g and sum g*y when this GEMM produces the gradient g at a BN output)
  int a_slots, w_slots;          // slab ring slots; weight ring stages (two tiles each)
  int stg_px;                    // pixels per epilogue staging pass (64 or 32)
  int stg_ch;                    // channels per staging row: 128, or the tensor's own width for a single narrow tile
  int w_tile_bytes;              // one weight tile in the ring: 16 KB, or (K-major, single narrow tile) n_phys x 128 B
  int items_per_tile;            // (group, chunk, tap) items of one output tile = weight tiles streamed per tile
  long long* dbg;                // optional [8] cycle counters written by CTA 0 (bring-up profiling)
  int dbg_flags;                 // bring-up only (env MPU_FWD_DEBUG): 1 = epilogue drains nothing, 2 = no slab loads,
                                 // 4 = no weight loads (results are garbage; isolates the pipeline stages' cost)
};

// Host-side description of one forward-type GEMM; fwd_setup builds tensor maps + tap groups from it.
struct FwdDesc {
  const void* A0; long long rowsA0; int C0, ldA0;
  const void* A1; long long rowsA1; int C1, ldA1;   // optional concat partner
  const void* W; int w_taps, n_phys, k_total;       // bf16 [w_taps][n_phys][k_total]
  int w_mn, w_rows;                                  // w_mn: W is [w_taps][w_rows = K][n_phys] (forward layout)
  int ntaps; const int* tap_a_off; const int* tap_w;
  int M_rows;
  int BN;                                            // ignored (kept for the bring-up API)
  RowMap map;
  void* out; int ldo;
  const float* bias; const void* mask; int ldm; int relu;
  double* stats;                                     // optional [2][n_phys]
  float* csum_f;                                     // optional column sums (see FwdParams)
  double* red_d; const void* red_y; int red_ldy, red_col0, red_C;
};

// wgrad tap group: taps of one kernel row share one X slab (64+8 pixel rows) and one dY tile
struct WgradGroup {
  int x_off;                     // row offset of the X slab relative to the K block start
  int dy_off;                    // row offset applied to dY (phase plane of the upsample-conv)
  int ntaps;                     // 1..3
  int shift[3];                  // extra row shift of each tap inside the slab (0..7)
  int w_idx[3];                  // tap index inside dW
};

constexpr int kWgMaxEntries = 512;

struct WgradParams {
  CUtensorMap tmX, tmDY;         // X box {64 ch, 72 rows}; dY box {64 ch, 64 rows}
  int ngroups;
  WgradGroup groups[kMaxTaps];
  int CA;                        // 64-channel ci atoms per CTA (1 or 2)
  int ci_atoms;                  // 64-channel atoms of X in total (the last ci tile may hold fewer than CA)
  int ci_tiles, co_tiles, splits;  // ci tiles of CA*64 channels, co tiles of 128 channels; K splits of the full ci tiles
  int ci_tiles_full;             // ci tiles that hold all CA atoms (the partial last one, if any, follows)
  int splits_part;               // K splits of the partial ci tile (fewer: its CTAs do less MMA work per K block)
  int kblocks;                   // total 64-row K blocks
  int kblocks_per_split;
  int kblocks_per_split_part;
  float* dW;                     // [tap][co][ldw] fp32, accumulated with vector reductions
  int ldw;
  int w_rows_per_tap;
  int dw_col0;                   // column offset (concat source 1)
  float* bias_grad;              // optional [co_valid] += column sums of dY over all K rows (the conv's bias gradient)
  int ci_valid, co_valid;
  int stages;
  int BN;                        // unused (kept for the bring-up API)
  long long* dbg;                // optional [8] role timings of CTA 0 (bring-up profiling, see the kernel)
  int n_entries;                 // > 0: launch order table below is used (else full tiles first, partial tile last)
  uint32_t order[kWgMaxEntries]; // per group of ngroups x co_tiles CTAs: bit 31 partial tile, bits 16-30 ci tile, 0-15 split
};

struct WgradDesc {
  const void* X; long long rowsX; int Cx, ldX;
  const void* dY; long long rowsDY; int Cy, ldDY;
  int ntaps; const int* tap_x_off; const int* tap_dy_off; const int* tap_w;
  long long rows_total;          // K extent (anchor rows)
  int BN;                        // 0 = choose
  int splits;                    // 0 = choose
  float* dW; int ldw, w_rows_per_tap, dw_col0;
  float* bias_grad;              // optional: += column sums of dY (see WgradParams)
};

// Host helpers -----------------------------------------------------------------------------------
// 2-D bf16 row-major tensor map with 128B swizzle; box = {box_cols (<=64), box_rows}.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows);

int fwd_setup(FwdParams& p, const FwdDesc& d);
int launch_fwd(FwdParams& p, cudaStream_t stream);
int pick_bn(int n_valid);
int wgrad_setup(WgradParams& p, const WgradDesc& d);
int wgrad_plan(WgradParams& p, const WgradDesc& d);   // the shape-only part of wgrad_setup (p zero-initialised)
int launch_wgrad(WgradParams& p, cudaStream_t stream);
int num_sms();

}  // namespace mpu
