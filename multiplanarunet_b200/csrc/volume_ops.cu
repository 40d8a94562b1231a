// Volume-side kernels of the mpunet hot path (CUDA cores, gather-bound; the 256^3 volume sits in L2):
//   * oblique-plane trilinear / nearest sampler  (mpunet/interpolation/sample_grid.py:192-244,
//     regular_grid_interpolator.py:204-223,252-270, view_interpolator.py:62-101,
//     sequences/isotrophic_live_view_sequence_2d.py:103-117, preprocessing/scaling.py:75-88)
//   * multi-view nearest mapping + fusion + argmax (utils/fusion/fuse_and_predict.py:92-137,
//     models/fusion_model.py:38-39, bin/predict.py:349-366, utils/utils.py:311-328)
//   * fusion-layer training step (evaluate/loss_functions.py:207-246, models/fusion_model.py:9-11)
// Index math is float64 in the reference's operation order (explicit _rn intrinsics: no FMA
// contraction) so gathers are bit-exact; see oracle/sampler.py and oracle/fusion.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mpunet_b200.h"
#include "common.h"

namespace mpu {

namespace {

constexpr int kMaxCh = 8;
constexpr int kMaxViews = 16;
constexpr int kMaxClasses = 16;

// i = searchsorted(g, x, 'left') - 1 clamped to [0, n-2]  ==  largest i with g[i] < x (clamped).
template <typename T>
__device__ __forceinline__ int find_cell(const T* __restrict__ g, int n, double x, double inv_step) {
  int i = (int)floor((x - (double)g[0]) * inv_step);
  i = max(0, min(n - 2, i));
  while (i < n - 2 && (double)g[i + 1] < x) ++i;
  while (i > 0 && !((double)g[i] < x)) --i;
  return i;
}

// (x - g0) / (g1 - g0) <= 0.5 evaluated exactly as numpy does (division rounded to nearest, then the
// comparison) but without the division in all but a vanishing sliver of cases: with a = x - g0 and
// b = g1 - g0 > 0, fl(a/b) <= 0.5 holds iff a/b <= 0.5*(1 + 2^-53); 2a <= b decides "yes" and
// 2a > b*(1 + 2^-51) decides "no"; only values in between take the division.
__device__ __forceinline__ bool lower_half(double a, double b) {
  const double d = __dsub_rn(__dadd_rn(a, a), b);
  if (d <= 0.0) return true;
  if (d > b * 4.5e-16) return false;
  return __ddiv_rn(a, b) <= 0.5;
}

__device__ __forceinline__ double dot3(const double* m, double a, double b, double c) {
  // BLAS-style accumulate: ((m0*a) + m1*b) + m2*c with fused multiply-adds
  return fma(m[2], c, fma(m[1], b, __dmul_rn(m[0], a)));
}

struct SamplerParams {
  const float* vol;       // [X][Y][Z][C]
  const uint8_t* labels;  // [X][Y][Z] or null
  int X, Y, Z, C;
  const float *gx, *gy, *gz;  // float32 voxel axes (sample_grid.py:93-98)
  double inv_step[3];
  const double* planes;   // [n][10]: basis row-major 3x3 (columns u v n), offset
  int n_planes, dim;
  double ax_start, ax_step;  // in-plane axis: idx*step + start (np.mgrid semantics)
  int has_rot;
  double rot[9];
  float bg_value[kMaxCh];
  int bg_class;
  int apply_scaler;
  double center[kMaxCh], scale[kMaxCh];
  float* out_f32;          // [n][dim][dim][C] or null
  __nv_bfloat16* out_pad;  // zero-bordered [n][dim+2][dim+2][cpad] or null
  int cpad;
  uint8_t* out_lab;        // [n][dim][dim] or null
};

// real-space coordinates of pixel (i, j) of plane pl (sample_grid.py:227-239 + optional view_interpolator.py:54-60)
__device__ __forceinline__ void plane_point(const SamplerParams& p, int pl, int i, int j, double* q) {
  const double* P = p.planes + (size_t)pl * 10;
  const double a = __dadd_rn(__dmul_rn((double)i, p.ax_step), p.ax_start);
  const double b = __dadd_rn(__dmul_rn((double)j, p.ax_step), p.ax_start);
  const double off = P[9];
  q[0] = dot3(P + 0, a, b, off);
  q[1] = dot3(P + 3, a, b, off);
  q[2] = dot3(P + 6, a, b, off);
  if (p.has_rot) {
    const double r0 = dot3(p.rot + 0, q[0], q[1], q[2]);
    const double r1 = dot3(p.rot + 3, q[0], q[1], q[2]);
    const double r2 = dot3(p.rot + 6, q[0], q[1], q[2]);
    q[0] = r0; q[1] = r1; q[2] = r2;
  }
}

// regular_grid_interpolator.py:252-270 on the three float32 voxel axes: cell index, float64 weight, OOB flag
__device__ __forceinline__ bool locate(const SamplerParams& p, const double* q, int* ci, double* tt) {
  const float* gs[3] = {p.gx, p.gy, p.gz};
  const int ns[3] = {p.X, p.Y, p.Z};
  bool oob = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* g = gs[k];
    const int n = ns[k];
    const int c = find_cell(g, n, q[k], p.inv_step[k]);
    ci[k] = c;
    const float den = __fsub_rn(g[c + 1], g[c]);  // float32 axis difference, as numpy computes it
    tt[k] = __ddiv_rn(__dsub_rn(q[k], (double)g[c]), (double)den);
    oob = oob || q[k] < (double)g[0] || q[k] > (double)g[n - 1];
  }
  return oob;
}

// regular_grid_interpolator.py:204-217: 8 corners in itertools.product order, float64 weights and accumulation
__device__ __forceinline__ double trilinear(const SamplerParams& p, const int* ci, const double* tt, int c) {
  const long long sY = (long long)p.Z * p.C, sX = (long long)p.Y * sY;
  const long long base = (long long)ci[0] * sX + (long long)ci[1] * sY + (long long)ci[2] * p.C;
  double acc = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int dx = (e >> 2) & 1, dy = (e >> 1) & 1, dz = e & 1;
    double w = dx ? tt[0] : __dsub_rn(1.0, tt[0]);
    w = __dmul_rn(w, dy ? tt[1] : __dsub_rn(1.0, tt[1]));
    w = __dmul_rn(w, dz ? tt[2] : __dsub_rn(1.0, tt[2]));
    const float v = __ldg(p.vol + base + dx * sX + dy * sY + dz * p.C + c);
    acc = __dadd_rn(acc, __dmul_rn((double)v, w));
  }
  return acc;
}

// regular_grid_interpolator.py:219-223: per axis t <= .5 -> i else i + 1
__device__ __forceinline__ int nearest_label(const SamplerParams& p, const int* ci, const double* tt) {
  const int s0 = tt[0] <= 0.5 ? ci[0] : ci[0] + 1;
  const int s1 = tt[1] <= 0.5 ? ci[1] : ci[1] + 1;
  const int s2 = tt[2] <= 0.5 ? ci[2] : ci[2] + 1;
  return p.labels[((long long)s0 * p.Y + s1) * p.Z + s2];
}

__global__ void sample_planes_kernel(const SamplerParams p) {
  const long long total = (long long)p.n_planes * p.dim * p.dim;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % p.dim);
    const long long t0 = idx / p.dim;
    const int i = (int)(t0 % p.dim);
    const int pl = (int)(t0 / p.dim);
    double q[3], tt[3];
    int ci[3];
    plane_point(p, pl, i, j, q);
    const bool oob = locate(p, q, ci, tt);
    for (int c = 0; c < p.C; ++c) {
      const double acc = oob ? (double)p.bg_value[c] : trilinear(p, ci, tt, c);
      float val = (float)acc;
      if (p.apply_scaler) {
        const float t = (float)__dsub_rn((double)val, p.center[c]);
        val = (float)__ddiv_rn((double)t, p.scale[c]);
      }
      if (p.out_f32) p.out_f32[idx * p.C + c] = val;
      if (p.out_pad) {
        const long long row = (long long)pl * (p.dim + 2) * (p.dim + 2) + (long long)(i + 1) * (p.dim + 2) + (j + 1);
        p.out_pad[row * p.cpad + c] = __float2bfloat16_rn(val);
      }
    }
    if (p.out_lab) {
      int lab = p.bg_class;
      if (!oob && p.labels) lab = nearest_label(p, ci, tt);
      p.out_lab[idx] = (uint8_t)lab;
    }
  }
}

// Candidate probe of the training-batch rejection sampler: what validate_lab / validate_lab_vec
// (sequences/isotrophic_live_view_sequence.py:98-128) and is_valid_im (:91-96) need to know about a candidate
// plane, without materialising it.  grid = (blocks per plane, planes).
//   class_mask[pl] |= 1u << nearest label of every pixel (out-of-bounds pixels carry bg_class)
//   valid[pl]       = 1 once any pixel of any channel of the UNSCALED trilinear image is not
//                     np.isclose(value, bg_value) (|v - bg| <= 1e-8 + 1e-5 |bg|); threads skip the image work
//                     as soon as the plane's flag is up, so the pass costs little more than the label gather.
__global__ void __launch_bounds__(256) probe_planes_kernel(const SamplerParams p, unsigned int* __restrict__ class_mask,
                                                           unsigned int* __restrict__ valid) {
  const int pl = blockIdx.y;
  __shared__ unsigned int s_mask, s_valid;
  if (threadIdx.x == 0) {
    s_mask = 0u;
    s_valid = 0u;
  }
  __syncthreads();
  unsigned int mask = 0u;
  volatile unsigned int* gvalid = valid + pl;
  const int npix = p.dim * p.dim;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const int i = pix / p.dim, j = pix - i * p.dim;
    double q[3], tt[3];
    int ci[3];
    plane_point(p, pl, i, j, q);
    const bool oob = locate(p, q, ci, tt);
    int lab = p.bg_class;
    if (!oob && p.labels) lab = nearest_label(p, ci, tt);
    mask |= 1u << (lab > 31 ? 31 : lab);
    if (valid && !oob && !s_valid && !*gvalid) {
      bool differs = false;
      for (int c = 0; c < p.C; ++c) {
        const float v = (float)trilinear(p, ci, tt, c);
        const double bg = (double)p.bg_value[c];
        differs = differs || !(fabs((double)v - bg) <= 1e-8 + 1e-5 * fabs(bg));
      }
      if (differs) s_valid = 1u;
    }
  }
  mask = __reduce_or_sync(0xffffffffu, mask);
  if ((threadIdx.x & 31) == 0 && mask) atomicOr(&s_mask, mask);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (class_mask && s_mask) atomicOr(class_mask + pl, s_mask);
    if (valid && s_valid) atomicOr(valid + pl, 1u);
  }
}

// ------------------------------------------------------------------------------------------------
struct MapFuseParams {
  int X, Y, Z, C, V;
  double affine[9];      // voxel -> real (3x3 part of the NIfTI affine)
  double mean[3];        // centre subtracted by get_voxel_grid_real_space (sample_grid.py:117-118)
  const float* pred[kMaxViews];      // per view [n_planes][dim][dim][C] softmax probabilities
  const double* inv_basis;           // [V][9]
  const double* ax;                  // [dim] in-plane axis (float64 linspace)
  const double* offsets;             // [V][n_planes]
  int dim, n_planes;
  double inv_step_ax;
  double inv_step_off[kMaxViews];
  const float* W;                    // [V][C] fusion weights (null for sum fusion)
  const float* b;                    // [C]
  int sum_fusion;
  uint8_t* labels_out;               // [X][Y][Z] or null
  float* probs_out;                  // [X][Y][Z][C] or null
  float* combined_out;               // [V][X][Y][Z][C] or null
  // Analytic axes (fast path): every axis is an np.linspace, g[i] = fl(fl(i * step) + start) for i < n-1 and
  // g[n-1] = stop (numpy/_core/function_base.py); the host verifies the tables bit for bit before enabling it.
  int analytic;
  double ax_start, ax_step, ax_stop;
  double off_start[kMaxViews], off_step[kMaxViews], off_stop[kMaxViews];
  double ib[kMaxViews][9];           // host copy of inv_basis for the fast path
  long long pred_elems;              // floats per view stack (guards the 16-byte over-read at the very end)
};

// Persistent blocks walk (i, j) lines of the voxel grid, threads walk k: no per-voxel integer division.  The
// axis tables every lookup reads (in-plane axis, per-view plane offsets) and the fusion weights are staged in
// shared memory once per block when they fit (SMEM), else read through L1.  NC > 0: class count known at
// compile time (loops fully unrolled without predication); NC == 0: generic, up to kMaxClasses.
template <int NC, bool SMEM>
__global__ void __launch_bounds__(256) map_fuse_kernel(const MapFuseParams p) {
  constexpr int MC = NC > 0 ? NC : kMaxClasses;
  const int C = NC > 0 ? NC : p.C;
  extern __shared__ double mf_sm[];
  const double* ax = p.ax;
  const double* offs = p.offsets;
  const double* ibv = p.inv_basis;
  if (SMEM) {
    double* s_ax = mf_sm;
    double* s_off = s_ax + p.dim;
    double* s_ib = s_off + p.V * p.n_planes;
    for (int t = threadIdx.x; t < p.dim; t += blockDim.x) s_ax[t] = p.ax[t];
    for (int t = threadIdx.x; t < p.V * p.n_planes; t += blockDim.x) s_off[t] = p.offsets[t];
    for (int t = threadIdx.x; t < p.V * 9; t += blockDim.x) s_ib[t] = p.inv_basis[t];
    __syncthreads();
    ax = s_ax;
    offs = s_off;
    ibv = s_ib;
  }
  const long long plane = (long long)p.Y * p.Z;
  const long long total = (long long)p.X * plane;
  const int lines = p.X * p.Y;
  for (int line = blockIdx.x; line < lines; line += gridDim.x) {
    const int i = line / p.Y, j = line - i * p.Y;
    for (int k = threadIdx.x; k < p.Z; k += blockDim.x) {
      const long long idx = (long long)line * p.Z + k;
      double xr[3];
      xr[0] = __dsub_rn(dot3(p.affine + 0, (double)i, (double)j, (double)k), p.mean[0]);
      xr[1] = __dsub_rn(dot3(p.affine + 3, (double)i, (double)j, (double)k), p.mean[1]);
      xr[2] = __dsub_rn(dot3(p.affine + 6, (double)i, (double)j, (double)k), p.mean[2]);
      float z[MC];
#pragma unroll
      for (int c = 0; c < MC; ++c) z[c] = 0.f;
      for (int v = 0; v < p.V; ++v) {
        const double* ib = ibv + v * 9;
        double q[3];
        q[0] = dot3(ib + 0, xr[0], xr[1], xr[2]);
        q[1] = dot3(ib + 3, xr[0], xr[1], xr[2]);
        q[2] = dot3(ib + 6, xr[0], xr[1], xr[2]);
        const double* off = offs + (size_t)v * p.n_planes;
        int sel[3];
        bool oob = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double* g = a < 2 ? ax : off;
          const int n = a < 2 ? p.dim : p.n_planes;
          const double inv = a < 2 ? p.inv_step_ax : p.inv_step_off[v];
          const int c = find_cell(g, n, q[a], inv);
          sel[a] = lower_half(__dsub_rn(q[a], g[c]), __dsub_rn(g[c + 1], g[c])) ? c : c + 1;
          oob = oob || q[a] < g[0] || q[a] > g[n - 1];
        }
        const float* src = p.pred[v] + (((long long)sel[2] * p.dim + sel[0]) * p.dim + sel[1]) * C;
#pragma unroll
        for (int c = 0; c < MC; ++c) {
          if (NC > 0 || c < C) {
            const float x = oob ? (c == 0 ? 1.f : 0.f) : __ldg(src + c);
            if (p.combined_out) p.combined_out[((long long)v * total + idx) * C + c] = x;
            if (p.sum_fusion) z[c] = __fadd_rn(z[c], x);
            else z[c] = __fadd_rn(z[c], __fmul_rn(__ldg(p.W + v * C + c), x));
          }
        }
      }
      float pr[MC];
      if (p.sum_fusion) {
#pragma unroll
        for (int c = 0; c < MC; ++c) pr[c] = z[c];
      } else {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (NC > 0 || c < C) {
            z[c] = __fadd_rn(z[c], __ldg(p.b + c));
            m = fmaxf(m, z[c]);
          }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (NC > 0 || c < C) {
            pr[c] = expf(__fsub_rn(z[c], m));
            sum = __fadd_rn(sum, pr[c]);
          }
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (NC > 0 || c < C) pr[c] = __fdiv_rn(pr[c], sum);
      }
      int best = 0;
      float bv = pr[0];
#pragma unroll
      for (int c = 1; c < MC; ++c)
        if ((NC > 0 || c < C) && pr[c] > bv) {
          bv = pr[c];
          best = c;
        }
      if (p.labels_out) p.labels_out[idx] = (uint8_t)best;
      if (p.probs_out) {
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (NC > 0 || c < C) p.probs_out[idx * C + c] = pr[c];
      }
    }
  }
}

// ---- fast path: analytic axes, no table lookups, 16-byte gathers -------------------------------------------
// The table version above is bound by the L1/TEX pipe (97 % busy, profiles/r02_ncu_full_volume_kernels_summary.txt):
// per voxel and view it issues ~24 shared-memory table reads for the three index searches and NC scattered 4-byte
// loads.  Here the axis values are recomputed in registers exactly as numpy's linspace produced them, and the NC
// probabilities of a pixel (a 4-byte-aligned run of NC floats) are fetched with ceil((NC+3)/4) aligned 16-byte loads.
struct LinAxis {
  double start, step, stop, inv;
  int n;
  __device__ __forceinline__ double at(int i) const {
    return i == n - 1 ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
  }
};

// searchsorted(g, x, 'left') - 1 clamped to [0, n-2], the nearest-neighbour decision t <= .5 and the OOB test of
// regular_grid_interpolator.py:219-223,252-270 on an analytic axis.  Returns the selected index; *oob accumulates.
// Which cell x falls in only matters through the midpoint rule (a point on a node selects that node from either
// side), so the continuous index r = (x - start) / step decides everything unless its fractional part is within 1e-6
// of .5 - six orders of magnitude above every rounding difference between r and the reference's t (|r| < 1e3,
// double precision) - and only that sliver takes the reference's exact sequence of operations.
__device__ __noinline__ int nearest_on_axis_exact(double start, double step, double stop, double inv, int n, double x) {
  // scalars by value: a struct passed by reference would force the caller to keep it in local memory
  const LinAxis a = {start, step, stop, inv, n};
  int i = (int)floor((x - a.start) * a.inv);
  i = max(0, min(a.n - 2, i));
  double g0 = a.at(i), g1 = a.at(i + 1);
  while (i < a.n - 2 && g1 < x) {
    ++i;
    g0 = g1;
    g1 = a.at(i + 1);
  }
  while (i > 0 && !(g0 < x)) {
    --i;
    g1 = g0;
    g0 = a.at(i);
  }
  return lower_half(__dsub_rn(x, g0), __dsub_rn(g1, g0)) ? i : i + 1;
}

__device__ __forceinline__ int nearest_on_axis(const LinAxis& a, double x, bool* oob) {
  *oob = *oob || x < a.start || x > a.stop;
  const double r = (x - a.start) * a.inv;
  const double fl = floor(r);
  const double f = r - fl;
  if (fabs(f - 0.5) > 1e-6) {
    const int s = (int)fl + (f > 0.5 ? 1 : 0);
    return max(0, min(a.n - 1, s));
  }
  return nearest_on_axis_exact(a.start, a.step, a.stop, a.inv, a.n, x);
}

template <int NC, bool VEC>
__global__ void __launch_bounds__(256, 4) map_fuse_lin_kernel(const __grid_constant__ MapFuseParams p) {
  constexpr int NW = (NC + 3 + 3) / 4;  // aligned 16-byte words covering NC floats at any 4-byte phase
  const long long plane = (long long)p.Y * p.Z;
  const long long total = (long long)p.X * plane;
  const int lines = p.X * p.Y;
  __shared__ float s_w[kMaxViews * NC + NC];  // fusion weights and bias: broadcast reads
  if (!p.sum_fusion) {
    for (int t = threadIdx.x; t < p.V * NC; t += blockDim.x) s_w[t] = p.W[t];
    for (int t = threadIdx.x; t < NC; t += blockDim.x) s_w[kMaxViews * NC + t] = p.b[t];
  }
  __syncthreads();
  // lines are dealt round-robin: all resident blocks then work inside the same ~2-voxel-thick slab of the volume,
  // whose pixels of every view stay in L2 (measured: giving each block a run of consecutive lines instead spreads
  // the blocks over the whole volume and multiplies the DRAM traffic by 2.5)
  for (int line = blockIdx.x; line < lines; line += gridDim.x) {
    const int i = line / p.Y, j = line - i * p.Y;
    for (int k = threadIdx.x; k < p.Z; k += blockDim.x) {
      const long long idx = (long long)line * p.Z + k;
      double xr[3];
      xr[0] = __dsub_rn(dot3(p.affine + 0, (double)i, (double)j, (double)k), p.mean[0]);
      xr[1] = __dsub_rn(dot3(p.affine + 3, (double)i, (double)j, (double)k), p.mean[1]);
      xr[2] = __dsub_rn(dot3(p.affine + 6, (double)i, (double)j, (double)k), p.mean[2]);
      float z[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) z[c] = 0.f;
      for (int v = 0; v < p.V; ++v) {
        const double* ib = p.ib[v];
        const double q0 = dot3(ib + 0, xr[0], xr[1], xr[2]);
        const double q1 = dot3(ib + 3, xr[0], xr[1], xr[2]);
        const double q2 = dot3(ib + 6, xr[0], xr[1], xr[2]);
        const LinAxis ax = {p.ax_start, p.ax_step, p.ax_stop, p.inv_step_ax, p.dim};
        const LinAxis of = {p.off_start[v], p.off_step[v], p.off_stop[v], p.inv_step_off[v], p.n_planes};
        bool oob = false;
        const int s0 = nearest_on_axis(ax, q0, &oob);
        const int s1 = nearest_on_axis(ax, q1, &oob);
        const int s2 = nearest_on_axis(of, q2, &oob);
        float x[NC];
        if (oob) {
#pragma unroll
          for (int c = 0; c < NC; ++c) x[c] = c == 0 ? 1.f : 0.f;
        } else {
          const long long e = (((long long)s2 * p.dim + s0) * p.dim + s1) * NC;  // first float of the pixel
          const float* src = p.pred[v] + e;
          const int ph = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);     // 4-byte phase inside 16 bytes
          if (VEC && e - ph + 4 * NW <= p.pred_elems) {
            const uint4* a16 = reinterpret_cast<const uint4*>(src - ph);
            uint32_t w[4 * NW];
#pragma unroll
            for (int t = 0; t < NW; ++t) {
              const uint4 q = __ldg(a16 + t);
              w[4 * t] = q.x; w[4 * t + 1] = q.y; w[4 * t + 2] = q.z; w[4 * t + 3] = q.w;
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const uint32_t u = ph == 0 ? w[c] : (ph == 1 ? w[c + 1] : (ph == 2 ? w[c + 2] : w[c + 3]));
              x[c] = __uint_as_float(u);
            }
          } else {  // last pixels of the stack: stay inside the buffer
#pragma unroll
            for (int c = 0; c < NC; ++c) x[c] = __ldg(src + c);
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (p.combined_out) p.combined_out[((long long)v * total + idx) * NC + c] = x[c];
          if (p.sum_fusion) z[c] = __fadd_rn(z[c], x[c]);
          else z[c] = __fadd_rn(z[c], __fmul_rn(s_w[v * NC + c], x[c]));
        }
      }
      float pr[NC];
      if (p.sum_fusion) {
#pragma unroll
        for (int c = 0; c < NC; ++c) pr[c] = z[c];
      } else {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          z[c] = __fadd_rn(z[c], s_w[kMaxViews * NC + c]);
          m = fmaxf(m, z[c]);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          pr[c] = expf(__fsub_rn(z[c], m));
          sum = __fadd_rn(sum, pr[c]);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) pr[c] = __fdiv_rn(pr[c], sum);
      }
      int best = 0;
      float bv = pr[0];
#pragma unroll
      for (int c = 1; c < NC; ++c)
        if (pr[c] > bv) {
          bv = pr[c];
          best = c;
        }
      if (p.labels_out) p.labels_out[idx] = (uint8_t)best;
      if (p.probs_out) {
#pragma unroll
        for (int c = 0; c < NC; ++c) p.probs_out[idx * NC + c] = pr[c];
      }
    }
  }
}

template <int NC>
static int launch_map_fuse_lin(const MapFuseParams& p, cudaStream_t st) {
  int occ = 1;
  const int sms = sm_count();
  static int vec = -1;  // bring-up knob: MPU_MAPFUSE_VEC=0 gathers with 4-byte loads
  if (vec < 0) {
    const char* e = getenv("MPU_MAPFUSE_VEC");
    vec = e ? atoi(e) : 1;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, map_fuse_lin_kernel<NC, true>, 256, 0) != cudaSuccess) occ = 1;
  const int lines = p.X * p.Y;
  const int grid = std::max(1, std::min(lines, sms * std::max(occ, 1)));
  if (vec) map_fuse_lin_kernel<NC, true><<<grid, 256, 0, st>>>(p);
  else map_fuse_lin_kernel<NC, false><<<grid, 256, 0, st>>>(p);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

template <int NC>
int launch_map_fuse(const MapFuseParams& p, cudaStream_t st) {
  if constexpr (NC > 0) {
    if (p.analytic) return launch_map_fuse_lin<NC>(p, st);
  }
  const size_t smem = sizeof(double) * ((size_t)p.dim + (size_t)p.V * p.n_planes + (size_t)p.V * 9);
  const int lines = p.X * p.Y;
  int occ = 1;
  if (smem <= 40 * 1024) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, map_fuse_kernel<NC, true>, 256, smem) != cudaSuccess)
      occ = 1;
    const int grid = std::max(1, std::min(lines, sm_count() * std::max(occ, 1)));
    map_fuse_kernel<NC, true><<<grid, 256, smem, st>>>(p);
  } else {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, map_fuse_kernel<NC, false>, 256, 0) != cudaSuccess)
      occ = 1;
    const int grid = std::max(1, std::min(lines, sm_count() * std::max(occ, 1)));
    map_fuse_kernel<NC, false><<<grid, 256, 0, st>>>(p);
  }
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

// ------------------------------------------------------------------------------------------------
// Fusion layer training: per point softmax(sum_v W*x + b), generalized dice (uniform weights),
// analytic gradients; accum = [dW (V*C) | db (C) | loss] in double via block reduction + atomics.
__global__ void fusion_grad_kernel(const float* __restrict__ X, const uint8_t* __restrict__ y,
                                   long long n, int V, int C, const float* __restrict__ W,
                                   const float* __restrict__ b, double* __restrict__ accum) {
  extern __shared__ double fsm[];  // [warps][V*C + C + 1]
  const int nacc = V * C + C + 1;
  float loc[kMaxViews * kMaxClasses / 2 + kMaxClasses + 1];  // supports V*C <= 128
  for (int k = 0; k < nacc; ++k) loc[k] = 0.f;
  float w[kMaxViews * kMaxClasses / 2];
  for (int k = 0; k < V * C; ++k) w[k] = W[k];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float* x = X + i * V * C;
    float z[kMaxClasses], pr[kMaxClasses];
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += w[v * C + c] * x[v * C + c];
      z[c] = s + b[c];
    }
    float m = z[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) {
      pr[c] = expf(z[c] - m);
      se += pr[c];
    }
    const float inv = 1.f / se;
    const int lab = y[i];
    float dice_sum = 0.f, dp[kMaxClasses], dpp = 0.f;
    for (int c = 0; c < C; ++c) {
      pr[c] *= inv;
      const float o = c == lab ? 1.f : 0.f;
      const float num = 2.f * o * pr[c];
      const float den = pr[c] + o + 1e-6f;
      dice_sum += num / den;
      dp[c] = -((2.f * o * den - num) / (den * den)) / (float)C;
      dpp += dp[c] * pr[c];
    }
    loc[nacc - 1] += 1.f - dice_sum / (float)C;
    for (int c = 0; c < C; ++c) {
      const float dz = pr[c] * (dp[c] - dpp);
      loc[V * C + c] += dz;
      for (int v = 0; v < V; ++v) loc[v * C + c] += dz * x[v * C + c];
    }
  }
  // warp reduce then block reduce in double
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = 0; k < nacc; ++k) {
    double v = (double)loc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) fsm[warp * nacc + k] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nacc; k += blockDim.x) {
    double s = 0;
    for (int ww = 0; ww < nwarps; ++ww) s += fsm[ww * nacc + k];
    atomicAdd(accum + k, s);
  }
}

// Compile-time (V, C) variant: everything lives in registers (the generic kernel above indexes its accumulators
// dynamically, which puts them in local memory: 687 MB of DRAM writes per 256^3 pass, profiles/r02_ncu_full_volume_
// kernels_summary.txt).  One point = V*C consecutive floats, read with 8-byte loads (V*C even) - each thread walks
// its own row, the warp covers 32 consecutive rows, so every fetched line is fully used out of L1.
// One Adam step on the fusion parameters from the accumulated gradient sums (shared by the stand-alone Adam kernel
// and the fused single-process train step).  k indexes [W (V*C) | b (C)].
__device__ __forceinline__ void fusion_adam_update(int k, float* W, float* b, float* m, float* v, double grad_sum,
                                                   double n_points, int nW, int C, float reg, float lr_t, float b1,
                                                   float b2, float eps) {
  float* prm = k < nW ? W + k : b + (k - nW);
  const float cnt = k < nW ? (float)nW : (float)C;
  const float g = (float)(grad_sum / n_points) + reg * 2.f * (*prm) / cnt;
  const float mi = b1 * m[k] + (1.f - b1) * g;
  const float vi = b2 * v[k] + (1.f - b2) * g * g;
  m[k] = mi;
  v[k] = vi;
  *prm = *prm - lr_t * mi / (sqrtf(vi) + eps);
}

constexpr int kMaxPeers = 8;
constexpr int kMailDoubles = 40;  // per (parity, rank) slot: V*C + C + 1 sums (<= 38 used by the instantiations) + count
// Mailbox every rank owns in peer-mapped (symmetric) memory: data[2][kMaxPeers][kMailDoubles] doubles, then
// flag[2][kMaxPeers] unsigned 64-bit sequence numbers.
constexpr size_t kMailFlagOffset = sizeof(double) * 2 * kMaxPeers * kMailDoubles;
constexpr int kFusionMaxBlocks = 2048;  // per-block partial-sum slots behind the arrival counter (see mpu_fusion_scratch_bytes)

struct FusionStepArgs {  // optional fused Adam: the last block to finish applies the update
  float *Wp, *bp, *m, *v;
  unsigned int* counter;   // zero before the launch; reset by the kernel
  double* loss_out;        // receives the batch's mean loss (without regulariser), or null
  float reg, lr_t, b1, b2, eps;
  int fused;
  double* scratch;         // [gridDim.x][kMailDoubles] per-block partial sums (fused mode): no atomics - 592 blocks
                           // adding to the same 36 addresses cost ~20 us per batch - and a fixed summation order
  // multi-rank exchange fused into the same kernel (world > 1): every rank's last block stores its sums into every
  // peer's mailbox over NVLink, raises the peer's flag to `seq`, waits for all flags of its own mailbox, adds the
  // contributions in rank order (identical bits on every rank) and applies Adam.  No NCCL call, no host round trip.
  int world, rank;
  unsigned long long seq;  // strictly increasing per step, same on all ranks
  unsigned char* mail[kMaxPeers];
};

// One warp-chunk of 32 points of a fusion-training batch: gather / stream the rows into the warp's staging buffer,
// then one point per lane: softmax(sum_v W.x + b), generalised-dice loss and its gradient sums into loc[].
template <int V, int C, bool IDX>
__device__ __forceinline__ void fusion_chunk(const float* __restrict__ X, const uint8_t* __restrict__ y,
                                             const long long* __restrict__ index, long long p0, long long n,
                                             float* __restrict__ st, int lane, const float (&w)[V * C],
                                             const float (&bb)[C], float (&loc)[V * C + C + 1]) {
  constexpr int VC = V * C, NACC = VC + C + 1;
  constexpr int kChunkFloats = 32 * VC;
  constexpr int kVec = kChunkFloats / 4;
    __syncwarp();
    if (IDX) {
      // rows are 8-byte aligned when V*C is even: gather float2 pieces, consecutive lanes inside one row
      const long long mine = p0 + lane < n ? index[p0 + lane] : 0;
      if (VC % 2 == 0) {
        constexpr int kPieces = VC / 2;
        // all kPieces gathers of a lane are issued before the first one is consumed (a rolled loop serialises the
        // DRAM latency: 15 x ~0.8 us per chunk, which WAS the 37 us per 2^17-point batch)
        constexpr int kHalf = (kPieces + 1) / 2;
#pragma unroll
        for (int h0 = 0; h0 < kPieces; h0 += kHalf) {
          float2 t[kHalf];
#pragma unroll
          for (int j = 0; j < kHalf; ++j) {
            const int e = (h0 + j) * 32 + lane;
            const int row = e / kPieces, col = e - row * kPieces;
            const long long src_row = __shfl_sync(0xffffffffu, mine, row);
            t[j] = (h0 + j < kPieces && p0 + row < n)
                       ? __ldg(reinterpret_cast<const float2*>(X + src_row * VC) + col) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < kHalf; ++j)
            if (h0 + j < kPieces) reinterpret_cast<float2*>(st)[(h0 + j) * 32 + lane] = t[j];
        }
      } else {
        for (int e = lane; e < kChunkFloats; e += 32) {
          const int row = e / VC, col = e - row * VC;
          const long long src_row = __shfl_sync(0xffffffffu, mine, row);
          st[e] = p0 + row < n ? __ldg(X + src_row * VC + col) : 0.f;
        }
      }
    } else {
      const long long f0 = p0 * VC;               // first float of the chunk: 16-byte aligned (32 * VC * 4 bytes)
      const long long fend = n * VC;
      if (f0 + kChunkFloats <= fend) {
        const float4* src = reinterpret_cast<const float4*>(X + f0);
#pragma unroll
        for (int j = 0; j < (kVec + 31) / 32; ++j) {
          const int v4 = j * 32 + lane;
          if (v4 < kVec) reinterpret_cast<float4*>(st)[v4] = __ldg(src + v4);
        }
      } else {
        for (int k = lane; k < kChunkFloats; k += 32) st[k] = f0 + k < fend ? __ldg(X + f0 + k) : 0.f;
      }
    }
    __syncwarp();
    const long long i = p0 + lane;
    if (i < n) {
      float x[VC];
#pragma unroll
      for (int k = 0; k < VC; ++k) x[k] = st[lane * VC + k];
      const int lab = y[IDX ? index[i] : i];
      float z[C], pr[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) s += w[v * C + c] * x[v * C + c];
        z[c] = s + bb[c];
      }
      float m = z[0];
#pragma unroll
      for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        pr[c] = __expf(z[c] - m);
        se += pr[c];
      }
      const float inv = __fdividef(1.f, se);
      float pl = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        pr[c] *= inv;
        pl = c == lab ? pr[c] : pl;
      }
      // dice_c = 2 o_c p_c / (p_c + o_c + eps) vanishes (with its derivative) for every class but the label's:
      // loss = 1 - dice_l / C,  d loss / d p_l = -2 (1 + eps) / (p_l + 1 + eps)^2 / C,
      // dz_c = p_c (dp_c - sum_k dp_k p_k) = dp_l p_c ([c == l] - p_l)
      const float den = pl + 1.f + 1e-6f;
      const float rden = __fdividef(1.f, den);
      loc[NACC - 1] += 1.f - 2.f * pl * rden * (1.f / (float)C);
      const float dpl = -2.f * (1.f + 1e-6f) * rden * rden * (1.f / (float)C);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float dz = dpl * pr[c] * ((c == lab ? 1.f : 0.f) - pl);
        loc[VC + c] += dz;
#pragma unroll
        for (int v = 0; v < V; ++v) loc[v * C + c] += dz * x[v * C + c];
      }
    }
  }

template <int V, int C, bool IDX>
__global__ void __launch_bounds__(128, 4) fusion_grad_kernel_t(const float* __restrict__ X, const uint8_t* __restrict__ y,
                                                               const long long* __restrict__ index, long long n,
                                                               const float* __restrict__ W, const float* __restrict__ b,
                                                               double* __restrict__ accum, const FusionStepArgs fs) {
  constexpr int VC = V * C, NACC = VC + C + 1;
  constexpr int kChunkFloats = 32 * VC;          // one warp iteration = 32 points
  constexpr int kVec = kChunkFloats / 4;         // float4 per chunk (32 * VC is a multiple of 4)
  __shared__ double fsm[4][NACC];
  // Each warp stages its 32 points (32 * VC floats) through shared memory with coalesced loads: per-thread row walks
  // touch 32 different lines per load instruction and saturate the L1 tag stage (measured 0.83 ms per 256^3 pass
  // against 0.31 ms of HBM time).  With `index` (a shuffled epoch: point i is row index[i], the reference's
  // fit(shuffle=True)) the rows are gathered 8 bytes per lane, 15 lanes per row for V*C = 30.
  __shared__ __align__(16) float stage[4][kChunkFloats];
  __shared__ int s_last;
  float loc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) loc[k] = 0.f;
  float w[VC], bb[C];
#pragma unroll
  for (int k = 0; k < VC; ++k) w[k] = W[k];
#pragma unroll
  for (int c = 0; c < C; ++c) bb[c] = b[c];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nchunks = (n + 31) / 32;
  const long long wstride = (long long)gridDim.x * 4;
  float* st = stage[warp];
  for (long long ch = (long long)blockIdx.x * 4 + warp; ch < nchunks; ch += wstride) {
    const long long p0 = ch * 32;
    fusion_chunk<V, C, IDX>(X, y, index, p0, n, st, lane, w, bb, loc);
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    double v = (double)loc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) fsm[warp][k] = v;
  }
  __syncthreads();
  const bool to_scratch = fs.fused && fs.scratch != nullptr;
  for (int k = threadIdx.x; k < NACC; k += blockDim.x) {
    double s = 0;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) s += fsm[ww][k];
    if (to_scratch) fs.scratch[(size_t)blockIdx.x * kMailDoubles + k] = s;
    else atomicAdd(accum + k, s);
  }
  if (!fs.fused) return;
  // fused Adam: the block that arrives last sees every block's contribution
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fs.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (to_scratch) {
    // deterministic tree: thread t sums blocks t, t + 128, ... of every accumulator, then the warps and the block
    // combine through shared memory in a fixed order; the totals land in accum
    const volatile double* sc = fs.scratch;
    for (int k = 0; k < NACC; ++k) {
      double part = 0.0;
      for (unsigned int bq = threadIdx.x; bq < gridDim.x; bq += blockDim.x) part += sc[(size_t)bq * kMailDoubles + k];
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) fsm[warp][k] = part;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NACC; k += blockDim.x) accum[k] = (fsm[0][k] + fsm[1][k]) + (fsm[2][k] + fsm[3][k]);
    __threadfence();
    __syncthreads();
  }
  volatile double* acc_v = accum;
  if (fs.world <= 1) {
    if (threadIdx.x < VC + C)
      fusion_adam_update(threadIdx.x, fs.Wp, fs.bp, fs.m, fs.v, acc_v[threadIdx.x], (double)n, VC, C, fs.reg, fs.lr_t,
                         fs.b1, fs.b2, fs.eps);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (fs.loss_out) *fs.loss_out = acc_v[NACC - 1] / (double)n;
      *fs.counter = 0u;
    }
  } else {
    // ---- exchange over peer memory ----
    const int par = (int)(fs.seq & 1ull);
    const int slot = (par * kMaxPeers + fs.rank) * kMailDoubles;
    for (int t = threadIdx.x; t < fs.world * (NACC + 1); t += blockDim.x) {
      const int peer = t / (NACC + 1), k = t - peer * (NACC + 1);
      volatile double* dst = reinterpret_cast<volatile double*>(fs.mail[peer]) + slot;
      dst[k] = k < NACC ? acc_v[k] : (double)n;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < fs.world) {
      volatile unsigned long long* flag =
          reinterpret_cast<volatile unsigned long long*>(fs.mail[threadIdx.x] + kMailFlagOffset) + par * kMaxPeers + fs.rank;
      *flag = fs.seq;
      __threadfence_system();
      // wait for rank `threadIdx.x`'s contribution to MY mailbox
      volatile unsigned long long* mine =
          reinterpret_cast<volatile unsigned long long*>(fs.mail[fs.rank] + kMailFlagOffset) + par * kMaxPeers + threadIdx.x;
      const long long t0 = clock64();
      while (*mine != fs.seq) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: a missing peer must not hang the GPU
          printf("mpu: fusion peer exchange timeout (rank %d waiting for rank %d, seq %llu)\n", fs.rank, (int)threadIdx.x,
                 fs.seq);
          __trap();
        }
      }
      __threadfence_system();
    }
    __syncthreads();
    volatile double* box = reinterpret_cast<volatile double*>(fs.mail[fs.rank]) + par * kMaxPeers * kMailDoubles;
    double ntot = 0.0;
    for (int q = 0; q < fs.world; ++q) ntot += box[q * kMailDoubles + NACC];
    if (threadIdx.x < VC + C && ntot > 0.0) {  // a step in which no rank had points left changes nothing
      double gsum = 0.0;
      for (int q = 0; q < fs.world; ++q) gsum += box[q * kMailDoubles + threadIdx.x];  // rank order: same bits everywhere
      fusion_adam_update(threadIdx.x, fs.Wp, fs.bp, fs.m, fs.v, gsum, ntot, VC, C, fs.reg, fs.lr_t, fs.b1, fs.b2, fs.eps);
    }
    if (threadIdx.x == 0) {
      double lsum = 0.0;
      for (int q = 0; q < fs.world; ++q) lsum += box[q * kMailDoubles + NACC - 1];
      if (fs.loss_out) *fs.loss_out = ntot > 0.0 ? lsum / ntot : 0.0;
      *fs.counter = 0u;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NACC; k += blockDim.x) accum[k] = 0.0;  // ready for the next batch
}

// ---- persistent epoch kernel -----------------------------------------------------------------------------------
// FusionModel.fit's epoch (bin/train_fusion.py:196-213) is a chain of dependent Adam steps over 2^17-point batches
// (128 per 256^3 volume): one launch per batch costs ~35 us of launch + tail + last-block latency for ~3 us of HBM
// time.  This kernel is launched ONCE per epoch (cooperatively: every block resident) and walks all batches:
//   per batch: every block adds its gradient sums into accum[k % 3] (fp64 atomics) -> grid barrier (one arrival per
//   block on a monotonic counter) -> every block reads the 36 totals and applies the SAME Adam update to its own copy
//   of (W, b, m, v) in shared memory - no broadcast of the new weights, no second barrier.  Block 0 writes the batch
//   loss, clears the accumulator two batches ahead and stores the parameters at the end.
// Multi-rank (world > 1): after the grid barrier block 0 exchanges the totals with every peer's mailbox over NVLink
// (same protocol as the per-batch kernel), adds them in rank order and publishes them in tot[k % 3] behind a flag
// the other blocks wait on.
struct FusionEpochArgs {
  const float* X;
  const uint8_t* y;
  const long long* perm;      // null: rows in order
  long long n, batch, n_batches;
  float *W, *b, *m, *v;       // parameters and Adam moments (read at start, written by block 0 at the end)
  double* acc3;               // [3][kMailDoubles], zero at launch
  double* tot3;               // [3][kMailDoubles] (multi-rank totals)
  unsigned long long* bar;    // [0] arrivals (monotonic), [1] blocks past the last barrier, [2] published-totals flag
  double* losses_out;         // [n_batches] or null
  float reg, lr, b1, b2, eps;
  int first_step;
  int world, rank;
  unsigned long long first_seq;
  unsigned char* mail[kMaxPeers];
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int V, int C, bool IDX>
__global__ void __launch_bounds__(128, 3) fusion_epoch_kernel_t(const FusionEpochArgs a) {
  constexpr int VC = V * C, NP = VC + C, NACC = VC + C + 1;
  constexpr int kChunkFloats = 32 * VC;
  __shared__ double fsm[4][NACC];
  __shared__ __align__(16) float stage[4][kChunkFloats];
  __shared__ float sP[NP], sM[NP], sV[NP];   // this block's copy of the parameters and Adam moments
  __shared__ double sTot[NACC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NP) {
    sP[threadIdx.x] = threadIdx.x < VC ? a.W[threadIdx.x] : a.b[threadIdx.x - VC];
    sM[threadIdx.x] = a.m[threadIdx.x];
    sV[threadIdx.x] = a.v[threadIdx.x];
  }
  __syncthreads();
  float* st = stage[warp];
  const long long wstride = (long long)gridDim.x * 4;
  // beta^step of Keras' bias correction, carried from step to step (pow() per batch would sit on the critical path)
  double p1 = pow((double)a.b1, (double)(a.first_step - 1)), p2 = pow((double)a.b2, (double)(a.first_step - 1));
  for (long long k = 0; k < a.n_batches; ++k) {
    p1 *= (double)a.b1;
    p2 *= (double)a.b2;
    const long long s0 = k * a.batch;
    const long long nb = s0 < a.n ? (a.n - s0 < a.batch ? a.n - s0 : a.batch) : 0;  // (a rank may run out of points)
    double* acc = a.acc3 + (k % 3) * kMailDoubles;
    float w[VC], bb[C], loc[NACC];
#pragma unroll
    for (int j = 0; j < VC; ++j) w[j] = sP[j];
#pragma unroll
    for (int c = 0; c < C; ++c) bb[c] = sP[VC + c];
#pragma unroll
    for (int j = 0; j < NACC; ++j) loc[j] = 0.f;
    const long long nchunks = (nb + 31) / 32;
    const float* Xb = IDX ? a.X : a.X + s0 * VC;
    const uint8_t* yb = IDX ? a.y : a.y + s0;
    const long long* ib = IDX ? a.perm + s0 : nullptr;
    for (long long ch = (long long)blockIdx.x * 4 + warp; ch < nchunks; ch += wstride)
      fusion_chunk<V, C, IDX>(Xb, yb, ib, ch * 32, nb, st, lane, w, bb, loc);
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      double v = (double)loc[j];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) fsm[warp][j] = v;
    }
    __syncthreads();
    if (nchunks > (long long)blockIdx.x * 4)  // blocks without a chunk of this batch add nothing
      for (int j = threadIdx.x; j < NACC; j += blockDim.x)
        atomicAdd(acc + j, (fsm[0][j] + fsm[1][j]) + (fsm[2][j] + fsm[3][j]));
    // ---- grid barrier k ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(a.bar, 1ull);
      const unsigned long long target = (unsigned long long)(k + 1) * gridDim.x;
      const long long t0 = clock64();
      while (ld_acquire_u64(a.bar) < target) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: never hang the GPU
          printf("mpu: fusion epoch barrier timeout (block %d, batch %lld)\n", (int)blockIdx.x, k);
          __trap();
        }
      }
    }
    __syncthreads();
    const double* src = acc;
    double ntot = (double)nb;
    if (a.world > 1) {
      // block 0: exchange with the peers, publish the rank-ordered totals; everybody else waits for the flag
      double* tot = a.tot3 + (k % 3) * kMailDoubles;
      const unsigned long long seq = a.first_seq + (unsigned long long)k;
      if (blockIdx.x == 0) {
        const int par = (int)(seq & 1ull);
        const int slot = (par * kMaxPeers + a.rank) * kMailDoubles;
        for (int t = threadIdx.x; t < a.world * (NACC + 1); t += blockDim.x) {
          const int peer = t / (NACC + 1), j = t - peer * (NACC + 1);
          volatile double* dst = reinterpret_cast<volatile double*>(a.mail[peer]) + slot;
          dst[j] = j < NACC ? __ldcg(acc + j) : (double)nb;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < a.world) {
          volatile unsigned long long* flag =
              reinterpret_cast<volatile unsigned long long*>(a.mail[threadIdx.x] + kMailFlagOffset) + par * kMaxPeers + a.rank;
          *flag = seq;
          __threadfence_system();
          volatile unsigned long long* mine =
              reinterpret_cast<volatile unsigned long long*>(a.mail[a.rank] + kMailFlagOffset) + par * kMaxPeers + threadIdx.x;
          const long long t0 = clock64();
          while (*mine != seq) {
            if (clock64() - t0 > 4000000000ll) {
              printf("mpu: fusion peer exchange timeout (rank %d waiting for rank %d, seq %llu)\n", a.rank,
                     (int)threadIdx.x, seq);
              __trap();
            }
          }
          __threadfence_system();
        }
        __syncthreads();
        volatile double* box = reinterpret_cast<volatile double*>(a.mail[a.rank]) + par * kMaxPeers * kMailDoubles;
        for (int j = threadIdx.x; j <= NACC; j += blockDim.x) {
          double sum = 0.0;
          for (int q = 0; q < a.world; ++q) sum += box[q * kMailDoubles + j];  // rank order: same bits on every rank
          tot[j] = sum;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(a.bar + 2, (unsigned long long)(k + 1));
      }
      if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_u64(a.bar + 2) < (unsigned long long)(k + 1)) {
          if (clock64() - t0 > 4000000000ll) {
            printf("mpu: fusion epoch totals timeout (block %d, batch %lld)\n", (int)blockIdx.x, k);
            __trap();
          }
        }
      }
      __syncthreads();
      src = tot;
    }
    for (int j = threadIdx.x; j <= NACC; j += blockDim.x)
      sTot[j] = (j < NACC || a.world > 1) ? __ldcg(src + j) : 0.0;
    __syncthreads();
    if (a.world > 1) ntot = sTot[NACC];
    if (threadIdx.x < NP && ntot > 0.0) {  // a step in which no rank had points left changes nothing
      const float lr_t = (float)((double)a.lr * sqrt(1.0 - p2) / (1.0 - p1));
      const int j = threadIdx.x;
      const float cnt = j < VC ? (float)VC : (float)C;
      const float g = (float)(sTot[j] / ntot) + a.reg * 2.f * sP[j] / cnt;
      const float mi = a.b1 * sM[j] + (1.f - a.b1) * g;
      const float vi = a.b2 * sV[j] + (1.f - a.b2) * g * g;
      sM[j] = mi;
      sV[j] = vi;
      sP[j] = sP[j] - lr_t * mi / (sqrtf(vi) + a.eps);
    }
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0 && a.losses_out) a.losses_out[k] = ntot > 0.0 ? sTot[NACC - 1] / ntot : 0.0;
      // accumulator of batch k + 2 (== the one of batch k - 1, which every block has read before arriving at
      // barrier k) is cleared now; its next atomics come after barrier k + 1, which needs this block's arrival
      double* nxt = a.acc3 + ((k + 2) % 3) * kMailDoubles;
      for (int j = threadIdx.x; j < kMailDoubles; j += blockDim.x) nxt[j] = 0.0;
    }
    __syncthreads();
  }
  // ---- epilogue: parameters out, barrier state back to zero for the next launch ----
  if (threadIdx.x == 0) atomicAdd(a.bar + 1, 1ull);
  if (blockIdx.x == 0) {
    if (threadIdx.x < NP) {
      if (threadIdx.x < VC) a.W[threadIdx.x] = sP[threadIdx.x];
      else a.b[threadIdx.x - VC] = sP[threadIdx.x];
      a.m[threadIdx.x] = sM[threadIdx.x];
      a.v[threadIdx.x] = sV[threadIdx.x];
    }
    if (threadIdx.x == 0) {  // every block has left the loop: nobody reads the barrier words or accumulators any more
      const long long t0 = clock64();
      while (ld_acquire_u64(a.bar + 1) < gridDim.x)
        if (clock64() - t0 > 4000000000ll) break;
      a.bar[0] = 0ull;
      a.bar[1] = 0ull;
      a.bar[2] = 0ull;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 3 * kMailDoubles; j += blockDim.x) a.acc3[j] = 0.0;
  }
}

template <int V, int C>
static int launch_fusion_epoch_t(const FusionEpochArgs& a, cudaStream_t st) {
  int occ = 0;
  const bool idx = a.perm != nullptr;
  const void* fn = idx ? (const void*)fusion_epoch_kernel_t<V, C, true> : (const void*)fusion_epoch_kernel_t<V, C, false>;
  cudaError_t e = idx ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fusion_epoch_kernel_t<V, C, true>, 128, 0)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fusion_epoch_kernel_t<V, C, false>, 128, 0);
  if (e != cudaSuccess || occ < 1) {
    set_error("fusion epoch kernel: occupancy query failed (%s)", cudaGetErrorString(e));
    return MPU_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(a.X) & 15) != 0) {
    set_error("mpu_fusion_train_epoch: X must be 16-byte aligned");
    return MPU_ERR_ARG;
  }
  // one 32-point chunk per warp and batch is the finest useful split; every block must be resident (spin barrier)
  long long blocks = (a.batch + 127) / 128;
  long long cap = (long long)sm_count() * occ;
  static int env_cap = -1;
  if (env_cap < 0) {
    const char* ev = getenv("MPU_FUSION_EPOCH_BLOCKS");
    env_cap = ev ? atoi(ev) : 0;
  }
  if (env_cap > 0 && env_cap < cap) cap = env_cap;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  FusionEpochArgs args = a;
  void* params[] = {&args};
  MPU_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(128), params, 0, st));
  count_launch();
  return MPU_OK;
}

static int fusion_epoch_dispatch(const FusionEpochArgs& a, int V, int C, cudaStream_t st, bool* handled) {
  *handled = true;
  if (V == 6) {
    switch (C) {
      case 2: return launch_fusion_epoch_t<6, 2>(a, st);
      case 3: return launch_fusion_epoch_t<6, 3>(a, st);
      case 4: return launch_fusion_epoch_t<6, 4>(a, st);
      case 5: return launch_fusion_epoch_t<6, 5>(a, st);
      case 6: return launch_fusion_epoch_t<6, 6>(a, st);
      case 7: return launch_fusion_epoch_t<6, 7>(a, st);
      case 8: return launch_fusion_epoch_t<6, 8>(a, st);
      default: break;
    }
  }
  if (V == 3 && C == 5) return launch_fusion_epoch_t<3, 5>(a, st);
  *handled = false;
  return MPU_OK;
}

template <int V, int C>
static int launch_fusion_grad_t(const float* X, const unsigned char* y, const long long* index, long long n,
                                const float* W, const float* b, double* accum, const FusionStepArgs& fs,
                                cudaStream_t st) {
  int occ = 2;
  const int sms = sm_count();
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fusion_grad_kernel_t<V, C, true>, 128, 0) != cudaSuccess ||
      occ < 1)
    occ = 1;
  if ((reinterpret_cast<uintptr_t>(X) & 15) != 0) {
    set_error("mpu_fusion_grad: X must be 16-byte aligned");
    return MPU_ERR_ARG;
  }
  long long blocks = (n + 127) / 128;
  long long cap = (long long)sms * occ;
  static int env_cap = -1;  // bring-up knob
  if (env_cap < 0) {
    const char* e = getenv("MPU_FUSION_BLOCKS");
    env_cap = e ? atoi(e) : 0;
  }
  if (env_cap > 0) cap = env_cap;
  if (cap > kFusionMaxBlocks) cap = kFusionMaxBlocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (index) fusion_grad_kernel_t<V, C, true><<<(int)blocks, 128, 0, st>>>(X, y, index, n, W, b, accum, fs);
  else fusion_grad_kernel_t<V, C, false><<<(int)blocks, 128, 0, st>>>(X, y, index, n, W, b, accum, fs);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

// Adam on the 35 fusion parameters (+ L2 regulariser gradients); one thread per parameter.
__global__ void fusion_adam_kernel(float* W, float* b, float* m, float* v, const double* accum,
                                   double n_points, int V, int C, float reg, float lr_t, float b1,
                                   float b2, float eps) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int nW = V * C;
  if (k >= nW + C) return;
  // n_points <= 0: the (all-reduced) point count travels with the sums, in accum[V*C + C + 1]
  const double np_ = n_points > 0 ? n_points : accum[nW + C + 1];
  fusion_adam_update(k, W, b, m, v, accum[k], np_, nW, C, reg, lr_t, b1, b2, eps);
}

inline int grid_for(long long work, int threads, int cap = 0) {
  if (cap <= 0) cap = sm_count() * 16;
  long long g = (work + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}


// ---- interpolation at explicit coordinates ------------------------------------------------------------------
// ViewInterpolator.__call__ / intrp_image / intrp_labels on an arbitrary grid (view_interpolator.py:62-101): the
// same per-point arithmetic as sample_planes_kernel (index search on the float32 voxel axes, float64 trilinear
// weights in itertools.product order, nearest labels, whole-point out-of-bounds fill), with the real-space
// coordinates supplied by the caller as [3][n] float64 instead of derived from a plane description.
struct PointsParams {
  const float* vol;
  const uint8_t* labels;
  int X, Y, Z, C;
  const float *gx, *gy, *gz;
  double inv_step[3];
  const double* coords;  // [3][n]
  long long n;
  float bg_value[kMaxCh];
  int bg_class;
  float* out_f32;    // [n][C] or null
  uint8_t* out_lab;  // [n] or null
};

__global__ void sample_points_kernel(const PointsParams p) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.n;
       idx += (long long)gridDim.x * blockDim.x) {
    const double q[3] = {p.coords[idx], p.coords[p.n + idx], p.coords[2 * p.n + idx]};
    const float* gs[3] = {p.gx, p.gy, p.gz};
    const int ns[3] = {p.X, p.Y, p.Z};
    int ci[3];
    double tt[3];
    bool oob = false;
    for (int k = 0; k < 3; ++k) {
      const float* g = gs[k];
      const int n = ns[k];
      const int c = find_cell(g, n, q[k], p.inv_step[k]);
      ci[k] = c;
      const float den = __fsub_rn(g[c + 1], g[c]);
      tt[k] = __ddiv_rn(__dsub_rn(q[k], (double)g[c]), (double)den);
      oob = oob || q[k] < (double)g[0] || q[k] > (double)g[n - 1];
    }
    const long long sY = (long long)p.Z * p.C, sX = (long long)p.Y * sY;
    const long long base = (long long)ci[0] * sX + (long long)ci[1] * sY + (long long)ci[2] * p.C;
    if (p.out_f32) {
      for (int c = 0; c < p.C; ++c) {
        double acc = 0.0;
        if (!oob) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int dx = (e >> 2) & 1, dy = (e >> 1) & 1, dz = e & 1;
            double w = dx ? tt[0] : __dsub_rn(1.0, tt[0]);
            w = __dmul_rn(w, dy ? tt[1] : __dsub_rn(1.0, tt[1]));
            w = __dmul_rn(w, dz ? tt[2] : __dsub_rn(1.0, tt[2]));
            const float v = __ldg(p.vol + base + dx * sX + dy * sY + dz * p.C + c);
            acc = __dadd_rn(acc, __dmul_rn((double)v, w));
          }
        } else {
          acc = (double)p.bg_value[c];
        }
        p.out_f32[idx * p.C + c] = (float)acc;
      }
    }
    if (p.out_lab) {
      int lab = p.bg_class;
      if (!oob && p.labels) {
        const int s0 = tt[0] <= 0.5 ? ci[0] : ci[0] + 1;
        const int s1 = tt[1] <= 0.5 ? ci[1] : ci[1] + 1;
        const int s2 = tt[2] <= 0.5 ? ci[2] : ci[2] + 1;
        lab = p.labels[((long long)s0 * p.Y + s1) * p.Z + s2];
      }
      p.out_lab[idx] = (uint8_t)lab;
    }
  }
}

// ---- exact centre of the voxel grid ------------------------------------------------------------------------------
// get_voxel_grid_real_space subtracts np.mean(A . (i, j, k)) over ALL voxels (sample_grid.py:117-118).  numpy sums each
// coordinate with its pairwise scheme: blocks of <= 128 elements are summed with 8 interleaved accumulators
// (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum) and the block sums are combined by a binary tree whose
// split points the host derives from the element count.  This kernel produces the block ("leaf") sums in exactly that
// operation order, from elements evaluated exactly as the mapping kernel evaluates them (BLAS-style fma chain);
// the host combines the tree (multiplanarunet_b200/interpolation/voxel_center.py).  One thread per (leaf, coordinate).
__global__ void voxel_leaf_sums_kernel(int Y, int Z, double a0, double a1, double a2, double b0, double b1, double b2,
                                       double c0, double c1, double c2, const long long* __restrict__ leaf_start,
                                       const int* __restrict__ leaf_len, long long n_leaves, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_leaves) return;
  const int r = (int)(t / n_leaves);
  const long long leaf = t - (long long)r * n_leaves;
  const double m[3] = {r == 0 ? a0 : (r == 1 ? b0 : c0), r == 0 ? a1 : (r == 1 ? b1 : c1),
                       r == 0 ? a2 : (r == 1 ? b2 : c2)};
  const long long s = leaf_start[leaf];
  const int n = leaf_len[leaf];
  const long long yz = (long long)Y * Z;
  auto elem = [&](long long idx) {
    const long long i = idx / yz;
    const long long rem = idx - i * yz;
    const long long j = rem / Z, k = rem - j * Z;
    return dot3(m, (double)i, (double)j, (double)k);
  };
  double res;
  if (n < 8) {
    res = 0.0;
    for (int q = 0; q < n; ++q) res = __dadd_rn(res, elem(s + q));
  } else {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = elem(s + q);
    int q = 8;
    for (; q < n - (n % 8); q += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = __dadd_rn(acc[u], elem(s + q + u));
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                    __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
    for (; q < n; ++q) res = __dadd_rn(res, elem(s + q));
  }
  out[t] = res;
}

// ---- confusion-matrix counts ---------------------------------------------------------------------------
// counts[0][c] = #(true == c & pred == c), counts[1][c] = #(true == c), counts[2][c] = #(pred == c)
// (TP / relevant / selected of callbacks/validation.py:117-131; the three sums of evaluate/metrics.py:12-23).
// pred comes either as labels (u8) or as per-class scores [n][ncls] f32 whose first maximum is the label
// (numpy argmax).  Integer work: exact.
__global__ void label_counts_kernel(const unsigned char* __restrict__ yt, const unsigned char* __restrict__ yp,
                                    const float* __restrict__ scores, long long n, int ncls,
                                    unsigned long long* __restrict__ counts) {
  __shared__ unsigned int h[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = yt[i];
    int p;
    if (scores) {
      const float* s = scores + i * ncls;
      float best = s[0];
      p = 0;
      for (int c = 1; c < ncls; ++c) {
        const float v = s[c];
        if (v > best) {
          best = v;
          p = c;
        }
      }
    } else {
      p = yp[i];
    }
    if (t < ncls) {
      atomicAdd(&h[256 + t], 1u);
      if (t == p) atomicAdd(&h[t], 1u);
    }
    if (p < ncls) atomicAdd(&h[512 + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * ncls; i += blockDim.x) {
    const unsigned int v = h[(i / ncls) * 256 + i % ncls];
    if (v) atomicAdd(counts + i, (unsigned long long)v);
  }
}

// ---- Elastic2D augmentation ------------------------------------------------------------------------------
// (mpunet/augmentation/elastic_deformation.py:6-69.)  The displacement fields are scipy's
// gaussian_filter(noise, sigma, mode="constant") - two separable float64 passes, axis 0 then axis 1 - whose
// inner loop (scipy/ndimage/src/ni_filters.c, NI_Correlate1D, symmetric branch) is
//     out = in[p]*w[0];  for j = -r .. -1:  out += (in[p+j] + in[p-j]) * w[j]
// evaluated here in the same order with explicitly rounded double operations (no FMA contraction).
__global__ void gauss1d_kernel(const double* __restrict__ in, double* __restrict__ out, int H, int W, int axis,
                               const double* __restrict__ weights, int wstride, const int* __restrict__ radius) {
  const int field = blockIdx.y;          // slice * 2 + {dx, dy}
  const int slice = field >> 1;
  const int r = radius[slice];
  const double* fw = weights + (long long)slice * wstride + r;  // centre tap
  const double* src = in + (long long)field * H * W;
  double* dst = out + (long long)field * H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const int y = p / W, x = p - y * W;
    const int pos = axis == 0 ? y : x, len = axis == 0 ? H : W, step = axis == 0 ? W : 1;
    double acc = __dmul_rn(src[p], fw[0]);
    for (int j = -r; j < 0; ++j) {
      const int a = pos + j, b = pos - j;
      const double va = a >= 0 ? src[p + j * step] : 0.0;
      const double vb = b < len ? src[p - j * step] : 0.0;
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(va, vb), fw[j]));
    }
    dst[p] = acc;
  }
}

// one axis of the reference's RegularGridInterpolator on the integer grid arange(n)
// (regular_grid_interpolator.py:252-270): i = searchsorted(grid, x) - 1 clamped to [0, n-2], t = x - i
__device__ __forceinline__ void find_index_arange(double x, int n, int* i, double* t, bool* oob) {
  int k = (int)ceil(fmin(fmax(x, -1.0e9), 1.0e9)) - 1;
  k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
  *i = k;
  *t = __dsub_rn(x, (double)k);
  *oob = (x < 0.0) || (x > (double)(n - 1));
}

__global__ void elastic_resample_kernel(const float* __restrict__ xin, const unsigned char* __restrict__ yin,
                                        const double* __restrict__ fields, const double* __restrict__ alpha,
                                        const float* __restrict__ bg, int H, int W, int C,
                                        float* __restrict__ xout, unsigned char* __restrict__ yout) {
  const int slice = blockIdx.y;
  const double al = alpha[slice];
  const double* dxf = fields + (long long)(2 * slice) * H * W;
  const double* dyf = dxf + (long long)H * W;
  const float* xi = xin + (long long)slice * H * W * C;
  float* xo = xout + (long long)slice * H * W * C;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const int r = p / W, c = p - r * W;
    const double px = __dadd_rn((double)r, __dmul_rn(dxf[p], al));
    const double py = __dadd_rn((double)c, __dmul_rn(dyf[p], al));
    int i0, i1;
    double t0, t1;
    bool o0, o1;
    find_index_arange(px, H, &i0, &t0, &o0);
    find_index_arange(py, W, &i1, &t1, &o1);
    const bool oob = o0 || o1;
    // corner weights in itertools.product order, weight = (1. * w_axis0) * w_axis1
    const double w0[2] = {__dsub_rn(1.0, t0), t0};
    const double w1[2] = {__dsub_rn(1.0, t1), t1};
    for (int ch = 0; ch < C; ++ch) {
      float outv;
      if (oob) {
        outv = bg[slice * C + ch];
      } else {
        double acc = 0.0;
#pragma unroll
        for (int e0 = 0; e0 < 2; ++e0)
#pragma unroll
          for (int e1 = 0; e1 < 2; ++e1) {
            const double w = __dmul_rn(w0[e0], w1[e1]);
            const double v = (double)xi[((long long)(i0 + e0) * W + (i1 + e1)) * C + ch];
            acc = __dadd_rn(acc, __dmul_rn(v, w));
          }
        outv = __double2float_rn(acc);
      }
      xo[(long long)p * C + ch] = outv;
    }
    if (yin) {
      unsigned char lab = 0;
      if (!oob) {
        const int j0 = t0 <= 0.5 ? i0 : i0 + 1;
        const int j1 = t1 <= 0.5 ? i1 : i1 + 1;
        lab = yin[(long long)slice * H * W + (long long)j0 * W + j1];
      }
      yout[(long long)slice * H * W + p] = lab;
    }
  }
}
}  // namespace
}  // namespace mpu

using namespace mpu;

extern "C" {

static int fill_sampler_params(SamplerParams& p, const char* who, const float* vol, const unsigned char* labels,
                               const int* h_dims, int C, const float* gx, const float* gy, const float* gz,
                               const double* h_inv_step, const double* h_rot, const double* planes, int n_planes,
                               int dim, double span, const float* h_bg_value, int bg_class) {
  if (!vol || !h_dims || !gx || !gy || !gz || !h_inv_step || !planes || n_planes < 1 || dim < 2) {
    set_error("%s: bad arguments", who);
    return MPU_ERR_ARG;
  }
  if (C < 1 || C > kMaxCh) {
    set_error("%s: %d channels unsupported (max %d)", who, C, kMaxCh);
    return MPU_ERR_ARG;
  }
  memset(&p, 0, sizeof(p));
  p.vol = vol;
  p.labels = labels;
  p.X = h_dims[0];
  p.Y = h_dims[1];
  p.Z = h_dims[2];
  p.C = C;
  p.gx = gx;
  p.gy = gy;
  p.gz = gz;
  p.planes = planes;
  p.n_planes = n_planes;
  p.dim = dim;
  // np.mgrid[-hd:hd:dim*1j] with hd = span // 2 (sample_grid.py:227-233)
  const double hd = floor(span / 2.0);
  p.ax_start = -hd;
  p.ax_step = (hd - (-hd)) / (double)(dim - 1);
  p.has_rot = h_rot != nullptr;
  if (h_rot) memcpy(p.rot, h_rot, sizeof(double) * 9);
  for (int c = 0; c < C; ++c) p.bg_value[c] = h_bg_value ? h_bg_value[c] : 0.f;
  p.bg_class = bg_class;
  // 1/spacing of each voxel axis: only seeds the cell search, the axis tables decide the result
  for (int k = 0; k < 3; ++k) p.inv_step[k] = h_inv_step[k];
  return MPU_OK;
}

int mpu_sample_planes(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                      const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                      const double* h_rot,
                      const double* planes, int n_planes, int dim, double span,
                      const float* h_bg_value, int bg_class, const double* h_center,
                      const double* h_scale, float* out_f32, void* out_padded_bf16, int cpad,
                      unsigned char* out_labels, void* stream) {
  SamplerParams p;
  MPU_TRY(fill_sampler_params(p, "mpu_sample_planes", vol, labels, h_dims, C, gx, gy, gz, h_inv_step, h_rot, planes,
                              n_planes, dim, span, h_bg_value, bg_class));
  if (out_padded_bf16 && cpad < C) {
    set_error("mpu_sample_planes: padded channel count %d < C=%d", cpad, C);
    return MPU_ERR_ARG;
  }
  p.apply_scaler = (h_center && h_scale) ? 1 : 0;
  for (int c = 0; c < C; ++c) {
    p.center[c] = h_center ? h_center[c] : 0.0;
    p.scale[c] = h_scale ? h_scale[c] : 1.0;
  }
  p.out_f32 = out_f32;
  p.out_pad = reinterpret_cast<__nv_bfloat16*>(out_padded_bf16);
  p.cpad = cpad;
  p.out_lab = out_labels;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)n_planes * dim * dim;
  sample_planes_kernel<<<grid_for(total, 256), 256, 0, st>>>(p);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_probe_planes(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                     const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                     const double* h_rot, const double* planes, int n_planes, int dim, double span,
                     const float* h_bg_value, int bg_class, unsigned int* class_mask, unsigned int* valid,
                     void* stream) {
  SamplerParams p;
  MPU_TRY(fill_sampler_params(p, "mpu_probe_planes", vol, labels, h_dims, C, gx, gy, gz, h_inv_step, h_rot, planes,
                              n_planes, dim, span, h_bg_value, bg_class));
  if (!class_mask && !valid) {
    set_error("mpu_probe_planes: no output requested");
    return MPU_ERR_ARG;
  }
  if (n_planes > 65535) {
    set_error("mpu_probe_planes: at most 65535 candidate planes per call");
    return MPU_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (class_mask) MPU_CUDA(cudaMemsetAsync(class_mask, 0, sizeof(unsigned int) * n_planes, st));
  if (valid) MPU_CUDA(cudaMemsetAsync(valid, 0, sizeof(unsigned int) * n_planes, st));
  int bx = (dim * dim + 255) / 256;
  if (bx > 16) bx = 16;
  probe_planes_kernel<<<dim3(bx, n_planes), 256, 0, st>>>(p, class_mask, valid);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

static int map_fuse_common(MapFuseParams& p, const char* who, const void* const* h_pred_ptrs, int V, int C, int dim,
                           int n_planes, const int* h_dims, const double* h_affine3x3, const double* h_mean,
                           const float* W, const float* b, int sum_fusion, unsigned char* labels_out,
                           float* probs_out, float* combined_out) {
  if (!h_pred_ptrs || V < 1 || V > kMaxViews || C < 1 || C > kMaxClasses || !h_dims || !h_affine3x3 || !h_mean) {
    set_error("%s: bad arguments (V=%d C=%d)", who, V, C);
    return MPU_ERR_ARG;
  }
  if (!sum_fusion && (!W || !b)) {
    set_error("%s: fusion weights missing", who);
    return MPU_ERR_ARG;
  }
  memset(&p, 0, sizeof(p));
  p.X = h_dims[0];
  p.Y = h_dims[1];
  p.Z = h_dims[2];
  p.C = C;
  p.V = V;
  memcpy(p.affine, h_affine3x3, sizeof(double) * 9);
  memcpy(p.mean, h_mean, sizeof(double) * 3);
  for (int v = 0; v < V; ++v) p.pred[v] = reinterpret_cast<const float*>(h_pred_ptrs[v]);
  p.dim = dim;
  p.n_planes = n_planes;
  p.W = W;
  p.b = b;
  p.sum_fusion = sum_fusion;
  p.labels_out = labels_out;
  p.probs_out = probs_out;
  p.combined_out = combined_out;
  if ((long long)p.X * p.Y > 0x7fffffffLL) {
    set_error("%s: volume too large", who);
    return MPU_ERR_ARG;
  }
  return MPU_OK;
}

static int map_fuse_dispatch(const MapFuseParams& p, cudaStream_t st) {
  switch (p.C) {
    case 2: return launch_map_fuse<2>(p, st);
    case 3: return launch_map_fuse<3>(p, st);
    case 4: return launch_map_fuse<4>(p, st);
    case 5: return launch_map_fuse<5>(p, st);
    case 6: return launch_map_fuse<6>(p, st);
    case 7: return launch_map_fuse<7>(p, st);
    case 8: return launch_map_fuse<8>(p, st);
    default: return launch_map_fuse<0>(p, st);
  }
}

int mpu_map_fuse(const void* const* h_pred_ptrs, int V, int C, int dim, int n_planes,
                 const double* inv_basis, const double* ax, const double* offsets,
                 const double* h_inv_step, const int* h_dims, const double* h_affine3x3,
                 const double* h_mean, const float* W, const float* b, int sum_fusion,
                 unsigned char* labels_out, float* probs_out, float* combined_out, void* stream) {
  if (!inv_basis || !ax || !offsets || !h_inv_step) {
    set_error("mpu_map_fuse: bad arguments (null axis tables)");
    return MPU_ERR_ARG;
  }
  MapFuseParams p;
  MPU_TRY(map_fuse_common(p, "mpu_map_fuse", h_pred_ptrs, V, C, dim, n_planes, h_dims, h_affine3x3, h_mean, W, b,
                          sum_fusion, labels_out, probs_out, combined_out));
  p.inv_basis = inv_basis;
  p.ax = ax;
  p.offsets = offsets;
  // h_inv_step = [1/spacing of the in-plane axis, 1/spacing of each view's offset axis]
  p.inv_step_ax = h_inv_step[0];
  for (int v = 0; v < V; ++v) p.inv_step_off[v] = h_inv_step[1 + v];
  return map_fuse_dispatch(p, reinterpret_cast<cudaStream_t>(stream));
}

int mpu_map_fuse_linspace(const void* const* h_pred_ptrs, int V, int C, int dim, int n_planes,
                          const double* h_inv_basis, const double* h_ax_lin, const double* h_off_lin,
                          const int* h_dims, const double* h_affine3x3, const double* h_mean, const float* W,
                          const float* b, int sum_fusion, unsigned char* labels_out, float* probs_out,
                          float* combined_out, void* stream) {
  if (!h_inv_basis || !h_ax_lin || !h_off_lin || dim < 2 || n_planes < 2) {
    set_error("mpu_map_fuse_linspace: bad arguments (null axis descriptions or an axis shorter than 2)");
    return MPU_ERR_ARG;
  }
  if (C > 8) {
    set_error("mpu_map_fuse_linspace: at most 8 classes on the analytic path (use mpu_map_fuse)");
    return MPU_ERR_ARG;
  }
  MapFuseParams p;
  MPU_TRY(map_fuse_common(p, "mpu_map_fuse_linspace", h_pred_ptrs, V, C, dim, n_planes, h_dims, h_affine3x3, h_mean,
                          W, b, sum_fusion, labels_out, probs_out, combined_out));
  p.analytic = 1;
  p.ax_start = h_ax_lin[0];
  p.ax_step = h_ax_lin[1];
  p.ax_stop = h_ax_lin[2];
  p.inv_step_ax = (double)(dim - 1) / (p.ax_stop - p.ax_start);
  for (int v = 0; v < V; ++v) {
    p.off_start[v] = h_off_lin[3 * v];
    p.off_step[v] = h_off_lin[3 * v + 1];
    p.off_stop[v] = h_off_lin[3 * v + 2];
    p.inv_step_off[v] = (double)(n_planes - 1) / (p.off_stop[v] - p.off_start[v]);
    memcpy(p.ib[v], h_inv_basis + 9 * v, sizeof(double) * 9);
  }
  p.pred_elems = (long long)n_planes * dim * dim * C;
  if (C == 1) {  // a single class has nothing to gather in vectors; the table kernel's template covers 2..8
    set_error("mpu_map_fuse_linspace: needs at least 2 classes");
    return MPU_ERR_ARG;
  }
  return map_fuse_dispatch(p, reinterpret_cast<cudaStream_t>(stream));
}

static int fusion_grad_dispatch(const float* X, const unsigned char* y, const long long* index, long long n, int V, int C,
                                const float* W, const float* b, double* accum, const FusionStepArgs& fs,
                                cudaStream_t st, bool* handled) {
  *handled = true;
  // the reference's configurations (6 views; bin/train_fusion.py) with 2..8 classes run fully in registers
  if (V == 6) {
    switch (C) {
      case 2: return launch_fusion_grad_t<6, 2>(X, y, index, n, W, b, accum, fs, st);
      case 3: return launch_fusion_grad_t<6, 3>(X, y, index, n, W, b, accum, fs, st);
      case 4: return launch_fusion_grad_t<6, 4>(X, y, index, n, W, b, accum, fs, st);
      case 5: return launch_fusion_grad_t<6, 5>(X, y, index, n, W, b, accum, fs, st);
      case 6: return launch_fusion_grad_t<6, 6>(X, y, index, n, W, b, accum, fs, st);
      case 7: return launch_fusion_grad_t<6, 7>(X, y, index, n, W, b, accum, fs, st);
      case 8: return launch_fusion_grad_t<6, 8>(X, y, index, n, W, b, accum, fs, st);
      default: break;
    }
  }
  if (V == 3 && C == 5) return launch_fusion_grad_t<3, 5>(X, y, index, n, W, b, accum, fs, st);
  *handled = false;
  return MPU_OK;
}

int mpu_fusion_grad(const float* X, const unsigned char* y, long long n, int V, int C, const float* W,
                    const float* b, double* accum, void* stream) {
  return mpu_fusion_grad_indexed(X, y, nullptr, n, V, C, W, b, accum, stream);
}

int mpu_fusion_grad_indexed(const float* X, const unsigned char* y, const long long* index, long long n, int V, int C,
                            const float* W, const float* b, double* accum, void* stream) {
  if (n <= 0) return MPU_OK;  // an empty shard contributes nothing (its rank still joins the all-reduce)
  if (!X || !y || !W || !b || !accum || V < 1 || C < 1 || C > kMaxClasses || V * C > 128) {
    set_error("mpu_fusion_grad: bad arguments (V=%d C=%d)", V, C);
    return MPU_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FusionStepArgs fs;
  memset(&fs, 0, sizeof(fs));
  bool handled = false;
  MPU_TRY(fusion_grad_dispatch(X, y, index, n, V, C, W, b, accum, fs, st, &handled));
  if (handled) return MPU_OK;
  if (index) {
    set_error("mpu_fusion_grad_indexed: (V=%d, C=%d) has no gather variant (supported: V=6 with 2..8 classes)", V, C);
    return MPU_ERR_ARG;
  }
  const int threads = 256;
  const int nacc = V * C + C + 1;
  const size_t smem = sizeof(double) * (threads / 32) * nacc;
  fusion_grad_kernel<<<grid_for(n, threads, sm_count() * 4), threads, smem, st>>>(X, y, n, V, C, W, b, accum);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_fusion_train_step(const float* X, const unsigned char* y, const long long* index, long long n, int V, int C,
                          float* W, float* b, float* m, float* v, double* accum, unsigned int* counter,
                          double* loss_out, float reg, float lr, float beta1, float beta2, float eps, int step,
                          void* stream) {
  if (!X || !y || !W || !b || !m || !v || !accum || !counter || n <= 0 || step < 1) {
    set_error("mpu_fusion_train_step: bad arguments");
    return MPU_ERR_ARG;
  }
  FusionStepArgs fs;
  memset(&fs, 0, sizeof(fs));
  fs.Wp = W; fs.bp = b; fs.m = m; fs.v = v;
  fs.counter = counter;
  // (the per-block partial-sum scratch behind the counter is allocated but not used: a single last block summing
  //  592 x 37 partials measured 7.7 ms per epoch against 4.7 ms with fp64 atomics)
  fs.scratch = nullptr;
  fs.loss_out = loss_out;
  fs.reg = reg;
  fs.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
  fs.b1 = beta1; fs.b2 = beta2; fs.eps = eps;
  fs.fused = 1;
  bool handled = false;
  MPU_TRY(fusion_grad_dispatch(X, y, index, n, V, C, W, b, accum, fs, reinterpret_cast<cudaStream_t>(stream), &handled));
  if (!handled) {
    set_error("mpu_fusion_train_step: (V=%d, C=%d) not instantiated (supported: V=6 with 2..8 classes, V=3 C=5)", V, C);
    return MPU_ERR_ARG;
  }
  return MPU_OK;
}

// layout of the caller's `counter` buffer (mpu_fusion_scratch_bytes() bytes, zeroed once): 16-byte arrival counter of
// the per-batch kernel, then the persistent epoch kernel's barrier words and rotating accumulators
static void fusion_epoch_buffers(unsigned int* counter, FusionEpochArgs& a) {
  unsigned char* base = reinterpret_cast<unsigned char*>(counter) + 16;
  a.bar = reinterpret_cast<unsigned long long*>(base);                       // 3 x u64 (64 bytes reserved)
  a.acc3 = reinterpret_cast<double*>(base + 64);                             // [3][kMailDoubles]
  a.tot3 = reinterpret_cast<double*>(base + 64 + 3 * kMailDoubles * sizeof(double));
}

int mpu_fusion_train_epoch(const float* X, const unsigned char* y, const long long* perm, long long n,
                           long long batch, int V, int C, float* W, float* b, float* m, float* v, double* accum,
                           unsigned int* counter, double* losses_out, float reg, float lr, float beta1, float beta2,
                           float eps, int first_step, void* stream) {
  if (!X || !y || !W || !b || !m || !v || !accum || !counter || batch < 1 || n < 1 || first_step < 1) {
    set_error("mpu_fusion_train_epoch: bad arguments");
    return MPU_ERR_ARG;
  }
  // MPU_FUSION_EPOCH_PERSISTENT=0: one launch per batch instead of the persistent epoch kernel (read per call)
  const char* env_p = getenv("MPU_FUSION_EPOCH_PERSISTENT");
  const bool per_batch = env_p && atoi(env_p) == 0;
  if (!per_batch) {
    FusionEpochArgs a;
    memset(&a, 0, sizeof(a));
    a.X = X; a.y = y; a.perm = perm; a.n = n; a.batch = batch; a.n_batches = (n + batch - 1) / batch;
    a.W = W; a.b = b; a.m = m; a.v = v;
    fusion_epoch_buffers(counter, a);
    a.losses_out = losses_out;
    a.reg = reg; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
    a.first_step = first_step;
    a.world = 1;
    bool handled = false;
    MPU_TRY(fusion_epoch_dispatch(a, V, C, reinterpret_cast<cudaStream_t>(stream), &handled));
    if (handled) return MPU_OK;
  }
  int step = first_step;
  long long k = 0;
  for (long long s = 0; s < n; s += batch, ++k, ++step) {
    const long long nb = n - s < batch ? n - s : batch;
    MPU_TRY(mpu_fusion_train_step(X, y, perm ? perm + s : nullptr, nb, V, C, W, b, m, v, accum, counter,
                                  losses_out ? losses_out + k : nullptr, reg, lr, beta1, beta2, eps, step, stream));
    if (!perm) {  // contiguous rows: advance the base pointers instead
      X += nb * (long long)V * C;
      y += nb;
    }
  }
  return MPU_OK;
}

int mpu_fusion_scratch_bytes(void) { return 16 + kFusionMaxBlocks * kMailDoubles * (int)sizeof(double); }

int mpu_fusion_mailbox_bytes(void) { return (int)(kMailFlagOffset + sizeof(unsigned long long) * 2 * kMaxPeers); }

int mpu_fusion_train_epoch_peer(const float* X, const unsigned char* y, const long long* perm, long long n,
                                long long batch, long long n_batches, int V, int C, float* W, float* b, float* m,
                                float* v, double* accum, unsigned int* counter, double* losses_out, float reg, float lr,
                                float beta1, float beta2, float eps, int first_step, const void* const* h_peer_mail,
                                int world, int rank, unsigned long long first_seq, void* stream) {
  if (!X || !y || !perm || !W || !b || !m || !v || !accum || !counter || !h_peer_mail || batch < 1 || n_batches < 1 ||
      world < 2 || world > kMaxPeers || rank < 0 || rank >= world || first_step < 1 || first_seq < 1) {
    set_error("mpu_fusion_train_epoch_peer: bad arguments (world=%d rank=%d)", world, rank);
    return MPU_ERR_ARG;
  }
  if (V * C + C + 2 > kMailDoubles) {
    set_error("mpu_fusion_train_epoch_peer: V*C + C + 2 = %d exceeds the mailbox slot (%d doubles)", V * C + C + 2,
              kMailDoubles);
    return MPU_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    const char* env_p = getenv("MPU_FUSION_EPOCH_PERSISTENT");
    const bool per_batch = env_p && atoi(env_p) == 0;
    if (!per_batch) {
      FusionEpochArgs a;
      memset(&a, 0, sizeof(a));
      a.X = X; a.y = y; a.perm = perm; a.n = n; a.batch = batch; a.n_batches = n_batches;
      a.W = W; a.b = b; a.m = m; a.v = v;
      fusion_epoch_buffers(counter, a);
      a.losses_out = losses_out;
      a.reg = reg; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
      a.first_step = first_step;
      a.world = world; a.rank = rank; a.first_seq = first_seq;
      for (int q = 0; q < world; ++q) a.mail[q] = reinterpret_cast<unsigned char*>(const_cast<void*>(h_peer_mail[q]));
      bool handled = false;
      MPU_TRY(fusion_epoch_dispatch(a, V, C, st, &handled));
      if (handled) return MPU_OK;
    }
  }
  for (long long k = 0; k < n_batches; ++k) {
    const long long s0 = k * batch;
    // every rank launches exactly n_batches exchanges; a rank that ran out of points contributes one (zero-weight
    // impossible: count 0 is fine) empty batch by pointing at its first row with n = 0 handled below
    long long nb = s0 < n ? (n - s0 < batch ? n - s0 : batch) : 0;
    FusionStepArgs fs;
    memset(&fs, 0, sizeof(fs));
    fs.Wp = W; fs.bp = b; fs.m = m; fs.v = v;
    fs.counter = counter;
    fs.scratch = nullptr;
    fs.loss_out = losses_out ? losses_out + k : nullptr;
    fs.reg = reg;
    const int step = first_step + (int)k;
    fs.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
    fs.b1 = beta1; fs.b2 = beta2; fs.eps = eps;
    fs.fused = 1;
    fs.world = world;
    fs.rank = rank;
    fs.seq = first_seq + (unsigned long long)k;
    for (int q = 0; q < world; ++q) fs.mail[q] = reinterpret_cast<unsigned char*>(const_cast<void*>(h_peer_mail[q]));
    bool handled = false;
    // n = 0 still launches one block: it takes part in the exchange with zero sums and a zero count
    MPU_TRY(fusion_grad_dispatch(X, y, perm + (nb > 0 ? s0 : 0), nb, V, C, W, b, accum, fs, st, &handled));
    if (!handled) {
      set_error("mpu_fusion_train_epoch_peer: (V=%d, C=%d) not instantiated", V, C);
      return MPU_ERR_ARG;
    }
  }
  return MPU_OK;
}

int mpu_fusion_adam(float* W, float* b, float* m, float* v, const double* accum, double n_points,
                    int V, int C, float reg, float lr, float beta1, float beta2, float eps, int step,
                    void* stream) {
  if (!W || !b || !m || !v || !accum) {
    set_error("mpu_fusion_adam: bad arguments");
    return MPU_ERR_ARG;
  }
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  const int n = V * C + C;
  fusion_adam_kernel<<<(n + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, b, m, v, accum, n_points, V, C, reg, (float)lr_t, beta1, beta2, eps);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_voxel_leaf_sums(const int* h_dims, const double* h_affine3x3, const long long* leaf_start, const int* leaf_len,
                        long long n_leaves, double* out, void* stream) {
  if (!h_dims || !h_affine3x3 || !leaf_start || !leaf_len || !out || n_leaves < 1) {
    set_error("mpu_voxel_leaf_sums: bad arguments");
    return MPU_ERR_ARG;
  }
  const double* A = h_affine3x3;
  const long long work = 3 * n_leaves;
  voxel_leaf_sums_kernel<<<(unsigned)((work + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      h_dims[1], h_dims[2], A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], A[8], leaf_start, leaf_len, n_leaves, out);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_label_counts(const unsigned char* y_true, const unsigned char* y_pred, const float* scores,
                     long long n, int n_classes, long long* counts, void* stream) {
  if (!y_true || (!y_pred && !scores) || !counts || n < 0 || n_classes < 1 || n_classes > 256) {
    set_error("label_counts: bad argument (n=%lld, n_classes=%d)", n, n_classes);
    return MPU_ERR_ARG;
  }
  if (n == 0) return MPU_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  label_counts_kernel<<<(int)blocks, 256, 0, st>>>(y_true, scores ? nullptr : y_pred, scores, n, n_classes,
                                                   reinterpret_cast<unsigned long long*>(counts));
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_elastic_2d(const float* x_in, const unsigned char* y_in, double* fields, double* scratch,
                   const double* d_weights, int weight_stride, const int* d_radius, const double* d_alpha,
                   const float* d_bg, int n, int H, int W, int C, float* x_out, unsigned char* y_out,
                   void* stream) {
  if (!x_in || !fields || !scratch || !d_weights || !d_radius || !d_alpha || !d_bg || !x_out || n < 0 ||
      H < 2 || W < 2 || C < 1 || (y_in && !y_out)) {
    set_error("elastic_2d: bad argument (n=%d, H=%d, W=%d, C=%d)", n, H, W, C);
    return MPU_ERR_ARG;
  }
  if (n == 0) return MPU_OK;
  if (x_in == x_out || (y_in && y_in == y_out)) {
    set_error("elastic_2d: in-place operation is not supported");
    return MPU_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int bx = (H * W + 255) / 256;
  if (bx > sm_count() * 4) bx = sm_count() * 4;
  const dim3 gf(bx, 2 * n), gs(bx, n);
  gauss1d_kernel<<<gf, 256, 0, st>>>(fields, scratch, H, W, 0, d_weights, weight_stride, d_radius);
  gauss1d_kernel<<<gf, 256, 0, st>>>(scratch, fields, H, W, 1, d_weights, weight_stride, d_radius);
  elastic_resample_kernel<<<gs, 256, 0, st>>>(x_in, y_in, fields, d_alpha, d_bg, H, W, C, x_out, y_out);
  count_launch(3);
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int mpu_interp_points(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                      const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                      const double* coords, long long n, const float* h_bg_value, int bg_class,
                      float* out_f32, unsigned char* out_labels, void* stream) {
  if (n == 0) return MPU_OK;
  if (!vol || !h_dims || !gx || !gy || !gz || !h_inv_step || !coords || n < 0 || (!out_f32 && !out_labels)) {
    set_error("mpu_interp_points: bad arguments");
    return MPU_ERR_ARG;
  }
  if (C < 1 || C > kMaxCh) {
    set_error("mpu_interp_points: %d channels unsupported (max %d)", C, kMaxCh);
    return MPU_ERR_ARG;
  }
  if (h_dims[0] < 2 || h_dims[1] < 2 || h_dims[2] < 2) {
    set_error("mpu_interp_points: every volume axis needs at least 2 voxels");
    return MPU_ERR_ARG;
  }
  PointsParams p;
  memset(&p, 0, sizeof(p));
  p.vol = vol;
  p.labels = labels;
  p.X = h_dims[0];
  p.Y = h_dims[1];
  p.Z = h_dims[2];
  p.C = C;
  p.gx = gx;
  p.gy = gy;
  p.gz = gz;
  for (int k = 0; k < 3; ++k) p.inv_step[k] = h_inv_step[k];
  p.coords = coords;
  p.n = n;
  for (int c = 0; c < C; ++c) p.bg_value[c] = h_bg_value ? h_bg_value[c] : 0.f;
  p.bg_class = bg_class;
  p.out_f32 = out_f32;
  p.out_lab = out_labels;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sample_points_kernel<<<grid_for(n, 256), 256, 0, st>>>(p);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

}  // extern "C"
