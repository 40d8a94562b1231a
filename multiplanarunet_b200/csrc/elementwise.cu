// See elementwise.cuh.  HBM-bound CUDA-core kernels; 16-byte vector access, channel-group-per-thread.
#include "elementwise.cuh"
#include "common.h"

#include <cstdio>
#include <cstdlib>

namespace mpu {

namespace {

constexpr int kMaxBlocks = 148 * 8;

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Vec8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    r.v[2 * j] = f.x;
    r.v[2 * j + 1] = f.y;
  }
  return r;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(r.v[2 * j], r.v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// thread layout for per-channel reductions: blockDim = CG * RL, tid -> (cg = tid % CG, rl = tid / CG)
inline void reduce_layout(int C, int* CG, int* RL, int* threads) {
  *CG = C / 8;
  int rl = 256 / *CG;
  if (rl < 1) rl = 1;
  *RL = rl;
  *threads = *CG * rl;
}

// Block-level reduction of per-thread 8-channel partials over the RL row lanes, then atomics.
template <int NV, typename OutT>
__device__ __forceinline__ void block_reduce_store(float (&acc)[NV][8], int CG, int RL, int C,
                                                   OutT* out /*[NV][C]*/) {
  extern __shared__ float red_smem[];  // [NV][RL][CG*8]
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
#pragma unroll
  for (int n = 0; n < NV; ++n)
#pragma unroll
    for (int j = 0; j < 8; ++j) red_smem[(n * RL + rl) * (CG * 8) + cg * 8 + j] = acc[n][j];
  __syncthreads();
  for (int i = tid; i < NV * CG * 8; i += blockDim.x) {
    const int n = i / (CG * 8), c = i % (CG * 8);
    float s = 0.f;
    for (int r = 0; r < RL; ++r) s += red_smem[(n * RL + r) * (CG * 8) + c];
    atomicAdd(out + (size_t)n * C + c, (OutT)s);
  }
}

// ------------------------------------------------------------------------------------------------
// Row loops of the pure reductions keep 4 independent 16-byte loads in flight per thread (these kernels
// are latency-bound HBM streams: bytes in flight per SM, not instruction count, sets their speed).
__global__ void channel_stats_kernel(const __nv_bfloat16* __restrict__ y, long long rows, int C, int ld,
                                     int CG, int RL, double* __restrict__ sums) {
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const long long stride = (long long)gridDim.x * RL;
  for (long long r = (long long)blockIdx.x * RL + rl; r < rows; r += 4 * stride) {
    Vec8 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long rr = r + u * stride;
      if (rr < rows) {
        x[u] = load8(y + rr * ld + cg * 8);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u].v[j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += x[u].v[j];
        acc[1][j] = fmaf(x[u].v[j], x[u].v[j], acc[1][j]);
      }
  }
  block_reduce_store<2, double>(acc, CG, RL, C, sums);
}

__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, int ld, int CG,
                              int RL, float* __restrict__ out) {
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  const long long stride = (long long)gridDim.x * RL;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  auto issue = [&](long long r, uint4 (&v)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long rr = r + u * stride;
      v[u] = rr < rows ? *reinterpret_cast<const uint4*>(x + rr * ld + cg * 8) : zero4;
    }
  };
  // the loads of the next iteration are in flight while this one is summed (see bn_bwd_kernel)
  uint4 cur[4], nxt[4];
  long long r = (long long)blockIdx.x * RL + rl;
  if (r < rows) issue(r, cur);
  for (; r < rows; r += 4 * stride) {
    const bool more = r + 4 * stride < rows;
    if (more) issue(r + 4 * stride, nxt);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        acc[0][2 * j] += f.x;
        acc[0][2 * j + 1] += f.y;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
  }
  block_reduce_store<1, float>(acc, CG, RL, C, out);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ mmean, float* __restrict__ mvar, float eps,
                                   float momentum, int training, int C, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = sums[c] / count;
    double v = sums[C + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    const double unbiased = count > 1 ? v * count / (count - 1) : v;
    mmean[c] = mmean[c] * momentum + mean * (1.f - momentum);
    mvar[c] = mvar[c] * momentum + (float)unbiased * (1.f - momentum);
  } else {
    mean = mmean[c];
    var = mvar[c];
  }
  const float rstd = 1.0f / sqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

// BN coefficients of one channel, computed by every thread for its own channels (the arithmetic of bn_finalize_kernel,
// which this replaces on the U-Net's schedule: 13 tiny launches on the forward critical path per step).  `writer`
// threads (one per channel: block 0, row lane 0) also store scale / shift / mean / rstd for the backward pass and
// update the moving statistics; in training mode nobody else reads the moving statistics, so there is no race.
__device__ __forceinline__ void bn_coef(const BnFin& f, int c, bool writer, float& sc, float& sh) {
  if (f.no_write) {  // coefficients already finalised by a separate launch (bring-up comparison)
    sc = f.scale[c];
    sh = f.shift[c];
    return;
  }
  float mean, var;
  if (f.training) {
    const double m = f.sums[c] / f.count;
    double v = f.sums[f.C + c] / f.count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    if (writer) {
      const double unbiased = f.count > 1 ? v * f.count / (f.count - 1) : v;
      f.mmean[c] = f.mmean[c] * f.momentum + mean * (1.f - f.momentum);
      f.mvar[c] = f.mvar[c] * f.momentum + (float)unbiased * (1.f - f.momentum);
    }
  } else {
    mean = f.mmean[c];
    var = f.mvar[c];
  }
  const float rstd = 1.0f / sqrtf(var + f.eps);
  sc = f.gamma[c] * rstd;
  sh = f.beta[c] - mean * sc;
  if (writer) {
    f.scale[c] = sc;
    f.shift[c] = sh;
    if (f.mean_out) f.mean_out[c] = mean;
    if (f.rstd_out) f.rstd_out[c] = rstd;
  }
}

// interior pixel index -> padded row
__device__ __forceinline__ long long padded_row(const Geo& g, long long pix) {
  const int x = (int)(pix % g.W);
  const long long t = pix / g.W;
  const int yy = (int)(t % g.H);
  const long long n = t / g.H;
  return n * (long long)(g.H + 2) * (g.W + 2) + (long long)(yy + 1) * (g.W + 2) + (x + 1);
}

// Work distribution of the per-pixel kernels: a work item is a segment of one image line - U*RL consecutive
// pixels (or 2x2 windows) handled by the RL row lanes of a block, every thread keeping its channel group.
// Items are walked with a mixed-radix counter (image, line, segment), so the hot loops contain no integer
// division (the padded-row arithmetic used to cost more instructions than the BN math itself).
struct SegIter {
  int n, yy, seg;
  int dn, dy, ds;
  int H, nseg;
  __device__ __forceinline__ void init(int start, int stride, int H_, int nseg_) {
    H = H_;
    nseg = nseg_;
    seg = start % nseg;
    int t = start / nseg;
    yy = t % H;
    n = t / H;
    ds = stride % nseg;
    t = stride / nseg;
    dy = t % H;
    dn = t / H;
  }
  __device__ __forceinline__ void next() {
    seg += ds;
    yy += dy;
    n += dn;
    if (seg >= nseg) {
      seg -= nseg;
      ++yy;
    }
    if (yy >= H) {
      yy -= H;
      ++n;
    }
  }
};

__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ y, BnFin f, __nv_bfloat16* __restrict__ b, Geo g,
                                int C, int CG, int RL) {
  constexpr int U = 4;  // independent 16-byte loads per thread and work item
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bn_coef(f, cg * 8 + j, blockIdx.x == 0 && rl == 0, sc[j], sh[j]);
  const int Wp = g.W + 2;
  auto base_of = [&](const SegIter& it) {
    return ((long long)it.n * (g.H + 2) + it.yy + 1) * Wp + 1 + it.seg * (U * RL) + rl;
  };
  auto issue = [&](const SegIter& it, uint4 (&v)[U]) {
    const long long base = base_of(it);
    const int x0 = it.seg * (U * RL) + rl;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (x0 + u * RL < g.W) v[u] = *reinterpret_cast<const uint4*>(y + (base + u * RL) * C + cg * 8);
  };
  // the loads of the next work item are in flight while this one is scaled and stored (see bn_bwd_kernel)
  SegIter it;
  it.init(blockIdx.x, gridDim.x, g.H, (g.W + U * RL - 1) / (U * RL));
  uint4 cur[U], nxt[U];
  bool have = it.n < g.B;
  if (have) issue(it, cur);
  while (have) {
    SegIter in = it;
    in.next();
    const bool have_n = in.n < g.B;
    if (have_n) issue(in, nxt);
    const long long base = base_of(it);
    const int x0 = it.seg * (U * RL) + rl;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (x0 + u * RL < g.W) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur[u]);
        Vec8 v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          v.v[2 * j] = fmaf(f.x, sc[2 * j], sh[2 * j]);
          v.v[2 * j + 1] = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
        }
        store8(b + (base + u * RL) * C + cg * 8, v);
      }
    it = in;
    have = have_n;
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
}

// window (2x2) version: writes the four BN outputs and their max into the half-resolution tensor.
__global__ void bn_apply_pool_kernel(const __nv_bfloat16* __restrict__ y, BnFin f,
                                     __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ pooled,
                                     Geo g, int C, int CG, int RL) {
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  const int h = g.H / 2, w = g.W / 2;
  const int Wp = g.W + 2;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bn_coef(f, cg * 8 + j, blockIdx.x == 0 && rl == 0, sc[j], sh[j]);
  SegIter it;
  it.init(blockIdx.x, gridDim.x, h, (w + RL - 1) / RL);
  for (; it.n < g.B; it.next()) {
    const int jx = it.seg * RL + rl, iy = it.yy;
    if (jx >= w) continue;
    const long long base = ((long long)it.n * (g.H + 2) + 2 * iy + 1) * Wp + (2 * jx + 1);
    Vec8 v[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) v[d] = load8(y + (base + (d >> 1) * Wp + (d & 1)) * C + cg * 8);
    Vec8 mx;
#pragma unroll
    for (int j = 0; j < 8; ++j) mx.v[j] = -INFINITY;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[d].v[j] = bf16_round(fmaf(v[d].v[j], sc[j], sh[j]));
        mx.v[j] = fmaxf(mx.v[j], v[d].v[j]);
      }
      store8(b + (base + (d >> 1) * Wp + (d & 1)) * C + cg * 8, v[d]);
    }
    const long long prow = ((long long)it.n * (h + 2) + iy + 1) * (w + 2) + (jx + 1);
    store8(pooled + prow * C + cg * 8, mx);
  }
}

// ---- BN backward ---------------------------------------------------------------------------------
// Work item = segment of pixels (or 2x2 windows) x channel group.  g(p,c) = gA[p] + (p is the first arg-max
// of its window ? gP[window] : 0).  Window arg-max is recomputed from the bf16-rounded BN outputs, scanning
// row-major so the first maximum wins (the forward max-pool kept no index).
// Pass 1 accumulates sum(g) and sum(g*y); with xhat = (y-mu)*rstd the BN backward is then the per-channel
// affine map dz = relu'(y) * (a*g + b*y + c):  a = gamma*rstd, b = -a*rstd*mgx, c = -a*mg - b*mu  where
// mg = mean(g), mgx = mean(g*xhat) = rstd*(mean(g*y) - mu*mg).  Few per-thread coefficients keep the
// register count low enough for 3 resident blocks per SM (these kernels are pure HBM streams).
// V channels (8 or 4) per thread as one raw 16- or 8-byte register group
template <int V>
struct RawV;
template <>
struct RawV<8> {
  typedef uint4 type;
};
template <>
struct RawV<4> {
  typedef uint2 type;
};
template <int V>
__device__ __forceinline__ void unpackv(const typename RawV<V>::type& u, float (&r)[V]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < V / 2; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    r[2 * j] = f.x;
    r[2 * j + 1] = f.y;
  }
}
template <int V>
__device__ __forceinline__ typename RawV<V>::type packv(const float (&r)[V]) {
  typename RawV<V>::type u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < V / 2; ++j) h[j] = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
  return u;
}

// Block-level reduction of per-thread V-channel partials over the RL row lanes, then atomics.
template <int NVAL, typename OutT, int V>
__device__ __forceinline__ void block_reduce_store_v(float (&acc)[NVAL][V], int CG, int RL, int C,
                                                     OutT* out /*[NVAL][C]*/) {
  extern __shared__ float red_smem[];  // [NVAL][RL][CG*V]
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
#pragma unroll
  for (int n = 0; n < NVAL; ++n)
#pragma unroll
    for (int j = 0; j < V; ++j) red_smem[(n * RL + rl) * (CG * V) + cg * V + j] = acc[n][j];
  __syncthreads();
  for (int i = tid; i < NVAL * CG * V; i += blockDim.x) {
    const int n = i / (CG * V), c = i % (CG * V);
    float s = 0.f;
    for (int r = 0; r < RL; ++r) s += red_smem[(n * RL + r) * (CG * V) + c];
    atomicAdd(out + (size_t)n * C + c, (OutT)s);
  }
}

// The loads of work item k+1 are issued (raw registers) BEFORE item k is processed: these kernels are pure HBM
// streams, and with load -> wait -> compute -> store per item the bytes in flight per SM dropped to zero once per
// iteration (3.1-4.6 TB/s under ncu where the simple streaming Adam kernel reaches 5.7).  The 2x2-window (POOL)
// variants keep 9 loads per item and buffer set; they process a window in two halves of 4 channels so that two buffer
// sets, the coefficients and the arg-max temporaries stay below the 170 registers that allow two resident blocks.
// (Measured alternative: 4 channels per thread with 8-byte accesses - half the state per thread - ran 2x SLOWER.)
template <bool POOL, bool APPLY>
__global__ void __launch_bounds__(256) __maxnreg__(POOL ? 168 : 104)
    bn_bwd_kernel(BnBwdArgs a, int CG, int RL, const double* __restrict__ sums_in,
                  double* __restrict__ sums_out, __nv_bfloat16* __restrict__ dz, int phase_major,
                  float* __restrict__ dbias) {
  constexpr int V = 8;              // channels per thread
  constexpr int U = POOL ? 1 : 2;
  constexpr int NV = POOL ? 4 : U;  // pixels per work item and thread
  typedef typename RawV<V>::type raw_t;
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  const int C = a.C;
  const Geo g = a.g;
  const int Wp = g.W + 2;
  const int hh = g.H / 2, wh = g.W / 2;
  float sc[POOL ? V : 1], sh[POOL ? V : 1];
  float ca[APPLY ? V : 1], cb[APPLY ? V : 1], cc[APPLY ? V : 1];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int c = cg * V + j;
    if (POOL) {
      sc[j] = a.scale[c];
      sh[j] = a.shift[c];
    }
    if (APPLY) {
      const double cnt = (double)g.pixels();
      const float mu = a.mean[c], rs = a.rstd[c];
      if (a.dgamma != nullptr && blockIdx.x == 0 && rl == 0) {
        // dbeta += sum g ; dgamma += sum g*xhat = rstd * (sum g*y - mu * sum g)   (was bn_bwd_params_kernel)
        a.dbeta[c] += (float)sums_in[c];
        a.dgamma[c] += (float)((double)rs * (sums_in[C + c] - (double)mu * sums_in[c]));
      }
      const float mg = (float)(sums_in[c] / cnt);
      const float mgy = (float)(sums_in[C + c] / cnt);
      const float mgx = rs * (mgy - mu * mg);
      const float k1 = a.gamma[c] * rs;
      ca[j] = k1;
      cb[j] = -k1 * rs * mgx;
      cc[j] = -k1 * mg - cb[j] * mu;
    }
  }
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[0][j] = acc[1][j] = 0.f;
  const long long rows_lo = (long long)g.B * (hh + 2) * (wh + 2);

  // one pixel: accumulate the reduction terms (pass 1) or write dz (pass 2)
  auto emit = [&](const float (&gv)[V], const float (&yv)[V], int n, int yy, int xx, long long row) {
    if (!APPLY) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        acc[0][j] += gv[j];
        acc[1][j] = fmaf(gv[j], yv[j], acc[1][j]);
      }
    } else {
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float v = fmaf(ca[j], gv[j], fmaf(cb[j], yv[j], cc[j]));
        if (!(yv[j] > 0.f)) v = 0.f;
        o[j] = v;
      }
      // round to the stored bf16 values first (packed conversion): the bias-gradient sums are taken over those
      const raw_t u = packv<V>(o);
      float r[V];
      unpackv<V>(u, r);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[0][j] += r[j];
      long long orow = row;
      if (phase_major)  // pixel (n, yy, xx) of this level -> [phase][rows of the half-resolution level]
        orow = ((yy & 1) * 2 + (xx & 1)) * rows_lo + ((long long)n * (hh + 2) + (yy >> 1) + 1) * (wh + 2) +
               ((xx >> 1) + 1);
      *reinterpret_cast<raw_t*>(dz + orow * C + cg * V) = u;
    }
  };

  // raw loads of one work item
  struct Raw {
    raw_t y[NV], g[NV], gp;
  };
  raw_t zero;
  memset(&zero, 0, sizeof(zero));
  auto live = [&](const SegIter& it, int u) {  // does position u of the item exist?
    return POOL ? (it.seg * RL + rl < wh) : (it.seg * (U * RL) + rl + u * RL < g.W);
  };
  auto row_of = [&](const SegIter& it, int u) -> long long {
    if (POOL) {
      const long long base = ((long long)it.n * (g.H + 2) + 2 * it.yy + 1) * Wp + (2 * (it.seg * RL + rl) + 1);
      return base + (u >> 1) * Wp + (u & 1);
    }
    return ((long long)it.n * (g.H + 2) + it.yy + 1) * Wp + 1 + it.seg * (U * RL) + rl + u * RL;
  };
  auto issue = [&](const SegIter& it, Raw& r) {
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      r.y[u] = zero;
      r.g[u] = zero;
      if (live(it, u)) {
        const long long row = row_of(it, u);
        r.y[u] = *reinterpret_cast<const raw_t*>(a.y + row * C + cg * V);
        if (a.gA) r.g[u] = *reinterpret_cast<const raw_t*>(a.gA + row * (long long)a.ldA + cg * V);
      }
    }
    r.gp = zero;
    if (POOL && a.gP && live(it, 0)) {
      const long long prow = ((long long)it.n * (hh + 2) + it.yy + 1) * (wh + 2) + (it.seg * RL + rl + 1);
      r.gp = *reinterpret_cast<const raw_t*>(a.gP + prow * C + cg * V);
    }
  };
  auto consume = [&](const SegIter& it, const Raw& r) {
    if (!live(it, 0)) return;  // (POOL: the whole window; line kernels: position 0 exists whenever any does)
    if (POOL) {
      uint4 outw[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // channels hf*4 .. hf*4+3 of the thread's group
        float yv[4][4], gv[4][4], gp[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          unpackv<4>(reinterpret_cast<const uint2*>(&r.y[d])[hf], yv[d]);
          unpackv<4>(reinterpret_cast<const uint2*>(&r.g[d])[hf], gv[d]);
        }
        unpackv<4>(reinterpret_cast<const uint2*>(&r.gp)[hf], gp);
        if (a.gP) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float best = -INFINITY;
            int am = 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              const float bv = bf16_round(fmaf(yv[d][j], sc[hf * 4 + j], sh[hf * 4 + j]));
              if (bv > best) {
                best = bv;
                am = d;
              }
            }
#pragma unroll
            for (int d = 0; d < 4; ++d)
              if (am == d) gv[d][j] += gp[j];
          }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          if (!APPLY) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[0][hf * 4 + j] += gv[d][j];
              acc[1][hf * 4 + j] = fmaf(gv[d][j], yv[d][j], acc[1][hf * 4 + j]);
            }
          } else {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v = fmaf(ca[hf * 4 + j], gv[d][j], fmaf(cb[hf * 4 + j], yv[d][j], cc[hf * 4 + j]));
              if (!(yv[d][j] > 0.f)) v = 0.f;
              o[j] = v;
            }
            const uint2 u = packv<4>(o);
            float rr[4];
            unpackv<4>(u, rr);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[0][hf * 4 + j] += rr[j];
            reinterpret_cast<uint2*>(&outw[d])[hf] = u;
          }
        }
      }
      if (APPLY) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          long long orow = row_of(it, d);
          if (phase_major) {
            const int yy = 2 * it.yy + (d >> 1), xx = 2 * (it.seg * RL + rl) + (d & 1);
            orow = ((yy & 1) * 2 + (xx & 1)) * rows_lo + ((long long)it.n * (hh + 2) + (yy >> 1) + 1) * (wh + 2) +
                   ((xx >> 1) + 1);
          }
          *reinterpret_cast<uint4*>(dz + orow * C + cg * 8) = outw[d];
        }
      }
      return;
    }
    float yv[NV][V], gv[NV][V];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      unpackv<V>(r.y[u], yv[u]);
      unpackv<V>(r.g[u], gv[u]);
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      if (!live(it, u)) continue;
      emit(gv[u], yv[u], it.n, it.yy, it.seg * (U * RL) + rl + u * RL, row_of(it, u));
    }
  };

  SegIter it;
  it.init(blockIdx.x, gridDim.x, POOL ? hh : g.H, ((POOL ? wh : g.W) + U * RL - 1) / (U * RL));
  Raw cur, nxt;
  bool have = it.n < g.B;
  if (have) issue(it, cur);
  while (have) {
    SegIter in = it;
    in.next();
    const bool have_n = in.n < g.B;
    if (have_n) issue(in, nxt);
    consume(it, cur);
    it = in;
    have = have_n;
    cur = nxt;
  }
  if (!APPLY) {
    block_reduce_store_v<2, double, V>(acc, CG, RL, C, sums_out);
  } else if (dbias) {
    float acc1[1][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc1[0][j] = acc[0][j];
    block_reduce_store_v<1, float, V>(acc1, CG, RL, C, dbias);
  }
}

// sums = [sum g | sum g*y]  ->  dbeta += sum g ; dgamma += sum g*xhat = rstd * (sum g*y - mu * sum g)
__global__ void bn_bwd_params_kernel(const double* __restrict__ sums, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] += (float)sums[c];
  dgamma[c] += (float)((double)rstd[c] * (sums[C + c] - (double)mean[c] * sums[c]));
}

// ---- first conv (CUDA cores) -------------------------------------------------------------------------
// thread = (8-channel group cg, row lane rl) like the other per-pixel kernels, a work item = a band of kRows image rows
// x RL consecutive pixels: the thread holds the (kRows+2) x 3 x CIN inputs of its column of kRows pixels in registers
// and reads the weights of ITS channel group from shared memory (fp32 [9*CIN][co_phys]; a warp reads 32-byte runs of
// consecutive groups, conflict-free): one pair of 16-byte weight reads feeds kRows x 8 FMAs.  A warp stores
// 512 contiguous bytes (the round-1 mapping, thread = pixel with a loop over channel groups, wrote 16-byte pieces
// 192 bytes apart and ran at 1.6 TB/s; this kernel only writes: 0.4 GB at 256^2, batch 32).
constexpr int kFirstRows = 4;

template <int CIN>
__global__ void __launch_bounds__(256)
    conv_first_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, Geo g, int co_phys, int CG,
                      int RL) {
  constexpr int R = kFirstRows;
  extern __shared__ float4 cw4[];  // [9*CIN][co_phys] + bias[co_phys]
  float* cw = reinterpret_cast<float*>(cw4);
  for (int i = threadIdx.x; i < 9 * CIN * co_phys; i += blockDim.x) {
    const int co = i % co_phys;
    const int t2 = i / co_phys;
    const int ci = t2 % CIN, tap = t2 / CIN;
    cw[i] = __bfloat162float(w[((long long)tap * co_phys + co) * 8 + ci]);
  }
  float* sb = cw + 9 * CIN * co_phys;
  for (int i = threadIdx.x; i < co_phys; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int Wp = g.W + 2;
  const int bands = (g.H + R - 1) / R;
  const float4 b0 = *reinterpret_cast<const float4*>(sb + cg * 8);
  const float4 b1 = *reinterpret_cast<const float4*>(sb + cg * 8 + 4);
  SegIter it;
  it.init(blockIdx.x, gridDim.x, bands, (g.W + RL - 1) / RL);
  for (; it.n < g.B; it.next()) {
    const int xx = it.seg * RL + rl;
    if (xx >= g.W) continue;
    const int y0 = it.yy * R;  // first image row of the band
    // padded row of pixel (y0 - 1, xx - 1): top-left input of the band's first pixel
    const long long row0 = ((long long)it.n * (g.H + 2) + y0) * Wp + xx;
    float xs[R + 2][3 * CIN];
#pragma unroll
    for (int r = 0; r < R + 2; ++r) {
      const bool live = y0 + r - 1 <= g.H;  // rows past the bottom border belong to the next image
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const long long rr = row0 + (long long)r * Wp + dx;
        if (CIN == 1) {
          xs[r][dx] = live ? __bfloat162float(x[rr * 8]) : 0.f;
        } else {
          Vec8 xv;
          if (live) xv = load8(x + rr * 8);
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) xs[r][dx * CIN + ci] = live ? xv.v[ci] : 0.f;
        }
      }
    }
    float acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
      acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int k = 0; k < 3 * CIN; ++k) {
        const float* wr = cw + ((ky * 3) * CIN + k) * co_phys + cg * 8;  // tap (ky, k / CIN), channel k % CIN
        const float4 w0 = *reinterpret_cast<const float4*>(wr);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float xv = xs[r + ky][k];
          acc[r][0] = fmaf(xv, w0.x, acc[r][0]);
          acc[r][1] = fmaf(xv, w0.y, acc[r][1]);
          acc[r][2] = fmaf(xv, w0.z, acc[r][2]);
          acc[r][3] = fmaf(xv, w0.w, acc[r][3]);
          acc[r][4] = fmaf(xv, w1.x, acc[r][4]);
          acc[r][5] = fmaf(xv, w1.y, acc[r][5]);
          acc[r][6] = fmaf(xv, w1.z, acc[r][6]);
          acc[r][7] = fmaf(xv, w1.w, acc[r][7]);
        }
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (y0 + r < g.H) {
        Vec8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaxf(acc[r][j], 0.f);
        store8(out + (row0 + (long long)(r + 1) * Wp + 1) * co_phys + cg * 8, o);
      }
    }
  }
}

// Weight gradient of the first conv: dW[tap][co][ci] += sum_p dz[p][co] * x[p + tap][ci] for one input channel
// ci = blockIdx.y.  Thread = (8-channel group of dz, row lane); the 9 x 8 partial sums live in registers and are
// reduced over the block's row lanes through shared memory, then added with one atomic per value.
__global__ void __launch_bounds__(256)
    conv_first_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz, Geo g,
                            int co_phys, int CG, int RL, float* __restrict__ dW, int ldw,
                            float* __restrict__ dbias) {
  constexpr int U = 4;  // pixels per thread and iteration: U dz rows + 9 U input taps in flight (with one pixel per
                        // iteration the kernel ran at 1.45 TB/s, bound by the latency of its single 16-byte load)
  extern __shared__ float wred[];  // [RL][CG*8]
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  const int ci = blockIdx.y;
  const int Wp = g.W + 2;
  float acc[9][8];
  float bs[8];  // column sums of dz: the conv's bias gradient (input channel 0's blocks only)
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  SegIter it;
  it.init(blockIdx.x, gridDim.x, g.H, (g.W + U * RL - 1) / (U * RL));
  for (; it.n < g.B; it.next()) {
    const int x0 = it.seg * (U * RL) + rl;
    const long long row0 = ((long long)it.n * (g.H + 2) + it.yy + 1) * Wp + x0 + 1;
    uint4 dr[U];
    __nv_bfloat16 xr[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (x0 + u * RL < g.W) {
        const long long row = row0 + u * RL;
        dr[u] = *reinterpret_cast<const uint4*>(dz + row * co_phys + cg * 8);
#pragma unroll
        for (int t = 0; t < 9; ++t) xr[u][t] = x[(row + (t / 3 - 1) * Wp + (t % 3 - 1)) * 8 + ci];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (x0 + u * RL < g.W) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&dr[u]);
        float d[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          d[2 * j] = f.x;
          d[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) bs[j] += d[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xv = __bfloat162float(xr[u][t]);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(d[j], xv, acc[t][j]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wred[tid * 8 + j] = acc[t][j];
    __syncthreads();
    for (int c = tid; c < CG * 8; c += blockDim.x) {
      float sum = 0.f;
      for (int r = 0; r < RL; ++r) sum += wred[r * (CG * 8) + c];
      atomicAdd(dW + ((long long)t * co_phys + c) * ldw + ci, sum);
    }
    __syncthreads();
  }
  if (dbias != nullptr && ci == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wred[tid * 8 + j] = bs[j];
    __syncthreads();
    for (int c = tid; c < CG * 8; c += blockDim.x) {
      float sum = 0.f;
      for (int r = 0; r < RL; ++r) sum += wred[r * (CG * 8) + c];
      atomicAdd(dbias + c, sum);
    }
  }
}

// ---- head ----------------------------------------------------------------------------------------
constexpr int kMaxCls = 16;
constexpr int kHeadTile = 256;  // pixels per block iteration (= blockDim)

// NC > 0: number of classes known at compile time (class loops fully unrolled, no predication);
// NC == 0: generic path for up to kMaxCls classes.
// A block iteration handles a tile of 256 pixels.  The tile's x rows are copied into shared memory COOPERATIVELY
// (consecutive threads fetch consecutive 16-byte pieces: x rows of one image line are contiguous in memory), then
// thread = pixel computes logits, softmax, loss and dlogits from its shared-memory row; TRAIN: thread = (channel pair,
// pixel segment) accumulates dWh[c][ci] += sum_p dlogit[p][c] * x[p][ci] from the tile, thread = pixel overwrites its
// row of the tile with dx = dlogit * Wh, and the tile is written back cooperatively.  (Round 1's version read and
// wrote the 192-byte rows per thread - 16-byte pieces 192 bytes apart across a warp - and ran at 1.9 TB/s.)
// tile_x == 0 (inference with very wide inputs whose tile does not fit shared memory): x is read per thread.
template <bool TRAIN, int NC>
__global__ void __launch_bounds__(kHeadTile)
    head_kernel(const __nv_bfloat16* __restrict__ x, Geo g, int C, const float* __restrict__ Wh,
                const float* __restrict__ bh, int ncls_rt, const uint8_t* __restrict__ labels,
                const float* __restrict__ sample_w, float grad_scale, __nv_bfloat16* __restrict__ dx,
                float* __restrict__ dWh, float* __restrict__ dbh, double* __restrict__ loss_sum,
                float* __restrict__ probs, int tile_x) {
  constexpr int MC = NC > 0 ? NC : kMaxCls;  // unrolled class loop bound
  constexpr int DLS = (MC + 3) & ~3;         // dlogit row stride (float4 reads)
  const int ncls = NC > 0 ? NC : ncls_rt;
  extern __shared__ float4 hsm4[];
  // smem: Wh [ncls][C] f32 | bh [DLS] | rows [256] i64 | (TRAIN) dl [256][DLS] f32 | xs [256][C+8] bf16
  float* sW = reinterpret_cast<float*>(hsm4);
  float* sB = sW + ((ncls * C + 3) & ~3);
  long long* rows_s = reinterpret_cast<long long*>(sB + DLS);
  float* dl = reinterpret_cast<float*>(rows_s + kHeadTile);
  const int XS = C + 8;  // row stride of the x tile: +16 B keeps the 16-byte row accesses off the same banks
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(dl + (TRAIN ? kHeadTile * DLS : 0));
  for (int i = threadIdx.x; i < ncls * C; i += blockDim.x) sW[i] = Wh[i];
  for (int i = threadIdx.x; i < DLS; i += blockDim.x) sB[i] = i < ncls ? bh[i] : 0.f;
  __syncthreads();

  const long long npix = g.pixels();
  const int Wp = g.W + 2;
  const int CG = C / 8;
  const int C2 = C / 2;                          // channel pairs
  const int nseg = max(1, kHeadTile / C2);       // pixel segments of the weight-gradient pass
  const int rows_per_seg = (kHeadTile + nseg - 1) / nseg;
  float wacc[2][MC];
  float bsum[MC];
#pragma unroll
  for (int c = 0; c < MC; ++c) wacc[0][c] = wacc[1][c] = bsum[c] = 0.f;
  double lacc = 0.0;
  const int hw = g.H * g.W;

  for (long long p0 = (long long)blockIdx.x * kHeadTile; p0 < npix; p0 += (long long)gridDim.x * kHeadTile) {
    const long long pix = p0 + threadIdx.x;
    const bool valid = pix < npix;
    const int nrows = (int)min((long long)kHeadTile, npix - p0);
    long long row = 0;
    int n = 0;
    if (valid) {
      // one division chain per tile and thread (tiles are 256x fewer than pixels)
      n = (int)(pix / hw);
      const int rem = (int)(pix - (long long)n * hw);
      const int yy = rem / g.W, xx = rem - yy * g.W;
      row = ((long long)n * (g.H + 2) + yy + 1) * Wp + xx + 1;
    }
    if (tile_x) {
      rows_s[threadIdx.x] = row;
      __syncthreads();
      // cooperative, coalesced copy of the tile's x rows
      for (int id = threadIdx.x; id < nrows * CG; id += kHeadTile) {
        const int px = id / CG, cgi = id - px * CG;
        *reinterpret_cast<uint4*>(xs + px * XS + cgi * 8) =
            *reinterpret_cast<const uint4*>(x + rows_s[px] * C + cgi * 8);
      }
      __syncthreads();
    }
    float d[MC];
    if (valid) {
      float z[MC];
#pragma unroll
      for (int c = 0; c < MC; ++c) z[c] = sB[c];
      for (int k8 = 0; k8 < C; k8 += 8) {
        const uint4 u = tile_x ? *reinterpret_cast<const uint4*>(xs + threadIdx.x * XS + k8)
                               : *reinterpret_cast<const uint4*>(x + row * C + k8);
        const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&u);
        float xv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(hv[j]);
          xv[2 * j] = f.x;
          xv[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int c = 0; c < MC; ++c) {
          if (NC > 0 || c < ncls) {
            const float4 w0 = *reinterpret_cast<const float4*>(sW + c * C + k8);
            const float4 w1 = *reinterpret_cast<const float4*>(sW + c * C + k8 + 4);
            float s = z[c];
            s = fmaf(xv[0], w0.x, s);
            s = fmaf(xv[1], w0.y, s);
            s = fmaf(xv[2], w0.z, s);
            s = fmaf(xv[3], w0.w, s);
            s = fmaf(xv[4], w1.x, s);
            s = fmaf(xv[5], w1.y, s);
            s = fmaf(xv[6], w1.z, s);
            s = fmaf(xv[7], w1.w, s);
            z[c] = s;
          }
        }
      }
      float mx = z[0];
#pragma unroll
      for (int c = 1; c < MC; ++c)
        if (NC > 0 || c < ncls) mx = fmaxf(mx, z[c]);
      float e[MC];
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < MC; ++c) {
        e[c] = (NC > 0 || c < ncls) ? expf(z[c] - mx) : 0.f;
        se += e[c];
      }
      const float inv = 1.f / se;
      if (probs) {
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (NC > 0 || c < ncls) probs[pix * ncls + c] = e[c] * inv;
      }
      if (TRAIN) {
        const int lab = labels[pix];
        const float w = sample_w ? sample_w[n] : 1.f;
        float zy = 0.f;
#pragma unroll
        for (int c = 0; c < MC; ++c)
          if (c == lab) zy = z[c];
        lacc += (double)((logf(se) + mx - zy) * w);
#pragma unroll
        for (int c = 0; c < MC; ++c) {
          d[c] = (NC > 0 || c < ncls) ? (e[c] * inv - (c == lab ? 1.f : 0.f)) * w * grad_scale : 0.f;
          bsum[c] += d[c];
          dl[threadIdx.x * DLS + c] = d[c];
        }
#pragma unroll
        for (int c = MC; c < DLS; ++c) dl[threadIdx.x * DLS + c] = 0.f;
      }
    }
    if (TRAIN) {
      __syncthreads();
      // dWh partial sums from the tile
      if (threadIdx.x < nseg * C2) {
        const int cp = threadIdx.x % C2, seg = threadIdx.x / C2;
        const int r_lo = seg * rows_per_seg, r_hi = min(nrows, r_lo + rows_per_seg);
        for (int r = r_lo; r < r_hi; ++r) {
          const float2 xv =
              __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xs + r * XS + 2 * cp));
          float dv[DLS];
#pragma unroll
          for (int q = 0; q < DLS / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(dl + r * DLS + 4 * q);
            dv[4 * q] = t.x;
            dv[4 * q + 1] = t.y;
            dv[4 * q + 2] = t.z;
            dv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int c = 0; c < MC; ++c) {
            wacc[0][c] = fmaf(dv[c], xv.x, wacc[0][c]);
            wacc[1][c] = fmaf(dv[c], xv.y, wacc[1][c]);
          }
        }
      }
      __syncthreads();
      // dx[ci] = sum_c dlogit_c * Wh[c][ci], written over the thread's own x row of the tile
      if (valid) {
        for (int k8 = 0; k8 < C; k8 += 8) {
          Vec8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
#pragma unroll
          for (int c = 0; c < MC; ++c) {
            if (NC > 0 || c < ncls) {
              const float4 w0 = *reinterpret_cast<const float4*>(sW + c * C + k8);
              const float4 w1 = *reinterpret_cast<const float4*>(sW + c * C + k8 + 4);
              o.v[0] = fmaf(d[c], w0.x, o.v[0]);
              o.v[1] = fmaf(d[c], w0.y, o.v[1]);
              o.v[2] = fmaf(d[c], w0.z, o.v[2]);
              o.v[3] = fmaf(d[c], w0.w, o.v[3]);
              o.v[4] = fmaf(d[c], w1.x, o.v[4]);
              o.v[5] = fmaf(d[c], w1.y, o.v[5]);
              o.v[6] = fmaf(d[c], w1.z, o.v[6]);
              o.v[7] = fmaf(d[c], w1.w, o.v[7]);
            }
          }
          store8(xs + threadIdx.x * XS + k8, o);
        }
      }
      __syncthreads();
      for (int id = threadIdx.x; id < nrows * CG; id += kHeadTile) {
        const int px = id / CG, cgi = id - px * CG;
        *reinterpret_cast<uint4*>(dx + rows_s[px] * C + cgi * 8) =
            *reinterpret_cast<const uint4*>(xs + px * XS + cgi * 8);
      }
      __syncthreads();  // the tile and the row table are rewritten by the next iteration
    }
  }
  if (TRAIN) {
    if (threadIdx.x < nseg * C2) {
      const int cp = threadIdx.x % C2;
#pragma unroll
      for (int c = 0; c < MC; ++c)
        if (NC > 0 || c < ncls) {
          atomicAdd(dWh + c * C + 2 * cp, wacc[0][c]);
          atomicAdd(dWh + c * C + 2 * cp + 1, wacc[1][c]);
        }
    }
#pragma unroll
    for (int c = 0; c < MC; ++c) {
      if (NC > 0 || c < ncls) {
        float v = bsum[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(dbh + c, v);
      }
    }
    // block reduce loss
    __shared__ double lred[kHeadTile];
    lred[threadIdx.x] = lacc;
    __syncthreads();
    for (int s2 = kHeadTile / 2; s2 > 0; s2 >>= 1) {
      if (threadIdx.x < s2) lred[threadIdx.x] += lred[threadIdx.x + s2];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(loss_sum, lred[0]);
  }
}

// ---- optimizer / weight layout ---------------------------------------------------------------------
__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float lr_t, float b1, float b2,
                                          float eps, float gscale) {
  const float gi = g * gscale;
  m = b1 * m + (1.f - b1) * gi;
  v = b2 * v + (1.f - b2) * gi * gi;
  return p - lr_t * m / (sqrtf(v) + eps);
}

// 16-byte accesses on the aligned body ([head, head + 4 * nvec)), scalar on the at most 3 + 3 edge elements
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            float gscale, __nv_bfloat16* __restrict__ shadow) {
  const long long head = min(n, (long long)((4 - ((reinterpret_cast<uintptr_t>(p) >> 2) & 3)) & 3));
  const long long nvec = (n - head) >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < nvec; i += nth) {
    const long long e = head + 4 * i;
    float4 pv = *reinterpret_cast<const float4*>(p + e);
    const float4 gv = *reinterpret_cast<const float4*>(g + e);
    float4 mv = *reinterpret_cast<const float4*>(m + e);
    float4 vv = *reinterpret_cast<const float4*>(v + e);
    pv.x = adam_one(pv.x, gv.x, mv.x, vv.x, lr_t, b1, b2, eps, gscale);
    pv.y = adam_one(pv.y, gv.y, mv.y, vv.y, lr_t, b1, b2, eps, gscale);
    pv.z = adam_one(pv.z, gv.z, mv.z, vv.z, lr_t, b1, b2, eps, gscale);
    pv.w = adam_one(pv.w, gv.w, mv.w, vv.w, lr_t, b1, b2, eps, gscale);
    *reinterpret_cast<float4*>(p + e) = pv;
    *reinterpret_cast<float4*>(m + e) = mv;
    *reinterpret_cast<float4*>(v + e) = vv;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + e) = pk;
    }
  }
  const long long edge = head + (n - head - 4 * nvec);  // scalar elements: [0, head) and [head + 4 nvec, n)
  for (long long k = tid; k < edge; k += nth) {
    const long long i = k < head ? k : head + 4 * nvec + (k - head);
    float mi = m[i], vi = v[i];
    const float pn = adam_one(p[i], g[i], mi, vi, lr_t, b1, b2, eps, gscale);
    m[i] = mi;
    v[i] = vi;
    p[i] = pn;
    if (shadow) shadow[i] = __float2bfloat16_rn(pn);
  }
}

// Keras kernel_regularizer=l2(lambda): g += coef * w on one conv kernel, and sum w^2 (warp -> block -> one double atomic)
__global__ void l2_penalty_kernel(const float* __restrict__ w, float* __restrict__ g, long long n, float coef,
                                  double* __restrict__ sumsq) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float wi = w[i];
    g[i] += coef * wi;
    s += wi * wi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += (double)ws[k];
    atomicAdd(sumsq, t);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ p, __nv_bfloat16* __restrict__ shadow, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    shadow[i] = __float2bfloat16_rn(p[i]);
}

__global__ void prep_conv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                 __nv_bfloat16* __restrict__ wd, int ntap, int co, int k, int flip) {
  const long long total = (long long)ntap * co * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % k);
    long long t = i / k;
    const int c = (int)(t % co);
    const int tap = (int)(t / co);
    const __nv_bfloat16 b = __float2bfloat16_rn(w[i]);
    if (wf) wf[i] = b;
    if (wd) {
      const int td = flip ? ntap - 1 - tap : tap;
      wd[((long long)td * k + kk) * co + c] = b;
    }
  }
}

__device__ __forceinline__ float collapsed_weight(const float* w, int pair, int c, int kk, int co, int k) {
  const UpPair p = up_pair(pair);
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
      if (((p.a + dy) >> 1) == p.di && ((p.b + dx) >> 1) == p.dj)
        s += w[((long long)(dy * 2 + dx) * co + c) * k + kk];
  return s;
}

__global__ void prep_upconv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, int co, int k) {
  const long long total = 9ll * co * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % k);
    long long t = i / k;
    const int c = (int)(t % co);
    const int pair = (int)(t / co);
    const __nv_bfloat16 b = __float2bfloat16_rn(collapsed_weight(w, pair, c, kk, co, k));
    wf[i] = b;
    if (wd) wd[((long long)pair * k + kk) * co + c] = b;
  }
}

__global__ void fold_upconv_grad_kernel(const float* __restrict__ dwc, float* __restrict__ dw, int co,
                                        int k) {
  const long long per = (long long)co * k;
  const long long total = 4 * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i / per);
    const long long rem = i % per;
    const int dy = tap >> 1, dx = tap & 1;
    float s = 0.f;
#pragma unroll
    for (int pair = 0; pair < 9; ++pair) {
      const UpPair p = up_pair(pair);
      if (((p.a + dy) >> 1) == p.di && ((p.b + dx) >> 1) == p.dj) s += dwc[(long long)pair * per + rem];
    }
    dw[i] += s;
  }
}

__global__ void pack_input_kernel(const float* __restrict__ x, Geo g, int cin, int cin_phys,
                                  __nv_bfloat16* __restrict__ out) {
  const long long total = g.pixels();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = padded_row(g, i);
    for (int c = 0; c < cin_phys; ++c)
      out[row * cin_phys + c] = __float2bfloat16_rn(c < cin ? x[i * cin + c] : 0.f);
  }
}

inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  if (b > kMaxBlocks) b = kMaxBlocks;
  if (b < 1) b = 1;
  return (int)b;
}

// Grid of a grid-stride kernel = the number of blocks that are resident at once (no partial last wave),
// capped by the number of work items.
template <typename K>
int resident_grid(K kernel, int threads, size_t smem, long long items) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 1;
  }
  long long b = 148ll * occ;
  if (b > items) b = items;
  if (b < 1) b = 1;
  return (int)b;
}

// thread layout of the per-pixel kernels: blockDim = CG * RL with RL a power of two (row lanes), at most
// `span` pixels (or windows) per line segment
inline void line_layout(int C, int span, int* CG, int* RL, int* threads) {
  *CG = C / 8;
  int rl = 1;
  while (2 * rl * *CG <= 256 && 2 * rl <= span) rl *= 2;
  *RL = rl;
  *threads = *CG * rl;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int launch_channel_stats(const __nv_bfloat16* y, long long rows, int C, int ld, double* sums,
                         cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(C, &CG, &RL, &threads);
  MPU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  const size_t smem = sizeof(float) * 2 * RL * CG * 8;
  const int grid = resident_grid(channel_stats_kernel, threads, smem, (rows + 4 * RL - 1) / (4 * RL));
  channel_stats_kernel<<<grid, threads, smem, st>>>(y, rows, C, ld, CG, RL, sums);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_colsum(const __nv_bfloat16* x, long long rows, int C, int ld, float* out, cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(C, &CG, &RL, &threads);
  const size_t smem = sizeof(float) * RL * CG * 8;
  const int grid = resident_grid(colsum_kernel, threads, smem, (rows + 4 * RL - 1) / (4 * RL));
  colsum_kernel<<<grid, threads, smem, st>>>(x, rows, C, ld, CG, RL, out);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                       float* moving_mean, float* moving_var, float eps, float momentum, int training,
                       int C, float* scale, float* shift, float* mean, float* rstd, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, count, gamma, beta, moving_mean, moving_var,
                                                     eps, momentum, training, C, scale, shift, mean,
                                                     rstd);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

// bring-up knob for same-box A/B measurements: MPU_BN_SEPARATE=1 runs bn_finalize / bn_bwd_params as the separate
// single-block launches they were before being folded into the apply kernels
static int bn_separate() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPU_BN_SEPARATE");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int launch_bn_apply(const __nv_bfloat16* y, const BnFin& f0, __nv_bfloat16* b,
                    __nv_bfloat16* pooled, Geo g, int C, cudaStream_t st) {
  BnFin f = f0;
  f.no_write = 0;
  if (bn_separate()) {
    MPU_TRY(launch_bn_finalize(f.sums, f.count, f.gamma, f.beta, f.mmean, f.mvar, f.eps, f.momentum, f.training, f.C,
                               f.scale, f.shift, f.mean_out, f.rstd_out, st));
    f.no_write = 1;   // (the apply kernel must then read the stored coefficients: the moving statistics have moved on)
  }
  int CG, RL, threads;
  if (C / 8 > 256) {
    set_error("bn_apply: C=%d exceeds 2048 channels", C);
    return MPU_ERR_ARG;
  }
  if (pooled) {
    line_layout(C, g.W / 2, &CG, &RL, &threads);
    const long long items = (long long)g.B * (g.H / 2) * ((g.W / 2 + RL - 1) / RL);
    const int grid = resident_grid(bn_apply_pool_kernel, threads, 0, items);
    bn_apply_pool_kernel<<<grid, threads, 0, st>>>(y, f, b, pooled, g, C, CG, RL);
  } else {
    line_layout(C, (g.W + 3) / 4, &CG, &RL, &threads);
    const long long items = (long long)g.B * g.H * ((g.W + 4 * RL - 1) / (4 * RL));
    const int grid = resident_grid(bn_apply_kernel, threads, 0, items);
    bn_apply_kernel<<<grid, threads, 0, st>>>(y, f, b, g, C, CG, RL);
  }
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

namespace {
// pass selection + launch of bn_bwd_kernel<POOL, APPLY>
template <bool POOL, bool APPLY>
int launch_bn_bwd(const BnBwdArgs& a, const double* sums_in, double* sums_out, __nv_bfloat16* dz,
                  int phase_major, float* dbias, cudaStream_t st) {
  int CG, RL, threads;
  const int span = POOL ? a.g.W / 2 : (a.g.W + 1) / 2;
  const int V = 8;  // channels per thread (see the kernel)
  line_layout(a.C * 8 / V, span, &CG, &RL, &threads);
  const int U = POOL ? 1 : 2;
  const long long items =
      (long long)a.g.B * (POOL ? a.g.H / 2 : a.g.H) * (((POOL ? a.g.W / 2 : a.g.W) + U * RL - 1) / (U * RL));
  const size_t smem = sizeof(float) * (APPLY ? 1 : 2) * RL * CG * V;
  const int grid = resident_grid(bn_bwd_kernel<POOL, APPLY>, threads, smem, items);
  bn_bwd_kernel<POOL, APPLY><<<grid, threads, smem, st>>>(a, CG, RL, sums_in, sums_out, dz, phase_major, dbias);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}
}  // namespace

int launch_bn_bwd_reduce(const BnBwdArgs& a, double* sums, cudaStream_t st) {
  if (a.C / 4 > 256 && a.gP != nullptr) {
    set_error("bn_bwd: C=%d exceeds 1024 channels on a pooled level", a.C);
    return MPU_ERR_ARG;
  }
  if (a.C / 8 > 256) {
    set_error("bn_bwd: C=%d exceeds 2048 channels", a.C);
    return MPU_ERR_ARG;
  }
  MPU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * a.C, st));
  if (a.gP != nullptr) return launch_bn_bwd<true, false>(a, nullptr, sums, nullptr, 0, nullptr, st);
  return launch_bn_bwd<false, false>(a, nullptr, sums, nullptr, 0, nullptr, st);
}

int launch_bn_bwd_apply(const BnBwdArgs& a0, const double* sums, __nv_bfloat16* dz, int phase_major,
                        float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  BnBwdArgs a = a0;  // the parameter gradients are added by block 0 of the apply kernel itself
  const bool sep = bn_separate() != 0;
  a.dgamma = sep ? nullptr : dgamma;
  a.dbeta = sep ? nullptr : dbeta;
  if (a.gP != nullptr)
    MPU_TRY((launch_bn_bwd<true, true>(a, sums, nullptr, dz, phase_major, dbias, st)));
  else
    MPU_TRY((launch_bn_bwd<false, true>(a, sums, nullptr, dz, phase_major, dbias, st)));
  if (sep) {
    bn_bwd_params_kernel<<<(a.C + 127) / 128, 128, 0, st>>>(sums, a.mean, a.rstd, a.C, dgamma, dbeta);
    count_launch();
    MPU_CUDA(cudaGetLastError());
  }
  return MPU_OK;
}

int launch_conv_first(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out,
                      Geo g, int cin, int co_phys, cudaStream_t st) {
  if (cin < 1 || cin > 4 || co_phys % 8) {
    set_error("conv_first: cin=%d co_phys=%d unsupported", cin, co_phys);
    return MPU_ERR_ARG;
  }
  const size_t smem = sizeof(float) * (9 * cin * co_phys + co_phys);
  if (smem > 48 * 1024) {
    set_error("conv_first: weights do not fit in shared memory");
    return MPU_ERR_ARG;
  }
  if (co_phys / 8 > 256) {
    set_error("conv_first: co_phys=%d exceeds 2048 channels", co_phys);
    return MPU_ERR_ARG;
  }
  int CG, RL, threads;
  line_layout(co_phys, g.W, &CG, &RL, &threads);
  const long long items = (long long)g.B * ((g.H + kFirstRows - 1) / kFirstRows) * ((g.W + RL - 1) / RL);
#define MPU_CONV_FIRST(CIN)                                                                        \
  conv_first_kernel<CIN><<<resident_grid(conv_first_kernel<CIN>, threads, smem, items), threads, smem, st>>>( \
      x, w, bias, out, g, co_phys, CG, RL)
  switch (cin) {
    case 1: MPU_CONV_FIRST(1); break;
    case 2: MPU_CONV_FIRST(2); break;
    case 3: MPU_CONV_FIRST(3); break;
    default: MPU_CONV_FIRST(4); break;
  }
#undef MPU_CONV_FIRST
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_conv_first_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dz, Geo g, int cin, int co_phys,
                            float* dW, int ldw, float* dbias, cudaStream_t st) {
  if (cin < 1 || cin > 8 || co_phys % 8 || co_phys / 8 > 256) {
    set_error("conv_first_wgrad: cin=%d co_phys=%d unsupported", cin, co_phys);
    return MPU_ERR_ARG;
  }
  int CG, RL, threads;
  line_layout(co_phys, (g.W + 3) / 4, &CG, &RL, &threads);
  const size_t smem = sizeof(float) * threads * 8;
  const long long items = (long long)g.B * g.H * ((g.W + 4 * RL - 1) / (4 * RL));
  const int grid = resident_grid(conv_first_wgrad_kernel, threads, smem, items);
  conv_first_wgrad_kernel<<<dim3(grid, cin), threads, smem, st>>>(x, dz, g, co_phys, CG, RL, dW, ldw, dbias);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

namespace {
size_t head_smem(int C, int ncls, int mc, bool train, bool tile_x) {
  const int dls = (mc + 3) & ~3;
  size_t s = sizeof(float) * ((size_t)((ncls * C + 3) & ~3) + dls) + sizeof(long long) * kHeadTile;
  if (train) s += sizeof(float) * kHeadTile * dls;
  if (tile_x) s += (size_t)kHeadTile * (C + 8) * 2;
  return s;
}

template <bool TRAIN, int NC>
int launch_head(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                const uint8_t* labels, const float* sample_w, float grad_scale, __nv_bfloat16* dx,
                float* dWh, float* dbh, double* loss_sum, float* probs, cudaStream_t st) {
  // inference with a very wide input: the x tile does not fit shared memory, rows are read per thread
  const bool tile_x = TRAIN || head_smem(C, ncls, NC > 0 ? NC : kMaxCls, TRAIN, true) <= 96 * 1024;
  const size_t smem = head_smem(C, ncls, NC > 0 ? NC : kMaxCls, TRAIN, tile_x);
  if (smem > 200 * 1024) {
    set_error("head: shared memory %zu too large (C=%d, n_classes=%d)", smem, C, ncls);
    return MPU_ERR_ARG;
  }
  static bool attr[64] = {false};  // one flag per instantiation and device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (!attr[dev]) {
    MPU_CUDA(cudaFuncSetAttribute(head_kernel<TRAIN, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    attr[dev] = true;
  }
  const long long tiles = (g.pixels() + kHeadTile - 1) / kHeadTile;
  const int grid = resident_grid(head_kernel<TRAIN, NC>, kHeadTile, smem, tiles);
  head_kernel<TRAIN, NC><<<grid, kHeadTile, smem, st>>>(x, g, C, Wh, bh, ncls, labels, sample_w, grad_scale,
                                                       dx, dWh, dbh, loss_sum, probs, tile_x ? 1 : 0);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

template <bool TRAIN>
int dispatch_head(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                  const uint8_t* labels, const float* sample_w, float grad_scale, __nv_bfloat16* dx,
                  float* dWh, float* dbh, double* loss_sum, float* probs, cudaStream_t st) {
#define MPU_HEAD(NC) \
  return launch_head<TRAIN, NC>(x, g, C, Wh, bh, ncls, labels, sample_w, grad_scale, dx, dWh, dbh, loss_sum, probs, st)
  switch (ncls) {
    case 2: MPU_HEAD(2);
    case 3: MPU_HEAD(3);
    case 4: MPU_HEAD(4);
    case 5: MPU_HEAD(5);
    case 6: MPU_HEAD(6);
    case 7: MPU_HEAD(7);
    case 8: MPU_HEAD(8);
    default: MPU_HEAD(0);
  }
#undef MPU_HEAD
}
}  // namespace

int launch_head_infer(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                      float* probs, cudaStream_t st) {
  if (ncls < 1 || ncls > kMaxCls) {
    set_error("head: n_classes=%d outside [1,%d]", ncls, kMaxCls);
    return MPU_ERR_ARG;
  }
  return dispatch_head<false>(x, g, C, Wh, bh, ncls, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr,
                              probs, st);
}

int launch_head_train(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                      const uint8_t* labels, const float* sample_w, float grad_scale,
                      __nv_bfloat16* dx, float* dWh, float* dbh, double* loss_sum, float* probs_opt,
                      cudaStream_t st) {
  if (ncls < 1 || ncls > kMaxCls || C > 256) {
    set_error("head(train): n_classes=%d, C=%d not supported (max %d classes, C <= 256)", ncls, C, kMaxCls);
    return MPU_ERR_ARG;
  }
  return dispatch_head<true>(x, g, C, Wh, bh, ncls, labels, sample_w, grad_scale, dx, dWh, dbh, loss_sum,
                             probs_opt, st);
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1,
                float b2, float eps, float gscale, __nv_bfloat16* shadow, cudaStream_t st) {
  if (n <= 0) return MPU_OK;
  adam_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, gscale, shadow);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_l2_penalty(const float* w, float* g, long long n, float coef, double* sumsq, cudaStream_t st) {
  if (n <= 0) return MPU_OK;
  l2_penalty_kernel<<<grid_for(n, 256), 256, 0, st>>>(w, g, n, coef, sumsq);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_cast_bf16(const float* p, __nv_bfloat16* shadow, long long n, cudaStream_t st) {
  cast_bf16_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, shadow, n);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_prep_conv(const float* w, __nv_bfloat16* wf, __nv_bfloat16* wd, int ntap, int co, int k,
                     int flip, cudaStream_t st) {
  prep_conv_kernel<<<grid_for((long long)ntap * co * k, 256), 256, 0, st>>>(w, wf, wd, ntap, co, k, flip);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_prep_upconv(const float* w, __nv_bfloat16* wf, __nv_bfloat16* wd, int co, int k,
                       cudaStream_t st) {
  prep_upconv_kernel<<<grid_for(9ll * co * k, 256), 256, 0, st>>>(w, wf, wd, co, k);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_pack_input(const float* x, Geo g, int cin, int cin_phys, __nv_bfloat16* out, cudaStream_t st) {
  pack_input_kernel<<<grid_for(g.pixels(), 256), 256, 0, st>>>(x, g, cin, cin_phys, out);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_fold_upconv_grad(const float* dwc, float* dw, int co, int k, cudaStream_t st) {
  fold_upconv_grad_kernel<<<grid_for(4ll * co * k, 256), 256, 0, st>>>(dwc, dw, co, k);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

}  // namespace mpu
