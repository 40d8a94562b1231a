// See elementwise.cuh.  HBM-bound CUDA-core kernels; 16-byte vector access, channel-group-per-thread.
#include "elementwise.cuh"
#include "common.h"

#include <cstdio>

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
__global__ void channel_stats_kernel(const __nv_bfloat16* __restrict__ y, long long rows, int C, int ld,
                                     int CG, int RL, double* __restrict__ sums) {
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  for (long long r = (long long)blockIdx.x * RL + rl; r < rows; r += (long long)gridDim.x * RL) {
    const Vec8 x = load8(y + r * ld + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += x.v[j];
      acc[1][j] += x.v[j] * x.v[j];
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
  for (long long r = (long long)blockIdx.x * RL + rl; r < rows; r += (long long)gridDim.x * RL) {
    const Vec8 v = load8(x + r * ld + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += v.v[j];
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

// interior pixel index -> padded row
__device__ __forceinline__ long long padded_row(const Geo& g, long long pix) {
  const int x = (int)(pix % g.W);
  const long long t = pix / g.W;
  const int yy = (int)(t % g.H);
  const long long n = t / g.H;
  return n * (long long)(g.H + 2) * (g.W + 2) + (long long)(yy + 1) * (g.W + 2) + (x + 1);
}

__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, __nv_bfloat16* __restrict__ b, Geo g,
                                int C) {
  const int CG = C / 8;
  const long long total = g.pixels() * CG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const long long row = padded_row(g, i / CG);
    Vec8 v = load8(y + row * C + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], scale[cg * 8 + j], shift[cg * 8 + j]);
    store8(b + row * C + cg * 8, v);
  }
}

// window (2x2) version: writes the four BN outputs and their max into the half-resolution tensor.
__global__ void bn_apply_pool_kernel(const __nv_bfloat16* __restrict__ y,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ pooled,
                                     Geo g, int C) {
  const int CG = C / 8;
  const int h = g.H / 2, w = g.W / 2;
  const long long total = (long long)g.B * h * w * CG;
  const int Wp = g.W + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    long long t = i / CG;
    const int jx = (int)(t % w);
    t /= w;
    const int iy = (int)(t % h);
    const long long n = t / h;
    const long long base = n * (long long)(g.H + 2) * Wp + (long long)(2 * iy + 1) * Wp + (2 * jx + 1);
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale[cg * 8 + j];
      sh[j] = shift[cg * 8 + j];
    }
    Vec8 mx;
#pragma unroll
    for (int j = 0; j < 8; ++j) mx.v[j] = -INFINITY;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const long long row = base + (d >> 1) * Wp + (d & 1);
      Vec8 v = load8(y + row * C + cg * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v.v[j] = bf16_round(fmaf(v.v[j], sc[j], sh[j]));
        mx.v[j] = fmaxf(mx.v[j], v.v[j]);
      }
      store8(b + row * C + cg * 8, v);
    }
    const long long prow = n * (long long)(h + 2) * (w + 2) + (long long)(iy + 1) * (w + 2) + (jx + 1);
    store8(pooled + prow * C + cg * 8, mx);
  }
}

// ---- BN backward ---------------------------------------------------------------------------------
// Work item = (pixel or 2x2 window, channel group).  g(p,c) = gA[p] + (p is the first arg-max of its
// window ? gP[window] : 0).  Window arg-max is recomputed from the bf16-rounded BN outputs, scanning
// row-major so the first maximum wins (the forward max-pool kept no index).
// Pass 1 accumulates sum(g) and sum(g*y); with xhat = (y-mu)*rstd the BN backward is then the per-channel
// affine map dz = relu'(y) * (a*g + b*y + c):  a = gamma*rstd, b = -a*rstd*mgx, c = -a*mg - b*mu  where
// mg = mean(g), mgx = mean(g*xhat) = rstd*(mean(g*y) - mu*mg).  Few per-thread coefficients keep the
// register count low enough for 4+ resident blocks per SM (these kernels are pure HBM streams).
template <bool POOL, bool APPLY>
__global__ void __launch_bounds__(256, POOL ? 2 : 3)
    bn_bwd_kernel(BnBwdArgs a, int CG, int RL, const double* __restrict__ sums_in,
                  double* __restrict__ sums_out, __nv_bfloat16* __restrict__ dz, int phase_major,
                  float* __restrict__ dbias) {
  const int tid = threadIdx.x;
  const int cg = tid % CG, rl = tid / CG;
  const int C = a.C;
  const Geo g = a.g;
  const int Wp = g.W + 2;
  const int hh = g.H / 2, wh = g.W / 2;
  const long long items = POOL ? (long long)g.B * hh * wh : g.pixels();
  float sc[POOL ? 8 : 1], sh[POOL ? 8 : 1];
  float ca[APPLY ? 8 : 1], cb[APPLY ? 8 : 1], cc[APPLY ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    if (POOL) {
      sc[j] = a.scale[c];
      sh[j] = a.shift[c];
    }
    if (APPLY) {
      const double cnt = (double)g.pixels();
      const float mu = a.mean[c], rs = a.rstd[c];
      const float mg = (float)(sums_in[c] / cnt);
      const float mgy = (float)(sums_in[C + c] / cnt);
      const float mgx = rs * (mgy - mu * mg);
      const float k1 = a.gamma[c] * rs;
      ca[j] = k1;
      cb[j] = -k1 * rs * mgx;
      cc[j] = -k1 * mg - cb[j] * mu;
    }
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;

  for (long long it = (long long)blockIdx.x * RL + rl; it < items; it += (long long)gridDim.x * RL) {
    long long rows4[4];
    int nsub;
    long long prow = 0;
    if (POOL) {
      long long t = it;
      const int jx = (int)(t % wh);
      t /= wh;
      const int iy = (int)(t % hh);
      const long long n = t / hh;
      const long long base = n * (long long)(g.H + 2) * Wp + (long long)(2 * iy + 1) * Wp + (2 * jx + 1);
      rows4[0] = base;
      rows4[1] = base + 1;
      rows4[2] = base + Wp;
      rows4[3] = base + Wp + 1;
      nsub = 4;
      prow = n * (long long)(hh + 2) * (wh + 2) + (long long)(iy + 1) * (wh + 2) + (jx + 1);
    } else {
      rows4[0] = padded_row(g, it);
      nsub = 1;
    }
    Vec8 yv[POOL ? 4 : 1];
    int amax[8];
    if (POOL) {
      float best[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        best[j] = -INFINITY;
        amax[j] = 0;
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        yv[d] = load8(a.y + rows4[d] * C + cg * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float bv = bf16_round(fmaf(yv[d].v[j], sc[j], sh[j]));
          if (bv > best[j]) {
            best[j] = bv;
            amax[j] = d;
          }
        }
      }
    } else {
      yv[0] = load8(a.y + rows4[0] * C + cg * 8);
    }
    Vec8 gp;
    if (POOL && a.gP) gp = load8(a.gP + prow * C + cg * 8);
#pragma unroll
    for (int d = 0; d < (POOL ? 4 : 1); ++d) {
      if (d < nsub) {
        Vec8 gv;
        if (a.gA) {
          gv = load8(a.gA + rows4[d] * (long long)a.ldA + cg * 8);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) gv.v[j] = 0.f;
        }
        if (POOL && a.gP) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (amax[j] == d) gv.v[j] += gp.v[j];
        }
        if (!APPLY) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0][j] += gv.v[j];
            acc[1][j] = fmaf(gv.v[j], yv[d].v[j], acc[1][j]);
          }
        } else {
          Vec8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = fmaf(ca[j], gv.v[j], fmaf(cb[j], yv[d].v[j], cc[j]));
            if (!(yv[d].v[j] > 0.f)) v = 0.f;
            v = bf16_round(v);
            o.v[j] = v;
            acc[0][j] += v;
          }
          long long orow = rows4[d];
          if (phase_major) {
            // pixel (n, y, x) of this level -> [phase][rows of the half-resolution level]
            const long long plane = (long long)(g.H + 2) * Wp;
            const long long n = orow / plane;
            const long long rem = orow - n * plane;
            const int yy = (int)(rem / Wp) - 1, xx = (int)(rem % Wp) - 1;
            const int ph = (yy & 1) * 2 + (xx & 1);
            const long long rows_lo = (long long)g.B * (hh + 2) * (wh + 2);
            orow = ph * rows_lo + n * (long long)(hh + 2) * (wh + 2) +
                   (long long)((yy >> 1) + 1) * (wh + 2) + ((xx >> 1) + 1);
          }
          store8(dz + orow * C + cg * 8, o);
        }
      }
    }
  }
  if (!APPLY) {
    block_reduce_store<2, double>(acc, CG, RL, C, sums_out);
  } else if (dbias) {
    float acc1[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc1[0][j] = acc[0][j];
    block_reduce_store<1, float>(acc1, CG, RL, C, dbias);
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
// thread = (pixel, 8-channel output group); the 9 x cin inputs of the pixel sit in registers, weights in
// shared memory as fp32 [9][cin][co_phys]
__global__ void conv_first_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, Geo g,
                                  int cin, int co_phys) {
  extern __shared__ float cw[];  // [9][cin][co_phys] + bias[co_phys]
  for (int i = threadIdx.x; i < 9 * cin * co_phys; i += blockDim.x) {
    const int co = i % co_phys;
    const int t2 = i / co_phys;
    const int ci = t2 % cin, tap = t2 / cin;
    cw[i] = __bfloat162float(w[((long long)tap * co_phys + co) * 8 + ci]);
  }
  float* sb = cw + 9 * cin * co_phys;
  for (int i = threadIdx.x; i < co_phys; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int CG = co_phys / 8;
  const int Wp = g.W + 2;
  const long long total = g.pixels() * CG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const long long row = padded_row(g, i / CG);
    Vec8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = sb[cg * 8 + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const long long r = row + (tap / 3 - 1) * Wp + (tap % 3 - 1);
      const Vec8 xv = load8(x + r * 8);
      for (int ci = 0; ci < cin; ++ci) {
        const float xs = xv.v[ci];
        const float* wr = cw + (tap * cin + ci) * co_phys + cg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(xs, wr[j], acc.v[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = fmaxf(acc.v[j], 0.f);
    store8(out + row * co_phys + cg * 8, acc);
  }
}

// ---- head ----------------------------------------------------------------------------------------
constexpr int kMaxCls = 16;

template <bool TRAIN>
__global__ void head_kernel(const __nv_bfloat16* __restrict__ x, Geo g, int C,
                            const float* __restrict__ Wh, const float* __restrict__ bh, int ncls,
                            const uint8_t* __restrict__ labels, const float* __restrict__ sample_w,
                            float grad_scale, __nv_bfloat16* __restrict__ dx, float* __restrict__ dWh,
                            float* __restrict__ dbh, double* __restrict__ loss_sum,
                            float* __restrict__ probs) {
  extern __shared__ unsigned char hsm[];
  // smem: Wh [ncls][C] f32 | bh [ncls] | (TRAIN) xs [256][C] bf16 | dl [256][ncls] f32
  float* sW = reinterpret_cast<float*>(hsm);
  float* sB = sW + ncls * C;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(sB + ((ncls + 3) & ~3));
  float* dl = reinterpret_cast<float*>(xs + (TRAIN ? 256 * C : 0));
  for (int i = threadIdx.x; i < ncls * C; i += blockDim.x) sW[i] = Wh[i];
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) sB[i] = bh[i];
  __syncthreads();

  const long long npix = g.pixels();
  const int nseg = max(1, 256 / C);                  // pixel segments for the weight-gradient pass
  const int rows_per_seg = (256 + nseg - 1) / nseg;
  float wacc[kMaxCls];
#pragma unroll
  for (int c = 0; c < kMaxCls; ++c) wacc[c] = 0.f;
  float bacc = 0.f;
  double lacc = 0.0;

  for (long long p0 = (long long)blockIdx.x * 256; p0 < npix; p0 += (long long)gridDim.x * 256) {
    const long long pix = p0 + threadIdx.x;
    const bool valid = pix < npix;
    float z[kMaxCls];
    long long row = 0;
    if (valid) {
      row = padded_row(g, pix);
#pragma unroll
      for (int c = 0; c < kMaxCls; ++c) z[c] = c < ncls ? sB[c] : 0.f;
      for (int k8 = 0; k8 < C; k8 += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + k8);
        if (TRAIN) *reinterpret_cast<uint4*>(xs + threadIdx.x * C + k8) = u;
        const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&u);
        float xv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(hv[j]);
          xv[2 * j] = f.x;
          xv[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c) {
          if (c < ncls) {
            float s = z[c];
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(xv[j], sW[c * C + k8 + j], s);
            z[c] = s;
          }
        }
      }
      float mx = z[0];
#pragma unroll
      for (int c = 1; c < kMaxCls; ++c)
        if (c < ncls) mx = fmaxf(mx, z[c]);
      float e[kMaxCls];
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxCls; ++c) {
        e[c] = c < ncls ? expf(z[c] - mx) : 0.f;
        se += e[c];
      }
      const float inv = 1.f / se;
      if (probs) {
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c)
          if (c < ncls) probs[pix * ncls + c] = e[c] * inv;
      }
      if (TRAIN) {
        const int lab = labels[pix];
        const int n = (int)(pix / ((long long)g.H * g.W));
        const float w = sample_w ? sample_w[n] : 1.f;
        float zy = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c)
          if (c == lab) zy = z[c];
        lacc += (double)((logf(se) + mx - zy) * w);
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c)
          if (c < ncls) dl[threadIdx.x * ncls + c] = (e[c] * inv - (c == lab ? 1.f : 0.f)) * w * grad_scale;
        // dx[ci] = sum_c dlogit_c * Wh[c][ci]
        for (int k8 = 0; k8 < C; k8 += 8) {
          Vec8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
#pragma unroll
          for (int c = 0; c < kMaxCls; ++c) {
            if (c < ncls) {
              const float d = dl[threadIdx.x * ncls + c];
#pragma unroll
              for (int j = 0; j < 8; ++j) o.v[j] = fmaf(d, sW[c * C + k8 + j], o.v[j]);
            }
          }
          store8(dx + row * C + k8, o);
        }
      }
    } else if (TRAIN) {
      for (int c = 0; c < ncls; ++c) dl[threadIdx.x * ncls + c] = 0.f;
    }
    if (TRAIN) {
      __syncthreads();
      // dWh[c][ci] += sum_p dl[p][c] * x[p][ci]: thread = (input channel ci, pixel segment); all classes
      // accumulate in registers, one x read + ncls broadcast reads per pixel
      const int nrows = (int)min((long long)256, npix - p0);
      if (threadIdx.x < nseg * C) {
        const int ci = threadIdx.x % C, seg = threadIdx.x / C;
        const int r_lo = seg * rows_per_seg, r_hi = min(nrows, r_lo + rows_per_seg);
        for (int r = r_lo; r < r_hi; ++r) {
          const float xv = __bfloat162float(xs[r * C + ci]);
#pragma unroll
          for (int c = 0; c < kMaxCls; ++c)
            if (c < ncls) wacc[c] = fmaf(dl[r * ncls + c], xv, wacc[c]);
        }
      }
      if (threadIdx.x < ncls) {
        float s = 0.f;
        for (int r = 0; r < nrows; ++r) s += dl[r * ncls + threadIdx.x];
        bacc += s;
      }
      __syncthreads();
    }
  }
  if (TRAIN) {
    if (threadIdx.x < nseg * C) {
      const int ci = threadIdx.x % C;
#pragma unroll
      for (int c = 0; c < kMaxCls; ++c)
        if (c < ncls) atomicAdd(dWh + c * C + ci, wacc[c]);
    }
    if (threadIdx.x < ncls) atomicAdd(dbh + threadIdx.x, bacc);
    // block reduce loss
    __shared__ double lred[256];
    lred[threadIdx.x] = lacc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) lred[threadIdx.x] += lred[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(loss_sum, lred[0]);
  }
}

// ---- optimizer / weight layout ---------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            float gscale, __nv_bfloat16* __restrict__ shadow) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pn = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    p[i] = pn;
    if (shadow) shadow[i] = __float2bfloat16_rn(pn);
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

}  // namespace

// ------------------------------------------------------------------------------------------------
int launch_channel_stats(const __nv_bfloat16* y, long long rows, int C, int ld, double* sums,
                         cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(C, &CG, &RL, &threads);
  MPU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  const int grid = grid_for(rows, RL * 8);
  const size_t smem = sizeof(float) * 2 * RL * CG * 8;
  channel_stats_kernel<<<grid, threads, smem, st>>>(y, rows, C, ld, CG, RL, sums);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_colsum(const __nv_bfloat16* x, long long rows, int C, int ld, float* out, cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(C, &CG, &RL, &threads);
  const int grid = grid_for(rows, RL * 8);
  const size_t smem = sizeof(float) * RL * CG * 8;
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

int launch_bn_apply(const __nv_bfloat16* y, const float* scale, const float* shift, __nv_bfloat16* b,
                    __nv_bfloat16* pooled, Geo g, int C, cudaStream_t st) {
  if (pooled) {
    const long long work = (long long)g.B * (g.H / 2) * (g.W / 2) * (C / 8);
    bn_apply_pool_kernel<<<grid_for(work, 256), 256, 0, st>>>(y, scale, shift, b, pooled, g, C);
    count_launch();
  } else {
    const long long work = g.pixels() * (C / 8);
    bn_apply_kernel<<<grid_for(work, 256), 256, 0, st>>>(y, scale, shift, b, g, C);
    count_launch();
  }
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_bn_bwd_reduce(const BnBwdArgs& a, double* sums, cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(a.C, &CG, &RL, &threads);
  MPU_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * a.C, st));
  const bool pool = a.gP != nullptr;
  const long long items = pool ? (long long)a.g.B * (a.g.H / 2) * (a.g.W / 2) : a.g.pixels();
  const int grid = grid_for(items, RL * 4);
  const size_t smem = sizeof(float) * 2 * RL * CG * 8;
  if (pool)
    bn_bwd_kernel<true, false><<<grid, threads, smem, st>>>(a, CG, RL, nullptr, sums, nullptr, 0, nullptr);
  else
    bn_bwd_kernel<false, false><<<grid, threads, smem, st>>>(a, CG, RL, nullptr, sums, nullptr, 0, nullptr);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_bn_bwd_apply(const BnBwdArgs& a, const double* sums, __nv_bfloat16* dz, int phase_major,
                        float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  int CG, RL, threads;
  reduce_layout(a.C, &CG, &RL, &threads);
  const bool pool = a.gP != nullptr;
  const long long items = pool ? (long long)a.g.B * (a.g.H / 2) * (a.g.W / 2) : a.g.pixels();
  const int grid = grid_for(items, RL * 4);
  const size_t smem = sizeof(float) * 2 * RL * CG * 8;
  if (pool)
    bn_bwd_kernel<true, true><<<grid, threads, smem, st>>>(a, CG, RL, sums, nullptr, dz, phase_major, dbias);
  else
    bn_bwd_kernel<false, true><<<grid, threads, smem, st>>>(a, CG, RL, sums, nullptr, dz, phase_major, dbias);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  bn_bwd_params_kernel<<<(a.C + 127) / 128, 128, 0, st>>>(sums, a.mean, a.rstd, a.C, dgamma, dbeta);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

static size_t head_smem(int C, int ncls, bool train) {
  size_t s = sizeof(float) * ((size_t)ncls * C + ((ncls + 3) & ~3));
  if (train) s += (size_t)256 * C * 2 + sizeof(float) * 256 * ncls;
  return s;
}

int launch_conv_first(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out,
                      Geo g, int cin, int co_phys, cudaStream_t st) {
  if (cin < 1 || cin > 8 || co_phys % 8) {
    set_error("conv_first: cin=%d co_phys=%d unsupported", cin, co_phys);
    return MPU_ERR_ARG;
  }
  const size_t smem = sizeof(float) * (9 * cin * co_phys + co_phys);
  if (smem > 48 * 1024) {
    set_error("conv_first: weights do not fit in shared memory");
    return MPU_ERR_ARG;
  }
  const long long work = g.pixels() * (co_phys / 8);
  conv_first_kernel<<<grid_for(work, 256), 256, smem, st>>>(x, w, bias, out, g, cin, co_phys);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_head_infer(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                      float* probs, cudaStream_t st) {
  if (ncls > kMaxCls) {
    set_error("head: n_classes=%d exceeds the supported maximum %d", ncls, kMaxCls);
    return MPU_ERR_ARG;
  }
  const size_t smem = head_smem(C, ncls, false);
  static bool attr = false;
  if (!attr) {
    MPU_CUDA(cudaFuncSetAttribute(head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int grid = grid_for(g.pixels(), 256);
  head_kernel<false><<<grid, 256, smem, st>>>(x, g, C, Wh, bh, ncls, nullptr, nullptr, 0.f, nullptr,
                                             nullptr, nullptr, nullptr, probs);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_head_train(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                      const uint8_t* labels, const float* sample_w, float grad_scale,
                      __nv_bfloat16* dx, float* dWh, float* dbh, double* loss_sum, float* probs_opt,
                      cudaStream_t st) {
  if (ncls > kMaxCls || C > 256) {
    set_error("head(train): n_classes=%d, C=%d not supported (max %d classes, C <= 256)", ncls, C, kMaxCls);
    return MPU_ERR_ARG;
  }
  const size_t smem = head_smem(C, ncls, true);
  static bool attr = false;
  if (!attr) {
    MPU_CUDA(cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  if (smem > 200 * 1024) {
    set_error("head(train): shared memory %zu too large", smem);
    return MPU_ERR_ARG;
  }
  int grid = grid_for(g.pixels(), 256);
  if (grid > 148 * 2) grid = 148 * 2;
  head_kernel<true><<<grid, 256, smem, st>>>(x, g, C, Wh, bh, ncls, labels, sample_w, grad_scale, dx,
                                            dWh, dbh, loss_sum, probs_opt);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1,
                float b2, float eps, float gscale, __nv_bfloat16* shadow, cudaStream_t st) {
  adam_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, gscale, shadow);
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
