// tcgen05 / TMA / TMEM multi-tap GEMM kernels (sm_100a). See mtgemm.cuh for the math.
//
// Warp roles (256 threads, 1 CTA per SM):
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : tcgen05.mma issuer (one elected lane)
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue (TMEM -> registers -> global), warp w owns TMEM lanes 32*(w%4)..+31
#include "mtgemm.cuh"
#include "ptx.cuh"
#include "common.h"

#include <cstdio>
#include <cstring>

namespace mpu {

using namespace ptx;

static constexpr int kThreads = 256;
static constexpr int kSmemBudget = 232448 - 2048;  // 227 KB minus alignment slack + barrier block
static constexpr int kABytes = 128 * 128;          // 128 rows x 64 bf16 (fwd A tile)

// ------------------------------------------------------------------------------------------------
// Forward-type kernel (also used for dgrad): persistent over (m_tile, n_tile), TMEM double-buffered.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) mtgemm_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const uint32_t bar_base = smem_base + (uint32_t)S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * S + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA0);
    if (p.chunks1 > 0) prefetch_tmap(&p.tmA1);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kchunks = p.chunks0 + p.chunks1;
  const int nk = p.ntaps * kchunks;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * 128;
        const int n0 = (tile % p.n_tiles) * p.BN;
        for (int t = 0; t < p.ntaps; ++t) {
          const int arow = m0 + p.tap_a_off[t];
          const int wrow = p.tap_w[t] * p.w_rows_per_tap + n0;
          for (int c = 0; c < kchunks; ++c) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t b_dst = a_dst + kABytes;
            mbar_expect_tx(full_bar(s), stage_bytes);
            int kcol;
            if (c < p.chunks0) {
              tma_load_2d(&p.tmA0, full_bar(s), a_dst, c * 64, arow);
              kcol = c * 64;
            } else {
              tma_load_2d(&p.tmA1, full_bar(s), a_dst, (c - p.chunks0) * 64, arow);
              kcol = p.kofs1 + (c - p.chunks0) * 64;
            }
            tma_load_2d(&p.tmB, full_bar(s), b_dst, kcol, wrow);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        for (int k = 0; k < nk; ++k) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
          const uint64_t a_desc = make_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = make_desc_sw128(a_addr + kABytes, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // advance 16 bf16 (32 B) along K inside the 128 B swizzle row: +2 in 16 B units
            mma_bf16_ss(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc,
                        (k | kk) != 0 ? 1u : 0u);
          }
          mma_commit(empty_bar(s));
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        mma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    const int plane = p.map.Hp * p.map.Wp;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles) * 128;
      const int n0 = (tile % p.n_tiles) * p.BN;
      const int m = m0 + q * 32 + lane;
      bool valid = m < p.M_rows;
      long long orow = 0;
      {
        const int img = m / plane;
        const int rem = m - img * plane;
        const int ya = rem / p.map.Wp;
        const int xa = rem - ya * p.map.Wp;
        valid = valid && ya >= 1 && ya <= p.map.Hp - 2 && xa >= 1 && xa <= p.map.Wp - 2;
        orow = (long long)img * p.map.oHp * p.map.oWp +
               (long long)(p.map.s * (ya - 1) + p.map.py + 1) * p.map.oWp +
               (p.map.s * (xa - 1) + p.map.px + 1);
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + (uint32_t)c0, r);
        tmem_ld_wait();
        const int n = n0 + c0;
        if (valid && n < p.n_valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.n_valid) v[j] += __ldg(p.bias + n + j);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (n + h * 8 < p.n_valid) {
              if (p.mask) {
                const uint4 mk = *reinterpret_cast<const uint4*>(p.mask + orow * p.ldm + n + h * 8);
                const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mk);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (!(__bfloat162float(mb[j]) > 0.f)) v[h * 8 + j] = 0.f;
              }
              uint4 o;
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                ob[j] = __floats2bfloat162_rn(v[h * 8 + 2 * j], v[h * 8 + 2 * j + 1]);
              *reinterpret_cast<uint4*>(p.out + orow * p.ldo + n + h * 8) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Wgrad kernel: one CTA = (ci tile of 128, co tile of BN, tap group, K split). MN-major operands
// straight from the NHWC tensors (pixels are the contraction dim), fp32 atomics into dW.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    mtgemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  int gmax = 1;
  for (int g = 0; g < p.ngroups; ++g) gmax = max(gmax, p.groups[g].count);
  const int nb = (p.BN + 63) / 64;
  const uint32_t a_tap_bytes = 2u * 8192u;  // two 64-channel atoms x 64 K rows
  const uint32_t stage_bytes = (uint32_t)gmax * a_tap_bytes + (uint32_t)nb * 8192u;
  const uint32_t bar_base = smem_base + (uint32_t)S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decomposition
  int w = blockIdx.x;
  const int split = w % p.splits; w /= p.splits;
  const int gi = w % p.ngroups;   w /= p.ngroups;
  const int co_t = w % p.co_tiles; w /= p.co_tiles;
  const int ci_t = w;
  const WgradGroup grp = p.groups[gi];
  const int ci0 = ci_t * 128, co0 = co_t * p.BN;
  const int kb0 = split * p.kblocks_per_split;
  const int kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
  const int nkb = max(0, kb1 - kb0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmX);
    prefetch_tmap(&p.tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)grp.count * a_tap_bytes + (uint32_t)nb * 8192u;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int r0 = kb * 64;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        mbar_expect_tx(full_bar(s), tx);
        for (int g = 0; g < grp.count; ++g) {
          const int xr = r0 + p.tap_x_off[grp.first + g];
          tma_load_2d(&p.tmX, full_bar(s), st + (uint32_t)g * a_tap_bytes, ci0, xr);
          tma_load_2d(&p.tmX, full_bar(s), st + (uint32_t)g * a_tap_bytes + 8192u, ci0 + 64, xr);
        }
        const uint32_t bst = st + (uint32_t)gmax * a_tap_bytes;
        for (int j = 0; j < nb; ++j)
          tma_load_2d(&p.tmDY, full_bar(s), bst + (uint32_t)j * 8192u, co0 + j * 64, r0 + grp.dy_off);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.BN, 1, 1);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t bst = st + (uint32_t)gmax * a_tap_bytes;
        for (int g = 0; g < grp.count; ++g) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(g * p.BN);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t a_desc = make_desc_sw128(
                st + (uint32_t)g * a_tap_bytes + (uint32_t)(kk * p.kstep_bytes), p.a_lbo, p.a_sbo);
            const uint64_t b_desc =
                make_desc_sw128(bst + (uint32_t)(kk * p.kstep_bytes), p.b_lbo, p.b_sbo);
            mma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
        }
        mma_commit(empty_bar(s));
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      mma_commit(done_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (nkb > 0) {
      const int ci = ci0 + q * 32 + lane;
      for (int g = 0; g < grp.count; ++g) {
        const int tw = p.tap_w[grp.first + g];
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.BN);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(t_row + (uint32_t)c0, r);
          tmem_ld_wait();
          if (ci < p.ci_valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = co0 + c0 + j;
              if (co < p.co_valid) {
                float* dst = p.dW + ((size_t)tw * p.w_rows_per_tap + co) * p.ldw + p.dw_col0 + ci;
                atomicAdd(dst, __uint_as_float(r[j]));
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return MPU_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u",
              (int)r, base, (unsigned long long)rows, (unsigned long long)cols,
              (unsigned long long)ld_elems, box_cols, box_rows);
    return MPU_ERR_CUDA;
  }
  return MPU_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static int g_fwd_attr_set = 0, g_wgrad_attr_set = 0;
static constexpr int kDynSmem = 232448 - 1024;  // leave room for static smem (none) and the driver

int launch_fwd(FwdParams& p, cudaStream_t stream) {
  if (p.BN % 16 != 0 || p.BN < 16 || p.BN > 256) {
    set_error("launch_fwd: BN=%d must be a multiple of 16 in [16,256]", p.BN);
    return MPU_ERR_ARG;
  }
  if (p.ntaps < 1 || p.ntaps > kMaxTaps) {
    set_error("launch_fwd: ntaps=%d out of range", p.ntaps);
    return MPU_ERR_ARG;
  }
  const int stage_bytes = kABytes + p.BN * 128;
  int S = kSmemBudget / stage_bytes;
  if (S > 8) S = 8;
  if (S < 2) {
    set_error("launch_fwd: not enough shared memory for 2 stages");
    return MPU_ERR_ARG;
  }
  p.stages = S;
  p.m_tiles = (p.M_rows + 127) / 128;
  p.n_tiles = (p.n_valid + p.BN - 1) / p.BN;
  if (!g_fwd_attr_set) {
    MPU_CUDA(cudaFuncSetAttribute(mtgemm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kDynSmem));
    g_fwd_attr_set = 1;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  mtgemm_fwd_kernel<<<grid, kThreads, kDynSmem, stream>>>(p);
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int launch_wgrad(WgradParams& p, cudaStream_t stream) {
  if (p.BN % 16 != 0 || p.BN < 16 || p.BN > 256) {
    set_error("launch_wgrad: BN=%d must be a multiple of 16 in [16,256]", p.BN);
    return MPU_ERR_ARG;
  }
  int gmax = 1;
  for (int g = 0; g < p.ngroups; ++g) gmax = p.groups[g].count > gmax ? p.groups[g].count : gmax;
  if (gmax * p.BN > 512) {
    set_error("launch_wgrad: group %d x BN %d exceeds 512 TMEM columns", gmax, p.BN);
    return MPU_ERR_ARG;
  }
  const int nb = (p.BN + 63) / 64;
  const int stage_bytes = gmax * 16384 + nb * 8192;
  int S = kSmemBudget / stage_bytes;
  if (S > 8) S = 8;
  if (S < 2) {
    set_error("launch_wgrad: not enough shared memory for 2 stages");
    return MPU_ERR_ARG;
  }
  p.stages = S;
  if (!g_wgrad_attr_set) {
    MPU_CUDA(cudaFuncSetAttribute(mtgemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kDynSmem));
    g_wgrad_attr_set = 1;
  }
  const int grid = p.ci_tiles * p.co_tiles * p.ngroups * p.splits;
  mtgemm_wgrad_kernel<<<grid, kThreads, kDynSmem, stream>>>(p);
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

}  // namespace mpu
