// tcgen05 / TMA / TMEM multi-tap GEMM kernels (sm_100a). See mtgemm.cuh for the math.
//
// Measured facts these kernels are built around (B200; ncu, in-kernel cycle counters and tests/mma_issue_bench.cu,
// evidence under profiles/, discussion in DESIGN.md section 3):
//   * the tensor pipe costs N/2 cycles per M=128, K=16 MMA at full rate down to N=128 (4095 MAC/cycle/SM), whatever
//     else the SM does (TMA refill, tcgen05.ld, row-shifted or MN-major descriptors); N=96 costs 64 cycles.  The
//     forward/dgrad kernel therefore puts the OUTPUT CHANNELS on the 128 MMA rows (weights = A operand) and 256 PIXELS
//     on the columns; wgrad stacks the three kx taps of a kernel row along N (N = 192);
//   * what limits a main loop is the ISSUE path of the single issuing thread (~15 uniform-datapath instructions per
//     MMA): one elect.sync at role entry and a plain single-thread loop inside, descriptor high words hoisted, K tails
//     issued through unrolled immediates (a rolled loop cost ~215 cycles per MMA);
//   * the 128B swizzle is applied on absolute shared-memory address bits: a descriptor whose start address is shifted
//     by whole 128 B rows inside a TMA-written tile reads the shifted rows correctly with the base-offset field left
//     at 0.  One activation "slab" load therefore serves the kx-1, kx, kx+1 taps of a 3x3 kernel row - or all nine taps
//     where the padded image rows are short;
//   * operand delivery L2 -> SM (~45-57 B/clk/SM through TMA) is the next bound on the fine levels: slabs are shared
//     across taps (and across the two CTAs of a cluster by multicast), rings of three slabs / three or four weight
//     stages cover the refill round trip;
//   * the SM clock of sustained operation is set by the ~1 kW power cap (1.62-1.80 GHz of 1.965): wasted MMA work
//     (padded rows) costs twice.
#include "mtgemm.cuh"
#include "ptx.cuh"
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace mpu {

using namespace ptx;

// mbarrier wait that optionally accumulates the stall time (bring-up profiling of the role pipeline)
__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, long long* acc) {
  if (acc) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

static constexpr int kThreads = 256;       // wgrad kernel
static constexpr int kFwdThreads = 384;    // fwd kernel: 4 control warps + 8 epilogue warps
static constexpr int kSmemBudget = 232448 - 1024 - 1024 - 512;  // 227 KB minus launch slack, alignment, barriers

static constexpr int kPT = 256;                              // pixels per CTA tile (MMA N)
static constexpr uint32_t kSlabRows = kPT + 8;               // 264: room for tap shifts of up to 7 rows
static constexpr uint32_t kSlabBytes = kSlabRows * 128u;     // 33792 (multiple of 1024)
static constexpr uint32_t kWTileBytes = 128u * 128u;         // 128 output channels x 64 K
static constexpr uint32_t kWStageBytes = 2u * kWTileBytes;    // one weight-ring stage = the tiles of TWO consecutive
                                                             // (chunk, tap) items behind one mbarrier: 8 MMAs per wait
static constexpr uint32_t kStageHalfBytes = 64u * 256u;      // epilogue staging per pixel half: up to 64 pixels x 128 ch bf16

// ------------------------------------------------------------------------------------------------
// Forward-type kernel (forward conv, dgrad, upsample-conv phases).  Persistent over CTA tiles of
// 256 pixels x 128 output channels:  D[co, px] = sum_{tap, k-chunk} W_tap[co, k] * Slab[px + shift_tap, k]^T
//   warp 0   : TMA producer   (activation slabs -> ring A, weight tiles -> ring W)
//   warp 1   : tcgen05.mma issuer, fp32 accumulators in TMEM (2 stages x 256 columns)
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue.  Thread = output channel (TMEM lane), 16 pixels per tcgen05.ld: bias + ReLU,
//              bf16 rounding, optional per-channel sum / sum-of-squares (BatchNorm statistics or bias
//              gradients, accumulated in registers across tiles), transpose through shared memory, then
//              row-wise coalesced 16-byte stores (optional ReLU-backward mask applied there).
// ------------------------------------------------------------------------------------------------
// W_MN: weights read MN-major from the forward layout (dgrad).  PROF: CTA 0 accumulates role stall cycles in p.dbg
// (bring-up only; the production instantiations carry no timing code).
// CL2: launched as clusters of two CTAs that work on the two 128-channel tiles of the SAME 256-pixel tile in lockstep:
// every activation slab is fetched from L2 once per cluster - each CTA issues half of the slab's TMA boxes with
// .multicast::cluster to both CTAs - which halves the slab traffic that starves level 1 (its 130-pixel padded rows are
// too long for the 9-tap slab; profiles/r02_fwd_ablation.txt).  A slab slot is released by the tcgen05.commit of BOTH
// CTAs' MMA threads (multicast arrive), so each producer knows both copies are free before it overwrites them.
// RED: the epilogue additionally takes column reductions of the stored tile (FwdParams::csum_f / red_d); a separate
// instantiation so that the plain kernels carry neither its registers nor its branches.
// CLM (cluster mode): 0 = independent CTAs; 1 = "CL2" above (the two channel tiles of one pixel tile share every slab);
// 2 = the two CTAs of a cluster work on two ADJACENT PIXEL TILES of the SAME channel tile and share the WEIGHT stream:
// each stage of the weight ring carries two (chunk, tap) tiles, CTA rank r fetches tile r and multicasts it to both
// (a conv with a single channel tile - level 0 - cannot share slabs, and its weights are re-streamed for every one of
// its 8256 pixel tiles); a weight stage is released by the commits of both MMA threads.
template <bool W_MN, bool PROF, int CLM, bool RED>
__global__ void __launch_bounds__(kFwdThreads, 1) mtgemm_fwd_kernel(const __grid_constant__ FwdParams p) {
  constexpr bool CL2 = CLM == 1;   // slab sharing
  constexpr bool CLW = CLM == 2;   // weight sharing
  constexpr bool CLUSTER = CLM != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int NA = p.a_slots, NW = p.w_slots;
  const uint32_t slab_bytes = (uint32_t)p.slab_rows * 128u;  // multiple of 1024
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + (uint32_t)NA * slab_bytes;
  const uint32_t w_tile_bytes = (uint32_t)p.w_tile_bytes;       // one (chunk, tap) weight tile: box rows x 128 B
  const uint32_t w_stage_bytes = 2u * w_tile_bytes;             // a ring stage carries two consecutive items
  const uint32_t stg_off = (uint32_t)NA * slab_bytes + (uint32_t)NW * w_stage_bytes;
  const uint32_t stg_half = (uint32_t)p.stg_px * (uint32_t)p.stg_ch * 2u;  // one half's staging: stg_px/2 pixel pairs x stg_ch x 4 B
  const uint32_t orow_off = stg_off + 2u * stg_half;                                 // int[256]
  const uint32_t bar_base = smem_base + orow_off + 1024u;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (NA + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * NA + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * NA + NW + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * NA + 2 * NW + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * NA + 2 * NW + 2 + s); };
  const uint32_t tmem_slot_off = orow_off + 1024u + 8u * (2 * NA + 2 * NW + 4);
  const uint32_t vmask_off = tmem_slot_off + 16u;  // uint32[2][4]: valid-pixel bits of the tile, 32 pixels per word
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + tmem_slot_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA0_hi);
    prefetch_tmap(&p.tmA0_lo);
    if (p.chunks1 > 0) {
      prefetch_tmap(&p.tmA1_hi);
      prefetch_tmap(&p.tmA1_lo);
    }
    prefetch_tmap(&p.tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), CL2 ? 2 : 1);
    }
    for (int s = 0; s < NW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), CLW ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_base + tmem_slot_off, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();  // the partner's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Work units: plain mode - one (pixel tile, channel tile) per CTA and step, channel tile fastest; cluster mode - one
  // (pixel tile, channel-tile PAIR) per cluster and step, CTA rank r takes channel tile 2 * pair + r.
  // weight-sharing mode - one (pixel-tile PAIR, channel tile) per cluster and step, CTA rank r takes pixel tile
  // 2 * pair + r (with an odd number of pixel tiles the last pair's second tile lies beyond the tensor: its CTA runs the
  // whole protocol on zero-filled slabs and stores nothing).
  const int crank = CLUSTER ? (int)cluster_ctarank() : 0;
  const int n_div = CL2 ? (p.n_tiles >> 1) : p.n_tiles;
  const int m_div = CLW ? ((p.m_tiles + 1) >> 1) : p.m_tiles;
  const int total_tiles = m_div * n_div;
  const int unit0 = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = CLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto pix_tile = [&](int unit) { return CLW ? (unit / n_div) * 2 + crank : unit / n_div; };
  auto ch_tile = [&](int unit) { return CL2 ? (unit % n_div) * 2 + crank : unit % n_div; };
  const int kchunks = p.chunks0 + p.chunks1;
  const bool prof = PROF && p.dbg != nullptr && blockIdx.x == 0;
  long long w0 = 0, w1 = 0, w2 = 0;
  const long long t_start = prof ? clock64() : 0;

  if (warp == 0) {
    // ===== TMA producer 1: activation slabs -> ring A =====
    // The two operand streams have their own producer threads (warp 0: slabs, warp 3: weight tiles): a single
    // sequential producer that blocks on a full slab ring cannot refill freed weight slots (and vice versa), which
    // showed up as the MMA thread waiting on w_full while the producer sat in a_empty.
    if (elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int m0 = pix_tile(tile) * kPT;
        for (int g = 0; g < p.ngroups; ++g) {
          const int arow = m0 + p.groups[g].a_off;
          for (int c = 0; c < kchunks; ++c) {
            timed_wait(a_empty(as), aph ^ 1u, prof ? &w0 : nullptr);
            const uint32_t a_dst = a_base + (uint32_t)as * slab_bytes;
            const bool src0 = c < p.chunks0;
            const int ccol = (src0 ? c : c - p.chunks0) * 64;
            const CUtensorMap* mhi = src0 ? &p.tmA0_hi : &p.tmA1_hi;
            const CUtensorMap* mlo = src0 ? &p.tmA0_lo : &p.tmA1_lo;
            if (PROF && !CL2 && (p.dbg_flags & 2)) {
              mbar_arrive(a_full(as));
            } else {
              mbar_expect_tx(a_full(as), slab_bytes);  // all boxes land in this CTA, whoever issues them
              if (!CL2) {
                tma_load_2d(mhi, a_full(as), a_dst, ccol, arow);
                tma_load_2d(mlo, a_full(as), a_dst + 136u * 128u, ccol, arow + 136);
                if (p.ext_rows > 0)
                  tma_load_2d(src0 ? &p.tmA0_ext : &p.tmA1_ext, a_full(as), a_dst + 264u * 128u, ccol, arow + 264);
              } else if (crank == 0) {
                tma_load_2d_multicast(mhi, a_full(as), a_dst, ccol, arow, (uint16_t)3);
              } else {
                tma_load_2d_multicast(mlo, a_full(as), a_dst + 136u * 128u, ccol, arow + 136, (uint16_t)3);
                if (p.ext_rows > 0)
                  tma_load_2d_multicast(src0 ? &p.tmA0_ext : &p.tmA1_ext, a_full(as), a_dst + 264u * 128u, ccol,
                                        arow + 264, (uint16_t)3);
              }
            }
            if (++as == NA) { as = 0; aph ^= 1u; }
          }
        }
      }
      if (prof) p.dbg[0] = w0;
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===== TMA producer 2: weight tiles -> ring W (two consecutive (chunk, tap) items per stage) =====
    if (elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int n0 = ch_tile(tile) * 128;
        int g = 0, c = 0, t = 0;
        int ntaps = p.groups[0].ntaps;
        const int rbase = W_MN ? 0 : n0;
        int left = p.items_per_tile;
        while (left > 0) {
          const int nit = left >= 2 ? 2 : 1;
          timed_wait(w_empty(ws), wph ^ 1u, prof ? &w1 : nullptr);
          const uint32_t w_dst = w_base + (uint32_t)ws * w_stage_bytes;
          if (!(PROF && (p.dbg_flags & 4))) mbar_expect_tx(w_full(ws), (uint32_t)nit * w_tile_bytes);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i < nit) {
              const int kcol = c < p.chunks0 ? c * 64 : p.kofs1 + (c - p.chunks0) * 64;
              // row (K-major: output channel row; MN-major: K row) of the tap's block in the weight matrix
              const int wr = p.groups[g].w_idx[t] * p.w_rows_per_tap + rbase;
              const uint32_t dst = w_dst + (uint32_t)i * w_tile_bytes;
              if (PROF && (p.dbg_flags & 4)) {
                // bring-up: no weight traffic
              } else if (CLW) {  // tile i of the stage is fetched by CTA rank i and delivered to both CTAs
                if (i == crank) {
                  if (!W_MN) {
                    tma_load_2d_multicast(&p.tmW, w_full(ws), dst, kcol, wr, (uint16_t)3);
                  } else {
                    tma_load_2d_multicast(&p.tmW, w_full(ws), dst, n0, wr + kcol, (uint16_t)3);
                    tma_load_2d_multicast(&p.tmW, w_full(ws), dst + 8192u, n0 + 64, wr + kcol, (uint16_t)3);
                  }
                }
              } else if (!W_MN) {
                tma_load_2d(&p.tmW, w_full(ws), dst, kcol, wr);
              } else {  // two 64-channel atoms of [64 K rows][64 channels]
                tma_load_2d(&p.tmW, w_full(ws), dst, n0, wr + kcol);
                tma_load_2d(&p.tmW, w_full(ws), dst + 8192u, n0 + 64, wr + kcol);
              }
              if (++t == ntaps) {
                t = 0;
                if (++c == kchunks) {
                  c = 0;
                  if (++g < p.ngroups) ntaps = p.groups[g].ntaps;
                  else g = 0;  // (tile finished; keeps the index of the look-ups above in range)
                }
              }
            }
          }
          if (PROF && (p.dbg_flags & 4)) mbar_arrive(w_full(ws));
          left -= nit;
          if (++ws == NW) { ws = 0; wph ^= 1u; }
        }
      }
      if (prof) p.dbg[1] = w1;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // ONE elected thread runs the whole role (waits, descriptor arithmetic, issue, commits): measured on B200
    // (tests/mma_issue_bench.cu, profiles/r02_mma_issue_cost.txt) the tensor pipe sustains its floor of N/2 cycles
    // per M=128 MMA with both operands in shared memory, concurrent TMA refill, tcgen05.ld traffic and row-shifted
    // descriptors - what held the round-1 loop at ~190 cycles per MMA was the ~100-instruction, latency-chained
    // issue path per tap (constant-bank lookups with register indices, R2UR chains, per-tap elect/reconverge).
    // Everything that does not change inside a tile is hoisted; a tap costs one mbarrier wait, two 64-bit adds
    // per MMA and one commit.
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, kPT, W_MN ? 1 : 0, 0);
      const uint64_t w_hi = (W_MN ? make_desc_sw128(0, 8192, 1024) : make_desc_sw128(0, 16, 1024));
      const uint64_t x_hi = make_desc_sw128(0, 16, 1024);
      constexpr uint64_t w_step = W_MN ? 128u : 2u;  // K = 16 step in 16-byte units (MN-major: 16 rows = 2048 B)
      const uint32_t a_lo0 = (a_base & 0x3FFFFu) >> 4, w_lo0 = (w_base & 0x3FFFFu) >> 4;
      // 64-channel chunks whose tail holds fewer than 4 K=16 steps of real channels (the rest is TMA zero fill)
      const int tail_c0 = (p.c0_valid & 63) ? p.chunks0 - 1 : -1;
      const int tail_k0 = ((p.c0_valid & 63) + 15) >> 4;
      const int tail_c1 = (p.chunks1 > 0 && (p.c1_valid & 63)) ? kchunks - 1 : -1;
      const int tail_k1 = ((p.c1_valid & 63) + 15) >> 4;
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      int acs = 0;
      uint32_t acph = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        timed_wait(tempty_bar(acs), acph ^ 1u, prof ? &w2 : nullptr);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acs * kPT);
        uint32_t acc = 0;
        int g = 0, c = 0, t = 0;
        int ntaps = p.groups[0].ntaps;
        uint32_t sh = (uint32_t)p.groups[0].shift[0] * 8u;  // row shift of the CURRENT tap in 16-byte units, looked up
                                                            // one tap ahead (the constant-bank load is off the issue path)
        uint32_t a_lo = 0;
        int ks = 4;
        int left = p.items_per_tile;
        while (left > 0) {
          const int nit = left >= 2 ? 2 : 1;
          timed_wait(w_full(ws), wph, prof ? &w1 : nullptr);
          tc_fence_after();
          const uint32_t w_lo = w_lo0 + (uint32_t)ws * (w_stage_bytes >> 4);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i < nit) {
              if (t == 0) {  // first tap of a (group, chunk): its slab must have landed
                timed_wait(a_full(as), aph, prof ? &w0 : nullptr);
                tc_fence_after();
                a_lo = a_lo0 + (uint32_t)as * (slab_bytes >> 4);
                ks = c == tail_c0 ? tail_k0 : (c == tail_c1 ? tail_k1 : 4);
              }
              const uint64_t wd = w_hi | (uint64_t)(w_lo + (uint32_t)i * (w_tile_bytes >> 4));
              const uint64_t xd = x_hi | (uint64_t)(a_lo + sh);
              // 1..4 K=16 steps, descriptor offsets as immediates.  (A rolled loop for the short tail chunks cost
              // ~215 cycles per MMA - R2UR + 64-bit multiply chains per iteration - which made level 0, where half of
              // the items are 2-step tails of the 96-channel K, issue-bound: 161 cycles per MMA with nothing else
              // running, profiles/r02_fwd_ablation.txt.)
              mma_bf16_ss(d_tmem, wd, xd, idesc, acc);
              if (ks >= 2) mma_bf16_ss(d_tmem, wd + w_step, xd + 2, idesc, 1u);
              if (ks >= 3) mma_bf16_ss(d_tmem, wd + 2 * w_step, xd + 4, idesc, 1u);
              if (ks >= 4) mma_bf16_ss(d_tmem, wd + 3 * w_step, xd + 6, idesc, 1u);
              acc = 1u;
              if (++t == ntaps) {  // last tap of the chunk: the slab is free once these MMAs retire
                if (CL2) mma_commit_multicast(a_empty(as), (uint16_t)3);
                else mma_commit(a_empty(as));
                if (++as == NA) { as = 0; aph ^= 1u; }
                t = 0;
                if (++c == kchunks) {
                  c = 0;
                  if (++g < p.ngroups) ntaps = p.groups[g].ntaps;
                  else g = 0;
                }
              }
              sh = (uint32_t)p.groups[g].shift[t] * 8u;
            }
          }
          if (CLW) mma_commit_multicast(w_empty(ws), (uint16_t)3);
          else mma_commit(w_empty(ws));
          left -= nit;
          if (++ws == NW) { ws = 0; wph ^= 1u; }
        }
        mma_commit(tfull_bar(acs));
        if (++acs == 2) { acs = 0; acph ^= 1u; }
      }
      if (prof) {
        p.dbg[2] = w0;
        p.dbg[3] = w1;
        p.dbg[4] = w2;
        p.dbg[7] = clock64() - t_start;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue =====
    // Thread = output channel (TMEM lane).  The 128 pixels of this half are read in four blocks of 32 columns, the
    // tcgen05.ld of block k+1 in flight while block k is converted (the accumulator is released right after the last
    // load has landed, not after the last conversion).  Conversion packs TWO pixels of the thread's channel per
    // cvt.rn.bf16x2 (F2FP; the scalar form compiles to the quarter-rate F2F) and stores them as one 32-bit word:
    // the staging buffer is laid out [pixel pair][channel] (4 B elements), so a warp's store is one conflict-free
    // 128 B wavefront instead of two 64 B ones.  The row-wise copy-out un-interleaves the pairs with byte permutes.
    const int q = warp & 3;              // TMEM lane quarter: channels q*32 .. q*32+31 of the tile
    const int half = (warp - 4) >> 2;    // pixel half: columns half*128 .. +127
    const int et = threadIdx.x - 128 - half * 128;  // 0..127 inside the half
    const int ch_local = q * 32 + lane;
    const int CS = p.stg_ch;             // channels per staging row (128, or the tensor's width for a single narrow tile)
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem_al + stg_off + (uint32_t)half * stg_half);
    int* orow_s = reinterpret_cast<int*>(smem_al + orow_off) + half * 128;
    volatile uint32_t* vm_s = reinterpret_cast<volatile uint32_t*>(smem_al + vmask_off) + half * 4;
    const bool st_ok = ch_local < CS;
    const int plane = p.map.Hp * p.map.Wp;
    const bool prof_e = prof && warp == 4 && lane == 0;
    const bool do_stats = p.stats != nullptr;
    const bool do_relu = p.relu != 0;
    int acs = 0;
    uint32_t acph = 0;
    float s_sum = 0.f, s_sq = 0.f;
    int s_ch = -1;
    // column reductions of the copy-out phase: this thread always handles the 8 channels n0 + (lane & 15) * 8 ..
    const bool do_red = RED && (p.csum_f != nullptr || p.red_d != nullptr);
    float cs[8], cy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = cy[j] = 0.f;
    int cs_ch = -1;
    auto flush_red = [&]() {
      if (cs_ch < 0) return;
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // the two row slots of a warp (lanes l, l + 16) hold the same channels
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
        cy[j] += __shfl_xor_sync(0xffffffffu, cy[j], 16);
      }
      if (lane < 16 && cs_ch < p.n_valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cs_ch + j;
          if (p.csum_f) atomicAdd(p.csum_f + c, cs[j]);
          const int rc = c - p.red_col0;
          if (p.red_d && rc >= 0 && rc < p.red_C) {
            atomicAdd(p.red_d + rc, (double)cs[j]);
            atomicAdd(p.red_d + p.red_C + rc, (double)cy[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] = cy[j] = 0.f;
    };
    const int SP = p.stg_px;
    for (int tile = unit0; tile < total_tiles; tile += unit_step) {
      const int m0 = pix_tile(tile) * kPT;
      const int n0 = ch_tile(tile) * 128;
      const int ch = n0 + ch_local;
      if (do_red && n0 + (lane & 15) * 8 != cs_ch) {  // (warp-uniform: n0 changes for the whole warp at once)
        flush_red();
        cs_ch = n0 + (lane & 15) * 8;
      }
      if (do_stats && ch != s_ch) {  // channel tile changed: flush the register accumulators
        if (s_ch >= 0 && s_ch < p.n_valid) {
          atomicAdd(p.stats + s_ch, (double)s_sum);
          atomicAdd(p.stats + p.n_valid + s_ch, (double)s_sq);
        }
        s_sum = s_sq = 0.f;
        s_ch = ch;
      }
      // output row (or -1) of pixel `et` of this half, and the half's valid-pixel bit masks
      {
        const int m = m0 + half * 128 + et;
        int orow = -1;
        if (m < p.M_rows) {
          const int img = m / plane;
          const int rem = m - img * plane;
          const int ya = rem / p.map.Wp;
          const int xa = rem - ya * p.map.Wp;
          if (ya >= 1 && ya <= p.map.Hp - 2 && xa >= 1 && xa <= p.map.Wp - 2)
            orow = img * p.map.oHp * p.map.oWp + (p.map.s * (ya - 1) + p.map.py + 1) * p.map.oWp +
                   (p.map.s * (xa - 1) + p.map.px + 1);
        }
        orow_s[et] = orow;
        const uint32_t bal = __ballot_sync(0xffffffffu, orow >= 0);
        if (lane == 0) vm_s[q] = bal;
      }
      named_bar_sync(1 + half, 128);
      timed_wait(tfull_bar(acs), acph, prof_e ? &w0 : nullptr);
      tc_fence_after();
      const long long te0 = prof_e ? clock64() : 0;
      const float bias = (p.bias && ch < p.n_valid) ? __ldg(p.bias + ch) : 0.f;
      const bool warp_live = n0 + q * 32 < p.n_valid;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acs * kPT + half * 128);
      if (PROF && (p.dbg_flags & 1)) {  // bring-up: release the accumulator untouched
        tc_fence_before();
        mbar_arrive(tempty_bar(acs));
        named_bar_sync(1 + half, 128);
        if (++acs == 2) { acs = 0; acph ^= 1u; }
        continue;
      }
      const int nchunk = min(16, (p.n_valid - n0 + 7) >> 3);
      const int sub = lane >> 4, chunk = lane & 15;
      const int pairs_w = SP >> 3;  // pixel pairs copied out by one warp per pass (SP / 2 pairs over 4 warps)

      // one 32-pixel block of this thread's channel: bias + ReLU, bf16x2 packing, staging store, BatchNorm statistics
      auto convert = [&](const uint32_t (&r)[2][16], int blk) {
        const int pp0 = ((blk * 32) & (SP - 1)) >> 1;  // first pixel pair of the block inside the pass
        const uint32_t vm = do_stats ? vm_s[blk] : 0u;
        const bool all_valid = vm == 0xffffffffu;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float v0 = __uint_as_float(r[h][2 * jj]) + bias;
            float v1 = __uint_as_float(r[h][2 * jj + 1]) + bias;
            if (do_relu) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);  // .x (low half) = even pixel
            const uint32_t u = *reinterpret_cast<const uint32_t*>(&pk);
            if (st_ok) stg[(pp0 + h * 8 + jj) * CS + ch_local] = u;
            if (do_stats) {
              const float f0 = __uint_as_float(u << 16), f1 = __uint_as_float(u & 0xffff0000u);
              if (all_valid) {
                s_sum += f0;
                s_sq = fmaf(f0, f0, s_sq);
                s_sum += f1;
                s_sq = fmaf(f1, f1, s_sq);
              } else {
                if ((vm >> (h * 16 + 2 * jj)) & 1u) {
                  s_sum += f0;
                  s_sq = fmaf(f0, f0, s_sq);
                }
                if ((vm >> (h * 16 + 2 * jj + 1)) & 1u) {
                  s_sum += f1;
                  s_sq = fmaf(f1, f1, s_sq);
                }
              }
            }
          }
        }
      };
      // one stored row chunk: ReLU-backward mask, 16-byte store, optional column reductions
      auto store_chunk = [&](int orow, uint4 val, uint4 mk) {
        const long long o = (long long)orow;
        if (p.mask) {
          const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mk);
          __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(&val);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!(__bfloat162float(mb[j]) > 0.f)) vb[j] = __float2bfloat16_rn(0.f);
        }
        *reinterpret_cast<uint4*>(p.out + o * p.ldo + n0 + chunk * 8) = val;
        if (do_red) {
          const __nv_bfloat16* vb = reinterpret_cast<const __nv_bfloat16*>(&val);
          float vf[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            vf[j] = __bfloat162float(vb[j]);
            cs[j] += vf[j];
          }
          const int rc = n0 + chunk * 8 - p.red_col0;
          if (p.red_y && rc >= 0 && rc < p.red_C) {
            const uint4 yk = *reinterpret_cast<const uint4*>(p.red_y + o * p.red_ldy + rc);
            const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(&yk);
#pragma unroll
            for (int j = 0; j < 8; ++j) cy[j] = fmaf(vf[j], __bfloat162float(yb[j]), cy[j]);
          }
        }
      };
      // row-wise copy-out of one pass: a warp instruction covers 2 pixel pairs x 16 chunks of 8 channels; every thread
      // reads its 8 channels x 2 pixels (32 B), splits them into the even and the odd pixel's 16 bytes and stores both
      auto copy_out = [&](int px0) {
#pragma unroll 2
        for (int pp = q * pairs_w + sub; pp < (q + 1) * pairs_w; pp += 2) {
          const int o0 = orow_s[px0 + 2 * pp], o1 = orow_s[px0 + 2 * pp + 1];
          if (chunk < nchunk && (o0 >= 0 || o1 >= 0)) {
            const uint4 lo = *reinterpret_cast<const uint4*>(stg + pp * CS + chunk * 8);
            const uint4 hi = *reinterpret_cast<const uint4*>(stg + pp * CS + chunk * 8 + 4);
            uint4 mk0 = make_uint4(0, 0, 0, 0), mk1 = mk0;
            if (p.mask) {  // both mask loads are issued before either is used
              if (o0 >= 0) mk0 = *reinterpret_cast<const uint4*>(p.mask + (long long)o0 * p.ldm + n0 + chunk * 8);
              if (o1 >= 0) mk1 = *reinterpret_cast<const uint4*>(p.mask + (long long)o1 * p.ldm + n0 + chunk * 8);
            }
            uint4 ev, od;
            ev.x = __byte_perm(lo.x, lo.y, 0x5410); od.x = __byte_perm(lo.x, lo.y, 0x7632);
            ev.y = __byte_perm(lo.z, lo.w, 0x5410); od.y = __byte_perm(lo.z, lo.w, 0x7632);
            ev.z = __byte_perm(hi.x, hi.y, 0x5410); od.z = __byte_perm(hi.x, hi.y, 0x7632);
            ev.w = __byte_perm(hi.z, hi.w, 0x5410); od.w = __byte_perm(hi.z, hi.w, 0x7632);
            if (o0 >= 0) store_chunk(o0, ev, mk0);
            if (o1 >= 0) store_chunk(o1, od, mk1);
          }
        }
      };
      // a warp whose 32 channels all lie beyond the tensor (90 channels in a 128-row tile: warp q = 3) has nothing to
      // load, convert or stage; it still takes part in the barriers and in the row-wise copy-out
      auto step = [&](const uint32_t (&cur)[2][16], uint32_t (&nxt)[2][16], int blk) {
        const int pxb = blk * 32;
        if (blk && (pxb & (SP - 1)) == 0) named_bar_sync(1 + half, 128);  // previous pass's copy-out is done with the staging buffer
        if (warp_live) {
          tmem_ld_wait();
          if (blk < 3) {
            tmem_ld16(t_row + (uint32_t)(pxb + 32), nxt[0]);
            tmem_ld16(t_row + (uint32_t)(pxb + 48), nxt[1]);
          }
        }
        if (blk == 3) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acs));  // accumulator drained: the MMA warp may start the tile after next
        }
        if (warp_live) convert(cur, blk);
        if (((pxb + 32) & (SP - 1)) == 0) {
          named_bar_sync(1 + half, 128);
          copy_out(pxb + 32 - SP);
        }
      };
      uint32_t ra[2][16], rb[2][16];
      if (warp_live) {
        tmem_ld16(t_row, ra[0]);
        tmem_ld16(t_row + 16u, ra[1]);
      }
      step(ra, rb, 0);
      step(rb, ra, 1);
      step(ra, rb, 2);
      step(rb, ra, 3);
      named_bar_sync(1 + half, 128);  // staging + row table are reused by the next tile
      if (prof_e) w1 += clock64() - te0;
      if (++acs == 2) { acs = 0; acph ^= 1u; }
    }
    if (p.stats && s_ch >= 0 && s_ch < p.n_valid) {
      atomicAdd(p.stats + s_ch, (double)s_sum);
      atomicAdd(p.stats + p.n_valid + s_ch, (double)s_sq);
    }
    if (do_red) flush_red();
    if (prof_e) {
      p.dbg[5] = w0;
      p.dbg[6] = w1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();  // no CTA leaves while its partner may still multicast data or arrivals into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Wgrad kernel.  dW_t[co, ci] = sum_m dY[m, co] * X[m + x_off + j_t, ci]  with pixels m as the
// contraction dim; both operands are read MN-major straight from the NHWC tensors.
//   * A = dY^T tile (M = 128 output channels, two 64-channel atoms);
//   * B = X slab of 72 rows x 64 input channels.  The kx taps of one kernel row are STACKED ALONG N of a
//     single MMA: the B descriptor's atom stride (LBO) is one 128-byte row, so N-atom j is the slab
//     shifted by j rows - one MMA of N = 64*ntaps computes all taps of the row at once;
//   * one CTA = (co tile of 128) x (CA ci atoms) x (tap row) x (K split); fp32 accumulators in TMEM
//     (CA x 64*ntaps columns); `red.global.add.v4.f32` into the flat gradient at the end.
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kWgSlabRows = 72;
static constexpr uint32_t kWgAtomBytes = kWgSlabRows * 128u;  // 9216: one 64-channel X slab
static constexpr uint32_t kWgABytes = 2u * 8192u;             // dY tile: 2 atoms x 64 rows x 128 B

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// PROF: CTA 0 writes role timings to p.dbg (bring-up only): [0] producer waits on empty slots, [1] MMA thread waits on
// full slots, [2] start -> all MMAs retired, [3] accumulator drain (TMEM -> red.global), [4] whole CTA, [5] prologue,
// [6] K blocks of the CTA.
template <bool PROF>
__global__ void __launch_bounds__(kThreads, 1)
    mtgemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const bool prof = PROF && p.dbg != nullptr && blockIdx.x == 0;
  const long long t_start = prof ? clock64() : 0;
  long long pw = 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int CA = p.CA;
  const uint32_t stage_bytes = kWgABytes + (uint32_t)CA * kWgAtomBytes;
  const uint32_t bar_base = smem_base + (uint32_t)S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decomposition: the CTAs that read the same pixel range (all kernel rows, output-channel tiles and
  // input-channel tiles of one K split) are neighbours in the grid, so they run in the same wave and share the
  // X / dY lines through L2 instead of re-streaming them from HBM (round 1: 2.4x the algorithmic DRAM bytes)
  // Two regions: CTAs of the FULL ci tiles (CA atoms) first, then those of the partial last ci tile (fewer atoms:
  // less work per K block), which get correspondingly longer K ranges so that every CTA of the grid
  // runs for about the same time (wgrad_setup sizes both so that the grid fills whole waves).
  // With an order table (wgrad_setup) consecutive groups of ngroups x co_tiles CTAs are (ci tile, K split) entries
  // sorted by the first pixel row they read, partial-tile entries interleaved with the full ones: launched last, as a
  // separate region, the partial tile's CTAs re-read the whole dY tensor from DRAM long after the full tiles streamed
  // it (level 1: 850 MB per launch for 400 MB of operands, profiles/r02_launches_trainstep_traffic.csv).
  int w = blockIdx.x;
  bool part;
  int gi, co_t, ci_t, split;
  if (p.n_entries > 0) {
    const int per = p.ngroups * p.co_tiles;
    const uint32_t ent = p.order[w / per];
    w %= per;
    gi = w % p.ngroups;
    co_t = w / p.ngroups;
    part = (ent >> 31) != 0;
    ci_t = part ? p.ci_tiles_full : (int)((ent >> 16) & 0x7fffu);
    split = (int)(ent & 0xffffu);
  } else {
    const int n_region1 = p.ngroups * p.co_tiles * p.ci_tiles_full * p.splits;
    part = w >= n_region1;
    if (part) w -= n_region1;
    gi = w % p.ngroups;   w /= p.ngroups;
    co_t = w % p.co_tiles; w /= p.co_tiles;
    if (!part) {
      ci_t = w % p.ci_tiles_full;
      split = w / p.ci_tiles_full;
    } else {
      ci_t = p.ci_tiles_full;
      split = w;
    }
  }
  const int kbps = part ? p.kblocks_per_split_part : p.kblocks_per_split;
  const WgradGroup& grp = p.groups[gi];
  const int ci0 = ci_t * CA * 64, co0 = co_t * 128;
  const int kb0 = split * kbps;
  const int kb1 = min(p.kblocks, kb0 + kbps);
  const int nkb = max(0, kb1 - kb0);
  const int ncol = 64 * grp.ntaps;  // accumulator columns per ci atom (MMA N)
  // the last ci tile may hold fewer than CA atoms (181 channels = 3 atoms = tiles of 2 + 1): no loads, MMAs or
  // reductions for atoms that do not exist
  const int CAh = min(CA, p.ci_atoms - ci_t * CA);
  const bool co_hi = co0 + 64 < p.co_valid;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmX);
    prefetch_tmap(&p.tmDY);
  }
  // Column sums of dY (= the bias gradient of the conv whose weight gradient this is) are taken from the dY tiles as
  // they pass through shared memory, by the four warps that otherwise idle until the drain - in the CTAs of the first
  // kernel row and first ci tile only, whose K ranges cover every pixel row exactly once.  A stage of such a CTA is
  // released by the MMA commit AND by these 128 readers.
  const bool do_cs = p.bias_grad != nullptr && !part && gi == 0 && ci_t == 0;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), do_cs ? 129 : 1);
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
  const long long t_setup = prof ? clock64() : 0;

  if (warp == 0) {
    // TMA producer: one elected thread runs the whole role
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int r0 = kb * 64;
        timed_wait(empty_bar(s), ph ^ 1u, prof ? &pw : nullptr);
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        // the second 64-channel half of the dY tile is not fetched when it lies beyond the tensor (its accumulator
        // rows are never stored; rows of an MMA are independent)
        mbar_expect_tx(full_bar(s), (co_hi ? kWgABytes : 8192u) + (uint32_t)CAh * kWgAtomBytes);
        tma_load_2d(&p.tmDY, full_bar(s), st, co0, r0 + grp.dy_off);
        if (co_hi) tma_load_2d(&p.tmDY, full_bar(s), st + 8192u, co0 + 64, r0 + grp.dy_off);
        for (int a = 0; a < CAh; ++a)
          tma_load_2d(&p.tmX, full_bar(s), st + kWgABytes + (uint32_t)a * kWgAtomBytes, ci0 + a * 64,
                      r0 + grp.x_off);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (prof) {
        p.dbg[0] = pw;
        p.dbg[5] = t_setup - t_start;
        p.dbg[6] = nkb;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // MMA issuer: one elected thread; descriptor high words hoisted (see mtgemm_fwd_kernel)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, ncol, 1, 1);
      // K step = 16 pixel rows = 2048 B.  A: co atoms 8192 B apart.  B: "atom" j = slab shifted by
      // j rows (LBO = 128 B) -> the taps of this kernel row side by side along N.
      const uint64_t a_hi = make_desc_sw128(0, 8192, 1024), b_hi = make_desc_sw128(0, 128, 1024);
      const uint32_t lo0 = (smem_base & 0x3FFFFu) >> 4;
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        timed_wait(full_bar(s), ph, prof ? &pw : nullptr);
        tc_fence_after();
        const uint32_t st_lo = lo0 + (uint32_t)s * (stage_bytes >> 4);
        const uint64_t ad = a_hi | (uint64_t)st_lo;
        for (int a = 0; a < CAh; ++a) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * ncol);
          const uint64_t bd = b_hi | (uint64_t)(st_lo + (kWgABytes >> 4) + (uint32_t)a * (kWgAtomBytes >> 4));
          mma_bf16_ss(d_tmem, ad, bd, idesc, acc);
          mma_bf16_ss(d_tmem, ad + 128, bd + 128, idesc, 1u);
          mma_bf16_ss(d_tmem, ad + 256, bd + 256, idesc, 1u);
          mma_bf16_ss(d_tmem, ad + 384, bd + 384, idesc, 1u);
        }
        acc = 1u;
        mma_commit(empty_bar(s));
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      mma_commit(done_bar);
      if (prof) p.dbg[1] = pw;
    }
    __syncwarp();
  } else if (warp >= 4 && do_cs) {
    // dY tile in shared memory: two atoms of [64 pixel rows][64 channels] bf16, 128 B rows, 16-byte chunk c of row r
    // stored at chunk position c ^ (r & 7).  Thread t sums channel chunk t & 15 (atom = chunk >> 3) over the rows
    // (t >> 4) + 8 i: its rows all have the same r & 7, so its physical chunk position is fixed.
    const int t = threadIdx.x - 128;
    const int chunk = t & 15, rg = t >> 4;
    const bool live = co0 + chunk * 8 < p.co_valid;  // (the second atom is not even loaded when it lies beyond the tensor)
    const uint32_t toff = (uint32_t)(chunk >> 3) * 8192u + (uint32_t)rg * 128u + (uint32_t)(((chunk & 7) ^ rg) << 4);
    float cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.f;
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(full_bar(s), ph);
      if (live) {
        const uint8_t* st = smem_raw + (smem_base + (uint32_t)s * stage_bytes - smem_u32(smem_raw)) + toff;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = *reinterpret_cast<const uint4*>(st + i * 1024);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            cs[2 * j] += f.x;
            cs[2 * j + 1] += f.y;
          }
        }
      }
      mbar_arrive(empty_bar(s));
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    // row groups rg and rg ^ 1 sit 16 lanes apart in the same warp
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
    if (live && lane < 16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(p.bias_grad + co0 + chunk * 8 + j, cs[j]);
    }
  }

  // Accumulator drain by ALL eight warps (the producer / MMA / allocator warps have nothing left to do): warp w reads
  // TMEM lane quarter w % 4 and every second 32-column block (w / 4), two tcgen05.ld in flight per wait, and adds the
  // values into the flat gradient with 16-byte vector reductions.  A CTA's drain is serial with its main loop (384
  // accumulator columns leave no room for a second buffer), so its length is pure overhead: 13-22k cycles with four
  // warps and one load per wait (profiles/r02_perf_wgrad_waves.txt).
  {
    const int q = warp & 3, hsel = warp >> 2;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const long long t_done = (prof && warp == 4 && lane == 0) ? clock64() : 0;
    const int co = co0 + q * 32 + lane;
    const bool warp_live = co0 + q * 32 < p.co_valid;
    if (nkb > 0 && warp_live) {
      for (int a = 0; a < CAh; ++a) {
        for (int j = 0; j < grp.ntaps; ++j) {
          // tap j sits at slab shift grp.shift[j] (shifts are consecutive 0..ntaps-1 by construction)
          const int tw = grp.w_idx[j];
          const int cb = hsel * 32;  // this warp's 32-column block of the tap's 64 channels
          const uint32_t t_row =
              tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * ncol + grp.shift[j] * 64 + cb);
          float* row = p.dW + ((size_t)tw * p.w_rows_per_tap + co) * p.ldw + p.dw_col0 + ci0 + a * 64 + cb;
          const int ci_base = ci0 + a * 64 + cb;
          if (ci_base >= p.ci_valid) continue;  // (warp-uniform)
          uint32_t r[2][16];
          tmem_ld16(t_row, r[0]);
          tmem_ld16(t_row + 16u, r[1]);
          tmem_ld_wait();
          if (co < p.co_valid) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int v4 = 0; v4 < 4; ++v4) {
                const int ci = ci_base + h * 16 + v4 * 4;
                if (ci < p.ci_valid)  // ci_valid is a multiple of 8
                  red_add_v4(row + h * 16 + v4 * 4, __uint_as_float(r[h][v4 * 4 + 0]), __uint_as_float(r[h][v4 * 4 + 1]),
                             __uint_as_float(r[h][v4 * 4 + 2]), __uint_as_float(r[h][v4 * 4 + 3]));
              }
          }
        }
      }
    }
    if (prof && warp == 4 && lane == 0) {
      p.dbg[2] = t_done - t_start;
      p.dbg[3] = clock64() - t_done;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (prof && threadIdx.x == 0) p.dbg[4] = clock64() - t_start;
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

// The encoded descriptor is a pure function of its arguments, and a U-Net schedule asks for the same ~250 descriptors
// every step: they are memoised (per process, mutex-protected) so a GEMM launch costs no driver call after the first step.
struct TmapKey {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_cols, box_rows;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2));
    h ^= (k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2));
    h ^= (((uint64_t)k.box_cols << 32 | k.box_rows) + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_cols, uint32_t box_rows) {
  const TmapKey key{base, rows, cols, ld_elems, box_cols, box_rows};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return MPU_OK;
    }
  }
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
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();  // long-lived processes that keep reallocating tensors
  g_tmap_cache.emplace(key, *out);
  return MPU_OK;
}

int num_sms() { return sm_count(); }

// cudaFuncSetAttribute is per device: one flag per device ordinal
static bool g_fwd_attr_set[64] = {false}, g_wgrad_attr_set[64] = {false};
static int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64) ? dev : 0;
}
typedef void (*FwdKernel)(const FwdParams);
// clm: cluster mode 0 / 1 (slab sharing) / 2 (weight sharing)
static FwdKernel fwd_kernel_for(bool w_mn, bool prof, int clm, bool red) {
  if (red) return w_mn ? mtgemm_fwd_kernel<true, false, 0, true> : mtgemm_fwd_kernel<false, false, 0, true>;
  if (clm == 1) {
    if (prof) return w_mn ? mtgemm_fwd_kernel<true, true, 1, false> : mtgemm_fwd_kernel<false, true, 1, false>;
    return w_mn ? mtgemm_fwd_kernel<true, false, 1, false> : mtgemm_fwd_kernel<false, false, 1, false>;
  }
  if (clm == 2) {
    if (prof) return w_mn ? mtgemm_fwd_kernel<true, true, 2, false> : mtgemm_fwd_kernel<false, true, 2, false>;
    return w_mn ? mtgemm_fwd_kernel<true, false, 2, false> : mtgemm_fwd_kernel<false, false, 2, false>;
  }
  if (prof) return w_mn ? mtgemm_fwd_kernel<true, true, 0, false> : mtgemm_fwd_kernel<false, true, 0, false>;
  return w_mn ? mtgemm_fwd_kernel<true, false, 0, false> : mtgemm_fwd_kernel<false, false, 0, false>;
}
// clusters of 2 that can be co-resident with this kernel's shared-memory footprint (per device; 0 = cluster launch
// not possible, fall back to independent CTAs)
static int g_max_clusters[64] = {0};
static bool g_max_clusters_set[64] = {false};
static constexpr int kDefaultSmemReserveKB = 0;
static constexpr int kDynSmem = 232448 - 1024;  // leave room for static smem (none) and the driver

int pick_bn(int n_valid) { return (n_valid + 15) / 16 * 16 > 256 ? 256 : (n_valid + 15) / 16 * 16; }

// bring-up knobs
static int smem_reserve();
static int g_fwd_no_slab = 0;
static int g_fwd_cl2 = 1;   // MPU_FWD_CL2: 0 = never launch slab-sharing clusters, 1 = whenever the conv has an even
                            // number of channel tiles (default: +5 % at level 1, +2 % at level 3, neutral at 16x16;
                            // profiles/r02_perf_gemm_cluster.txt), 2 = only where the 9-tap slab does not fit
static int g_fwd_clw = 0;   // MPU_FWD_CLW: weight-sharing clusters (see launch_fwd); measured neutral -> off by default
static int g_fwd_wide = 1;  // MPU_FWD_WIDE=0: never build slabs taller than 264 rows (bring-up comparison)
static long long* g_fwd_dbg = nullptr;
extern "C" void mpu_debug_set_fwd_mode(int no_slab) { g_fwd_no_slab = no_slab; }
// device pointer to 8 long longs: role stall cycles of CTA 0 for the next forward-type launches
extern "C" void mpu_debug_set_fwd_profile(long long* dev_counters) { g_fwd_dbg = dev_counters; }

int fwd_setup(FwdParams& p, const FwdDesc& d) {
  static bool env_read = false;
  if (!env_read) {  // bring-up override without a rebuild
    if (const char* e = getenv("MPU_FWD_NO_SLAB")) g_fwd_no_slab = atoi(e);
    if (const char* e = getenv("MPU_FWD_WIDE")) g_fwd_wide = atoi(e);
    if (const char* e = getenv("MPU_FWD_CL2")) g_fwd_cl2 = atoi(e);
    if (const char* e = getenv("MPU_FWD_CLW")) g_fwd_clw = atoi(e);
    env_read = true;
  }
  memset(&p, 0, sizeof(p));
  if (!d.A0 || !d.W || !d.out || d.ntaps < 1 || d.ntaps > kMaxTaps) {
    set_error("fwd_setup: bad arguments");
    return MPU_ERR_ARG;
  }
  MPU_TRY(make_tmap_2d(&p.tmA0_hi, d.A0, (uint64_t)d.rowsA0, (uint64_t)d.C0, (uint64_t)d.ldA0, 64, 136));
  MPU_TRY(make_tmap_2d(&p.tmA0_lo, d.A0, (uint64_t)d.rowsA0, (uint64_t)d.C0, (uint64_t)d.ldA0, 64, 128));
  p.chunks0 = (d.C0 + 63) / 64;
  p.c0_valid = d.C0;
  p.c1_valid = d.C1;
  if (d.A1) {
    MPU_TRY(make_tmap_2d(&p.tmA1_hi, d.A1, (uint64_t)d.rowsA1, (uint64_t)d.C1, (uint64_t)d.ldA1, 64, 136));
    MPU_TRY(make_tmap_2d(&p.tmA1_lo, d.A1, (uint64_t)d.rowsA1, (uint64_t)d.C1, (uint64_t)d.ldA1, 64, 128));
    p.chunks1 = (d.C1 + 63) / 64;
    p.kofs1 = d.C0;
  }
  p.w_mn = d.w_mn ? 1 : 0;
  // A conv with a single, narrow channel tile (level 0: 96 physical channels) loads weight tiles of exactly its own
  // rows and stages its outputs at its own width: 12 instead of 16 KB per tile and per 64-pixel staging pass, which is
  // what pays for a third activation slab there.  (The MMA still reads 128 A rows; accumulator rows beyond the tensor
  // are never stored, and rows of an MMA are independent.)
  const int narrow = (d.n_phys < 128 && d.n_phys % 8 == 0) ? d.n_phys : 128;
  p.stg_ch = narrow;
  p.w_tile_bytes = d.w_mn ? (int)kWTileBytes : narrow * 128;
  if (!d.w_mn) {
    MPU_TRY(make_tmap_2d(&p.tmW, d.W, (uint64_t)d.w_taps * d.n_phys, (uint64_t)d.k_total,
                         (uint64_t)d.k_total, 64, narrow));
  } else {
    MPU_TRY(make_tmap_2d(&p.tmW, d.W, (uint64_t)d.w_taps * d.w_rows, (uint64_t)d.n_phys,
                         (uint64_t)d.n_phys, 64, 64));
  }
  // Tap grouping.  Taps sorted by row offset; a group = consecutive taps whose offsets span at most S rows: they share
  // one slab of 256 + roundup8(S + 1) rows and differ only in the descriptor's row shift.  S = 7 puts the 3 kx taps of
  // a kernel row into one 264-row slab (3 slabs per K chunk of a 3x3 conv).  Where the padded image rows are short a
  // larger S covers SEVERAL kernel rows: the candidate spans are the pairwise offset differences; the one with the
  // fewest slab rows per K chunk that still leaves room for two slabs and three weight stages wins (ties: smaller S).
  int order[kMaxTaps];
  for (int i = 0; i < d.ntaps; ++i) order[i] = i;
  for (int i = 1; i < d.ntaps; ++i)
    for (int j = i; j > 0 && d.tap_a_off[order[j]] < d.tap_a_off[order[j - 1]]; --j) {
      const int t = order[j];
      order[j] = order[j - 1];
      order[j - 1] = t;
    }
  auto slab_rows_for = [](int S) { return 256 + ((S + 1 + 7) / 8) * 8; };
  auto count_groups = [&](int S, int max_taps) {
    int n = 0, i = 0;
    while (i < d.ntaps) {
      const int o = d.tap_a_off[order[i]];
      int k = 0;
      while (i < d.ntaps && k < max_taps && d.tap_a_off[order[i]] - o <= S) { ++i; ++k; }
      ++n;
    }
    return n;
  };
  // shared memory left for slabs with 3 weight stages: staging of 64-pixel passes, or 32-pixel passes if that is
  // what makes the tall slab fit
  const int budget = kSmemBudget - smem_reserve();
  auto fits = [&](int S, int stg_px) {
    return 2 * slab_rows_for(S) * 128 + 3 * 2 * p.w_tile_bytes + 2 * stg_px * p.stg_ch * 2 + 1024 <= budget;
  };
  int bestS = 7;
  long long best_cost = (long long)count_groups(7, 3) * slab_rows_for(7);
  if (!g_fwd_no_slab && g_fwd_wide) {
    for (int i = 0; i < d.ntaps; ++i)
      for (int j = i + 1; j < d.ntaps; ++j) {
        const int S = d.tap_a_off[order[j]] - d.tap_a_off[order[i]];
        if (S <= 7 || !fits(S, 32)) continue;
        const long long cost = (long long)count_groups(S, kMaxGroupTaps) * slab_rows_for(S);
        if (cost * 10 < best_cost * 9 || (bestS > 7 && (cost < best_cost || (cost == best_cost && S < bestS)))) {
          best_cost = cost;
          bestS = S;
        }
      }
  }
  const int max_taps = bestS > 7 ? kMaxGroupTaps : 3;
  p.slab_rows = slab_rows_for(bestS);
  p.ext_rows = p.slab_rows - (int)kSlabRows;
  p.stg_px = fits(bestS, 64) ? 64 : 32;
  if (p.ext_rows > 0) {
    MPU_TRY(make_tmap_2d(&p.tmA0_ext, d.A0, (uint64_t)d.rowsA0, (uint64_t)d.C0, (uint64_t)d.ldA0, 64, p.ext_rows));
    if (d.A1)
      MPU_TRY(make_tmap_2d(&p.tmA1_ext, d.A1, (uint64_t)d.rowsA1, (uint64_t)d.C1, (uint64_t)d.ldA1, 64, p.ext_rows));
  }
  p.ngroups = 0;
  int i = 0;
  while (i < d.ntaps) {
    TapGroup& G = p.groups[p.ngroups++];
    G.a_off = d.tap_a_off[order[i]];
    G.ntaps = 0;
    while (i < d.ntaps && G.ntaps < max_taps && d.tap_a_off[order[i]] - G.a_off <= bestS &&
           (!g_fwd_no_slab || G.ntaps < 1)) {
      G.shift[G.ntaps] = d.tap_a_off[order[i]] - G.a_off;
      G.w_idx[G.ntaps] = d.tap_w[order[i]];
      ++G.ntaps;
      ++i;
    }
  }
  p.w_rows_per_tap = d.w_mn ? d.w_rows : d.n_phys;
  p.M_rows = d.M_rows;
  p.n_valid = d.n_phys;
  p.map = d.map;
  p.out = reinterpret_cast<__nv_bfloat16*>(d.out);
  p.ldo = d.ldo;
  p.bias = d.bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(d.mask);
  p.ldm = d.ldm;
  p.relu = d.relu;
  p.stats = d.stats;
  p.csum_f = d.csum_f;
  p.red_d = d.red_d;
  p.red_y = reinterpret_cast<const __nv_bfloat16*>(d.red_y);
  p.red_ldy = d.red_ldy;
  p.red_col0 = d.red_col0;
  p.red_C = d.red_C;
  if (d.red_d && (d.red_col0 % 8 || d.red_C % 8 || (d.red_y && d.red_ldy % 8))) {
    set_error("fwd_setup: epilogue reduction columns must be multiples of 8");
    return MPU_ERR_ARG;
  }
  p.dbg = g_fwd_dbg;
  if (const char* e = getenv("MPU_FWD_DEBUG")) p.dbg_flags = atoi(e);
  const long long imgs = (d.M_rows + (long long)d.map.Hp * d.map.Wp - 1) / ((long long)d.map.Hp * d.map.Wp);
  if ((long long)d.map.oHp * d.map.oWp * imgs > 2147483647LL) {
    set_error("fwd_setup: output row index exceeds 32 bits");
    return MPU_ERR_ARG;
  }
  return MPU_OK;
}

// Shared memory the GEMM CTAs leave free on their SM so that blocks of the HBM-bound elementwise kernels running
// on the other stream can be co-resident (tensor pipe and LSU then work at the same time).
static int smem_reserve() {
  static int r = -1;
  if (r < 0) {
    const char* e = getenv("MPU_SMEM_RESERVE_KB");
    r = (e ? atoi(e) : kDefaultSmemReserveKB) * 1024;
    if (r < 0) r = 0;
    if (r > 96 * 1024) r = 96 * 1024;
  }
  return r;
}

int launch_fwd(FwdParams& p, cudaStream_t stream) {
  // Two slabs + three weight stages + 64-pixel epilogue passes (fwd_setup picked slab height and pass size).
  // Measured alternative (MPU_FWD_NA=3 MPU_FWD_SP=32: a third slab paid for by 32-pixel passes): the MMA thread's
  // a_full waits drop from 12-20 % to 4 %, but its w_full waits rise by as much - no net gain on levels 2-3, 5 %
  // slower at 16x16 (profiles/r02_perf_gemm_pairs.txt).
  // Round 2, session 2 (profiles/r02_perf_gemm_epilogue.txt): where a slab serves only the 3 taps of one kernel row
  // (levels 0-1: 6-12 MMAs per slab) two slab slots do not cover the refill round trip - the MMA thread waited 14-23 %
  // of its time on a_full - so those convs take a THIRD slab whenever three weight stages still fit, with 32-pixel
  // staging passes if that is what it takes.
  const int budget = kSmemBudget - smem_reserve();
  const int stage_bytes = 2 * p.w_tile_bytes;
  auto fixed_for = [&](int sp) { return 2 * sp * p.stg_ch * 2 + 1024; };  // epilogue staging + row table
  auto nw_for = [&](int na, int sp) { return (budget - fixed_for(sp) - na * p.slab_rows * 128) / stage_bytes; };
  int NA = 2;
  if (const char* e = getenv("MPU_FWD_NA")) {  // bring-up override
    NA = atoi(e);
    if (NA >= 3 && p.ext_rows == 0) p.stg_px = 32;
  } else if (p.ext_rows == 0) {
    if (nw_for(3, p.stg_px) >= 3) {
      NA = 3;
    } else if (nw_for(3, 32) >= 3) {
      NA = 3;
      p.stg_px = 32;
    }
  }
  if (const char* e = getenv("MPU_FWD_SP")) p.stg_px = atoi(e) == 32 ? 32 : 64;
  int NW = nw_for(NA, p.stg_px);   // stages of two weight tiles
  if (NW > 8) NW = 8;
  p.items_per_tile = 0;
  for (int g = 0; g < p.ngroups; ++g) p.items_per_tile += p.groups[g].ntaps * (p.chunks0 + p.chunks1);
  if (NA < 1 || NW < 2) {
    set_error("launch_fwd: not enough shared memory (NA=%d NW=%d)", NA, NW);
    return MPU_ERR_ARG;
  }
  p.a_slots = NA;
  p.w_slots = NW;
  p.m_tiles = (p.M_rows + kPT - 1) / kPT;
  p.n_tiles = (p.n_valid + 127) / 128;
  if (!g_fwd_attr_set[cur_dev()]) {
    for (int i = 0; i < 4; ++i)
      for (int clm = 0; clm < 3; ++clm)
        MPU_CUDA(cudaFuncSetAttribute(fwd_kernel_for(i & 1, i & 2, clm, false),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
    for (int i = 0; i < 2; ++i)
      MPU_CUDA(cudaFuncSetAttribute(fwd_kernel_for(i & 1, false, 0, true),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
    g_fwd_attr_set[cur_dev()] = true;
  }
  const int smem = kDynSmem - smem_reserve();
  // Cluster modes (pairs of CTAs sharing one operand stream through TMA multicast).  1: the two channel tiles of a pixel
  // tile share every activation slab - needs an even number of channel tiles (levels 1, 3, 4 at complexity_factor 2).
  // 2: two adjacent pixel tiles of one channel tile share the weight stream - used where there is a single channel
  // tile (level 0), whose weights are otherwise re-fetched for every pixel tile (MPU_FWD_CLW: 0 = never [default],
  // 1 = single channel tile, 2 = also instead of mode 1 on the 3-tap-slab levels).  Built, parity-green and measured
  // NEUTRAL (profiles/r02_perf_gemm_weight_sharing.txt: level 0 853 vs 847 TFLOP/s, 180->90 938 vs 932, level 1 with
  // mode 2 instead of mode 1 899 vs 907; step 26.47 vs 26.33-26.44 ms): the 7-8 % the MMA thread waits for weights at
  // level 0 is ring latency, not fetch bandwidth.  Off by default: independent CTAs do not run in lockstep.
  const bool red = p.csum_f != nullptr || p.red_d != nullptr;  // (own instantiation: neither profiled nor clustered)
  int clm = 0;
  if (!red && p.n_tiles >= 2 && (p.n_tiles & 1) == 0 && (g_fwd_cl2 == 1 || (g_fwd_cl2 == 2 && p.ext_rows == 0)))
    clm = 1;
  if (!red && g_fwd_clw >= 1 && p.n_tiles == 1 && p.m_tiles >= 2 && p.items_per_tile >= 2) clm = 2;
  if (!red && g_fwd_clw >= 2 && p.m_tiles >= 2 && p.items_per_tile >= 2 && p.ext_rows == 0) clm = 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kFwdThreads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  if (clm) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int dev = cur_dev();
    if (!g_max_clusters_set[dev]) {
      cfg.gridDim = dim3(2 * (num_sms() / 2), 1, 1);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fwd_kernel_for(p.w_mn != 0, false, 1, false), &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
      }
      g_max_clusters[dev] = n;
      g_max_clusters_set[dev] = true;
    }
    if (g_max_clusters[dev] < 1) clm = 0;
  }
  const int units = clm == 1 ? p.m_tiles * (p.n_tiles / 2)
                             : (clm == 2 ? ((p.m_tiles + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles);
  int grid;
  if (clm) {
    const int nc = units < g_max_clusters[cur_dev()] ? units : g_max_clusters[cur_dev()];
    grid = 2 * nc;
  } else {
    grid = units < num_sms() ? units : num_sms();
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
  }
  cfg.gridDim = dim3(grid, 1, 1);
  gemm_timer_begin(stream);
  MPU_CUDA(cudaLaunchKernelEx(&cfg, fwd_kernel_for(p.w_mn != 0, p.dbg != nullptr && !red, clm, red), p));
  gemm_timer_end(stream);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

int wgrad_setup(WgradParams& p, const WgradDesc& d) {
  memset(&p, 0, sizeof(p));
  if (!d.X || !d.dY || !d.dW || d.ntaps < 1 || d.ntaps > kMaxTaps) {
    set_error("wgrad_setup: bad arguments");
    return MPU_ERR_ARG;
  }
  if ((d.ldw % 4) || (d.dw_col0 % 4)) {
    set_error("wgrad_setup: dW row stride / column offset must be multiples of 4 floats");
    return MPU_ERR_ARG;
  }
  MPU_TRY(make_tmap_2d(&p.tmX, d.X, (uint64_t)d.rowsX, (uint64_t)d.Cx, (uint64_t)d.ldX, 64, kWgSlabRows));
  MPU_TRY(make_tmap_2d(&p.tmDY, d.dY, (uint64_t)d.rowsDY, (uint64_t)d.Cy, (uint64_t)d.ldDY, 64, 64));
  return wgrad_plan(p, d);
}

// Work decomposition of one weight-gradient GEMM (tap groups, tiles, split-K sizing, launch order): pure host
// arithmetic on the shapes - no device, no tensor maps - so that it can be tested on its own (mpu_debug_wgrad_plan,
// tests/test_cabi.py).
int wgrad_plan(WgradParams& p, const WgradDesc& d) {
  // sort taps by (dy_off, x_off); a group = taps of one dY plane at CONSECUTIVE rows (shift 0,1,2):
  // they become the N atoms of one MMA
  int order[kMaxTaps];
  for (int i = 0; i < d.ntaps; ++i) order[i] = i;
  auto less = [&](int a, int b) {
    const int da = d.tap_dy_off ? d.tap_dy_off[a] : 0, db = d.tap_dy_off ? d.tap_dy_off[b] : 0;
    if (da != db) return da < db;
    return d.tap_x_off[a] < d.tap_x_off[b];
  };
  for (int i = 1; i < d.ntaps; ++i)
    for (int j = i; j > 0 && less(order[j], order[j - 1]); --j) {
      const int t = order[j];
      order[j] = order[j - 1];
      order[j - 1] = t;
    }
  p.ngroups = 0;
  int i = 0;
  while (i < d.ntaps) {
    WgradGroup& G = p.groups[p.ngroups++];
    G.x_off = d.tap_x_off[order[i]];
    G.dy_off = d.tap_dy_off ? d.tap_dy_off[order[i]] : 0;
    G.ntaps = 0;
    while (i < d.ntaps && G.ntaps < 3 && d.tap_x_off[order[i]] - G.x_off == G.ntaps &&
           (d.tap_dy_off ? d.tap_dy_off[order[i]] : 0) == G.dy_off) {
      G.shift[G.ntaps] = G.ntaps;
      G.w_idx[G.ntaps] = d.tap_w[order[i]];
      ++G.ntaps;
      ++i;
    }
  }
  p.ci_valid = d.Cx;
  p.co_valid = d.Cy;
  const int atoms = (d.Cx + 63) / 64;
  p.ci_atoms = atoms;
  p.CA = atoms >= 2 ? 2 : 1;
  p.ci_tiles = (atoms + p.CA - 1) / p.CA;
  p.co_tiles = (d.Cy + 127) / 128;
  p.kblocks = (int)((d.rows_total + 63) / 64);
  // Split-K sizing.  A CTA owns its SM (shared memory + 384-512 TMEM columns), so the grid runs in waves of num_sms
  // CTAs and a wave lasts as long as its slowest CTA.  Measured (profiles/r02_perf_wgrad_waves.txt): the main loop
  // runs at 98-99 % of the tensor pipe (777-788 cycles per K block of 8 N=192 MMAs), but round 1's rule "about 2 CTAs
  // per SM" produced 297 CTAs on 148 SMs for two of the five levels - a third, almost empty wave, i.e. 2/3 of the rate.
  // Now: the partial last ci tile (181 channels = 2 + 1 atoms) gets K ranges longer by CA / its atom count, and the K
  // range L per full CTA is the one that minimises  waves(L) x (L x cycles per K block + per-CTA overhead)  over
  // the wave counts 1..12 (overhead = prologue + accumulator drain, ~12k cycles).
  const int part_atoms = atoms % p.CA;                     // atoms of the partial ci tile (0: none)
  p.ci_tiles_full = atoms / p.CA;
  const int units_full = p.ci_tiles_full * p.co_tiles * p.ngroups;
  const int units_part = part_atoms ? p.co_tiles * p.ngroups : 0;
  auto ceil_div = [](long long a, long long b) { return (int)((a + b - 1) / b); };
  // cycles per K block of a CTA with n atoms: the larger of its MMA time and of its operand bytes at the ~45 B/cycle a
  // single SM pulls through TMA (a one-atom CTA still loads the whole 16 KB dY tile: it is load-bound at ~570 cycles,
  // not the 384 its MMAs need)
  int max_taps = 1;
  for (int g = 0; g < p.ngroups; ++g) max_taps = p.groups[g].ntaps > max_taps ? p.groups[g].ntaps : max_taps;
  auto cyc_kb_for = [&](int n_atoms) {
    const double mma = n_atoms * 4.0 * (max_taps >= 2 ? 32.0 * max_taps : 64.0);
    const double load = ((d.Cy > 64 ? 16384.0 : 8192.0) + n_atoms * (double)kWgAtomBytes) / 45.0;
    return mma > load ? mma : load;
  };
  const double cyc_kb = cyc_kb_for(p.CA);
  auto part_len = [&](int L) { return part_atoms ? (int)(L * cyc_kb / cyc_kb_for(part_atoms)) : L; };
  auto ctas_for = [&](int L) {
    return (long long)units_full * ceil_div(p.kblocks, L) +
           (units_part ? (long long)units_part * ceil_div(p.kblocks, part_len(L)) : 0);
  };
  int L;
  static int old_rule = -1;  // MPU_WG_OLD_SPLITS=1: round 1's "about two CTAs per SM" (same-box A/B measurements)
  if (old_rule < 0) {
    const char* e = getenv("MPU_WG_OLD_SPLITS");
    old_rule = e ? atoi(e) : 0;
  }
  const int fixed_splits = d.splits > 0 ? d.splits
                           : (old_rule ? (2 * num_sms() + p.ci_tiles * p.co_tiles * p.ngroups - 1) /
                                             (p.ci_tiles * p.co_tiles * p.ngroups)
                                       : 0);
  if (fixed_splits > 0) {  // caller-fixed split count (tests): the same K range for every tile
    int splits = fixed_splits > p.kblocks ? p.kblocks : fixed_splits;
    L = ceil_div(p.kblocks, splits);
    p.kblocks_per_split = L;
    p.kblocks_per_split_part = L;
  } else {
    const double overhead = 12000.0;
    const int sms = num_sms();
    double best_t = 1e300;
    L = p.kblocks;
    for (int k = 1; k <= 12; ++k) {
      int lo = 1, hi = p.kblocks;  // smallest K range whose grid fits k waves
      if (ctas_for(hi) > (long long)k * sms) continue;
      while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (ctas_for(mid) <= (long long)k * sms) hi = mid;
        else lo = mid + 1;
      }
      const double t = (double)ceil_div(ctas_for(lo), sms) * (lo * cyc_kb + overhead);
      if (t < best_t * 0.999) {
        best_t = t;
        L = lo;
      }
    }
    p.kblocks_per_split = L;
    p.kblocks_per_split_part = part_len(L) > p.kblocks ? p.kblocks : part_len(L);
  }
  p.splits = ceil_div(p.kblocks, p.kblocks_per_split);
  p.splits_part = units_part ? ceil_div(p.kblocks, p.kblocks_per_split_part) : 0;
  // launch order of the (ci tile, K split) entries: by first K block, partial-tile entries interleaved
  p.n_entries = 0;
  {
    static int no_order = -1;  // MPU_WG_NO_ORDER=1: partial-tile CTAs as a separate region at the end (same-box A/B)
    if (no_order < 0) {
      const char* e = getenv("MPU_WG_NO_ORDER");
      no_order = e ? atoi(e) : 0;
    }
    const int n = p.ci_tiles_full * p.splits + p.splits_part;
    if (!no_order && p.splits_part > 0 && n <= kWgMaxEntries && p.splits < 65536 && p.ci_tiles_full < 32768) {
      struct Ent { int kb0; uint32_t code; };
      Ent ents[kWgMaxEntries];
      int m = 0;
      for (int sp = 0; sp < p.splits; ++sp)
        for (int ct = 0; ct < p.ci_tiles_full; ++ct)
          ents[m++] = {sp * p.kblocks_per_split, ((uint32_t)ct << 16) | (uint32_t)sp};
      for (int sp = 0; sp < p.splits_part; ++sp)
        ents[m++] = {sp * p.kblocks_per_split_part, 0x80000000u | (uint32_t)sp};
      for (int i = 1; i < m; ++i)  // stable insertion sort by first K block (full tiles before the partial one on ties)
        for (int j = i; j > 0 && ents[j].kb0 < ents[j - 1].kb0; --j) {
          const Ent t = ents[j];
          ents[j] = ents[j - 1];
          ents[j - 1] = t;
        }
      for (int i = 0; i < m; ++i) p.order[i] = ents[i].code;
      p.n_entries = m;
    }
  }
  p.dW = d.dW;
  p.ldw = d.ldw;
  p.w_rows_per_tap = d.w_rows_per_tap;
  p.dw_col0 = d.dw_col0;
  p.bias_grad = d.bias_grad;
  if (d.bias_grad) {
    for (int g = 0; g < p.ngroups; ++g)
      if (p.groups[g].dy_off != 0) {
        set_error("wgrad_setup: bias_grad needs an unshifted dY (no phase planes)");
        return MPU_ERR_ARG;
      }
    if (p.ci_tiles_full < 1) {
      set_error("wgrad_setup: bias_grad needs at least one full ci tile");
      return MPU_ERR_ARG;
    }
  }
  return MPU_OK;
}

int launch_wgrad(WgradParams& p, cudaStream_t stream) {
  const int stage_bytes = (int)kWgABytes + p.CA * (int)kWgAtomBytes;
  int S = (kSmemBudget - smem_reserve()) / stage_bytes;
  if (S > 8) S = 8;
  if (S < 2) {
    set_error("launch_wgrad: not enough shared memory for 2 stages");
    return MPU_ERR_ARG;
  }
  p.stages = S;
  if (!g_wgrad_attr_set[cur_dev()]) {
    MPU_CUDA(cudaFuncSetAttribute(mtgemm_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
    MPU_CUDA(cudaFuncSetAttribute(mtgemm_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
    g_wgrad_attr_set[cur_dev()] = true;
  }
  const int grid = p.co_tiles * p.ngroups * (p.ci_tiles_full * p.splits + (p.ci_tiles - p.ci_tiles_full) * p.splits_part);
  p.dbg = g_fwd_dbg;
  gemm_timer_begin(stream);
  if (p.dbg) mtgemm_wgrad_kernel<true><<<grid, kThreads, kDynSmem - smem_reserve(), stream>>>(p);
  else mtgemm_wgrad_kernel<false><<<grid, kThreads, kDynSmem - smem_reserve(), stream>>>(p);
  gemm_timer_end(stream);
  count_launch();
  MPU_CUDA(cudaGetLastError());
  return MPU_OK;
}

}  // namespace mpu
