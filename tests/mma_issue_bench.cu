// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, K = 16, both operands in shared
// memory, 128B swizzle, K-major) as a function of the instruction shape and of concurrent TMA traffic.
//   cta_group::1 : M = 128, N in {64, 96, 128, 192, 256}
//   cta_group::2 : M = 256 (128 rows per CTA of the pair), N in {64, 96, 128, 192, 256} (N/2 columns of B per CTA)
// Every SM runs one CTA (148 CTAs; 74 pairs for cta_group::2).  One elected thread issues R back-to-back MMAs that
// accumulate into one TMEM tile while walking a ring of operand stages, then commits and waits; cycles = clock64
// delta / R.  With `tma=1` a second warp streams 16 KB TMA boxes from global memory into a separate shared-memory
// ring for the whole duration (the operand refill traffic of a real GEMM main loop).
// Operand contents are whatever the shared memory holds: only timing is measured.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tests/mma_issue_bench tests/mma_issue_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000ll) return false;  // ~1 s watchdog: never hang the box
  return true;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;   // D fp32
  d |= 1u << 7;   // A bf16
  d |= 1u << 10;  // B bf16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

struct Params {
  int N;            // MMA N (whole instruction)
  int reps;         // MMAs per CTA (pair)
  int tma;          // TMA boxes kept in flight by the streaming warp (0 = off, up to 4)
  int nstage;       // operand ring stages
  int shift;        // 1: the B descriptor of consecutive MMA groups starts 0, 1, 2 rows (128 B) into the tile
                    //    (how one activation slab serves the three kx taps of a kernel row)
  int commit;       // 1: tcgen05.commit to a (never awaited) mbarrier after every group of 4 MMAs
  int epi;          // 1: warps 4-11 stream tcgen05.ld (the epilogue draining the other accumulator) ...
  int epi_smem;     // ... 1: and write/read 16 B per thread per load to shared memory (transposing epilogue)
  int a_mn, b_mn;   // operand read MN-major (dgrad weights / wgrad operands) instead of K-major
  int rnd;          // 1: fill the operand stages with random bf16 values in [-2, 2) first (0: zeros)
  long long* out;   // [grid][4]: cycles, reps, tma boxes, ok
  CUtensorMap tm;   // bf16 [rows][64] box {64, 128} = 16 KB
  int tm_rows;
};

static constexpr uint32_t kATile = 128 * 128;  // 128 rows x 64 bf16
static constexpr uint32_t kTmaBox = 128 * 128;

template <int CG>
__global__ void __launch_bounds__(384, 1) mma_bench_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_rows = (uint32_t)(CG == 2 ? p.N / 2 : p.N);     // rows of B this CTA holds
  const uint32_t b_tile = ((b_rows * 128u) + 1023u) & ~1023u;
  const uint32_t stage_bytes = kATile + b_tile;
  const uint32_t ring = base + (uint32_t)p.nstage * stage_bytes;    // TMA scratch ring (4 boxes)
  const uint32_t epi_buf = ring + 4 * kTmaBox;                      // 8 KB epilogue scratch
  const uint32_t bars = epi_buf + 8192;
  const uint32_t done_bar = bars, tma_bar0 = bars + 8;              // 4 TMA barriers, then 4 commit barriers
  const uint32_t cm_bar0 = bars + 40;
  const uint32_t slot = bars + 80;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + (slot + 8 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;

  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(tma_bar0 + 8 * i, 1);
      mbar_init(cm_bar0 + 8 * i, 1);
    }
    *stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<CG>(slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *slot_ptr;

  // operand contents: zeros or pseudo-random bf16 (sign, exponent 125..128, random mantissa)
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    const uint32_t nwords = (uint32_t)p.nstage * stage_bytes / 4;
    uint32_t x = 0x9E3779B9u * (blockIdx.x * 384u + threadIdx.x + 1u);
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) {
      x ^= x << 13; x ^= x >> 17; x ^= x << 5;
      uint32_t hi = (x & 0x807Fu) | ((125u + ((x >> 8) & 3u)) << 7);
      uint32_t y = x * 2654435761u;
      uint32_t lo = (y & 0x807Fu) | ((125u + ((y >> 8) & 3u)) << 7);
      w[i] = p.rnd ? ((hi << 16) | lo) : 0u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  long long cycles = 0, boxes = 0, ok = 1;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(CG == 2 ? 256 : 128, p.N);
    const long long t0 = clock64();
    int s = 0, sh = 0, cb = 0;
    const uint64_t a_step = p.a_mn ? 128u : 2u, b_step = p.b_mn ? 128u : 2u;  // K = 16 step in 16-byte units
    for (int r = 0; r < p.reps; r += 4) {  // warp-uniform loop, one elected lane issues (as the GEMM kernels do)
      const uint32_t st = base + (uint32_t)s * stage_bytes;
      if (elect_one()) {
        const uint64_t ad = p.a_mn ? make_desc_sw128(st, 8192, 1024) : make_desc_sw128(st, 16, 1024);
        const uint32_t bst = st + kATile + (p.shift ? (uint32_t)sh * 128u : 0u);
        const uint64_t bd = p.b_mn ? make_desc_sw128(bst, 8192, 1024) : make_desc_sw128(bst, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_ss<CG>(tmem_base, ad + (uint64_t)kk * a_step, bd + (uint64_t)kk * b_step, idesc, (r | kk) ? 1u : 0u);
        if (p.commit) mma_commit<CG>(cm_bar0 + 8u * (uint32_t)cb);
      }
      __syncwarp();
      if (++s == p.nstage) s = 0;
      if (++sh == 3) sh = 0;
      cb = (cb + 1) & 3;
    }
    if (elect_one()) mma_commit<CG>(done_bar);
    __syncwarp();
    ok = mbar_wait(done_bar, 0) ? 1 : 0;
    cycles = clock64() - t0;
    *stop_flag = 1;
  } else if (warp == 1 && p.tma) {
    // stream 16 KB boxes, p.tma in flight, until the MMA warp (of the leader CTA) is done
    int ph = 0;
    // a follower CTA of a pair polls the leader's flag through distributed shared memory
    uint32_t flag_addr = slot + 8;
    if (CG == 2) asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(flag_addr) : "r"(slot + 8));
    auto stopped = [&]() {
      uint32_t v;
      asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(flag_addr) : "memory");
      return v != 0;
    };
    int row = (int)((blockIdx.x * 977u) % (uint32_t)(p.tm_rows - 512));
    // prime the ring, then refill each slot as soon as its box has landed
    if (elect_one())
      for (int i = 0; i < p.tma; ++i) {
        mbar_expect_tx(tma_bar0 + 8 * i, kTmaBox);
        tma_load_2d(&p.tm, tma_bar0 + 8 * i, ring + i * kTmaBox, 0, row + 128 * i);
      }
    __syncwarp();
    row += 128 * p.tma;
    while (!stopped() && boxes < (1 << 28)) {
      for (int i = 0; i < p.tma; ++i) {
        if (!mbar_wait(tma_bar0 + 8 * i, ph)) { ok = 0; break; }
        if (elect_one()) {
          mbar_expect_tx(tma_bar0 + 8 * i, kTmaBox);
          tma_load_2d(&p.tm, tma_bar0 + 8 * i, ring + i * kTmaBox, 0, row);
        }
        __syncwarp();
        row += 128;
        if (row > p.tm_rows - 256) row = 0;
        ++boxes;
      }
      ph ^= 1;
      if (!ok) break;
    }
    for (int i = 0; i < p.tma; ++i) mbar_wait(tma_bar0 + 8 * i, ph);  // drain before the CTA exits
  } else if (warp >= 4 && p.epi) {
    // epilogue stand-in: tcgen05.ld 32 lanes x 16 columns from the upper half of TMEM, optional smem round trip
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    const uint32_t my = epi_buf + (uint32_t)(threadIdx.x - 128) * 16u;
    uint32_t acc = 0;
    uint32_t flag_addr = slot + 8;
    if (CG == 2) asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(flag_addr) : "r"(slot + 8));
    for (int it = 0;; ++it) {
      uint32_t v;
      asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(flag_addr) : "memory");
      if (v) break;
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(t_row + (uint32_t)((it & 15) * 16)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
      if (p.epi_smem) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
        uint32_t a0, a1, a2, a3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(my ^ 2048u) : "memory");
        acc ^= a0 ^ a1 ^ a2 ^ a3;
      }
    }
    if (acc == 0x12345678u) p.out[0] = 1;  // keep the loads alive
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc<CG>(tmem_base, 512);
  }
  if (lane == 0 && warp == 0 && rank == 0) {
    p.out[blockIdx.x * 4 + 0] = cycles;
    p.out[blockIdx.x * 4 + 1] = p.reps;
    p.out[blockIdx.x * 4 + 3] = ok;
  }
  if (lane == 0 && warp == 1) p.out[blockIdx.x * 4 + 2] = boxes;
}

// ---- a faithful mini main loop: TMA producer warp -> full/empty mbarrier ring -> MMA warp ---------------------
// Stage = A tile (16 KB, loaded every iteration) + B tile (32 KB, loaded every `b_every`-th iteration, as the
// slab that serves three taps).  `group` MMAs are issued per full-barrier wait (4 = one 64-wide K chunk).
struct PipeParams {
  int reps, nstage, b_every, group, fence;
  long long* out;
  CUtensorMap tm;
  int tm_rows;
};

__global__ void __launch_bounds__(128, 1) pipe_bench_kernel(const __grid_constant__ PipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = kATile + 2 * kTmaBox;
  const uint32_t bars = base + (uint32_t)p.nstage * stage_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (p.nstage + s); };
  const uint32_t done_bar = bars + 8u * (2 * p.nstage);
  const uint32_t slot = done_bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < p.nstage; ++s2) {
      mbar_init(full(s2), 1);
      mbar_init(empty(s2), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<1>(slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *slot_ptr;
  const int iters = p.reps / p.group;
  long long cycles = 0, w_full = 0, w_empty = 0, ok = 1;
  if (warp == 0) {
    int s2 = 0;
    uint32_t ph = 0;
    int row = (int)((blockIdx.x * 977u) % (uint32_t)(p.tm_rows - 1024));
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      if (!mbar_wait(empty(s2), ph ^ 1u)) { ok = 0; break; }
      w_empty += clock64() - t0;
      const uint32_t st = base + (uint32_t)s2 * stage_bytes;
      const bool with_b = (it % p.b_every) == 0;
      if (elect_one()) {
        mbar_expect_tx(full(s2), with_b ? 3 * kTmaBox : kTmaBox);
        tma_load_2d(&p.tm, full(s2), st, 0, row);
        if (with_b) {
          tma_load_2d(&p.tm, full(s2), st + kATile, 0, row + 128);
          tma_load_2d(&p.tm, full(s2), st + kATile + kTmaBox, 0, row + 256);
        }
      }
      __syncwarp();
      row += 384;
      if (row > p.tm_rows - 1024) row = 0;
      if (++s2 == p.nstage) { s2 = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, 256);
    int s2 = 0;
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const long long tw = clock64();
      if (!mbar_wait(full(s2), ph)) { ok = 0; break; }
      w_full += clock64() - tw;
      if (p.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st = base + (uint32_t)s2 * stage_bytes;
      if (elect_one()) {
        const uint64_t ad = make_desc_sw128(st, 16, 1024);
        const uint64_t bd = make_desc_sw128(st + kATile, 16, 1024);
        for (int g = 0; g < p.group; ++g)
          mma_ss<1>(tmem_base, ad + (uint64_t)(2 * (g & 3)), bd + (uint64_t)(2 * (g & 3)), idesc, (it | g) ? 1u : 0u);
        mma_commit<1>(empty(s2));
      }
      __syncwarp();
      if (++s2 == p.nstage) { s2 = 0; ph ^= 1u; }
    }
    if (elect_one()) mma_commit<1>(done_bar);
    __syncwarp();
    if (!mbar_wait(done_bar, 0)) ok = 0;
    cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc<1>(tmem_base, 512);
  }
  if (lane == 0 && warp == 1) {
    p.out[blockIdx.x * 4 + 0] = cycles;
    p.out[blockIdx.x * 4 + 1] = w_full;
    p.out[blockIdx.x * 4 + 3] = ok;
  }
  if (lane == 0 && warp == 0) p.out[blockIdx.x * 4 + 2] = w_empty;
}

static void run_pipe(PipeParams p, int sms, const char* tag) {
  const size_t smem = 232448 - 1024;
  CK(cudaFuncSetAttribute(pipe_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(p.out, 0, sizeof(long long) * 4 * sms));
  for (int it = 0; it < 2; ++it) {
    pipe_bench_kernel<<<sms, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(4 * sms);
  CK(cudaMemcpy(h.data(), p.out, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
  std::vector<double> cyc;
  double wf = 0, we = 0;
  int bad = 0;
  for (int b = 0; b < sms; ++b) {
    cyc.push_back((double)h[b * 4] / p.reps);
    wf += (double)h[b * 4 + 1] / (double)h[b * 4];
    we += (double)h[b * 4 + 2] / (double)h[b * 4];
    bad += h[b * 4 + 3] ? 0 : 1;
  }
  std::sort(cyc.begin(), cyc.end());
  printf("%-34s stages %d, %d MMAs / wait, B every %d | cycles/MMA min %.1f med %.1f max %.1f | MMA warp waits full %.0f%% | "
         "producer waits empty %.0f%%%s\n", tag, p.nstage, p.group, p.b_every, cyc.front(), cyc[cyc.size() / 2], cyc.back(),
         100 * wf / sms, 100 * we / sms, bad ? " [TIMEOUT]" : "");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CG>
static void run(Params p, int sms, const char* tag) {
  const int grid = CG == 2 ? (sms / 2) * 2 : sms;
  const size_t smem = 232448 - 1024;
  CK(cudaFuncSetAttribute(mma_bench_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(p.out, 0, sizeof(long long) * 4 * grid));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) {  // first run warms clocks / caches
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, mma_bench_kernel<CG>, p));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(4 * grid);
  CK(cudaMemcpy(h.data(), p.out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
  std::vector<double> cyc;
  long long boxes = 0, bad = 0;
  for (int b = 0; b < grid; ++b) {
    if (CG == 2 && (b & 1)) { boxes += h[b * 4 + 2]; continue; }
    if (!h[b * 4 + 3]) ++bad;
    cyc.push_back((double)h[b * 4 + 0] / (double)p.reps);
    boxes += h[b * 4 + 2];
  }
  std::sort(cyc.begin(), cyc.end());
  const double med = cyc[cyc.size() / 2];
  const int M = CG == 2 ? 256 : 128;
  const double macs_per_cyc_sm = (double)M * p.N * 16 / med / CG;
  const double tflops = 2.0 * M * p.N * 16.0 * p.reps * (grid / CG) / (ms * 1e-3) / 1e12;
  printf("%-22s cta_group::%d M=%3d N=%3d tma=%d | cycles/MMA min %.1f med %.1f max %.1f | %.0f MAC/cyc/SM | "
         "%.0f TFLOP/s chip (%.3f ms) | TMA %.1f B/cyc/SM%s\n",
         tag, CG, M, p.N, p.tma, cyc.front(), med, cyc.back(), macs_per_cyc_sm, tflops, ms,
         (double)boxes * kTmaBox / grid / (med * p.reps), bad ? "  [TIMEOUT in some CTAs]" : "");
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int reps = argc > 1 ? atoi(argv[1]) : 8192;
  Params p = {};
  p.reps = reps;
  p.nstage = 3;
  CK(cudaMalloc(&p.out, sizeof(long long) * 4 * 256));
  // 256 MB bf16 source for the TMA stream (larger than L2)
  const int rows = 2 * 1024 * 1024;
  void* src = nullptr;
  CK(cudaMalloc(&src, (size_t)rows * 128));
  CK(cudaMemset(src, 0, (size_t)rows * 128));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)fn)(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstr, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    return 1;
  }
  p.tm_rows = rows;
  printf("# %d SMs, %d MMAs per CTA (pair); ideal = 4096 MAC/cyc/SM (128 cycles for M=128 N=256 K=16)\n", sms, reps);
  const int Ns[] = {64, 96, 128, 192, 256};
  printf("# operands all zero\n");
  for (int tma = 0; tma <= 2; tma += 2) {
    p.tma = tma;
    for (int N : Ns) {
      p.N = N;
      run<1>(p, sms, "SS");
    }
    for (int N : Ns) {
      p.N = N;
      run<2>(p, sms, "SS-pair");
    }
  }
  printf("# operands random bf16 in [-2, 2): long run to let the power management settle\n");
  p.rnd = 1;
  p.tma = 2;
  {
    Params q = p;
    q.reps = reps * 16;
    for (int N : Ns) {
      q.N = N;
      run<1>(q, sms, "SS random x16 reps");
    }
    q.N = 256;
    run<2>(q, sms, "SS-pair random x16");
    q.N = 96;
    run<2>(q, sms, "SS-pair random x16");
  }
  // what a real main loop adds, one feature at a time (N = 256)
  printf("# features of the GEMM main loop, one at a time and together (M=128 N=256 unless stated)\n");
  p.N = 256;
  p.tma = 0;
  struct { const char* tag; int tma, shift, commit, epi, epi_smem, a_mn, b_mn, N; } V[] = {
      {"base (random data)", 0, 0, 0, 0, 0, 0, 0, 256},
      {"tma x4", 4, 0, 0, 0, 0, 0, 0, 256},
      {"row-shifted B", 0, 1, 0, 0, 0, 0, 0, 256},
      {"commit / 4 MMAs", 0, 0, 1, 0, 0, 0, 0, 256},
      {"tcgen05.ld x8 warps", 0, 0, 0, 1, 0, 0, 0, 256},
      {"ld + smem round trip", 0, 0, 0, 1, 1, 0, 0, 256},
      {"all (fwd-like)", 4, 1, 1, 1, 1, 0, 0, 256},
      {"A MN-major (dgrad)", 0, 0, 0, 0, 0, 1, 0, 256},
      {"all (dgrad-like)", 4, 1, 1, 1, 1, 1, 0, 256},
      {"A,B MN N=192 (wgrad)", 0, 0, 0, 0, 0, 1, 1, 192},
      {"wgrad-like + tma x4", 4, 0, 1, 0, 0, 1, 1, 192},
  };
  p.rnd = 1;
  for (auto& v : V) {
    p.tma = v.tma; p.shift = v.shift; p.commit = v.commit; p.epi = v.epi; p.epi_smem = v.epi_smem;
    p.a_mn = v.a_mn; p.b_mn = v.b_mn; p.N = v.N;
    run<1>(p, sms, v.tag);
  }
  printf("# mini main loop: TMA producer -> mbarrier ring -> MMA warp (M=128 N=256, operands streamed from a 256 MB buffer)\n");
  {
    PipeParams q = {};
    q.reps = reps;
    q.out = p.out;
    q.tm = p.tm;
    q.tm_rows = p.tm_rows;
    const int cfgs[][4] = {{4, 3, 4, 1}, {4, 3, 4, 0}, {3, 3, 4, 1}, {2, 3, 4, 1}, {4, 1, 4, 1}, {4, 3, 8, 1}, {4, 6, 4, 1}, {4, 1000000, 4, 1}};
    for (auto& c : cfgs) {
      q.nstage = c[0]; q.b_every = c[1]; q.group = c[2]; q.fence = c[3];
      run_pipe(q, sms, c[3] ? "ring (fence after wait)" : "ring (no tcgen05 fence)");
    }
  }
  for (auto& v : V) {
    if (v.a_mn || v.b_mn) continue;
    p.tma = v.tma; p.shift = v.shift; p.commit = v.commit; p.epi = v.epi; p.epi_smem = v.epi_smem;
    p.a_mn = 0; p.b_mn = 0; p.N = v.N;
    run<2>(p, sms, v.tag);
  }
  return 0;
}
