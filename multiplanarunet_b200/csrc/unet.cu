// U-Net engine: owns the layer table, the activation/gradient workspace layout and the kernel schedule
// for forward, train step (forward + sparse-CE + backward) and Adam.  Graph = mpunet/models/unet.py:
// encoder (:114-134), bottom (:136-146), up path (:148-180), 1x1 softmax head (:211-214).
//
// Data layout: every activation is a zero-bordered NHWC bf16 matrix [B*(H+2)*(W+2)][C_phys]
// (C_phys = channels rounded up to 8; padded channels stay exactly zero in forward and backward).
// All convolutions run as multi-tap GEMMs on tcgen05 (mtgemm.cu); BN / pool / head / Adam are the
// HBM-bound kernels of elementwise.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpunet_b200.h"
#include "common.h"
#include "elementwise.cuh"
#include "mtgemm.cuh"

namespace mpu {

typedef __nv_bfloat16 bf16;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct ConvL {
  std::string name;
  int ksize;            // 3, 2 (upsample-conv) or 1 (head)
  int cin, cout;        // logical (Keras) channels; cin counts both concat halves
  int k_phys, co_phys;  // physical K (input channels incl. padding, both concat halves) and outputs
  int c0_phys;          // physical channels of concat source 0 (== k_phys when single source)
  int ntap_master;      // 9 / 4 / 1
  int ntap_gemm;        // 9 / 9 pairs / 1
  long long w_off, b_off;  // float offsets into the flat parameter buffer
  bf16* wf = nullptr;   // forward operand [ntap_gemm][co_phys][k_phys]
  bf16* wd = nullptr;   // dgrad operand   [ntap_gemm][k_phys][co_phys]
};

struct BnL {
  std::string name;
  int c, c_phys;
  long long g_off, b_off;  // gamma / beta in the parameter buffer
  long long m_off, v_off;  // moving mean / variance in the bn_state buffer
  float *scale = nullptr, *shift = nullptr, *mean = nullptr, *rstd = nullptr;
  double* sums = nullptr;  // [2][c_phys]
};

struct Level {
  Geo g;
  int C;  // physical channels at this level
  // encoder / bottom
  bf16 *a1 = nullptr, *a2 = nullptr, *b = nullptr, *pooled = nullptr;  // pooled: geometry of level+1
  // up path (levels 0..depth-1)
  bf16 *u = nullptr, *bn1 = nullptr, *c2 = nullptr, *c3 = nullptr, *bn2 = nullptr;
  // gradients / scratch
  bf16 *gout = nullptr;   // grad wrt this level's block output (bn2_l, or bottom BN for the last level)
  bf16 *s1 = nullptr, *s2 = nullptr, *dcat = nullptr, *dzu = nullptr, *dpool = nullptr;
};

struct UNet {
  MpuUNetConfig cfg;
  int depth;
  int cin_phys;
  std::vector<ConvL> convs;  // enc l: [2l], [2l+1]; bottom: [2d], [2d+1]; up i: [2d+2+3i .. +2]; head last
  std::vector<BnL> bns;      // enc l: [l]; bottom: [d]; up i: [d+1+2i] (BN1), [d+2+2i] (BN2)
  std::vector<Level> lv;     // 0..depth
  long long n_params = 0, n_bn_state = 0;
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr, *bn_state = nullptr;
  char* ws = nullptr;
  long long ws_bytes = 0, ws_used = 0;
  bf16* x_in = nullptr;      // [rows0][cin_phys]
  bf16* shadow = nullptr;    // bf16 copy of the whole flat parameter buffer (written by Adam); the forward
                             // GEMM operand of every 3x3 conv is a view into it
  bool dgrad_mn = true;      // dgrad reads the forward weights MN-major (no transposed copies)
  int bias_colsum_pass = 0;  // 1 (MPU_BIAS_COLSUM): conv1 bias gradients by a separate column-sum pass over dz1
  int red_min_level = 99;    // levels >= this take BatchNorm-backward sums / bias column sums in the GEMM epilogue.
                             // OFF by default: measured slower at every level than the separate HBM passes, which
                             // overlap with the next GEMM, while the epilogue is on these GEMMs' critical path
                             // (MPU_EPI_RED_LEVEL, measured in profiles/r02_epilogue_reductions.txt)
  float* dwc = nullptr;      // scratch for collapsed upsample-conv weight gradients
  long long dwc_floats = 0;
  bool training_buffers = false;
  bool weights_synced = false;
  // Second stream for work that is off the critical path (weight gradients, two of the four upsample-conv
  // phases): its CTAs fill the SMs that the tail of the kernel on the main stream leaves idle.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  // Third stream: the optimizer update of a parameter range whose gradients are final runs while the next backward
  // stage computes (mpu_unet_train_step_adam); HBM-bound Adam under tensor-bound GEMMs.
  cudaStream_t opt = nullptr;
  cudaEvent_t ev_stage = nullptr, ev_opt = nullptr;
  bool overlap = true;
  int blk = 0;  // backward block counter within a stage (ring index of ev_join)

  ~UNet() {
    if (side) cudaStreamDestroy(side);
    if (opt) cudaStreamDestroy(opt);
    if (ev_stage) cudaEventDestroy(ev_stage);
    if (ev_opt) cudaEventDestroy(ev_opt);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (cudaEvent_t e : ev_join)
      if (e) cudaEventDestroy(e);
  }
  // side stream to use for this call (the main stream itself when overlap is off or GEMMs are being timed)
  cudaStream_t side_for(cudaStream_t st) const { return (overlap && side && !gemm_timer_on()) ? side : st; }

  ConvL& enc_conv(int l, int j) { return convs[2 * l + j]; }
  ConvL& up_conv(int i, int j) { return convs[2 * (depth + 1) + 3 * i + j]; }
  ConvL& head() { return convs.back(); }
  BnL& enc_bn(int l) { return bns[l]; }
  BnL& up_bn(int i, int j) { return bns[depth + 1 + 2 * i + j]; }
};

// ---- construction ---------------------------------------------------------------------------------
static int build_tables(UNet& u) {
  const MpuUNetConfig& c = u.cfg;
  u.depth = c.depth;
  if (c.depth < 1 || c.depth > 6) {
    set_error("unet: depth %d not supported", c.depth);
    return MPU_ERR_ARG;
  }
  if (c.H % (1 << c.depth) || c.W % (1 << c.depth)) {
    set_error("unet: H=%d W=%d must be divisible by 2^depth", c.H, c.W);
    return MPU_ERR_ARG;
  }
  u.cin_phys = round_up(c.n_channels, 8);
  long long off = 0, soff = 0;
  auto add_conv = [&](const std::string& name, int ks, int cin, int cout, int k_phys, int co_phys,
                      int c0_phys) {
    ConvL L;
    L.name = name;
    L.ksize = ks;
    L.cin = cin;
    L.cout = cout;
    L.k_phys = k_phys;
    L.co_phys = co_phys;
    L.c0_phys = c0_phys;
    L.ntap_master = ks * ks;
    L.ntap_gemm = ks == 2 ? 9 : ks * ks;
    L.w_off = off;
    off += (long long)L.ntap_master * co_phys * k_phys;
    L.b_off = off;
    off += co_phys;
    u.convs.push_back(L);
  };
  auto add_bn = [&](const std::string& name, int ch) {
    BnL B;
    B.name = name;
    B.c = ch;
    B.c_phys = round_up(ch, 8);
    B.g_off = off;
    off += B.c_phys;
    B.b_off = off;
    off += B.c_phys;
    B.m_off = soff;
    soff += B.c_phys;
    B.v_off = soff;
    soff += B.c_phys;
    u.bns.push_back(B);
  };
  char nm[64];
  int cin = c.n_channels, cin_p = u.cin_phys;
  for (int l = 0; l <= c.depth; ++l) {
    const int f = c.filters[l], fp = round_up(f, 8);
    const char* pre = l < c.depth ? "encoder_L%d" : "bottom";
    char base[48];
    snprintf(base, sizeof(base), pre, l);
    snprintf(nm, sizeof(nm), "%s_conv1", base);
    add_conv(nm, 3, cin, f, cin_p, fp, cin_p);
    snprintf(nm, sizeof(nm), "%s_conv2", base);
    add_conv(nm, 3, f, f, fp, fp, fp);
    cin = f;
    cin_p = fp;
  }
  for (int l = 0; l <= c.depth; ++l) {
    if (l < c.depth) snprintf(nm, sizeof(nm), "encoder_L%d_BN", l);
    else snprintf(nm, sizeof(nm), "bottom_BN");
    add_bn(nm, c.filters[l]);
  }
  for (int i = 0; i < c.depth; ++i) {
    const int l = c.depth - 1 - i;
    const int f = c.filters[l], fp = round_up(f, 8);
    snprintf(nm, sizeof(nm), "upsample_L%d_conv1", i);
    add_conv(nm, 2, cin, f, cin_p, fp, cin_p);
    snprintf(nm, sizeof(nm), "upsample_L%d_conv2", i);
    add_conv(nm, 3, 2 * f, f, 2 * fp, fp, fp);
    snprintf(nm, sizeof(nm), "upsample_L%d_conv3", i);
    add_conv(nm, 3, f, f, fp, fp, fp);
    snprintf(nm, sizeof(nm), "upsample_L%d_BN1", i);
    add_bn(nm, f);
    snprintf(nm, sizeof(nm), "upsample_L%d_BN2", i);
    add_bn(nm, f);
    cin = f;
    cin_p = fp;
  }
  add_conv("conv2d", 1, cin, c.n_classes, cin_p, c.n_classes, cin_p);
  u.n_params = off;
  u.n_bn_state = soff;
  return MPU_OK;
}

template <typename T>
static T* bump(UNet& u, long long count, bool dry) {
  long long bytes = (count * (long long)sizeof(T) + 1023) / 1024 * 1024;
  T* p = dry ? nullptr : reinterpret_cast<T*>(u.ws + u.ws_used);
  u.ws_used += bytes;
  return p;
}

static void layout_workspace(UNet& u, bool dry) {
  const MpuUNetConfig& c = u.cfg;
  u.ws_used = 0;
  const int B = c.max_batch;
  u.lv.assign(c.depth + 1, Level());
  for (int l = 0; l <= c.depth; ++l) {
    Level& L = u.lv[l];
    L.g = Geo{B, c.H >> l, c.W >> l};
    L.C = round_up(c.filters[l], 8);
  }
  u.x_in = bump<bf16>(u, u.lv[0].g.rows() * u.cin_phys, dry);
  const bool tr = c.training != 0;
  for (int l = 0; l <= c.depth; ++l) {
    Level& L = u.lv[l];
    const long long n = L.g.rows() * L.C;
    L.a1 = bump<bf16>(u, n, dry);
    L.a2 = bump<bf16>(u, n, dry);
    L.b = bump<bf16>(u, n, dry);
    if (l < c.depth) {
      L.pooled = bump<bf16>(u, u.lv[l + 1].g.rows() * L.C, dry);
      L.u = bump<bf16>(u, n, dry);
      L.bn1 = bump<bf16>(u, n, dry);
      L.c2 = bump<bf16>(u, n, dry);
      L.c3 = bump<bf16>(u, n, dry);
      L.bn2 = bump<bf16>(u, n, dry);
    }
    if (tr) {
      L.gout = bump<bf16>(u, n, dry);
      L.s1 = bump<bf16>(u, n, dry);
      L.s2 = bump<bf16>(u, n, dry);
      if (l < c.depth) {
        L.dcat = bump<bf16>(u, 2 * n, dry);
        L.dzu = bump<bf16>(u, 4 * u.lv[l + 1].g.rows() * L.C, dry);
        L.dpool = bump<bf16>(u, u.lv[l + 1].g.rows() * L.C, dry);
      }
    }
  }
  long long dwc = 0;
  u.shadow = bump<bf16>(u, u.n_params, dry);
  for (ConvL& L : u.convs) {
    const long long n = (long long)L.ntap_gemm * L.co_phys * L.k_phys;
    if (L.ksize == 1) continue;  // head runs on CUDA cores from the fp32 master
    if (L.ksize == 3) {
      L.wf = dry ? nullptr : u.shadow + L.w_off;  // same [tap][co][k] indexing as the master
    } else {
      L.wf = bump<bf16>(u, n, dry);               // collapsed 9-pair weights of the upsample-conv
      if (n > dwc) dwc = n;
    }
    L.wd = u.dgrad_mn ? nullptr : bump<bf16>(u, n, dry);
  }
  u.dwc_floats = dwc;
  u.dwc = tr ? bump<float>(u, dwc, dry) : nullptr;
  for (BnL& b : u.bns) {
    b.scale = bump<float>(u, b.c_phys, dry);
    b.shift = bump<float>(u, b.c_phys, dry);
    b.mean = bump<float>(u, b.c_phys, dry);
    b.rstd = bump<float>(u, b.c_phys, dry);
    b.sums = bump<double>(u, 2 * b.c_phys, dry);
  }
}

// Column reductions a forward-type GEMM takes over its stored outputs in the epilogue (mtgemm.cuh FwdParams):
// the bias gradient of the conv whose dz this GEMM writes, and / or BatchNorm backward's [sum g | sum g*y].
struct EpiRed {
  float* csum_f = nullptr;
  double* red_d = nullptr;
  const bf16* red_y = nullptr;
  int red_ldy = 0, red_col0 = 0, red_C = 0;
};
static void apply_red(FwdDesc& d, const EpiRed* r) {
  if (!r) return;
  d.csum_f = r->csum_f;
  d.red_d = r->red_d;
  d.red_y = r->red_y;
  d.red_ldy = r->red_ldy;
  d.red_col0 = r->red_col0;
  d.red_C = r->red_C;
}

// ---- GEMM wrappers ----------------------------------------------------------------------------------
static void taps3x3(int Wp, int* off) {
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) off[ky * 3 + kx] = (ky - 1) * Wp + (kx - 1);
}

// out[rows][n_phys] = act( sum_taps A[m+off] . W[tap]^T + bias ), same-resolution (3x3 or dgrad).
static int gemm_same(const bf16* A0, int C0, int ldA0, const bf16* A1, int C1, int ldA1, const bf16* W,
                     int ntap, int n_phys, int k_total, Geo g, bf16* out, int ldo, const float* bias,
                     const bf16* mask, int ldm, int relu, cudaStream_t st, double* stats = nullptr,
                     int w_mn = 0, int w_rows = 0, bool flip = false, const EpiRed* red = nullptr) {
  int off[kMaxTaps] = {0}, widx[kMaxTaps];
  if (ntap == 9) taps3x3(g.Wp(), off);
  for (int t = 0; t < ntap; ++t) widx[t] = flip ? ntap - 1 - t : t;
  FwdDesc d;
  memset(&d, 0, sizeof(d));
  d.A0 = A0; d.rowsA0 = g.rows(); d.C0 = C0; d.ldA0 = ldA0;
  d.A1 = A1; d.rowsA1 = A1 ? g.rows() : 0; d.C1 = C1; d.ldA1 = ldA1;
  d.W = W; d.w_taps = ntap; d.n_phys = n_phys; d.k_total = k_total;
  d.ntaps = ntap; d.tap_a_off = off; d.tap_w = widx;
  d.M_rows = (int)g.rows();
  d.map = RowMap{g.Hp(), g.Wp(), g.Hp(), g.Wp(), 1, 0, 0};
  d.out = out; d.ldo = ldo; d.bias = bias; d.mask = mask; d.ldm = ldm; d.relu = relu;
  d.stats = stats;
  d.w_mn = w_mn; d.w_rows = w_rows;
  apply_red(d, red);
  FwdParams p;
  MPU_TRY(fwd_setup(p, d));
  return launch_fwd(p, st);
}

// dX = sum_taps dZ[m - off] . W[tap]  for a 3x3 conv layer L (K = L.co_phys, outputs = L.k_phys channels):
// either from the forward weights read MN-major with flipped tap index, or from the transposed copy
static int gemm_dgrad3x3(UNet& u, const bf16* dz, const ConvL& L, Geo g, bf16* out, const bf16* mask, int ldm,
                         cudaStream_t st, const EpiRed* red = nullptr);

// nearest-2x upsample + 2x2 SAME conv + bias + ReLU as four phase GEMMs on the low-res grid.
// side stream starts after everything enqueued on st so far
static int fork_side(UNet& u, cudaStream_t st, cudaStream_t sb) {
  if (sb == st) return MPU_OK;
  MPU_CUDA(cudaEventRecord(u.ev_fork, st));
  MPU_CUDA(cudaStreamWaitEvent(sb, u.ev_fork, 0));
  return MPU_OK;
}
// mark the side stream's position in ring slot `slot`
static int mark_side(UNet& u, cudaStream_t st, cudaStream_t sb, int slot) {
  if (sb == st) return MPU_OK;
  MPU_CUDA(cudaEventRecord(u.ev_join[slot], sb));
  return MPU_OK;
}
// st waits for the side-stream position marked in `slot`
static int wait_side(UNet& u, cudaStream_t st, cudaStream_t sb, int slot) {
  if (sb == st) return MPU_OK;
  MPU_CUDA(cudaStreamWaitEvent(st, u.ev_join[slot], 0));
  return MPU_OK;
}

static int gemm_upconv(UNet& u, const bf16* X, int Cx, Geo glo, const ConvL& L, const float* bias, Geo ghi,
                       bf16* out, cudaStream_t st, double* stats = nullptr) {
  // the four output phases are independent launches: phases (0,1) and (1,0) go to the side stream
  cudaStream_t sb = u.side_for(st);
  MPU_TRY(fork_side(u, st, sb));
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      int off[kMaxTaps], widx[kMaxTaps], n = 0;
      for (int i = 0; i < 9; ++i) {
        const UpPair pr = up_pair(i);
        if (pr.a == a && pr.b == b) {
          off[n] = pr.di * glo.Wp() + pr.dj;
          widx[n] = i;
          ++n;
        }
      }
      FwdDesc d;
      memset(&d, 0, sizeof(d));
      d.A0 = X; d.rowsA0 = glo.rows(); d.C0 = Cx; d.ldA0 = Cx;
      d.W = L.wf; d.w_taps = 9; d.n_phys = L.co_phys; d.k_total = L.k_phys;
      d.ntaps = n; d.tap_a_off = off; d.tap_w = widx;
      d.M_rows = (int)glo.rows();
      d.map = RowMap{glo.Hp(), glo.Wp(), ghi.Hp(), ghi.Wp(), 2, a, b};
      d.out = out; d.ldo = L.co_phys; d.bias = bias; d.relu = 1;
      d.stats = stats;
      FwdParams p;
      MPU_TRY(fwd_setup(p, d));
      MPU_TRY(launch_fwd(p, a != b ? sb : st));
    }
  MPU_TRY(mark_side(u, st, sb, 0));
  return wait_side(u, st, sb, 0);
}

// dIn[m_lo] = sum over the 9 (phase, tap) pairs of dZ[phase][m_lo - off] . Wc[pair]  (dZ phase-major)
static int gemm_upconv_dgrad(UNet& u, const bf16* dzu, const ConvL& L, Geo glo, bf16* out, cudaStream_t st,
                             const EpiRed* red = nullptr) {
  const long long rows_lo = glo.rows();
  int off[kMaxTaps], widx[kMaxTaps];
  for (int i = 0; i < 9; ++i) {
    const UpPair pr = up_pair(i);
    off[i] = (int)((pr.a * 2 + pr.b) * rows_lo) - (pr.di * glo.Wp() + pr.dj);
    widx[i] = i;
  }
  FwdDesc d;
  memset(&d, 0, sizeof(d));
  d.A0 = dzu; d.rowsA0 = 4 * rows_lo; d.C0 = L.co_phys; d.ldA0 = L.co_phys;
  d.w_taps = 9; d.n_phys = L.k_phys; d.k_total = L.co_phys;
  if (u.dgrad_mn) {
    d.W = L.wf; d.w_mn = 1; d.w_rows = L.co_phys;
  } else {
    d.W = L.wd;
  }
  d.ntaps = 9; d.tap_a_off = off; d.tap_w = widx;
  d.M_rows = (int)rows_lo;
  d.map = RowMap{glo.Hp(), glo.Wp(), glo.Hp(), glo.Wp(), 1, 0, 0};
  d.out = out; d.ldo = L.k_phys;
  apply_red(d, red);
  FwdParams p;
  MPU_TRY(fwd_setup(p, d));
  return launch_fwd(p, st);
}

static int gemm_dgrad3x3(UNet& u, const bf16* dz, const ConvL& L, Geo g, bf16* out, const bf16* mask, int ldm,
                         cudaStream_t st, const EpiRed* red) {
  if (u.dgrad_mn)
    return gemm_same(dz, L.co_phys, L.co_phys, nullptr, 0, 0, L.wf, 9, L.k_phys, L.co_phys, g, out, L.k_phys,
                     nullptr, mask, ldm, 0, st, nullptr, 1, L.co_phys, true, red);
  return gemm_same(dz, L.co_phys, L.co_phys, nullptr, 0, 0, L.wd, 9, L.k_phys, L.co_phys, g, out, L.k_phys,
                   nullptr, mask, ldm, 0, st, nullptr, 0, 0, false, red);
}

// dW[tap][co][col0 + ci] += sum_m X[m+off_tap][ci] * dZ[m][co]   (3x3 / 1-tap, same resolution)
// bias_grad (optional): += column sums of dZ, taken from the dZ tiles inside the same kernel
static int wgrad_same(const bf16* X, int Cx, int ldX, const bf16* dZ, int Cz, int ntap, Geo g,
                      float* dW, int ldw, int co_phys, int col0, cudaStream_t st, float* bias_grad = nullptr) {
  int off[kMaxTaps] = {0}, widx[kMaxTaps];
  if (ntap == 9) taps3x3(g.Wp(), off);
  for (int t = 0; t < ntap; ++t) widx[t] = t;
  WgradDesc d;
  memset(&d, 0, sizeof(d));
  d.X = X; d.rowsX = g.rows(); d.Cx = Cx; d.ldX = ldX;
  d.dY = dZ; d.rowsDY = g.rows(); d.Cy = Cz; d.ldDY = Cz;
  d.ntaps = ntap; d.tap_x_off = off; d.tap_dy_off = nullptr; d.tap_w = widx;
  d.rows_total = g.rows();
  d.dW = dW; d.ldw = ldw; d.w_rows_per_tap = co_phys; d.dw_col0 = col0;
  d.bias_grad = bias_grad;
  WgradParams p;
  MPU_TRY(wgrad_setup(p, d));
  return launch_wgrad(p, st);
}

// collapsed upsample-conv weight gradient: dWc[pair][co][ci] = sum_m X[m + off][ci] * dZ[phase][m][co]
static int wgrad_upconv(const bf16* X, int Cx, const bf16* dzu, const ConvL& L, Geo glo, float* dwc,
                        cudaStream_t st) {
  const long long rows_lo = glo.rows();
  int xoff[kMaxTaps], dyoff[kMaxTaps], widx[kMaxTaps];
  for (int t = 0; t < 9; ++t) {
    const UpPair pr = up_pair(t);
    xoff[t] = pr.di * glo.Wp() + pr.dj;
    dyoff[t] = (int)((pr.a * 2 + pr.b) * rows_lo);
    widx[t] = t;
  }
  WgradDesc d;
  memset(&d, 0, sizeof(d));
  d.X = X; d.rowsX = rows_lo; d.Cx = Cx; d.ldX = Cx;
  d.dY = dzu; d.rowsDY = 4 * rows_lo; d.Cy = L.co_phys; d.ldDY = L.co_phys;
  d.ntaps = 9; d.tap_x_off = xoff; d.tap_dy_off = dyoff; d.tap_w = widx;
  d.rows_total = rows_lo;
  d.dW = dwc; d.ldw = L.k_phys; d.w_rows_per_tap = L.co_phys; d.dw_col0 = 0;
  WgradParams p;
  MPU_TRY(wgrad_setup(p, d));
  return launch_wgrad(p, st);
}

// ---- schedule -------------------------------------------------------------------------------------
static Geo geo_b(const UNet& u, int l, int B) {
  Geo g = u.lv[l].g;
  g.B = B;
  return g;
}

// `stats_fused`: the producing GEMM's epilogue already accumulated sum / sum-of-squares into bn.sums
static int bn_forward(UNet& u, BnL& bn, const bf16* y, Geo g, bf16* b, bf16* pooled, int training,
                      cudaStream_t st, bool stats_fused = false) {
  const float* P = u.params;
  if (training && !stats_fused)
    MPU_TRY(launch_channel_stats(y, g.rows(), bn.c_phys, bn.c_phys, bn.sums, st));
  BnFin f;
  f.sums = bn.sums; f.count = (double)g.pixels();
  f.gamma = P + bn.g_off; f.beta = P + bn.b_off;
  f.mmean = u.bn_state + bn.m_off; f.mvar = u.bn_state + bn.v_off;
  f.eps = u.cfg.bn_eps; f.momentum = u.cfg.bn_momentum;
  f.training = training; f.C = bn.c_phys;
  f.scale = bn.scale; f.shift = bn.shift; f.mean_out = bn.mean; f.rstd_out = bn.rstd;
  return launch_bn_apply(y, f, b, pooled, g, bn.c_phys, st);
}

// bf16 operand copies that are NOT plain views of the shadow buffer: collapsed upsample-conv weights,
// and (only without MN-major dgrad) the transposed 3x3 copies
static int derive_weights(UNet& u, cudaStream_t st) {
  for (ConvL& L : u.convs) {
    if (L.ksize == 2)
      MPU_TRY(launch_prep_upconv(u.params + L.w_off, L.wf, L.wd, L.co_phys, L.k_phys, st));
    else if (L.ksize == 3 && L.wd)
      MPU_TRY(launch_prep_conv(u.params + L.w_off, nullptr, L.wd, 9, L.co_phys, L.k_phys, 1, st));
  }
  return MPU_OK;
}

static int sync_weights(UNet& u, cudaStream_t st) {
  MPU_TRY(launch_cast_bf16(u.params, u.shadow, u.n_params, st));
  MPU_TRY(derive_weights(u, st));
  u.weights_synced = true;
  return MPU_OK;
}

static int forward(UNet& u, int B, int training, cudaStream_t st) {
  if (!u.weights_synced) MPU_TRY(sync_weights(u, st));
  const int d = u.depth;
  const float* P = u.params;
  const bf16* x = u.x_in;
  int cx = u.cin_phys;
  for (int l = 0; l <= d; ++l) {
    Level& L = u.lv[l];
    const Geo g = geo_b(u, l, B);
    ConvL& c1 = u.enc_conv(l, 0);
    ConvL& c2 = u.enc_conv(l, 1);
    if (l == 0 && u.cin_phys == 8 && u.cfg.n_channels <= 4 && 9 * u.cfg.n_channels * c1.co_phys * 4 + c1.co_phys * 4 <= 48 * 1024) {
      // K = 9 * n_channels is too thin for the tensor cores: CUDA-core kernel, HBM-write bound
      MPU_TRY(launch_conv_first(x, c1.wf, P + c1.b_off, L.a1, g, u.cfg.n_channels, c1.co_phys, st));
    } else {
      MPU_TRY(gemm_same(x, cx, cx, nullptr, 0, 0, c1.wf, 9, c1.co_phys, c1.k_phys, g, L.a1, L.C,
                        P + c1.b_off, nullptr, 0, 1, st));
    }
    double* st2 = training ? u.enc_bn(l).sums : nullptr;
    if (st2) MPU_CUDA(cudaMemsetAsync(st2, 0, sizeof(double) * 2 * L.C, st));
    MPU_TRY(gemm_same(L.a1, L.C, L.C, nullptr, 0, 0, c2.wf, 9, c2.co_phys, c2.k_phys, g, L.a2, L.C,
                      P + c2.b_off, nullptr, 0, 1, st, st2));
    MPU_TRY(bn_forward(u, u.enc_bn(l), L.a2, g, L.b, l < d ? L.pooled : nullptr, training, st, st2 != nullptr));
    x = l < d ? L.pooled : L.b;
    cx = L.C;
  }
  for (int i = 0; i < d; ++i) {
    const int l = d - 1 - i;
    Level& L = u.lv[l];
    const Geo g = geo_b(u, l, B), glo = geo_b(u, l + 1, B);
    ConvL& c1 = u.up_conv(i, 0);
    ConvL& c2 = u.up_conv(i, 1);
    ConvL& c3 = u.up_conv(i, 2);
    double* st1 = training ? u.up_bn(i, 0).sums : nullptr;
    double* st3 = training ? u.up_bn(i, 1).sums : nullptr;
    if (st1) MPU_CUDA(cudaMemsetAsync(st1, 0, sizeof(double) * 2 * L.C, st));
    if (st3) MPU_CUDA(cudaMemsetAsync(st3, 0, sizeof(double) * 2 * L.C, st));
    MPU_TRY(gemm_upconv(u, x, cx, glo, c1, P + c1.b_off, g, L.u, st, st1));
    MPU_TRY(bn_forward(u, u.up_bn(i, 0), L.u, g, L.bn1, nullptr, training, st, st1 != nullptr));
    MPU_TRY(gemm_same(L.b, L.C, L.C, L.bn1, L.C, L.C, c2.wf, 9, c2.co_phys, c2.k_phys, g, L.c2, L.C,
                      P + c2.b_off, nullptr, 0, 1, st));
    MPU_TRY(gemm_same(L.c2, L.C, L.C, nullptr, 0, 0, c3.wf, 9, c3.co_phys, c3.k_phys, g, L.c3, L.C,
                      P + c3.b_off, nullptr, 0, 1, st, st3));
    MPU_TRY(bn_forward(u, u.up_bn(i, 1), L.c3, g, L.bn2, nullptr, training, st, st3 != nullptr));
    x = L.bn2;
    cx = L.C;
  }
  return MPU_OK;
}

// sums_ready: [sum g | sum g*y] were already accumulated into bn.sums by the epilogue of the GEMM that produced gA
static int bn_backward(UNet& u, BnL& bn, const bf16* y, const bf16* gA, int ldA, const bf16* gP, Geo g,
                       bf16* dz, int phase_major, float* dbias, cudaStream_t st, bool sums_ready = false) {
  BnBwdArgs a;
  a.y = y;
  a.gA = gA;
  a.ldA = ldA;
  a.gP = gP;
  a.scale = bn.scale;
  a.shift = bn.shift;
  a.mean = bn.mean;
  a.rstd = bn.rstd;
  a.gamma = u.params + bn.g_off;
  a.g = g;
  a.C = bn.c_phys;
  a.dgamma = a.dbeta = nullptr;  // (set by launch_bn_bwd_apply)
  if (!sums_ready) MPU_TRY(launch_bn_bwd_reduce(a, bn.sums, st));
  return launch_bn_bwd_apply(a, bn.sums, dz, phase_major, u.grads + bn.g_off, u.grads + bn.b_off, dbias,
                             st);
}

// conv-ReLU-conv-ReLU-BN block backward given dz2 (gradient at the second conv's pre-activation):
// wgrad conv2, dgrad conv2 (masked by a1 -> dz1), bias1, wgrad conv1, optional dgrad conv1.
static int block_tail_backward(UNet& u, ConvL& c1, ConvL& c2, const bf16* xin0, int cx0, const bf16* xin1,
                               int cx1, const bf16* a1, const bf16* dz2, bf16* dz1, bf16* dxin, int C,
                               Geo g, cudaStream_t st, cudaStream_t sb, bool fuse_red = false,
                               const EpiRed* dxin_red = nullptr) {
  float* G = u.grads;
  // weight / bias gradients run on the side stream sb; the dgrad chain (critical path) stays on st
  MPU_TRY(fork_side(u, st, sb));  // dz2 is ready
  MPU_TRY(wgrad_same(a1, C, C, dz2, C, 9, g, G + c2.w_off, c2.k_phys, c2.co_phys, 0, sb));
  // conv1's bias gradient = column sums of dz1: taken by the epilogue of the dgrad that writes dz1, or by a
  // separate pass on the side stream
  EpiRed bias_red;
  bias_red.csum_f = G + c1.b_off;
  MPU_TRY(gemm_dgrad3x3(u, dz2, c2, g, dz1, a1, C, st, fuse_red ? &bias_red : nullptr));
  MPU_TRY(fork_side(u, st, sb));  // dz1 is ready
  // ... or (default) by the weight-gradient kernel of conv1 itself, from the dz1 tiles it streams anyway: the separate
  // column-sum pass (9 launches, 1.56 GB of DRAM reads per step) is only kept as a bring-up comparison
  // (MPU_BIAS_COLSUM=1)
  float* bg = (fuse_red || u.bias_colsum_pass) ? nullptr : G + c1.b_off;
  if (!fuse_red && u.bias_colsum_pass) MPU_TRY(launch_colsum(dz1, g.rows(), C, C, G + c1.b_off, sb));
  if (!xin1 && cx0 == 8 && c1.k_phys == 8 && u.cfg.n_channels <= 4 && xin0 == u.x_in)
    // first conv of the network: K = 9 * n_channels is too thin for the tensor cores
    MPU_TRY(launch_conv_first_wgrad(xin0, dz1, g, u.cfg.n_channels, c1.co_phys, G + c1.w_off, c1.k_phys, bg, sb));
  else
    MPU_TRY(wgrad_same(xin0, cx0, cx0, dz1, C, 9, g, G + c1.w_off, c1.k_phys, c1.co_phys, 0, sb, bg));
  if (xin1)
    MPU_TRY(wgrad_same(xin1, cx1, cx1, dz1, C, 9, g, G + c1.w_off, c1.k_phys, c1.co_phys, cx0, sb));
  if (dxin) MPU_TRY(gemm_dgrad3x3(u, dz1, c1, g, dxin, nullptr, 0, st, dxin_red));
  return MPU_OK;
}

// Buffers the side stream reads (dz tensors of a level) are rewritten by the main stream at the earliest two
// blocks later (up block of level l -> encoder block of level l), so each block first waits for the side
// work of the block before the previous one; stages end with a full join.
static int block_begin(UNet& u, cudaStream_t st, cudaStream_t sb) {
  if (u.blk >= 2) MPU_TRY(wait_side(u, st, sb, (u.blk - 2) % 3));
  return MPU_OK;
}
static int block_end(UNet& u, cudaStream_t st, cudaStream_t sb) {
  MPU_TRY(mark_side(u, st, sb, u.blk % 3));
  ++u.blk;
  return MPU_OK;
}
static int stage_join(UNet& u, cudaStream_t st, cudaStream_t sb) {
  if (u.blk > 0) MPU_TRY(wait_side(u, st, sb, (u.blk - 1) % 3));
  u.blk = 0;
  return MPU_OK;
}

// Backward in three stages so the caller can start the gradient all-reduce of finished parameter ranges
// while the rest of backward still runs:  0 = up path (+ head, already done by the loss kernel),
// 1 = bottom block, 2 = encoder levels depth-1 .. 0.
static int backward_up(UNet& u, int B, cudaStream_t st, cudaStream_t sb) {
  const int d = u.depth;
  float* G = u.grads;
  for (int l = 0; l < d; ++l) {
    MPU_TRY(block_begin(u, st, sb));
    const int i = d - 1 - l;
    Level& L = u.lv[l];
    Level& Lo = u.lv[l + 1];
    const Geo g = geo_b(u, l, B), glo = geo_b(u, l + 1, B);
    ConvL& c1 = u.up_conv(i, 0);
    ConvL& c2 = u.up_conv(i, 1);
    ConvL& c3 = u.up_conv(i, 2);
    // BN2 backward: gradient wrt bn2_l is in L.gout; for l > 0 the upsample-conv dgrad of the previous block
    // already left [sum g | sum g*y] in the BN's sums (level 0's gradient comes from the head kernel)
    const bool fuse_here = l >= u.red_min_level, fuse_lo = l + 1 >= u.red_min_level;
    MPU_TRY(bn_backward(u, u.up_bn(i, 1), L.c3, L.gout, L.C, nullptr, g, L.s1, 0, G + c3.b_off, st,
                        l > 0 && fuse_here));
    // conv3 / conv2 ([skip | bn1] concat input) backward; dgrad of conv2 -> dcat [rows][2C]; its second half is the
    // gradient at BN1's output: the dgrad epilogue accumulates BN1's reduction sums against y = u
    BnL& bn1 = u.up_bn(i, 0);
    if (fuse_here) MPU_CUDA(cudaMemsetAsync(bn1.sums, 0, sizeof(double) * 2 * L.C, st));
    EpiRed r1;
    r1.red_d = bn1.sums; r1.red_y = L.u; r1.red_ldy = L.C; r1.red_col0 = L.C; r1.red_C = L.C;
    MPU_TRY(block_tail_backward(u, c2, c3, L.b, L.C, L.bn1, L.C, L.c2, L.s1, L.s2, L.dcat, L.C, g, st, sb, fuse_here,
                                fuse_here ? &r1 : nullptr));
    // BN1 backward on the second half of dcat -> dz of the upsample-conv, phase-major
    MPU_TRY(bn_backward(u, bn1, L.u, L.dcat + L.C, 2 * L.C, nullptr, g, L.dzu, 1, G + c1.b_off, st, fuse_here));
    // upsample-conv backward
    const bf16* xin = (l + 1 == d) ? Lo.b : Lo.bn2;
    MPU_TRY(fork_side(u, st, sb));  // dzu is ready
    MPU_CUDA(cudaMemsetAsync(u.dwc, 0, sizeof(float) * 9 * c1.co_phys * c1.k_phys, sb));
    MPU_TRY(wgrad_upconv(xin, Lo.C, L.dzu, c1, glo, u.dwc, sb));
    MPU_TRY(launch_fold_upconv_grad(u.dwc, G + c1.w_off, c1.co_phys, c1.k_phys, sb));
    // the gradient it writes sits at a BN output (BN2 of the next coarser up block, or the bottom BN): take that
    // BN's reduction sums in the epilogue
    BnL& bnn = (l + 1 == d) ? u.enc_bn(d) : u.up_bn(i - 1, 1);
    if (fuse_lo) MPU_CUDA(cudaMemsetAsync(bnn.sums, 0, sizeof(double) * 2 * Lo.C, st));
    EpiRed r2;
    r2.red_d = bnn.sums; r2.red_y = (l + 1 == d) ? Lo.a2 : Lo.c3; r2.red_ldy = Lo.C; r2.red_col0 = 0; r2.red_C = Lo.C;
    MPU_TRY(gemm_upconv_dgrad(u, L.dzu, c1, glo, Lo.gout, st, fuse_lo ? &r2 : nullptr));
    MPU_TRY(block_end(u, st, sb));
  }
  return MPU_OK;
}

static int backward_enc_level(UNet& u, int B, int l, cudaStream_t st, cudaStream_t sb) {
  const int d = u.depth;
  float* G = u.grads;
  MPU_TRY(block_begin(u, st, sb));
  Level& L = u.lv[l];
  const Geo g = geo_b(u, l, B);
  ConvL& c1 = u.enc_conv(l, 0);
  ConvL& c2 = u.enc_conv(l, 1);
  if (l == d) {  // (sums left by the upsample-conv dgrad of the last up block, backward stage 0)
    MPU_TRY(bn_backward(u, u.enc_bn(l), L.a2, L.gout, L.C, nullptr, g, L.s1, 0, G + c2.b_off, st,
                        l >= u.red_min_level));
  } else {
    // skip gradient = first half of dcat_l; pooled gradient = dpool_l (from level l+1's conv1 dgrad)
    MPU_TRY(bn_backward(u, u.enc_bn(l), L.a2, L.dcat, 2 * L.C, L.dpool, g, L.s1, 0, G + c2.b_off, st));
  }
  const bf16* xin = l == 0 ? u.x_in : u.lv[l - 1].pooled;
  const int cx = l == 0 ? u.cin_phys : u.lv[l - 1].C;
  bf16* dxin = l == 0 ? nullptr : u.lv[l - 1].dpool;
  MPU_TRY(block_tail_backward(u, c1, c2, xin, cx, nullptr, 0, L.a1, L.s1, L.s2, dxin, L.C, g, st, sb,
                              l >= u.red_min_level));
  return block_end(u, st, sb);
}

static int backward_stage(UNet& u, int B, int stage, cudaStream_t st) {
  cudaStream_t sb = u.side_for(st);
  u.blk = 0;
  if (stage == 0) {
    MPU_TRY(backward_up(u, B, st, sb));
  } else if (stage == 1) {
    MPU_TRY(backward_enc_level(u, B, u.depth, st, sb));
  } else if (stage == 2) {
    for (int l = u.depth - 1; l >= 0; --l) MPU_TRY(backward_enc_level(u, B, l, st, sb));
  } else {
    set_error("backward_stage: stage %d out of range", stage);
    return MPU_ERR_ARG;
  }
  return stage_join(u, st, sb);  // all gradients of the stage are complete on st
}

static int backward(UNet& u, int B, cudaStream_t st) {
  for (int s = 0; s < 3; ++s) MPU_TRY(backward_stage(u, B, s, st));
  return MPU_OK;
}

}  // namespace mpu

// ======================================================================================================
using namespace mpu;

extern "C" {

static int check_cfg(const MpuUNetConfig* cfg) {
  if (!cfg) {
    set_error("unet: null config");
    return MPU_ERR_ARG;
  }
  if (cfg->max_batch < 1 || cfg->n_classes < 1 || cfg->n_classes > 16 || cfg->n_channels < 1) {
    set_error("unet: bad config (max_batch=%d n_classes=%d n_channels=%d)", cfg->max_batch,
              cfg->n_classes, cfg->n_channels);
    return MPU_ERR_ARG;
  }
  return MPU_OK;
}

int mpu_unet_sizes(const MpuUNetConfig* cfg, long long* n_params, long long* n_bn_state,
                   long long* workspace_bytes) {
  MPU_TRY(check_cfg(cfg));
  UNet u;
  u.cfg = *cfg;
  if (const char* e = getenv("MPU_DGRAD_MN")) u.dgrad_mn = atoi(e) != 0;
  MPU_TRY(build_tables(u));
  layout_workspace(u, true);
  if (n_params) *n_params = u.n_params;
  if (n_bn_state) *n_bn_state = u.n_bn_state;
  if (workspace_bytes) *workspace_bytes = u.ws_used;
  return MPU_OK;
}

int mpu_unet_create(const MpuUNetConfig* cfg, float* params, float* grads, float* adam_m, float* adam_v,
                    float* bn_state, void* workspace, long long workspace_bytes, void* stream,
                    void** handle) {
  MPU_TRY(check_cfg(cfg));
  if (!params || !bn_state || !workspace || !handle) {
    set_error("unet_create: null buffer");
    return MPU_ERR_ARG;
  }
  if (cfg->training && (!grads || !adam_m || !adam_v)) {
    set_error("unet_create: training handle needs grads / adam buffers");
    return MPU_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("unet_create: no CUDA device (this library has no CPU fallback)");
    return MPU_ERR_CUDA;
  }
  UNet* u = new UNet();
  u->cfg = *cfg;
  if (const char* e = getenv("MPU_DGRAD_MN")) u->dgrad_mn = atoi(e) != 0;
  if (const char* e = getenv("MPU_EPI_RED_LEVEL")) u->red_min_level = atoi(e);
  if (const char* e = getenv("MPU_BIAS_COLSUM")) u->bias_colsum_pass = atoi(e);
  if (const char* e = getenv("MPU_OVERLAP")) u->overlap = atoi(e) != 0;
  if (u->overlap) {
    bool ok = cudaStreamCreateWithFlags(&u->side, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&u->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (cudaEvent_t& ev : u->ev_join) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&u->opt, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&u->ev_stage, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&u->ev_opt, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      set_error("unet_create: could not create the side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
      delete u;
      return MPU_ERR_CUDA;
    }
  }
  int rc = build_tables(*u);
  if (rc != MPU_OK) {
    delete u;
    return rc;
  }
  layout_workspace(*u, true);
  if (u->ws_used > workspace_bytes) {
    set_error("unet_create: workspace too small (%lld < %lld bytes)", workspace_bytes, u->ws_used);
    delete u;
    return MPU_ERR_NOMEM;
  }
  u->params = params;
  u->grads = grads;
  u->adam_m = adam_m;
  u->adam_v = adam_v;
  u->bn_state = bn_state;
  u->ws = reinterpret_cast<char*>(workspace);
  u->ws_bytes = workspace_bytes;
  const long long need = u->ws_used;
  layout_workspace(*u, false);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // zero borders / padded channels once; kernels only ever write interiors
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)need, st);
  if (e != cudaSuccess) {
    set_error("unet_create: memset failed: %s", cudaGetErrorString(e));
    delete u;
    return MPU_ERR_CUDA;
  }
  *handle = u;
  return MPU_OK;
}

int mpu_unet_destroy(void* handle) {
  delete reinterpret_cast<UNet*>(handle);
  return MPU_OK;
}

int mpu_unet_num_layers(void* handle) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  return (int)(u->convs.size() + u->bns.size());
}

int mpu_unet_layer_info(void* handle, int idx, MpuLayerInfo* out) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (!u || !out || idx < 0 || idx >= (int)(u->convs.size() + u->bns.size())) {
    set_error("layer_info: bad index %d", idx);
    return MPU_ERR_ARG;
  }
  memset(out, 0, sizeof(*out));
  if (idx < (int)u->convs.size()) {
    const ConvL& L = u->convs[idx];
    snprintf(out->name, sizeof(out->name), "%s", L.name.c_str());
    out->kind = 0;
    out->ksize = L.ksize;
    out->cin = L.cin;
    out->cout = L.cout;
    out->k_phys = L.k_phys;
    out->co_phys = L.co_phys;
    out->c0_phys = L.c0_phys;
    out->off0 = L.w_off;
    out->off1 = L.b_off;
  } else {
    const BnL& b = u->bns[idx - u->convs.size()];
    snprintf(out->name, sizeof(out->name), "%s", b.name.c_str());
    out->kind = 1;
    out->cout = b.c;
    out->co_phys = b.c_phys;
    out->off0 = b.g_off;
    out->off1 = b.b_off;
    out->off2 = b.m_off;
    out->off3 = b.v_off;
  }
  return MPU_OK;
}

int mpu_unet_sync_weights(void* handle, void* stream) {
  return sync_weights(*reinterpret_cast<UNet*>(handle), reinterpret_cast<cudaStream_t>(stream));
}

int mpu_unet_input_buffer(void* handle, void** ptr, int* cin_phys, long long* rows) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (ptr) *ptr = u->x_in;
  if (cin_phys) *cin_phys = u->cin_phys;
  if (rows) *rows = u->lv[0].g.rows();
  return MPU_OK;
}

int mpu_unet_pack_input(void* handle, const float* x_nhwc, int B, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (B < 1 || B > u->cfg.max_batch) {
    set_error("pack_input: batch %d outside [1,%d]", B, u->cfg.max_batch);
    return MPU_ERR_ARG;
  }
  return launch_pack_input(x_nhwc, geo_b(*u, 0, B), u->cfg.n_channels, u->cin_phys, u->x_in,
                           reinterpret_cast<cudaStream_t>(stream));
}

int mpu_unet_forward(void* handle, int B, int bn_training, float* probs_out, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B < 1 || B > u->cfg.max_batch) {
    set_error("forward: batch %d outside [1,%d]", B, u->cfg.max_batch);
    return MPU_ERR_ARG;
  }
  MPU_TRY(forward(*u, B, bn_training, st));
  const ConvL& H = u->head();
  return launch_head_infer(u->lv[0].bn2, geo_b(*u, 0, B), u->lv[0].C, u->params + H.w_off,
                           u->params + H.b_off, u->cfg.n_classes, probs_out, st);
}

static int train_forward(UNet* u, int B, const unsigned char* labels, const float* sample_w,
                         float grad_scale, double* loss_sum, float* probs_opt, cudaStream_t st) {
  if (!u->cfg.training) {
    set_error("train_step: handle was created with training=0");
    return MPU_ERR_STATE;
  }
  if (B < 1 || B > u->cfg.max_batch) {
    set_error("train_step: batch %d outside [1,%d]", B, u->cfg.max_batch);
    return MPU_ERR_ARG;
  }
  MPU_CUDA(cudaMemsetAsync(u->grads, 0, sizeof(float) * u->n_params, st));
  MPU_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
  MPU_TRY(forward(*u, B, 1, st));
  const ConvL& H = u->head();
  return launch_head_train(u->lv[0].bn2, geo_b(*u, 0, B), u->lv[0].C, u->params + H.w_off,
                           u->params + H.b_off, u->cfg.n_classes, labels, sample_w, grad_scale,
                           u->lv[0].gout, u->grads + H.w_off, u->grads + H.b_off, loss_sum, probs_opt, st);
}

int mpu_unet_train_step(void* handle, int B, const unsigned char* labels, const float* sample_w,
                        float grad_scale, double* loss_sum, float* probs_opt, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MPU_TRY(train_forward(u, B, labels, sample_w, grad_scale, loss_sum, probs_opt, st));
  return backward(*u, B, st);
}

// Whole single-process train step with the optimizer overlapped: after each backward stage the parameter ranges whose
// gradients are final (mpu_unet_grad_ranges) get their l2 penalty + Adam update on a third stream while the next stage
// computes on the main one (a later stage reads neither the weights nor the gradients of an earlier range).  Results
// are identical to mpu_unet_train_step + mpu_unet_l2_penalty + mpu_unet_adam.
int mpu_unet_train_step_adam(void* handle, int B, const unsigned char* labels, const float* sample_w,
                             float grad_scale, double* loss_sum, float* probs_opt, float lr, float beta1, float beta2,
                             float eps, int step, float l2_grad_coef, double* l2_sumsq, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MPU_TRY(train_forward(u, B, labels, sample_w, grad_scale, loss_sum, probs_opt, st));
  if (l2_grad_coef != 0.f && !l2_sumsq) {
    set_error("train_step_adam: l2 penalty requested without an output for sum(w^2)");
    return MPU_ERR_ARG;
  }
  long long r[8];
  MPU_TRY(mpu_unet_grad_ranges(handle, r));
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  const bool ov = u->overlap && u->opt && !gemm_timer_on();
  cudaStream_t so = ov ? u->opt : st;
  if (l2_grad_coef != 0.f) MPU_CUDA(cudaMemsetAsync(l2_sumsq, 0, sizeof(double), st));
  for (int stage = 0; stage < 3; ++stage) {
    MPU_TRY(backward_stage(*u, B, stage, st));
    if (ov) {
      MPU_CUDA(cudaEventRecord(u->ev_stage, st));
      MPU_CUDA(cudaStreamWaitEvent(so, u->ev_stage, 0));
    }
    for (int k = (stage < 2 ? stage : 2); k <= (stage < 2 ? stage : 3); ++k) {
      const long long a = r[2 * k], b = r[2 * k + 1];
      if (b <= a) continue;
      if (l2_grad_coef != 0.f) MPU_TRY(mpu_unet_l2_penalty(handle, a, b, l2_grad_coef, l2_sumsq, so));
      MPU_TRY(launch_adam(u->params + a, u->grads + a, u->adam_m + a, u->adam_v + a, b - a, (float)lr_t, beta1,
                          beta2, eps, 1.0f, u->shadow + a, so));
    }
  }
  if (ov) {
    MPU_CUDA(cudaEventRecord(u->ev_opt, so));
    MPU_CUDA(cudaStreamWaitEvent(st, u->ev_opt, 0));
  }
  return derive_weights(*u, st);
}

int mpu_unet_train_forward(void* handle, int B, const unsigned char* labels, const float* sample_w,
                           float grad_scale, double* loss_sum, float* probs_opt, void* stream) {
  return train_forward(reinterpret_cast<UNet*>(handle), B, labels, sample_w, grad_scale, loss_sum, probs_opt,
                       reinterpret_cast<cudaStream_t>(stream));
}

int mpu_unet_backward_stage(void* handle, int B, int stage, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (!u->cfg.training || B < 1 || B > u->cfg.max_batch) {
    set_error("backward_stage: bad handle state or batch");
    return MPU_ERR_ARG;
  }
  return backward_stage(*u, B, stage, reinterpret_cast<cudaStream_t>(stream));
}

int mpu_unet_grad_ranges(void* handle, long long* out8) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (!u || !out8) {
    set_error("grad_ranges: null argument");
    return MPU_ERR_ARG;
  }
  const long long up0 = u->up_conv(0, 0).w_off, bot0 = u->enc_conv(u->depth, 0).w_off, bn0 = u->bns[0].g_off;
  out8[0] = up0;  out8[1] = u->n_params;  // complete after stage 0 (up path + head)
  out8[2] = bot0; out8[3] = bn0;          // complete after stage 1 (bottom convs)
  out8[4] = 0;    out8[5] = bot0;         // complete after stage 2 (encoder convs)
  out8[6] = bn0;  out8[7] = up0;          // complete after stage 2 (encoder + bottom BatchNorm)
  return MPU_OK;
}

int mpu_unet_adam(void* handle, float lr, float beta1, float beta2, float eps, int step,
                  float grad_scale, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!u->cfg.training) {
    set_error("adam: handle was created with training=0");
    return MPU_ERR_STATE;
  }
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  MPU_TRY(launch_adam(u->params, u->grads, u->adam_m, u->adam_v, u->n_params, (float)lr_t, beta1, beta2,
                      eps, grad_scale, u->shadow, st));
  return derive_weights(*u, st);
}

int mpu_unet_adam_range(void* handle, long long begin, long long end, float lr, float beta1, float beta2, float eps,
                        int step, float grad_scale, int finish, void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!u->cfg.training || begin < 0 || end > u->n_params || begin > end) {
    set_error("adam_range: bad handle state or range [%lld, %lld)", begin, end);
    return MPU_ERR_ARG;
  }
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  MPU_TRY(launch_adam(u->params + begin, u->grads + begin, u->adam_m + begin, u->adam_v + begin, end - begin,
                      (float)lr_t, beta1, beta2, eps, grad_scale, u->shadow + begin, st));
  return finish ? derive_weights(*u, st) : MPU_OK;
}

// Keras kernel_regularizer=l2(l2) on every conv of the encoder / bottom / up path (not the 1x1 head:
// mpunet/models/unet.py:114-196): for the conv kernels inside the parameter range [begin, end) adds
// grad_coef * w to the gradient and accumulates sum w^2 into *sumsq_out (device double, NOT zeroed here).
int mpu_unet_l2_penalty(void* handle, long long begin, long long end, float grad_coef, double* sumsq_out,
                        void* stream) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!u->cfg.training || begin < 0 || end > u->n_params || begin > end || !sumsq_out) {
    set_error("l2_penalty: bad handle state, range [%lld, %lld) or null output", begin, end);
    return MPU_ERR_ARG;
  }
  for (size_t i = 0; i + 1 < u->convs.size(); ++i) {  // the head is last and carries no regulariser
    const ConvL& L = u->convs[i];
    const long long n = (long long)L.ntap_master * L.co_phys * L.k_phys;
    const long long a = std::max(begin, L.w_off), b = std::min(end, L.w_off + n);
    if (b > a) MPU_TRY(launch_l2_penalty(u->params + a, u->grads + a, b - a, grad_coef, sumsq_out, st));
  }
  return MPU_OK;
}

// debug / test access to internal activations: which = 0:a1 1:a2 2:b 3:pooled 4:u 5:bn1 6:c2 7:c3 8:bn2
// 9:gout 10:s1 11:s2 12:dcat 13:dzu 14:dpool
int mpu_unet_debug_buffer(void* handle, int level, int which, void** ptr, long long* rows, int* C) {
  UNet* u = reinterpret_cast<UNet*>(handle);
  if (level < 0 || level > u->depth) {
    set_error("debug_buffer: bad level");
    return MPU_ERR_ARG;
  }
  Level& L = u->lv[level];
  bf16* tab[15] = {L.a1, L.a2, L.b, L.pooled, L.u, L.bn1, L.c2, L.c3, L.bn2, L.gout, L.s1, L.s2,
                   L.dcat, L.dzu, L.dpool};
  if (which < 0 || which >= 15) {
    set_error("debug_buffer: bad selector");
    return MPU_ERR_ARG;
  }
  if (ptr) *ptr = tab[which];
  if (rows) *rows = L.g.rows();
  if (C) *C = L.C;
  return MPU_OK;
}

}  // extern "C"
