// CUDA-core kernels around the tensor-core GEMMs of the 2D U-Net: BatchNorm (stats / apply / backward),
// 2x2 max-pool (fused into BN apply / backward), 1x1 softmax head (+ sparse CE fwd/bwd), Adam,
// weight re-layout.  All activations: zero-bordered NHWC bf16, C multiple of 8, 16-byte vectors.
// Reference semantics: mpunet/models/unet.py:114-216 (Keras layers), TF/Keras 2.3 defaults.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mpu {

struct Geo {       // one resolution level
  int B, H, W;     // real dims; padded dims are H+2, W+2
  __host__ __device__ int Hp() const { return H + 2; }
  __host__ __device__ int Wp() const { return W + 2; }
  __host__ __device__ long long rows() const { return (long long)B * (H + 2) * (W + 2); }
  __host__ __device__ long long pixels() const { return (long long)B * H * W; }
};

// Per-channel sum and sum of squares over all rows (borders are zero, so they do not contribute).
int launch_channel_stats(const __nv_bfloat16* y, long long rows, int C, int ld, double* sums /*[2][C]*/,
                         cudaStream_t st);

// BN coefficients.  training: from batch sums (biased var for normalisation, unbiased for the moving
// update, momentum m); inference: from moving stats.  Outputs scale/shift (fwd) and mean/rstd (bwd).
int launch_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                       float* moving_mean, float* moving_var, float eps, float momentum, int training,
                       int C, float* scale, float* shift, float* mean, float* rstd, cudaStream_t st);

// What bn_finalize computes, handed to the apply kernels so that they derive the coefficients themselves (no separate
// launch between the stats-producing GEMM and the apply pass): batch sums / count (training) or moving statistics
// (inference) -> scale / shift; block 0 also stores scale / shift / mean / rstd and updates the moving statistics.
struct BnFin {
  const double* sums;            // [2][C] sum / sum of squares (training)
  double count;
  const float *gamma, *beta;
  float *mmean, *mvar;           // moving statistics (updated in training mode)
  float eps, momentum;
  int training, C;
  float *scale, *shift, *mean_out, *rstd_out;
  int no_write;                  // set by launch_bn_apply (bring-up: coefficients finalised by a separate launch)
};
// b = y*scale + shift on interior pixels; optionally also the 2x2 max-pooled tensor (next level).
int launch_bn_apply(const __nv_bfloat16* y, const BnFin& f, __nv_bfloat16* b,
                    __nv_bfloat16* pooled, Geo g, int C, cudaStream_t st);

struct BnBwdArgs {
  const __nv_bfloat16* y;        // BN input (post-ReLU conv output), [rows][C]
  const __nv_bfloat16* gA;       // direct gradient wrt BN output (may be null), row stride ldA
  int ldA;
  const __nv_bfloat16* gP;       // gradient wrt the pooled tensor (may be null), [rows_lo][C]
  const float *scale, *shift, *mean, *rstd, *gamma;
  Geo g;
  int C;
  float *dgamma, *dbeta;         // set by launch_bn_bwd_apply: parameter gradients, added by block 0 of the apply pass
};
// pass 1: sums[0][c] = sum g, sums[1][c] = sum g*xhat
int launch_bn_bwd_reduce(const BnBwdArgs& a, double* sums, cudaStream_t st);
// pass 2: dz = relu'(y) * gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); also dgamma/dbeta and the bias
// gradient of the conv that produced y (sum dz).  phase_major: write dz as [4][rows_lo][C] (for the
// upsample-conv backward) instead of the normal layout.
int launch_bn_bwd_apply(const BnBwdArgs& a, const double* sums, __nv_bfloat16* dz, int phase_major,
                        float* dgamma, float* dbeta, float* dbias, cudaStream_t st);

// Per-channel column sum of a bf16 matrix, accumulated into fp32 out[C] (bias gradients).
int launch_colsum(const __nv_bfloat16* x, long long rows, int C, int ld, float* out, cudaStream_t st);

// First conv of the network (tiny K = 9 * n_channels): CUDA-core 3x3 conv + bias + ReLU from the
// zero-bordered bf16 input [rows][cin_phys] (cin_phys == 8) with the bf16 forward weights
// [9][co_phys][8]; writes the zero-bordered bf16 output.  The tensor-core path would spend its time on
// zero-filled TMA tiles here.
int launch_conv_first(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out,
                      Geo g, int cin, int co_phys, cudaStream_t st);

// Its weight gradient: dW[tap][co][ci] (row stride ldw floats per (tap, co)) += sum_p dz[p][co] * x[p+tap][ci].
// dbias (optional): += column sums of dz, the conv's bias gradient, taken in the same pass
int launch_conv_first_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dz, Geo g, int cin, int co_phys,
                            float* dW, int ldw, float* dbias, cudaStream_t st);

// 1x1 conv + softmax head.  x = BN2 output of the last up block.
int launch_head_infer(const __nv_bfloat16* x, Geo g, int C, const float* Wh /*[ncls][C]*/,
                      const float* bh, int ncls, float* probs /*[B,H,W,ncls]*/, cudaStream_t st);
// forward + sparse-CE + backward in one pass.  grad_scale multiplies dlogits (1 = Keras "sum").
int launch_head_train(const __nv_bfloat16* x, Geo g, int C, const float* Wh, const float* bh, int ncls,
                      const uint8_t* labels /*[B,H,W]*/, const float* sample_w /*[B] or null*/,
                      float grad_scale, __nv_bfloat16* dx, float* dWh, float* dbh,
                      double* loss_sum, float* probs_opt, cudaStream_t st);

// Adam on the flat parameter buffer; also writes the bf16 shadow copy (same indexing) the GEMMs read.
int launch_l2_penalty(const float* w, float* g, long long n, float coef, double* sumsq, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1,
                float b2, float eps, float gscale, __nv_bfloat16* shadow, cudaStream_t st);
int launch_cast_bf16(const float* p, __nv_bfloat16* shadow, long long n, cudaStream_t st);

// master fp32 [ntap][co][k] -> bf16 forward copy (same layout) and dgrad copy [ntap][k][co].
// flip: dgrad tap index = ntap-1-t (3x3 spatial flip).
int launch_prep_conv(const float* w, __nv_bfloat16* wf, __nv_bfloat16* wd, int ntap, int co, int k,
                     int flip, cudaStream_t st);
// master fp32 [4][co][k] (dy*2+dx) -> collapsed 9-pair bf16 copies, forward [9][co][k] + dgrad [9][k][co].
int launch_prep_upconv(const float* w, __nv_bfloat16* wf, __nv_bfloat16* wd, int co, int k,
                       cudaStream_t st);
// dWc [9][co][k] fp32 -> accumulate into master-layout gradient [4][co][k].
int launch_fold_upconv_grad(const float* dwc, float* dw, int co, int k, cudaStream_t st);

// The 9 (phase, tap) pairs of the collapsed upsample-conv, fixed order shared by host and device.
struct UpPair { int a, b, di, dj; };
__host__ __device__ inline UpPair up_pair(int i) {
  // (a, b, di, dj) packed 4 bits each: out[2i+a, 2j+b] += Wc[pair] . in[i+di, j+dj]
  const unsigned short tab[9] = {0x0000, 0x0100, 0x0101, 0x1000, 0x1010, 0x1100, 0x1101, 0x1110, 0x1111};
  const unsigned short t = tab[i];
  return UpPair{(t >> 12) & 1, (t >> 8) & 1, (t >> 4) & 1, t & 1};
}

// fp32 NHWC [B,H,W,cin] (unpadded, as the reference feeds Keras) -> zero-bordered bf16 [rows][cin_phys]
int launch_pack_input(const float* x, Geo g, int cin, int cin_phys, __nv_bfloat16* out, cudaStream_t st);

}  // namespace mpu
