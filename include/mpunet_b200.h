/* libmpunet_b200.so — C ABI of the B200-native mpunet hot path.
 *
 * The reference (perslev/MultiPlanarUNet, `mpunet` 0.2.12) has no native layer at all: its extension
 * points are Python objects selected by name.  This header is the boundary a
 * maintainer binds with ctypes from those Python plug points; each entry cites the reference code it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (a torch.Tensor's data_ptr()) unless the
 *     argument name starts with `h_` (host pointer);
 *   - every call takes the CUDA stream to enqueue on (`void* stream` = cudaStream_t; NULL = default);
 *   - return value 0 = OK, negative = error; mpu_last_error() gives a thread-local message;
 *   - nothing is allocated or freed behind the caller's back except inside opaque handles;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef MPUNET_B200_H_
#define MPUNET_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MPU_OK 0
#define MPU_ERR_ARG (-1)
#define MPU_ERR_CUDA (-2)
#define MPU_ERR_STATE (-3)
#define MPU_ERR_NOMEM (-4)

const char* mpu_last_error(void);
int mpu_version(void);
/* number of CUDA kernels this library has launched so far in this process */
long long mpu_launch_count(void);
/* live CUDA-event timing of the tensor-core GEMM launches (bench.py roofline): enable, run a step,
 * then read the summed device time and launch count (read also resets) */
int mpu_profile_gemm(int enable);
int mpu_profile_gemm_read(double* total_ms, int* launches);

/* ---- kernel-level surface: multi-tap GEMM on tcgen05 (bring-up + kernel parity tests) ------------
 * The contraction inside tf.keras Conv2D / its gradients, reference call sites
 * mpunet/models/unet.py:120-127,137-144,159-163,171-178 (forward) and the Keras autodiff of
 * train/trainer.py:246 (backward).  Activations are bf16 [rows][channels] matrices in zero-bordered
 * NHWC layout; W is bf16 [w_taps][n_phys][k_total]. */
int mpu_mtgemm_fwd(const void* A0, long long rowsA0, int C0, int ldA0, const void* A1,
                   long long rowsA1, int C1, int ldA1, const void* W, int w_taps, int n_phys,
                   int k_total, int ntaps, const int* h_tap_a_off, const int* h_tap_w, int M_rows,
                   int Hp, int Wp, int oHp, int oWp, int s, int py, int px, void* out, int ldo,
                   const float* bias, const void* mask, int ldm, int relu, void* stream);

/* dW[tap][co][dw_col0 + ci] += sum_m X[m + h_tap_x_off[tap]][ci] * dY[m + h_tap_dy_off[tap]][co] over rows_total
 * anchor rows m (h_tap_dy_off may be NULL = 0; taps at consecutive X rows of one dY offset share one MMA).
 * dW is fp32 [taps][w_rows_per_tap][ldw], accumulated with vector reductions (zero it first); splits = 0 chooses
 * the K split. */
int mpu_mtgemm_wgrad(const void* X, long long rowsX, int Cx, int ldX, const void* dY,
                     long long rowsDY, int Cy, int ldDY, int ntaps, const int* h_tap_x_off,
                     const int* h_tap_dy_off, const int* h_tap_w, long long rows_total, int splits, float* dW,
                     int ldw, int w_rows_per_tap, int dw_col0, void* stream);

/* ---- 2D U-Net engine ------------------------------------------------------------------------------
 * Replaces the Keras model built by mpunet/models/unet.py:26-216 (`UNet`, selected by name through
 * mpunet/models/model_init.py:10-13) and what the callers do with it: `model.predict`
 * (utils/fusion/fuse_and_predict.py:88), `model.fit` train step (train/trainer.py:246-257 with the
 * loss/optimizer of bin/defaults/MultiPlanar/train_hparams.yaml:108,125-126).
 * filters[l] = int(64 * 2^l * sqrt(complexity_factor)) for l = 0..depth (unet.py:91,120,195). */
typedef struct MpuUNetConfig {
  int H, W;            /* slice size ("dim"), divisible by 2^depth */
  int n_channels;      /* input image channels */
  int n_classes;       /* <= 16 */
  int depth;           /* number of 2x2 max-pools (reference: 4) */
  int filters[8];      /* encoder levels 0..depth-1, then bottom */
  int max_batch;
  int training;        /* 1: allocate gradient buffers */
  float bn_eps;        /* Keras default 1e-3 */
  float bn_momentum;   /* Keras default 0.99 */
} MpuUNetConfig;

typedef struct MpuLayerInfo {
  char name[64];       /* Keras layer name (unet.py:119-179; head = "conv2d") */
  int kind;            /* 0 = conv, 1 = batch-norm */
  int ksize, cin, cout, k_phys, co_phys, c0_phys;
  long long off0, off1; /* conv: kernel [k*k][co_phys][k_phys], bias [co_phys]; bn: gamma, beta (param buffer) */
  long long off2, off3; /* bn: moving mean, moving variance (bn_state buffer) */
} MpuLayerInfo;

int mpu_unet_sizes(const MpuUNetConfig* cfg, long long* n_params, long long* n_bn_state,
                   long long* workspace_bytes);
/* params / grads / adam_m / adam_v: float[n_params]; bn_state: float[n_bn_state]; all caller-owned. */
int mpu_unet_create(const MpuUNetConfig* cfg, float* params, float* grads, float* adam_m, float* adam_v,
                    float* bn_state, void* workspace, long long workspace_bytes, void* stream,
                    void** handle);
int mpu_unet_destroy(void* handle);
int mpu_unet_num_layers(void* handle);
int mpu_unet_layer_info(void* handle, int idx, MpuLayerInfo* out);
/* rebuild the bf16 GEMM operand copies after `params` changed (load_weights / external update) */
int mpu_unet_sync_weights(void* handle, void* stream);
/* the zero-bordered bf16 input tensor [B*(H+2)*(W+2)][cin_phys] the plane sampler writes into */
int mpu_unet_input_buffer(void* handle, void** ptr, int* cin_phys, long long* rows);
/* fp32 NHWC [B,H,W,n_channels] -> input buffer (what Keras' predict/fit receives) */
int mpu_unet_pack_input(void* handle, const float* x_nhwc, int B, void* stream);
/* model.predict_on_batch: softmax probabilities fp32 [B,H,W,n_classes]; bn_training=0 uses moving stats */
int mpu_unet_forward(void* handle, int B, int bn_training, float* probs_out, void* stream);
/* forward + sparse categorical cross-entropy (x sample weight) + backward; gradients land in `grads`.
 * labels uint8 [B,H,W]; loss_sum (device double) receives the SUM of per-pixel weighted losses;
 * grad_scale multiplies dlogits (1 = Keras' sum-of-unreduced-losses semantics). */
int mpu_unet_train_step(void* handle, int B, const unsigned char* labels, const float* sample_w,
                        float grad_scale, double* loss_sum, float* probs_opt, void* stream);
/* Single-process train step with the optimizer inside: mpu_unet_train_step, then per parameter range the l2 penalty
 * (l2_grad_coef != 0: see mpu_unet_l2_penalty; l2_sumsq is zeroed here) and Keras Adam step `step`, each range as soon
 * as its gradients are final - on an internal stream, overlapped with the remaining backward stages.  Identical
 * results to the three separate calls (trainer.py:246-257's model.fit step). */
int mpu_unet_train_step_adam(void* handle, int B, const unsigned char* labels, const float* sample_w,
                             float grad_scale, double* loss_sum, float* probs_opt, float lr, float beta1, float beta2,
                             float eps, int step, float l2_grad_coef, double* l2_sumsq, void* stream);
/* The same step split for data-parallel overlap: forward + loss, then backward stage 0 (up path),
 * 1 (bottom block), 2 (encoder).  mpu_unet_grad_ranges fills 4 [begin,end) float ranges of `grads`:
 * [0] complete after stage 0, [1] after stage 1, [2] and [3] after stage 2 - each can be all-reduced
 * (replacing MirroredStrategy's aggregation, bin/train.py:349) while the next stage runs. */
int mpu_unet_train_forward(void* handle, int B, const unsigned char* labels, const float* sample_w,
                           float grad_scale, double* loss_sum, float* probs_opt, void* stream);
int mpu_unet_backward_stage(void* handle, int B, int stage, void* stream);
int mpu_unet_grad_ranges(void* handle, long long* h_out8);
/* Keras Adam step t (1-based) over the whole parameter buffer; refreshes the bf16 GEMM operand copies */
int mpu_unet_adam(void* handle, float lr, float beta1, float beta2, float eps, int step,
                  float grad_scale, void* stream);
/* The same update on one [begin, end) float range of the parameter buffer (the ranges of mpu_unet_grad_ranges): a
 * data-parallel step applies it to each range as soon as that range's all-reduce has completed, so the update of
 * the early ranges overlaps the reduction of the last one.  finish != 0 (on the last range of a step) refreshes the
 * derived bf16 operand copies.  The bf16 shadow of the range is written by the same kernel. */
int mpu_unet_adam_range(void* handle, long long begin, long long end, float lr, float beta1, float beta2, float eps,
                        int step, float grad_scale, int finish, void* stream);
/* kernel_regularizer = l2(l2_reg) of every encoder / bottom / up-path conv (mpunet/models/unet.py:39,95,122-189; the
 * 1x1 head carries none): for the conv kernels inside the float range [begin, end) of the parameter buffer, adds
 * grad_coef * w to the gradient buffer and accumulates sum(w^2) into *sumsq_out (device double, caller-zeroed).
 * Call between the gradient all-reduce and Adam.  Keras adds the scalar penalty to every element of the unreduced
 * loss tensor, so grad_coef = 2 * l2_reg * grad_scale * (B*H*W) and the reported loss gains l2_reg * sumsq. */
int mpu_unet_l2_penalty(void* handle, long long begin, long long end, float grad_coef, double* sumsq_out,
                        void* stream);
int mpu_unet_debug_buffer(void* handle, int level, int which, void** ptr, long long* rows, int* C);

/* ---- oblique-plane sampler -------------------------------------------------------------------------
 * Replaces ViewInterpolator.__call__ + scaler.transform per plane: mpunet/interpolation/
 * view_interpolator.py:62-101, regular_grid_interpolator.py:204-223,252-270, sample_grid.py:227-239 and
 * sequences/isotrophic_live_view_sequence_2d.py:103-117 (inference stacks :29-101, train batches :163-216).
 *   vol [X][Y][Z][C] f32, labels [X][Y][Z] u8 (may be NULL); gx/gy/gz: float32 voxel axes
 *   (sample_grid.py:93-98); h_inv_step[3]: 1/axis spacing (search seed only); h_rot: optional 3x3
 *   (view_interpolator.py:54-60); planes: device double [n][10] = row-major basis [u v n] (9) + offset,
 *   computed on the host exactly as sample_plane_at does; span = real_space_span;
 *   h_bg_value[C] (f32), bg_class; h_center/h_scale[C]: RobustScaler statistics or NULL.
 * Outputs (each optional): out_f32 [n][dim][dim][C]; out_padded_bf16 = the U-Net input tensor
 * [n][dim+2][dim+2][cpad]; out_labels [n][dim][dim] u8. */
int mpu_sample_planes(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                      const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                      const double* h_rot, const double* planes, int n_planes, int dim, double span,
                      const float* h_bg_value, int bg_class, const double* h_center,
                      const double* h_scale, float* out_f32, void* out_padded_bf16, int cpad,
                      unsigned char* out_labels, void* stream);

/* Candidate probe of the training-batch rejection sampler (sequences/isotrophic_live_view_sequence_2d.py:119-161):
 * same plane description as mpu_sample_planes, nothing is materialised.  Per candidate plane (device uint32 [n],
 * overwritten):  class_mask = OR over pixels of (1u << nearest label), out-of-bounds pixels = bg_class - the
 * np.isin(fg_classes, lab) of validate_lab / validate_lab_vec (isotrophic_live_view_sequence.py:98-128);
 * valid = 1 iff some pixel of some channel of the UNSCALED trilinear image is not np.isclose(bg_value) - is_valid_im
 * (isotrophic_live_view_sequence.py:91-96).  Either output may be NULL.  Labels above 31 share bit 31. */
int mpu_probe_planes(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                     const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                     const double* h_rot, const double* planes, int n_planes, int dim, double span,
                     const float* h_bg_value, int bg_class, unsigned int* class_mask, unsigned int* valid,
                     void* stream);

/* ---- multi-view mapping + fusion ---------------------------------------------------------------------
 * Replaces map_real_space_pred per view (mpunet/utils/fusion/fuse_and_predict.py:92-137) fused with
 * FusionLayer.call + argmax (models/fusion_model.py:38-39, bin/predict.py:349-366, utils/utils.py:311-328).
 *   h_pred_ptrs[V]: device pointers, each [n_planes][dim][dim][C] f32 softmax of one view's plane stack;
 *   inv_basis: device double [V][9]; ax: device double [dim] (in-plane axis); offsets: device double
 *   [V][n_planes]; h_inv_step[1+V]: 1/spacing of ax and of each offsets row; h_dims[3] = volume shape;
 *   h_affine3x3 / h_mean[3]: voxel->real transform and the centre get_voxel_grid_real_space subtracts
 *   (sample_grid.py:101-130); W [V][C], b [C] fusion weights (sum_fusion=1: plain sum over views).
 * Outputs (each optional): labels_out [X][Y][Z] u8; probs_out [X][Y][Z][C] f32; combined_out
 * [V][X][Y][Z][C] f32 (the per-view mapped volumes, what the reference materialises). */
int mpu_map_fuse(const void* const* h_pred_ptrs, int V, int C, int dim, int n_planes,
                 const double* inv_basis, const double* ax, const double* offsets,
                 const double* h_inv_step, const int* h_dims, const double* h_affine3x3,
                 const double* h_mean, const float* W, const float* b, int sum_fusion,
                 unsigned char* labels_out, float* probs_out, float* combined_out, void* stream);

/* The same operation when every axis is an np.linspace (what the reference's grids are: in-plane axis
 * np.linspace(-hd, hd, dim), sample_grid.py:227-233 in test mode; plane offsets np.linspace(-bounds, bounds, n),
 * sequences/isotrophic_live_view_sequence_2d.py:62): the axis values g[i] = fl(fl(i*step) + start), g[n-1] = stop are
 * recomputed in registers instead of looked up, and the C probabilities of a pixel are gathered with aligned 16-byte
 * loads.  The CALLER guarantees the tables equal that formula bit for bit (the Python shim checks and otherwise calls
 * mpu_map_fuse); results are then identical to mpu_map_fuse.  All h_ arguments are host pointers:
 * h_inv_basis [V][9], h_ax_lin [3] = start, step, stop, h_off_lin [V][3].  2 <= C <= 8. */
int mpu_map_fuse_linspace(const void* const* h_pred_ptrs, int V, int C, int dim, int n_planes,
                          const double* h_inv_basis, const double* h_ax_lin, const double* h_off_lin,
                          const int* h_dims, const double* h_affine3x3, const double* h_mean, const float* W,
                          const float* b, int sum_fusion, unsigned char* labels_out, float* probs_out,
                          float* combined_out, void* stream);

/* ---- fusion-layer training ---------------------------------------------------------------------------
 * Replaces FusionModel.fit's train step (bin/train_fusion.py:196-213; loss evaluate/loss_functions.py:
 * 207-246 with uniform weights; regulariser models/fusion_model.py:9-11).  X [n][V][C] f32, y [n] u8.
 * mpu_fusion_grad ADDS into accum (double [V*C + C + 1] = dW | db | sum of per-point losses); all-reduce
 * accum across ranks, then mpu_fusion_adam applies mean gradient + regulariser with Keras Adam.
 * mpu_fusion_adam with n_points <= 0 takes the point count from accum[V*C + C + 1] (device): the count is all-reduced in
 * the same message as the sums, so a multi-rank step needs no host synchronisation. */
int mpu_fusion_grad(const float* X, const unsigned char* y, long long n, int V, int C, const float* W,
                    const float* b, double* accum, void* stream);
/* The same with a shuffled epoch: point i of the batch is row index[i] of X / y (device int64 [n]) - Keras'
 * fit(shuffle=True) without materialising X[perm].  index == NULL: rows 0..n-1. */
int mpu_fusion_grad_indexed(const float* X, const unsigned char* y, const long long* index, long long n, int V, int C,
                            const float* W, const float* b, double* accum, void* stream);
/* Single-process train step in ONE launch: gradient sums, then the last block to finish applies the Adam update
 * (+ regulariser), writes the batch's mean dice loss to loss_out (device double, optional), and zeroes accum and
 * counter for the next batch.  accum (double [V*C + C + 1]) and counter (see mpu_fusion_train_epoch) must be zero
 * before the first call.
 * With several ranks use mpu_fusion_grad_indexed + all-reduce + mpu_fusion_adam instead. */
int mpu_fusion_train_step(const float* X, const unsigned char* y, const long long* index, long long n, int V, int C,
                          float* W, float* b, float* m, float* v, double* accum, unsigned int* counter,
                          double* loss_out, float reg, float lr, float beta1, float beta2, float eps, int step,
                          void* stream);
/* One single-process EPOCH: ceil(n / batch) fused train steps over the rows perm[0..n) (device int64; NULL = rows in
 * order), Adam step numbers first_step, first_step + 1, ...; losses_out (device double [ceil(n / batch)], optional)
 * receives every batch's mean dice loss.  The host loop of FusionModel.fit (bin/train_fusion.py:196-213) in C: one
 * launch per batch, nothing else between them.
 * `counter` (here and in mpu_fusion_train_step / _epoch_peer) is a device buffer of mpu_fusion_scratch_bytes() bytes,
 * zeroed once: the 16-byte arrival counter followed by space reserved for per-block partial sums.  (Combining the
 * blocks' partials in a fixed order by the last block was measured slower than the 36 double atomics per block -
 * 7.7 vs 4.7 ms per epoch - and is switched off; sums are accumulated with fp64 atomics, whose order varies.) */
int mpu_fusion_scratch_bytes(void);
int mpu_fusion_train_epoch(const float* X, const unsigned char* y, const long long* perm, long long n,
                           long long batch, int V, int C, float* W, float* b, float* m, float* v, double* accum,
                           unsigned int* counter, double* losses_out, float reg, float lr, float beta1, float beta2,
                           float eps, int first_step, void* stream);
/* One MULTI-RANK epoch with the gradient exchange fused into the train-step kernel (the compute step followed by a
 * collective of bin/train_fusion.py's data-parallel fit): after its local sums, every rank's last block stores them
 * into every peer's mailbox over NVLink peer memory, raises the peer's flag to the step's sequence number, waits for
 * all flags of its own mailbox, adds the contributions in rank order and applies Adam - one launch per batch, no
 * NCCL call and no host round trip; parameters stay bit-identical on all ranks.
 *   h_peer_mail[world]: device pointers to every rank's mailbox (mpu_fusion_mailbox_bytes() bytes each, zeroed once,
 *   allocated in peer-mapped memory, e.g. torch.distributed._symmetric_memory); every rank calls this with the same
 *   n_batches / first_step / first_seq (sequence numbers must be >= 1 and never reused); a rank whose points run out
 *   before n_batches contributes empty batches.  world <= 8. */
int mpu_fusion_mailbox_bytes(void);
int mpu_fusion_train_epoch_peer(const float* X, const unsigned char* y, const long long* perm, long long n,
                                long long batch, long long n_batches, int V, int C, float* W, float* b, float* m,
                                float* v, double* accum, unsigned int* counter, double* losses_out, float reg, float lr,
                                float beta1, float beta2, float eps, int first_step, const void* const* h_peer_mail,
                                int world, int rank, unsigned long long first_seq, void* stream);
int mpu_fusion_adam(float* W, float* b, float* m, float* v, const double* accum, double n_points,
                    int V, int C, float reg, float lr, float beta1, float beta2, float eps, int step,
                    void* stream);

/* Interpolation of the resident volume at explicit real-space coordinates: ViewInterpolator.__call__ /
 * intrp_image / intrp_labels on an arbitrary grid (interpolation/view_interpolator.py:62-101, after the optional
 * apply_rotation, which the caller does on the host).  coords: device [3][n] float64 (x | y | z).  out_f32 [n][C]
 * (trilinear, float64 weights, cast to f32, out-of-bounds -> h_bg_value) and/or out_labels [n] (nearest,
 * out-of-bounds -> bg_class); no scaling is applied (the reference scales after interpolation). */
int mpu_interp_points(const float* vol, const unsigned char* labels, const int* h_dims, int C,
                      const float* gx, const float* gy, const float* gz, const double* h_inv_step,
                      const double* coords, long long n, const float* h_bg_value, int bg_class,
                      float* out_f32, unsigned char* out_labels, void* stream);

/* Leaf sums of numpy's pairwise summation of the real-space voxel coordinates A . (i, j, k) (C order over the volume),
 * for the exact centre that get_voxel_grid_real_space subtracts: np.mean(grid_points_real_space, axis=0),
 * interpolation/sample_grid.py:117-118.  leaf_start / leaf_len (device, [n_leaves]) are the <= 128-element blocks of
 * numpy's pairwise_sum recursion; out (device double [3][n_leaves]) receives each block's sum per coordinate, summed
 * in numpy's 8-accumulator order.  The host combines the binary tree (interpolation/voxel_center.py). */
int mpu_voxel_leaf_sums(const int* h_dims, const double* h_affine3x3, const long long* leaf_start, const int* leaf_len,
                        long long n_leaves, double* out, void* stream);

/* Confusion-matrix counts for the validation callback and dice_all: ADDS into counts [3][n_classes] int64
 * (device): [0] true == c & pred == c (TP), [1] true == c (relevant), [2] pred == c (selected).
 * pred is either y_pred (u8 labels) or, when scores != NULL, the first arg-max of scores [n][n_classes] f32.
 * Replaces callbacks/validation.py:117-131 (np.bincount x3) and the sums of evaluate/metrics.py:12-52. */
int mpu_label_counts(const unsigned char* y_true, const unsigned char* y_pred, const float* scores,
                     long long n, int n_classes, long long* counts, void* stream);

/* Elastic2D augmentation of n slices (augmentation/elastic_deformation.py:6-69, called from
 * augmentation/augmenters.py:87-109 after scaling).  All pointers are device pointers.
 *   x_in [n][H][W][C] f32, y_in [n][H][W] u8 or NULL -> x_out, y_out (not in place)
 *   fields [n][2][H][W] f64: in = the uniform(-1,1) noise images (dx, dy) drawn by the caller's RNG;
 *                            out = gaussian_filter(noise, sigma, mode="constant") (before the alpha factor)
 *   scratch: same size as fields
 *   d_weights [n][weight_stride] f64: normalised Gaussian taps of each slice, 2*radius+1 values from index 0
 *   d_radius [n] int (= int(4*sigma + 0.5)), d_alpha [n] f64, d_bg [n][C] f32 (fill value per channel)
 * Image: bilinear with float64 weights -> f32; labels: nearest, out-of-bounds -> 0. */
int mpu_elastic_2d(const float* x_in, const unsigned char* y_in, double* fields, double* scratch,
                   const double* d_weights, int weight_stride, const int* d_radius, const double* d_alpha,
                   const float* d_bg, int n, int H, int W, int C, float* x_out, unsigned char* y_out,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPUNET_B200_H_ */
