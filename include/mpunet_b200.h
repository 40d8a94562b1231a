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

/* ---- kernel-level surface: multi-tap GEMM on tcgen05 (bring-up + kernel parity tests) ------------
 * The contraction inside tf.keras Conv2D / its gradients, reference call sites
 * mpunet/models/unet.py:120-127,137-144,159-163,171-178 (forward) and the Keras autodiff of
 * train/trainer.py:246 (backward).  Activations are bf16 [rows][channels] matrices in zero-bordered
 * NHWC layout; W is bf16 [w_taps][n_phys][k_total]. */
int mpu_mtgemm_fwd(const void* A0, long long rowsA0, int C0, int ldA0, const void* A1,
                   long long rowsA1, int C1, int ldA1, const void* W, int w_taps, int n_phys,
                   int k_total, int ntaps, const int* h_tap_a_off, const int* h_tap_w, int M_rows,
                   int BN, int Hp, int Wp, int oHp, int oWp, int s, int py, int px, void* out, int ldo,
                   const float* bias, const void* mask, int ldm, int relu, void* stream);

int mpu_mtgemm_wgrad(const void* X, long long rowsX, int Cx, int ldX, const void* dY,
                     long long rowsDY, int Cy, int ldDY, int ntaps, const int* h_tap_x_off,
                     const int* h_tap_w, int ngroups, const int* h_group_first,
                     const int* h_group_count, const int* h_group_dy_off, int rows_total, int BN,
                     int splits, float* dW, int ldw, int w_rows_per_tap, int dw_col0, int ci_valid,
                     int co_valid, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int kstep_bytes,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPUNET_B200_H_ */
