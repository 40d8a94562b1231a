// Error plumbing shared by all translation units of libmpunet_b200.so.
#pragma once
#include <cuda_runtime.h>

#define MPU_OK 0
#define MPU_ERR_ARG (-1)
#define MPU_ERR_CUDA (-2)
#define MPU_ERR_STATE (-3)
#define MPU_ERR_NOMEM (-4)

namespace mpu {
// printf-style; stores a thread-local message retrievable through mpu_last_error().
void set_error(const char* fmt, ...);
const char* last_error();
// every kernel launch of this library bumps this counter (bench.py reports it as gpu_launches)
extern long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }
// optional live timing of the tensor-core GEMM launches with CUDA events (roofline.achieved in bench.py)
void gemm_timer_begin(cudaStream_t st);
void gemm_timer_end(cudaStream_t st);
bool gemm_timer_on();
// multiprocessor count of the CURRENT device (cached per device ordinal)
int sm_count();
}  // namespace mpu

#define MPU_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::mpu::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                       cudaGetErrorString(_e));                                          \
      return MPU_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MPU_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != MPU_OK) return _r; \
  } while (0)
