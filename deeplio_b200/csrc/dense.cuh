// Internal launchers of the skinny dense kernels (dense.cu), shared with rnn.cu.
#pragma once
#include <cuda_runtime.h>

namespace dlio {
// y = act(x w^T + b + b2)
int linear_fwd_launch(const float *x, int ldx, const float *w, const float *b, const float *b2, int M, int N, int K,
                      int act, float *y, int ldy, cudaStream_t st);
// dx += dz w        (dx must be zeroed by the caller; accumulates atomically)
int linear_dx_launch(const float *dz, int lddz, const float *w, int M, int N, int K, float *dx, int lddx,
                     cudaStream_t st);
// dw = dz^T x       (overwrites); db (optional): out[n] = sum_m dz[m, n] from the same launch
int linear_dw_launch(const float *dz, int lddz, const float *x, int ldx, int M, int N, int K, float *dw,
                     cudaStream_t st, float *db = nullptr);
// out[n] = sum_m dz[m, n]
int colsum_launch(const float *dz, int lddz, int M, int N, float *out, cudaStream_t st);
}  // namespace dlio
