// check.cuh — internal interface of the on-device SVD checker (check.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
// A0 (m x n, the ORIGINAL matrix), sigma[nc], U (m x nc), V (n x nc), all on the device; nc = min(m,n) for a
// whole SVD, fewer for a column block of the factors.  out_dev[6]: see check.cu.  Enqueued on `st`.
size_t check_workspace_bytes(int m, int n, int nc);
void check_device(int m, int n, const double *A0, long lda, const double *sigma, const double *U, long ldu,
                  const double *V, long ldv, int nc, double *out_dev, void *workspace, cudaStream_t st);
}
