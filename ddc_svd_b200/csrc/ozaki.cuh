// ozaki.cuh — internal interface of the tcgen05 (int8, TMEM) FP64-accurate update GEMM (ozaki.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
// C (M x N, ldc) += sign * A (M x 128, lda) * B (128 x N, ldb), all FP64 column-major on the device; the
// product is formed from 8 x 8 int8 slice products on the 5th-generation tensor cores and recombined to
// FP64 accuracy (error-free slicing).  K must be 128 (the compact-WY panel width).
bool ozaki_update_supported(int M, int N, int K);
size_t ozaki_workspace_bytes(int M, int N);
void ozaki_update_device(int M, int N, double sign, const double *A, long lda, const double *B, long ldb,
                         double *C, long ldc, void *workspace, cudaStream_t st);
}
