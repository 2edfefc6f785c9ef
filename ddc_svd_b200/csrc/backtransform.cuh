// backtransform.cuh — internal interface of the compact-WY back-transform (backtransform.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
constexpr int WY_MAX_SPLIT = 16;
// C (rows x nc, ldc, device) <- H_0 H_1 ... H_{nref-1} C with the reflectors stored in A
// (device, lda) the way bidiag leaves them: left=1 column reflectors A[j:rows, j] (multU,
// bidiag_par.c:1046-1095), left=0 row reflectors A[j, j+1:rows] (multV, :990-1043).
size_t backtransform_workspace_bytes(int rows, int nref, int nc);
void wy_apply_device(int left, int rows, int nref, const double *A, long lda, double *C, long ldc, int nc,
                     void *workspace, cudaStream_t st);
// Householder QR of a tall matrix in the same reflector convention (for the QR-first path): A is
// overwritten with the reflectors (diagonal and below) and R's strict upper triangle, R (n x n)
// receives the triangular factor; wy_apply_device(left = 1, rows = m, nref = n, A, ...) applies Q.
size_t qr_workspace_bytes(int m, int n);
void qr_device(int m, int n, double *A, long lda, double *R, long ldr, void *workspace, cudaStream_t st);
}
