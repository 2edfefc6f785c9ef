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
struct ProgressHook;
void qr_device(int m, int n, double *A, long lda, double *R, long ldr, void *workspace, cudaStream_t st,
               const ProgressHook *hook = nullptr);
// The same back-transform in two halves, so that the panel set-up (which depends on the reflectors only)
// can run while the factorization is still producing the later reflectors, on another stream, and so
// that other devices can be handed the prepared panels instead of the reflector matrix:
//   wy_setup_device      panels [pb, pe) of NBW reflectors: gather, Gram, T, VT = V T  -> `panels`
//   wy_apply_prepared    C <- H_0 ... H_{nref-1} C reading only V and VT of `panels`
//   wy_panel_slices      the V / VT column ranges of panels [pb, pe) (contiguous: what gets broadcast)
int wy_panel_width();
int wy_panel_count(int nref);
size_t wy_panels_bytes(int rows, int nref);
size_t wy_apply_workspace_bytes(int nc);
void wy_setup_device(int left, int rows, int nref, const double *A, long lda, void *panels, int pb, int pe,
                     cudaStream_t st);
void wy_apply_prepared(int left, int rows, int nref, const void *panels, double *C, long ldc, int nc,
                       void *workspace, cudaStream_t st);
void wy_panel_slices(void *panels, int rows, int nref, int pb, int pe, double **V, double **VT, size_t *count);
}
