// twisted.cuh — internal interface of the twisted-factorization vector phase (twisted.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
// Bidiagonal B: n rows, mb = n or n+1 columns, diagonal a[n], super-diagonal b[mb-1].
// Computes, for the ns singular values sigma_all[i0 .. i0+ns) (sigma_all has ntot entries,
// ascending), the right vectors X(:,t) (length mb, X[t*ldx + j] — the reference's layout,
// parallel-twisted.c:94,497) and, if Y != NULL, the left vectors Y[t*ldy + j] = (B x_t)_j/sigma_t
// (parallel-twisted.c:545-549).  sigma_out (optional, ns entries) receives the Rayleigh-quotient
// polished values.  Device pointers; everything is enqueued on `st`.
size_t twisted_workspace_bytes(int n, int mb, int ns);
void twisted_vectors_device(int n, int mb, const double *a, const double *b, const double *sigma_all,
                            int ntot, int i0, int ns, double *X, long ldx, double *Y, long ldy,
                            double *sigma_out, int rqi_steps, void *workspace, cudaStream_t st);
}
