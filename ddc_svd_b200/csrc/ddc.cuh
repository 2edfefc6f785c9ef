// ddc.cuh — internal interface of the dDC singular-value phase (ddc.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
// Singular values (ascending) of the N x (N+1) upper bidiagonal with diagonal b1[N] and
// super-diagonal b2[N] (b2[N-1] = 0 for a square B) — the contract of the reference's
// GetSingularValues_Parallel (Calculations-Parallel.c:852-874).  Device pointers.
// Only enqueues on `st` (the merge-tree tables are cached on the device per N).
size_t ddc_workspace_bytes(int N);
void ddc_values_device(int N, const double *b1, const double *b2, double *sigma, void *workspace,
                       cudaStream_t st);
}
