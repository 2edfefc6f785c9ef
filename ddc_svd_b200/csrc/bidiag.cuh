// bidiag.cuh — internal interface of the bidiagonalization phase (bidiag.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace svdgpu {
constexpr int BIDIAG_MAX_SPLIT = 64;
// A: m x n column-major on the device, leading dimension lda (even, >= round_up(m,2); rows
// [m, lda) must hold finite values).  On return A holds the reflectors exactly as the
// reference's bidiag_par() leaves them (bidiag_par.c:310-397), alpha[min(m,n)],
// beta[n-1 if m >= n else m].  All work is enqueued on `st`; no host synchronisation.
size_t bidiag_workspace_bytes(int m, int n, long lda);
// Progress hook of the factorizations (bidiag_device, qr_device): fn(user, done, st) is called on the HOST,
// at enqueue time, whenever the work enqueued on `st` so far leaves reflectors [0, done) final in A
// (every `every` steps, and once at the end with done = the reflector count).  The callee typically
// records an event on `st` and lets another stream prepare / ship those reflectors (svd_gpu.c).
struct ProgressHook { void (*fn)(void *user, int done, cudaStream_t st); void *user; int every; };
void bidiag_device(int m, int n, double *A, long lda, double *alpha, double *beta, void *workspace,
                   int nb, cudaStream_t st, const ProgressHook *hook = nullptr);
// which = 0: one gemvT pass, 1: one gemvN pass over the full matrix (workspace as above; its
// vector buffers must have been initialised, e.g. by a previous bidiag_device call or a memset)
// first step of the on-chip tail (bidiag_tail.cuh) for `ctas` co-resident CTAs; min(m,n) when it never fits
int bidiag_tail_start(int m, int n, int nb, int ctas);
void bidiag_pass_probe(int m, int n, const double *A, long lda, void *workspace, int which, cudaStream_t st);
}
