// scaling.cu — power-of-two pre-scaling of the input so that squares (a^2, sigma^2, pole
// products) neither overflow nor underflow.  The reference has no such guard: with entries near
// 1e+-150 its norms (normsq_matcol.cl) and dDC terms (Calculations-Parallel.c:76) already leave
// the double range.  The factor is an exact power of two, so the stored reflectors — which are
// scale invariant — and every rounding are unchanged; sigma is scaled back at the end.
#include "common.cuh"
#include <cfloat>

namespace svdgpu {

__global__ void absmax_kernel(const double *__restrict__ A, long lda, int m, int n, double *__restrict__ part)
{
    __shared__ double red[32];
    double mx = 0.0;
    const long total = (long)m * n;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e % m, c = e / m;
        mx = fmax(mx, fabs(A[r + c * lda]));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, red[w]);
        part[blockIdx.x] = mx;
    }
}

// sc[0] = factor applied to A, sc[1] = its inverse (applied to sigma); 1 when no scaling is needed
__global__ void scale_decide_kernel(const double *__restrict__ part, int nparts, double *__restrict__ sc)
{
    double mx = 0.0;
    for (int p = threadIdx.x; p < nparts; p += 32) mx = fmax(mx, part[p]);
    mx = warp_max(mx);
    if (threadIdx.x == 0) {
        double f = 1.0, fi = 1.0;
        if (mx > 0.0 && isfinite(mx) && (mx < 1e-100 || mx > 1e100)) {
            const int e = ilogb(mx);
            f = scalbn(1.0, -e);
            fi = scalbn(1.0, e);
        }
        sc[0] = f; sc[1] = fi;
    }
}

__global__ void scale_apply_kernel(double *__restrict__ A, long lda, int m, int n, const double *__restrict__ sc)
{
    const double f = sc[0];
    if (f == 1.0) return;
    const long total = (long)m * n;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e % m, c = e / m;
        A[r + c * lda] *= f;
    }
}

__global__ void scale_vec_kernel(double *__restrict__ x, int n, const double *__restrict__ factor)
{
    const double f = *factor;
    if (f == 1.0) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t] *= f;
}

void scale_matrix_device(int m, int n, double *A, long lda, double *sc, double *work, cudaStream_t st)
{
    const int nblk = 1024;
    absmax_kernel<<<nblk, 256, 0, st>>>(A, lda, m, n, work);
    SVD_KERNEL_CHECK();
    scale_decide_kernel<<<1, 32, 0, st>>>(work, nblk, sc);
    SVD_KERNEL_CHECK();
    scale_apply_kernel<<<nblk, 256, 0, st>>>(A, lda, m, n, sc);
    SVD_KERNEL_CHECK();
}
void scale_vector_device(int n, double *x, const double *factor, cudaStream_t st)
{
    scale_vec_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, n, factor);
    SVD_KERNEL_CHECK();
}

// At (n x m, ldat) = A^T (A is m x n, lda): 32 x 32 tiles through padded shared memory
__global__ void transpose_kernel(const double *__restrict__ A, long lda, int m, int n, double *__restrict__ At, long ldat)
{
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        if (r < m && c < n) tile[j][threadIdx.x] = A[r + (long)c * lda];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + threadIdx.x, r = r0 + j;
        if (r < m && c < n) At[c + (long)r * ldat] = tile[threadIdx.x][j];
    }
}
void transpose_device(int m, int n, const double *A, long lda, double *At, long ldat, cudaStream_t st)
{
    if (m <= 0 || n <= 0) return;
    const dim3 grid(ceil_div(m, 32), ceil_div(n, 32));
    if (grid.y > 65535) { fprintf(stderr, "transpose_device: n too large (%d)\n", n); abort(); }
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(A, lda, m, n, At, ldat);
    SVD_KERNEL_CHECK();
}

} // namespace svdgpu
