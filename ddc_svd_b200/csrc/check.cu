// check.cu — residual / orthogonality of a computed SVD, on the device.
//
// The reference's driver carries this check switched off (test-whole-svd.c:81-96: transpose V, form
// diag(sigma), two dgemm_simple calls, Frobenius norm of A - U S V^T over the norm of A; it does not
// link there because l2_norm_mat is defined nowhere).  This is its enabled form for matrices that live
// on the GPU: the same quantities — plus the two orthogonality defects the north star names — from
// blocked FP64 DMMA GEMMs (dgemm_dmma.cu) and fixed-order reductions, in O(block) extra memory.
//   out[0] = ||U^T U - I||_F        out[1] = ||V^T V - I||_F
//   out[2] = ||A - U S V^T||_F / ||A||_F      (all min(m,n) columns given)
//          = ||A V - U S||_F / ||A||_F        (a column block of the factors: what one rank of a sharded run holds)
//   out[3] = |sum sigma^2 - ||A||_F^2| / ||A||_F^2   (checksum of checksums; all columns only)
//   out[4] = ||A||_F                out[5] = 1 if sigma is ascending (Calculations-Parallel.c:56) else 0
#include "common.cuh"
#include "check.cuh"

namespace svdgpu {

constexpr int CK_PARTS = 1024;          // per-CTA partial sums, accumulated across launches in stream order

// parts[blockIdx.x] += sum over the rows x cols block of (X[r,c] - (r == c + diag ? 1 : 0))^2   (diag < 0: no identity)
__global__ void __launch_bounds__(256)
ck_sq_accum_kernel(const double *__restrict__ X, long ldx, int rows, int cols, int diag, double *__restrict__ parts)
{
    __shared__ double red[8];
    double acc = 0.0;
    const long total = (long)rows * cols;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int r = (int)(e % rows), c = (int)(e / rows);
        double v = X[r + (long)c * ldx];
        if (diag >= 0 && r == c + diag) v -= 1.0;
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        parts[blockIdx.x] += s;
    }
}

// B[k, jj] = sigma[k] * V[j0 + jj, k]     (nc x w): the block of S V^T that multiplies U
__global__ void ck_form_svt_kernel(int nc, int w, int j0, const double *__restrict__ sigma, const double *__restrict__ V,
                                   long ldv, double *__restrict__ B, long ldb)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nc * w) return;
    const int jj = (int)(idx % w), k = (int)(idx / w);          // consecutive threads walk a column of V
    B[k + (long)jj * ldb] = sigma[k] * V[(j0 + jj) + (long)k * ldv];
}

// R[:, jj] -= sigma[j0 + jj] * U[:, j0 + jj]
__global__ void ck_sub_scaled_kernel(int m, int w, int j0, const double *__restrict__ sigma, const double *__restrict__ U,
                                     long ldu, double *__restrict__ R, long ldr)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)m * w) return;
    const int r = (int)(idx % m), jj = (int)(idx / m);
    R[r + (long)jj * ldr] -= sigma[j0 + jj] * U[r + (long)(j0 + jj) * ldu];
}

__global__ void ck_copy_block_kernel(int m, int w, const double *__restrict__ A, long lda, double *__restrict__ R, long ldr)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)m * w) return;
    const int r = (int)(idx % m), jj = (int)(idx / m);
    R[r + (long)jj * ldr] = A[r + (long)jj * lda];
}

// one CTA: fixed-order sums of the four partial arrays, sigma checks, final figures
__global__ void __launch_bounds__(256)
ck_final_kernel(const double *__restrict__ parts, const double *__restrict__ sigma, int nc, int full, double *__restrict__ out)
{
    __shared__ double red[8];
    __shared__ double tot[6];
    const int t = threadIdx.x;
    for (int q = 0; q < 4; ++q) {
        double a = 0.0;
        for (int p = t; p < CK_PARTS; p += 256) a += parts[q * CK_PARTS + p];
        a = warp_sum(a);
        if ((t & 31) == 0) red[t >> 5] = a;
        __syncthreads();
        if (t == 0) { double s = 0.0; for (int w = 0; w < 8; ++w) s += red[w]; tot[q] = s; }
        __syncthreads();
    }
    double s2 = 0.0, bad = 0.0;
    for (int k = t; k < nc; k += 256) {
        s2 += sigma[k] * sigma[k];
        if (k > 0 && !(sigma[k] >= sigma[k - 1])) bad += 1.0;
    }
    s2 = warp_sum(s2);
    if ((t & 31) == 0) red[t >> 5] = s2;
    __syncthreads();
    if (t == 0) { double s = 0.0; for (int w = 0; w < 8; ++w) s += red[w]; tot[4] = s; }
    __syncthreads();
    bad = warp_sum(bad);
    if ((t & 31) == 0) red[t >> 5] = bad;
    __syncthreads();
    if (t == 0) {
        double b = 0.0;
        for (int w = 0; w < 8; ++w) b += red[w];
        const double a2 = tot[3];
        out[0] = sqrt(tot[0]);
        out[1] = sqrt(tot[1]);
        out[2] = (a2 > 0.0) ? sqrt(tot[2] / a2) : sqrt(tot[2]);
        out[3] = (full && a2 > 0.0) ? fabs(tot[4] - a2) / a2 : 0.0;
        out[4] = sqrt(a2);
        out[5] = (b == 0.0) ? 1.0 : 0.0;
    }
}

static int ck_block(int nc) { return nc < 1024 ? nc : 1024; }

size_t check_workspace_bytes(int m, int n, int nc)
{
    const int w = ck_block(nc);
    const size_t big = (size_t)(m > n ? m : n) > (size_t)nc ? (size_t)(m > n ? m : n) : (size_t)nc;
    return ((size_t)big * w + (size_t)nc * w + 4 * CK_PARTS + 8) * sizeof(double) + 256;
}

static void sq_accum(const double *X, long ldx, int rows, int cols, int diag, double *parts, cudaStream_t st)
{
    long total = (long)rows * cols;
    int grid = (int)((total + 255) / 256);
    if (grid > CK_PARTS) grid = CK_PARTS;
    if (grid < 1) grid = 1;
    ck_sq_accum_kernel<<<grid, 256, 0, st>>>(X, ldx, rows, cols, diag, parts);
    SVD_KERNEL_CHECK();
}

void check_device(int m, int n, const double *A0, long lda, const double *sigma, const double *U, long ldu,
                  const double *V, long ldv, int nc, double *out_dev, void *workspace, cudaStream_t st)
{
    const int mn = m < n ? m : n;
    const int full = (nc == mn);
    const int w = ck_block(nc);
    const size_t big = (size_t)(m > n ? m : n) > (size_t)nc ? (size_t)(m > n ? m : n) : (size_t)nc;
    double *R = (double *)workspace;               // big x w
    double *B = R + big * w;                       // nc x w
    double *parts = B + (size_t)nc * w;            // [4][CK_PARTS]: orthU, orthV, resid, |A|^2
    SVD_CUDA_CHECK(cudaMemsetAsync(parts, 0, sizeof(double) * 4 * CK_PARTS, st));
    auto gemm = [&](int tA, int tB, int M, int N, int K, double alpha, const double *a, long la, const double *b, long lb,
                    double beta, double *c, long lc) {
        GemmArgs g = {};
        g.M = M; g.N = N; g.K = K; g.A = a; g.lda = la; g.transA = tA; g.B = b; g.ldb = lb; g.transB = tB;
        g.C = c; g.ldc = lc; g.alpha = alpha; g.beta = beta; g.batch = 1; g.splitk = 1;
        dgemm_dmma(g, st);
    };
    // orthogonality, by column blocks:  G = Q^T Q[:, J]  (nc x wj), identity on rows j0 + jj
    for (int side = 0; side < 2; ++side) {
        const double *Q = side ? V : U;
        const long ldq = side ? ldv : ldu;
        const int rows = side ? n : m;
        for (int j0 = 0; j0 < nc; j0 += w) {
            const int wj = (nc - j0 < w) ? nc - j0 : w;
            gemm(1, 0, nc, wj, rows, 1.0, Q, ldq, Q + (long)j0 * ldq, ldq, 0.0, R, nc);
            sq_accum(R, nc, nc, wj, j0, parts + side * CK_PARTS, st);
        }
    }
    if (full) {
        // residual by column blocks of A:  R = A[:, J] - U (S V[J, :]^T)
        for (int j0 = 0; j0 < n; j0 += w) {
            const int wj = (n - j0 < w) ? n - j0 : w;
            ck_copy_block_kernel<<<ceil_div((long)m * wj, 256), 256, 0, st>>>(m, wj, A0 + (long)j0 * lda, lda, R, m);
            SVD_KERNEL_CHECK();
            sq_accum(R, m, m, wj, -1, parts + 3 * CK_PARTS, st);
            ck_form_svt_kernel<<<ceil_div((long)nc * wj, 256), 256, 0, st>>>(nc, wj, j0, sigma, V, ldv, B, nc);
            SVD_KERNEL_CHECK();
            gemm(0, 0, m, wj, nc, -1.0, U, ldu, B, nc, 1.0, R, m);
            sq_accum(R, m, m, wj, -1, parts + 2 * CK_PARTS, st);
        }
    } else {
        // a block of the factors:  R = A V[:, J] - U[:, J] S_J ; ||A||_F separately
        for (int j0 = 0; j0 < n; j0 += w) {
            const int wj = (n - j0 < w) ? n - j0 : w;
            sq_accum(A0 + (long)j0 * lda, lda, m, wj, -1, parts + 3 * CK_PARTS, st);
        }
        for (int j0 = 0; j0 < nc; j0 += w) {
            const int wj = (nc - j0 < w) ? nc - j0 : w;
            gemm(0, 0, m, wj, n, 1.0, A0, lda, V + (long)j0 * ldv, ldv, 0.0, R, m);
            ck_sub_scaled_kernel<<<ceil_div((long)m * wj, 256), 256, 0, st>>>(m, wj, j0, sigma, U, ldu, R, m);
            SVD_KERNEL_CHECK();
            sq_accum(R, m, m, wj, -1, parts + 2 * CK_PARTS, st);
        }
    }
    ck_final_kernel<<<1, 256, 0, st>>>(parts, sigma, nc, full, out_dev);
    SVD_KERNEL_CHECK();
}

} // namespace svdgpu
