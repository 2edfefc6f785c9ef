// backtransform.cu — U = Q_L [Y;0], V = Q_R [X;0] in compact-WY form on the FP64 DMMA pipe.
//
// Replaces the reference's BLAS1 back-transform: svd_gpu.c:117-121 calls multU / multV
// (bidiag_par.c:1046-1095 / :990-1043) once per singular vector, each applying the
// Householder reflectors one at a time (dot + axpy), after a host transpose() of the
// reflector matrix (matrix_helper.c:166-174, svd_gpu.c:103).
//
// Here the reflectors (unit-norm v_j, H_j = I - 2 v_j v_j^T, exactly as bidiag leaves them
// in A) are grouped into panels of NB: H_p0 ... H_{p0+NB-1} = I - V T V^T with
//     T = (striu(V^T V) + 1/2 I)^{-1}            (all tau_j = 2),
// and every panel is applied to all vectors at once with two GEMMs
//     W = V^T C ,  C -= (V T) W
// on the DMMA kernel of dgemm_dmma.cu.  The reflector matrix is never transposed on the
// host: extract_right_kernel gathers the row reflectors into column form in one tiled pass.
// Flop count is the reference's 4*(m-j) per reflector per vector, but as BLAS3.
#include "common.cuh"
#include "backtransform.cuh"

namespace svdgpu {

constexpr int NBW = 128;    // WY panel width (K of the update GEMM: 128 keeps it DMMA-bound, not C-traffic-bound)

// VL[r, j] = A[r, j] for r >= j (left reflector j lives in column j from the diagonal down)
__global__ void extract_left_kernel(int m, int nL, int nLpad, const double *__restrict__ A, long lda,
                                    double *__restrict__ VL, long ldv)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)ldv * nLpad) return;
    int r = (int)(idx % ldv), j = (int)(idx / ldv);
    double v = 0.0;
    if (j < nL && r < m && r >= j) v = A[r + (long)j * lda];
    VL[idx] = v;
}

// VR[c, j] = A[j, c] for c >= j+1 (right reflector j lives in row j right of the super-diagonal)
__global__ void __launch_bounds__(256)
extract_right_kernel(int n, int nR, int nRpad, const double *__restrict__ A, long lda,
                     double *__restrict__ VR, long ldv)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {          // read A[j0+tx, c0+q] : rows contiguous over tx
        int j = j0 + tx, c = c0 + q;
        double v = 0.0;
        if (j < nR && c < n && c >= j + 1) v = A[j + (long)c * lda];
        tile[q][tx] = v;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {          // write VR[c0+tx, j0+q]
        int c = c0 + tx, j = j0 + q;
        if (c < ldv && j < nRpad) VR[c + (long)j * ldv] = (c < n) ? tile[tx][q] : 0.0;
    }
}

// T = (striu(G) + 1/2 I)^{-1}, one CTA per panel, thread j owns column j.
__global__ void __launch_bounds__(NBW) wy_tinv_kernel(const double *__restrict__ G, double *__restrict__ T)
{
    extern __shared__ double Rsm[];                 // R[i][k] at Rsm[i*(NBW+1)+k]
    const double *g = G + (size_t)blockIdx.x * NBW * NBW;
    double *tcol = T + (size_t)blockIdx.x * NBW * NBW + (size_t)threadIdx.x * NBW;   // column j of T
    const int j = threadIdx.x;
    for (int i = 0; i < NBW; ++i) Rsm[i * (NBW + 1) + j] = (i < j) ? g[i + j * NBW] : (i == j ? 0.5 : 0.0);
    __syncthreads();
    // column j of the inverse of an upper-triangular matrix, bottom-up; the column is private
    // to this thread, so it can live in (L1-cached) global memory
    for (int i = j + 1; i < NBW; ++i) tcol[i] = 0.0;
    tcol[j] = 2.0;
    for (int i = j - 1; i >= 0; --i) {
        double s = 0.0;
        const double *Ri = Rsm + i * (NBW + 1);
        for (int k = i + 1; k <= j; ++k) s += Ri[k] * tcol[k];
        tcol[i] = -2.0 * s;
    }
}

size_t backtransform_workspace_bytes(int rows, int nref, int nc)
{
    long ld = round_up(rows, 2);
    long npad = round_up(nref > 0 ? nref : 1, NBW);
    long np = npad / NBW;
    size_t d = 0;
    d += 2 * (size_t)ld * npad;                 // V, VT
    d += 2 * (size_t)np * NBW * NBW;            // G, T
    d += (size_t)NBW * nc;                      // W
    d += (size_t)WY_MAX_SPLIT * NBW * nc;       // split-K partials
    return d * sizeof(double) + 4096;
}

// Apply Q = H_0 H_1 ... H_{nref-1} to C (rows x nc, ldc) in place.  Reflector j is
// column j of the (rows x nref) trapezoid described by `left`:
//   left = 1: v_j = A[j:rows, j]        (column reflectors, rows = m)
//   left = 0: v_j = A[j, j+1:rows]^T    (row reflectors,    rows = n)
void wy_apply_device(int left, int rows, int nref, const double *A, long lda, double *C, long ldc, int nc,
                     void *workspace, cudaStream_t st)
{
    if (nref <= 0 || nc <= 0) return;
    const long ld = round_up(rows, 2);
    const int npad = (int)round_up(nref, NBW), np = npad / NBW;
    const int ro = left ? 0 : 1;                // first nonzero row of reflector j is j + ro
    double *w = (double *)workspace;
    double *V = w;   w += (size_t)ld * npad;
    double *VT = w;  w += (size_t)ld * npad;
    double *G = w;   w += (size_t)np * NBW * NBW;
    double *T = w;   w += (size_t)np * NBW * NBW;
    double *W = w;   w += (size_t)NBW * nc;
    double *Wp = w;

    if (left) {
        extract_left_kernel<<<ceil_div(ld * npad, 256), 256, 0, st>>>(rows, nref, npad, A, lda, V, ld);
    } else {
        dim3 grid(ceil_div(ld, 32), ceil_div(npad, 32));
        extract_right_kernel<<<grid, 256, 0, st>>>(rows, nref, npad, A, lda, V, ld);
    }
    SVD_KERNEL_CHECK();

    // Gram matrices of all panels in one batched launch: G_p = V_p^T V_p, K = rows - p0 - ro
    {
        GemmArgs g = {};
        g.M = NBW; g.N = NBW; g.K = rows - ro;
        g.A = V + ro; g.lda = ld; g.transA = 1;
        g.B = V + ro; g.ldb = ld; g.transB = 0;
        g.C = G; g.ldc = NBW; g.alpha = 1.0; g.beta = 0.0;
        g.batch = np; g.sA = (long)NBW * (ld + 1); g.sB = g.sA; g.sC = (long)NBW * NBW; g.dK = NBW;
        g.splitk = 1;
        dgemm_dmma(g, st);
    }
    {
        const int smem = NBW * (NBW + 1) * (int)sizeof(double);
        SVD_CUDA_CHECK(cudaFuncSetAttribute(wy_tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        wy_tinv_kernel<<<np, NBW, smem, st>>>(G, T);
    }
    SVD_KERNEL_CHECK();
    // VT_p = V_p T_p for all panels: (rows - p0 - ro) x NBW
    {
        SVD_CUDA_CHECK(cudaMemsetAsync(VT, 0, sizeof(double) * (size_t)ld * npad, st));
        GemmArgs g = {};
        g.M = rows - ro; g.N = NBW; g.K = NBW;
        g.A = V + ro; g.lda = ld; g.transA = 0;
        g.B = T; g.ldb = NBW; g.transB = 0;
        g.C = VT + ro; g.ldc = ld; g.alpha = 1.0; g.beta = 0.0;
        g.batch = np; g.sA = (long)NBW * (ld + 1); g.sB = (long)NBW * NBW; g.sC = g.sA; g.dM = NBW;
        g.splitk = 1;
        dgemm_dmma(g, st);
    }
    int dev = 0, nsm = 148;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    // panels last to first:  W = V_p^T C[p0+ro:, :] ;  C[p0+ro:, :] -= VT_p W
    for (int p = np - 1; p >= 0; --p) {
        const int r0 = p * NBW + ro;
        const int K = rows - r0;
        if (K <= 0) continue;
        const double *Vp = V + r0 + (long)p * NBW * ld;
        const double *VTp = VT + r0 + (long)p * NBW * ld;
        int tiles = ceil_div(nc, 64) * ceil_div(NBW, 128);
        int split = 1;
        if (tiles < 2 * nsm) {
            split = (2 * nsm) / tiles;
            int maxs = K / 256;                 // at least 256 of K per slice
            if (split > maxs) split = maxs;
            if (split > WY_MAX_SPLIT) split = WY_MAX_SPLIT;
            if (split < 1) split = 1;
        }
        GemmArgs g1 = {};
        g1.M = NBW; g1.N = nc; g1.K = K;
        g1.A = Vp; g1.lda = ld; g1.transA = 1;
        g1.B = C + r0; g1.ldb = ldc; g1.transB = 0;
        g1.alpha = 1.0; g1.beta = 0.0; g1.batch = 1;
        if (split > 1) {
            g1.C = Wp; g1.ldc = NBW; g1.splitk = split; g1.sSplit = (long)NBW * nc;
            dgemm_dmma(g1, st);
            sum_partials(W, NBW, Wp, NBW, (long)NBW * nc, split, NBW, nc, 1.0, 0.0, st);
        } else {
            g1.C = W; g1.ldc = NBW; g1.splitk = 1;
            dgemm_dmma(g1, st);
        }
        GemmArgs g2 = {};
        g2.M = K; g2.N = nc; g2.K = NBW;
        g2.A = VTp; g2.lda = ld; g2.transA = 0;
        g2.B = W; g2.ldb = NBW; g2.transB = 0;
        g2.C = C + r0; g2.ldc = ldc; g2.alpha = -1.0; g2.beta = 1.0; g2.batch = 1; g2.splitk = 1;
        dgemm_dmma(g2, st);
    }
}

} // namespace svdgpu
