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
#include "bidiag.cuh"
#include "ozaki.cuh"
#include <cstdlib>

namespace svdgpu {

constexpr int NBW = 128;    // WY panel width (K of the update GEMM: 128 keeps it DMMA-bound, not C-traffic-bound)

// VL[r, jj] = A[r, j0 + jj] for r >= j0 + jj (left reflector j lives in column j from the diagonal down);
// columns [j0, j0 + ncols) of the gathered array, VL already offset to column j0
__global__ void extract_left_kernel(int m, int nL, int j0, int ncols, const double *__restrict__ A, long lda,
                                    double *__restrict__ VL, long ldv)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)ldv * ncols) return;
    int r = (int)(idx % ldv), j = j0 + (int)(idx / ldv);
    double v = 0.0;
    if (j < nL && r < m && r >= j) v = A[r + (long)j * lda];
    VL[idx] = v;
}

// VR[c, jj] = A[j0 + jj, c] for c >= j0 + jj + 1 (right reflector j lives in row j right of the super-diagonal)
__global__ void __launch_bounds__(256)
extract_right_kernel(int n, int nR, int jbase, int ncols, const double *__restrict__ A, long lda,
                     double *__restrict__ VR, long ldv)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, j0 = jbase + blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {          // read A[j0+tx, c0+q] : rows contiguous over tx
        int j = j0 + tx, c = c0 + q;
        double v = 0.0;
        if (j < nR && c < n && c >= j + 1) v = A[j + (long)c * lda];
        tile[q][tx] = v;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {          // write VR[c0+tx, j0+q-jbase]
        int c = c0 + tx, j = j0 + q;
        if (c < ldv && j - jbase < ncols) VR[c + (long)(j - jbase) * ldv] = (c < n) ? tile[tx][q] : 0.0;
    }
}

// T = (striu(G) + 1/2 I)^{-1}, one CTA per panel.  Column j of the inverse of an upper-triangular matrix by
// back-substitution, t_j = 2, t_i = -2 sum_{k=i+1..j} R[i][k] t_k: a WARP per column (lanes split the dot
// product, the column lives in the lanes' registers: lane l holds t_l, t_{l+32}, t_{l+64}, t_{l+96}), 32 warps
// per CTA take the 128 columns four at a time.  (Round 1 ran one THREAD per column: 171 us per launch.)
constexpr int TINV_WARPS = 32;
__global__ void __launch_bounds__(32 * TINV_WARPS) wy_tinv_kernel(const double *__restrict__ G, double *__restrict__ T)
{
    extern __shared__ double Rsm[];                 // R[i][k] at Rsm[i*(NBW+1)+k]
    const double *g = G + (size_t)blockIdx.x * NBW * NBW;
    double *Tp = T + (size_t)blockIdx.x * NBW * NBW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < NBW * NBW; e += 32 * TINV_WARPS) {
        const int i = e % NBW, k = e / NBW;         // g is column-major: coalesced reads
        Rsm[i * (NBW + 1) + k] = (i < k) ? g[e] : (i == k ? 0.5 : 0.0);
    }
    __syncthreads();
    static_assert(NBW == 128, "four entries of the column per lane");
    for (int j = warp; j < NBW; j += TINV_WARPS) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};        // t[c] = entry lane + 32 c of column j
#pragma unroll
        for (int c = 0; c < 4; ++c) if (lane + 32 * c == j) t[c] = 2.0;
        for (int i = j - 1; i >= 0; --i) {
            const double *Ri = Rsm + i * (NBW + 1);
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = lane + 32 * c;
                if (k > i && k <= j) s = fma(Ri[k], t[c], s);
            }
            s = warp_sum(s);
#pragma unroll
            for (int c = 0; c < 4; ++c) if (lane + 32 * c == i) t[c] = -2.0 * s;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) Tp[(size_t)j * NBW + lane + 32 * c] = t[c];
    }
}

// K split of the long-K products W = V^T C (few output tiles, K = the reflector length).
// One-tile-per-CTA kernel: fill the 2 x nsm CTA slots (3x / 4x / 6x the SM count measured the same).
// Persistent kernel (dgemm_ws.cu): its CTAs walk tiles x slices round-robin, so the smallest split that
// leaves the last round at least 95 % full (else the fullest one); at least 256 of K per slice.
static int pick_split(int tiles, int K, int nsm)
{
    int maxs = K / 256;
    if (maxs > WY_MAX_SPLIT) maxs = WY_MAX_SPLIT;
    if (maxs < 1) maxs = 1;
    if (!dgemm_ws_enabled()) {
        int split = 1;
        if (tiles < 2 * nsm) split = (2 * nsm) / tiles;
        if (split > maxs) split = maxs;
        return split < 1 ? 1 : split;
    }
    int best = 1;
    double best_eff = 0.0;
    for (int sp = 1; sp <= maxs; ++sp) {
        const long units = (long)tiles * sp;
        const double eff = (double)units / (double)((units + nsm - 1) / nsm * nsm);
        if (eff >= 0.95) return sp;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
    }
    return best;
}

// ---- storage of the prepared panels of one reflector set: V | VT | G | T -----------------------
struct WyPanels { double *V, *VT, *G, *T; long ld; int npad, np; };
static WyPanels wy_carve(void *panels, int rows, int nref)
{
    WyPanels w;
    w.ld = round_up(rows, 2);
    w.npad = (int)round_up(nref > 0 ? nref : 1, NBW);
    w.np = w.npad / NBW;
    double *p = (double *)panels;
    w.V = p;   p += (size_t)w.ld * w.npad;
    w.VT = p;  p += (size_t)w.ld * w.npad;
    w.G = p;   p += (size_t)w.np * NBW * NBW;
    w.T = p;
    return w;
}
int wy_panel_count(int nref) { return (int)(round_up(nref > 0 ? nref : 1, NBW) / NBW); }
int wy_panel_width() { return NBW; }
size_t wy_panels_bytes(int rows, int nref)
{
    const long ld = round_up(rows, 2), npad = round_up(nref > 0 ? nref : 1, NBW), np = npad / NBW;
    return (2 * (size_t)ld * npad + 2 * (size_t)np * NBW * NBW) * sizeof(double) + 256;
}
// the part of the panel storage other devices need for wy_apply_prepared: columns [pb*NBW, pe*NBW) of V and of VT
void wy_panel_slices(void *panels, int rows, int nref, int pb, int pe, double **V, double **VT, size_t *count)
{
    const WyPanels w = wy_carve(panels, rows, nref);
    *V = w.V + (size_t)pb * NBW * w.ld;
    *VT = w.VT + (size_t)pb * NBW * w.ld;
    *count = (size_t)(pe - pb) * NBW * w.ld;
}
// The short-K update C -= (V T) W can run on the 5th-generation tensor cores (ozaki.cu: int8 slice products in
// TMEM, FP64-accurate) instead of the FP64 DMMA pipe: SVD_GPU_OZAKI = 0 never, 1 whenever the shape is large
// enough to fill the machine, unset = the measured default below.
constexpr int OZAKI_DEFAULT = 0;
constexpr int OZAKI_MIN_ROWS = 1024, OZAKI_MIN_COLS = 1024;
static int ozaki_mode()
{
    const char *e = getenv("SVD_GPU_OZAKI");
    return e ? atoi(e) : OZAKI_DEFAULT;
}
// rows of the tallest update a (rows x nref) reflector set can ask for is not known here: size for 32768
constexpr int OZAKI_MAX_ROWS = 65536;
size_t wy_apply_workspace_bytes(int nc)
{
    return ((size_t)NBW * nc + (size_t)WY_MAX_SPLIT * NBW * nc) * sizeof(double) + 256 +
           ozaki_workspace_bytes(OZAKI_MAX_ROWS, nc);
}
size_t backtransform_workspace_bytes(int rows, int nref, int nc)
{
    return wy_panels_bytes(rows, nref) + wy_apply_workspace_bytes(nc) + 4096;
}

// Prepare panels [pb, pe) of a reflector set: gather the reflectors into column form (V), Gram matrices,
// T = (striu(V^T V) + I/2)^-1 and VT = V T.  Reflector j is column j of the (rows x nref) trapezoid described
// by `left`:
//   left = 1: v_j = A[j:rows, j]        (column reflectors, rows = m)
//   left = 0: v_j = A[j, j+1:rows]^T    (row reflectors,    rows = n)
// Only reflectors [pb*NBW, pe*NBW) of A have to be final, so the panels can be prepared while the
// factorization that produces the later ones is still running (and shipped to other devices).
void wy_setup_device(int left, int rows, int nref, const double *A, long lda, void *panels, int pb, int pe,
                     cudaStream_t st)
{
    if (nref <= 0) return;
    const WyPanels w = wy_carve(panels, rows, nref);
    if (pe > w.np) pe = w.np;
    if (pb >= pe) return;
    const long ld = w.ld;
    const int ro = left ? 0 : 1;                // first nonzero row of reflector j is j + ro
    const int j0 = pb * NBW, ncols = (pe - pb) * NBW, nbatch = pe - pb;
    double *Vp = w.V + (size_t)j0 * ld;
    if (left) {
        extract_left_kernel<<<ceil_div(ld * ncols, 256), 256, 0, st>>>(rows, nref, j0, ncols, A, lda, Vp, ld);
    } else {
        dim3 grid(ceil_div(ld, 32), ceil_div(ncols, 32));
        extract_right_kernel<<<grid, 256, 0, st>>>(rows, nref, j0, ncols, A, lda, Vp, ld);
    }
    SVD_KERNEL_CHECK();
    // Gram matrices of the panels in one batched launch: G_p = V_p^T V_p, K = rows - p0 - ro.  The long K is
    // split (partials in the panels' own VT columns, which are written last) so that the launch is wide and
    // short: it may run on a low-priority stream next to the factorization and must not hold SMs for long.
    const long sP = (long)NBW * (ld + 1);       // panel p starts NBW rows and NBW columns further
    {
        const int Kmax = rows - ro - j0;
        int gs = Kmax / 512;
        if (gs > 16) gs = 16;
        if (gs > (int)(ld / NBW)) gs = (int)(ld / NBW);
        if (gs < 1) gs = 1;
        GemmArgs g = {};
        g.M = NBW; g.N = NBW; g.K = Kmax;
        g.A = w.V + ro + pb * sP; g.lda = ld; g.transA = 1;
        g.B = g.A; g.ldb = ld; g.transB = 0;
        g.alpha = 1.0; g.beta = 0.0;
        g.batch = nbatch; g.sA = sP; g.sB = sP; g.sC = (long)NBW * NBW; g.dK = NBW;
        double *Gp = w.G + (size_t)pb * NBW * NBW;
        if (gs > 1) {
            double *part = w.VT + (size_t)j0 * ld;              // >= gs * nbatch * NBW^2 doubles
            g.C = part; g.ldc = NBW; g.splitk = gs; g.sSplit = (long)nbatch * NBW * NBW;
            dgemm_dmma(g, st);
            sum_partials(Gp, NBW, part, NBW, g.sSplit, gs, NBW, NBW * nbatch, 1.0, 0.0, st);
        } else {
            g.C = Gp; g.ldc = NBW; g.splitk = 1;
            dgemm_dmma(g, st);
        }
    }
    {
        const int smem = NBW * (NBW + 1) * (int)sizeof(double);
        static DeviceOnce once;
        if (first_on_device(once))
            SVD_CUDA_CHECK(cudaFuncSetAttribute(wy_tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        wy_tinv_kernel<<<nbatch, 32 * TINV_WARPS, smem, st>>>(w.G + (size_t)pb * NBW * NBW, w.T + (size_t)pb * NBW * NBW);
    }
    SVD_KERNEL_CHECK();
    // VT_p = V_p T_p: (rows - p0 - ro) x NBW
    {
        SVD_CUDA_CHECK(cudaMemsetAsync(w.VT + (size_t)j0 * ld, 0, sizeof(double) * (size_t)ld * ncols, st));
        GemmArgs g = {};
        g.M = rows - ro - j0; g.N = NBW; g.K = NBW;
        g.A = w.V + ro + pb * sP; g.lda = ld; g.transA = 0;
        g.B = w.T + (size_t)pb * NBW * NBW; g.ldb = NBW; g.transB = 0;
        g.C = w.VT + ro + pb * sP; g.ldc = ld; g.alpha = 1.0; g.beta = 0.0;
        g.batch = nbatch; g.sA = sP; g.sB = (long)NBW * NBW; g.sC = sP; g.dM = NBW;
        g.splitk = 1;
        dgemm_dmma(g, st);
    }
}

// C (rows x nc, ldc) <- H_0 H_1 ... H_{nref-1} C from prepared panels (only V and VT are read):
// panels last to first,  W = V_p^T C[p0+ro:, :] ;  C[p0+ro:, :] -= VT_p W
void wy_apply_prepared(int left, int rows, int nref, const void *panels, double *C, long ldc, int nc,
                       void *workspace, cudaStream_t st)
{
    if (nref <= 0 || nc <= 0) return;
    const WyPanels w = wy_carve(const_cast<void *>(panels), rows, nref);
    const long ld = w.ld;
    const int ro = left ? 0 : 1;
    double *W = (double *)workspace;
    double *Wp = W + (size_t)NBW * nc;
    void *oz_ws = (void *)(Wp + (size_t)WY_MAX_SPLIT * NBW * nc + 32);
    const int oz = ozaki_mode();
    static int nsm = 0;
    if (nsm == 0) {
        int dev = 0;
        SVD_CUDA_CHECK(cudaGetDevice(&dev));
        SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    }
    for (int p = w.np - 1; p >= 0; --p) {
        const int r0 = p * NBW + ro;
        const int K = rows - r0;
        if (K <= 0) continue;
        const double *Vp = w.V + r0 + (long)p * NBW * ld;
        const double *VTp = w.VT + r0 + (long)p * NBW * ld;
        const int split = pick_split(ceil_div(nc, 64) * ceil_div(NBW, 128), K, nsm);
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
        if (oz && K >= OZAKI_MIN_ROWS && K <= OZAKI_MAX_ROWS && nc >= OZAKI_MIN_COLS && ozaki_update_supported(K, nc, NBW)) {
            ozaki_update_device(K, nc, -1.0, VTp, ld, W, NBW, C + r0, ldc, oz_ws, st);
            continue;
        }
        GemmArgs g2 = {};
        g2.M = K; g2.N = nc; g2.K = NBW;
        g2.A = VTp; g2.lda = ld; g2.transA = 0;
        g2.B = W; g2.ldb = NBW; g2.transB = 0;
        g2.C = C + r0; g2.ldc = ldc; g2.alpha = -1.0; g2.beta = 1.0; g2.batch = 1; g2.splitk = 1;
        dgemm_dmma(g2, st);
    }
}

// set-up + apply in one call (phase-level entry points, QR-first U = Q [U_R; 0])
void wy_apply_device(int left, int rows, int nref, const double *A, long lda, double *C, long ldc, int nc,
                     void *workspace, cudaStream_t st)
{
    if (nref <= 0 || nc <= 0) return;
    char *apply_ws = (char *)workspace + wy_panels_bytes(rows, nref);
    wy_setup_device(left, rows, nref, A, lda, workspace, 0, wy_panel_count(nref), st);
    wy_apply_prepared(left, rows, nref, workspace, C, ldc, nc, apply_ws, st);
}


// =============================================================================================
// QR first: for m >> n, A = Q R with Householder reflectors in the library's
// own convention (unit-norm v, H = I - 2 v v^T, stored from the diagonal down), so that the SVD
// only has to bidiagonalize the n x n factor R and U = Q [U_R; 0] goes through wy_apply_device.
// Panels of NBW columns: BLAS2 factorization inside the (L2-resident) panel, then one compact-WY
// update of the trailing columns on the DMMA GEMM:  A2 <- (I - V T^T V^T) A2.
// =============================================================================================
constexpr int QR_GRAM_SPLIT = 64;
constexpr int QR_CHUNK = 512;       // rows per CTA in the column-dot kernel

// partial column dots of panel column i with the later panel columns, and its norm:
//   part[chunk][jj] = sum_{r in chunk, r >= i} A[r, i] * A[r, i+1+jj]   (jj < nj),  part[chunk][NBW] = sum A[r,i]^2
__global__ void __launch_bounds__(256)
qr_coldots_kernel(const double *__restrict__ A, long lda, int m, int i, int nj, double *__restrict__ part,
                  double *__restrict__ rowi)
{
    // row i of the panel is rewritten by one CTA of the apply kernel while the others still need it
    if (blockIdx.x == 0 && threadIdx.x <= nj)
        rowi[threadIdx.x < nj ? threadIdx.x : NBW] = A[i + (long)(threadIdx.x < nj ? i + 1 + threadIdx.x : i) * lda];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = i + blockIdx.x * QR_CHUNK, r1 = min(m, r0 + QR_CHUNK);
    const double *ci = A + (long)i * lda;
    double *out = part + (long)blockIdx.x * (NBW + 1);
    for (int jj = warp; jj <= nj; jj += 8) {                 // jj == nj: the norm
        const double *cj = (jj < nj) ? A + (long)(i + 1 + jj) * lda : ci;
        double acc = 0.0;
        for (int r = r0 + lane; r < r1; r += 32) acc += ci[r] * cj[r];
        acc = warp_sum(acc);
        if (lane == 0) out[jj < nj ? jj : NBW] = acc;
    }
}

// One launch per panel column i: finish the reduction of the dots the previous launch left behind,
// form reflector i, apply it to the later panel columns (rows >= i), and — on the freshly updated
// values still in registers — leave the dots of column i+1 with the columns after it for the next
// launch.   v = (c + s*nu*e_i)*inv in place, alpha[i] = -s*nu, A[r,j] -= v_r y_j, y_j = 2 (t_j + s*nu*A[i,j]) inv.
// part/rowi are double-buffered by the caller (other CTAs of this launch still read the old ones).
template <int RPT>
__global__ void __launch_bounds__(256)
qr_step_kernel(double *__restrict__ A, long lda, int m, int i, int nj, const double *__restrict__ part, int nchunk,
               const double *__restrict__ rowi, double *__restrict__ part_out, double *__restrict__ rowi_out,
               double *__restrict__ alpha)
{
    __shared__ double s_y[NBW];
    __shared__ double s_red[8][NBW];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t <= nj) {
        double a2 = 0.0;
        const int slot = (t < nj) ? t : NBW;
        for (int c = 0; c < nchunk; ++c) a2 += part[(long)c * (NBW + 1) + slot];
        s_y[t < nj ? t : NBW - 1] = a2;                      // raw dots; the norm parks in the last slot
    }
    __syncthreads();
    const double ci = rowi[NBW];
    const double nrm2 = s_y[NBW - 1];
    const double nu = sqrt(nrm2), sg = (ci < 0.0) ? -1.0 : 1.0, snu = sg * nu;
    const double sc = sqrt(2.0) * sqrt(nu * nu + fabs(nu * ci));
    const double inv = (sc > 0.0) ? 1.0 / sc : 0.0;
    __syncthreads();
    if (t < nj) s_y[t] = 2.0 * (s_y[t] + snu * rowi[t]) * inv;
    __syncthreads();

    const int rbase = i + blockIdx.x * (256 * RPT) + t;
    double v[RPT], n1[RPT];
    double *col[RPT];
    bool live[RPT], nxt[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = rbase + 256 * k;
        live[k] = r < m;
        nxt[k] = live[k] && r > i;                           // rows of the next reflector
        col[k] = A + (live[k] ? r : i) + (long)i * lda;
        v[k] = live[k] ? (*col[k] + (r == i ? snu : 0.0)) * inv : 0.0;
        n1[k] = 0.0;
    }
    double *po = part_out + (long)blockIdx.x * (NBW + 1);
    constexpr int JB = 8;                                    // columns loaded ahead of the dependent stores
    if (nj > 0) {                                            // column i+1: the next reflector's column
        const double y = s_y[0];
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            double *pa = col[k] + lda;
            double a = 0.0;
            if (live[k]) { a = *pa - v[k] * y; *pa = a; }
            n1[k] = nxt[k] ? a : 0.0;
            acc += n1[k] * a;
            if (rbase + 256 * k == i + 1) rowi_out[NBW] = a;
        }
        acc = warp_sum(acc);
        if (lane == 0) s_red[warp][0] = acc;
    }
    for (int j0 = 1; j0 < nj; j0 += JB) {
        double a[RPT][JB];
#pragma unroll
        for (int k = 0; k < RPT; ++k)
#pragma unroll
            for (int q = 0; q < JB; ++q)
                a[k][q] = (live[k] && j0 + q < nj) ? col[k][(long)(j0 + q + 1) * lda] : 0.0;
#pragma unroll
        for (int q = 0; q < JB; ++q) {
            if (j0 + q < nj) {
                const double y = s_y[j0 + q];
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const double an = a[k][q] - v[k] * y;
                    if (live[k]) col[k][(long)(j0 + q + 1) * lda] = an;
                    acc += n1[k] * an;
                    if (rbase + 256 * k == i + 1) rowi_out[j0 + q - 1] = an;      // row i+1 for the next launch
                }
                acc = warp_sum(acc);
                if (lane == 0) s_red[warp][j0 + q] = acc;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k)
        if (live[k]) *col[k] = v[k];
    __syncthreads();
    if (t < nj) {
        double a2 = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) a2 += s_red[w8][t];
        po[t == 0 ? NBW : t - 1] = a2;                       // jj = 0 is the next column's own norm
    }
    if (blockIdx.x == 0 && t == 0) alpha[i] = -snu;
}

// clean copy of panel p0: Vp[r - p0, jj] = A[r, p0 + jj] for r >= p0 + jj, else 0   ((m - p0) x NBW)
__global__ void qr_extract_panel_kernel(const double *__restrict__ A, long lda, int m, int n, int p0,
                                        double *__restrict__ Vp, long ldv)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long rows = m - p0;
    if (idx >= rows * NBW) return;
    const int rr = (int)(idx % rows), jj = (int)(idx / rows);
    double v = 0.0;
    if (p0 + jj < n && rr >= jj) v = A[(p0 + rr) + (long)(p0 + jj) * lda];
    Vp[rr + (long)jj * ldv] = v;
}

// R (n x n, ldr): strict upper triangle from A, diagonal from alpha, zero below
__global__ void qr_copy_r_kernel(const double *__restrict__ A, long lda, const double *__restrict__ alpha, int n,
                                 double *__restrict__ R, long ldr)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)ldr * n) return;
    const int r = (int)(idx % ldr), c = (int)(idx / ldr);
    double v = 0.0;
    if (r < n) v = (r < c) ? A[r + (long)c * lda] : (r == c ? alpha[c] : 0.0);
    R[r + (long)c * ldr] = v;
}

size_t qr_workspace_bytes(int m, int n)
{
    const long ldv = round_up(m, 2);
    size_t d = 0;
    d += 2 * (size_t)ldv * NBW;                          // Vp, Vp T^T
    d += 2 * (size_t)NBW * NBW;                          // G, T
    d += (size_t)NBW * n;                                // W
    d += (size_t)WY_MAX_SPLIT * NBW * n;                 // split-K partials of W
    d += 2 * (size_t)(ceil_div(m, 256) + 1) * (NBW + 1); // column-dot partials + stashed row, double-buffered
    d += (size_t)QR_GRAM_SPLIT * NBW * NBW;              // split-K partials of the Gram matrix
    d += (size_t)n + 8;                                  // alpha
    return d * sizeof(double) + 4096;
}

// A (m x n, m >= n) <- reflectors (from the diagonal down) and, above the diagonal, R's strict upper
// triangle; R (n x n, ldr) receives the full triangular factor.
void qr_device(int m, int n, double *A, long lda, double *R, long ldr, void *workspace, cudaStream_t st,
               const ProgressHook *hook)
{
    const long ldv = round_up(m, 2);
    double *w = (double *)workspace;
    double *Vp = w;    w += (size_t)ldv * NBW;
    double *VTt = w;   w += (size_t)ldv * NBW;
    double *G = w;     w += (size_t)NBW * NBW;
    double *T = w;     w += (size_t)NBW * NBW;
    double *W = w;     w += (size_t)NBW * n;
    double *Wp = w;    w += (size_t)WY_MAX_SPLIT * NBW * n;
    const size_t pstride = (size_t)(ceil_div(m, 256) + 1) * (NBW + 1);
    double *partb[2], *rowib[2];
    partb[0] = w; rowib[0] = w + pstride - (NBW + 1); w += pstride;
    partb[1] = w; rowib[1] = w + pstride - (NBW + 1); w += pstride;
    double *Gp = w;    w += (size_t)QR_GRAM_SPLIT * NBW * NBW;
    double *alpha = w;
    int dev = 0, nsm = 148;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int tsmem = NBW * (NBW + 1) * (int)sizeof(double);
    SVD_CUDA_CHECK(cudaFuncSetAttribute(wy_tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tsmem));

    for (int p0 = 0; p0 < n; p0 += NBW) {
        const int pw = (n - p0 < NBW) ? n - p0 : NBW;
        // ---- panel factorization (BLAS2, panel resident in L2): one launch per column
        {
            const int nj0 = pw - 1, L0 = m - p0;
            qr_coldots_kernel<<<ceil_div(L0, QR_CHUNK), 256, 0, st>>>(A, lda, m, p0, nj0, partb[0], rowib[0]);
            SVD_KERNEL_CHECK();
            int nchunk = ceil_div(L0, QR_CHUNK), cur = 0;
            for (int i = p0; i < p0 + pw; ++i) {
                const int nj = p0 + pw - i - 1, L = m - i;
                if (L >= 256 * 2 * nsm / 2) {
                    const int nb = ceil_div(L, 512);
                    qr_step_kernel<2><<<nb, 256, 0, st>>>(A, lda, m, i, nj, partb[cur], nchunk, rowib[cur],
                                                          partb[cur ^ 1], rowib[cur ^ 1], alpha);
                    nchunk = nb;
                } else {
                    const int nb = ceil_div(L, 256);
                    qr_step_kernel<1><<<nb, 256, 0, st>>>(A, lda, m, i, nj, partb[cur], nchunk, rowib[cur],
                                                          partb[cur ^ 1], rowib[cur ^ 1], alpha);
                    nchunk = nb;
                }
                SVD_KERNEL_CHECK();
                cur ^= 1;
            }
        }
        const int ntrail = n - p0 - pw;
        // the panel's reflectors are final (the trailing update below only touches later columns)
        if (hook && hook->fn) hook->fn(hook->user, p0 + pw, st);
        if (ntrail <= 0) break;
        // ---- trailing update  A2 <- A2 - (V T^T) (V^T A2),  T = (striu(V^T V) + I/2)^-1
        const int rows = m - p0;
        qr_extract_panel_kernel<<<ceil_div((long)rows * NBW, 256), 256, 0, st>>>(A, lda, m, n, p0, Vp, ldv);
        SVD_KERNEL_CHECK();
        GemmArgs g = {};
        g.M = NBW; g.N = NBW; g.K = rows; g.A = Vp; g.lda = ldv; g.transA = 1; g.B = Vp; g.ldb = ldv; g.transB = 0;
        g.alpha = 1.0; g.beta = 0.0; g.batch = 1;
        int gs = rows / 512;
        if (gs > QR_GRAM_SPLIT) gs = QR_GRAM_SPLIT;
        if (gs > 1) {
            g.C = Gp; g.ldc = NBW; g.splitk = gs; g.sSplit = (long)NBW * NBW;
            dgemm_dmma(g, st);
            sum_partials(G, NBW, Gp, NBW, (long)NBW * NBW, gs, NBW, NBW, 1.0, 0.0, st);
        } else {
            g.C = G; g.ldc = NBW; g.splitk = 1;
            dgemm_dmma(g, st);
        }
        wy_tinv_kernel<<<1, 32 * TINV_WARPS, tsmem, st>>>(G, T);
        SVD_KERNEL_CHECK();
        GemmArgs g2 = {};
        g2.M = rows; g2.N = NBW; g2.K = NBW; g2.A = Vp; g2.lda = ldv; g2.transA = 0; g2.B = T; g2.ldb = NBW; g2.transB = 1;
        g2.C = VTt; g2.ldc = ldv; g2.alpha = 1.0; g2.beta = 0.0; g2.batch = 1; g2.splitk = 1;
        dgemm_dmma(g2, st);
        double *A2 = A + p0 + (long)(p0 + pw) * lda;
        const int split = pick_split(ceil_div(ntrail, 64), rows, nsm);
        GemmArgs g3 = {};
        g3.M = NBW; g3.N = ntrail; g3.K = rows; g3.A = Vp; g3.lda = ldv; g3.transA = 1; g3.B = A2; g3.ldb = lda; g3.transB = 0;
        g3.alpha = 1.0; g3.beta = 0.0; g3.batch = 1;
        if (split > 1) {
            g3.C = Wp; g3.ldc = NBW; g3.splitk = split; g3.sSplit = (long)NBW * ntrail;
            dgemm_dmma(g3, st);
            sum_partials(W, NBW, Wp, NBW, (long)NBW * ntrail, split, NBW, ntrail, 1.0, 0.0, st);
        } else {
            g3.C = W; g3.ldc = NBW; g3.splitk = 1;
            dgemm_dmma(g3, st);
        }
        GemmArgs g4 = {};
        g4.M = rows; g4.N = ntrail; g4.K = NBW; g4.A = VTt; g4.lda = ldv; g4.transA = 0; g4.B = W; g4.ldb = NBW; g4.transB = 0;
        g4.C = A2; g4.ldc = lda; g4.alpha = -1.0; g4.beta = 1.0; g4.batch = 1; g4.splitk = 1;
        dgemm_dmma(g4, st);
    }
    qr_copy_r_kernel<<<ceil_div(ldr * n, 256), 256, 0, st>>>(A, lda, alpha, n, R, ldr);
    SVD_KERNEL_CHECK();
}

} // namespace svdgpu
