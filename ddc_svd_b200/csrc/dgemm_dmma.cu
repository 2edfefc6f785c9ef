// dgemm_dmma.cu — FP64 GEMM on the sm_100a DMMA pipe (mma.sync.m8n8k4.f64 -> DMMA.8x8x4).
//
// Used for the two BLAS3 pieces of the svd_gpu() path:
//   * the compact-WY back-transform that replaces the reference's BLAS1 multU/multV loops
//     (bidiag_par.c:990-1095, svd_gpu.c:117-121);
//   * the deferred rank-2nb trailing update of the panel bidiagonalization
//     (the work of left_update_mat.cl / right_update_mat.cl, applied once per panel).
// tcgen05 has no f64 kind, so the only tensor path that keeps the reference's working
// precision bit-for-bit FP64 is DMMA; every operand stays FP64 end to end.
//
// Tiling: CTA tile BM x BN, BK = 16, warp tile 32 x 32 (4 x 4 DMMA fragments, 16 DMMA per
// 8 LDS.64), 3-stage cp.async pipeline.  Shared-memory tiles keep the global-contiguous
// dimension contiguous and pad the leading dimension to == 4 (mod 16) doubles, which makes
// every fragment load (8 rows x 4 k) hit 16 distinct 8-byte banks per half-warp.
#include "common.cuh"

namespace svdgpu {

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int BK = 16;
constexpr int STAGES = 3;

template <int BM, int BN, bool TA, bool TB> struct TileCfg {
    static constexpr int LDA_S = TA ? (BK + 4) : (BM + 4);
    static constexpr int LDB_S = TB ? (BN + 4) : (BK + 4);
    static constexpr int A_ELEMS = TA ? BM * LDA_S : BK * LDA_S;
    static constexpr int B_ELEMS = TB ? BK * LDB_S : BN * LDB_S;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 8;
    static constexpr int THREADS = (BM / 32) * (BN / 32) * 32;
};

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(TileCfg<BM, BN, TA, TB>::THREADS)
dgemm_dmma_kernel(GemmArgs g)
{
    using Cfg = TileCfg<BM, BN, TA, TB>;
    constexpr int NT = Cfg::THREADS;
    extern __shared__ __align__(16) double smem[];

    const int zb = blockIdx.z / g.splitk;
    const int zs = blockIdx.z - zb * g.splitk;
    const int M = g.M - zb * g.dM;
    const int Kfull = g.K - zb * g.dK;
    const int N = g.N;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M || n0 >= N) return;

    int kbeg = 0, kend = Kfull;
    if (g.splitk > 1) {
        int chunk = ((Kfull + g.splitk - 1) / g.splitk + BK - 1) / BK * BK;
        kbeg = zs * chunk;
        kend = min(Kfull, kbeg + chunk);
    }
    const double *A = g.A + zb * g.sA;
    const double *B = g.B + zb * g.sB;
    double *C = g.C + zb * g.sC + (g.splitk > 1 ? zs * g.sSplit : 0);
    const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int wm0 = (warp % (BM / 32)) * 32, wn0 = (warp / (BM / 32)) * 32;

    auto load_stage = [&](int stage, int kt) {
        double *As = smem + stage * Cfg::STAGE_ELEMS;
        double *Bs = As + Cfg::A_ELEMS;
        const int k0 = kbeg + kt * BK;
        if (!TA) {   // A is M x K, m contiguous -> As[k][m]
            for (int e = tid; e < BK * BM; e += NT) {
                int mm = e % BM, kk = e / BM;
                bool ok = (m0 + mm < M) && (k0 + kk < kend);
                const double *src = ok ? A + (m0 + mm) + (long)(k0 + kk) * g.lda : A;
                cp_async8(As + kk * Cfg::LDA_S + mm, src, ok);
            }
        } else {     // A stored K x M, k contiguous -> As[m][k]
            for (int e = tid; e < BK * BM; e += NT) {
                int kk = e % BK, mm = e / BK;
                bool ok = (m0 + mm < M) && (k0 + kk < kend);
                const double *src = ok ? A + (k0 + kk) + (long)(m0 + mm) * g.lda : A;
                cp_async8(As + mm * Cfg::LDA_S + kk, src, ok);
            }
        }
        if (!TB) {   // B is K x N, k contiguous -> Bs[n][k]
            for (int e = tid; e < BK * BN; e += NT) {
                int kk = e % BK, nn = e / BK;
                bool ok = (n0 + nn < N) && (k0 + kk < kend);
                const double *src = ok ? B + (k0 + kk) + (long)(n0 + nn) * g.ldb : B;
                cp_async8(Bs + nn * Cfg::LDB_S + kk, src, ok);
            }
        } else {     // B stored N x K, n contiguous -> Bs[k][n]
            for (int e = tid; e < BK * BN; e += NT) {
                int nn = e % BN, kk = e / BN;
                bool ok = (n0 + nn < N) && (k0 + kk < kend);
                const double *src = ok ? B + (n0 + nn) + (long)(k0 + kk) * g.ldb : B;
                cp_async8(Bs + kk * Cfg::LDB_S + nn, src, ok);
            }
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // C -= A*B style updates (alpha = +-1, beta != 0): start the accumulators from C so that the
    // read of the C tile overlaps the pipeline prologue instead of trailing the main loop
    const bool cinit = (g.splitk <= 1) && (g.beta != 0.0) && (g.alpha == 1.0 || g.alpha == -1.0);
    if (cinit) {
        const double sc = g.beta * g.alpha;          // beta/alpha for alpha = +-1
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int mm = m0 + wm0 + mi * 8 + gq;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int nn = n0 + wn0 + ni * 8 + 2 * tq + e;
                    if (mm < M && nn < N) acc[mi][ni][e] = sc * C[mm + (long)nn * g.ldc];
                }
        }
    }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const double *As = smem + (kt % STAGES) * Cfg::STAGE_ELEMS;
        const double *Bs = As + Cfg::A_ELEMS;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[4], b[4];
            const int kk = k4 * 4 + tq;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                int mm = wm0 + mi * 8 + gq;
                a[mi] = TA ? As[mm * Cfg::LDA_S + kk] : As[kk * Cfg::LDA_S + mm];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                int nn = wn0 + ni * 8 + gq;
                b[ni] = TB ? Bs[kk * Cfg::LDB_S + nn] : Bs[nn * Cfg::LDB_S + kk];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

    const double alpha = g.alpha;
    const double beta = (g.splitk > 1 || cinit) ? 0.0 : g.beta;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int mm = m0 + wm0 + mi * 8 + gq;
        if (mm >= M) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nn = n0 + wn0 + ni * 8 + 2 * tq + e;
                if (nn >= N) continue;
                double *c = C + mm + (long)nn * g.ldc;
                double v = alpha * acc[mi][ni][e];
                if (beta != 0.0) v += beta * (*c);
                *c = v;
            }
        }
    }
}

template <int BM, int BN, bool TA, bool TB>
static void launch_cfg(const GemmArgs &g, cudaStream_t st)
{
    using Cfg = TileCfg<BM, BN, TA, TB>;
    static DeviceOnce once;
    if (first_on_device(once))
        SVD_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<BM, BN, TA, TB>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES));
    dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN), g.batch * g.splitk);
    dgemm_dmma_kernel<BM, BN, TA, TB><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(g);
    SVD_KERNEL_CHECK();
}

template <bool TA, bool TB> static void launch_t(const GemmArgs &g, cudaStream_t st)
{
    // SVD_GPU_GEMM_TILE=64 (experiments): 64 x 64 tiles, 4 CTAs of 4 warps per SM instead of 2 of 8
    const char *te = getenv("SVD_GPU_GEMM_TILE");
    const bool force64 = te && atoi(te) == 64;
    if (g.M > 64 && !force64) launch_cfg<128, 64, TA, TB>(g, st);
    else launch_cfg<64, 64, TA, TB>(g, st);
}

// SVD_GPU_GEMM_WS=0/1 overrides the default choice between the one-tile-per-CTA kernel of this file
// and the persistent warp-specialised kernel of dgemm_ws.cu (read on every call: tests flip it)
int dgemm_ws_mode()
{
    const char *e = getenv("SVD_GPU_GEMM_WS");
    return e ? atoi(e) : DGEMM_WS_DEFAULT;      // 0: one-tile-per-CTA kernel, otherwise the persistent one
}
bool dgemm_ws_enabled() { return dgemm_ws_mode() != 0; }

void dgemm_dmma(const GemmArgs &gin, cudaStream_t st)
{
    GemmArgs g = gin;
    if (g.batch < 1) g.batch = 1;
    if (g.splitk < 1) g.splitk = 1;
    if (g.M <= 0 || g.N <= 0) return;
    if (dgemm_ws_enabled() && dgemm_ws_eligible(g)) { dgemm_ws(g, st); return; }
    if (g.transA) { if (g.transB) launch_t<true, true>(g, st); else launch_t<true, false>(g, st); }
    else          { if (g.transB) launch_t<false, true>(g, st); else launch_t<false, false>(g, st); }
}

__global__ void sum_partials_kernel(double *out, long ldo, const double *part, long ldp,
                                    long sSplit, int nsplit, int M, int N, double alpha, double beta)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * N) return;
    int r = (int)(idx % M), c = (int)(idx / M);
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += part[k * sSplit + r + (long)c * ldp];
    double *o = out + r + (long)c * ldo;
    *o = (beta != 0.0 ? beta * (*o) : 0.0) + alpha * s;
}

void sum_partials(double *out, long ldo, const double *part, long ldp, long sSplit, int nsplit,
                  int M, int N, double alpha, double beta, cudaStream_t st)
{
    long tot = (long)M * N;
    if (tot <= 0) return;
    sum_partials_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(out, ldo, part, ldp, sSplit, nsplit, M, N,
                                                            alpha, beta);
    SVD_KERNEL_CHECK();
}

} // namespace svdgpu
