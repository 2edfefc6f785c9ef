// ozaki.cu — FP64-accurate GEMM updates on the 5th-generation tensor cores (tcgen05, int8, TMEM).
//
// The back-transform that replaces multU / multV (bidiag_par.c:990-1095, svd_gpu.c:117-121) spends half of
// its flops in the short-K updates  C(rows x nc) -= (V T)(rows x 128) * W(128 x nc).  tcgen05 has no FP64
// kind, and the FP64 DMMA pipe (37 TFLOP/s) is the ceiling of dgemm_ws.cu.  This file computes the same
// update to FP64 accuracy from INTEGER tensor-core products (the Ozaki scheme, error-free slicing):
//     x = 2^e * sum_{i<8} s_i 2^-(6+7i),  s_i in [-64, 64]  (int8),   exact for |x| < 2^e up to 2^(e-56)
//     A B = 2^(ea+eb) * sum_g 2^-(12+7g) * sum_{i+j=g} A_i B_j,       A_i B_j exact in int32
// The 36 slice products with i + j <= 7 go to 8 int32 accumulators in tensor memory (one per g; |sum| <
// 2^12 * K * 8 << 2^31), the epilogue recombines them in FP64 from the smallest to the largest term and
// applies the update to C.  Operands here have bounded entries (unit-norm reflector panels and vectors), so
// one power-of-two scale per operand (from its max-abs) keeps the dropped terms (g >= 8) below 2^-56
// relative to the operand scales, i.e. at the level of FP64 rounding of the same product.
//
// sm_100a mapping (one CTA per SM, persistent):
//   * the sliced operands are int8 planes in global memory, K-major (128 bytes of K per row), written by
//     oz_slice_kernel; TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) brings them into shared memory;
//   * A-stationary: K = 128 is the WHOLE contraction, so a CTA loads the 8 planes of its 128-row block once
//     (128 KB) and streams 32-column B tiles (8 planes x 4 KB) through a 2-stage ring;
//   * one elected thread issues tcgen05.mma.kind::i8 (M 128, N 32, K 32 per instruction): 36 plane pairs x 4
//     K steps per tile into the 8 accumulators of one of two TMEM buffers (2 x 8 x 32 columns = all 512);
//   * four epilogue warps read the accumulators with tcgen05.ld (32x32b), recombine, and read-modify-write
//     the FP64 tile of C while the MMA thread already works on the other TMEM buffer;
//   * mbarriers connect the roles (TMA complete_tx, tcgen05.commit), no CTA-wide barrier in the loop.
#include "common.cuh"
#include "ozaki.cuh"
#include <cuda.h>
#include <cfloat>

namespace svdgpu {

namespace {

constexpr int OZ_SL = 8;                  // slices per operand
constexpr int OZ_BM = 128, OZ_BN = 32, OZ_K = 128;
constexpr int OZ_BSTAGES = 2;
constexpr int OZ_A_BYTES = OZ_SL * OZ_BM * OZ_K;            // 128 KB
constexpr int OZ_B_BYTES = OZ_SL * OZ_BN * OZ_K;            // 32 KB per stage
constexpr int OZ_SMEM = OZ_A_BYTES + OZ_BSTAGES * OZ_B_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int OZ_THREADS = 256;           // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int OZ_NT = 16;                 // n-tiles per work unit (512 columns)

__device__ __forceinline__ unsigned oz_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_u32(bar)) : "memory");
}
// a wait that lasts seconds is a protocol bug: trap instead of hanging the device
__device__ __forceinline__ void oz_mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned a = oz_u32(bar);
    unsigned ok = 0;
    long long t0 = 0;
    for (int spin = 0;; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (spin == 64) t0 = clock64();
        if (spin > 64 && (spin & 255) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void oz_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(oz_u32(dst)), "l"(map), "r"(oz_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100): rows of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t oz_smem_desc(unsigned saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);            // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = S32, A = B = S8, both K-major, N = 32, M = 128
constexpr unsigned OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(OZ_BN >> 3) << 17) | ((unsigned)(OZ_BM >> 4) << 24);

__device__ __forceinline__ void oz_mma_i8(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(OZ_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(unsigned taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}

struct OzArgs {
    int M, N;                 // C is M x N; K = 128
    double *C; long ldc;
    long Mpad, Npad;          // rows per plane of the sliced operands
    const int *expo;          // expo[0] + expo[1] = exponent of the product scale
    double sign;              // C += sign * A B
};

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_update_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const OzArgs g)
{
    extern __shared__ unsigned char oz_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)oz_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char *sA = base;                                  // [8][128 rows][128 B]
    unsigned char *sB = base + OZ_A_BYTES;                     // [stage][8][32 rows][128 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + OZ_BSTAGES * OZ_B_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 1, *b_full = bars + 2, *b_empty = bars + 4, *t_full = bars + 6, *t_empty = bars + 8;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mblocks = (g.M + OZ_BM - 1) / OZ_BM;
    const int ntiles = (g.N + OZ_BN - 1) / OZ_BN;
    const int ngroups = (ntiles + OZ_NT - 1) / OZ_NT;
    const int nunits = mblocks * ngroups;

    if (threadIdx.x == 0) {
        oz_mbar_init(a_full, 1); oz_mbar_init(a_empty, 1);
        for (int s = 0; s < OZ_BSTAGES; ++s) { oz_mbar_init(b_full + s, 1); oz_mbar_init(b_empty + s, 1); }
        for (int b = 0; b < 2; ++b) { oz_mbar_init(t_full + b, 1); oz_mbar_init(t_empty + b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // all 512 columns of tensor memory: two buffers of 8 accumulators x 32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(oz_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            unsigned ua = 0, ub = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ua) {
                const int mb = u % mblocks, ng = u / mblocks;
                oz_mbar_wait(a_empty, (ua & 1u) ^ 1u);
                oz_mbar_expect_tx(a_full, OZ_A_BYTES);
                for (int i = 0; i < OZ_SL; ++i)
                    oz_tma_2d(sA + i * OZ_BM * OZ_K, &mapA, 0, (int)(i * g.Mpad) + mb * OZ_BM, a_full);
                const int t1 = min(ntiles, (ng + 1) * OZ_NT);
                for (int t = ng * OZ_NT; t < t1; ++t, ++ub) {
                    const unsigned s = ub % OZ_BSTAGES;
                    oz_mbar_wait(b_empty + s, ((ub / OZ_BSTAGES) & 1u) ^ 1u);
                    oz_mbar_expect_tx(b_full + s, OZ_B_BYTES);
                    for (int j = 0; j < OZ_SL; ++j)
                        oz_tma_2d(sB + s * OZ_B_BYTES + j * OZ_BN * OZ_K, &mapB, 0, (int)(j * g.Npad) + t * OZ_BN, b_full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            unsigned ua = 0, ub = 0, ut = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ua) {
                const int ng = u / mblocks;
                oz_mbar_wait(a_full, ua & 1u);
                const int t1 = min(ntiles, (ng + 1) * OZ_NT);
                for (int t = ng * OZ_NT; t < t1; ++t, ++ub, ++ut) {
                    const unsigned s = ub % OZ_BSTAGES, buf = ut & 1u;
                    oz_mbar_wait(t_empty + buf, ((ut >> 1) & 1u) ^ 1u);
                    oz_mbar_wait(b_full + s, (ub / OZ_BSTAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t a0 = oz_smem_desc(oz_u32(sA));
                    const uint64_t b0 = oz_smem_desc(oz_u32(sB + s * OZ_B_BYTES));
                    for (int gsum = 0; gsum < OZ_SL; ++gsum) {
                        const unsigned d = tmem + buf * 256 + gsum * OZ_BN;
                        for (int i = 0; i <= gsum; ++i) {
                            const int j = gsum - i;
                            // plane i of A starts i*16 KB, plane j of B j*4 KB further; 32 bytes of K per instruction
                            const uint64_t ad = a0 + (uint64_t)((i * OZ_BM * OZ_K) >> 4);
                            const uint64_t bd = b0 + (uint64_t)((j * OZ_BN * OZ_K) >> 4);
#pragma unroll
                            for (int k4 = 0; k4 < OZ_K / 32; ++k4)
                                oz_mma_i8(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), (i > 0 || k4 > 0) ? 1u : 0u);
                        }
                    }
                    oz_commit(b_empty + s);              // the B stage is free once these MMAs have read it
                    oz_commit(t_full + buf);             // ... and the accumulators are complete
                }
                oz_commit(a_empty);                      // the A block is free once every MMA of the unit is done
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ====================================
        const int ew = warp - 4;                          // TMEM lanes [32 ew, 32 ew + 32)
        const double sc0 = scalbn(g.sign, g.expo[0] + g.expo[1] - 12);
        unsigned ut = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int mb = u % mblocks, ng = u / mblocks;
            const int row = mb * OZ_BM + ew * 32 + lane;
            const int t1 = min(ntiles, (ng + 1) * OZ_NT);
            for (int t = ng * OZ_NT; t < t1; ++t, ++ut) {
                const unsigned buf = ut & 1u;
                oz_mbar_wait(t_full + buf, (ut >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tbase = tmem + ((unsigned)(ew * 32) << 16) + buf * 256;
#pragma unroll 1
                for (int c8 = 0; c8 < OZ_BN; c8 += 8) {
                    int v[OZ_SL][8];
#pragma unroll
                    for (int gsum = 0; gsum < OZ_SL; ++gsum) oz_tmem_ld8(tbase + gsum * OZ_BN + c8, v[gsum]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c8 + 8 >= OZ_BN) {
                        // every accumulator of this buffer has been read by this warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) oz_mbar_arrive(t_empty + buf);
                    }
                    const int n0 = t * OZ_BN + c8;
                    if (row < g.M) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (n0 + q < g.N) {
                                // smallest terms first; every term is an exact double
                                constexpr double wgt[OZ_SL] = {1.0, 0x1p-7, 0x1p-14, 0x1p-21, 0x1p-28, 0x1p-35, 0x1p-42, 0x1p-49};
                                double acc = 0.0;
#pragma unroll
                                for (int gsum = OZ_SL - 1; gsum >= 0; --gsum)
                                    acc = fma((double)v[gsum][q], wgt[gsum], acc);
                                double *c = g.C + row + (long)(n0 + q) * g.ldc;
                                *c = fma(acc, sc0, *c);
                            }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- slicing -----------------------------------------------------------------------------------
// max |X| over an R x K block (element (r,k) at X[r*sr + k*sk]) -> power-of-two exponent e with |x| 2^-e < 1
__global__ void __launch_bounds__(256) oz_absmax_kernel(const double *__restrict__ X, long sr, long sk, int R, int K, double *__restrict__ part)
{
    __shared__ double red[8];
    double mx = 0.0;
    const long total = (long)R * K;
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long)gridDim.x * 256) {
        // walk the contiguous direction fastest
        const long r = (sr <= sk) ? e % R : e / K, k = (sr <= sk) ? e / R : e % K;
        mx = fmax(mx, fabs(X[r * sr + k * sk]));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        part[blockIdx.x] = mx;
    }
}
__global__ void oz_expo_kernel(const double *__restrict__ part, int nparts, int *__restrict__ expo)
{
    double mx = 0.0;
    for (int p = threadIdx.x; p < nparts; p += 32) mx = fmax(mx, part[p]);
    mx = warp_max(mx);
    if (threadIdx.x == 0) *expo = (mx > 0.0 && isfinite(mx)) ? ilogb(mx) + 1 : 0;
}
// planes[i][r][k] (r < Rpad, k < 128, int8): 8 signed 7-bit digits of X[r,k] 2^-e; rows >= R are zero.
// One thread per (r, 16 consecutive k): eight 16-byte stores.
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double *__restrict__ X, long sr, long sk, int R, long Rpad, const int *__restrict__ expo,
                signed char *__restrict__ planes)
{
    const long idx = (long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= Rpad * (OZ_K / 16)) return;
    // consecutive threads take consecutive rows when rows are contiguous in X, consecutive k-chunks otherwise
    long r; int kc;
    if (sr <= sk) { r = idx % Rpad; kc = (int)(idx / Rpad); } else { kc = (int)(idx % (OZ_K / 16)); r = idx / (OZ_K / 16); }
    const int e = *expo;
    union { signed char b[16]; int4 v; } out[OZ_SL];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = kc * 16 + q;
        double x = (r < R) ? scalbn(X[r * sr + (long)k * sk], -e) : 0.0;
#pragma unroll
        for (int i = 0; i < OZ_SL; ++i) {
            const double d = rint(scalbn(x, 6 + 7 * i));
            out[i].b[q] = (signed char)(int)d;
            x -= scalbn(d, -(6 + 7 * i));
        }
    }
#pragma unroll
    for (int i = 0; i < OZ_SL; ++i)
        *reinterpret_cast<int4 *>(planes + ((size_t)i * Rpad + r) * OZ_K + kc * 16) = out[i].v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SVD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "*** libsvdgpu: cuTensorMapEncodeTiled not available\n"); abort(); }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}
// int8 planes [rows_total][128], box = 128 bytes x box_rows, SWIZZLE_128B
CUtensorMap make_map(const signed char *planes, long rows_total, int box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)OZ_K, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {(cuuint64_t)OZ_K};
    cuuint32_t box[2] = {(cuuint32_t)OZ_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)planes, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "*** libsvdgpu: cuTensorMapEncodeTiled failed (%d)\n", (int)r); abort(); }
    return m;
}

void slice_operand(const double *X, long sr, long sk, int R, long Rpad, int *expo, double *scratch, signed char *planes,
                   cudaStream_t st)
{
    const int nparts = 256;
    oz_absmax_kernel<<<nparts, 256, 0, st>>>(X, sr, sk, R, OZ_K, scratch);
    SVD_KERNEL_CHECK();
    oz_expo_kernel<<<1, 32, 0, st>>>(scratch, nparts, expo);
    SVD_KERNEL_CHECK();
    oz_slice_kernel<<<ceil_div(Rpad * (OZ_K / 16), 256), 256, 0, st>>>(X, sr, sk, R, Rpad, expo, planes);
    SVD_KERNEL_CHECK();
}

} // namespace

static long pad_to(long v, long a) { return (v + a - 1) / a * a; }

size_t ozaki_workspace_bytes(int M, int N)
{
    return (size_t)OZ_SL * (pad_to(M, OZ_BM) + pad_to(N, OZ_BN)) * OZ_K + 4096 + 512 * sizeof(double);
}

bool ozaki_update_supported(int M, int N, int K)
{
    return K == OZ_K && M >= 1 && N >= 1;
}

// C (M x N, ldc) += sign * A (M x 128, lda, column-major) * B (128 x N, ldb, column-major), FP64 in and out,
// the product formed on the int8 tensor cores (see the top of the file).
void ozaki_update_device(int M, int N, double sign, const double *A, long lda, const double *B, long ldb, double *C, long ldc,
                         void *workspace, cudaStream_t st)
{
    const long Mpad = pad_to(M, OZ_BM), Npad = pad_to(N, OZ_BN);
    char *w = (char *)workspace;
    double *scratch = (double *)w;                    w += 512 * sizeof(double);
    int *expo = (int *)w;                             w += 256;
    signed char *pa = (signed char *)(((uintptr_t)w + 1023) & ~(uintptr_t)1023);
    signed char *pb = pa + (size_t)OZ_SL * Mpad * OZ_K;
    // A[m][k] at A[m + k*lda]: rows contiguous; B[k][n] at B[k + n*ldb]: for the "row" n of the plane, k is contiguous
    slice_operand(A, 1, lda, M, Mpad, expo, scratch, pa, st);
    slice_operand(B, ldb, 1, N, Npad, expo + 1, scratch + 256, pb, st);
    const CUtensorMap mapA = make_map(pa, OZ_SL * Mpad, OZ_BM);
    const CUtensorMap mapB = make_map(pb, OZ_SL * Npad, OZ_BN);
    OzArgs g;
    g.M = M; g.N = N; g.C = C; g.ldc = ldc; g.Mpad = Mpad; g.Npad = Npad; g.expo = expo; g.sign = sign;
    int dev = 0, nsm = 148;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SVD_CUDA_CHECK(cudaFuncSetAttribute(oz_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    const int mblocks = (int)(Mpad / OZ_BM), ngroups = ceil_div(ceil_div(N, OZ_BN), OZ_NT);
    const int nunits = mblocks * ngroups;
    oz_update_kernel<<<nunits < nsm ? nunits : nsm, OZ_THREADS, OZ_SMEM, st>>>(mapA, mapB, g);
    SVD_KERNEL_CHECK();
}

} // namespace svdgpu
