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
//   * eight epilogue warps read the accumulators with tcgen05.ld (32x32b), recombine, and read-modify-write
//     the FP64 tile of C (whose values were requested two tiles earlier) while the MMA thread already works on
//     the other TMEM buffer;
//   * mbarriers connect the roles (TMA complete_tx, tcgen05.commit), no CTA-wide barrier in the loop.
#include "common.cuh"
#include "ozaki.cuh"
#include <cuda.h>
#include <cfloat>

namespace svdgpu {

namespace {

#define OZ_TR(slot, n) do { if (g.trace && blockIdx.x == 0 && (n) < 128) g.trace[(n) * 8 + (slot)] = clock64(); } while (0)
constexpr int OZ_SL = 8;                  // slices per operand
constexpr int OZ_BM = 128, OZ_BN = 32, OZ_K = 128;
constexpr int OZ_BSTAGES = 2;
constexpr int OZ_A_BYTES = OZ_SL * OZ_BM * OZ_K;            // 128 KB
constexpr int OZ_B_BYTES = OZ_SL * OZ_BN * OZ_K;            // 32 KB per stage
constexpr int OZ_SMEM = OZ_A_BYTES + OZ_BSTAGES * OZ_B_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int OZ_THREADS = 384;           // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
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
// one box of the plane stack: coordinates (k, row, plane)
__device__ __forceinline__ void oz_tma_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(oz_u32(dst)), "l"(map), "r"(oz_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
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

// D[tmem] (+)= A[smem] * B[smem]^T; the descriptors differ between the 144 instructions of a tile only in the
// start-address field (low word), so the issuing thread adds compile-time constants to two 32-bit values
template <bool ACC>
__device__ __forceinline__ void oz_mma_i8(unsigned tmem_d, unsigned a_lo, unsigned b_lo, unsigned desc_hi)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, {%6, %6, %6, %6}, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(OZ_IDESC), "n"(ACC ? 1 : 0), "r"(0u) : "memory");
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
    unsigned long long *trace; // SVD_GPU_OZ_TRACE: clock64 stamps of CTA 0, [tile][8]
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
        for (int b = 0; b < 2; ++b) { oz_mbar_init(t_full + b, 1); oz_mbar_init(t_empty + b, 8); }
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
                for (int i = 0; i < OZ_SL; i += 4)           // 4 planes (64 KB) per copy
                    oz_tma_3d(sA + i * OZ_BM * OZ_K, &mapA, 0, mb * OZ_BM, i, a_full);
                const int t1 = min(ntiles, (ng + 1) * OZ_NT);
                for (int t = ng * OZ_NT; t < t1; ++t, ++ub) {
                    const unsigned s = ub % OZ_BSTAGES;
                    oz_mbar_wait(b_empty + s, ((ub / OZ_BSTAGES) & 1u) ^ 1u);
                    OZ_TR(0, ub);
                    oz_mbar_expect_tx(b_full + s, OZ_B_BYTES);
                    oz_tma_3d(sB + s * OZ_B_BYTES, &mapB, 0, t * OZ_BN, 0, b_full + s);      // all 8 planes in one copy
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
                    OZ_TR(1, ut);
                    oz_mbar_wait(b_full + s, (ub / OZ_BSTAGES) & 1u);
                    OZ_TR(2, ut);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t a0 = oz_smem_desc(oz_u32(sA));
                    const uint64_t b0 = oz_smem_desc(oz_u32(sB + s * OZ_B_BYTES));
                    const unsigned a_lo = (unsigned)a0, b_lo = (unsigned)b0, d_hi = (unsigned)(a0 >> 32);
                    const unsigned d0 = tmem + buf * 256;
                    {
                        // 36 plane pairs x 4 K steps, fully unrolled: plane i of A starts i*16 KB, plane j of B j*4 KB
                        // further, 32 bytes of K per instruction (descriptor addresses count 16-byte units)
#pragma unroll
                        for (int gsum = 0; gsum < OZ_SL; ++gsum) {
#pragma unroll
                            for (int i = 0; i <= gsum; ++i) {
#pragma unroll
                                for (int k4 = 0; k4 < OZ_K / 32; ++k4) {
                                    const unsigned ao = (unsigned)((i * OZ_BM * OZ_K) >> 4) + 2u * k4;
                                    const unsigned bo = (unsigned)(((gsum - i) * OZ_BN * OZ_K) >> 4) + 2u * k4;
                                    if (i == 0 && k4 == 0) oz_mma_i8<false>(d0 + gsum * OZ_BN, a_lo + ao, b_lo + bo, d_hi);
                                    else oz_mma_i8<true>(d0 + gsum * OZ_BN, a_lo + ao, b_lo + bo, d_hi);
                                }
                            }
                        }
                    }
                    OZ_TR(3, ut);
                    oz_commit(b_empty + s);              // the B stage is free once these MMAs have read it
                    oz_commit(t_full + buf);             // ... and the accumulators are complete
                }
                oz_commit(a_empty);                      // the A block is free once every MMA of the unit is done
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ====================================
        // 8 warps: warp e reads TMEM lanes [32 (e%4), +32) (the hardware ties a warp to the lane quarter e%4) and
        // the 16 accumulator columns [16 (e/4), +16) of every group.  A thread owns one row of the tile; its 16
        // values of C for the NEXT TWO tiles are already in flight while it recombines the current one (the read
        // of C is the long pole: 2 x 16 x 8 bytes x 256 threads = 64 KB in flight per SM).
        const int ew = warp - 4, quarter = ew & 3, chalf = ew >> 2;
        const double sc0 = scalbn(g.sign, g.expo[0] + g.expo[1] - 12);
        struct It { int u, t, t1, mb; };
        auto first = [&](It &it) {
            it.u = blockIdx.x;
            if (it.u < nunits) { it.mb = it.u % mblocks; const int ng = it.u / mblocks; it.t = ng * OZ_NT; it.t1 = min(ntiles, (ng + 1) * OZ_NT); }
        };
        auto next = [&](It &it) {
            if (++it.t < it.t1) return;
            it.u += gridDim.x;
            if (it.u < nunits) { it.mb = it.u % mblocks; const int ng = it.u / mblocks; it.t = ng * OZ_NT; it.t1 = min(ntiles, (ng + 1) * OZ_NT); }
        };
        auto fetch = [&](const It &it, double (&c)[16]) {
            const int row = it.mb * OZ_BM + quarter * 32 + lane;
            const int n0 = it.t * OZ_BN + chalf * 16;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                c[q] = (it.u < nunits && row < g.M && n0 + q < g.N) ? __ldcg(g.C + row + (long)(n0 + q) * g.ldc) : 0.0;
        };
        It cur, pf;
        first(cur);
        pf = cur;
        // one tile: wait for its accumulators, recombine, update C from the values in `cc`
        auto do_tile = [&](const It &it, const double (&cc)[16], unsigned ut) {
            const unsigned buf = ut & 1u;
            oz_mbar_wait(t_full + buf, (ut >> 1) & 1u);
            if (ew == 0 && lane == 0) OZ_TR(4, ut);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tbase = tmem + ((unsigned)(quarter * 32) << 16) + buf * 256 + chalf * 16;
            const int row = it.mb * OZ_BM + quarter * 32 + lane;
#pragma unroll
            for (int c8 = 0; c8 < 16; c8 += 8) {
                int v[OZ_SL][8];
#pragma unroll
                for (int gsum = 0; gsum < OZ_SL; ++gsum) oz_tmem_ld8(tbase + gsum * OZ_BN + c8, v[gsum]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c8 == 8) {
                    // every accumulator column of this warp has been read: hand the buffer back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) oz_mbar_arrive(t_empty + buf);
                    if (ew == 0 && lane == 0) OZ_TR(5, ut);
                }
                const int n0 = it.t * OZ_BN + chalf * 16 + c8;
                if (row < g.M) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (n0 + q < g.N) {
                            // neighbouring groups are 2^-7 apart and |acc_g| <= 8 * 128 * 64 * 64 = 2^22: a pair fits an
                            // int32 exactly; four exact conversions, smallest term first
                            const int h0 = v[0][q] * 128 + v[1][q], h1 = v[2][q] * 128 + v[3][q];
                            const int h2 = v[4][q] * 128 + v[5][q], h3 = v[6][q] * 128 + v[7][q];
                            double acc = (double)h3 * 0x1p-49;
                            acc = fma((double)h2, 0x1p-35, acc);
                            acc = fma((double)h1, 0x1p-21, acc);
                            acc = fma((double)h0, 0x1p-7, acc);
                            g.C[row + (long)(n0 + q) * g.ldc] = fma(acc, sc0, cc[c8 + q]);
                        }
                    }
                }
            }
            if (ew == 0 && lane == 0) OZ_TR(6, ut);
        };
        // software pipeline over the CTA's tiles, three register sets in rotation: while tile k is recombined
        // the values of C for tiles k+1 and k+2 are in flight (no register copies: a copy would wait for the load)
        double c0[16], c1[16], c2[16];
        fetch(pf, c0);
        if (pf.u < nunits) next(pf);
        fetch(pf, c1);
        unsigned ut = 0;
        while (cur.u < nunits) {
            if (pf.u < nunits) next(pf);
            fetch(pf, c2);
            do_tile(cur, c0, ut++); next(cur);
            if (cur.u >= nunits) break;
            if (pf.u < nunits) next(pf);
            fetch(pf, c0);
            do_tile(cur, c1, ut++); next(cur);
            if (cur.u >= nunits) break;
            if (pf.u < nunits) next(pf);
            fetch(pf, c1);
            do_tile(cur, c2, ut++); next(cur);
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
    const double sc = scalbn(64.0, -*expo);              // x 2^-e 2^6: the first digit is rint of this
    union { signed char b[16]; int4 v; } out[OZ_SL];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = kc * 16 + q;
        double y = (r < R) ? X[r * sr + (long)k * sk] * sc : 0.0;
#pragma unroll
        for (int i = 0; i < OZ_SL; ++i) {
            // digit i = rint(residual * 2^(6+7i)); the residual is carried pre-scaled: every step is exact
            const double d = rint(y);
            out[i].b[q] = (signed char)__double2int_rn(d);
            y = (y - d) * 128.0;
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
// int8 planes [plane][row][128], box = 128 bytes x box_rows x box_planes, SWIZZLE_128B
CUtensorMap make_map(const signed char *planes, long rows_per_plane, int box_rows, int box_planes)
{
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)OZ_K, (cuuint64_t)rows_per_plane, (cuuint64_t)OZ_SL};
    cuuint64_t strides[2] = {(cuuint64_t)OZ_K, (cuuint64_t)OZ_K * (cuuint64_t)rows_per_plane};
    cuuint32_t box[3] = {(cuuint32_t)OZ_K, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)planes, dims, strides, box, estr,
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
    const CUtensorMap mapA = make_map(pa, Mpad, OZ_BM, 4);
    const CUtensorMap mapB = make_map(pb, Npad, OZ_BN, OZ_SL);
    OzArgs g;
    g.M = M; g.N = N; g.C = C; g.ldc = ldc; g.Mpad = Mpad; g.Npad = Npad; g.expo = expo; g.sign = sign;
    g.trace = nullptr;
    static unsigned long long *d_trace = nullptr;
    const bool tracing = getenv("SVD_GPU_OZ_TRACE") != nullptr;
    if (tracing) {
        if (!d_trace) SVD_CUDA_CHECK(cudaMalloc(&d_trace, 128 * 8 * sizeof(unsigned long long)));
        SVD_CUDA_CHECK(cudaMemsetAsync(d_trace, 0, 128 * 8 * sizeof(unsigned long long), st));
        g.trace = d_trace;
    }
    int dev = 0, nsm = 148;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    static DeviceOnce once;
    if (first_on_device(once))
        SVD_CUDA_CHECK(cudaFuncSetAttribute(oz_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    const int mblocks = (int)(Mpad / OZ_BM), ngroups = ceil_div(ceil_div(N, OZ_BN), OZ_NT);
    const int nunits = mblocks * ngroups;
    oz_update_kernel<<<nunits < nsm ? nunits : nsm, OZ_THREADS, OZ_SMEM, st>>>(mapA, mapB, g);
    SVD_KERNEL_CHECK();
    if (tracing) {
        static unsigned long long h[128 * 8];
        SVD_CUDA_CHECK(cudaMemcpyAsync(h, d_trace, sizeof h, cudaMemcpyDeviceToHost, st));
        SVD_CUDA_CHECK(cudaStreamSynchronize(st));
        unsigned long long t0 = ~0ull;
        for (int z = 0; z < 128 * 8; ++z) if (h[z] && h[z] < t0) t0 = h[z];
        fprintf(stderr, "OZTRACE tile: producer-issue | mma: tmem-free, B-landed, issued | epilogue: acc-ready, tmem-released, done\n");
        for (int nt = 0; nt < 48; ++nt) {
            fprintf(stderr, "OZTRACE %3d", nt);
            for (int z = 0; z < 7; ++z) fprintf(stderr, " %8lld", h[nt * 8 + z] ? (long long)(h[nt * 8 + z] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
    }
}

} // namespace svdgpu
