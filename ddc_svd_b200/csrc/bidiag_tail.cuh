// bidiag_tail.cuh — the END of the bidiagonalization with the trailing matrix resident on chip.
//
// Once the trailing block (rows and columns >= i0) fits into the shared memory of all SMs together
// (148 x ~220 KB: up to ~2000 x 2000 doubles), the remaining steps of bidiag_par (bidiag_par.c:310-397,
// the loop of bidiag.c:73-183) are latency-bound in the streaming path: two kernel boundaries per step
// cost more than the data movement.  This kernel finishes the job in ONE cooperative launch:
//   * the trailing columns are dealt round-robin to the CTAs and stay in shared memory, updated in
//     place (no deferred panel: shared-memory sweeps are cheap);
//   * a step is  sweep 1 (pending right update of the previous step applied on the fly, column dots
//     v^T a_j) -> sweep 2 (left update, row i, partial A r)  ->  exchange  ->  row slices reduce the
//     partials: beta_i, x = 2 A u, next column c'  ->  exchange;
//   * only vectors cross CTAs (through L2): per-CTA partials of A r, x, c', a handful of scalars - each
//     value carries its step tag, so an exchange is one store and one polled load, not a grid barrier
//     (a counter barrier measured ~2 us: fence + atomic + poll, twice per step);
//   * reflectors, alpha and beta leave in the reference's layout (v_i in A[i:m, i], u_i in A[i, i+1:n],
//     unit norm, sign rule of update_scale_matcol.cl:58-82 via make_refl).
// Same algebra as the fused streaming step (row dots taken with the un-normalised row, norm and
// scale applied afterwards), so a step needs exactly two grid-wide exchanges.
// Square / tall inputs only (m >= n): wide ones reach the library transposed.
#pragma once

namespace svdgpu {

constexpr int TL_THREADS = 512;
constexpr int TL_WARPS = TL_THREADS / 32;
constexpr int TL_RPT = 4;                         // rows per thread
constexpr int TL_MAXROWS = TL_THREADS * TL_RPT;   // 2048
constexpr int TL_CPC = 16;                        // columns per CTA (register arrays)
constexpr int TL_MAXG = 148;                      // CTAs (= partial vectors per row; <= 160: 5 per lane)
constexpr int TL_CAP = 28000;                     // doubles of shared memory for the matrix (224 KB)
constexpr int TL_AUX = 4 * TL_CPC + TL_WARPS * (TL_CPC + 1) + 64;

// One value of a cross-CTA exchange: the 8 data bytes travel with a step tag in the same 16-byte
// word ({lo, tag, hi, tag}: each 8-byte half is single-copy atomic and self-validating), so the
// reader polls the data itself - no fence, no grid barrier, no second round trip.  Buffers are
// zeroed before the launch; tags start at 1 and grow with the step, a slot is rewritten only after
// every reader of its previous value has (transitively) delivered what the writer needed first.
struct __align__(16) TlSlot { unsigned lo, t0, hi, t1; };

struct TailArgs {
    double *A; long lda; int m, n, i0;
    double *alpha, *beta;
    TlSlot *W;                                    // [G][TL_MAXROWS] per-CTA partial A r, by local row
    TlSlot *X, *C, *A1;                           // [TL_MAXROWS] x = 2 A u, next column c', column i+1 after H_i
    TlSlot *RR, *R1;                              // [2][TL_MAXG] partial r.r, [2] r_{i+1} (by step parity)
    unsigned *counter;                            // grid barrier used once at entry (zeroed by the host)
    int Lp;                                       // rows of a shared-memory column (even)
};
constexpr size_t TL_WS_SLOTS = (size_t)TL_MAXG * TL_MAXROWS + 3 * TL_MAXROWS + 2 * TL_MAXG + 8;

// Both halves travel as 64-bit elements of one 16-byte vector access: the PTX memory model makes every
// naturally aligned element of a vector access single-copy atomic, so {lo, tag} and {hi, tag} can each
// only be seen whole (as 32-bit elements a torn {lo, tag} pair would be legal, if unseen on this hardware).
__device__ __forceinline__ void tl_put(TlSlot *s, double v, unsigned tag)
{
    const unsigned long long a = (unsigned long long)(unsigned)__double2loint(v) | ((unsigned long long)tag << 32);
    const unsigned long long b = (unsigned long long)(unsigned)__double2hiint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(s), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ bool tl_try(const TlSlot *s, unsigned tag, double &v)
{
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(s) : "memory");
    v = __hiloint2double((int)(unsigned)b, (int)(unsigned)a);
    return (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag;
}
// a poll that lasts seconds is a protocol bug: trap instead of hanging the device
struct TlWatch {
    long long t0; int spin;
    __device__ __forceinline__ TlWatch() : t0(0), spin(0) {}
    __device__ __forceinline__ void tick()
    {
        if (++spin == 64) t0 = clock64();
        if (spin > 64 && (spin & 63) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
};

__device__ __forceinline__ void tl_grid_barrier(unsigned *counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        TlWatch wd;
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if ((int)(v - target) >= 0) break;
            wd.tick();
        }
        __threadfence();
    }
    __syncthreads();
}

// sum of one value per thread over the CTA, same order in every CTA (all CTAs must agree bitwise)
__device__ __forceinline__ double tl_block_sum(double v, double *s_w)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TL_WARPS; ++w) s += s_w[w];
    return s;
}
// G tagged values (one per CTA), summed by ONE warp in a fixed order; lane 0 may fetch one extra slot.
// Only one warp per CTA reads them: 148 CTAs x 16 warps on the same few lines is an L2 hot spot.
__device__ __forceinline__ double tl_poll_sum(const TlSlot *base, int G, unsigned tag, const TlSlot *extra, double &xv)
{
    const int lane = threadIdx.x & 31;
    double sv[5];
    TlWatch wd;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            const int bb = lane + 32 * e;
            sv[e] = 0.0;
            if (bb < G) ok &= tl_try(base + bb, tag, sv[e]);
        }
        if (extra != nullptr && lane == 0) ok &= tl_try(extra, tag, xv);
        if (__all_sync(0xffffffffu, ok)) break;
        wd.tick();
    }
    xv = __shfl_sync(0xffffffffu, xv, 0);
    return warp_sum(((sv[0] + sv[1]) + (sv[2] + sv[3])) + sv[4]);
}

// one reflector from (first entry, squared norm).  VER 2: one sqrt and one rsqrt on the critical path instead of
// two square roots and a divide (same value up to rounding: inv = 1 / (sqrt(2) sqrt(nu^2 + |nu x0|)))
template <int VER> __device__ __forceinline__ Refl tl_refl(double x0, double nrm2)
{
    if constexpr (VER == 1) {
        return make_refl(x0, nrm2);
    } else {
        Refl f;
        const double nu = sqrt(nrm2);
        f.snu = (x0 < 0.0) ? -nu : nu;
        const double q = nu * nu + fabs(nu * x0);
        f.inv = (q > 0.0) ? rsqrt(2.0 * q) : 0.0;
        return f;
    }
}

// VER 1: the version validated and measured in round 1 (default).
// VER 2 (SVD_GPU_TAIL=2, experiment): sweeps as loops over the live columns only (the 16 x 4 unrolled code of VER 1
// is 2 900 instructions and misses the instruction cache), the cheaper reflector, and polls that re-read only the
// slots that had not arrived (every retry of VER 1 re-reads all of x and c': 4.9 MB of L2 traffic per round at 1024 rows).
template <int VER>
__global__ void __launch_bounds__(TL_THREADS, 1) bidiag_tail_kernel(TailArgs p)
{
    extern __shared__ __align__(16) double tl_sm[];
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int m = p.m, n = p.n, i0 = p.i0, Lp = p.Lp;
    const int L0 = m - i0;
    const int ncol = (n - i0 - b + G - 1) / G;          // own columns: j = i0 + b + q*G, q < ncol (may be <= 0)
    double *a = tl_sm;                                  // [ncol][Lp]
    double *s_t = a + (size_t)TL_CPC * 0 + (size_t)(ncol > 0 ? ncol : 0) * Lp;
    double *s_r = s_t + TL_CPC, *s_u = s_r + TL_CPC, *s_w = s_u + TL_CPC;     // s_w: TL_WARPS + 8
    double *s_red = s_w + TL_WARPS + 8;                                          // [TL_WARPS][TL_CPC+1]
    double *A = p.A;
    const long lda = p.lda;

    // ---- trailing columns into shared memory (rows i0..m-1), round-robin over the CTAs
    for (int q = 0; q < ncol; ++q) {
        const double *src = A + i0 + (long)(i0 + b + q * G) * lda;
        for (int rho = t; rho < L0; rho += TL_THREADS) a[q * Lp + rho] = src[rho];
    }
    if (t < TL_CPC) { s_u[t] = 0.0; s_r[t] = 0.0; s_t[t] = 0.0; }
    // current column c (rows >= i0 of column i0) and c.c, identical in every CTA
    double cr[TL_RPT], xr[TL_RPT];
    double part = 0.0;
#pragma unroll
    for (int z = 0; z < TL_RPT; ++z) {
        const int rho = t + z * TL_THREADS;
        cr[z] = (rho < L0) ? A[i0 + rho + (long)i0 * lda] : 0.0;
        xr[z] = 0.0;
        part += cr[z] * cr[z];
    }
    double cc = tl_block_sum(part, s_w);
    double ci = A[i0 + (long)i0 * lda];
    // every CTA has read column i0 (and its own columns) from A before anyone stores a reflector there
    tl_grid_barrier(p.counter, (unsigned)G);

    for (int i = i0; i < n; ++i) {
        const int lo = i - i0;                           // first active local row
        // ---- column reflector from (c, c.c): v = (c + s*nu*e_i) * inv, alpha_i = -s*nu
        const Refl f = tl_refl<VER>(ci, cc);
        double v[TL_RPT];
#pragma unroll
        for (int z = 0; z < TL_RPT; ++z) {
            const int rho = t + z * TL_THREADS;
            v[z] = (rho >= lo && rho < L0) ? (cr[z] + (rho == lo ? f.snu : 0.0)) * f.inv : 0.0;
        }
        if (b == lo % G) {                               // the CTA that owned column i stores the reflector
#pragma unroll
            for (int z = 0; z < TL_RPT; ++z) {
                const int rho = t + z * TL_THREADS;
                if (rho >= lo && rho < L0) A[i0 + rho + (long)i * lda] = v[z];
            }
            if (t == 0) p.alpha[i] = -f.snu;
        }
        if (i == n - 1) break;                           // tall / square tail: last column reflector only
        const double vi = (ci + f.snu) * f.inv;          // v at row i
        const int qlo = (lo - b + G) / G;                // first own column with j > i  (j = i0 + b + q*G)
        const int qlo_c = qlo < 0 ? 0 : qlo;

        // ---- sweep 1: pending right update of step i-1 (a -= x u_j), column dots t_j = v^T a_j
        if constexpr (VER == 1) {
            double pq[TL_CPC];
#pragma unroll
            for (int q = 0; q < TL_CPC; ++q) {
                pq[q] = 0.0;
                if (q >= qlo_c && q < ncol) {
                    const double uq = s_u[q];
                    double *col = a + q * Lp;
#pragma unroll
                    for (int z = 0; z < TL_RPT; ++z) {
                        const int rho = t + z * TL_THREADS;
                        if (rho >= lo && rho < L0) {
                            const double aa = col[rho] - xr[z] * uq;
                            col[rho] = aa;
                            pq[q] += v[z] * aa;
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < TL_CPC; ++q)
                if (q >= qlo_c && q < ncol) {                // CTA-uniform
                    const double s = warp_sum(pq[q]);
                    if (lane == 0) s_red[warp * (TL_CPC + 1) + q] = s;
                }
        } else {
            for (int q = qlo_c; q < ncol; ++q) {
                const double uq = s_u[q];
                double *col = a + q * Lp;
                double pp = 0.0;
#pragma unroll
                for (int z = 0; z < TL_RPT; ++z) {
                    const int rho = t + z * TL_THREADS;
                    if (rho >= lo && rho < L0) {
                        const double aa = col[rho] - xr[z] * uq;
                        col[rho] = aa;
                        pp += v[z] * aa;
                    }
                }
                pp = warp_sum(pp);
                if (lane == 0) s_red[warp * (TL_CPC + 1) + q] = pp;
            }
        }
        __syncthreads();
        if (t < TL_CPC && t >= qlo_c && t < ncol) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < TL_WARPS; ++w) s += s_red[w * (TL_CPC + 1) + t];
            s_t[t] = s;
            // row i after H_i (un-normalised row reflector entry of this column)
            s_r[t] = a[t * Lp + lo] - 2.0 * vi * s;
        }
        __syncthreads();

        // ---- sweep 2: left update a -= 2 v t_j (rows > i), partial w = A[i+1:, own cols] r
        const unsigned tag = (unsigned)(lo + 1);
        double w[TL_RPT];
#pragma unroll
        for (int z = 0; z < TL_RPT; ++z) w[z] = 0.0;
        double rr = 0.0;
        if constexpr (VER == 1) {
#pragma unroll
            for (int q = 0; q < TL_CPC; ++q) {
                if (q >= qlo_c && q < ncol) {
                    const double tq2 = 2.0 * s_t[q], rq = s_r[q];
                    double *col = a + q * Lp;
                    rr += rq * rq;
#pragma unroll
                    for (int z = 0; z < TL_RPT; ++z) {
                        const int rho = t + z * TL_THREADS;
                        if (rho > lo && rho < L0) {
                            const double aa = col[rho] - v[z] * tq2;
                            col[rho] = aa;
                            w[z] += aa * rq;
                        }
                    }
                }
            }
        } else {
            for (int q = qlo_c; q < ncol; ++q) {
                const double tq2 = 2.0 * s_t[q], rq = s_r[q];
                double *col = a + q * Lp;
                rr += rq * rq;
#pragma unroll
                for (int z = 0; z < TL_RPT; ++z) {
                    const int rho = t + z * TL_THREADS;
                    if (rho > lo && rho < L0) {
                        const double aa = col[rho] - v[z] * tq2;
                        col[rho] = aa;
                        w[z] += aa * rq;
                    }
                }
            }
        }
#pragma unroll
        for (int z = 0; z < TL_RPT; ++z) {
            const int rho = t + z * TL_THREADS;
            if (rho > lo && rho < L0) tl_put(p.W + (size_t)b * TL_MAXROWS + rho, w[z], tag);
        }
        // RR / R1 are read by EVERY CTA, also by those that own no row and therefore deliver nothing the
        // others wait for: two slots by step parity keep a fast CTA from overwriting what a slow one
        // still polls (nobody can be two steps ahead: each step needs everybody's RR)
        TlSlot *RRs = p.RR + (lo & 1) * TL_MAXG, *R1s = p.R1 + (lo & 1);
        if (t == 0) tl_put(RRs + b, rr, tag);            // (rr is the same in every thread)
        __syncthreads();                                 // sweep 2 stores of column i+1 visible below
        if (b == (lo + 1) % G) {                         // owner of column i+1: a1 = that column after H_i, r_{i+1}
            const int q1 = (lo + 1) / G;
            const double *col = a + q1 * Lp;
#pragma unroll
            for (int z = 0; z < TL_RPT; ++z) {
                const int rho = t + z * TL_THREADS;
                if (rho > lo && rho < L0) tl_put(p.A1 + rho, col[rho], tag);
            }
            if (t == 0) tl_put(R1s, s_r[q1], tag);
        }

        // ---- row slices: warp w of CTA b owns one row below row i; warp 0 also fetches the scalars
        const bool has_row = (i < n - 2);
        const int Lb = m - i - 1;
        const int S = (Lb + G - 1) / G;                  // rows per CTA (<= 14 < TL_WARPS: 2048 rows over 148 CTAs)
        const int rl = b * S + warp;                     // this warp's row within the block below row i
        const bool myrow = (warp < S && rl < Lb);
        const int rrow = lo + 1 + rl;                    // local row
        double acc = 0.0, a1r = 0.0;
        if (myrow) {
            double wv[5];
            TlWatch wd;
            if constexpr (VER == 1) {
                for (;;) {
                    bool ok = true;
#pragma unroll
                    for (int e = 0; e < 5; ++e) {
                        const int bb = lane + 32 * e;
                        wv[e] = 0.0;
                        if (bb < G) ok &= tl_try(p.W + (size_t)bb * TL_MAXROWS + rrow, tag, wv[e]);
                    }
                    if (lane == 0) ok &= tl_try(p.A1 + rrow, tag, a1r);
                    if (__all_sync(0xffffffffu, ok)) break;
                    wd.tick();
                }
            } else {
                unsigned pend = 0;                       // bit e: partial of CTA lane + 32 e, bit 5: a1 (lane 0)
#pragma unroll
                for (int e = 0; e < 5; ++e) { wv[e] = 0.0; if (lane + 32 * e < G) pend |= 1u << e; }
                if (lane == 0) pend |= 1u << 5;
                for (;;) {
#pragma unroll
                    for (int e = 0; e < 5; ++e)
                        if ((pend >> e) & 1u) {
                            double got;
                            if (tl_try(p.W + (size_t)(lane + 32 * e) * TL_MAXROWS + rrow, tag, got)) { wv[e] = got; pend &= ~(1u << e); }
                        }
                    if ((pend >> 5) & 1u) {
                        double got;
                        if (tl_try(p.A1 + rrow, tag, got)) { a1r = got; pend &= ~(1u << 5); }
                    }
                    if (__all_sync(0xffffffffu, pend == 0)) break;
                    wd.tick();
                }
            }
            a1r = __shfl_sync(0xffffffffu, a1r, 0);
            acc = warp_sum(((wv[0] + wv[1]) + (wv[2] + wv[3])) + wv[4]);
        }
        if (warp == TL_WARPS - 1) {                      // never a row owner (S <= 14): polls alongside the row warps
            double r1v = 0.0;
            const double tot = tl_poll_sum(RRs, G, tag, R1s, r1v);
            if (lane == 0) { s_w[TL_WARPS] = tot; s_w[TL_WARPS + 1] = r1v; }
        }
        __syncthreads();
        const double rr_tot = s_w[TL_WARPS], r1 = s_w[TL_WARPS + 1];
        Refl g;
        if (has_row) g = tl_refl<VER>(r1, rr_tot); else { g.snu = 0.0; g.inv = 0.0; }
        const double u1 = (r1 + g.snu) * g.inv;
        if (b == 0 && t == 0) p.beta[i] = has_row ? -g.snu : r1;
        if (t < TL_CPC && t >= qlo_c && t < ncol) {
            const int j = i0 + b + t * G;
            const double u = (s_r[t] + (j == i + 1 ? g.snu : 0.0)) * g.inv;
            s_u[t] = u;
            A[i + (long)j * lda] = u;                    // u_i in place (zero when there is no row reflector)
        }
        // x = 2 A u, next column c' = a1 - x u_{i+1}
        if (myrow && lane == 0) {
            const double x = 2.0 * g.inv * (acc + g.snu * a1r);
            tl_put(p.X + rrow, x, tag);
            tl_put(p.C + rrow, a1r - x * u1, tag);
        }

        // ---- next step's inputs: x (pending right update), c', c'.c'
        if constexpr (VER == 1) {
            TlWatch wd;
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int z = 0; z < TL_RPT; ++z) {
                    const int rho = t + z * TL_THREADS;
                    xr[z] = 0.0; cr[z] = 0.0;
                    if (rho > lo && rho < L0) {
                        ok &= tl_try(p.X + rho, tag, xr[z]);
                        ok &= tl_try(p.C + rho, tag, cr[z]);
                    }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                wd.tick();
            }
        } else {
            TlWatch wd;
            unsigned pend = 0;                           // bit 2z: x of row t + 512 z, bit 2z+1: c'
#pragma unroll
            for (int z = 0; z < TL_RPT; ++z) {
                const int rho = t + z * TL_THREADS;
                xr[z] = 0.0; cr[z] = 0.0;
                if (rho > lo && rho < L0) pend |= 3u << (2 * z);
            }
            for (;;) {
#pragma unroll
                for (int z = 0; z < TL_RPT; ++z) {
                    const int rho = t + z * TL_THREADS;
                    double got;
                    if ((pend >> (2 * z)) & 1u)
                        if (tl_try(p.X + rho, tag, got)) { xr[z] = got; pend &= ~(1u << (2 * z)); }
                    if ((pend >> (2 * z + 1)) & 1u)
                        if (tl_try(p.C + rho, tag, got)) { cr[z] = got; pend &= ~(2u << (2 * z)); }
                }
                if (__all_sync(0xffffffffu, pend == 0)) break;
                wd.tick();
            }
        }
        // c'.c' from the full vector every CTA now holds (same thread-to-row map and reduction order in
        // every CTA, so all agree bitwise) - one exchange less than shipping per-CTA partials
        double part2 = 0.0;
#pragma unroll
        for (int z = 0; z < TL_RPT; ++z) {
            part2 += cr[z] * cr[z];
            if (t + z * TL_THREADS == lo + 1) s_w[TL_WARPS + 3] = cr[z];
        }
        cc = tl_block_sum(part2, s_w);
        ci = s_w[TL_WARPS + 3];
    }
}

} // namespace svdgpu
