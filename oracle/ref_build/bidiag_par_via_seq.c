/* TEST INFRASTRUCTURE ONLY — routes the reference's device entry point
 * bidiag_par() (bidiag_par.h:30) to its own serial prototype bidiag_seq()
 * (bidiag.c:33), which north_star names as the CPU path ("serial bidiag.c plus
 * host dDC/twisted code").  bidiag_par.c itself is compiled with
 * -Dbidiag_par=bidiag_par_opencl so its multU/multV stay available while the
 * OpenCL entry (which cannot run here) gets out of the way.
 */
#include "bidiag.h"
void bidiag_par(int m, int n, double *restrict A, double *restrict alpha, double *restrict beta)
{
    bidiag_seq(m, n, A, alpha, beta);
}
