// Per-candidate LMMSE SINR core shared by the PMI kernels (comm.cu, comm_fused.cu): packed lower-triangle Cholesky of
// (HW)'(HW) + nVar I and the diagonal of its inverse (getPrecodedSINR, dlPMISelect.m:1825-1834; precodedSINR.m:11-18).
#pragma once
#include "isac_common.cuh"

namespace isac {

__device__ __forceinline__ double round4(double x) { return copysign(floor(fabs(x) * 1e4 + 0.5) / 1e4, x); }

#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))   /* packed lower triangle, i >= j: stays in registers */

// A (packed lower triangle of (HW)'(HW)) -> sinr_l = 1/(nVar [ (A + nVar I)^-1 ]_ll) - 1, l = 0..NU-1, written with stride
// (dlPMISelect.m:1831-1833).  Right-looking Cholesky A = L L^H in place (after step j the trailing block is updated, so
// every update of a step is independent of the others), then L is inverted in place column by column from the last one
// (T = L^-1, T_ij = -T_jj sum_{k=j+1..i} T_ik L_kj: the entries of one column are independent of each other) and
// [A^-1]_cc = sum_{i>=c} |T_ic|^2.  Same operation count as the left-looking factorisation + one forward substitution per
// column this replaces, but short dependency chains: the FP64 pipe is fed by instruction-level parallelism, the register
// budget (NU(NU+1) doubles for A) leaves no room for more resident warps.
// Every loop runs 0..NU with a compile-time guard, so each one unrolls on its own (trip counts that depend on an outer
// induction variable made the unroller fall back to local memory for some NU).
template <int NU>
__device__ __forceinline__ void chol_sinr(double2 (&A)[NU * (NU + 1) / 2], double nVar, double* __restrict__ out, long long stride) {
#pragma unroll
    for (int i = 0; i < NU; ++i) A[TRI(i, i)].x += nVar;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        // store 1/L_jj on the diagonal (MUFU seed + Newton instead of sqrt and two divisions)
        const double inv = fast_rsqrt(A[TRI(j, j)].x);
        A[TRI(j, j)] = make_double2(inv, 0.0);
#pragma unroll
        for (int i = 0; i < NU; ++i)
            if (i > j) A[TRI(i, j)] = make_double2(A[TRI(i, j)].x * inv, A[TRI(i, j)].y * inv);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            if (i > j) {
#pragma unroll
                for (int k = 0; k < NU; ++k)
                    if (k > j && k < i) A[TRI(i, k)] = zfmsc(A[TRI(i, k)], A[TRI(i, j)], A[TRI(k, j)]);
                A[TRI(i, i)].x = fma(-A[TRI(i, j)].x, A[TRI(i, j)].x, fma(-A[TRI(i, j)].y, A[TRI(i, j)].y, A[TRI(i, i)].x));
            }
        }
    }
#pragma unroll
    for (int jj = 0; jj < NU; ++jj) {
        const int j = NU - 1 - jj;
        double2 col[NU];
#pragma unroll
        for (int k = 0; k < NU; ++k)
            if (k > j) col[k] = A[TRI(k, j)];
        const double ninv = -A[TRI(j, j)].x;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            if (i > j) {
                double2 s = make_double2(A[TRI(i, i)].x * col[i].x, A[TRI(i, i)].x * col[i].y);   // T_ii is real
#pragma unroll
                for (int k = 0; k < NU; ++k)
                    if (k > j && k < i) s = zfma(s, A[TRI(i, k)], col[k]);
                A[TRI(i, j)] = make_double2(s.x * ninv, s.y * ninv);
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < NU; ++cc) {
        double nrm = A[TRI(cc, cc)].x * A[TRI(cc, cc)].x;
#pragma unroll
        for (int i = 0; i < NU; ++i)
            if (i > cc) nrm = fma(A[TRI(i, cc)].x, A[TRI(i, cc)].x, fma(A[TRI(i, cc)].y, A[TRI(i, cc)].y, nrm));
        out[(long long)cc * stride] = fast_rcp(nVar * nrm) - 1.0;
    }
}

}  // namespace isac
