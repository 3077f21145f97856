// qb_kernels.cuh -- device building blocks: complex128 row products for the three operator
// formats, warp/block reductions with a fixed (deterministic) order.
//
// Mapping used everywhere: one warp owns one 32-row slice, lane == row inside the slice.
// * DIAM  (diagonal-masked slices): the slice's entries are (diagonal offset, lane mask)
//   pairs; the values of an entry are packed in lane order, so the 32 lanes read one
//   contiguous run of the value array (coalesced, 16 B per lane) and x[r + offset] is a
//   contiguous run of the state too.  No per-nonzero column index is read: 16 B per
//   non-zero + 8 B per (slice, diagonal) instead of CSR's 20 B per non-zero.
// * CSR   : generic fallback, one lane per row (handles unsorted / duplicate indices).
// * DENSE : column-major A, lanes read consecutive rows of a column (coalesced).
#pragma once
#include <cuda_runtime.h>
#include "qb_types.h"

__device__ __forceinline__ double2 qb_ld(const qb_c128* p) {
    return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ void qb_fma(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 qb_mul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// (A x)[r] for the lane's row r of slice sl.  `active` lanes have r < nrows.
// x is a plain vector (stride 1).
__device__ __forceinline__ double2 qb_rowdot(const QbOpDev& A, int sl, int lane, long long r,
                                             bool active, const double2* __restrict__ x)
{
    double2 acc = make_double2(0.0, 0.0);
    if (A.fmt == QB_FMT_DIAM) {
        const int e0 = A.slice_ptr[sl], e1 = A.slice_ptr[sl + 1];
        long long vb = A.slice_vbase[sl];
        const unsigned lt = (1u << lane) - 1u;
        const int2* __restrict__ ent = reinterpret_cast<const int2*>(A.ent_off);
        const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val);
        int e = e0;
        for (; e + 4 <= e1; e += 4) {
            int2 d0 = ent[e], d1 = ent[e + 1], d2 = ent[e + 2], d3 = ent[e + 3];
            const unsigned m0 = (unsigned)d0.y, m1 = (unsigned)d1.y, m2 = (unsigned)d2.y,
                           m3 = (unsigned)d3.y;
            const long long b0 = vb, b1 = b0 + __popc(m0), b2 = b1 + __popc(m1),
                            b3 = b2 + __popc(m2);
            vb = b3 + __popc(m3);
            const bool h0 = (m0 >> lane) & 1u, h1 = (m1 >> lane) & 1u, h2 = (m2 >> lane) & 1u,
                       h3 = (m3 >> lane) & 1u;
            const double2 zero = make_double2(0.0, 0.0);
            double2 v0 = h0 ? val[b0 + __popc(m0 & lt)] : zero;
            double2 v1 = h1 ? val[b1 + __popc(m1 & lt)] : zero;
            double2 v2 = h2 ? val[b2 + __popc(m2 & lt)] : zero;
            double2 v3 = h3 ? val[b3 + __popc(m3 & lt)] : zero;
            double2 x0 = h0 ? x[r + d0.x] : zero;
            double2 x1 = h1 ? x[r + d1.x] : zero;
            double2 x2 = h2 ? x[r + d2.x] : zero;
            double2 x3 = h3 ? x[r + d3.x] : zero;
            qb_fma(acc, v0, x0); qb_fma(acc, v1, x1); qb_fma(acc, v2, x2); qb_fma(acc, v3, x3);
        }
        for (; e < e1; e++) {
            int2 d0 = ent[e];
            const unsigned m0 = (unsigned)d0.y;
            if ((m0 >> lane) & 1u) {
                double2 v0 = val[vb + __popc(m0 & lt)];
                double2 x0 = x[r + d0.x];
                qb_fma(acc, v0, x0);
            }
            vb += __popc(m0);
        }
    } else if (A.fmt == QB_FMT_CSR) {
        if (active) {
            const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val);
            const int p1 = A.rowptr[r + 1];
            for (int p = A.rowptr[r]; p < p1; p++) qb_fma(acc, val[p], x[A.col[p]]);
        }
    } else {
        if (active) {
            const double2* __restrict__ a = reinterpret_cast<const double2*>(A.dense) + r;
            const long long ld = A.nrows;
            int c = 0;
            for (; c + 4 <= A.ncols; c += 4) {
                double2 a0 = a[(long long)c * ld], a1 = a[(long long)(c + 1) * ld],
                        a2 = a[(long long)(c + 2) * ld], a3 = a[(long long)(c + 3) * ld];
                qb_fma(acc, a0, x[c]); qb_fma(acc, a1, x[c + 1]);
                qb_fma(acc, a2, x[c + 2]); qb_fma(acc, a3, x[c + 3]);
            }
            for (; c < A.ncols; c++) qb_fma(acc, a[(long long)c * ld], x[c]);
        }
    }
    return acc;
}

// butterfly sum: every lane ends with the same value, order fixed -> deterministic
__device__ __forceinline__ double qb_warp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
