// qb_kernels.cuh -- device building blocks: complex128 row products for the three operator
// formats, warp/block reductions with a fixed (deterministic) order.
//
// Mapping used everywhere: one warp owns one 32-row slice, lane == row inside the slice.
// * DIAM  (diagonal-masked slices): the slice's entries are (diagonal offset, lane mask)
//   pairs; the values of an entry are packed in lane order, so the 32 lanes read one
//   contiguous run of the value array (coalesced, 16 B per lane) and x[r + offset] is a
//   contiguous run of the state too.  No per-nonzero column index is read: 16 B per
//   non-zero + 8 B per (slice, diagonal) instead of CSR's 20 B per non-zero.
//   The entry list of a slice is fetched once by the warp (lane e holds entry e), the
//   value offsets come from a warp prefix sum of the mask popcounts, and the main loop
//   broadcasts (offset, mask, base) by shuffle -- so no global load in the loop depends on
//   another one and the unrolled loop keeps 8 x (1 + G) independent 16-byte loads per lane
//   in flight.  G state vectors (trajectories) can share one pass over the operator.
// * SELL  (sliced ELLPACK, slot-major, explicit column indices): for operators that stay
//   L2-resident while many trajectories re-read them; three coalesced loads + four FMAs per
//   slot and no mask arithmetic -> several times fewer instructions than DIAM.
// * CSR   : generic fallback, one lane per row (handles unsorted / duplicate indices).
// * DENSE : column-major A, lanes read consecutive rows of a column (coalesced).
#pragma once
#include <cuda_runtime.h>
#include "qb_types.h"

// Entries kept in flight per lane by the DIAM sweep.  Measured on B200 (tools/perf_ab.sh):
// occupancy beats deeper unrolling -- U=3 at 64 registers (32 warps/SM) gives the best
// stand-alone SpMV, U=2 the best fused pass kernel.
#ifndef QB_U1
#define QB_U1 4     // stand-alone SpMV / matmul / expect kernels
#endif
#ifndef QB_UP
#define QB_UP 2     // inside the fused pass kernel
#endif

__device__ __forceinline__ void qb_fma(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 qb_mul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc[g] += (A x_g)[r] for the lane's row of slice sl, g < G.  DIAM only.  Whole-warp call.
// A.pad_ != 0 marks an operator far larger than L2: its values are read with the
// cache-streaming policy so they do not evict the (re-used) state vectors.
template <int G, int QB_UNROLL>
__device__ __forceinline__ void qb_rowdot_diam(const QbOpDev& A, int sl, int lane, long long r64,
                                               const double2* const* x, double2 (&acc)[G])
{
    const int e0 = A.slice_ptr[sl], e1 = A.slice_ptr[sl + 1];
    long long vbase = A.slice_vbase[sl];
    const int r = (int)r64;                 // 32-bit indexing: rows, cols < 2^31
    const unsigned lt = (1u << lane) - 1u;
    const bool stream = A.pad_ != 0;
    const int2* __restrict__ ent = reinterpret_cast<const int2*>(A.ent_off);
    const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val);
    const double2 zero = make_double2(0.0, 0.0);
    for (int c0 = e0; c0 < e1; c0 += 32) {
        const int ne = min(32, e1 - c0);
        int2 my = make_int2(0, 0);
        if (lane < ne) my = ent[c0 + lane];
        // exclusive prefix sum of popc(mask) over the chunk's entries
        const int cnt = __popc((unsigned)my.y);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int my_vb = incl - cnt;
        const int chunk_total = __shfl_sync(0xffffffffu, incl, 31);
        const double2* vchunk = val + vbase;
        // every iteration is a full unrolled batch: entries past `ne` come from lanes whose
        // mask is zero (lanes >= ne hold (0, 0)), so they are predicated off -- no serial tail
        for (int e = 0; e < ne; e += QB_UNROLL) {
            double2 v[QB_UNROLL];
            double2 xv[QB_UNROLL][G];
#pragma unroll
            for (int u = 0; u < QB_UNROLL; u++) {
                const int src = (e + u) & 31;
                const int off = __shfl_sync(0xffffffffu, my.x, src);
                unsigned m = (unsigned)__shfl_sync(0xffffffffu, my.y, src);
                const int vb = __shfl_sync(0xffffffffu, my_vb, src);
                if (e + u >= ne) m = 0u;
                const bool hit = (m >> lane) & 1u;
                const double2* vp = vchunk + (vb + __popc(m & lt));
                if (stream) v[u] = hit ? __ldcs(vp) : zero;
                else v[u] = hit ? __ldg(vp) : zero;
#pragma unroll
                for (int g = 0; g < G; g++) xv[u][g] = hit ? x[g][r + off] : zero;
            }
#pragma unroll
            for (int u = 0; u < QB_UNROLL; u++)
#pragma unroll
                for (int g = 0; g < G; g++) qb_fma(acc[g], v[u], xv[u][g]);
        }
        vbase += chunk_total;
    }
}

#ifndef QB_KU
#define QB_KU 6      // entries of A's row j in flight in the right product rho A^dagger
#endif
// Matrix-free Kronecker operators on the column-stacked n x n state (row r = i + j n):
//   kside 0:  z[i,j] = sum_k A[i,k] x[k,j]            (I (x) A,        rho -> A rho)
//   kside 1:  z[i,j] = sum_k conj(A[j,k]) x[i,k]      (conj(A) (x) I,  rho -> rho A^dagger)
//   kside 2:  z[i,j] = sum_c (C_c x C_c^dagger)[i,j]  (sum conj(C_c) (x) C_c, the jump part)
// The right product reads row j of A, which is the same for all lanes of a warp when n is a
// multiple of 32 (uniform loads) and gathers x[i + k n] -- consecutive lanes, coalesced.
__device__ __forceinline__ double2 qb_rowdot_kron(const QbOpDev& A, int sl, int lane, long long r,
                                                  bool active, const double2* __restrict__ x)
{
    double2 acc = make_double2(0.0, 0.0);
    const int n = A.kn;
    const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val);
    if (A.kside == 0) {
        if (A.kval) {                       // n % 32 == 0: SELL sweep of A on column j of the state
            const int spc = n >> 5;         // slices per column
            const int j = sl / spc, isl = sl - j * spc;
            const double2* __restrict__ xc = x + (long long)j * n;
            const int s0 = A.slice_ptr[isl], w = A.slice_ptr[isl + 1] - s0;
            const double2* __restrict__ v = reinterpret_cast<const double2*>(A.kval) + ((size_t)s0 * 32 + lane);
            const int* __restrict__ c = A.kcol + ((size_t)s0 * 32 + lane);
            for (int k = 0; k < w; k += 4) {       // full predicated batches, no serial tail
                int cc[4];
                double2 vv[4], xx[4];
#pragma unroll
                for (int u = 0; u < 4; u++) cc[u] = (k + u < w) ? __ldg(c + (k + u) * 32) : 0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    vv[u] = (k + u < w) ? __ldg(v + (k + u) * 32) : make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < 4; u++) xx[u] = xc[cc[u]];
#pragma unroll
                for (int u = 0; u < 4; u++) qb_fma(acc, vv[u], xx[u]);
            }
        } else if (active) {
            const int j = (int)(r / n), i = (int)(r - (long long)j * n);
            const double2* __restrict__ xc = x + (long long)j * n;
            const int p1 = A.rowptr[i + 1];
            for (int p = A.rowptr[i]; p < p1; p++) qb_fma(acc, val[p], xc[A.col[p]]);
        }
    } else if (A.kside == 2) {
        // sandwich  z[i,j] = sum_c sum_{k,l} C_c[i,k] conj(C_c[j,l]) x[k,l]  (rho -> sum C rho C^dagger);
        // the C_c are stacked row-wise in one CSR: row c*n + i
        if (active) {
            const int j = (int)(r / n), i = (int)(r - (long long)j * n);
            const int nstack = A.kstack;
            for (int c = 0; c < nstack; c++) {
                const int* __restrict__ rp = A.rowptr + (long long)c * n;
                // row j is the same for the whole warp when n is a multiple of 32: an operator
                // with an empty row j is skipped before any per-lane load
                const int pj0 = __ldg(rp + j), pj1 = __ldg(rp + j + 1);
                if (pj0 == pj1) continue;
                const int pi0 = __ldg(rp + i), pi1 = __ldg(rp + i + 1);
                if (pi0 == pi1) continue;
                for (int q = pj0; q < pj1; q++) {
                    double2 b = __ldg(val + q);
                    b.y = -b.y;
                    const double2* __restrict__ xl = x + (long long)__ldg(A.col + q) * n;
                    double2 t = make_double2(0.0, 0.0);
                    for (int p = pi0; p < pi1; p++) qb_fma(t, __ldg(val + p), xl[__ldg(A.col + p)]);
                    qb_fma(acc, b, t);
                }
            }
        }
    } else if (active) {
        const int j = (int)(r / n), i = (int)(r - (long long)j * n);
        const double2* __restrict__ xr = x + i;
        const int p1 = __ldg(A.rowptr + j + 1);
        // full predicated batches of QB_KU entries: the (warp-uniform, L1-resident) column
        // indices first, then all gathers in flight, the values only when they are consumed
        for (int p = __ldg(A.rowptr + j); p < p1; p += QB_KU) {
            int cc[QB_KU];
            double2 xv[QB_KU];
#pragma unroll
            for (int u = 0; u < QB_KU; u++) cc[u] = (p + u < p1) ? __ldg(A.col + p + u) : 0;
#pragma unroll
            for (int u = 0; u < QB_KU; u++) xv[u] = xr[(long long)cc[u] * n];
#pragma unroll
            for (int u = 0; u < QB_KU; u++) {
                double2 a = (p + u < p1) ? __ldg(val + p + u) : make_double2(0.0, 0.0);
                a.y = -a.y;
                qb_fma(acc, a, xv[u]);
            }
        }
    }
    return acc;
}

// RSELL sweep (qb_types.h).  Per slot the descriptor gives the column rule (row + delta,
// row ^ delta, or an explicit block) and the value (one constant or an explicit block), so
// diagonal-structured operators issue no per-element index loads.
#ifndef QB_RS_U
#define QB_RS_U 2
#endif
__device__ __forceinline__ double2 qb_lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
// COH: x is re-written by other SMs while this kernel runs (the cooperative multi-round kernel),
// so its gathers must not take the non-coherent (ld.global.nc) path; every out-of-tile row is
// gathered by exactly one lane per slot, so bypassing L1 (ld.global.cg) costs no reuse
template <bool COH>
__device__ __forceinline__ double2 qb_ldx(const double2* p) {
    if (COH) {
        double2 v;
        asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    }
    return __ldg(p);
}
// generic form (stand-alone kernels, rare passes): (A x)[r] for the lane's row r of slice sl,
// descriptors and x from global memory
template <int U = QB_RS_U, bool COH = false>
__device__ __forceinline__ double2 qb_rowdot_rsell(const QbOpDev& A, int sl, int lane, int r,
                                                   const double2* __restrict__ x)
{
    double2 acc = make_double2(0.0, 0.0);
    const int4 si = __ldg(reinterpret_cast<const int4*>(A.sinfo) + sl);
    const int4* __restrict__ dsc = reinterpret_cast<const int4*>(A.sdesc);
    const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val) + ((long long)si.z * 32 + lane);
    const int* __restrict__ col = A.col + ((long long)si.w * 32 + lane);
    const int s1 = si.x + (si.y & 4095);
    auto slot = [&](int k, int& c, double2& v) {
        const int4 d = __ldg(dsc + 2 * k);
        const int cr = d.x & QB_RS_COL_MASK;
        c = (cr == QB_RS_COL_XOR) ? (r ^ d.y) : (r + d.y);
        if (cr == QB_RS_COL_EXPL) c = __ldg(col + d.z * 32);
        v = __ldg(reinterpret_cast<const double2*>(dsc + 2 * k + 1));
        if (!(d.x & QB_RS_VAL_CONST)) v = __ldg(val + d.w * 32);
    };
    int k = si.x;
    for (; k + U <= s1; k += U) {
        int cc[U];
        double2 vv[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; u++) slot(k + u, cc[u], vv[u]);
#pragma unroll
        for (int u = 0; u < U; u++) xx[u] = qb_ldx<COH>(x + cc[u]);
#pragma unroll
        for (int u = 0; u < U; u++) qb_fma(acc, vv[u], xx[u]);
    }
    for (; k < s1; k++) {
        int c; double2 v;
        slot(k, c, v);
        qb_fma(acc, v, qb_ldx<COH>(x + c));
    }
    return acc;
}

// Fused-pass form: RB slices that SHARE one descriptor list (same first descriptor and
// count; base[j] = the slice's explicit value / column block bases), descriptors from the
// constant bank (CD, kernel parameter) or global memory.  Rows [lo, lo + trows) of x were
// staged in shared memory (shared-window address sxa, by a TMA bulk copy): columns inside
// that window are gathered with conflict-free LDS.128 (no L1 tag stage / replays), the
// others from global memory.
struct QbTileElem { const int* sinfo; const qb_c128* val; const int* col; const QbSlotDesc* sdesc; };
template <int RB, bool CD, bool COH = false>
__device__ __forceinline__ void qb_rowdot_rsell_tile(const QbTileElem& A, const QbSlotDesc* __restrict__ cdesc,
                                                     int dstart, int dcount, const int (&vb)[RB],
                                                     const int (&cb)[RB], const int (&r)[RB], int lane,
                                                     const double2* __restrict__ x, unsigned sxa,
                                                     int lo, int trows, int tnom, bool pow2, double2 (&acc)[RB])
{
#pragma unroll
    for (int j = 0; j < RB; j++) acc[j] = make_double2(0.0, 0.0);
    const QbSlotDesc* __restrict__ dsc = CD ? cdesc + dstart : A.sdesc + dstart;
    const int ntot = dcount & 4095, nfast = (dcount >> 12) & 4095;
    const int nxor = pow2 ? (dcount >> 24) & 255 : 0;
    auto fetch = [&](int k, int& rule, int& delta, int& cpos, int& vpos, double2& cv) {
        if (CD) {
            rule = dsc[k].rule; delta = dsc[k].delta; cpos = dsc[k].cpos; vpos = dsc[k].vpos;
            cv = make_double2(dsc[k].vre, dsc[k].vim);
        } else {
            const int4 d = __ldg(reinterpret_cast<const int4*>(dsc + k));
            cv = __ldg(reinterpret_cast<const double2*>(dsc + k) + 1);
            rule = d.x; delta = d.y; cpos = d.z; vpos = d.w;
        }
    };
    auto gather = [&](int c) -> double2 {
        const unsigned o = (unsigned)(c - lo);
        if (o < (unsigned)trows) return qb_lds128(sxa + o * 16u);
        return qb_ldx<COH>(x + c);
    };
    // xor slots with a constant value on a power-of-two tile: the partner row is at the own
    // tile offset ^ (delta * 16) when delta < tile rows (warp-uniform branch), else in global
    // memory -- no per-lane bounds test, no select between two loads
    unsigned sa[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) sa[j] = (unsigned)(r[j] - lo) * 16u;
    int k = 0;
#ifdef QB_XSHFL
    double2 own[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) own[j] = nxor > 0 ? qb_lds128(sxa + sa[j]) : make_double2(0.0, 0.0);
#endif
    auto xslot = [&](int kk, double2 (&a)[RB]) {
        int delta;
        double2 cv, xv[RB];
        if (CD) { delta = dsc[kk].delta; cv = make_double2(dsc[kk].vre, dsc[kk].vim); }
        else {
            delta = __ldg(&dsc[kk].delta);
            cv = __ldg(reinterpret_cast<const double2*>(dsc + kk) + 1);
        }
#ifdef QB_XSHFL     // partner rows inside the warp's own 32 rows: warp shuffles instead of gathers
        if (delta < 32) {
#pragma unroll
            for (int j = 0; j < RB; j++) {
                xv[j].x = __shfl_xor_sync(0xffffffffu, own[j].x, delta);
                xv[j].y = __shfl_xor_sync(0xffffffffu, own[j].y, delta);
            }
        } else
#endif
        if (delta < tnom) {        // tnom: nominal (power-of-two) tile rows
#pragma unroll
            for (int j = 0; j < RB; j++) xv[j] = qb_lds128(sxa + (sa[j] ^ ((unsigned)delta << 4)));
        } else {
#pragma unroll
            for (int j = 0; j < RB; j++) xv[j] = qb_ldx<COH>(x + (r[j] ^ delta));
        }
#ifdef QB_XIMAG     // constants of Hamiltonian parts are purely imaginary (-i H): two DFMA instead of four
        if (cv.x == 0.0) {             // warp-uniform (constant bank / uniform load)
#pragma unroll
            for (int j = 0; j < RB; j++) {
                a[j].x = fma(-cv.y, xv[j].y, a[j].x); a[j].y = fma(cv.y, xv[j].x, a[j].y);
            }
        } else
#endif
#pragma unroll
        for (int j = 0; j < RB; j++) qb_fma(a[j], cv, xv[j]);
    };
#ifdef QB_XPF       // measured neutral on C3 (tools/tile_ab.sh): off by default
    // the partner rows outside the tile (the largest deltas, processed last) are L2 hits a few
    // hundred cycles away: request them into L1 before the in-tile slots are swept
    for (int kk = nxor - 1; kk >= 0; kk--) {
        const int delta = CD ? dsc[kk].delta : __ldg(&dsc[kk].delta);
        if (delta < tnom) break;
#pragma unroll
        for (int j = 0; j < RB; j++)
            asm volatile("prefetch.global.L1 [%0];" :: "l"(x + (r[j] ^ delta)));
    }
#endif
    // two independent accumulator sets halve the dependent DFMA chain of the sweep
    double2 acc2[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) acc2[j] = make_double2(0.0, 0.0);
#ifdef QB_XACC2     // measured neutral on C3 (tools/tile_ab.sh): off by default
    for (; k + 2 <= nxor; k += 2) { xslot(k, acc); xslot(k + 1, acc2); }
#endif
    for (; k < nxor; k++) xslot(k, acc);
#pragma unroll
    for (int j = 0; j < RB; j++) { acc[j].x += acc2[j].x; acc[j].y += acc2[j].y; }
    // other fast slots: column = (row ^ xd) + ad, one constant value
    for (; k < nfast; k++) {
        int ru, de, cp, vp;
        double2 cv;
        fetch(k, ru, de, cp, vp, cv);
        const bool isx = (ru & QB_RS_COL_MASK) == QB_RS_COL_XOR;
        const int xd = isx ? de : 0, ad = isx ? 0 : de;
#pragma unroll
        for (int j = 0; j < RB; j++) qb_fma(acc[j], cv, gather((r[j] ^ xd) + ad));
    }
    // general slots: explicit column and / or value blocks
    const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val) + lane;
    const int* __restrict__ col = A.col + lane;
    for (; k < ntot; k++) {
        int rule, delta, cpos, vpos;
        double2 cv;
        fetch(k, rule, delta, cpos, vpos, cv);
        const int cr = rule & QB_RS_COL_MASK;
        int c[RB];
        double2 v[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) {
            c[j] = (cr == QB_RS_COL_XOR) ? (r[j] ^ delta) : (r[j] + delta);
            if (cr == QB_RS_COL_EXPL) c[j] = __ldg(col + (long long)(cb[j] + cpos) * 32);
            v[j] = cv;
            if (!(rule & QB_RS_VAL_CONST)) v[j] = __ldg(val + (long long)(vb[j] + vpos) * 32);
        }
#pragma unroll
        for (int j = 0; j < RB; j++) qb_fma(acc[j], v[j], gather(c[j]));
    }
}

// (A x)[r] for the lane's row r of slice sl, any format.  `active` lanes have r < nrows.
template <int U = QB_U1, bool COH = false>
__device__ __forceinline__ double2 qb_rowdot(const QbOpDev& A, int sl, int lane, long long r,
                                             bool active, const double2* __restrict__ x)
{
    double2 acc = make_double2(0.0, 0.0);
    if (A.fmt == QB_FMT_DIAM) {
        const double2* xs[1] = {x};
        double2 a1[1] = {acc};
        qb_rowdot_diam<1, U>(A, sl, lane, r, xs, a1);
        acc = a1[0];
    } else if (A.fmt == QB_FMT_SELL) {
        const int s0 = A.slice_ptr[sl], w = A.slice_ptr[sl + 1] - s0;
        const double2* __restrict__ v = reinterpret_cast<const double2*>(A.val) + ((size_t)s0 * 32 + lane);
        const int* __restrict__ c = A.col + ((size_t)s0 * 32 + lane);
        int k = 0;
        for (; k + 4 <= w; k += 4) {          // 4 slots in flight: 12 coalesced loads per lane
            const int c0 = __ldg(c + k * 32), c1 = __ldg(c + (k + 1) * 32),
                      c2 = __ldg(c + (k + 2) * 32), c3 = __ldg(c + (k + 3) * 32);
            const double2 v0 = __ldg(v + k * 32), v1 = __ldg(v + (k + 1) * 32),
                          v2 = __ldg(v + (k + 2) * 32), v3 = __ldg(v + (k + 3) * 32);
            const double2 x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
            qb_fma(acc, v0, x0); qb_fma(acc, v1, x1); qb_fma(acc, v2, x2); qb_fma(acc, v3, x3);
        }
        for (; k < w; k++) qb_fma(acc, __ldg(v + k * 32), x[__ldg(c + k * 32)]);
    } else if (A.fmt == QB_FMT_CSR) {
        if (active) {
            const double2* __restrict__ val = reinterpret_cast<const double2*>(A.val);
            const int p1 = A.rowptr[r + 1];
            for (int p = A.rowptr[r]; p < p1; p++) qb_fma(acc, val[p], x[A.col[p]]);
        }
    } else if (A.fmt == QB_FMT_KRON) {
        acc = qb_rowdot_kron(A, sl, lane, r, active, x);
    } else if (A.fmt == QB_FMT_RSELL) {
        acc = qb_rowdot_rsell<QB_RS_U, COH>(A, sl, lane, (int)r, x);
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

// SELL-only row product (hot body of the fused pass kernel), U slots in flight
#ifndef QB_HOT_U
#define QB_HOT_U 4
#endif
template <int U = QB_HOT_U>
__device__ __forceinline__ double2 qb_rowdot_sell(const QbOpDev& A, int sl, int lane,
                                                  const double2* __restrict__ x)
{
    double2 acc = make_double2(0.0, 0.0);
    const int s0 = A.slice_ptr[sl], w = A.slice_ptr[sl + 1] - s0;
    const double2* __restrict__ v = reinterpret_cast<const double2*>(A.val) + ((size_t)s0 * 32 + lane);
    const int* __restrict__ c = A.col + ((size_t)s0 * 32 + lane);
    int k = 0;
    for (; k + U <= w; k += U) {
        int cc[U];
        double2 vv[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; u++) cc[u] = __ldg(c + (k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; u++) vv[u] = __ldg(v + (k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; u++) xx[u] = x[cc[u]];
#pragma unroll
        for (int u = 0; u < U; u++) qb_fma(acc, vv[u], xx[u]);
    }
    for (; k < w; k++) qb_fma(acc, __ldg(v + k * 32), x[__ldg(c + k * 32)]);
    return acc;
}

// the same sweep for G states at once: every operator slot is loaded once and applied to G
// gathered vectors (register blocking over trajectories)
#ifndef QB_HOT_UG
#define QB_HOT_UG 2
#endif
template <int G, int U = QB_HOT_UG>
__device__ __forceinline__ void qb_rowdot_sell_multi(const QbOpDev& A, int sl, int lane,
                                                     const double2* const (&x)[G], double2 (&acc)[G])
{
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = make_double2(0.0, 0.0);
    const int s0 = A.slice_ptr[sl], w = A.slice_ptr[sl + 1] - s0;
    const double2* __restrict__ v = reinterpret_cast<const double2*>(A.val) + ((size_t)s0 * 32 + lane);
    const int* __restrict__ c = A.col + ((size_t)s0 * 32 + lane);
    int k = 0;
    for (; k + U <= w; k += U) {
        int cc[U];
        double2 vv[U], xx[U][G];
#pragma unroll
        for (int u = 0; u < U; u++) cc[u] = __ldg(c + (k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; u++) vv[u] = __ldg(v + (k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int g = 0; g < G; g++) xx[u][g] = x[g][cc[u]];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int g = 0; g < G; g++) qb_fma(acc[g], vv[u], xx[u][g]);
    }
    for (; k < w; k++) {
        const int c0 = __ldg(c + k * 32);
        const double2 v0 = __ldg(v + k * 32);
#pragma unroll
        for (int g = 0; g < G; g++) qb_fma(acc[g], v0, x[g][c0]);
    }
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
