// qb_engine.cu -- fused evolution engine: pass kernel (vector unit), control kernel
// (scalar unit, qb_control.h) and the host driver behind the C ABI.
//
// Data layout in HBM
//   pool[nslots][V][N] complex128 : per trajectory slot V = S + 5 state-sized vectors:
//       k[0..S-1] (RK stage derivatives), y_prev, y_front, y_interp, tmpA, tmpB.
//       "y_prev <- y_front" etc. are slot relabels in QbTraj, never copies.
//   pass[nslots], traj[nslots]    : the next vector instruction and the controller state
//   partials[nslots][nslices][red_stride] : per-warp partial reductions, summed in a fixed order
//   linmap[nslots]                : weights of multi-output passes (Adams prediction / update)
// A "round" = pass kernel (+ qb_linmap_kernel for Adams engines, + the partial-sum reduction
// for N > 65536) + control kernel; the host enqueues rounds back to back (CUDA graphs of 16
// rounds) and only looks at a device counter every chunk: no host synchronisation per step.
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "qb_host.h"
#include "qb_kernels.cuh"
#include "qb_tableaux.h"
#include <cooperative_groups.h>

// the tableaux (and the generated Adams coefficient table) live in constant memory: the controller's single active lane reads them
// through the constant cache instead of serial global loads
__constant__ QbTableau c_tabs[4];

struct QbEngineDev {
    QbCtl ctl;
    int tableau_id, pad0_;
    QbOpDev elem[QB_MAX_ELEMS];
    const QbOpDev* cops;
    const QbOpDev* nops;
    const QbOpDev* eops;
    double2* pool;
    int V, nslots;
    const double2* init_states;
    const int* init_map;
    double2* out_states;
    QbTraj* traj;
    QbPass* pass;
    QbLinMap* linmap;      // [nslots] weights of LINMAP passes (Adams prediction / update)
    qb_c128* coef;
    double* probs;
    double* partials;
    int* queue_head;       // next trajectory id to start
    int* n_active;         // slots still working
    int* work;             // work counter of the persistent tile kernel (reset here every round)
    int* act_list;         // slots whose pass the tile kernel executes this round, in slot order
    int* act_count;
    int ntraj_total;
    int mode;
    int* out_status;
    int* out_stats;
    unsigned long long* vec_count;   // state-sized vector accesses issued (algorithmic traffic)
    int nslices, red_stride;
    int all_sell, all_lean;     // every RHS element is SELL / is SELL, DIAM or KRON (lean pass body)
    int all_rsell, pad1_;       // every RHS element is RSELL: tile-staged pass kernel
    double* red_final;          // [nslots][QB_MAXRED]: pre-reduced partials (large systems) or null
    // dense batched path (qb_dense.cu): z of every slot precomputed by one DMMA ZGEMM
    double2* zbuf;              // [nslots][N] or null
    const double2** xcols;      // [nslots] column pointers of the GEMM's right operand
    double2** zcols;            // [nslots]
};

int qb_launch_dense_rhs(cudaStream_t stream, const qb_c128* A, int N, const void* const* xcols,
                        void* const* zcols, int ncols);

// ------------------------------------------------------------------ pass kernel
__device__ __forceinline__ const double2* qb_vsrc(const QbEngineDev* E, int slot, int idx,
                                                  int init_idx) {
    if (idx >= 0) return E->pool + ((size_t)slot * E->V + idx) * (size_t)E->ctl.N;
    return E->init_states + (size_t)init_idx * (size_t)E->ctl.N;
}

#ifndef QB_GRAPH_ROUNDS
#define QB_GRAPH_ROUNDS 16   // rounds per captured CUDA graph
#endif
#ifndef QB_NO_STREAM_VEC
// epilogue sources are read once and k / y outputs are written once per pass: streaming
// cache policy keeps them from evicting the operator and the gathered state in L1/L2
#define QB_LDV(p) __ldcs(p)
#define QB_STV(p, v) __stcs((p), (v))
#else
#define QB_LDV(p) (*(p))
#define QB_STV(p, v) (*(p) = (v))
#endif
#ifndef QB_PF
#define QB_PF 4      // epilogue source vectors prefetched before the operator sweep
#endif
#ifndef QB_UL
#define QB_UL 4      // DIAM entries in flight per warp in the lean body (HBM streaming)
#endif
#ifndef QB_MINB
#define QB_MINB 4
#endif
#define QB_RED_CTAS 32
#ifndef QB_COOP_MAX_SLOTS
#define QB_COOP_MAX_SLOTS 63     // the cooperative form serves runs below the compaction threshold
#endif
#define QB_SH_MAXW 40   // widest SELL slice the shared kernel stages (40*32*20 B = 25.6 KB)
#ifndef QB_SH_T
#define QB_SH_T 4       // consecutive slices one CTA of the shared kernel walks through
#endif

// super-operator H mcsolve (trn = n > 0): part[3] = sum over the diagonal rows of rho of Re o1,
// part[4] the same of z -- tr(rho) replaces ||psi||^2 in the jump logic (mcsolve.py:311-319)
__device__ __forceinline__ void qb_trace_partials(int trn, long long r, bool active, double2 o1, double2 z,
                                                  int lane, double* __restrict__ part)
{
    if (!trn) return;                                     // warp-uniform
    const bool diag = active && (r % (trn + 1) == 0);
    const double t1 = qb_warp_sum(diag ? o1.x : 0.0), t2 = qb_warp_sum(diag ? z.x : 0.0);
    if (lane == 0) { part[3] = t1; part[4] = t2; }
}

// per-warp view of a slot's pass descriptor (read straight from L1/L2-resident global memory)
struct QbWarpHdr {
    int kind, nsrc, zdst, dst1, red, xslot, my_src;
    double my_w1, my_w2, zscale, w1z, w2z;
    const double2* slot_base;
    const double2* init_ptr;
};

__device__ __forceinline__ void qb_load_hdr(const QbEngineDev* __restrict__ E, int slot, int lane,
                                            QbWarpHdr& h)
{
    const QbPass* __restrict__ gp = &E->pass[slot];
    const size_t N_ = (size_t)E->ctl.N;
    h.kind = gp->kind; h.nsrc = gp->nsrc; h.zdst = gp->zdst; h.dst1 = gp->dst1; h.red = gp->red;
    h.xslot = gp->x;
    h.my_src = 0; h.my_w1 = 0.0; h.my_w2 = 0.0;
    if (lane < h.nsrc) { h.my_src = gp->sw[lane].src; h.my_w1 = gp->sw[lane].w1; h.my_w2 = gp->w2[lane]; }
    h.zscale = gp->zscale; h.w1z = gp->w1z; h.w2z = gp->w2z;
    h.slot_base = E->pool + (size_t)slot * E->V * N_;
    h.init_ptr = E->init_states + (size_t)E->traj[slot].init_idx * N_;
}

// One 32-row slice of one trajectory slot, executed by one warp without any block-level
// synchronisation: epilogue sources are requested BEFORE the operator sweep so their HBM
// latency overlaps it; partial reductions go to partials[slot][slice][k].
// sval/scol != nullptr: the slice of the (single, SELL) RHS operator is staged in shared
// memory by the caller and shared by the CTA's warps (= 8 trajectories).
template <bool COH = false>
__device__ __forceinline__ void qb_pass_one_slice(
    const QbEngineDev* __restrict__ E, int slot, int sl, int lane, const QbWarpHdr& h,
    const double2* __restrict__ sval, const int* __restrict__ scol, int sw)
{
    const QbPass* __restrict__ gp = &E->pass[slot];
    const int N = E->ctl.N;
    const size_t N_ = (size_t)N;
    const long long r = (long long)sl * 32 + lane;
    const bool active = r < N;
    const int kind = h.kind, nsrc = h.nsrc;
#define QB_VS(idx) ((idx) >= 0 ? h.slot_base + (long long)(idx) * N : h.init_ptr)
    double* __restrict__ part = E->partials + ((size_t)slot * E->nslices + sl) * E->red_stride;

    if (kind == QB_PASS_EXPECT) {
        const int opset = gp->opset, op_lo = gp->op_lo, nops = gp->op_hi - gp->op_lo;
        const QbOpDev* ops = (opset == QB_OPSET_EOPS) ? E->eops : E->nops;
        const bool functional = (opset == QB_OPSET_EOPS) && E->ctl.eop_functional;
        // super-operator H (mcsolve.py:481-490): probabilities are tr(n_k rho), the diagonal
        // rows of the column-stacked n_k rho
        const int trn = (opset == QB_OPSET_NOPS) ? E->ctl.mc_trace : 0;
        const double2* x = QB_VS(h.xslot);
        double2 xr = make_double2(0.0, 0.0);
        if (active) xr = x[r];
        for (int m = 0; m < nops; m++) {
            const double2 q = qb_rowdot<QB_UP, COH>(ops[op_lo + m], sl, lane, r, active, x);
            double2 pr;
            if (functional) pr = q;
            else if (trn) pr = (r % (trn + 1) == 0) ? q : make_double2(0.0, 0.0);
            else pr = make_double2(xr.x * q.x + xr.y * q.y, xr.x * q.y - xr.y * q.x);  // conj(x)*q
            if (!active) pr = make_double2(0.0, 0.0);
            const double sre = qb_warp_sum(pr.x), sim = qb_warp_sum(pr.y);
            if (lane == 0) { part[2 * m] = sre; part[2 * m + 1] = sim; }
        }
        return;
    }

    // ---- epilogue sources: lane i holds (slot, w1, w2) of source i; prefetch the first QB_PF
#if QB_PF > 0
    double2 pv[QB_PF];
#pragma unroll
    for (int u = 0; u < QB_PF; u++) {
        const int sidx = __shfl_sync(0xffffffffu, h.my_src, u);
        const double2* p = QB_VS(sidx);
        pv[u] = (u < nsrc && active) ? QB_LDV(p + r) : make_double2(0.0, 0.0);
    }
#endif

    // ---- operator application ----
    double2 z = make_double2(0.0, 0.0);
    if (kind == QB_PASS_RHS) {
        const double2* x = QB_VS(h.xslot);
        const int nelem = E->ctl.nelem;
        const qb_c128* cf = E->coef + (size_t)slot * E->ctl.maxcoef;
        if (sval) {                // SELL slice staged in shared memory, shared by 8 slots
            double2 q = make_double2(0.0, 0.0);
            const double2* v = sval + lane;
            const int* c = scol + lane;
            int k = 0;
            for (; k + 4 <= sw; k += 4) {
                const int c0 = c[k * 32], c1 = c[(k + 1) * 32], c2 = c[(k + 2) * 32], c3 = c[(k + 3) * 32];
                const double2 x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
                qb_fma(q, v[k * 32], x0); qb_fma(q, v[(k + 1) * 32], x1);
                qb_fma(q, v[(k + 2) * 32], x2); qb_fma(q, v[(k + 3) * 32], x3);
            }
            for (; k < sw; k++) qb_fma(q, v[k * 32], x[c[k * 32]]);
            const qb_c128 cc = cf[0];
            z.x = cc.re * q.x - cc.im * q.y;
            z.y = cc.re * q.y + cc.im * q.x;
        } else if (E->zbuf) {      // dense batched path: A x was computed by the ZGEMM pre-pass
            const double2 q = active ? E->zbuf[(size_t)slot * N_ + r] : make_double2(0.0, 0.0);
            const qb_c128 c = cf[0];
            z.x = c.re * q.x - c.im * q.y;
            z.y = c.re * q.y + c.im * q.x;
        } else
        for (int e = 0; e < nelem; e++) {
            const double2 q = qb_rowdot<QB_UP, COH>(E->elem[e], sl, lane, r, active, x);
            const qb_c128 c = cf[e];
            z.x += c.re * q.x - c.im * q.y;
            z.y += c.re * q.y + c.im * q.x;
        }
    } else if (kind == QB_PASS_APPLY) {
        const double2 q = qb_rowdot<QB_UP, COH>(E->cops[gp->op_lo], sl, lane, r, active, QB_VS(h.xslot));
        const qb_c128 c = E->coef[(size_t)slot * E->ctl.maxcoef];
        z = make_double2(c.re * q.x - c.im * q.y, c.re * q.y + c.im * q.x);
    }
    z.x *= h.zscale; z.y *= h.zscale;

    // ---- fused linear combinations (sources in order, z last), stores, reductions ----
    double2 o1 = make_double2(0.0, 0.0), o2 = make_double2(0.0, 0.0);
#if QB_PF > 0
#pragma unroll
    for (int u = 0; u < QB_PF; u++) {
        const double a = __shfl_sync(0xffffffffu, h.my_w1, u), b = __shfl_sync(0xffffffffu, h.my_w2, u);
        o1.x = fma(a, pv[u].x, o1.x); o1.y = fma(a, pv[u].y, o1.y);
        o2.x = fma(b, pv[u].x, o2.x); o2.y = fma(b, pv[u].y, o2.y);
    }
#endif
    for (int i = QB_PF; i < nsrc; i += 4) {            // rarely taken (dense-output rows)
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int sidx = __shfl_sync(0xffffffffu, h.my_src, (i + u) & 31);
            const double2* p = QB_VS(sidx);
            v[u] = (i + u < nsrc && active) ? QB_LDV(p + r) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const double a = __shfl_sync(0xffffffffu, h.my_w1, (i + u) & 31);
            const double b = __shfl_sync(0xffffffffu, h.my_w2, (i + u) & 31);
            o1.x = fma(a, v[u].x, o1.x); o1.y = fma(a, v[u].y, o1.y);
            o2.x = fma(b, v[u].x, o2.x); o2.y = fma(b, v[u].y, o2.y);
        }
    }
    o1.x = fma(h.w1z, z.x, o1.x); o1.y = fma(h.w1z, z.y, o1.y);
    o2.x = fma(h.w2z, z.x, o2.x); o2.y = fma(h.w2z, z.y, o2.y);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    if (active) {
        double2* base = const_cast<double2*>(h.slot_base);
        if (h.zdst >= 0) QB_STV(base + (size_t)h.zdst * N_ + r, z);
        if (h.dst1 >= 0) QB_STV(base + (size_t)h.dst1 * N_ + r, o1);
        else if (h.dst1 == QB_SLOT_OUT)
            E->out_states[((size_t)E->traj[slot].traj_id * E->ctl.nt + gp->out_index) * N_ + r] = o1;
        const double n1 = o1.x * o1.x + o1.y * o1.y;
        r0 = n1;
        if (h.red & QB_RED_WRMS) {
            const double q = sqrt(o2.x * o2.x + o2.y * o2.y)
                             / (E->ctl.opt.atol + E->ctl.opt.rtol * sqrt(n1));
            r1 = q * q;
        }
        r2 = z.x * z.x + z.y * z.y;
    }
    if (h.red) {
        r0 = qb_warp_sum(r0); r1 = qb_warp_sum(r1); r2 = qb_warp_sum(r2);
        if (lane == 0) { part[0] = r0; part[1] = r1; part[2] = r2; }
        qb_trace_partials(E->ctl.mc_trace, r, active, o1, z, lane, part);
    }
#undef QB_VS
}

// The common case -- COMBINE passes and RHS passes whose elements are all SELL -- as a lean
// separate body: nothing but the operand pointer is fetched before the operator sweep, the
// pass descriptor is read after it (it is L1/L2 resident), so no value other than the
// accumulator lives across the sweep and the 64-register budget holds without spilling.
// Everything else (EXPECT, APPLY, DIAM / CSR / dense elements) takes qb_pass_slice_generic.
__device__ __forceinline__ const double2* qb_hot_x(const QbEngineDev* __restrict__ E, int slot)
{
    // (slot * V + vector) fits 32 bits: ONE widening multiply per address instead of 64-bit
    // multiply chains (address arithmetic was a quarter of the epilogue's instructions)
    const int xs = E->pass[slot].x;
    if (xs >= 0) return E->pool + (long long)(slot * E->V + xs) * E->ctl.N;
    return E->init_states + (long long)E->traj[slot].init_idx * E->ctl.N;
}

// fused linear combinations (sources in order, z last), stores, reductions of one slot
__device__ __forceinline__ void qb_hot_epilogue(const QbEngineDev* __restrict__ E, int slot,
                                                int sl, int lane, double2 z)
{
    const QbPass* __restrict__ gp = &E->pass[slot];
    const int N = E->ctl.N;
    const long long r = (long long)sl * 32 + lane;
    const bool active = r < N;
    const int nsrc = gp->nsrc;
    const int red = gp->red;
    const bool werr = (red & QB_RED_WRMS) != 0;      // o2 (the error combination) is only
    double2* const pool_r = E->pool + r;             // consumed by WRMS
    const int vbase = slot * E->V;
    double2 o1 = make_double2(0.0, 0.0), o2 = make_double2(0.0, 0.0);
    for (int i = 0; i < nsrc; i += 4) {
        double2 v[4];
        double a[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            // (slot, weight) of source i + u in one uniform 16-byte load
            const int4 sw = *reinterpret_cast<const int4*>(&gp->sw[min(i + u, QB_MAXSRC - 1)]);
            a[u] = __hiloint2double(sw.w, sw.z);
            const double2* p = pool_r + (long long)(vbase + max(sw.x, 0)) * N;
            if (sw.x < 0)                                 // the initial-state buffer (first pass)
                p = E->init_states + ((long long)E->traj[slot].init_idx * N + r);
            v[u] = (i + u < nsrc && active) ? QB_LDV(p) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { o1.x = fma(a[u], v[u].x, o1.x); o1.y = fma(a[u], v[u].y, o1.y); }
        if (werr) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double b = gp->w2[min(i + u, QB_MAXSRC - 1)];
                o2.x = fma(b, v[u].x, o2.x); o2.y = fma(b, v[u].y, o2.y);
            }
        }
    }
    {
        const double2 hz = *reinterpret_cast<const double2*>(&gp->w1z);  // w1z, w2z
        o1.x = fma(hz.x, z.x, o1.x); o1.y = fma(hz.x, z.y, o1.y);
        o2.x = fma(hz.y, z.x, o2.x); o2.y = fma(hz.y, z.y, o2.y);
    }
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    if (active) {
        const int zdst = gp->zdst, dst1 = gp->dst1;
        if (zdst >= 0) QB_STV(pool_r + (long long)(vbase + zdst) * N, z);
        if (dst1 >= 0) QB_STV(pool_r + (long long)(vbase + dst1) * N, o1);
        else if (dst1 == QB_SLOT_OUT)
            E->out_states[((size_t)E->traj[slot].traj_id * E->ctl.nt + gp->out_index) * (size_t)N + r] = o1;
        const double n1 = o1.x * o1.x + o1.y * o1.y;
        r0 = n1;
        if (red & QB_RED_WRMS) {
            const double q = sqrt(o2.x * o2.x + o2.y * o2.y)
                             / (E->ctl.opt.atol + E->ctl.opt.rtol * sqrt(n1));
            r1 = q * q;
        }
        r2 = z.x * z.x + z.y * z.y;
    }
    if (red) {
        r0 = qb_warp_sum(r0); r1 = qb_warp_sum(r1); r2 = qb_warp_sum(r2);
        if (lane == 0) {
            double* __restrict__ part = E->partials + (long long)(slot * E->nslices + sl) * E->red_stride;
            part[0] = r0; part[1] = r1; part[2] = r2;
        }
        if (E->ctl.mc_trace)
            qb_trace_partials(E->ctl.mc_trace, r, active, o1, z, lane,
                              E->partials + (long long)(slot * E->nslices + sl) * E->red_stride);
    }
}

__device__ __forceinline__ void qb_pass_slice_hot(const QbEngineDev* __restrict__ E, int slot,
                                                  int sl, int lane, int kind)
{
    double2 z = make_double2(0.0, 0.0);
    if (kind == QB_PASS_RHS) {
        const double2* x = qb_hot_x(E, slot);
        const int nelem = E->ctl.nelem;
        for (int e = 0; e < nelem; e++) {
            double2 q;
            if (E->elem[e].fmt == QB_FMT_SELL) {
                q = qb_rowdot_sell(E->elem[e], sl, lane, x);
            } else if (E->elem[e].fmt == QB_FMT_KRON) {   // matrix-free Lindblad products
                const long long rr = (long long)sl * 32 + lane;
                q = qb_rowdot_kron(E->elem[e], sl, lane, rr, rr < E->ctl.N, x);
            } else {                       // DIAM (HBM-streamed operators)
                const double2* xs1[1] = {x};
                double2 a1[1] = {make_double2(0.0, 0.0)};
                qb_rowdot_diam<1, QB_UL>(E->elem[e], sl, lane, (long long)sl * 32 + lane, xs1, a1);
                q = a1[0];
            }
            const qb_c128 c = E->coef[(size_t)slot * E->ctl.maxcoef + e];
            z.x += c.re * q.x - c.im * q.y;
            z.y += c.re * q.y + c.im * q.x;
        }
        const double zs = E->pass[slot].zscale;
        z.x *= zs; z.y *= zs;
    }
    qb_hot_epilogue(E, slot, sl, lane, z);
}

// QB_G consecutive slots whose passes are all SELL RHS passes: one warp sweeps the slice ONCE
// for all of them (operator values / columns loaded once, G gathered states), then runs each
// slot's own epilogue.
template <int G>
__device__ __forceinline__ void qb_pass_slice_hot_multi(const QbEngineDev* __restrict__ E, int slot0,
                                                        int sl, int lane)
{
    const double2* x[G];
    double2 z[G];
#pragma unroll
    for (int g = 0; g < G; g++) { x[g] = qb_hot_x(E, slot0 + g); z[g] = make_double2(0.0, 0.0); }
    const int nelem = E->ctl.nelem;
    for (int e = 0; e < nelem; e++) {
        double2 q[G];
        qb_rowdot_sell_multi<G>(E->elem[e], sl, lane, x, q);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const qb_c128 c = E->coef[(size_t)(slot0 + g) * E->ctl.maxcoef + e];
            z[g].x += c.re * q[g].x - c.im * q[g].y;
            z[g].y += c.re * q[g].y + c.im * q[g].x;
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        const double zs = E->pass[slot0 + g].zscale;
        z[g].x *= zs; z[g].y *= zs;
        qb_hot_epilogue(E, slot0 + g, sl, lane, z[g]);
    }
}

template <bool COH = false>
static __device__ __noinline__ void qb_pass_slice_generic(const QbEngineDev* E, int slot, int sl,
                                                          int lane)
{
    QbWarpHdr h;
    qb_load_hdr(E, slot, lane, h);
    qb_pass_one_slice<COH>(E, slot, sl, lane, h, nullptr, nullptr, 0);
}

#ifndef QB_G
#define QB_G 4       // trajectory slots sharing one operator sweep (register blocking)
#endif

// Warp-autonomous pass kernel: one warp = one 32-row slice of QB_G consecutive trajectory
// slots; no shared memory, no block-level barrier.
__global__ void __launch_bounds__(QB_TILE_ROWS, QB_MINB)
qb_pass_kernel(const QbEngineDev* __restrict__ E, int nslots_used)
{
    const int ntiles = E->ctl.ntiles;
    const int grp = blockIdx.x / ntiles;
    const int tile = blockIdx.x - grp * ntiles;
    const int slot0 = grp * QB_G;
    int kinds[QB_G];
    bool any = false, all_rhs = E->all_sell != 0;
#pragma unroll
    for (int g = 0; g < QB_G; g++) {
        kinds[g] = (slot0 + g < nslots_used) ? E->pass[slot0 + g].kind : QB_PASS_NONE;
        if (kinds[g] == QB_PASS_LINMAP) kinds[g] = QB_PASS_NONE;      // qb_linmap_kernel's
        any |= kinds[g] != QB_PASS_NONE;
        all_rhs &= kinds[g] == QB_PASS_RHS;
    }
    if (!any) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = tile * (QB_TILE_ROWS / 32) + warp;
    if ((long long)sl * 32 >= E->ctl.N) return;          // warp-uniform
    if (QB_G > 1 && all_rhs) {
        qb_pass_slice_hot_multi<QB_G>(E, slot0, sl, lane);
        return;
    }
#if QB_G == 4
    // not all four in an RHS pass: still share the sweep inside each pair that is
    const bool p0 = E->all_sell && kinds[0] == QB_PASS_RHS && kinds[1] == QB_PASS_RHS;
    const bool p1 = E->all_sell && kinds[2] == QB_PASS_RHS && kinds[3] == QB_PASS_RHS;
    if (p0) { qb_pass_slice_hot_multi<2>(E, slot0, sl, lane); kinds[0] = kinds[1] = QB_PASS_NONE; }
    if (p1) { qb_pass_slice_hot_multi<2>(E, slot0 + 2, sl, lane); kinds[2] = kinds[3] = QB_PASS_NONE; }
#endif
#pragma unroll
    for (int g = 0; g < QB_G; g++) {
        const int kind = kinds[g];
        if (kind == QB_PASS_NONE) continue;
        if (kind == QB_PASS_COMBINE || (kind == QB_PASS_RHS && E->all_lean))
            qb_pass_slice_hot(E, slot0 + g, sl, lane, kind);
        else
            qb_pass_slice_generic(E, slot0 + g, sl, lane);
    }
}

// ------------------------------------------------------------------ tile-staged pass kernel
// For systems whose RHS elements are all RSELL.  One CTA = one tile of `trows` consecutive
// rows of ONE trajectory slot; its 8 warps walk the tile's 32-row slices.
//  * The tile's rows of the operand state x are staged in shared memory by ONE TMA bulk copy
//    (cp.async.bulk + mbarrier, issued by thread 0): every gather whose column falls inside
//    the tile is a conflict-free LDS.128 instead of an L1 lookup (the warp-autonomous kernel
//    above is bound by L1 wavefronts: 4 per gathered 512 bytes, replayed); columns outside
//    the tile are gathered from global memory (L2 hits: the slot's other tiles run at the
//    same time).  The operator costs two warp-uniform loads per slot (RSELL descriptors).
//  * The first QB_TP epilogue sources of a slice are requested BEFORE the sweep and held in
//    registers (the shared-memory sweep needs few), so their HBM latency overlaps the sweep
//    instead of following it in dependent batches -- the latency chain per slice is
//    max(HBM, sweep) instead of their sum, which is what bounded the old kernel's bytes in
//    flight.
// Partial reductions, stores and every other pass kind are the same as in qb_pass_kernel.
#ifndef QB_TT_MINB
#define QB_TT_MINB 3
#endif
#ifndef QB_TP
#define QB_TP 6      // epilogue sources prefetched into registers before the sweep
#endif
__device__ __forceinline__ unsigned qb_smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool qb_mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// engine constants of the tile kernel as a kernel parameter (constant bank): warp-uniform
// operands that cost neither a register nor an L1 wavefront
struct QbTileArgs {
    double2* pool;
    const QbPass* pass;
    const qb_c128* coef;
    double* partials;
    int* work;                  // persistent mode: next (slot, tile) work item
    const int* act_list;        // compacted list of the slots with a pass this round (qb_compact_kernel)
    const int* act_count;
    int* coop_rounds;           // cooperative form: rounds executed (added up by block 0)
    int N, V, nslices, red_stride, nelem, maxcoef, mc_trace, tpad_;
    double atol, rtol;
    QbTileElem elem[QB_MAX_ELEMS];
};

// COOP: the cooperative multi-round form for systems in few slots (a single mesolve, a handful
// of trajectories): ONE launch runs up to max_rounds rounds of (pass, partial sums, controller)
// separated by grid barriers instead of three kernel launches per round.  Correct (the engine
// tests pass with it) but not faster -- see qb_drive; opt-in only.  State written in one phase is read in the next by other
// SMs: no non-coherent (ld.global.nc) loads of mutable data in this instantiation.
__device__ void qb_control_slot_coop(QbEngineDev* E, int slot, int lane, double* sred_row);
__device__ void qb_partials_reduce_coop(const QbEngineDev* E, int b);

template <bool CD, bool COOP>
__global__ void __launch_bounds__(256, QB_TT_MINB)
qb_pass_tile_kernel(const QbEngineDev* __restrict__ E, int nslots_used, int trows, int nsb, int persist,
                    const __grid_constant__ QbTileArgs ta, const __grid_constant__ QbConstDesc cd,
                    int max_rounds)
{
    extern __shared__ __align__(128) unsigned char qb_tile_smem[];
    __shared__ __align__(8) unsigned long long mbar;          // x tile
    __shared__ __align__(8) unsigned long long wbar[8];       // per warp: staged epilogue sources
    __shared__ int s_work;
    __shared__ double s_red[COOP ? 8 : 1][QB_MAXRED];
    const int N = ta.N;
    const int ntiles = (N + trows - 1) / trows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // shared-window addresses, computed ONCE (volatile: the compiler would otherwise re-derive
    // them from SR_CgaCtaId in front of every LDS to save a register)
    unsigned sxa;
    asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}"
                 : "=r"(sxa) : "l"(qb_tile_smem));
    const unsigned wba = sxa + (unsigned)trows * 16u + (unsigned)(warp * nsb) * 1024u;   // this warp's source buffer
    const unsigned bar = qb_smem_u32(&mbar);
    const unsigned wb = qb_smem_u32(&wbar[warp]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        for (int w = 0; w < nw; w++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(qb_smem_u32(&wbar[w])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool pow2 = (trows & (trows - 1)) == 0;       // tiles are aligned power-of-two blocks
    unsigned parity = 0, xparity = 0;
    int round = 0;
    for (;; round++) {                            // one iteration unless COOP
    // only the slots that have a pass this round are visited (their list is compacted by
    // qb_compact_kernel after the controller): finished trajectories cost nothing in the tail
    const int total = min(*ta.act_count, nslots_used) * ntiles;
    // persist: the grid is one wave of CTAs that draw (slot, tile) work items from a counter
    // (reset by the control kernel) -- no CTA launch / barrier set-up per tile
    for (bool first = true;; first = false) {
    if (!first) __syncthreads();                  // everybody is done with the previous tile's smem
    if (threadIdx.x == 0) s_work = persist ? atomicAdd(ta.work, 1) : (first ? (int)blockIdx.x : total);
    __syncthreads();              // also: the barrier objects are initialised before anybody uses them
    const int work = s_work;
    if (work >= total) break;
    const int item = work / ntiles;
    const int tile = work - item * ntiles;
    const int slot = ta.act_list[item];
    const QbPass* __restrict__ gp = &ta.pass[slot];
    const int kind = gp->kind;
    if (kind == QB_PASS_NONE || kind == QB_PASS_LINMAP) continue;     // CTA-uniform
    const int lo = tile * trows;
    const int rows = min(trows, N - lo);
    const int sl0 = lo >> 5, sl1 = (lo + rows + 31) >> 5;
    if (kind != QB_PASS_RHS && kind != QB_PASS_COMBINE) {             // EXPECT / APPLY: rare
        for (int sl = sl0 + warp; sl < sl1; sl += nw) qb_pass_slice_generic<COOP>(E, slot, sl, lane);
        continue;
    }
    const int vbase = slot * ta.V;
    const double2* const initp = E->init_states + (long long)E->traj[slot].init_idx * N;
    const double2* gx = nullptr;
    if (kind == QB_PASS_RHS) {
        const int xs = gp->x;
        gx = xs >= 0 ? ta.pool + (long long)(vbase + xs) * N : initp;
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)rows * 16u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier reads of the tile are done
            if (COOP) asm volatile("fence.proxy.async.global;" ::: "memory");   // x was written in this launch
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(sxa), "l"(gx + lo), "r"(bytes), "r"(bar) : "memory");
        }
    }
    const int nsrc = gp->nsrc;
    const int nb = min(nsrc, nsb);                      // sources staged by TMA; the rest is loaded directly
#ifdef QB_SRC_EVICT_FIRST
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
#endif
    const int red = gp->red;
    const bool werr = (red & QB_RED_WRMS) != 0;
    const int nelem = ta.nelem;
    bool staged = false;
    // a warp takes PAIRS of adjacent slices (64 consecutive rows): everything warp-uniform
    // (descriptor lists, pass descriptor, weights) is read once per pair, and the pair's rows
    // of every epilogue source are ONE contiguous kilobyte -- fetched by one TMA bulk copy per
    // source into the warp's own buffer, all in flight during the sweep, no register held
    const int npairs = (sl1 - sl0 + 1) >> 1;
    for (int pr = warp; pr < npairs; pr += nw) {
        const int sla = sl0 + 2 * pr;
        const bool hasb = sla + 1 < sl1;
        const int slb = hasb ? sla + 1 : sla;
        const int r[2] = {sla * 32 + lane, slb * 32 + lane};
        const bool act[2] = {r[0] < N, hasb && r[1] < N};
        if (nb > 0) {
            // lane i issues the copy of source i: the source indices are fetched with one
            // coalesced load and all copies leave in one instruction
            const int r0 = sla * 32;
            const unsigned bytes = (unsigned)min(hasb ? 64 : 32, N - r0) * 16u;
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier reads of the buffer are done
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(wb), "r"(bytes * (unsigned)nb) : "memory");
            }
            __syncwarp();
            if (lane < nb) {
                if (COOP) asm volatile("fence.proxy.async.global;" ::: "memory");   // sources were written in this launch
                const int s = gp->sw[lane].src;
                const double2* p = (s >= 0 ? ta.pool + (long long)(vbase + s) * N : initp) + r0;
#ifdef QB_SRC_EVICT_FIRST
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                             "[%0], [%1], %2, [%3], %4;"
                             :: "r"(wba + (unsigned)lane * 1024u), "l"(p), "r"(bytes), "r"(wb), "l"(pol_stream) : "memory");
#else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(wba + (unsigned)lane * 1024u), "l"(p), "r"(bytes), "r"(wb) : "memory");
#endif
            }
        }
#ifdef QB_TAIL_PF
        for (int i = nb; i < nsrc; i++) {       // sources beyond the staged ones: request them into L2
            const int s = gp->sw[i].src;
            const double2* p = s >= 0 ? ta.pool + (long long)(vbase + max(s, 0)) * N : initp;
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (act[j]) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + r[j]));
        }
#endif
        // ---- operator sweep (x from the staged tile / global memory)
        double2 z[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        if (kind == QB_PASS_RHS) {
            if (!staged) { while (!qb_mbar_try_wait(bar, xparity)) { } staged = true; }
            for (int e = 0; e < nelem; e++) {
                const QbTileElem& A = ta.elem[e];
                const QbSlotDesc* cdp = cd.d + cd.elem_off[e];
                const int4 sa = __ldg(reinterpret_cast<const int4*>(A.sinfo) + sla);
                const int4 sb = __ldg(reinterpret_cast<const int4*>(A.sinfo) + slb);
                double2 q[2];
                if (hasb && sa.x == sb.x && sa.y == sb.y) {
                    const int vb[2] = {sa.z, sb.z}, cb[2] = {sa.w, sb.w};
                    qb_rowdot_rsell_tile<2, CD, COOP>(A, cdp, sa.x, sa.y, vb, cb, r, lane, gx, sxa, lo, rows, trows, pow2, q);
                } else {
                    double2 q1[1];
                    const int va[1] = {sa.z}, ca[1] = {sa.w}, ra[1] = {r[0]};
                    qb_rowdot_rsell_tile<1, CD, COOP>(A, cdp, sa.x, sa.y, va, ca, ra, lane, gx, sxa, lo, rows, trows, pow2, q1);
                    q[0] = q1[0]; q[1] = make_double2(0.0, 0.0);
                    if (hasb) {
                        const int vb1[1] = {sb.z}, cb1[1] = {sb.w}, rb1[1] = {r[1]};
                        qb_rowdot_rsell_tile<1, CD, COOP>(A, cdp, sb.x, sb.y, vb1, cb1, rb1, lane, gx, sxa, lo, rows, trows, pow2, q1);
                        q[1] = q1[0];
                    }
                }
                const qb_c128 c = ta.coef[(size_t)slot * ta.maxcoef + e];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    z[j].x += c.re * q[j].x - c.im * q[j].y;
                    z[j].y += c.re * q[j].y + c.im * q[j].x;
                }
            }
            const double zs = gp->zscale;
#pragma unroll
            for (int j = 0; j < 2; j++) { z[j].x *= zs; z[j].y *= zs; }
        }
        // ---- fused linear combinations (sources in order, z last), stores, reductions
        double2 o1[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        double2 o2[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        if (nb > 0) {
            while (!qb_mbar_try_wait(wb, parity)) { }
            parity ^= 1u;
            const unsigned la = wba + (unsigned)lane * 16u;
            for (int i = 0; i < nb; i++) {
                const double a = gp->sw[i].w1;
                const double b = werr ? gp->w2[i] : 0.0;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double2 v = act[j] ? qb_lds128(la + (unsigned)i * 1024u + (unsigned)j * 512u)
                                             : make_double2(0.0, 0.0);
                    o1[j].x = fma(a, v.x, o1[j].x); o1[j].y = fma(a, v.y, o1[j].y);
                    o2[j].x = fma(b, v.x, o2[j].x); o2[j].y = fma(b, v.y, o2[j].y);
                }
            }
            __syncwarp();         // every lane has read its rows before the buffer is refilled
        }
        for (int i = nb; i < nsrc; i += 2) {
            double2 v[2][2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int s = gp->sw[min(i + u, QB_MAXSRC - 1)].src;
                const double2* p = s >= 0 ? ta.pool + (long long)(vbase + max(s, 0)) * N : initp;
#pragma unroll
                for (int j = 0; j < 2; j++)
                    v[u][j] = (i + u < nsrc && act[j]) ? QB_LDV(p + r[j]) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double a = gp->sw[min(i + u, QB_MAXSRC - 1)].w1;
                const double b = werr ? gp->w2[min(i + u, QB_MAXSRC - 1)] : 0.0;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    o1[j].x = fma(a, v[u][j].x, o1[j].x); o1[j].y = fma(a, v[u][j].y, o1[j].y);
                    o2[j].x = fma(b, v[u][j].x, o2[j].x); o2[j].y = fma(b, v[u][j].y, o2[j].y);
                }
            }
        }
        const double2 hz = *reinterpret_cast<const double2*>(&gp->w1z);  // w1z, w2z
        const int zdst = gp->zdst, dst1 = gp->dst1;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (j == 1 && !hasb) break;                                  // warp-uniform
            o1[j].x = fma(hz.x, z[j].x, o1[j].x); o1[j].y = fma(hz.x, z[j].y, o1[j].y);
            o2[j].x = fma(hz.y, z[j].x, o2[j].x); o2[j].y = fma(hz.y, z[j].y, o2[j].y);
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            if (act[j]) {
                double2* const pool_r = ta.pool + r[j];
                if (zdst >= 0) QB_STV(pool_r + (long long)(vbase + zdst) * N, z[j]);
                if (dst1 >= 0) QB_STV(pool_r + (long long)(vbase + dst1) * N, o1[j]);
                else if (dst1 == QB_SLOT_OUT)
                    E->out_states[((size_t)E->traj[slot].traj_id * E->ctl.nt + gp->out_index) * (size_t)N + r[j]] = o1[j];
                const double n1 = o1[j].x * o1[j].x + o1[j].y * o1[j].y;
                r0 = n1;
                if (werr) {
                    const double q = sqrt(o2[j].x * o2[j].x + o2[j].y * o2[j].y) /
                                     (ta.atol + ta.rtol * sqrt(n1));
                    r1 = q * q;
                }
                r2 = z[j].x * z[j].x + z[j].y * z[j].y;
            }
            if (red) {
                r0 = qb_warp_sum(r0); r1 = qb_warp_sum(r1); r2 = qb_warp_sum(r2);
                if (lane == 0) {
                    double* __restrict__ part = ta.partials +
                        (long long)(slot * ta.nslices + (j ? slb : sla)) * ta.red_stride;
                    part[0] = r0; part[1] = r1; part[2] = r2;
                }
                if (ta.mc_trace)
                    qb_trace_partials(ta.mc_trace, r[j], act[j], o1[j], z[j], lane,
                                      ta.partials + (long long)(slot * ta.nslices + (j ? slb : sla)) * ta.red_stride);
            }
        }
    }
    if (kind == QB_PASS_RHS) xparity ^= 1u;       // the x-tile barrier completed one more phase
    }   // work loop
    if (!COOP) break;
    // ---- the rest of the round: partial sums, controller, each behind a grid barrier
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    grid.sync();
    QbEngineDev* Ew = const_cast<QbEngineDev*>(E);
    if (Ew->red_final) {
        for (int b = blockIdx.x; b < nslots_used * QB_RED_CTAS; b += gridDim.x) qb_partials_reduce_coop(E, b);
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *ta.work = 0;
    for (int slot = blockIdx.x * nw + warp; slot < nslots_used; slot += gridDim.x * nw)
        qb_control_slot_coop(Ew, slot, lane, s_red[COOP ? warp : 0]);
    grid.sync();
    if (*reinterpret_cast<volatile int*>(Ew->n_active) <= 0 || round + 1 >= max_rounds) break;
    }   // rounds
    // (every CTA leaves the loop in the same round)
    if (COOP && blockIdx.x == 0 && threadIdx.x == 0 && ta.coop_rounds) *ta.coop_rounds += round + 1;
}

// LINMAP passes (Adams prediction / update): V[dst[j]] = sum_k w[j][k] V[src[k]] for up to 14
// outputs of up to 15 sources.  Every source row is loaded ONCE into registers (that is the
// point of the pass: O(q) instead of O(q^2) vector reads per step), which needs ~100
// registers -- hence a kernel of its own, launched only by Adams engines; the weights are
// staged in shared memory.
#ifndef QB_LM_MINB
#define QB_LM_MINB 2
#endif
__global__ void __launch_bounds__(QB_TILE_ROWS, QB_LM_MINB)
qb_linmap_kernel(const QbEngineDev* __restrict__ E, int nslots_used)
{
    __shared__ double s_w[QB_LM_MAXOUT][QB_LM_MAXSRC];
    __shared__ int s_dst[QB_LM_MAXOUT];
    const int ntiles = E->ctl.ntiles;
    const int slot = blockIdx.x / ntiles;
    const int tile = blockIdx.x - slot * ntiles;
    if (slot >= nslots_used) return;
    const QbPass* __restrict__ gp = &E->pass[slot];
    if (gp->kind != QB_PASS_LINMAP) return;               // CTA-uniform
    const QbLinMap* __restrict__ lm = &E->linmap[slot];
    const int nout = lm->nout, nsrc = gp->nsrc;
    for (int i = threadIdx.x; i < nout * QB_LM_MAXSRC; i += QB_TILE_ROWS)
        s_w[i / QB_LM_MAXSRC][i % QB_LM_MAXSRC] = lm->w[i / QB_LM_MAXSRC][i % QB_LM_MAXSRC];
    if (threadIdx.x < nout) s_dst[threadIdx.x] = lm->dst[threadIdx.x];
    __syncthreads();
    const int N = E->ctl.N;
    const size_t N_ = (size_t)N;
    const long long r = (long long)tile * QB_TILE_ROWS + threadIdx.x;
    const bool active = r < N;
    double2* slot_base = E->pool + (size_t)slot * E->V * N_;
    const double2* init_ptr = E->init_states + (size_t)E->traj[slot].init_idx * N_;
    double2 v[QB_LM_MAXSRC];
#pragma unroll
    for (int k = 0; k < QB_LM_MAXSRC; k++) {
        const int sidx = gp->sw[k].src;
        const double2* p = sidx >= 0 ? slot_base + (long long)sidx * N : init_ptr;
        v[k] = (k < nsrc && active) ? QB_LDV(p + r) : make_double2(0.0, 0.0);
    }
    double n0 = 0.0, tr0 = 0.0;
    for (int j = 0; j < nout; j++) {
        double2 o = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < QB_LM_MAXSRC; k++) {
            const double w = s_w[j][k];
            o.x = fma(w, v[k].x, o.x); o.y = fma(w, v[k].y, o.y);
        }
        if (active) {
            QB_STV(slot_base + (size_t)s_dst[j] * N_ + r, o);
            if (j == 0) { n0 = o.x * o.x + o.y * o.y; tr0 = o.x; }
        }
    }
    if (gp->red) {
        const int lane = threadIdx.x & 31, sl = (int)(((long long)tile * QB_TILE_ROWS + threadIdx.x) >> 5);
        if ((long long)sl * 32 < N) {                      // warp-uniform
            n0 = qb_warp_sum(n0);
            if (lane == 0) {
                double* __restrict__ part = E->partials + ((size_t)slot * E->nslices + sl) * E->red_stride;
                part[0] = n0; part[1] = 0.0; part[2] = 0.0;
            }
            if (E->ctl.mc_trace) {
                const bool diag = active && (r % (E->ctl.mc_trace + 1) == 0);
                const double t1 = qb_warp_sum(diag ? tr0 : 0.0);
                if (lane == 0) {
                    double* __restrict__ part = E->partials + ((size_t)slot * E->nslices + sl) * E->red_stride;
                    part[3] = t1; part[4] = 0.0;
                }
            }
        }
    }
}

// Shared-operator variant for systems whose RHS is ONE SELL operator (mcsolve H_eff): the 8
// warps of a CTA are 8 DIFFERENT trajectory slots working on the SAME slice, so the slice's
// values and column indices are staged in shared memory once and re-used 8 times -- the
// operator's L2->SM traffic drops 8x.  A CTA walks QB_SH_T consecutive slices so that each
// warp's state gathers keep their L1 locality.
__global__ void __launch_bounds__(QB_TILE_ROWS, QB_MINB)
qb_pass_kernel_shared(const QbEngineDev* __restrict__ E)
{
    __shared__ double2 s_val[QB_SH_MAXW * 32];
    __shared__ int s_col[QB_SH_MAXW * 32];
    const int nslices = E->nslices;
    const int nchunks = (nslices + QB_SH_T - 1) / QB_SH_T;
    const int group = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - group * nchunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = group * (QB_TILE_ROWS / 32) + warp;
    const bool slot_ok = slot < E->nslots && E->pass[slot].kind != QB_PASS_NONE &&
                         E->pass[slot].kind != QB_PASS_LINMAP;
    if (!__syncthreads_or(slot_ok)) return;
    QbWarpHdr h;
    if (slot_ok) qb_load_hdr(E, slot, lane, h);
    const QbOpDev& A = E->elem[0];
    const double2* __restrict__ gval = reinterpret_cast<const double2*>(A.val);
    for (int it = 0; it < QB_SH_T; it++) {
        const int sl = chunk * QB_SH_T + it;
        if (sl >= nslices) break;                        // CTA-uniform
        const int s0 = A.slice_ptr[sl], w = A.slice_ptr[sl + 1] - s0;
        const bool staged = w <= QB_SH_MAXW;
        if (staged) {
            for (int i = threadIdx.x; i < w * 32; i += QB_TILE_ROWS) {
                s_val[i] = __ldg(gval + (size_t)s0 * 32 + i);
                s_col[i] = __ldg(A.col + (size_t)s0 * 32 + i);
            }
        }
        __syncthreads();
        if (slot_ok) qb_pass_one_slice(E, slot, sl, lane, h, staged ? s_val : nullptr, s_col, w);
        __syncthreads();
    }
}

// Large systems (N/32 > 2048 slices): QB_RED_CTAS CTAs per slot sum the per-warp partials in
// a fixed order (the control kernel's warp then adds the QB_RED_CTAS sub-sums), so that the
// control kernel's single warp does not walk tens of thousands of partials serially.

__device__ __forceinline__ void qb_partials_reduce_body(const QbEngineDev* __restrict__ E, int b)
{
    const int slot = b / QB_RED_CTAS, c = b - slot * QB_RED_CTAS;
    const QbPass* gp = &E->pass[slot];
    const int kind = gp->kind;
    int nred = 0;
    if (kind == QB_PASS_EXPECT) nred = 2 * (gp->op_hi - gp->op_lo);
    else if (kind != QB_PASS_NONE && gp->red) nred = E->ctl.mc_trace ? 5 : 3;
    if (nred == 0) return;                           // CTA-uniform
    __shared__ double sh[32];
    const int nslices = E->nslices, stride = E->red_stride;
    const int chunk = (nslices + QB_RED_CTAS - 1) / QB_RED_CTAS;
    const int i0 = c * chunk, i1 = min(nslices, i0 + chunk);
    const double* part = E->partials + (size_t)slot * nslices * stride;
    const int nthr = (int)blockDim.x, nwarps = nthr >> 5;
    for (int k = 0; k < nred; k++) {
        double s = 0.0;
        for (int i = i0 + threadIdx.x; i < i1; i += nthr) s += part[(size_t)i * stride + k];
        s = qb_warp_sum(s);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < nwarps; w++) t += sh[w];
            E->red_final[((size_t)slot * QB_RED_CTAS + c) * QB_MAXRED + k] = t;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
qb_partials_reduce_kernel(const QbEngineDev* __restrict__ E)
{
    qb_partials_reduce_body(E, (int)blockIdx.x);
}

// ------------------------------------------------------------------ control kernel
__device__ void qb_start_traj(QbEngineDev* E, QbTraj& c, int traj_id) {
    const int S = E->ctl.tab.S;
    c.traj_id = traj_id;
    c.init_idx = E->init_map ? E->init_map[traj_id] : 0;
    c.mode = E->mode;
    c.tl_idx = 0; c.tl_end = E->ctl.nt;
    c.sP = S; c.sF = S + 1; c.sI = S + 2; c.sTA = S + 3; c.sTB = S + 4; c.sY = S + 1;
    c.status = QB_ST_NORMAL; c.done = 0;
    c.pc = E->mode ? QB_PC_MC_BEGIN : QB_PC_ME_BEGIN;
}

// Slots with a pass for the tile kernel, compacted in slot order (one block; runs after the
// controller of every round).
__global__ void __launch_bounds__(1024)
qb_compact_kernel(QbEngineDev* __restrict__ E, int nslots_used)
{
    __shared__ int s_cnt[1024];
    const int per = (nslots_used + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(nslots_used, lo + per);
    int c = 0;
    for (int s = lo; s < hi; s++) {
        const int k = E->pass[s].kind;
        c += (k != QB_PASS_NONE && k != QB_PASS_LINMAP);
    }
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                  // inclusive scan
        const int v = threadIdx.x >= d ? s_cnt[threadIdx.x - d] : 0;
        __syncthreads();
        s_cnt[threadIdx.x] += v;
        __syncthreads();
    }
    int pos = s_cnt[threadIdx.x] - c;
    for (int s = lo; s < hi; s++) {
        const int k = E->pass[s].kind;
        if (k != QB_PASS_NONE && k != QB_PASS_LINMAP) E->act_list[pos++] = s;
    }
    if (threadIdx.x == 1023) *E->act_count = s_cnt[1023];
}

// the scalar part of a round for one slot: consume the reduced partials, advance the state
// machine until it issues the next pass (or retires / refills the slot)
__device__ __forceinline__ void qb_control_advance(QbEngineDev* __restrict__ E, int slot, double* sred_row)
{
    QbTraj* gc = &E->traj[slot];
    QbPass* gp = &E->pass[slot];
    QbTraj c = *gc;
    QbPass p;
    qb_c128* coef = E->coef + (size_t)slot * E->ctl.maxcoef;
    double* probs = E->probs + (size_t)slot * (E->ctl.ncops > 0 ? E->ctl.ncops : 1);
    for (;;) {
        const int issued = qb_advance(E->ctl, c_tabs[E->tableau_id], c, p, sred_row, coef, probs,
                                      &E->linmap[slot]);
        if (issued) {
            c.n_pass++;
            if (E->vec_count) {
                // algorithmic state traffic of the pass: x once, every source once, stores
                unsigned long long nv = (unsigned long long)p.nsrc + (p.zdst >= 0) + (p.dst1 != -1)
                                        + (p.kind != QB_PASS_COMBINE ? 1 : 0);
                if (p.kind == QB_PASS_LINMAP) nv = (unsigned long long)p.nsrc + E->linmap[slot].nout;
                atomicAdd(E->vec_count, nv);
            }
            break;
        }
        // finished, paused or failed
        if (c.done == 2) { atomicSub(E->n_active, 1); break; }   // waits for host coefficients
        if (E->out_status && c.traj_id >= 0) E->out_status[c.traj_id] = c.done;
        if (E->out_stats && c.traj_id >= 0) {
            int* st = E->out_stats + (size_t)c.traj_id * 4;
            st[0] = c.n_rhs; st[1] = c.n_accept; st[2] = c.n_reject; st[3] = c.n_pass;
        }
        int next = E->ntraj_total;
        if (E->queue_head) next = atomicAdd(E->queue_head, 1);
        if (next >= E->ntraj_total) { atomicSub(E->n_active, 1); break; }
        qb_start_traj(E, c, next);
    }
    *gc = c;
    *gp = p;
    if (E->zbuf) {
        double2* zrow = E->zbuf + (size_t)slot * (size_t)E->ctl.N;
        E->zcols[slot] = zrow;
        E->xcols[slot] = (p.kind == QB_PASS_RHS) ? qb_vsrc(E, slot, p.x, c.init_idx) : zrow;
    }
}

// BIG (systems with more than 2048 slices): one CTA of 1024 threads per slot sums the slot's
// partials cooperatively in a fixed order before thread 0 runs the controller -- one launch
// instead of qb_partials_reduce_kernel + a single-warp controller.
// one warp = one slot: sum the slot's partials (or the pre-reduced sub-sums) in a fixed order,
// then lane 0 runs the controller
__device__ __forceinline__ void qb_control_slot_warp(QbEngineDev* __restrict__ E, int slot, int lane,
                                                     double* sred_row)
{
    if (E->traj[slot].pc == QB_PC_IDLE) return;
    const QbPass* gp = &E->pass[slot];
    const int kind = gp->kind;
    int nred = 0;
    if (kind == QB_PASS_EXPECT) nred = 2 * (gp->op_hi - gp->op_lo);
    else if (kind != QB_PASS_NONE && gp->red) nred = E->ctl.mc_trace ? 5 : 3;
    const int nslices = E->nslices, stride = E->red_stride;
    const double* part = E->partials + (size_t)slot * nslices * stride;
    if (E->red_final) {
        for (int k = 0; k < nred; k++) {      // lane c holds CTA c's sub-sum
            double sv = E->red_final[((size_t)slot * QB_RED_CTAS + lane) * QB_MAXRED + k];
            sv = qb_warp_sum(sv);
            if (lane == 0) sred_row[k] = sv;
        }
    } else
    for (int k = 0; k < nred; k++) {
        double s = 0.0;
        for (int i = lane; i < nslices; i += 32) s += part[(size_t)i * stride + k];
        s = qb_warp_sum(s);
        if (lane == 0) sred_row[k] = s;
    }
    __syncwarp();
    if (lane != 0) return;
    qb_control_advance(E, slot, sred_row);
}

__device__ __noinline__ void qb_control_slot_coop(QbEngineDev* E, int slot, int lane, double* sred_row)
{
    qb_control_slot_warp(E, slot, lane, sred_row);
}
__device__ __noinline__ void qb_partials_reduce_coop(const QbEngineDev* E, int b)
{
    qb_partials_reduce_body(E, b);
}

template <bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 128)
qb_control_kernel(QbEngineDev* __restrict__ E)
{
    const int slot = BIG ? (int)blockIdx.x : (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31, w = BIG ? 0 : (int)(threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0 && E->work) *E->work = 0;
    if (slot >= E->nslots) return;
    __shared__ double sred[4][QB_MAXRED];
    __shared__ double swarp[BIG ? 32 : 1];
    if (!BIG) { qb_control_slot_warp(E, slot, lane, sred[w]); return; }
    QbTraj* gc = &E->traj[slot];
    if (gc->pc == QB_PC_IDLE) return;
    QbPass* gp = &E->pass[slot];
    const int kind = gp->kind;
    int nred = 0;
    if (kind == QB_PASS_EXPECT) nred = 2 * (gp->op_hi - gp->op_lo);
    else if (kind != QB_PASS_NONE && gp->red) nred = E->ctl.mc_trace ? 5 : 3;
    const int nslices = E->nslices, stride = E->red_stride;
    const double* __restrict__ part = E->partials + (size_t)slot * nslices * stride;
    for (int k = 0; k < nred; k++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nslices; i += 1024) s += part[(size_t)i * stride + k];
        s = qb_warp_sum(s);
        if (lane == 0) swarp[BIG ? (threadIdx.x >> 5) : 0] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < 32; q++) t += swarp[BIG ? q : 0];
            sred[0][k] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    qb_control_advance(E, slot, sred[0]);
}

// one RHS evaluation on plain device vectors (micro-benchmark / data-layer matmul of a
// whole QobjEvo): out = sum_k coef_k A_k x
__global__ void __launch_bounds__(QB_TILE_ROWS, 4)
qb_rhs_kernel(const QbEngineDev* __restrict__ E, const double2* __restrict__ x,
              double2* __restrict__ out, const qb_c128* __restrict__ coef)
{
    const int N = E->ctl.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x * (QB_TILE_ROWS / 32) + warp;
    const long long r = (long long)sl * 32 + lane;
    if ((long long)sl * 32 >= N) return;
    const bool active = r < N;
    double2 z = make_double2(0.0, 0.0);
    for (int e = 0; e < E->ctl.nelem; e++) {
        const double2 q = qb_rowdot(E->elem[e], sl, lane, r, active, x);
        const qb_c128 c = coef[e];
        z.x += c.re * q.x - c.im * q.y;
        z.y += c.re * q.y + c.im * q.x;
    }
    if (active) out[r] = z;
}

// sum over trajectories: sums[0][e][t] = sum_j v, sums[1][e][t] = sum_j (re^2, im^2)
__global__ void qb_reduce_expect_kernel(const double2* __restrict__ ex, long long ntraj,
                                        int n, double2* __restrict__ sums)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 s = make_double2(0.0, 0.0), s2 = make_double2(0.0, 0.0);
    for (long long j = 0; j < ntraj; j++) {
        const double2 v = ex[(size_t)j * n + i];
        s.x += v.x; s.y += v.y; s2.x += v.x * v.x; s2.y += v.y * v.y;
    }
    sums[i] = s; sums[n + i] = s2;
}

// ================================================================== host side
struct QbEngH : QbObj {
    QbSysH* sys = nullptr;
    int tableau = 0, nslots = 0, V = 0;
    QbOptions opt;
    QbEngineDev h;                 // host mirror of the device descriptor
    QbEngineDev* d = nullptr;
    void* io_p[12] = {nullptr}; size_t io_cap[12] = {0};   // staging buffers of qb_engine_run
    // Integrator protocol (slot 0): host mirrors that save a blocking copy per call
    QbTraj traj0; bool traj0_valid = false;          // slot 0's controller state after the last call
    QbEngineDev d_copy; bool d_copy_valid = false;   // what *d currently holds
    std::vector<void*> owned;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int* h_active = nullptr;       // pinned
    // buffers re-allocated per run when sizes change
    void* d_tlist = nullptr; int cap_nt = 0;
    void* d_init = nullptr; size_t cap_init = 0;
    void* d_args = nullptr; size_t cap_args = 0;
    long long last_rounds = 0;
    double last_ms = 0.0;
    int profiling = 0;
    cudaGraphExec_t graph = nullptr;
    int graph_slots = 0;
    cudaEvent_t ev_chunk[2] = {nullptr, nullptr};
    int no_shared = 1;          // qb_pass_kernel_shared only when QB_SHARED is set
    int tile_g = 0, tile_rows = 0, tile_xw = 0, tile_ns = 0, tile_threads = 256;   // TMA-staged kernel (tile_g == 0: off)
    size_t tile_smem = 0;
    int tile_persist = 0;        // CTAs per SM of the persistent grid (0: one CTA per tile)
    int coop_occ = 0;            // CTAs per SM of the cooperative multi-round form (0: unavailable)
    int act_identity = 0;        // act_list currently holds the identity list of this many slots
    QbConstDesc cdesc;          // descriptor lists in the constant bank (n == 0: read from global memory)
    QbTileArgs targs;
    double prof_pass_ms = 0.0;
    long long prof_pass_launches = 0;
    unsigned long long prof_vec_count = 0;
    std::vector<cudaEvent_t> prof_events;
    std::vector<double> prof_round_ms;                  // per round of the last profiled run
    std::vector<unsigned long long> prof_round_vec;     // cumulative vector accesses after the round
    unsigned long long* prof_vec_host = nullptr;        // pinned
    size_t prof_vec_cap = 0;
    int maxcoef = 1;
    int64_t last_ntraj = 0; int last_nt = 0;     // shape of the expectation values of the last qb_engine_run
    QbEngH() : QbObj(QB_TAG_ENG) { memset(&h, 0, sizeof h); }
    ~QbEngH() override {
        for (void* p : owned) cudaFree(p);
        for (void* p : io_p) if (p) cudaFree(p);
        if (d_tlist) cudaFree(d_tlist);
        if (d_init) cudaFree(d_init);
        if (d_args) cudaFree(d_args);
        if (h_active) cudaFreeHost(h_active);
        if (prof_vec_host) cudaFreeHost(prof_vec_host);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (graph) cudaGraphExecDestroy(graph);
        for (auto ev : ev_chunk) if (ev) cudaEventDestroy(ev);
        for (auto ev : prof_events) cudaEventDestroy(ev);
        if (stream) cudaStreamDestroy(stream);
    }
};

template <class T> static int qb_dev_array(QbEngH* e, const std::vector<T>& v, const T** out) {
    *out = nullptr;
    if (v.empty()) return QB_OK;
    void* p = nullptr;
    QB_CUDA(cudaMalloc(&p, v.size() * sizeof(T)));
    e->owned.push_back(p);
    QB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return QB_OK;
}
template <class T> static int qb_dev_alloc(QbEngH* e, size_t n, T** out, bool zero = true) {
    void* p = nullptr;
    if (n == 0) n = 1;
    QB_CUDA(cudaMalloc(&p, n * sizeof(T)));
    e->owned.push_back(p);
    if (zero) QB_CUDA(cudaMemset(p, 0, n * sizeof(T)));
    *out = static_cast<T*>(p);
    return QB_OK;
}

extern "C" int qb_options_default(qb_options* o) {
    if (!o) QB_FAIL(QB_E_ARG, "null options");
    o->atol = 1e-8; o->rtol = 1e-6; o->nsteps = 1000;
    o->first_step = 0; o->min_step = 0; o->max_step = 0; o->interpolate = 1;
    o->norm_steps = 25; o->norm_t_tol = 1e-6; o->norm_tol = 1e-4; o->norm_min_step = 0.1;
    o->mc_corr_eps = 1e-10; o->store_states = 0; o->max_collapses = 64; o->no_jump = 0;
    o->jump_prob_floor = 0.0;
    o->max_order = 0; o->pad_ = 0;
    return QB_OK;
}

// ---- system ----
extern "C" int qb_system_create(int64_t N, int nargs, qb_handle* out) {
    if (N <= 0 || N > 0x7fffffff || nargs < 0 || !out) QB_FAIL(QB_E_ARG, "bad system size");
    QbSysH* s = new QbSysH();
    s->N = N; s->nargs = nargs;
    *out = s;
    return QB_OK;
}
static int qb_get_opdev(qb_handle op, int64_t N, QbOpDev* out) {
    if (QbOpH* o = qb_cast<QbOpH>(op, QB_TAG_OP)) {
        if (o->dev.nrows != N || o->dev.ncols != N)
            QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes (%d, %d) and (%lld, 1)",
                    o->dev.nrows, o->dev.ncols, (long long)N);
        *out = o->dev; return QB_OK;
    }
    if (QbDenseH* dn = qb_cast<QbDenseH>(op, QB_TAG_DENSE)) {
        if (dn->rows != N || dn->cols != N)
            QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes (%lld, %lld) and (%lld, 1)",
                    (long long)dn->rows, (long long)dn->cols, (long long)N);
        if (!dn->fortran) QB_FAIL(QB_E_TYPE, "dense operators must be column-major (fortran)");
        memset(out, 0, sizeof *out);
        out->fmt = QB_FMT_DENSE; out->nrows = (int)N; out->ncols = (int)N;
        out->nnz = N * N; out->dense = reinterpret_cast<const qb_c128*>(dn->d);
        return QB_OK;
    }
    QB_FAIL(QB_E_TYPE, "handle is not an operator");
}
static std::vector<QbInstr> qb_prog(const qb_instr* p, int n) {
    std::vector<QbInstr> v;
    for (int i = 0; i < n; i++) { QbInstr q; q.op = p[i].op; q.iarg = p[i].iarg; q.re = p[i].re; q.im = p[i].im; v.push_back(q); }
    return v;
}
extern "C" int qb_system_add_element(qb_handle sys, qb_handle op, const qb_instr* prog, int nprog) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    if ((int)s->elems.size() >= QB_MAX_ELEMS) QB_FAIL(QB_E_ARG, "too many elements (max %d)", QB_MAX_ELEMS);
    QbOpDev d; int rc = qb_get_opdev(op, s->N, &d); if (rc) return rc;
    s->elems.push_back(d); s->elem_prog.push_back(qb_prog(prog, nprog));
    return QB_OK;
}
extern "C" int qb_system_add_collapse(qb_handle sys, qb_handle c_op, const qb_instr* cprog, int ncprog,
                                      qb_handle n_op, const qb_instr* nprog, int nnprog) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    QbOpDev c, n; int rc = qb_get_opdev(c_op, s->N, &c); if (rc) return rc;
    rc = qb_get_opdev(n_op, s->N, &n); if (rc) return rc;
    s->cops.push_back(c); s->nops.push_back(n);
    s->cop_prog.push_back(qb_prog(cprog, ncprog)); s->nop_prog.push_back(qb_prog(nprog, nnprog));
    return QB_OK;
}
extern "C" int qb_system_add_eop(qb_handle sys, qb_handle op, const qb_instr* prog, int nprog) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    QbOpDev d; int rc = qb_get_opdev(op, s->N, &d); if (rc) return rc;
    s->eops.push_back(d); s->eop_prog.push_back(qb_prog(prog, nprog));
    return QB_OK;
}
extern "C" int qb_system_set_eop_functional(qb_handle sys, int functional) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    s->eop_functional = functional ? 1 : 0;
    return QB_OK;
}
extern "C" int qb_system_set_mc_trace(qb_handle sys, int n) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    if (n < 0 || (n > 0 && (int64_t)n * n != s->N)) QB_FAIL(QB_E_SHAPE, "mc_trace needs a system of size n*n");
    s->mc_trace = n;
    return QB_OK;
}
extern "C" int qb_system_add_spline(qb_handle sys, const double* tlist, const void* poly,
                                    int n, int order, double dt, int* id) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s || !tlist || !poly || n < 1 || order < 0) QB_FAIL(QB_E_ARG, "bad spline");
    QbSpline sp; sp.n = n; sp.order = order; sp.uniform = dt != 0.0; sp.pad_ = 0; sp.dt = dt;
    sp.t_off = (long long)s->spool.size();
    s->spool.insert(s->spool.end(), tlist, tlist + n);
    sp.p_off = (long long)s->spool.size();
    const double* pp = static_cast<const double*>(poly);
    s->spool.insert(s->spool.end(), pp, pp + 2 * (size_t)(order + 1) * n);
    if (id) *id = (int)s->splines.size();
    s->splines.push_back(sp);
    return QB_OK;
}

// ---- engine ----
extern "C" int qb_engine_create(qb_handle sys, int tableau, int nslots, const qb_options* opt,
                                qb_handle* out) {
    QbSysH* s = qb_cast<QbSysH>(sys, QB_TAG_SYS);
    if (!s) QB_FAIL(QB_E_TYPE, "not a system handle");
    if (tableau < 0 || tableau > 3) QB_FAIL(QB_E_ARG, "unknown tableau id %d (0 vern7, 1 vern9, 2 tsit5, 3 adams)", tableau);
    if (nslots < 1 || !out || !opt) QB_FAIL(QB_E_ARG, "bad engine arguments");
    if (s->elems.empty()) QB_FAIL(QB_E_STATE, "system has no elements");
    QbEngH* e = new QbEngH();
    e->sys = s; e->tableau = tableau; e->nslots = nslots;
    // the shared-operator variant measured 7 % slower than the warp-autonomous kernel on C3
    // (its two barriers per slice cost more than the 8x smaller operator traffic saves): opt-in
    e->no_shared = getenv("QB_SHARED") == nullptr;
    static_assert(sizeof(qb_options) == sizeof(QbOptions), "options layout");
    memcpy(&e->opt, opt, sizeof(QbOptions));
    if (e->opt.max_collapses < 1) e->opt.max_collapses = 1;
    if (tableau == 3) { e->opt.atol *= QB_AD_TOL_SCALE; e->opt.rtol *= QB_AD_TOL_SCALE; }   // qb_adams.h
    QbEngineDev& h = e->h;
    if (tableau == 3) qb_adams_table(&h.ctl.tab);
    else h.ctl.tab = *QB_TABLEAUX[tableau];
    h.tableau_id = tableau;
    {
        static bool tabs_uploaded_dev[64] = {false};     // constant memory is per device
        bool& tabs_uploaded = tabs_uploaded_dev[(e->device >= 0 && e->device < 64) ? e->device : 0];
        if (!tabs_uploaded) {
            static QbTableau both[4];
            both[0] = *QB_TABLEAUX[0]; both[1] = *QB_TABLEAUX[1]; both[2] = *QB_TABLEAUX[2];
            qb_adams_table(&both[3]);
            cudaError_t ce = cudaMemcpyToSymbol(c_tabs, both, sizeof(both));
            if (ce != cudaSuccess) { delete e; QB_FAIL(QB_E_CUDA, "tableau upload: %s", cudaGetErrorString(ce)); }
            tabs_uploaded = true;
        }
    }
    h.ctl.opt = e->opt;
    h.ctl.N = (int)s->N;
    h.ctl.ntiles = (int)((s->N + QB_TILE_ROWS - 1) / QB_TILE_ROWS);
    h.ctl.nelem = (int)s->elems.size();
    h.ctl.ncops = (int)s->cops.size();
    h.ctl.neops = (int)s->eops.size();
    h.ctl.nargs = s->nargs;
    h.ctl.eop_functional = s->eop_functional;
    h.ctl.mc_trace = s->mc_trace;
    h.ctl.has_host_coef = 0;
    for (auto& pr : s->elem_prog)
        if (pr.size() == 1 && pr[0].op == QB_I_HOST) h.ctl.has_host_coef = 1;
    e->maxcoef = std::max(1, h.ctl.nelem);
    h.ctl.maxcoef = e->maxcoef;
    e->V = h.ctl.tab.S + 5;
    h.V = e->V; h.nslots = nslots;
    int rc;
#define QB_TRY(x) do { rc = (x); if (rc) { delete e; return rc; } } while (0)
    // programs
    std::vector<QbInstr> instr;
    auto pack = [&](const std::vector<std::vector<QbInstr>>& progs) {
        std::vector<QbProgRef> refs;
        for (auto& p : progs) { QbProgRef r; r.off = (int)instr.size(); r.len = (int)p.size(); refs.push_back(r); instr.insert(instr.end(), p.begin(), p.end()); }
        return refs;
    };
    std::vector<QbProgRef> r_el = pack(s->elem_prog), r_c = pack(s->cop_prog),
                           r_n = pack(s->nop_prog), r_e = pack(s->eop_prog);
    QB_TRY(qb_dev_array(e, r_el, &h.ctl.elem_prog));
    QB_TRY(qb_dev_array(e, r_c, &h.ctl.cop_prog));
    QB_TRY(qb_dev_array(e, r_n, &h.ctl.nop_prog));
    QB_TRY(qb_dev_array(e, r_e, &h.ctl.eop_prog));
    QB_TRY(qb_dev_array(e, instr, &h.ctl.instr));
    QB_TRY(qb_dev_array(e, s->splines, &h.ctl.splines));
    QB_TRY(qb_dev_array(e, s->spool, &h.ctl.spool));
    for (size_t i = 0; i < s->elems.size(); i++) h.elem[i] = s->elems[i];
    QB_TRY(qb_dev_array(e, s->cops, &h.cops));
    QB_TRY(qb_dev_array(e, s->nops, &h.nops));
    QB_TRY(qb_dev_array(e, s->eops, &h.eops));
    // state
    const size_t N = (size_t)s->N;
    {
        cudaError_t ce = cudaMalloc((void**)&h.pool, (size_t)nslots * e->V * N * sizeof(double2));
        if (ce != cudaSuccess) { delete e; QB_FAIL(QB_E_ALLOC, "cannot allocate %.1f MB of state buffers: %s",
            (double)nslots * e->V * N * 16 / 1e6, cudaGetErrorString(ce)); }
        e->owned.push_back(h.pool);
    }
    QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.traj));
    QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.pass));
    QB_TRY(qb_dev_alloc(e, (size_t)nslots * e->maxcoef, &h.coef));
    QB_TRY(qb_dev_alloc(e, (size_t)nslots * std::max(1, h.ctl.ncops), &h.probs));
    h.nslices = (int)((s->N + 31) / 32);
    {
        const int nops = std::max(h.ctl.ncops, h.ctl.neops);
        h.red_stride = std::max(6, 2 * std::min(QB_MAXRED / 2, nops));
    }
    QB_TRY(qb_dev_alloc(e, (size_t)nslots * h.nslices * h.red_stride, &h.partials));
    h.linmap = nullptr;
    if (h.ctl.tab.method == 1) QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.linmap));
    QB_TRY(qb_dev_alloc(e, 1, &h.queue_head));
    QB_TRY(qb_dev_alloc(e, 1, &h.n_active));
    QB_TRY(qb_dev_alloc(e, 1, &h.work));
    QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.act_list));
    QB_TRY(qb_dev_alloc(e, 1, &h.act_count));
    QB_TRY(qb_dev_alloc(e, 1, &h.vec_count));
    h.zbuf = nullptr; h.xcols = nullptr; h.zcols = nullptr;
    h.red_final = nullptr;
    h.all_sell = getenv("QB_NO_HOT") ? 0 : 1;
    h.all_lean = h.all_sell;
    for (auto& el : s->elems) {
        if (el.fmt != QB_FMT_SELL) h.all_sell = 0;
        if (el.fmt != QB_FMT_SELL && el.fmt != QB_FMT_DIAM && el.fmt != QB_FMT_KRON) h.all_lean = 0;
        if (el.fmt == QB_FMT_KRON && getenv("QB_KRON_GENERIC")) h.all_lean = 0;
    }
    h.all_rsell = 1;
    for (auto& el : s->elems) if (el.fmt != QB_FMT_RSELL) h.all_rsell = 0;
    if (h.all_rsell && !getenv("QB_NO_TILE")) {
        // tile geometry: QB_TILE_ROWS rows of one slot per CTA (x tile = rows * 16 B of smem)
        const char* er = getenv("QB_TILE_ROWS"); const char* et = getenv("QB_TILE_THREADS");
        int rows = er ? atoi(er) : 1024, thr = et ? atoi(et) : 256;
        const int nround = (int)((s->N + 31) / 32 * 32);
        rows = std::max(32, std::min(rows, 8192)) & ~31;
        if (rows > nround) rows = nround;
        // few slots of a large system: smaller tiles, so that the (persistent) grid draws enough
        // work items to balance its last wave (C2: 1024 tiles of 1024 rows for 444 CTAs -> 4096 of 256)
        if (!er)
            while (rows > 256 && (rows & (rows - 1)) == 0 &&
                   (long long)nslots * ((s->N + rows - 1) / rows) < 3552) rows >>= 1;
        thr = std::max(32, std::min(256, thr)) & ~31;
        if (thr > rows) thr = rows;
        if (!et) thr = std::min(thr, std::max(32, (rows / 2) & ~31));     // a warp takes pairs of slices
        const char* eb = getenv("QB_TILE_NSB");
        int nsb = eb ? atoi(eb) : 6;                   // epilogue sources staged per warp (1 KB each)
        nsb = std::max(0, std::min(nsb, QB_MAXSRC));
        e->tile_g = 1; e->tile_rows = rows; e->tile_threads = thr; e->tile_ns = nsb;
        e->tile_smem = (size_t)rows * 16 + (size_t)(thr / 32) * nsb * 1024;
        cudaError_t ce = cudaFuncSetAttribute(qb_pass_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tile_smem);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qb_pass_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tile_smem);
#ifdef QB_ENABLE_COOP
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qb_pass_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tile_smem);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qb_pass_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tile_smem);
#endif
        if (ce != cudaSuccess) { delete e; QB_FAIL(QB_E_CUDA, "tile kernel shared memory: %s", cudaGetErrorString(ce)); }
        memset(&e->targs, 0, sizeof e->targs);
        e->targs.pool = h.pool; e->targs.pass = h.pass; e->targs.coef = h.coef; e->targs.partials = h.partials;
        e->targs.work = h.work; e->targs.act_list = h.act_list; e->targs.act_count = h.act_count;
        // persistent work-queue grid with as many CTAs per SM as fit (measured -3 % on C3 against one
        // CTA per tile); QB_TILE_PERSIST=0 restores the latter, =n forces n CTAs per SM
        e->tile_persist = getenv("QB_TILE_PERSIST") ? atoi(getenv("QB_TILE_PERSIST")) : -1;
        e->targs.N = h.ctl.N; e->targs.V = h.V; e->targs.nslices = h.nslices; e->targs.red_stride = h.red_stride;
        e->targs.nelem = h.ctl.nelem; e->targs.maxcoef = h.ctl.maxcoef; e->targs.mc_trace = h.ctl.mc_trace;
        e->targs.atol = e->opt.atol; e->targs.rtol = e->opt.rtol;
        for (size_t i = 0; i < s->elems.size(); i++) {
            e->targs.elem[i].sinfo = s->elems[i].sinfo; e->targs.elem[i].val = s->elems[i].val;
            e->targs.elem[i].col = s->elems[i].col; e->targs.elem[i].sdesc = s->elems[i].sdesc;
        }
        // descriptor lists of all elements into the constant-bank table when they fit
        memset(&e->cdesc, 0, sizeof e->cdesc);
        int total = 0;
        for (auto& el : s->elems) total += el.ndesc;
        if (e->tile_persist < 0) {
            int occ = 0;
            cudaError_t co = (total <= QB_CD_MAX && !getenv("QB_NO_CDESC"))
                ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qb_pass_tile_kernel<true, false>, e->tile_threads, e->tile_smem)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qb_pass_tile_kernel<false, false>, e->tile_threads, e->tile_smem);
            e->tile_persist = (co == cudaSuccess && occ > 0) ? occ : 0;
        }
#ifdef QB_ENABLE_COOP
        {   // cooperative multi-round form: its grid must be co-resident
            int occ = 0, dev = 0, coop_ok = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev);
            cudaError_t co = (total <= QB_CD_MAX && !getenv("QB_NO_CDESC"))
                ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qb_pass_tile_kernel<true, true>, e->tile_threads, e->tile_smem)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qb_pass_tile_kernel<false, true>, e->tile_threads, e->tile_smem);
            e->coop_occ = (co == cudaSuccess && coop_ok && !getenv("QB_NO_COOP")) ? occ : 0;
            if (e->coop_occ > 0) {
                if (qb_dev_alloc(e, 1, &e->targs.coop_rounds)) e->coop_occ = 0;
            }
        }
#endif
        if (total <= QB_CD_MAX && !getenv("QB_NO_CDESC")) {
            int off = 0;
            for (size_t i = 0; i < s->elems.size(); i++) {
                e->cdesc.elem_off[i] = off;
                if (s->elems[i].ndesc > 0 &&
                    cudaMemcpy(&e->cdesc.d[off], s->elems[i].sdesc, (size_t)s->elems[i].ndesc * sizeof(QbSlotDesc),
                               cudaMemcpyDeviceToHost) != cudaSuccess) { delete e; QB_FAIL(QB_E_CUDA, "descriptor read-back failed"); }
                off += s->elems[i].ndesc;
            }
            e->cdesc.n = total > 0 ? total : 0;
        }
    }
    if (h.nslices > 2048) QB_TRY(qb_dev_alloc(e, (size_t)nslots * QB_RED_CTAS * QB_MAXRED, &h.red_final));
    if (s->elems.size() == 1 && s->elems[0].fmt == QB_FMT_DENSE && nslots >= 8) {
        QB_TRY(qb_dev_alloc(e, (size_t)nslots * N, &h.zbuf));
        QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.xcols));
        QB_TRY(qb_dev_alloc(e, (size_t)nslots, &h.zcols));
        std::vector<double2*> rows(nslots);
        for (int i = 0; i < nslots; i++) rows[i] = h.zbuf + (size_t)i * N;
        if (cudaMemcpy(h.xcols, rows.data(), nslots * sizeof(void*), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h.zcols, rows.data(), nslots * sizeof(void*), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete e; QB_FAIL(QB_E_CUDA, "column pointer upload failed");
        }
    }
    {
        void* p = nullptr;
        cudaError_t ce = cudaMalloc(&p, sizeof(QbEngineDev));
        if (ce != cudaSuccess) { delete e; QB_FAIL(QB_E_ALLOC, "cudaMalloc failed"); }
        e->owned.push_back(p); e->d = static_cast<QbEngineDev*>(p);
    }
    if (cudaStreamCreate(&e->stream) != cudaSuccess || cudaEventCreate(&e->ev0) != cudaSuccess ||
        cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreate(&e->ev_chunk[0]) != cudaSuccess || cudaEventCreate(&e->ev_chunk[1]) != cudaSuccess ||
        cudaMallocHost((void**)&e->h_active, 2 * sizeof(int)) != cudaSuccess) {
        delete e; QB_FAIL(QB_E_CUDA, "stream/event creation failed");
    }
#undef QB_TRY
    *out = e;
    return QB_OK;
}

// enqueue rounds until every slot is idle; one host sync per chunk
static int qb_drive(QbEngH* e, int nslots_used, bool short_call = false) {
    const int ntiles = e->h.ctl.ntiles;
    const long long grid1 = (long long)((nslots_used + QB_G - 1) / QB_G) * ntiles;
    const int grid2 = (nslots_used * 32 + 127) / 128;
    if (grid1 > 0x7fffffffLL) QB_FAIL(QB_E_ARG, "grid too large");
    // one SELL operator shared by >= 8 trajectory slots: stage it per CTA (qb_pass_kernel_shared)
    const bool use_shared = e->h.ctl.nelem == 1 && e->h.elem[0].fmt == QB_FMT_SELL && !e->h.zbuf &&
                            nslots_used >= 8 && !e->no_shared;
    const long long grid_sh = (long long)((nslots_used + 7) / 8) * ((e->h.nslices + QB_SH_T - 1) / QB_SH_T);
    // large systems in few slots: partial sums and controller fused in one launch per round.
    // Measured on C2 (tools/prof_run.py c2): 4 us per round SLOWER than the 32-CTA reduction +
    // single-warp controller (one SM reads 786 KB of partials alone) -- opt-in only.
    const bool big_control = e->h.red_final != nullptr && nslots_used <= 64 && getenv("QB_BIG_CONTROL");
    // the list of active slots is compacted every round when there are many slots (the tail of an
    // mcsolve batch); few-slot systems keep the identity list and skip the extra launch
    const bool compact = e->tile_g && nslots_used >= 64;
    if (e->tile_g && !compact && e->act_identity != nslots_used) {
        std::vector<int> idl(nslots_used);
        for (int i = 0; i < nslots_used; i++) idl[i] = i;
        QB_CUDA(cudaMemcpy(e->h.act_list, idl.data(), idl.size() * sizeof(int), cudaMemcpyHostToDevice));
        QB_CUDA(cudaMemcpy(e->h.act_count, &nslots_used, sizeof(int), cudaMemcpyHostToDevice));
        e->act_identity = nslots_used;
    }
    if (compact) e->act_identity = 0;
    long long rounds = 0;
    QB_CUDA(cudaMemsetAsync(e->h.vec_count, 0, sizeof(unsigned long long), e->stream));
    QB_CUDA(cudaEventRecord(e->ev0, e->stream));
    // the very first control launch turns the *_BEGIN entry points into passes
    if (big_control) qb_control_kernel<true><<<nslots_used, 1024, 0, e->stream>>>(e->d);
    else qb_control_kernel<false><<<grid2, 128, 0, e->stream>>>(e->d);
    QB_LAUNCH_CHECK();
    if (compact) { qb_compact_kernel<<<1, 1024, 0, e->stream>>>(e->d, nslots_used); QB_LAUNCH_CHECK(); }
    auto enqueue_round = [&](bool timed) -> int {
        cudaEvent_t pa = nullptr, pb = nullptr;
        if (timed) {
            if (cudaEventCreate(&pa) != cudaSuccess || cudaEventCreate(&pb) != cudaSuccess)
                QB_FAIL(QB_E_CUDA, "event creation failed");
            e->prof_events.push_back(pa); e->prof_events.push_back(pb);
            cudaEventRecord(pa, e->stream);
        }
        if (e->h.zbuf) {
            int rcg = qb_launch_dense_rhs(e->stream, e->h.elem[0].dense, e->h.ctl.N,
                                          (const void* const*)e->h.xcols, (void* const*)e->h.zcols,
                                          nslots_used);
            if (rcg) return rcg;
        }
        if (e->tile_g) {
            const int nt = (e->h.ctl.N + e->tile_rows - 1) / e->tile_rows;
            unsigned grid = (unsigned)(nslots_used * nt);
            if (e->tile_persist > 0) {
                int dev = 0, sms = 148;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                grid = std::min<unsigned>(grid, (unsigned)(sms * e->tile_persist));
            }
            if (e->cdesc.n > 0)
                qb_pass_tile_kernel<true, false><<<grid, e->tile_threads, e->tile_smem, e->stream>>>(
                    e->d, nslots_used, e->tile_rows, e->tile_ns, e->tile_persist > 0, e->targs, e->cdesc, 1);
            else
                qb_pass_tile_kernel<false, false><<<grid, e->tile_threads, e->tile_smem, e->stream>>>(
                    e->d, nslots_used, e->tile_rows, e->tile_ns, e->tile_persist > 0, e->targs, e->cdesc, 1);
        }
        else if (use_shared) qb_pass_kernel_shared<<<(unsigned)grid_sh, QB_TILE_ROWS, 0, e->stream>>>(e->d);
        else qb_pass_kernel<<<(unsigned)grid1, QB_TILE_ROWS, 0, e->stream>>>(e->d, nslots_used);
        QB_LAUNCH_CHECK();
        if (e->h.linmap) {
            qb_linmap_kernel<<<(unsigned)((long long)nslots_used * ntiles), QB_TILE_ROWS, 0, e->stream>>>(e->d, nslots_used);
            QB_LAUNCH_CHECK();
        }
        if (timed) {
            cudaEventRecord(pb, e->stream);
            // cumulative algorithmic vector accesses ISSUED up to here: the controller that follows
            // adds the next round's, so the difference of two samples belongs to one pass launch
            const size_t idx = e->prof_events.size() / 2 - 1;
            if (idx < e->prof_vec_cap)
                cudaMemcpyAsync(e->prof_vec_host + idx, e->h.vec_count, sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, e->stream);
        }
        if (big_control) qb_control_kernel<true><<<nslots_used, 1024, 0, e->stream>>>(e->d);
        else {
            if (e->h.red_final) {
                qb_partials_reduce_kernel<<<nslots_used * QB_RED_CTAS, 256, 0, e->stream>>>(e->d);
                QB_LAUNCH_CHECK();
            }
            qb_control_kernel<false><<<grid2, 128, 0, e->stream>>>(e->d);
        }
        QB_LAUNCH_CHECK();
        if (compact) { qb_compact_kernel<<<1, 1024, 0, e->stream>>>(e->d, nslots_used); QB_LAUNCH_CHECK(); }
        return QB_OK;
    };
    if (e->profiling && !e->prof_vec_host) {
        e->prof_vec_cap = 8192;
        if (cudaMallocHost((void**)&e->prof_vec_host, e->prof_vec_cap * sizeof(unsigned long long)) != cudaSuccess) {
            e->prof_vec_host = nullptr; e->prof_vec_cap = 0;
        }
    }
    // few slots of an all-RSELL system: the cooperative multi-round kernel -- one launch runs
    // rounds (pass, partial sums, controller) behind grid barriers until every slot has retired,
    // for batched runs and for the Integrator protocol alike (no idle rounds, no host polling)
    // Measured (tools/coop_ab.sh, tools/plugin_c1.py) and left OFF: a grid barrier among tens to
    // hundreds of CTAs costs as much as a kernel boundary inside a CUDA graph -- C4 (57 tiles)
    // 58.5 -> 71.8 ms, C2 (4096 tiles) 8.9 k -> 6.8 k RHS evaluations/s -- and for a one-CTA grid
    // (C1, N = 400) the plug-in wall time is unchanged within noise (13.9-14.0 ms either way: the
    // round is bound by its dependent loads, not by the launches).  QB_COOP_MAX_CTAS=n enables it
    // for grids of up to n CTAs in a build with -DQB_ENABLE_COOP (QB_NVCC_EXTRA); the default
    // build does not instantiate it.
    const int coop_nt = e->tile_g ? (e->h.ctl.N + e->tile_rows - 1) / e->tile_rows : 0;
    static const int coop_max_ctas = getenv("QB_COOP_MAX_CTAS") ? atoi(getenv("QB_COOP_MAX_CTAS")) : 0;
    const bool coop = e->tile_g && e->coop_occ > 0 && !compact && !e->profiling && !e->h.linmap &&
                      !e->h.zbuf && !big_control && nslots_used <= QB_COOP_MAX_SLOTS &&
                      (long long)nslots_used * coop_nt <= coop_max_ctas;
#ifdef QB_ENABLE_COOP
    if (coop) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int nt = coop_nt;
        unsigned grid = (unsigned)std::min<long long>((long long)nslots_used * nt, (long long)sms * e->coop_occ);
        grid = std::max(grid, 1u);
        QB_CUDA(cudaMemsetAsync(e->targs.coop_rounds, 0, sizeof(int), e->stream));
        for (;;) {
            const QbEngineDev* dE = e->d;
            int a_slots = nslots_used, a_rows = e->tile_rows, a_nsb = e->tile_ns, a_persist = 1, a_max = 4096;
            void* args[] = {(void*)&dE, (void*)&a_slots, (void*)&a_rows, (void*)&a_nsb, (void*)&a_persist,
                            (void*)&e->targs, (void*)&e->cdesc, (void*)&a_max};
            const void* fn = e->cdesc.n > 0 ? (const void*)qb_pass_tile_kernel<true, true>
                                            : (const void*)qb_pass_tile_kernel<false, true>;
            QB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(e->tile_threads), args, e->tile_smem, e->stream));
            g_qb_launches++;
            QB_CUDA(cudaMemcpyAsync(e->h_active, e->h.n_active, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
            QB_CUDA(cudaStreamSynchronize(e->stream));
            if (*e->h_active <= 0) break;
            if (rounds > 2000000000LL) QB_FAIL(QB_E_STATE, "engine did not terminate");
            rounds += a_max;
        }
        int done = 0;
        QB_CUDA(cudaMemcpy(&done, e->targs.coop_rounds, sizeof(int), cudaMemcpyDeviceToHost));
        rounds = done;
    } else
#else
    (void)coop;
#endif
    if (e->profiling || short_call) {
        // plain launches, one host look at the counter per chunk.  Used for per-pass
        // CUDA-event timing and for the Integrator protocol (qb_integ_*), whose calls often
        // need only one or two rounds (an interpolation) -- a 16-round graph would be waste
        // (measured: continuing on the graph after the first chunk costs more in idle rounds
        // than it saves in launches).
        int chunk = e->profiling ? 8 : 2;
        for (;;) {
            for (int i = 0; i < chunk; i++) { int rc = enqueue_round(e->profiling != 0); if (rc) return rc; }
            rounds += chunk;
            QB_CUDA(cudaMemcpyAsync(e->h_active, e->h.n_active, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
            QB_CUDA(cudaStreamSynchronize(e->stream));
            if (*e->h_active <= 0) break;
            if (chunk < 64) chunk *= 2;
            if (rounds > 2000000000LL) QB_FAIL(QB_E_STATE, "engine did not terminate");
        }
    } else {
        // CUDA graph of QB_GRAPH_ROUNDS rounds (all state lives in device memory, so the
        // launch parameters never change), re-launched with a depth-2 pipeline: chunk k+1 is
        // enqueued before the host looks at chunk k's counter, so the GPU never idles on the
        // host round trip and the CPU pays one launch per QB_GRAPH_ROUNDS rounds.
        if (!e->graph || e->graph_slots != nslots_used) {
            if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
            cudaGraph_t g = nullptr;
            QB_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            int rcc = QB_OK;
            for (int i = 0; i < QB_GRAPH_ROUNDS && !rcc; i++) rcc = enqueue_round(false);
            cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
            if (rcc) { if (g) cudaGraphDestroy(g); return rcc; }
            if (ce != cudaSuccess) QB_FAIL(QB_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&e->graph, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) { e->graph = nullptr; QB_FAIL(QB_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce)); }
            e->graph_slots = nslots_used;
            g_qb_launches -= (long long)QB_GRAPH_ROUNDS * (2 + (compact ? 1 : 0) + (e->h.zbuf ? 1 : 0) + ((e->h.red_final && !big_control) ? 1 : 0) + (e->h.linmap ? 1 : 0));
        }
        int pending = 0;              // chunks enqueued whose counter has not been read
        for (;;) {
            const int buf = (int)((rounds / QB_GRAPH_ROUNDS) & 1);
            QB_CUDA(cudaGraphLaunch(e->graph, e->stream));
            g_qb_launches += (long long)QB_GRAPH_ROUNDS * (2 + (compact ? 1 : 0) + (e->h.zbuf ? 1 : 0) + ((e->h.red_final && !big_control) ? 1 : 0) + (e->h.linmap ? 1 : 0));
            QB_CUDA(cudaMemcpyAsync(e->h_active + buf, e->h.n_active, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
            QB_CUDA(cudaEventRecord(e->ev_chunk[buf], e->stream));
            rounds += QB_GRAPH_ROUNDS;
            pending++;
            if (pending == 2) {      // wait for the older chunk only
                QB_CUDA(cudaEventSynchronize(e->ev_chunk[buf ^ 1]));
                pending--;
                if (e->h_active[buf ^ 1] <= 0) break;
            }
            if (rounds > 2000000000LL) QB_FAIL(QB_E_STATE, "engine did not terminate");
        }
        QB_CUDA(cudaStreamSynchronize(e->stream));
    }
    QB_CUDA(cudaEventRecord(e->ev1, e->stream));
    QB_CUDA(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->last_rounds = rounds; e->last_ms = ms;
    if (e->profiling) {
        double tot = 0.0;
        for (size_t i = 0; i + 1 < e->prof_events.size(); i += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, e->prof_events[i], e->prof_events[i + 1]);
            tot += t;
        }
        e->prof_round_ms.clear(); e->prof_round_vec.clear();
        for (size_t i = 0; i + 1 < e->prof_events.size(); i += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, e->prof_events[i], e->prof_events[i + 1]);
            e->prof_round_ms.push_back(t);
            e->prof_round_vec.push_back(i / 2 < e->prof_vec_cap ? e->prof_vec_host[i / 2] : 0ull);
        }
        e->prof_pass_ms = tot; e->prof_pass_launches = (long long)(e->prof_events.size() / 2);
        for (auto ev : e->prof_events) cudaEventDestroy(ev);
        e->prof_events.clear();
        QB_CUDA(cudaMemcpy(&e->prof_vec_count, e->h.vec_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    return QB_OK;
}

static int qb_run_common(QbEngH* e, int mode, int64_t ntraj,
                         const void* d_init, const int32_t* d_init_map,
                         const double* d_tlist, int nt, const void* d_args,
                         const double* d_draws, int ndraws,
                         void* d_expect, int32_t* d_status, int32_t* d_ncol, double* d_col_t,
                         int32_t* d_col_which, int32_t* d_stats, void* d_states) {
    QbEngineDev& h = e->h;
    e->traj0_valid = false; e->d_copy_valid = false;
    if (h.ctl.has_host_coef) QB_FAIL(QB_E_STATE, "host-evaluated (python) coefficients are only supported by the integrator protocol, not by batched runs");
    if (mode == 1 && h.ctl.ncops == 0) QB_FAIL(QB_E_STATE, "mcsolve mode needs collapse operators");
    if (mode == 1 && !d_draws && !e->opt.no_jump) QB_FAIL(QB_E_ARG, "mcsolve mode needs the threshold table");
    if (h.ctl.neops > 0 && !d_expect) QB_FAIL(QB_E_ARG, "system has e_ops but no expect output buffer");
    if (e->opt.store_states && !d_states) QB_FAIL(QB_E_ARG, "store_states set but no states buffer");
    if (mode == 1 && (!d_ncol || !d_col_t || !d_col_which)) QB_FAIL(QB_E_ARG, "mcsolve mode needs collapse output buffers");
    h.ctl.nt = nt; h.ctl.ndraws = ndraws;
    h.ctl.neops = (int)e->sys->eops.size();      // (the Integrator protocol zeroes these)
    h.ctl.opt = e->opt;
    h.ctl.tlist = d_tlist; h.ctl.draws = d_draws;
    h.ctl.args = static_cast<const qb_c128*>(d_args);
    h.ctl.out_expect = static_cast<qb_c128*>(d_expect);
    h.ctl.out_ncol = d_ncol; h.ctl.out_col_t = d_col_t; h.ctl.out_col_which = d_col_which;
    h.init_states = static_cast<const double2*>(d_init);
    h.init_map = d_init_map;
    h.out_states = static_cast<double2*>(d_states);
    h.out_status = d_status; h.out_stats = d_stats;
    h.ntraj_total = (int)ntraj; h.mode = mode;
    const int used = (int)std::min<int64_t>(e->nslots, ntraj);
    // initial slots: trajectory i in slot i
    std::vector<QbTraj> tr(used);
    const int S = h.ctl.tab.S;
    for (int i = 0; i < used; i++) {
        QbTraj& c = tr[i];
        memset(&c, 0, sizeof c);
        c.traj_id = i; c.init_idx = 0; c.mode = mode; c.tl_idx = 0; c.tl_end = nt;
        c.sP = S; c.sF = S + 1; c.sI = S + 2; c.sTA = S + 3; c.sTB = S + 4; c.sY = S + 1;
        c.pc = mode ? QB_PC_MC_BEGIN : QB_PC_ME_BEGIN;
    }
    if (d_init_map) {
        std::vector<int32_t> im(used);
        QB_CUDA(cudaMemcpy(im.data(), d_init_map, used * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < used; i++) tr[i].init_idx = im[i];
    }
    QB_CUDA(cudaMemsetAsync(h.traj, 0, (size_t)e->nslots * sizeof(QbTraj), e->stream));
    QB_CUDA(cudaMemsetAsync(h.pass, 0, (size_t)e->nslots * sizeof(QbPass), e->stream));
    QB_CUDA(cudaMemcpyAsync(h.traj, tr.data(), used * sizeof(QbTraj), cudaMemcpyHostToDevice, e->stream));
    int head = used, act = used;
    QB_CUDA(cudaMemcpyAsync(h.queue_head, &head, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    QB_CUDA(cudaMemcpyAsync(h.n_active, &act, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    QB_CUDA(cudaMemcpyAsync(e->d, &h, sizeof(QbEngineDev), cudaMemcpyHostToDevice, e->stream));
    QB_CUDA(cudaStreamSynchronize(e->stream));   // host vectors above go out of scope
    return qb_drive(e, used);
}

extern "C" int qb_engine_run_device(qb_handle eng, int mode, int64_t ntraj,
                         const void* d_init_states, int64_t ninit, const int32_t* d_init_map,
                         const double* d_tlist, int nt, const void* d_args,
                         const double* d_draws, int ndraws,
                         void* d_expect, int32_t* d_status, int32_t* d_ncol, double* d_col_t,
                         int32_t* d_col_which, int32_t* d_stats, void* d_states) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    if (ntraj < 1 || nt < 1 || !d_init_states || !d_tlist || ninit < 1) QB_FAIL(QB_E_ARG, "bad run arguments");
    if (e->device >= 0) QB_CUDA(cudaSetDevice(e->device));
    if (e->sys->nargs > 0 && !d_args) QB_FAIL(QB_E_ARG, "system has args but none were given");
    return qb_run_common(e, mode, ntraj, d_init_states, d_init_map, d_tlist, nt, d_args, d_draws,
                         ndraws, d_expect, d_status, d_ncol, d_col_t, d_col_which, d_stats, d_states);
}

namespace {
// input / output staging of qb_engine_run: owned by the engine and only ever grown, so that
// repeated runs do not pay a dozen (synchronising) cudaMalloc / cudaFree pairs each
struct DevBuf {
    void*& p;
    size_t& cap;
    DevBuf(void*& p_, size_t& cap_) : p(p_), cap(cap_) {}
    int alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return -1; }
        cap = bytes;
        return 0;
    }
};
}

extern "C" int qb_engine_run(qb_handle eng, int mode, int64_t ntraj,
                  const void* init_states, int64_t ninit, const int32_t* init_map,
                  const double* tlist, int nt,
                  const void* args, const double* draws, int ndraws,
                  void* expect, int32_t* status, int32_t* ncol, double* col_t,
                  int32_t* col_which, int32_t* stats, void* states, void* final_states) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    if (ntraj < 1 || nt < 1 || !init_states || !tlist || ninit < 1) QB_FAIL(QB_E_ARG, "bad run arguments");
    if (e->device >= 0) QB_CUDA(cudaSetDevice(e->device));
    const size_t N = (size_t)e->sys->N;
    // (the Integrator protocol zeroes h.ctl.neops; the system's own count is authoritative)
    const int neops = (int)e->sys->eops.size(), nargs = e->sys->nargs, maxcol = e->opt.max_collapses;
    e->last_ntraj = 0; e->last_nt = 0;
    if (nargs > 0 && !args) QB_FAIL(QB_E_ARG, "system has args but none were given");
    if (final_states && ntraj > e->nslots) QB_FAIL(QB_E_ARG, "final_states needs nslots >= ntraj");
    DevBuf b_init(e->io_p[0], e->io_cap[0]), b_map(e->io_p[1], e->io_cap[1]), b_tl(e->io_p[2], e->io_cap[2]),
        b_args(e->io_p[3], e->io_cap[3]), b_draws(e->io_p[4], e->io_cap[4]), b_exp(e->io_p[5], e->io_cap[5]),
        b_st(e->io_p[6], e->io_cap[6]), b_ncol(e->io_p[7], e->io_cap[7]), b_ct(e->io_p[8], e->io_cap[8]),
        b_cw(e->io_p[9], e->io_cap[9]), b_stats(e->io_p[10], e->io_cap[10]), b_states(e->io_p[11], e->io_cap[11]);
#define QB_A(buf, bytes) do { if (buf.alloc(bytes)) QB_FAIL(QB_E_ALLOC, "device allocation of %zu bytes failed", (size_t)(bytes)); } while (0)
    QB_A(b_init, (size_t)ninit * N * 16);
    QB_CUDA(cudaMemcpy(b_init.p, init_states, (size_t)ninit * N * 16, cudaMemcpyHostToDevice));
    if (init_map) { QB_A(b_map, ntraj * 4); QB_CUDA(cudaMemcpy(b_map.p, init_map, ntraj * 4, cudaMemcpyHostToDevice)); }
    QB_A(b_tl, (size_t)nt * 8);
    QB_CUDA(cudaMemcpy(b_tl.p, tlist, (size_t)nt * 8, cudaMemcpyHostToDevice));
    if (nargs > 0) { QB_A(b_args, (size_t)ntraj * nargs * 16); QB_CUDA(cudaMemcpy(b_args.p, args, (size_t)ntraj * nargs * 16, cudaMemcpyHostToDevice)); }
    if (draws && ndraws > 0) { QB_A(b_draws, (size_t)ntraj * ndraws * 8); QB_CUDA(cudaMemcpy(b_draws.p, draws, (size_t)ntraj * ndraws * 8, cudaMemcpyHostToDevice)); }
    QB_A(b_exp, (size_t)ntraj * std::max(1, neops) * nt * 16);
    QB_CUDA(cudaMemset(b_exp.p, 0, (size_t)ntraj * std::max(1, neops) * nt * 16));
    QB_A(b_st, ntraj * 4); QB_CUDA(cudaMemset(b_st.p, 0, ntraj * 4));
    QB_A(b_ncol, ntraj * 4); QB_CUDA(cudaMemset(b_ncol.p, 0, ntraj * 4));
    QB_A(b_ct, (size_t)ntraj * maxcol * 8); QB_A(b_cw, (size_t)ntraj * maxcol * 4);
    QB_CUDA(cudaMemset(b_ct.p, 0, (size_t)ntraj * maxcol * 8));
    QB_CUDA(cudaMemset(b_cw.p, 0, (size_t)ntraj * maxcol * 4));
    QB_A(b_stats, ntraj * 16); QB_CUDA(cudaMemset(b_stats.p, 0, ntraj * 16));
    if (e->opt.store_states) {
        if (!states) QB_FAIL(QB_E_ARG, "store_states set but states == NULL");
        QB_A(b_states, (size_t)ntraj * nt * N * 16);
    }
#undef QB_A
    // the staging buffers are cached: a buffer this run did not fill must not be passed on
    int rc = qb_run_common(e, mode, ntraj, b_init.p, init_map ? static_cast<int32_t*>(b_map.p) : nullptr,
                           static_cast<double*>(b_tl.p), nt, nargs > 0 ? b_args.p : nullptr,
                           (draws && ndraws > 0) ? static_cast<double*>(b_draws.p) : nullptr, ndraws, b_exp.p,
                           static_cast<int32_t*>(b_st.p), static_cast<int32_t*>(b_ncol.p),
                           static_cast<double*>(b_ct.p), static_cast<int32_t*>(b_cw.p),
                           static_cast<int32_t*>(b_stats.p), e->opt.store_states ? b_states.p : nullptr);
    if (rc) return rc;
    e->last_ntraj = ntraj; e->last_nt = nt;
    if (expect && neops > 0) QB_CUDA(cudaMemcpy(expect, b_exp.p, (size_t)ntraj * neops * nt * 16, cudaMemcpyDeviceToHost));
    if (status) QB_CUDA(cudaMemcpy(status, b_st.p, ntraj * 4, cudaMemcpyDeviceToHost));
    if (ncol) QB_CUDA(cudaMemcpy(ncol, b_ncol.p, ntraj * 4, cudaMemcpyDeviceToHost));
    if (col_t) QB_CUDA(cudaMemcpy(col_t, b_ct.p, (size_t)ntraj * maxcol * 8, cudaMemcpyDeviceToHost));
    if (col_which) QB_CUDA(cudaMemcpy(col_which, b_cw.p, (size_t)ntraj * maxcol * 4, cudaMemcpyDeviceToHost));
    if (stats) QB_CUDA(cudaMemcpy(stats, b_stats.p, ntraj * 16, cudaMemcpyDeviceToHost));
    if (states && e->opt.store_states) QB_CUDA(cudaMemcpy(states, b_states.p, (size_t)ntraj * nt * N * 16, cudaMemcpyDeviceToHost));
    if (final_states) {
        // slot i holds trajectory i (ntraj <= nslots): copy V[sY]
        std::vector<QbTraj> tr(ntraj);
        QB_CUDA(cudaMemcpy(tr.data(), e->h.traj, ntraj * sizeof(QbTraj), cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < ntraj; j++) {
            const QbTraj& c = tr[j];
            const double2* src = e->h.pool + ((size_t)j * e->V + c.sY) * N;
            QB_CUDA(cudaMemcpy(static_cast<char*>(final_states) + (size_t)c.traj_id * N * 16, src, N * 16, cudaMemcpyDeviceToHost));
        }
    }
    return QB_OK;
}

int qb_reduce_expect_on(cudaStream_t stream, const void* d_expect, int64_t ntraj, int neops, int nt, void* d_sums) {
    const int n = neops * nt;
    if (n <= 0 || ntraj < 0 || !d_expect || !d_sums) QB_FAIL(QB_E_ARG, "bad reduce arguments");
    qb_reduce_expect_kernel<<<(n + 127) / 128, 128, 0, stream>>>(static_cast<const double2*>(d_expect), ntraj, n,
                                                                  static_cast<double2*>(d_sums));
    QB_LAUNCH_CHECK();
    return QB_OK;
}
extern "C" int qb_reduce_expect(const void* d_expect, int64_t ntraj, int neops, int nt, void* d_sums) {
    return qb_reduce_expect_on(nullptr, d_expect, ntraj, neops, nt, d_sums);
}
// where the per-trajectory expectation values of the last qb_engine_run live (qb_comm.cu)
int qb_engine_expect_view(qb_handle eng, const void** d_expect, int* neops, int* device, int64_t* ntraj, int* nt) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    if (e->last_ntraj < 1 || !e->io_p[5]) QB_FAIL(QB_E_STATE, "engine has no finished batched run");
    *d_expect = e->io_p[5]; *neops = (int)e->sys->eops.size(); *device = e->device;
    *ntraj = e->last_ntraj; *nt = e->last_nt;
    return QB_OK;
}

extern "C" int qb_engine_last_run_info(qb_handle eng, int64_t* rounds, double* gpu_ms) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    if (rounds) *rounds = e->last_rounds;
    if (gpu_ms) *gpu_ms = e->last_ms;
    return QB_OK;
}

// ---- Integrator protocol on slot 0 ----
static int qb_integ_traj(QbEngH* e, QbTraj* c) {
    if (!e->traj0_valid) {
        QB_CUDA(cudaMemcpy(&e->traj0, e->h.traj, sizeof(QbTraj), cudaMemcpyDeviceToHost));
        e->traj0_valid = true;
    }
    *c = e->traj0;
    return QB_OK;
}
static int qb_integ_prepare(QbEngH* e) {
    QbEngineDev& h = e->h;
    if (e->device >= 0) QB_CUDA(cudaSetDevice(e->device));
    if (!e->d_tlist) { QB_CUDA(cudaMalloc(&e->d_tlist, 8)); }
    if (!e->d_init) { QB_CUDA(cudaMalloc(&e->d_init, (size_t)e->sys->N * 16)); }
    if (e->sys->nargs > 0 && !e->d_args) {
        QB_CUDA(cudaMalloc(&e->d_args, (size_t)e->sys->nargs * 16));
        QB_CUDA(cudaMemset(e->d_args, 0, (size_t)e->sys->nargs * 16));
    }
    h.ctl.nt = 1; h.ctl.ndraws = 0; h.ctl.tlist = static_cast<const double*>(e->d_tlist);
    h.ctl.draws = nullptr; h.ctl.args = static_cast<const qb_c128*>(e->d_args);
    h.ctl.out_expect = nullptr; h.ctl.out_ncol = nullptr; h.ctl.out_col_t = nullptr;
    h.ctl.out_col_which = nullptr;
    h.ctl.neops = 0;                   // e_ops are evaluated by the caller in this protocol
    h.ctl.opt.store_states = 0;
    h.init_states = static_cast<const double2*>(e->d_init); h.init_map = nullptr;
    h.out_states = nullptr; h.out_status = nullptr; h.out_stats = nullptr;
    h.ntraj_total = 0; h.mode = 0;
    return QB_OK;
}
static int qb_integ_launch(QbEngH* e, QbTraj& c) {
    QbEngineDev& h = e->h;
    int act = 1, head = 1;
    QB_CUDA(cudaMemcpyAsync(h.traj, &c, sizeof(QbTraj), cudaMemcpyHostToDevice, e->stream));
    QB_CUDA(cudaMemsetAsync(h.pass, 0, sizeof(QbPass), e->stream));
    QB_CUDA(cudaMemcpyAsync(h.queue_head, &head, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    QB_CUDA(cudaMemcpyAsync(h.n_active, &act, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (!e->d_copy_valid || memcmp(&e->d_copy, &h, sizeof(QbEngineDev)) != 0) {
        QB_CUDA(cudaMemcpyAsync(e->d, &h, sizeof(QbEngineDev), cudaMemcpyHostToDevice, e->stream));
        QB_CUDA(cudaStreamSynchronize(e->stream));      // &h must stay intact until copied
        memcpy(&e->d_copy, &h, sizeof(QbEngineDev));
        e->d_copy_valid = true;
    }
    e->traj0_valid = false;
    int rc = qb_drive(e, 1, true);
    if (rc) return rc;
    QB_CUDA(cudaMemcpy(&c, h.traj, sizeof(QbTraj), cudaMemcpyDeviceToHost));
    e->traj0 = c; e->traj0_valid = true;
    return QB_OK;
}

extern "C" int qb_integ_set_args(qb_handle eng, const void* args) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    int rc = qb_integ_prepare(e); if (rc) return rc;
    if (e->sys->nargs > 0) {
        if (!args) QB_FAIL(QB_E_ARG, "null args");
        QB_CUDA(cudaMemcpy(e->d_args, args, (size_t)e->sys->nargs * 16, cudaMemcpyHostToDevice));
    }
    return QB_OK;
}

extern "C" int qb_integ_set_state(qb_handle eng, double t, const void* y) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e || !y) QB_FAIL(QB_E_TYPE, "not an engine handle / null state");
    int rc = qb_integ_prepare(e); if (rc) return rc;
    QB_CUDA(cudaMemcpy(e->d_init, y, (size_t)e->sys->N * 16, cudaMemcpyHostToDevice));
    QB_CUDA(cudaMemcpy(e->d_tlist, &t, 8, cudaMemcpyHostToDevice));
    QbTraj c; memset(&c, 0, sizeof c);
    const int S = e->h.ctl.tab.S;
    c.traj_id = 0; c.init_idx = 0; c.mode = 0; c.tl_idx = 0; c.tl_end = 0;
    c.sP = S; c.sF = S + 1; c.sI = S + 2; c.sTA = S + 3; c.sTB = S + 4; c.sY = S + 1;
    c.status = QB_ST_NORMAL;
    // ME_BEGIN with an empty target list: set_state, then the QL_RECORD/ME_NEXT chain
    // finds nothing to do and pauses.
    c.pc = QB_PC_ME_BEGIN;
    rc = qb_integ_launch(e, c); if (rc) return rc;
    if (c.done < 0) QB_FAIL(QB_E_STATE, "set_state failed with status %d", c.done);
    return c.done == 2 ? 3 : QB_OK;      // 3: coefficient values needed (qb_integ_resume)
}

extern "C" int qb_integ_integrate(qb_handle eng, double t, int step, double* t_out, int* status) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    int rc = qb_integ_prepare(e); if (rc) return rc;
    QbTraj c;
    { int rct = qb_integ_traj(e, &c); if (rct) return rct; }
    if (c.sP == 0 && c.sF == 0) { if (status) *status = QB_ST_NOT_INITIATED; return QB_OK; }
    QB_CUDA(cudaMemcpy(e->d_tlist, &t, 8, cudaMemcpyHostToDevice));
    c.tl_idx = 0; c.tl_end = 1; c.done = 0;
    c.pc = step ? QB_PC_STEP_ENTRY : QB_PC_ME_NEXT;
    rc = qb_integ_launch(e, c); if (rc) return rc;
    if (t_out) *t_out = c.t;
    if (status) *status = c.done == 2 ? 3 : (c.done < 0 ? c.done : c.status);
    return QB_OK;
}

// ---- host-evaluated coefficients (python callables) ----
// After qb_integ_set_state / qb_integ_integrate returned status QB_ST_NEED_COEF (3) the
// controller waits for the values of the host-evaluated elements at *t:
//   qb_integ_pending_coef(eng, &t) ; qb_integ_resume(eng, vals[nelem] complex, &t_out, &status)
extern "C" int qb_integ_pending_coef(qb_handle eng, double* t) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e || !t) QB_FAIL(QB_E_TYPE, "not an engine handle");
    QbTraj c;
    { int rct = qb_integ_traj(e, &c); if (rct) return rct; }
    if (c.done != 2) QB_FAIL(QB_E_STATE, "no coefficient request pending");
    *t = c.hc_t;
    return QB_OK;
}
extern "C" int qb_integ_resume(qb_handle eng, const void* vals, double* t_out, int* status) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e || !vals) QB_FAIL(QB_E_TYPE, "not an engine handle");
    QbTraj c;
    { int rct = qb_integ_traj(e, &c); if (rct) return rct; }
    if (c.done != 2) QB_FAIL(QB_E_STATE, "no coefficient request pending");
    // values of ALL elements are accepted; the controller overwrites the device-evaluated ones
    QB_CUDA(cudaMemcpy(e->h.coef, vals, (size_t)e->h.ctl.nelem * 16, cudaMemcpyHostToDevice));
    c.hc_valid = 1; c.done = 0; c.pc = c.hc_resume;
    int rc = qb_integ_launch(e, c); if (rc) return rc;
    if (t_out) *t_out = c.t;
    if (status) *status = c.done == 2 ? 3 : (c.done < 0 ? c.done : c.status);
    return QB_OK;
}

extern "C" int qb_integ_get_state(qb_handle eng, double* t, void* y) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    QbTraj c;
    { int rct = qb_integ_traj(e, &c); if (rct) return rct; }
    if (t) *t = c.t;
    if (y) {
        const size_t N = (size_t)e->sys->N;
        QB_CUDA(cudaMemcpy(y, e->h.pool + (size_t)c.sY * N, N * 16, cudaMemcpyDeviceToHost));
    }
    return QB_OK;
}

extern "C" int qb_integ_stats(qb_handle eng, int64_t stats[4]) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e || !stats) QB_FAIL(QB_E_TYPE, "not an engine handle");
    QbTraj c;
    { int rct = qb_integ_traj(e, &c); if (rct) return rct; }
    stats[0] = c.n_rhs; stats[1] = c.n_accept; stats[2] = c.n_reject; stats[3] = c.n_pass;
    return QB_OK;
}

// out = sum_k vals[k] A_k x with coefficient values supplied by the caller (python-callable
// coefficients evaluated on the host; used by the zvode-driven Adams path)
extern "C" int qb_engine_rhs_coef(qb_handle eng, const void* vals, qb_handle xh, qb_handle outh) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* o = qb_cast<QbDenseH>(outh, QB_TAG_DENSE);
    if (!e || !x || !o || !vals) QB_FAIL(QB_E_TYPE, "bad handles");
    if (x->size() != e->sys->N || o->size() != e->sys->N) QB_FAIL(QB_E_SHAPE, "incompatible shapes");
    QB_CUDA(cudaMemcpyAsync(e->h.coef, vals, (size_t)e->h.ctl.nelem * 16, cudaMemcpyHostToDevice, e->stream));
    e->d_copy_valid = false;
    QB_CUDA(cudaMemcpyAsync(e->d, &e->h, sizeof(QbEngineDev), cudaMemcpyHostToDevice, e->stream));
    qb_rhs_kernel<<<(e->h.ctl.N + QB_TILE_ROWS - 1) / QB_TILE_ROWS, QB_TILE_ROWS, 0, e->stream>>>(e->d, x->d, o->d, e->h.coef);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaStreamSynchronize(e->stream));
    return QB_OK;
}

extern "C" int qb_engine_rhs(qb_handle eng, double t, qb_handle xh, qb_handle outh) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* o = qb_cast<QbDenseH>(outh, QB_TAG_DENSE);
    if (!e || !x || !o) QB_FAIL(QB_E_TYPE, "bad handles");
    if (x->size() != e->sys->N || o->size() != e->sys->N) QB_FAIL(QB_E_SHAPE, "incompatible shapes");
    // coefficients on the host (same byte-code interpreter)
    std::vector<qb_c128> coef(e->h.ctl.nelem);
    std::vector<qb_c128> zero_args(std::max(1, e->sys->nargs));
    if (e->d_args && e->sys->nargs > 0)
        QB_CUDA(cudaMemcpy(zero_args.data(), e->d_args, (size_t)e->sys->nargs * 16, cudaMemcpyDeviceToHost));
    for (int k = 0; k < e->h.ctl.nelem; k++) {
        const auto& p = e->sys->elem_prog[k];
        if (p.empty()) { coef[k].re = 1.0; coef[k].im = 0.0; }
        else if (qb_eval_prog(p.data(), (int)p.size(), t, zero_args.data(), e->sys->splines.data(),
                              e->sys->spool.data(), &coef[k]))
            QB_FAIL(QB_E_STATE, "coefficient program failed");
    }
    QB_CUDA(cudaMemcpyAsync(e->h.coef, coef.data(), coef.size() * 16, cudaMemcpyHostToDevice, e->stream));
    e->d_copy_valid = false;
    QB_CUDA(cudaMemcpyAsync(e->d, &e->h, sizeof(QbEngineDev), cudaMemcpyHostToDevice, e->stream));
    qb_rhs_kernel<<<(e->h.ctl.N + QB_TILE_ROWS - 1) / QB_TILE_ROWS, QB_TILE_ROWS, 0, e->stream>>>(e->d, x->d, o->d, e->h.coef);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaStreamSynchronize(e->stream));
    return QB_OK;
}

// ---- measurement hooks ----
extern "C" int qb_engine_set_profiling(qb_handle eng, int on) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    e->profiling = on ? 1 : 0;
    return QB_OK;
}
extern "C" int qb_engine_profile(qb_handle eng, double* pass_ms, int64_t* pass_launches,
                                 double* state_vector_accesses) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e) QB_FAIL(QB_E_TYPE, "not an engine handle");
    if (pass_ms) *pass_ms = e->prof_pass_ms;
    if (pass_launches) *pass_launches = e->prof_pass_launches;
    if (state_vector_accesses) *state_vector_accesses = (double)e->prof_vec_count;
    return QB_OK;
}
// per pass launch of the last profiled run: CUDA-event time and the cumulative number of
// algorithmic state-vector accesses issued when it ran (differences = that launch's accesses)
extern "C" int qb_engine_profile_rounds(qb_handle eng, double* ms, double* cum_vec, int64_t max, int64_t* n) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    if (!e || !n) QB_FAIL(QB_E_TYPE, "not an engine handle");
    const int64_t m = std::min<int64_t>(max, (int64_t)e->prof_round_ms.size());
    for (int64_t i = 0; i < m; i++) {
        if (ms) ms[i] = e->prof_round_ms[i];
        if (cum_vec) cum_vec[i] = (double)e->prof_round_vec[i];
    }
    *n = (int64_t)e->prof_round_ms.size();
    return QB_OK;
}
// `iters` back-to-back RHS evaluations out = sum_k c_k(t) A_k x, timed with CUDA events on
// the engine's stream (ms_total covers all iterations)
extern "C" int qb_engine_rhs_bench(qb_handle eng, double t, qb_handle xh, qb_handle outh,
                                   int iters, double* ms_total) {
    QbEngH* e = qb_cast<QbEngH>(eng, QB_TAG_ENG);
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* o = qb_cast<QbDenseH>(outh, QB_TAG_DENSE);
    if (!e || !x || !o || iters < 1) QB_FAIL(QB_E_TYPE, "bad arguments");
    int rc = qb_engine_rhs(eng, t, xh, outh);      // sets coefficients, warm-up
    if (rc) return rc;
    QB_CUDA(cudaEventRecord(e->ev0, e->stream));
    for (int i = 0; i < iters; i++) {
        qb_rhs_kernel<<<(e->h.ctl.N + QB_TILE_ROWS - 1) / QB_TILE_ROWS, QB_TILE_ROWS, 0, e->stream>>>(e->d, x->d, o->d, e->h.coef);
        QB_LAUNCH_CHECK();
    }
    QB_CUDA(cudaEventRecord(e->ev1, e->stream));
    QB_CUDA(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    if (ms_total) *ms_total = ms;
    return QB_OK;
}
