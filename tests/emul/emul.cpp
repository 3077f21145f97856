// tests/emul/emul.cpp -- TEST INFRASTRUCTURE ONLY (never part of libqutip_b200.so).
//
// Host-side unit-test harness for the engine's *control logic and data formats*: it
// compiles the very same headers the CUDA kernels use
//     qutip_b200/csrc/qb_control.h   (step controller / Monte-Carlo state machine)
//     qutip_b200/csrc/qb_coeff.h     (coefficient byte-code)
//     qutip_b200/csrc/qb_diam.h      (CSR/Dia -> diagonal-masked slices)
// with g++ and executes each QbPass with a plain serial loop that follows the pass
// kernel's semantics (qb_engine.cu:qb_pass_kernel).  This lets the no-GPU CI check the
// flattened state machine against the oracle; it is not a product path and is not
// importable from the package.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../qutip_b200/csrc/qb_types.h"
#include "../../qutip_b200/csrc/qb_coeff.h"
#include "../../qutip_b200/csrc/qb_control.h"
#include "../../qutip_b200/csrc/qb_diam.h"
#include "../../qutip_b200/csrc/qb_tableaux.h"

namespace {
struct HostOp {
    int fmt; int nrows, ncols;
    std::vector<qb_c128> val; std::vector<int> col, rowptr;      // CSR
    qbdiam::DiamHost dh;                                           // DIAM
    qbdiam::SellHost sh;                                           // SELL
    qbdiam::RsellHost rs;                                          // RSELL
};
struct Sys {
    int64_t N = 0; int nargs = 0;
    std::vector<HostOp> elems, cops, nops, eops;
    std::vector<std::vector<QbInstr>> elem_prog, cop_prog, nop_prog, eop_prog;
    int eop_functional = 0;
    int mc_trace = 0;
};
Sys g_sys;
// optional log of the pass classes each trajectory issues (scheduler studies): per pass one
// 64-bit key built from roles (P, F, I, TA, TB, k_j, INIT), not physical slot numbers
std::vector<std::vector<uint64_t>> g_class_log;
bool g_log_classes = false;
int g_tile_mode = 0, g_exp_chunk = 0;

static int popc(unsigned x) { return __builtin_popcount(x); }

// mirrors qb_kernels.cuh:qb_rowdot lane by lane
static qb_c128 rowdot(const HostOp& A, int64_t r, const qb_c128* x) {
    qb_c128 acc = {0.0, 0.0};
    if (A.fmt == QB_FMT_CSR) {
        for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; p++) {
            const qb_c128 a = A.val[p], b = x[A.col[p]];
            acc.re += a.re * b.re - a.im * b.im; acc.im += a.re * b.im + a.im * b.re;
        }
        return acc;
    }
    const int sl = (int)(r / 32), lane = (int)(r % 32);
    if (A.fmt == QB_FMT_SELL) {
        for (int k = A.sh.slice_ptr[sl]; k < A.sh.slice_ptr[sl + 1]; k++) {
            const qb_c128 a = A.sh.val[(size_t)k * 32 + lane], b = x[A.sh.col[(size_t)k * 32 + lane]];
            acc.re += a.re * b.re - a.im * b.im; acc.im += a.re * b.im + a.im * b.re;
        }
        return acc;
    }
    if (A.fmt == QB_FMT_RSELL) {
        const int* si = &A.rs.sinfo[(size_t)sl * 4];
        for (int k = si[0]; k < si[0] + (si[1] & 4095); k++) {
            const QbSlotDesc& d = A.rs.desc[k];
            const int cr = d.rule & QB_RS_COL_MASK;
            const long long c = cr == QB_RS_COL_ADD ? r + d.delta : cr == QB_RS_COL_XOR ? (r ^ (long long)d.delta)
                                                                   : A.rs.col[(size_t)(si[3] + d.cpos) * 32 + lane];
            qb_c128 a = {d.vre, d.vim};
            if (!(d.rule & QB_RS_VAL_CONST)) a = A.rs.val[(size_t)(si[2] + d.vpos) * 32 + lane];
            const qb_c128 b = x[c];
            acc.re += a.re * b.re - a.im * b.im; acc.im += a.re * b.im + a.im * b.re;
        }
        return acc;
    }
    long long vb = A.dh.slice_vbase[sl];
    const unsigned lt = (1u << lane) - 1u;
    for (int e = A.dh.slice_ptr[sl]; e < A.dh.slice_ptr[sl + 1]; e++) {
        const unsigned m = (unsigned)A.dh.ent[e].y;
        if ((m >> lane) & 1u) {
            const qb_c128 a = A.dh.val[vb + popc(m & lt)], b = x[r + A.dh.ent[e].x];
            acc.re += a.re * b.re - a.im * b.im; acc.im += a.re * b.im + a.im * b.re;
        }
        vb += popc(m);
    }
    return acc;
}

static HostOp make_op(const qb_c128* data, const int32_t* col, const int32_t* rowptr,
                      int64_t rows, int64_t cols, int fmt) {
    HostOp o; o.fmt = fmt; o.nrows = (int)rows; o.ncols = (int)cols;
    const int64_t nnz = rowptr[rows];
    if (fmt == QB_FMT_CSR) {
        o.val.assign(data, data + nnz); o.col.assign(col, col + nnz);
        o.rowptr.assign(rowptr, rowptr + rows + 1);
    } else if (fmt == QB_FMT_SELL) {
        qbdiam::build_sell(rows, cols, [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o2) {
            for (int p = rowptr[r]; p < rowptr[r + 1]; p++) o2.push_back({col[p], data[p]});
        }, o.sh);
    } else if (fmt == QB_FMT_RSELL) {
        qbdiam::build_rsell(rows, cols, [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o2) {
            for (int p = rowptr[r]; p < rowptr[r + 1]; p++) o2.push_back({col[p], data[p]});
        }, o.rs);
    } else {
        qbdiam::build_diam(rows, [&](int64_t sl, std::vector<qbdiam::Entry>& es) {
            const int64_t r0 = sl * 32, r1 = std::min<int64_t>(rows, r0 + 32);
            for (int64_t r = r0; r < r1; r++)
                for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                    qbdiam::Entry e; e.off = (int)(col[p] - r); e.lane = (int)(r - r0); e.v = data[p];
                    es.push_back(e);
                }
        }, o.dh);
    }
    return o;
}
static std::vector<QbInstr> prog(const QbInstr* p, int n) { return std::vector<QbInstr>(p, p + n); }
}  // namespace

extern "C" {
void emul_reset(int64_t N, int nargs) { g_sys = Sys(); g_sys.N = N; g_sys.nargs = nargs; }
void emul_add_element(const void* data, const int32_t* col, const int32_t* rowptr, int fmt,
                      const QbInstr* p, int np) {
    g_sys.elems.push_back(make_op((const qb_c128*)data, col, rowptr, g_sys.N, g_sys.N, fmt));
    g_sys.elem_prog.push_back(prog(p, np));
}
void emul_add_collapse(const void* cd, const int32_t* cc, const int32_t* cr,
                       const void* nd, const int32_t* nc, const int32_t* nr, int fmt) {
    g_sys.cops.push_back(make_op((const qb_c128*)cd, cc, cr, g_sys.N, g_sys.N, fmt));
    g_sys.nops.push_back(make_op((const qb_c128*)nd, nc, nr, g_sys.N, g_sys.N, fmt));
    g_sys.cop_prog.push_back({}); g_sys.nop_prog.push_back({});
}
// collapse operator with time-dependent coefficients: c(t) C and n(t) C^dagger C
void emul_add_collapse_td(const void* cd, const int32_t* cc, const int32_t* cr,
                          const void* nd, const int32_t* nc, const int32_t* nr, int fmt,
                          const QbInstr* cp, int ncp, const QbInstr* npg, int nnp) {
    g_sys.cops.push_back(make_op((const qb_c128*)cd, cc, cr, g_sys.N, g_sys.N, fmt));
    g_sys.nops.push_back(make_op((const qb_c128*)nd, nc, nr, g_sys.N, g_sys.N, fmt));
    g_sys.cop_prog.push_back(prog(cp, ncp)); g_sys.nop_prog.push_back(prog(npg, nnp));
}
void emul_add_eop(const void* d, const int32_t* c, const int32_t* r, int fmt) {
    g_sys.eops.push_back(make_op((const qb_c128*)d, c, r, g_sys.N, g_sys.N, fmt));
    g_sys.eop_prog.push_back({});
}
void emul_set_functional(int f) { g_sys.eop_functional = f; }
void emul_set_mc_trace(int n) { g_sys.mc_trace = n; }
double emul_diam_avg_lanes(int which) {
    const HostOp& o = g_sys.elems[which];
    return o.dh.ent.empty() ? 0.0 : (double)o.dh.val.size() / (double)o.dh.ent.size();
}
// RSELL statistics of element `which`: out = {slots, explicit column blocks, explicit value
// blocks, device bytes, xor-rule slots, de-duplicated descriptors}
void emul_rsell_stats(int which, long long* out) {
    const HostOp& o = g_sys.elems[which];
    long long nx = 0;
    for (size_t sl = 0; sl * 4 < o.rs.sinfo.size(); sl++)
        for (int k = o.rs.sinfo[sl * 4]; k < o.rs.sinfo[sl * 4] + (o.rs.sinfo[sl * 4 + 1] & 4095); k++)
            nx += (o.rs.desc[k].rule & QB_RS_COL_MASK) == QB_RS_COL_XOR;
    out[0] = o.rs.stored(); out[1] = (long long)o.rs.col.size() / 32;
    out[2] = (long long)o.rs.val.size() / 32; out[3] = o.rs.bytes(); out[4] = nx;
    out[5] = (long long)o.rs.desc.size();
}
// Adams corrector coefficients / error constants of order nq (qb_adams.h)
void emul_adams_table(int nq, double* el, double* tq) {
    static QbTableau T;
    qb_adams_table(&T);
    for (int j = 0; j <= nq; j++) el[j] = T.a[nq][j];
    for (int j = 0; j < 3; j++) tq[j] = T.bi[nq][j];
}
// y = A x with the element `which` (format check)
void emul_matvec(int which, const void* x, void* y) {
    const HostOp& o = g_sys.elems[which];
    for (int64_t r = 0; r < g_sys.N; r++) ((qb_c128*)y)[r] = rowdot(o, r, (const qb_c128*)x);
}
void emul_set_tile_mode(int on, int exp_chunk) { g_tile_mode = on; g_exp_chunk = exp_chunk; }
void emul_log_classes(int on) { g_log_classes = on != 0; g_class_log.clear(); }
int64_t emul_class_log_len(int traj) { return traj < (int)g_class_log.size() ? (int64_t)g_class_log[traj].size() : 0; }
void emul_class_log_get(int traj, uint64_t* out) { for (size_t i = 0; i < g_class_log[traj].size(); i++) out[i] = g_class_log[traj][i]; }
int emul_eval_prog(const QbInstr* p, int np, double t, const void* args, double out[2]) {
    qb_c128 r; int rc = qb_eval_prog(p, np, t, (const qb_c128*)args, nullptr, nullptr, &r);
    out[0] = r.re; out[1] = r.im; return rc;
}

int emul_run(int mode, int tableau, const QbOptions* opt, int64_t ntraj, int nslots,
             const void* init_states, const int32_t* init_map, const double* tlist, int nt,
             const void* args, const double* draws, int ndraws,
             void* expect, int32_t* status, int32_t* ncol, double* col_t, int32_t* col_which,
             int32_t* stats, void* states, int64_t max_rounds)
{
    Sys& s = g_sys;
    const int64_t N = s.N;
    QbCtl g; memset(&g, 0, sizeof g);
    if (tableau == 3) qb_adams_table(&g.tab);
    else g.tab = *QB_TABLEAUX[tableau];
    g.opt = *opt;
    if (tableau == 3) { g.opt.atol *= QB_AD_TOL_SCALE; g.opt.rtol *= QB_AD_TOL_SCALE; }   // as qb_engine_create
    g.N = (int)N; g.ntiles = 1;
    g.nelem = (int)s.elems.size(); g.ncops = (int)s.cops.size(); g.neops = (int)s.eops.size();
    g.nargs = s.nargs; g.eop_functional = s.eop_functional; g.mc_trace = s.mc_trace;
    g.maxcoef = std::max(1, g.nelem);
    g.nt = nt; g.ndraws = ndraws;
    g.tile_mode = g_tile_mode; g.exp_chunk = g_exp_chunk;
    std::vector<QbInstr> instr;
    auto pack = [&](const std::vector<std::vector<QbInstr>>& ps) {
        std::vector<QbProgRef> refs;
        for (auto& p : ps) { QbProgRef r; r.off = (int)instr.size(); r.len = (int)p.size(); refs.push_back(r); instr.insert(instr.end(), p.begin(), p.end()); }
        return refs;
    };
    auto r_el = pack(s.elem_prog), r_c = pack(s.cop_prog), r_n = pack(s.nop_prog), r_e = pack(s.eop_prog);
    g.elem_prog = r_el.data(); g.cop_prog = r_c.data(); g.nop_prog = r_n.data(); g.eop_prog = r_e.data();
    g.instr = instr.data();
    g.args = (const qb_c128*)args; g.tlist = tlist; g.draws = draws;
    g.out_expect = (qb_c128*)expect; g.out_ncol = ncol; g.out_col_t = col_t; g.out_col_which = col_which;
    const int S = g.tab.S, V = S + 5;
    std::vector<qb_c128> pool((size_t)nslots * V * N);
    std::vector<QbTraj> traj(nslots); std::vector<QbPass> pass(nslots);
    std::vector<QbLinMap> linmap(nslots);
    std::vector<qb_c128> coef((size_t)nslots * g.maxcoef);
    std::vector<double> probs((size_t)nslots * std::max(1, g.ncops));
    std::vector<double> red((size_t)nslots * QB_MAXRED);
    memset(traj.data(), 0, traj.size() * sizeof(QbTraj));
    memset(pass.data(), 0, pass.size() * sizeof(QbPass));
    int64_t head = 0; int active = 0;
    auto start = [&](QbTraj& c, int id) {
        c.traj_id = id; c.init_idx = init_map ? init_map[id] : 0; c.mode = mode;
        c.tl_idx = 0; c.tl_end = nt;
        c.sP = S; c.sF = S + 1; c.sI = S + 2; c.sTA = S + 3; c.sTB = S + 4; c.sY = S + 1;
        c.status = QB_ST_NORMAL; c.done = 0; c.pc = mode ? QB_PC_MC_BEGIN : QB_PC_ME_BEGIN;
    };
    for (int i = 0; i < nslots && head < ntraj; i++) { start(traj[i], (int)head++); active++; }
    auto vsrc = [&](int slot, int idx) -> const qb_c128* {
        if (idx >= 0) return pool.data() + ((size_t)slot * V + idx) * N;
        return (const qb_c128*)init_states + (size_t)traj[slot].init_idx * N;
    };
    auto control = [&]() {
        for (int slot = 0; slot < nslots; slot++) {
            QbTraj& c = traj[slot];
            if (c.pc == QB_PC_IDLE) continue;
            for (;;) {
                int issued = qb_advance(g, g.tab, c, pass[slot], red.data() + (size_t)slot * QB_MAXRED,
                                        coef.data() + (size_t)slot * g.maxcoef,
                                        probs.data() + (size_t)slot * std::max(1, g.ncops),
                                        &linmap[slot]);
                if (issued) {
                    c.n_pass++;
                    if (g_log_classes) {
                        const QbPass& q = pass[slot];
                        auto role = [&](int sidx) -> uint64_t {
                            if (sidx < 0) return 40 + (uint64_t)(-sidx);
                            if (sidx == c.sP) return 30; if (sidx == c.sF) return 31; if (sidx == c.sI) return 32;
                            if (sidx == c.sTA) return 33; if (sidx == c.sTB) return 34;
                            return (uint64_t)sidx;
                        };
                        uint64_t h = 1469598103934665603ull;
                        auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
                        mix(q.kind); mix(q.opset); mix(q.op_lo); mix(q.op_hi);
                        mix(role(q.x)); mix(role(q.zdst)); mix(role(q.dst1)); mix(q.nsrc);
                        for (int i = 0; i < q.nsrc; i++) mix(role(q.sw[i].src));
                        if (q.kind == QB_PASS_RHS && q.zdst == 0 && q.x == c.sP) h = 1;   // stage 0 of a step
                        if ((int)g_class_log.size() <= c.traj_id) g_class_log.resize(c.traj_id + 1);
                        g_class_log[c.traj_id].push_back(h);
                    }
                    break;
                }
                if (status) status[c.traj_id] = c.done;
                if (stats) { int* st = stats + (size_t)c.traj_id * 4; st[0] = c.n_rhs; st[1] = c.n_accept; st[2] = c.n_reject; st[3] = c.n_pass; }
                if (head >= ntraj) { active--; break; }
                start(c, (int)head++);
            }
        }
    };
    std::vector<qb_c128> zbuf(N), o1buf(N);
    control();
    int64_t rounds = 0;
    while (active > 0) {
        if (++rounds > max_rounds) return -1;
        for (int slot = 0; slot < nslots; slot++) {
            const QbPass& p = pass[slot];
            if (p.kind == QB_PASS_NONE) continue;
            double* rd = red.data() + (size_t)slot * QB_MAXRED;
            const qb_c128* cf = coef.data() + (size_t)slot * g.maxcoef;
            if (p.kind == QB_PASS_EXPECT) {
                const std::vector<HostOp>& ops = p.opset == QB_OPSET_EOPS ? s.eops : s.nops;
                const bool fun = p.opset == QB_OPSET_EOPS && s.eop_functional;
                const qb_c128* x = vsrc(slot, p.x);
                for (int m = p.op_lo; m < p.op_hi; m++) {
                    double sre = 0, sim = 0;
                    for (int64_t r = 0; r < N; r++) {
                        qb_c128 q = rowdot(ops[m], r, x);
                        if (fun) { sre += q.re; sim += q.im; }
                        else if (p.opset == QB_OPSET_NOPS && g.mc_trace) {
                            if (r % (g.mc_trace + 1) == 0) { sre += q.re; sim += q.im; }   // tr(n_k rho)
                        }
                        else { sre += x[r].re * q.re + x[r].im * q.im; sim += x[r].re * q.im - x[r].im * q.re; }
                    }
                    rd[2 * (m - p.op_lo)] = sre; rd[2 * (m - p.op_lo) + 1] = sim;
                }
                continue;
            }
            if (p.kind == QB_PASS_LINMAP) {
                const QbLinMap& lm = linmap[slot];
                qb_c128* base = pool.data() + (size_t)slot * V * N;
                std::vector<std::vector<qb_c128>> outs(lm.nout, std::vector<qb_c128>(N));
                for (int j = 0; j < lm.nout; j++)
                    for (int64_t r = 0; r < N; r++) {
                        qb_c128 o = {0, 0};
                        for (int k = 0; k < p.nsrc; k++) {
                            const qb_c128 v = vsrc(slot, p.sw[k].src)[r];
                            o.re += lm.w[j][k] * v.re; o.im += lm.w[j][k] * v.im;
                        }
                        outs[j][r] = o;
                    }
                double n0 = 0, tr0 = 0;
                for (int64_t r = 0; r < N; r++) {
                    n0 += outs[0][r].re * outs[0][r].re + outs[0][r].im * outs[0][r].im;
                    if (g.mc_trace && r % (g.mc_trace + 1) == 0) tr0 += outs[0][r].re;
                }
                for (int j = 0; j < lm.nout; j++) memcpy(base + (size_t)lm.dst[j] * N, outs[j].data(), N * sizeof(qb_c128));
                rd[0] = n0; rd[1] = 0; rd[2] = 0; rd[3] = tr0; rd[4] = 0;
                continue;
            }
            // operator application into zbuf (x must not alias any destination)
            for (int64_t r = 0; r < N; r++) { zbuf[r].re = 0; zbuf[r].im = 0; }
            if (p.kind == QB_PASS_RHS || p.kind == QB_PASS_APPLY) {
                const qb_c128* x = vsrc(slot, p.x);
                if (p.x >= 0 && (p.x == p.zdst || p.x == p.dst1)) return -2;   // gather hazard
                for (int64_t r = 0; r < N; r++) {
                    qb_c128 z = {0, 0};
                    if (p.kind == QB_PASS_RHS) {
                        for (int e = 0; e < g.nelem; e++) {
                            qb_c128 q = rowdot(s.elems[e], r, x);
                            z.re += cf[e].re * q.re - cf[e].im * q.im; z.im += cf[e].re * q.im + cf[e].im * q.re;
                        }
                    } else {
                        qb_c128 q = rowdot(s.cops[p.op_lo], r, x);
                        z.re = cf[0].re * q.re - cf[0].im * q.im; z.im = cf[0].re * q.im + cf[0].im * q.re;
                    }
                    z.re *= p.zscale; z.im *= p.zscale;
                    zbuf[r] = z;
                }
            }
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
            qb_c128* base = pool.data() + (size_t)slot * V * N;
            for (int64_t r = 0; r < N; r++) {
                const qb_c128 z = zbuf[r];
                qb_c128 o1 = {0, 0}, o2 = {0, 0};
                for (int i = 0; i < p.nsrc; i++) {
                    const qb_c128 v = vsrc(slot, p.sw[i].src)[r];
                    o1.re += p.sw[i].w1 * v.re; o1.im += p.sw[i].w1 * v.im;
                    o2.re += p.w2[i] * v.re; o2.im += p.w2[i] * v.im;
                }
                o1.re += p.w1z * z.re; o1.im += p.w1z * z.im;
                o2.re += p.w2z * z.re; o2.im += p.w2z * z.im;
                o1buf[r] = o1;
                const double n1 = o1.re * o1.re + o1.im * o1.im;
                r0 += n1;
                if (p.red & QB_RED_WRMS) {
                    const double q = sqrt(o2.re * o2.re + o2.im * o2.im) / (g.opt.atol + g.opt.rtol * sqrt(n1));
                    r1 += q * q;
                }
                r2 += z.re * z.re + z.im * z.im;
                if (g.mc_trace && r % (g.mc_trace + 1) == 0) { r3 += o1.re; r4 += z.re; }
            }
            if (p.zdst >= 0) memcpy(base + (size_t)p.zdst * N, zbuf.data(), N * sizeof(qb_c128));
            if (p.dst1 >= 0) memcpy(base + (size_t)p.dst1 * N, o1buf.data(), N * sizeof(qb_c128));
            else if (p.dst1 == QB_SLOT_OUT)
                memcpy((qb_c128*)states + ((size_t)traj[slot].traj_id * nt + p.out_index) * N, o1buf.data(), N * sizeof(qb_c128));
            rd[0] = r0; rd[1] = r1; rd[2] = r2; rd[3] = r3; rd[4] = r4;
        }
        control();
    }
    return (int)std::min<int64_t>(rounds, 0x7fffffff);
}
}
