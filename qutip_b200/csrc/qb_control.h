// qb_control.h -- per-trajectory step controller ("scalar unit" of the engine).
//
// The engine advances every trajectory by alternating two kernels:
//   pass kernel  : executes ONE vector instruction (QbPass) per trajectory -- an operator
//                  application fused with the RK linear combinations and reductions;
//   control kernel: reduces the partial sums and runs qb_advance() below, which consumes
//                  the reductions, takes every scalar decision of the reference's
//                  integrator / Monte-Carlo logic and emits the next QbPass.
// No host round trip happens between steps.
//
// qb_advance() is a flattened, resumable restatement of
//   Explicit_RungeKutta       qutip/solver/integrator/explicit_rk.pyx:204-467
//   IntegratorVern7/9 wrapper qutip/solver/integrator/qutip_integrator.py:69-92
//   MCIntegrator              qutip/solver/mcsolve.py:245-406
//   Solver.run / _run_one_traj loops  solver_base.py:159-226, multitraj.py:260-283
// Every label cites the reference lines it mirrors.  The code is host+device so that the
// control flow can be unit-tested without a GPU (tests/emul); it contains no vector work.
#pragma once
#include <math.h>
#include "qb_types.h"
#include "qb_coeff.h"
#include "qb_adams.h"

struct QbProgRef { int off, len; };   // len == 0: constant coefficient 1

struct QbCtl {
    QbTableau tab;
    QbOptions opt;
    int N, ntiles;
    int nelem, ncops, neops, nargs;
    int eop_functional;         // e_ops are linear functionals (mesolve tr(E rho))
    int has_host_coef;          // some RHS element's coefficient is evaluated by the host
    int tile_mode;              // trajectory-interleaved tiles: slot labels stay fixed (copies,
                                // not relabels) so that lanes in the same phase issue identical
                                // slot patterns and can execute one pass together
    int exp_chunk;              // operators per EXPECT pass (<= QB_MAXRED / 2)
    int maxcoef;
    int mc_trace;               // n > 0: mcsolve of a super-operator H on the column-stacked n x n rho:
                                // the "norm" is tr(rho).real (mcsolve.py:311-319), collapse probabilities
                                // are tr(n_k rho), states are renormalised by their trace
    int nt, ndraws;
    const QbProgRef* elem_prog;  // [nelem]
    const QbProgRef* cop_prog;   // [ncops]
    const QbProgRef* nop_prog;   // [ncops]
    const QbProgRef* eop_prog;   // [neops]
    const QbInstr* instr;
    const QbSpline* splines;
    const double* spool;
    const qb_c128* args;         // [ntraj][nargs]
    const double* tlist;         // [nt]
    const double* draws;         // [ntraj][ndraws]
    qb_c128* out_expect;         // [ntraj][neops][nt]
    int* out_ncol;               // [ntraj]
    double* out_col_t;           // [ntraj][max_collapses]
    int* out_col_which;          // [ntraj][max_collapses]
};

// ------------------------------------------------------------------ small helpers
// the quantity MCIntegrator compares with its threshold: ||y||^2 (red[0]) for kets, tr(rho)
// (red[3]) for a super-operator H; qb_mcz: the same for z (red[2] / red[4])
QB_HD double qb_mcn(const QbCtl& g, const double* red) { return g.mc_trace ? red[3] : red[0]; }
QB_HD double qb_mc_inv_norm(const QbCtl& g, double n) { return g.mc_trace ? 1.0 / n : 1.0 / sqrt(n); }
QB_HD void qb_pass_clear(QbPass& p) {
    p.kind = QB_PASS_NONE; p.opset = 0; p.op_lo = 0; p.op_hi = 0; p.x = -1; p.zdst = -1;
    p.dst1 = -1; p.nsrc = 0; p.red = 0; p.out_index = 0; p.zscale = 1.0; p.w1z = 0.0;
    p.w2z = 0.0;
}
// reference iadd_data skips exact-zero factors (explicit_rk.pyx:38-39)
QB_HD void qb_pass_src(QbPass& p, int slot, double w1, double w2) {
    if (w1 == 0.0 && w2 == 0.0) return;
    p.sw[p.nsrc].src = slot; p.sw[p.nsrc].w1 = w1; p.w2[p.nsrc] = w2; p.nsrc++;
}
QB_HD int qb_eval_ref(const QbCtl& g, const QbTraj& c, QbProgRef pr, double t, qb_c128* out) {
    if (pr.len == 0) { out->re = 1.0; out->im = 0.0; return 0; }
    return qb_eval_prog(g.instr + pr.off, pr.len, t, g.args + (size_t)c.traj_id * g.nargs,
                        g.splines, g.spool, out);
}
// coefficients of all RHS elements at time t -> coef[0..nelem)
// returns 0 ok, -1 bad program, 1 = values of host-evaluated elements are needed for time t
QB_HD int qb_eval_rhs_coefs(const QbCtl& g, QbTraj& c, double t, qb_c128* coef) {
    if (g.has_host_coef && !(c.hc_valid && c.hc_t == t)) { c.hc_t = t; c.hc_valid = 0; return 1; }
    for (int e = 0; e < g.nelem; e++) {
        const QbProgRef pr = g.elem_prog[e];
        if (pr.len == 1 && g.instr[pr.off].op == QB_I_HOST) continue;   // written by the host
        if (qb_eval_ref(g, c, pr, t, &coef[e])) return -1;
    }
    c.hc_valid = 0;
    return 0;
}
QB_HD int qb_ad_maxord(const QbCtl& g) {
    return (g.opt.max_order >= 1 && g.opt.max_order < QB_AD_MAXORD) ? g.opt.max_order : QB_AD_MAXORD;
}
// common handling of the three outcomes at a pass-issuing label `label`
#define QB_COEFS_OR_PAUSE(tval, label)                                              \
    {                                                                               \
        const int rcq_ = qb_eval_rhs_coefs(g, c, (tval), coef);                     \
        if (rcq_ < 0) { c.status = QB_ST_BAD_PROGRAM; L = QL_FAIL; break; }         \
        if (rcq_ > 0) { qb_pass_clear(p); c.pc = QB_PC_IDLE; c.hc_resume = (label); \
                        c.done = 2; return 0; }                                     \
    }
// slot index of k[j]: for FSAL tableaux k[0] and k[s-1] exchange their buffers every step
QB_HD int qb_ks(const QbTraj& c, int j, int s) {
    if (c.kswap) { if (j == 0) return s - 1; if (j == s - 1) return 0; }
    return j;
}
QB_HD int qb_stage_x(const QbTraj& c, int i, int first) {
    // stage `first` reads sTA (or y_prev for stage 0); then TB, TA, ... alternate
    return ((i - first) & 1) ? c.sTB : c.sTA;
}

// internal labels (not resumable)
enum {
    QL_RETURN = 1000, QL_FAIL, QL_SET_BEGIN, QL_SET_DONE, QL_INT_BEGIN, QL_RK_LOOP,
    QL_STEP_ATTEMPT, QL_STAGE_ISSUE, QL_AFTER_LOOP, QL_DENSE_BEGIN, QL_DENSE_ISSUE,
    QL_INTERP_ISSUE, QL_INT_DONE, QL_ME_REACHED, QL_ME_NEXT, QL_MC_ENTRY, QL_MC_LOOP,
    QL_MC_TARGET, QL_RECORD, QL_EXPECT_ISSUE, QL_AFTER_RECORD, QL_RF_LOOP, QL_RF_END,
    QL_COLLAPSE, QL_APPLY_ISSUE, QL_FINISH,
    // Adams (qb_adams.h)
    QL_AD_SET, QL_AD_F0, QL_AD_INT, QL_AD_LOOP, QL_AD_STEP, QL_AD_PRED_ISSUE, QL_AD_CORR_ISSUE,
    QL_AD_CONVFAIL, QL_AD_ERRTEST, QL_AD_UPD_ISSUE, QL_AD_O520, QL_AD_O540, QL_AD_O560,
    QL_AD_RESCALE, QL_AD_3FAIL, QL_AD_3FAIL_ISSUE, QL_AD_STEP_DONE, QL_AD_AFTER, QL_AD_INTERP
};

// Run the controller until it has emitted a pass (returns 1), the trajectory finished or
// paused (returns 0) or failed (returns 0 with c.done < 0).
//   T     : the Butcher tableau (g.tab on the host; a __constant__ copy on the device)
//   red   : reductions of the pass just executed (red[0]=|o1|^2, red[1]=wrms sum,
//           red[2]=|z|^2 ; EXPECT: red[2m], red[2m+1] = <x|O_m|x>)
//   coef  : per-slot coefficient buffer the next pass will read
//   probs : per-slot scratch [ncops]
//   lm    : per-slot weight table of LINMAP passes
QB_HD int qb_advance(const QbCtl& g, const QbTableau& T, QbTraj& c, QbPass& p,
                     const double* red, qb_c128* coef, double* probs, QbLinMap* lm)
{
    const int s = T.s, S = T.S;
    int L = c.pc;
    int dense_i = c.stage_arg, stage_i = c.stage_arg;   // used when resuming at an *_ISSUE label
    double set_t = 0.0;
    for (int guard = 0; guard < 4096; guard++) {
        switch (L) {
        case QB_PC_IDLE:
            qb_pass_clear(p);
            return 0;

        // ================================================================ entry points
        case QB_PC_ME_BEGIN:        // Solver.run: set_state, record t0 (solver_base.py:200-205)
            c.mode = 0; c.done = 0; c.ncol = 0; c.rng = 0;
            c.n_rhs = c.n_accept = c.n_reject = c.n_pass = 0;
            c.set_t = g.tlist[c.tl_idx]; c.set_x = QB_SLOT_INIT; c.set_scale = 1.0;
            c.after_set = QB_K_ME_START;
            L = QL_SET_BEGIN; break;
        case QB_PC_MC_BEGIN: {      // MCIntegrator.set_state (mcsolve.py:245-281)
            c.mode = 1; c.done = 0; c.ncol = 0; c.rng = 0;
            c.n_rhs = c.n_accept = c.n_reject = c.n_pass = 0;
            if (g.opt.no_jump) c.target_norm = 0.0;
            else {
                if (c.rng >= g.ndraws) { c.status = QB_ST_RNG_EXHAUSTED; L = QL_FAIL; break; }
                double u = g.draws[(size_t)c.traj_id * g.ndraws + c.rng++];
                c.target_norm = u * (1.0 - g.opt.jump_prob_floor) + g.opt.jump_prob_floor;
            }
            c.set_t = g.tlist[c.tl_idx]; c.set_x = QB_SLOT_INIT; c.set_scale = 1.0;
            c.after_set = QB_K_MC_START;
            L = QL_SET_BEGIN; break;
        }
        case QB_PC_ME_NEXT: L = QL_ME_NEXT; break;       // host appended targets
        case QB_PC_MC_ENTRY: L = QL_MC_ENTRY; break;
        case QB_PC_STEP_ENTRY:      // Integrator.mcstep(t) from the host (qutip_integrator.py:84-87)
            c.done = 0;
            c.int_t = g.tlist[c.tl_idx]; c.int_step = 1; c.after_int = QB_K_PAUSE;
            L = QL_INT_BEGIN; break;

        // ================================================================ set_initial_value
        case QL_SET_BEGIN: {        // explicit_rk.pyx:204-230 ; y0 = set_scale * V[set_x]
            if (T.method == 1) { L = QL_AD_SET; break; }
            set_t = c.set_t;
            c.t = c.t_prev = c.t_front = set_t;
            c.dt_int = 0.0;
            if (c.set_x == c.sF) {
                if (g.tile_mode) {                   // fixed labels: move the source out of the way
                    qb_pass_clear(p);
                    p.kind = QB_PASS_COMBINE; p.dst1 = c.sTA;
                    qb_pass_src(p, c.sF, 1.0, 0.0);
                    c.set_x = c.sTA;
                    c.pc = QB_PC_SETCOPY_DONE; return 1;
                }
                int tmp = c.sP; c.sP = c.sF; c.sF = tmp;
            }
            c.sY = c.sF;
            c.fsal_pending = 0;                      // k[0] will hold k_fsal = f(t, y0) (:222-225)
            qb_pass_clear(p);
            if (g.opt.first_step != 0.0 && !T.fsal) {   // :227-230, no estimate
                c.dt_safe = g.opt.first_step;
                p.kind = QB_PASS_COMBINE; p.dst1 = c.sF; p.red = QB_RED_NORM2_O1;
                qb_pass_src(p, c.set_x, c.set_scale, 0.0);
                c.pc = QB_PC_SET_DONE; return 1;
            }
            // _estimate_first_step (:232-276), first RHS evaluation k0 = f(t, y0)
            p.kind = QB_PASS_RHS; p.x = c.set_x; p.zscale = c.set_scale; p.zdst = qb_ks(c, 0, s);
            p.dst1 = c.sF; p.red = QB_RED_NORM2_O1 | QB_RED_NORM2_Z;
            qb_pass_src(p, c.set_x, c.set_scale, 0.0);
            QB_COEFS_OR_PAUSE(set_t, QL_SET_BEGIN)
            c.n_rhs++;
            c.pc = QB_PC_EST0_DONE; return 1;
        }
        case QB_PC_SETCOPY_DONE: L = QL_SET_BEGIN; break;
        case QB_PC_SET_DONE:
            c.norm2_y = qb_mcn(g, red);
            L = QL_SET_DONE; break;
        case QB_PC_EST0_DONE: {     // :237-256
            double norm = sqrt(red[0]);
            c.norm2_y = qb_mcn(g, red);
            if (g.opt.first_step != 0.0) {           // FSAL with a given first step: k_fsal only
                c.dt_safe = g.opt.first_step;
                L = QL_SET_DONE; break;
            }
            double tol = g.opt.atol + norm * g.opt.rtol;
            if (norm <= g.opt.atol) norm = 1.0;
            double tmp_norm = sqrt(red[2]);
            double fact = 1.0;
            for (int i = 1; i <= T.order; i++) fact *= i;
            double dt1;
            if (tmp_norm >= g.opt.atol * 1e-6)
                dt1 = pow(tol * fact * pow(norm, (double)T.order), 1.0 / (T.order + 1)) / tmp_norm;
            else
                dt1 = pow(tol * fact, 1.0 / (T.order + 1)) * norm * 0.5;
            c.est_norm = norm; c.est_tol = tol; c.est_dt1 = dt1;
            // y_temp = y0 + (dt1/100) k0   (:258-261)
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.dst1 = c.sTB;
            qb_pass_src(p, c.sF, 1.0, 0.0);
            qb_pass_src(p, qb_ks(c, 0, s), dt1 / 100, 0.0);
            c.pc = QB_PC_EST1IN_DONE; return 1;
        }
        case QB_PC_EST1IN_DONE: {   // k1 = f(t + dt1/100, y_temp)   (:262-263)
            qb_pass_clear(p);
            p.kind = QB_PASS_RHS; p.x = c.sTB; p.zdst = qb_ks(c, 1, s); p.red = QB_RED_NORM2_Z;
            QB_COEFS_OR_PAUSE(c.t + c.est_dt1 / 100, QB_PC_EST1IN_DONE)
            c.n_rhs++;
            c.pc = QB_PC_EST1_DONE; return 1;
        }
        case QB_PC_EST1_DONE: {     // :264-276
            double tmp_norm = sqrt(red[2]);
            double fact = 1.0;
            for (int i = 1; i <= T.order; i++) fact *= i;
            double dt2;
            if (tmp_norm >= g.opt.atol * 1e-6)
                dt2 = pow(c.est_tol * fact * pow(c.est_norm, (double)T.order),
                          1.0 / (T.order + 1)) / tmp_norm;
            else
                dt2 = c.est_dt1;
            double dt = c.est_dt1 < dt2 ? c.est_dt1 : dt2;
            double min_step = g.opt.min_step != 0.0 ? g.opt.min_step : 1e-15;   // :115
            if (g.opt.max_step != 0.0 && g.opt.max_step < dt) dt = g.opt.max_step;
            if (min_step > dt) dt = min_step;
            c.dt_safe = dt;
            L = QL_SET_DONE; break;
        }
        case QL_SET_DONE:
            switch (c.after_set) {
            case QB_K_ME_START: case QB_K_MC_START: L = QL_RECORD; break;   // record tlist[0]
            case QB_K_MC_AFTER_COLLAPSE:                  // mcsolve.py:297-298
                c.mc_t_old = c.t; c.mc_n_old = 1.0; L = QL_MC_LOOP; break;
            default: L = QL_FINISH; break;               // host-driven set_state
            }
            break;

        // ================================================================ integrate(t, step)
        case QL_INT_BEGIN: {        // explicit_rk.pyx:278-310
            if (T.method == 1) { L = QL_AD_INT; break; }
            double t = c.int_t;
            if (t == c.t) { L = QL_INT_DONE; break; }                       // :291
            if (t < c.t_prev) { c.status = QB_ST_OUTSIDE_RANGE; L = QL_FAIL; break; }   // :294
            if (g.opt.interpolate && t < c.t_front) {                       // :298-304
                if (c.status != QB_ST_INTERPOLATED) { L = QL_DENSE_BEGIN; break; }
                L = QL_INTERP_ISSUE; break;
            }
            c.status = QB_ST_NORMAL;
            if (c.int_step && c.t < c.t_front && t > c.t_front) c.int_t = c.t_front;  // :308-310
            c.nsteps_left = g.opt.nsteps;
            L = QL_RK_LOOP; break;
        }
        case QL_RK_LOOP:            // while self._t_front < t and self._status >= 0   (:312)
            if (c.t_front < c.int_t && c.status >= 0) {
                if (g.tile_mode) {                   // y_prev <- y_front (:313) as a real copy
                    qb_pass_clear(p);
                    p.kind = QB_PASS_COMBINE; p.dst1 = c.sP;
                    qb_pass_src(p, c.sF, 1.0, 0.0);
                    c.pc = QB_PC_COPY_DONE; return 1;
                }
                int tmp = c.sP; c.sP = c.sF; c.sF = tmp;     // y_prev <- y_front (:313), by relabel
                if (T.fsal && c.fsal_pending) { c.kswap ^= 1; c.fsal_pending = 0; }   // k[0] <- k_fsal
                c.t_prev = c.t_front;
                c.step_n = 0;
                L = QL_STEP_ATTEMPT; break;
            }
            L = QL_AFTER_LOOP; break;
        case QB_PC_COPY_DONE:
            if (T.fsal && c.fsal_pending) { c.kswap ^= 1; c.fsal_pending = 0; }   // k[0] <- k_fsal
            c.t_prev = c.t_front;
            c.step_n = 0;
            L = QL_STEP_ATTEMPT; break;
        case QL_STEP_ATTEMPT: {     // _step_in_err body (:339-343) + _get_timestep (:440-449)
            double dt_needed = c.int_t - c.t_prev, dt;
            if (g.opt.interpolate) dt = c.dt_safe;
            else if (dt_needed <= c.dt_safe) dt = dt_needed;
            else dt = dt_needed / ((int)(dt_needed / c.dt_safe) + 1);
            c.dt_cur = dt;
            stage_i = 0;
            L = QL_STAGE_ISSUE; break;
        }
        case QL_STAGE_ISSUE: {      // _compute_step (:358-389), one stage per pass
            const int i = stage_i;
            const double dt = c.dt_cur;
            qb_pass_clear(p);
            if (i == 0 && T.fsal) {
                // k[0] = k_fsal is already there (:365-366): only build the stage-1 input
                p.kind = QB_PASS_COMBINE; p.dst1 = qb_stage_x(c, 1, 1);
                qb_pass_src(p, c.sP, 1.0, 0.0);
                qb_pass_src(p, qb_ks(c, 0, s), dt * T.a[1][0], 0.0);
                c.stage = 0;
                c.pc = QB_PC_STAGE_DONE; return 1;
            }
            p.kind = QB_PASS_RHS;
            p.x = (i == 0) ? c.sP : qb_stage_x(c, i, 1);
            p.zdst = qb_ks(c, i, s);
            qb_pass_src(p, c.sP, 1.0, 0.0);
            if (i < s - 1) {        // epilogue builds the next stage input (:374-375)
                const int n = i + 1;
                p.dst1 = qb_stage_x(c, n, 1);
                for (int j = 0; j < i; j++) qb_pass_src(p, qb_ks(c, j, s), dt * T.a[n][j], 0.0);
                p.w1z = dt * T.a[n][i];
            } else {                // y_front, error vector (:380-397)
                p.dst1 = c.sF;
                for (int j = 0; j < i; j++) qb_pass_src(p, qb_ks(c, j, s), dt * T.b[j], dt * T.e[j]);
                p.w1z = dt * T.b[i]; p.w2z = dt * T.e[i];
                p.red = QB_RED_NORM2_O1 | QB_RED_WRMS;
            }
            c.stage_arg = i;
            QB_COEFS_OR_PAUSE(i == 0 ? c.t_prev : c.t_prev + T.c[i] * dt, QL_STAGE_ISSUE)
            c.stage = i; c.n_rhs++;
            c.pc = QB_PC_STAGE_DONE; return 1;
        }
        case QB_PC_STAGE_DONE: {
            if (c.stage < s - 1) { stage_i = c.stage + 1; L = QL_STAGE_ISSUE; break; }
            // step finished: error, controller (:341-352, :451-467)
            const double dt = c.dt_cur;
            const double err = sqrt(red[1] / (double)g.N);
            c.t_front = c.t_prev + dt;
            c.dt_int = dt;
            c.norm2_front = qb_mcn(g, red);
            double factor;
            if (err == 0.0) factor = 10.0;
            else {
                factor = 0.9 * pow(err, -1.0 / (T.order + 1));
                if (factor > 10.0) factor = 10.0;
                if (factor < 0.2) factor = 0.2;
            }
            double min_step = g.opt.min_step != 0.0 ? g.opt.min_step : 1e-15;
            double dts = dt * factor;
            if (g.opt.max_step != 0.0 && g.opt.max_step < dts) dts = g.opt.max_step;
            if (min_step > dts) dts = min_step;
            c.dt_safe = dts;
            if (err >= 1.0) c.n_reject++; else c.n_accept++;
            bool stop = false;
            if (dt == min_step && err > 1.0) { c.status = QB_ST_DT_UNDERFLOW; stop = true; }
            else {
                c.step_n++;
                if (c.step_n > c.nsteps_left) { c.status = QB_ST_TOO_MUCH_WORK; stop = true; }
            }
            if (!stop && err >= 1.0) { L = QL_STEP_ATTEMPT; break; }   // while error >= 1
            if (T.fsal) c.fsal_pending = 1;                             // k_fsal <- k[s-1] (:354-355)
            c.nsteps_left -= c.step_n;                                  // :315
            L = c.int_step ? QL_AFTER_LOOP : QL_RK_LOOP;                // :317-318
            break;
        }
        case QL_AFTER_LOOP:         // :320-331
            if (c.status < 0) { L = QL_FAIL; break; }
            if (c.t_front > c.int_t) { L = QL_DENSE_BEGIN; break; }
            c.status = QB_ST_AT_FRONT;
            c.t = c.t_front; c.sY = c.sF; c.norm2_y = c.norm2_front;
            L = QL_INT_DONE; break;

        // ---------------------------------------------------------------- dense output
        case QL_DENSE_BEGIN: {      // _prep_dense_out (:399-410), input of the first extra stage
            if (S == s) { c.status = QB_ST_INTERPOLATED; L = QL_INTERP_ISSUE; break; }
            const double dt = c.dt_int;
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.dst1 = c.sTA;
            qb_pass_src(p, c.sP, 1.0, 0.0);
            for (int j = 0; j < s; j++) qb_pass_src(p, qb_ks(c, j, s), dt * T.a[s][j], 0.0);
            c.pc = QB_PC_DENSEIN_DONE; return 1;
        }
        case QB_PC_DENSEIN_DONE: dense_i = s; L = QL_DENSE_ISSUE; break;
        case QL_DENSE_ISSUE: {
            const int i = dense_i;
            const double dt = c.dt_int;
            qb_pass_clear(p);
            p.kind = QB_PASS_RHS; p.x = qb_stage_x(c, i, s); p.zdst = qb_ks(c, i, s);
            qb_pass_src(p, c.sP, 1.0, 0.0);
            if (i < S - 1) {
                const int n = i + 1;
                p.dst1 = qb_stage_x(c, n, s);
                for (int j = 0; j < i; j++) qb_pass_src(p, qb_ks(c, j, s), dt * T.a[n][j], 0.0);
                p.w1z = dt * T.a[n][i];
            } else {                // last extra stage: fuse _interpolate_step(int_t) (:412-430)
                const double tau = (c.int_t - c.t_prev) / dt;
                p.dst1 = c.sI; p.red = QB_RED_NORM2_O1;
                for (int j = 0; j < S; j++) {
                    double bf = 0.0;
                    for (int q = T.dense_order - 1; q >= 0; q--) { bf += T.bi[j][q]; bf *= tau; }
                    if (j < i) qb_pass_src(p, qb_ks(c, j, s), dt * bf, 0.0); else p.w1z = dt * bf;
                }
            }
            c.stage_arg = i;
            QB_COEFS_OR_PAUSE(c.t_prev + T.c[i] * dt, QL_DENSE_ISSUE)
            c.stage = i; c.n_rhs++;
            c.pc = QB_PC_DENSE_DONE; return 1;
        }
        case QB_PC_DENSE_DONE:
            if (c.stage < S - 1) { dense_i = c.stage + 1; L = QL_DENSE_ISSUE; break; }
            c.status = QB_ST_INTERPOLATED;
            c.t = c.int_t; c.sY = c.sI; c.norm2_y = qb_mcn(g, red);
            L = QL_INT_DONE; break;
        case QL_INTERP_ISSUE: {     // _interpolate_step (:412-430) with k already complete
            const double dt = c.dt_int;
            const double tau = (c.int_t - c.t_prev) / dt;
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.dst1 = c.sI; p.red = QB_RED_NORM2_O1;
            qb_pass_src(p, c.sP, 1.0, 0.0);
            for (int j = 0; j < S; j++) {
                double bf = 0.0;
                for (int q = T.dense_order - 1; q >= 0; q--) { bf += T.bi[j][q]; bf *= tau; }
                qb_pass_src(p, qb_ks(c, j, s), dt * bf, 0.0);
            }
            c.pc = QB_PC_INTERP_DONE; return 1;
        }
        case QB_PC_INTERP_DONE:
            c.status = QB_ST_INTERPOLATED;
            c.t = c.int_t; c.sY = c.sI; c.norm2_y = qb_mcn(g, red);
            L = QL_INT_DONE; break;

        case QL_INT_DONE:
            switch (c.after_int) {
            case QB_K_ME_REACHED: L = QL_RECORD; break;
            case QB_K_MC_AFTER_STEP: {      // mcsolve.py:290-300
                const double norm = c.norm2_y;
                if (norm <= c.target_norm) {
                    c.rf_n_old = c.mc_n_old; c.rf_n = norm;
                    c.rf_t_prev = c.mc_t_old; c.rf_t_final = c.t;
                    c.rf_tries = 0;
                    L = QL_RF_LOOP;
                } else {
                    c.mc_t_old = c.t; c.mc_n_old = norm;
                    L = QL_MC_LOOP;
                }
                break;
            }
            case QB_K_RF_AFTER_GUESS: {     // mcsolve.py:348-361
                const double n2 = c.norm2_y;
                if (fabs(c.target_norm - n2) < g.opt.norm_tol * c.target_norm) { L = QL_RF_END; break; }
                if (n2 < c.target_norm) { c.rf_t_final = c.rf_t_guess; c.rf_n = n2; }
                else { c.rf_t_prev = c.rf_t_guess; c.rf_n_old = n2; }
                L = QL_RF_LOOP; break;
            }
            default: L = QL_FINISH; break;
            }
            break;

        // ================================================================ Adams-Moulton (qb_adams.h)
        case QL_AD_SET: {           // set_state: YH0 = set_scale * V[set_x] (in-place safe: no gather)
            c.t = c.t_prev = c.t_front = c.ad_tn = c.set_t;
            c.dt_int = 0.0;
            c.sY = QB_AD_YH(0);
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.dst1 = QB_AD_YH(0); p.red = QB_RED_NORM2_O1;
            qb_pass_src(p, c.set_x, c.set_scale, 0.0);
            c.pc = QB_PC_AD_SET0_DONE; return 1;
        }
        case QB_PC_AD_SET0_DONE:
            c.norm2_y = c.norm2_front = qb_mcn(g, red);
            L = QL_AD_F0; break;
        case QL_AD_F0: {            // YH1 = f(t0, y0), unscaled (ad_hyh = 1); its weighted norm
            qb_pass_clear(p);
            p.kind = QB_PASS_RHS; p.x = QB_AD_YH(0); p.zdst = QB_AD_YH(1); p.red = QB_RED_WRMS;
            qb_pass_src(p, QB_AD_YH(0), 1.0, 0.0);
            p.w2z = 1.0;
            QB_COEFS_OR_PAUSE(c.ad_tn, QL_AD_F0)
            c.n_rhs++;
            c.pc = QB_PC_AD_F0_DONE; return 1;
        }
        case QB_PC_AD_F0_DONE:
            c.ad_f0n2 = red[1] / (double)g.N;
            c.ad_nq = 1; c.ad_ialth = 2; c.ad_rmax = 1e4; c.ad_crate = 0.7;
            c.ad_kflag = 0; c.ad_ncf = 0;
            c.ad_hyh = 1.0; c.ad_hu = 0.0; c.ad_pred_nq = -1;
            c.ad_h = g.opt.first_step;               // 0: chosen when the first target is known
            L = QL_SET_DONE; break;

        case QL_AD_INT: {           // integrate(t, step) with the semantics of explicit_rk.pyx:278-331
            const double t = c.int_t;
            if (t == c.t) { L = QL_INT_DONE; break; }
            if (t < c.t_prev) { c.status = QB_ST_OUTSIDE_RANGE; L = QL_FAIL; break; }
            if (t < c.t_front) { L = QL_AD_INTERP; break; }
            c.status = QB_ST_NORMAL;
            if (c.int_step && c.t < c.t_front && t > c.t_front) c.int_t = c.t_front;
            c.step_n = 0;
            L = QL_AD_LOOP; break;
        }
        case QL_AD_LOOP:
            if (c.t_front < c.int_t && c.status >= 0) {
                if (c.ad_h == 0.0) {
                    // initial step: h0^2 = 1 / (1/(tol w0^2) + tol ||f0||^2), not past the target
                    double tol = g.opt.rtol;
                    if (tol < 100.0 * 2.220446049250313e-16) tol = 100.0 * 2.220446049250313e-16;
                    if (tol > 1e-3) tol = 1e-3;
                    double w0 = fabs(c.ad_tn) > fabs(c.int_t) ? fabs(c.ad_tn) : fabs(c.int_t);
                    double h0 = 1.0 / sqrt(1.0 / (tol * w0 * w0) + tol * c.ad_f0n2);
                    if (h0 > c.int_t - c.ad_tn) h0 = c.int_t - c.ad_tn;
                    if (g.opt.max_step != 0.0 && h0 > g.opt.max_step) h0 = g.opt.max_step;
                    if (h0 < g.opt.min_step) h0 = g.opt.min_step;
                    c.ad_h = h0;
                }
                L = QL_AD_STEP; break;
            }
            L = QL_AD_AFTER; break;
        case QL_AD_STEP:            // one attempt with step ad_h at order ad_nq
            if (c.ad_tn + c.ad_h == c.ad_tn) { c.status = QB_ST_DT_UNDERFLOW; L = QL_FAIL; break; }
            if (++c.step_n > g.opt.nsteps) { c.status = QB_ST_TOO_MUCH_WORK; L = QL_FAIL; break; }
            L = QL_AD_PRED_ISSUE; break;
        case QL_AD_PRED_ISSUE: {    // YP_j = sum_{k>=j} C(k,j) eta^k YH_k  (Pascal-triangle prediction,
            const int nq = c.ad_nq;  // every history vector read once: one LINMAP pass)
            if (c.ad_pred_nq == nq && c.ad_pred_h == c.ad_h && c.ad_hyh == c.ad_h) {
                c.ad_pred_nq = -1;   // already written by the previous step's update pass
                L = QB_PC_AD_PRED_DONE; break;
            }
            c.ad_pred_nq = -1;
            const double eta = c.ad_h / c.ad_hyh;
            qb_pass_clear(p);
            p.kind = QB_PASS_LINMAP; p.nsrc = nq + 1;
            for (int k = 0; k <= nq; k++) p.sw[k].src = QB_AD_YH(k);
            lm->nout = nq + 1;
            for (int jj = 0; jj <= nq; jj++) {
                lm->dst[jj] = QB_AD_YP(jj);
                for (int k = 0; k < QB_LM_MAXSRC; k++) lm->w[jj][k] = 0.0;
                double w = 1.0;
                for (int k = 0; k < jj; k++) w *= eta;
                for (int k = jj; k <= nq; k++) {
                    lm->w[jj][k] = w;
                    w *= eta * (double)(k + 1) / (double)(k + 1 - jj);
                }
            }
            c.pc = QB_PC_AD_PRED_DONE; return 1;
        }
        case QB_PC_AD_PRED_DONE:
            c.ad_m = 0; c.ad_fsel = 0; c.ad_ysel = 0; c.ad_delp = 0.0;
            L = QL_AD_CORR_ISSUE; break;
        case QL_AD_CORR_ISSUE: {
            // functional iteration m:  savf = h f(t_n + h, y_m);  acor_m = savf - YP1;
            // y_{m+1} = YP0 + l_0 acor_m;  del = || acor_m - acor_{m-1} ||  -- one fused pass
            const int m = c.ad_m;
            const double el0 = T.a[c.ad_nq][0];
            const int ycur = c.ad_ysel ? c.sTB : c.sTA, ynew = c.ad_ysel ? c.sTA : c.sTB;
            const int fcur = QB_AD_SAVF(c.ad_fsel), fnew = QB_AD_SAVF(c.ad_fsel ^ 1);
            qb_pass_clear(p);
            p.kind = QB_PASS_RHS; p.x = (m == 0) ? QB_AD_YP(0) : ycur;
            p.zscale = c.ad_h; p.zdst = fnew; p.dst1 = ynew; p.red = QB_RED_WRMS;
            qb_pass_src(p, QB_AD_YP(0), 1.0, 0.0);
            qb_pass_src(p, QB_AD_YP(1), -el0, m == 0 ? -1.0 : 0.0);
            if (m > 0) qb_pass_src(p, fcur, 0.0, -1.0);
            p.w1z = el0; p.w2z = 1.0;
            QB_COEFS_OR_PAUSE(c.ad_tn + c.ad_h, QL_AD_CORR_ISSUE)
            c.n_rhs++;
            c.pc = QB_PC_AD_CORR_DONE; return 1;
        }
        case QB_PC_AD_CORR_DONE: {
            const int nq = c.ad_nq;
            const double tq2 = T.bi[nq][1], conit = 0.5 / (nq + 2);
            const double del = sqrt(red[1] / (double)g.N);
            int m = c.ad_m;
            c.ad_ysel ^= 1; c.ad_fsel ^= 1;          // the buffers just written are current
            if (m != 0 && c.ad_delp > 0.0) {
                const double r = del / c.ad_delp;
                c.ad_crate = 0.2 * c.ad_crate > r ? 0.2 * c.ad_crate : r;
            }
            const double cr = 1.5 * c.ad_crate < 1.0 ? 1.5 * c.ad_crate : 1.0;
            const double dcon = del * cr / (tq2 * conit);
            c.ad_del = del;
            if (dcon <= 1.0) {
                if (m == 0) { c.ad_dsm = del / tq2; L = QL_AD_ERRTEST; break; }
                qb_pass_clear(p);                    // dsm = || savf - YP1 || / tq2
                p.kind = QB_PASS_COMBINE; p.red = QB_RED_WRMS;
                qb_pass_src(p, QB_AD_YP(0), 1.0, 0.0);
                qb_pass_src(p, QB_AD_SAVF(c.ad_fsel), 0.0, 1.0);
                qb_pass_src(p, QB_AD_YP(1), 0.0, -1.0);
                c.pc = QB_PC_AD_DSM_DONE; return 1;
            }
            m++;
            if (m == 3 || (m >= 2 && del > 2.0 * c.ad_delp)) { L = QL_AD_CONVFAIL; break; }
            c.ad_delp = del; c.ad_m = m;
            L = QL_AD_CORR_ISSUE; break;
        }
        case QB_PC_AD_DSM_DONE:
            c.ad_dsm = sqrt(red[1] / (double)g.N) / T.bi[c.ad_nq][1];
            L = QL_AD_ERRTEST; break;
        case QL_AD_CONVFAIL:        // corrector did not converge: quarter the step and retry
            c.ad_ncf++; c.ad_rmax = 2.0;
            if (fabs(c.ad_h) <= g.opt.min_step * 1.00001 || c.ad_ncf >= 10) {
                c.status = QB_ST_CORRECTOR_FAILED; L = QL_FAIL; break;
            }
            c.ad_rh = 0.25; c.ad_iredo = 1;
            L = QL_AD_RESCALE; break;
        case QL_AD_ERRTEST:
            if (c.ad_dsm > 1.0) {   // local error test failed
                c.ad_kflag--; c.ad_rmax = 2.0; c.n_reject++;
                if (fabs(c.ad_h) <= g.opt.min_step * 1.00001) { c.status = QB_ST_DT_UNDERFLOW; L = QL_FAIL; break; }
                if (c.ad_kflag <= -3) { L = QL_AD_3FAIL; break; }
                c.ad_iredo = 2; c.ad_rhup = 0.0;
                L = QL_AD_O540; break;
            }
            c.n_accept++;
            L = QL_AD_UPD_ISSUE; break;
        case QL_AD_UPD_ISSUE: {     // YH_j = YP_j + l_j (savf - YP1), all columns in one LINMAP pass;
            const int nq = c.ad_nq;  // when due, acor = savf - YP1 is kept for the order-increase estimate
            qb_pass_clear(p);
            p.kind = QB_PASS_LINMAP; p.nsrc = nq + 2; p.red = QB_RED_NORM2_O1;
            for (int k = 0; k <= nq; k++) p.sw[k].src = QB_AD_YP(k);
            p.sw[nq + 1].src = QB_AD_SAVF(c.ad_fsel);
            int nout = nq + 1;
            for (int jj = 0; jj <= nq; jj++) {
                const double elj = T.a[nq][jj];
                lm->dst[jj] = QB_AD_YH(jj);
                for (int k = 0; k < QB_LM_MAXSRC; k++) lm->w[jj][k] = 0.0;
                lm->w[jj][jj] = 1.0;
                lm->w[jj][1] -= elj;
                lm->w[jj][nq + 1] = elj;
            }
            if (c.ad_ialth == 2 && nq < qb_ad_maxord(g)) {       // ialth will be 1 after this step
                lm->dst[nout] = c.sP;
                for (int k = 0; k < QB_LM_MAXSRC; k++) lm->w[nout][k] = 0.0;
                lm->w[nout][1] = -1.0; lm->w[nout][nq + 1] = 1.0;
                nout++;
            }
            // Unless an order / step-size decision follows this step (it reads savf and YP1),
            // the same pass also writes the NEXT step's prediction YP'_j = sum_{k>=j} C(k,j) YH_k
            // (same h): one round per step less, and the new YH is not re-read.
            c.ad_pred_nq = -1;
            if (c.ad_ialth != 1 && nout + nq + 1 <= QB_LM_MAXOUT) {
                for (int jj = 0; jj <= nq; jj++) {
                    double* wr = lm->w[nout + jj];
                    for (int k = 0; k < QB_LM_MAXSRC; k++) wr[k] = 0.0;
                    double cb = 1.0;                         // C(k, jj)
                    for (int k = jj; k <= nq; k++) {
                        for (int m2 = 0; m2 <= nq + 1; m2++) wr[m2] += cb * lm->w[k][m2];
                        cb *= (double)(k + 1) / (double)(k + 1 - jj);
                    }
                    lm->dst[nout + jj] = QB_AD_YP(jj);
                }
                nout += nq + 1;
                c.ad_pred_h = c.ad_h; c.ad_pred_nq = nq;
            }
            lm->nout = nout;
            c.pc = QB_PC_AD_UPD_DONE; return 1;
        }
        case QB_PC_AD_UPD_DONE:
            c.norm2_front = qb_mcn(g, red);
            // the step is accepted
            c.ad_kflag = 0; c.ad_iredo = 0; c.ad_ncf = 0;
            c.ad_hu = c.ad_h; c.ad_tn += c.ad_h; c.ad_hyh = c.ad_h;
            c.t_front = c.ad_tn; c.t_prev = c.ad_tn - c.ad_hu; c.dt_int = c.ad_hu;
            c.ad_ialth--;
            if (c.ad_ialth == 0) { L = QL_AD_O520; break; }
            L = QL_AD_STEP_DONE; break;
        case QB_PC_AD_SAVE_DONE: L = QL_AD_STEP_DONE; break;

        // ---- order / step-size selection: candidates rh at order nq-1, nq, nq+1 ----
        case QL_AD_O520:
            c.ad_rhup = 0.0;
            if (c.ad_nq >= qb_ad_maxord(g)) { L = QL_AD_O540; break; }
            qb_pass_clear(p);       // dup = || acor - acor_saved || / tq3
            p.kind = QB_PASS_COMBINE; p.red = QB_RED_WRMS;
            qb_pass_src(p, QB_AD_YH(0), 1.0, 0.0);
            qb_pass_src(p, QB_AD_SAVF(c.ad_fsel), 0.0, 1.0);
            qb_pass_src(p, QB_AD_YP(1), 0.0, -1.0);
            qb_pass_src(p, c.sP, 0.0, -1.0);
            c.pc = QB_PC_AD_DUP_DONE; return 1;
        case QB_PC_AD_DUP_DONE: {
            const int l = c.ad_nq + 1;
            const double dup = sqrt(red[1] / (double)g.N) / T.bi[c.ad_nq][2];
            c.ad_rhup = 1.0 / (1.4 * pow(dup, 1.0 / (l + 1)) + 0.0000014);
            L = QL_AD_O540; break;
        }
        case QL_AD_O540: {
            c.ad_rhdn = 0.0;
            if (c.ad_nq == 1) { L = QL_AD_O560; break; }
            double w = 1.0;         // ddn = || YH_nq (at the current step size) || / tq1
            const double eta = c.ad_h / c.ad_hyh;
            for (int k = 0; k < c.ad_nq; k++) w *= eta;
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.red = QB_RED_WRMS;
            qb_pass_src(p, QB_AD_YH(0), 1.0, 0.0);
            qb_pass_src(p, QB_AD_YH(c.ad_nq), 0.0, w);
            c.pc = QB_PC_AD_DDN_DONE; return 1;
        }
        case QB_PC_AD_DDN_DONE: {
            const double ddn = sqrt(red[1] / (double)g.N) / T.bi[c.ad_nq][0];
            c.ad_rhdn = 1.0 / (1.3 * pow(ddn, 1.0 / c.ad_nq) + 0.0000013);
            L = QL_AD_O560; break;
        }
        case QL_AD_O560: {
            const int nq = c.ad_nq, l = nq + 1;
            const double rhsm = 1.0 / (1.2 * pow(c.ad_dsm, 1.0 / l) + 0.0000012);
            const double rhup = c.ad_rhup, rhdn = c.ad_rhdn;
            int newq; double rh;
            if (rhsm >= rhup) {
                if (rhsm < rhdn) { newq = nq - 1; rh = rhdn; }
                else { newq = nq; rh = rhsm; }
            } else if (rhup > rhdn) { newq = l; rh = rhup; }
            else { newq = nq - 1; rh = rhdn; }
            if (newq == nq - 1 && c.ad_kflag < 0 && rh > 1.0) rh = 1.0;
            if (newq == l) {        // order increase: new column YH_{nq+1} = acor l_nq / (nq+1)
                if (rh < 1.1) { c.ad_ialth = 3; L = QL_AD_STEP_DONE; break; }
                const double r = T.a[nq][nq] / (double)l;
                c.ad_newq = newq; c.ad_rh = rh;
                qb_pass_clear(p);
                p.kind = QB_PASS_COMBINE; p.dst1 = QB_AD_YH(l);
                qb_pass_src(p, QB_AD_SAVF(c.ad_fsel), r, 0.0);
                qb_pass_src(p, QB_AD_YP(1), -r, 0.0);
                c.pc = QB_PC_AD_NEWCOL_DONE; return 1;
            }
            if (c.ad_kflag == 0 && rh < 1.1) { c.ad_ialth = 3; L = QL_AD_STEP_DONE; break; }
            if (c.ad_kflag <= -2 && rh > 0.2) rh = 0.2;
            c.ad_nq = newq; c.ad_rh = rh;
            L = QL_AD_RESCALE; break;
        }
        case QB_PC_AD_NEWCOL_DONE:
            c.ad_nq = c.ad_newq;
            L = QL_AD_RESCALE; break;
        case QL_AD_RESCALE: {       // h <- rh h (the array keeps its scaling, see QL_AD_PRED_ISSUE)
            double rh = c.ad_rh;
            const double h = c.ad_h;
            if (g.opt.min_step > 0.0 && rh < g.opt.min_step / fabs(h)) rh = g.opt.min_step / fabs(h);
            if (rh > c.ad_rmax) rh = c.ad_rmax;
            if (g.opt.max_step != 0.0) {
                const double q = fabs(h) * rh / g.opt.max_step;
                if (q > 1.0) rh /= q;
            }
            c.ad_h = h * rh;
            c.ad_ialth = c.ad_nq + 1;
            if (c.ad_iredo == 0) { c.ad_rmax = 10.0; L = QL_AD_STEP_DONE; }
            else L = QL_AD_STEP;
            break;
        }
        case QL_AD_3FAIL: {         // three error-test failures: restart at order 1 with h / 10
            if (c.ad_kflag <= -10) { c.status = QB_ST_DT_UNDERFLOW; L = QL_FAIL; break; }
            double rh = 0.1;
            if (g.opt.min_step > 0.0 && rh < g.opt.min_step / fabs(c.ad_h)) rh = g.opt.min_step / fabs(c.ad_h);
            c.ad_h *= rh;
            L = QL_AD_3FAIL_ISSUE; break;
        }
        case QL_AD_3FAIL_ISSUE: {   // resumable (host-evaluated coefficients): the step is scaled once
            qb_pass_clear(p);       // YH1 = f(t_n, YH0), unscaled
            p.kind = QB_PASS_RHS; p.x = QB_AD_YH(0); p.zdst = QB_AD_YH(1);
            QB_COEFS_OR_PAUSE(c.ad_tn, QL_AD_3FAIL_ISSUE)
            c.n_rhs++;
            c.pc = QB_PC_AD_REF_DONE; return 1;
        }
        case QB_PC_AD_REF_DONE:
            c.ad_hyh = 1.0; c.ad_nq = 1; c.ad_ialth = 5; c.ad_pred_nq = -1;
            L = QL_AD_STEP; break;
        case QL_AD_STEP_DONE:
            L = c.int_step ? QL_AD_AFTER : QL_AD_LOOP; break;
        case QL_AD_AFTER:
            if (c.status < 0) { L = QL_FAIL; break; }
            if (c.t_front > c.int_t) { L = QL_AD_INTERP; break; }
            c.status = QB_ST_AT_FRONT;
            c.t = c.t_front; c.sY = QB_AD_YH(0); c.norm2_y = c.norm2_front;
            L = QL_INT_DONE; break;
        case QL_AD_INTERP: {        // dense output: y(t) = sum_j YH_j ((t - t_n) / h)^j
            const double sx = (c.int_t - c.ad_tn) / c.ad_hyh;
            qb_pass_clear(p);
            p.kind = QB_PASS_COMBINE; p.dst1 = c.sI; p.red = QB_RED_NORM2_O1;
            double w = 1.0;
            for (int j = 0; j <= c.ad_nq; j++) { qb_pass_src(p, QB_AD_YH(j), w, 0.0); w *= sx; }
            c.pc = QB_PC_AD_INTERP_DONE; return 1;
        }
        case QB_PC_AD_INTERP_DONE:
            c.status = QB_ST_INTERPOLATED;
            c.t = c.int_t; c.sY = c.sI; c.norm2_y = qb_mcn(g, red);
            L = QL_INT_DONE; break;

        // ================================================================ mesolve driver
        case QL_ME_NEXT:            // for t in tlist[1:]: integrate(t)   (integrator.py:197-212)
            if (c.tl_idx >= c.tl_end) { L = QL_FINISH; break; }
            c.done = 0;
            c.int_t = g.tlist[c.tl_idx]; c.int_step = 0; c.after_int = QB_K_ME_REACHED;
            L = QL_INT_BEGIN; break;

        // ---- record the state at tlist[tl_idx]: e_ops, optional stored state ----
        case QL_RECORD:
            c.exp_set = QB_OPSET_EOPS; c.exp_lo = 0; c.expect_mode = 0; c.exp_t = c.t;
            if (g.neops > 0) { L = QL_EXPECT_ISSUE; break; }
            L = QL_AFTER_RECORD; break;
        case QL_EXPECT_ISSUE: {
            const int nops = (c.exp_set == QB_OPSET_EOPS) ? g.neops : g.ncops;
            qb_pass_clear(p);
            p.kind = QB_PASS_EXPECT; p.opset = c.exp_set; p.x = c.sY;
            p.op_lo = c.exp_lo;
            const int chunk = (g.exp_chunk > 0 && g.exp_chunk < QB_MAXRED / 2) ? g.exp_chunk : QB_MAXRED / 2;
            p.op_hi = (c.exp_lo + chunk < nops) ? c.exp_lo + chunk : nops;
            c.pc = QB_PC_EXPECT_DONE; return 1;
        }
        case QB_PC_EXPECT_DONE: {
            const int nops = (c.exp_set == QB_OPSET_EOPS) ? g.neops : g.ncops;
            const int lo = c.exp_lo;
            const int chunk = (g.exp_chunk > 0 && g.exp_chunk < QB_MAXRED / 2) ? g.exp_chunk : QB_MAXRED / 2;
            const int hi = (lo + chunk < nops) ? lo + chunk : nops;
            bool bad = false;
            for (int m = lo; m < hi; m++) {
                qb_c128 v, cf; v.re = red[2 * (m - lo)]; v.im = red[2 * (m - lo) + 1];
                QbProgRef pr = (c.exp_set == QB_OPSET_EOPS) ? g.eop_prog[m] : g.nop_prog[m];
                if (qb_eval_ref(g, c, pr, c.exp_t, &cf)) { bad = true; break; }
                v = qb_cmul(cf, v);
                if (c.expect_mode == 0) {
                    // mcsolve returns y/||y|| (mcsolve.py:302); tlist[0] is recorded as given
                    if (c.mode == 1 && c.tl_idx > 0) { v.re /= c.norm2_y; v.im /= c.norm2_y; }
                    g.out_expect[((size_t)c.traj_id * g.neops + m) * g.nt + c.tl_idx] = v;
                } else {
                    probs[m] = v.re;        // mcsolve.py:384-387
                }
            }
            if (bad) { c.status = QB_ST_BAD_PROGRAM; L = QL_FAIL; break; }
            c.exp_lo = hi;
            if (hi < nops) { L = QL_EXPECT_ISSUE; break; }
            L = (c.expect_mode == 0) ? QL_AFTER_RECORD : QL_APPLY_ISSUE;
            break;
        }
        case QL_AFTER_RECORD:
            if (g.opt.store_states) {
                qb_pass_clear(p);
                p.kind = QB_PASS_COMBINE; p.dst1 = QB_SLOT_OUT; p.out_index = c.tl_idx;
                double sc = 1.0;
                if (c.mode == 1 && c.tl_idx > 0) sc = qb_mc_inv_norm(g, c.norm2_y);
                qb_pass_src(p, c.sY, sc, 0.0);
                c.pc = QB_PC_STORE_DONE; return 1;
            }
            /* fallthrough */
        case QB_PC_STORE_DONE:
            c.tl_idx++;
            L = (c.mode == 0) ? QL_ME_NEXT : QL_MC_ENTRY;
            break;

        // ================================================================ mcsolve driver
        case QL_MC_ENTRY:           // MCIntegrator.integrate(t)   (mcsolve.py:286-289)
            if (c.tl_idx >= c.tl_end) { L = QL_FINISH; break; }
            c.done = 0;
            c.mc_t_old = c.t; c.mc_n_old = c.norm2_y;
            L = QL_MC_LOOP; break;
        case QL_MC_LOOP:            // while t_old < t   (:289)
            if (c.mc_t_old < g.tlist[c.tl_idx]) {
                c.int_t = g.tlist[c.tl_idx]; c.int_step = 1; c.after_int = QB_K_MC_AFTER_STEP;
                L = QL_INT_BEGIN; break;
            }
            L = QL_RECORD; break;   // return t_old, y_old/||y_old||   (:302)

        // ---------------------------------------------------------------- _find_collapse_time
        case QL_RF_LOOP: {          // mcsolve.py:321-347
            if (c.rf_tries >= g.opt.norm_steps) { L = QL_RF_END; break; }
            c.rf_tries++;
            if ((c.rf_t_final - c.rf_t_prev) < g.opt.norm_t_tol) {
                c.rf_t_guess = c.rf_t_final;      // state = integrator.get_state()
                L = QL_RF_END; break;
            }
            const double dt = c.rf_t_final - c.rf_t_prev;
            double ratio = log(c.rf_n_old / c.target_norm) / log(c.rf_n_old / c.rf_n);
            if (ratio < g.opt.norm_min_step) ratio = g.opt.norm_min_step;
            if (ratio > (1.0 - g.opt.norm_min_step)) ratio = 1.0 - g.opt.norm_min_step;
            double t_guess = c.rf_t_prev + dt * ratio;
            if ((t_guess - c.rf_t_prev) < g.opt.norm_t_tol) t_guess = c.rf_t_prev + g.opt.norm_t_tol;
            c.rf_t_guess = t_guess;
            c.int_t = t_guess; c.int_step = 1; c.after_int = QB_K_RF_AFTER_GUESS;
            L = QL_INT_BEGIN; break;
        }
        case QL_RF_END:             // :363-369
            if (c.rf_tries >= g.opt.norm_steps) { c.status = QB_ST_ROOTFIND_FAILED; L = QL_FAIL; break; }
            L = QL_COLLAPSE; break;

        // ---------------------------------------------------------------- _do_collapse
        case QL_COLLAPSE:           // mcsolve.py:371-392 ; state = V[sY], time = rf_t_guess
            if (g.ncops == 1) { c.which = 0; L = QL_APPLY_ISSUE; c.expect_mode = 2; break; }
            c.exp_set = QB_OPSET_NOPS; c.exp_lo = 0; c.expect_mode = 1;
            // n_ops are evaluated at the collapse time (expect_data(collapse_time, state))
            c.exp_t = c.rf_t_guess;
            L = QL_EXPECT_ISSUE; break;
        case QL_APPLY_ISSUE: {
            if (c.expect_mode == 1) {       // choose the operator (:388-392)
                if (c.rng >= g.ndraws) { c.status = QB_ST_RNG_EXHAUSTED; L = QL_FAIL; break; }
                const double u = g.draws[(size_t)c.traj_id * g.ndraws + c.rng++];
                double sum = 0.0;
                for (int k = 0; k < g.ncops; k++) sum += probs[k];   // python sum(), left to right
                double target = sum * u - probs[0];
                int which = 0;
                bool bad = false;
                while (target > 0.0 && which <= g.ncops) {
                    which++;
                    if (which >= g.ncops) { bad = true; break; }      // reference: IndexError
                    target -= probs[which];
                }
                if (bad) { c.status = QB_ST_COLLAPSE_INDEX; L = QL_FAIL; break; }
                c.which = which;
            }
            qb_pass_clear(p);
            p.kind = QB_PASS_APPLY; p.opset = QB_OPSET_COPS; p.op_lo = c.which; p.op_hi = c.which + 1;
            p.x = c.sY; p.zdst = c.sTA; p.red = QB_RED_NORM2_Z;
            if (qb_eval_ref(g, c, g.cop_prog[c.which], c.rf_t_guess, &coef[0])) {
                c.status = QB_ST_BAD_PROGRAM; L = QL_FAIL; break;
            }
            c.pc = QB_PC_APPLY_DONE; return 1;
        }
        case QB_PC_PROBS_DONE: L = QL_FAIL; c.status = QB_ST_BAD_PROGRAM; break;   // unused
        case QB_PC_APPLY_DONE: {    // mcsolve.py:394-406
            const double new_norm = g.mc_trace ? red[4] : sqrt(red[2]);
            if (new_norm < g.opt.mc_corr_eps) {
                // numerical-error collapse: keep the state, renormalise, no record, no draw
                c.set_x = c.sY; c.set_scale = qb_mc_inv_norm(g, c.norm2_y);
            } else {
                c.set_x = c.sTA; c.set_scale = 1.0 / new_norm;
                if (c.ncol >= g.opt.max_collapses) { c.status = QB_ST_TOO_MANY_COLLAPSES; L = QL_FAIL; break; }
                g.out_col_t[(size_t)c.traj_id * g.opt.max_collapses + c.ncol] = c.rf_t_guess;
                g.out_col_which[(size_t)c.traj_id * g.opt.max_collapses + c.ncol] = c.which;
                c.ncol++;
                if (c.rng >= g.ndraws) { c.status = QB_ST_RNG_EXHAUSTED; L = QL_FAIL; break; }
                c.target_norm = g.draws[(size_t)c.traj_id * g.ndraws + c.rng++];
            }
            c.set_t = c.rf_t_guess;
            c.after_set = QB_K_MC_AFTER_COLLAPSE;
            L = QL_SET_BEGIN; break;
        }

        // ================================================================ exits
        case QL_FINISH:
            qb_pass_clear(p);
            c.pc = QB_PC_IDLE; c.done = 1;
            if (c.mode == 1) g.out_ncol[c.traj_id] = c.ncol;
            return 0;
        case QL_FAIL:
        default:
            qb_pass_clear(p);
            if (L != QL_FAIL) c.status = QB_ST_BAD_PROGRAM;
            c.pc = QB_PC_IDLE; c.done = c.status < 0 ? c.status : QB_ST_BAD_PROGRAM;
            if (c.mode == 1 && g.out_ncol) g.out_ncol[c.traj_id] = c.ncol;
            return 0;
        }
    }
    qb_pass_clear(p);
    c.pc = QB_PC_IDLE; c.done = QB_ST_BAD_PROGRAM;
    return 0;
}
