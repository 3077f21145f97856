// qb_types.h -- plain-old-data shared by the host API, the CUDA kernels and the
// host-side unit-test harness of the step controller (tests/emul).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define QB_HD __host__ __device__ __forceinline__
#else
#define QB_HD inline
#endif

#define QB_MAX_STAGES 26      // vern9 with dense output
#define QB_MAX_DENSE_ORDER 9
#define QB_MAX_ELEMS 16       // fused QobjEvo elements per RHS evaluation
#define QB_MAXSRC 28          // sources of one linear combination (y_prev + all k)
#define QB_MAXRED 64          // reduction outputs per pass (32 complex expectations)
#define QB_SLICE 32           // rows per DIAM slice == warp size
#ifndef QB_TILE_ROWS
#define QB_TILE_ROWS 256      // rows per CTA of the pass kernel (8 warps)
#endif

struct qb_c128 { double re, im; };

// ---- Butcher tableau (see qb_tableaux.h; reference explicit_rk.pyx:127-192) ----
struct QbTableau {
    int order, s, S, dense_order, fsal;
    double a[QB_MAX_STAGES][QB_MAX_STAGES];
    double b[QB_MAX_STAGES];
    double c[QB_MAX_STAGES];
    double e[QB_MAX_STAGES];
    double bi[QB_MAX_STAGES][QB_MAX_DENSE_ORDER];
    // 0: explicit Runge-Kutta (the fields above).  1: variable-order Adams-Moulton in
    // Nordsieck form (qb_adams.h): a[nq][0..nq] = corrector coefficients l_j of order nq,
    // bi[nq][0..2] = the three error-test constants of order nq, S = number of work vectors
    int method;
};

// ---- operator storage on the device ----
enum { QB_FMT_CSR = 0, QB_FMT_DIAM = 1, QB_FMT_DENSE = 2, QB_FMT_SELL = 3, QB_FMT_KRON = 4,
       QB_FMT_RSELL = 5 };

// RSELL ("rule" sliced ELLPACK): SELL whose slots carry a 32-byte descriptor read with two
// warp-uniform loads.  A slot of a 32-row slice is one diagonal of the slice: its column is
// given by a rule -- row + delta, row ^ delta (tensor-product / spin operators flip one bit
// of the row index) -- or by an explicit 32-lane column block, and its value is one constant
// for the whole slot or an explicit 32-lane value block.  Diagonal-structured operators
// shrink from 20 B per stored element to 32 B per (slice, diagonal) + the non-constant
// diagonals, and the sweep needs no per-element index loads.
enum { QB_RS_COL_EXPL = 0, QB_RS_COL_ADD = 1, QB_RS_COL_XOR = 2, QB_RS_COL_MASK = 3,
       QB_RS_VAL_CONST = 4 };
struct alignas(16) QbSlotDesc {
    int rule;        // QB_RS_COL_* | QB_RS_VAL_CONST
    int delta;       // operand of the column rule
    int cpos;        // explicit columns: col[(slice column base + cpos) * 32 + lane]
    int vpos;        // explicit values:  val[(slice value base + vpos) * 32 + lane]
    double vre, vim; // constant value
};

struct QbOpDev {
    int fmt, nrows, ncols, pad_;
    long long nnz;
    // CSR (reference layout, core/data/csr.pxd:17-30)
    const qb_c128* val;
    const int* col;
    const int* rowptr;
    // DIAM: per 32-row slice a list of (diagonal offset, lane mask); values packed in
    // (entry, lane) order.  slice_ptr[nslices+1] indexes ent[]; slice_vbase[nslices]
    // is the index of the slice's first packed value.
    const int* slice_ptr;
    const int* ent_off;
    const unsigned* ent_mask;
    const long long* slice_vbase;
    // SELL (sliced ELLPACK, 32-row slices, slot-major): slice_ptr[nslices+1] counts slots;
    // slot k of slice s holds val/col[(slice_ptr[s] + k) * 32 + lane]; padding has val = 0
    // DENSE: column-major A[nrows x ncols]
    const qb_c128* dense;
    // KRON (matrix-free superoperator of an n x n operator A acting on the column-stacked
    // n x n state, core/cy/lindblad_matrix_form.pyx:105-203): kside 0 = I (x) A  (rho -> A rho),
    // kside 1 = conj(A) (x) I  (rho -> rho A^dagger).  A is kept as CSR in val/col/rowptr and,
    // when n is a multiple of 32, also as SELL in kval/kcol/slice_ptr for the left product.
    int kn, kside;            // kside 2: sandwich sum_c C_c rho C_c^dagger, C_c stacked row-wise
    int kstack, kpad_;
    const qb_c128* kval;
    const int* kcol;
    // RSELL: sinfo[nslices][4] = (first descriptor, count | fast count << 12 | xor-fast count
    // << 24, value-block base, column-block base) of every slice -- a list holds the xor-rule
    // constant-value slots first, then the add-rule ones ("fast": no explicit block), then the rest; sdesc[ndesc] the de-duplicated descriptor lists; explicit blocks
    // live in val / col
    const QbSlotDesc* sdesc;
    const int* sinfo;
    int ndesc, rpad_;
};

// descriptor lists of a system's RSELL elements in the constant bank (kernel parameter of the
// fused pass kernel): warp-uniform reads that do not touch the L1 data pipe
#define QB_CD_MAX 80
struct QbConstDesc {
    int n;                          // 0: lists are read from global memory
    int elem_off[QB_MAX_ELEMS];     // first descriptor of element e in d[]
    int cpad_[3];
    QbSlotDesc d[QB_CD_MAX];
};

// ---- coefficient byte-code (qb_coeff.h) ----
enum {
    QB_I_CONST = 0, QB_I_T, QB_I_ARG, QB_I_ADD, QB_I_SUB, QB_I_MUL, QB_I_DIV, QB_I_NEG,
    QB_I_CONJ, QB_I_SIN, QB_I_COS, QB_I_TAN, QB_I_EXP, QB_I_LOG, QB_I_SQRT, QB_I_ABS,
    QB_I_REAL, QB_I_IMAG, QB_I_POW, QB_I_SINH, QB_I_COSH, QB_I_TANH, QB_I_SPLINE,
    QB_I_ASIN, QB_I_ACOS, QB_I_ATAN, QB_I_NORM2, QB_I_HEAVISIDE_GE,
    QB_I_HOST,    // value supplied by the host for the pending evaluation time (python callables)
    QB_I_MIN_RE,  // binary: min of the real parts (RateShiftCoefficient, solver/cy/nm_mcsolve.pyx:56-66)
    QB_I_SQRT_RE  // unary: real sqrt of the real part (SqrtRealCoefficient, nm_mcsolve.pyx:113-115)
};
struct QbInstr { int op; int iarg; double re, im; };

// ---- status codes: >= 0 as the reference Status enum (explicit_rk.pxd:5-12) ----
enum {
    QB_ST_AT_FRONT = 2, QB_ST_INTERPOLATED = 1, QB_ST_NORMAL = 0,
    QB_ST_TOO_MUCH_WORK = -1, QB_ST_DT_UNDERFLOW = -2, QB_ST_OUTSIDE_RANGE = -3,
    QB_ST_NOT_INITIATED = -4,
    QB_ST_ROOTFIND_FAILED = -10,   // mcsolve.py:363-367 RuntimeError
    QB_ST_COLLAPSE_INDEX = -11,    // mcsolve.py:388-392 IndexError corner
    QB_ST_RNG_EXHAUSTED = -12,     // host must supply a longer threshold table
    QB_ST_TOO_MANY_COLLAPSES = -13,
    QB_ST_BAD_PROGRAM = -14,
    QB_ST_CORRECTOR_FAILED = -15   // Adams corrector iteration did not converge (zvode istate -5)
};

// ---- pass descriptor: one vector instruction executed by the pass kernel ----
enum { QB_PASS_NONE = 0, QB_PASS_RHS, QB_PASS_APPLY, QB_PASS_EXPECT, QB_PASS_COMBINE,
       QB_PASS_LINMAP };   // LINMAP: several outputs, each a linear combination of src[0..nsrc)
enum { QB_RED_NORM2_O1 = 1, QB_RED_WRMS = 2, QB_RED_NORM2_Z = 4 };
enum { QB_OPSET_EOPS = 0, QB_OPSET_NOPS = 1, QB_OPSET_COPS = 2 };
#define QB_SLOT_INIT (-2)     // source vector = the trajectory's initial state buffer
#define QB_SLOT_OUT (-3)      // destination = the stored-states output buffer

// Laid out for 16-byte warp-uniform loads in the hot epilogue: (x, zdst, dst1, nsrc),
// (red, out_index, zscale), (w1z, w2z) and one (slot, weight) pair per source.
struct QbSrcW { int src; int pad_; double w1; };
struct alignas(16) QbPass {
    int kind;
    int opset, op_lo, op_hi;   // EXPECT: ops [lo,hi) of opset ; APPLY: c_ops[op_lo]
    int x;                     // operator input slot
    int zdst;                  // slot receiving z = zscale * Op(x), or -1
    int dst1;                  // slot receiving o1, or -1
    int nsrc;
    int red;                   // QB_RED_* mask
    int out_index;             // tlist index for QB_SLOT_OUT / expectation records
    double zscale;
    double w1z, w2z;           // weight of z in o1 / o2
    QbSrcW sw[QB_MAXSRC];      // source slot and its weight in o1
    double w2[QB_MAXSRC];      // weight of the source in o2
};
static_assert(offsetof(QbPass, x) == 16 && offsetof(QbPass, red) == 32 && offsetof(QbPass, w1z) == 48 &&
              offsetof(QbPass, sw) == 64 && sizeof(QbSrcW) == 16 && sizeof(QbPass) % 16 == 0,
              "QbPass layout is read with 16-byte loads");

// weights / destinations of a LINMAP pass (one per trajectory slot, next to the QbPass):
// V[dst[j]] = sum_k w[j][k] V[src[k]]; used by the Adams integrator for the Nordsieck
// prediction and update, which touch every history vector once
#define QB_LM_MAXSRC 15
#define QB_LM_MAXOUT 28
struct QbLinMap {
    int nout, pad_;
    int dst[QB_LM_MAXOUT];
    double w[QB_LM_MAXOUT][QB_LM_MAXSRC];
};

// ---- program counter of the per-trajectory controller (qb_control.h) ----
enum {
    QB_PC_IDLE = 0,
    QB_PC_ME_BEGIN, QB_PC_MC_BEGIN,          // entry points written by the host
    QB_PC_ME_NEXT, QB_PC_MC_ENTRY,           // resume points written by the host
    QB_PC_STEP_ENTRY,                        // Integrator.mcstep from the host
    QB_PC_SET_DONE, QB_PC_EST0_DONE, QB_PC_EST1IN_DONE, QB_PC_EST1_DONE,
    QB_PC_STAGE_DONE, QB_PC_DENSEIN_DONE, QB_PC_DENSE_DONE, QB_PC_INTERP_DONE,
    QB_PC_EXPECT_DONE, QB_PC_STORE_DONE, QB_PC_PROBS_DONE, QB_PC_APPLY_DONE,
    QB_PC_COPY_DONE, QB_PC_SETCOPY_DONE,     // tile mode: explicit y_prev <- y_front copies
    // Adams (qb_adams.h)
    QB_PC_AD_SET0_DONE, QB_PC_AD_F0_DONE, QB_PC_AD_PRED_DONE, QB_PC_AD_CORR_DONE,
    QB_PC_AD_DSM_DONE, QB_PC_AD_UPD_DONE, QB_PC_AD_SAVE_DONE, QB_PC_AD_DUP_DONE,
    QB_PC_AD_DDN_DONE, QB_PC_AD_NEWCOL_DONE, QB_PC_AD_REF_DONE, QB_PC_AD_INTERP_DONE
};
// continuations
enum {
    QB_K_PAUSE = 0, QB_K_ME_REACHED, QB_K_ME_START, QB_K_MC_START, QB_K_MC_AFTER_STEP,
    QB_K_RF_AFTER_GUESS, QB_K_MC_AFTER_COLLAPSE
};

struct QbTraj {
    // ---- mirror of Explicit_RungeKutta's scalars ----
    double t, t_prev, t_front, dt_int, dt_safe;
    double dt_cur;             // dt of the step being computed
    double int_t;              // target of the running integrate() call
    double norm2_y, norm2_front;
    double est_norm, est_tol, est_dt1;
    int status;                // reference Status (or QB_ST_* error)
    int int_step;              // step flag of the running integrate() call
    int nsteps_left, step_n;
    int stage;
    int cur_tmp;               // slot holding the current stage input
    int sP, sF, sI, sTA, sTB, sY;   // slot labels
    int after_set, after_int;  // continuations
    int pc;
    // ---- trajectory bookkeeping ----
    int traj_id;               // index into the per-trajectory output arrays
    int init_idx;              // which initial state
    int mode;                  // 0 mesolve, 1 mcsolve
    int tl_idx, tl_end;
    int done;                  // 1 = finished/paused, <0 = failed with that status
    // ---- MCIntegrator state (mcsolve.py:228-414) ----
    double target_norm, mc_t_old, mc_n_old;
    double rf_t_prev, rf_t_final, rf_n_old, rf_n, rf_t_guess;
    double set_t, set_scale;
    double exp_t;              // time at which the running EXPECT chain's coefficients are taken
    int set_x;
    int rf_tries;
    int rng;                   // next unread draw
    int ncol;
    int which;
    int exp_set;               // opset of the running EXPECT chain
    int exp_lo;
    int expect_mode;           // what the EXPECT chain feeds (0 e_ops record, 1 probs)
    // ---- host-evaluated coefficients (python callables): the controller pauses with
    //      done == 2 and hc_t = the time it needs them for; the host writes the values into
    //      the slot's coefficient buffer, sets hc_valid and resumes at pc = hc_resume ----
    double hc_t;
    int hc_valid, hc_resume, stage_arg;
    // ---- FSAL tableaux (tsit5): k[0] of a step is k[s-1] of the previous accepted step;
    //      done by exchanging the two slot labels at the start of the next step ----
    int kswap, fsal_pending;
    // ---- statistics ----
    int n_rhs, n_accept, n_reject, n_pass;
    // ---- Adams-Moulton / Nordsieck integrator state (qb_adams.h) ----
    double ad_tn;              // time of the Nordsieck array
    double ad_h;               // step size of the running / next attempt (0: not chosen yet)
    double ad_hyh;             // step size the stored array YH is scaled for
    double ad_hu;              // last successful step size
    double ad_rmax, ad_crate, ad_del, ad_delp, ad_dsm, ad_rhup, ad_rhdn, ad_rh, ad_f0n2;
    int ad_nq, ad_ialth, ad_kflag, ad_ncf, ad_m, ad_iredo, ad_j, ad_fsel, ad_ysel, ad_newq;
    // the update pass of an accepted step also wrote the prediction of the next one, valid
    // for step size ad_pred_h at order ad_pred_nq (-1: none)
    double ad_pred_h;
    int ad_pred_nq, ad_pad_;
};

// ---- options (defaults = reference: qutip_integrator.py:51-59, mcsolve.py:460-465) ----
struct QbOptions {
    double atol, rtol;
    int nsteps;
    double first_step, min_step, max_step;
    int interpolate;
    int norm_steps;
    double norm_t_tol, norm_tol, norm_min_step, mc_corr_eps;
    int store_states;          // 1: write state at every tlist point to out_states
    int max_collapses;         // capacity of the per-trajectory collapse record
    int no_jump;               // mcsolve: sample the no-jump trajectory (target 0)
    double jump_prob_floor;    // improved sampling floor (mcsolve.py:276-279)
    int max_order;             // Adams: highest order used (0 = QB_AD_MAXORD)
    int pad_;
};
