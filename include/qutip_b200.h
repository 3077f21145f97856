/*
 * qutip_b200.h -- C ABI of libqutip_b200.so (sm_100a).
 *
 * Drop-in boundary for QuTiP's time-evolution hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the qutip 5.4.0.dev tree).  All
 * pointers are plain host or device pointers, complex numbers are (re, im) pairs of
 * doubles (`double complex` / numpy complex128 layout), indices are int32 like the
 * reference's default `idxint`.  Every function returns 0 on success or a negative
 * QB_E_* code; qb_last_error() gives the message.  There is no CPU fallback: without a
 * CUDA device every compute call fails with QB_E_CUDA.
 *
 * Handles are opaque; the library owns the device memory behind them until qb_free().
 */
#ifndef QUTIP_B200_H
#define QUTIP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* qb_handle;

enum {
    QB_OK = 0, QB_E_CUDA = -1, QB_E_SHAPE = -2, QB_E_ARG = -3, QB_E_ALLOC = -4,
    QB_E_TYPE = -5, QB_E_STATE = -6
};

/* ---- library ---- */
int qb_version(void);
const char* qb_last_error(void);
int qb_device_count(int* n);
int qb_set_device(int dev);
int qb_device_mem_info(int64_t* free_bytes, int64_t* total_bytes);   /* current device */
int qb_synchronize(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t qb_launch_count(void);

/* ---- data-layer objects -------------------------------------------------------------
 * Dense  : core/data/dense.pxd:9-22  (contiguous complex128, `fortran` flag)
 * CSR    : core/data/csr.pxd:17-30   (data, col_index, row_index; may be unsorted)
 * Dia    : core/data/dia.pxd:17-29   (SciPy DIA: data[num_diag][ncols], offsets)
 * Uploads are the `to(DeviceType, HostType)` conversions registered with
 * qutip.core.data.to.add_conversions (core/data/convert.pyx:208-329). */
int qb_dense_upload(const void* host, int64_t rows, int64_t cols, int fortran, qb_handle* out);
int qb_dense_zeros(int64_t rows, int64_t cols, int fortran, qb_handle* out);
int qb_dense_download(qb_handle h, void* host);
int qb_dense_write(qb_handle h, const void* host);      /* overwrite from host memory */
int qb_dense_copy(qb_handle h, qb_handle* out);
int qb_dense_reshape(qb_handle h, int64_t rows, int64_t cols);   /* column-major buffer, same size, no copy */
int qb_dense_info(qb_handle h, int64_t* rows, int64_t* cols, int* fortran, void** devptr);
/* operator formats: 0 auto (rule-compressed sliced ELLPACK for small L2-resident operators
 * whose 32-row slices are diagonal structured -- no per-element column index, one constant per
 * constant diagonal --, plain sliced ELLPACK for other small operators with little padding,
 * diagonal-masked slices for large diagonal-structured operators, CSR otherwise),
 * 1 force CSR, 2 force DIAM, 3 force SELL, 5 force RSELL */
int qb_csr_upload(const void* data, const int32_t* col, const int32_t* rowptr,
                  int64_t rows, int64_t cols, int64_t nnz, int format, qb_handle* out);
int qb_dia_upload(const void* data, const int32_t* offsets, int64_t ndiag,
                  int64_t rows, int64_t cols, int format, qb_handle* out);
/* Matrix-free superoperator of an n x n CSR operator A on the column-stacked n x n state:
 * side 0 = I (x) A (rho -> A rho), side 1 = conj(A) (x) I (rho -> rho A^dagger) -- the two
 * products LindbladMatrixForm.matmul_data is made of (core/cy/lindblad_matrix_form.pyx:105-203,
 * imatmul_data_dense / imatmul_dag_dense_data core/data/matmul.pyx:1218-1248).  The handle is
 * an operator of shape (n^2, n^2) usable wherever a CSR/Dia upload is. */
int qb_kron_upload(const void* data, const int32_t* col, const int32_t* rowptr, int64_t n,
                   int64_t nnz, int side, qb_handle* out);
/* Matrix-free jump part of the Lindblad equation, rho -> sum_c C_c rho C_c^dagger
 * (= sum_c conj(C_c) (x) C_c on the column-stacked state; lindblad_matrix_form.pyx:177-181,
 * 195-199: c_op.matmul_data then c_op.adjoint_rmatmul_data): the nstack n x n operators are
 * passed stacked row-wise as ONE CSR matrix of shape (nstack*n, n). */
int qb_sandwich_upload(const void* data, const int32_t* col, const int32_t* rowptr,
                       int64_t n, int64_t nstack, int64_t nnz, qb_handle* out);
/* Liouvillian assembled on the device (replaces the host kron / add chain of
 * qutip.liouvillian, core/superoperator.py:116-142): A = -iH - 1/2 sum_k C_k^+ C_k (n x n CSR,
 * built by the caller) and the nstack collapse operators stacked row-wise as one
 * (nstack*n, n) CSR matrix give
 *   L = I (x) A + conj(A) (x) I + sum_k conj(C_k) (x) C_k      (column-stacked, canonical CSR)
 * one thread per row: count, scan, fill, sort, merge, tidy, compact.  tol: the reference's
 * auto_tidyup_atol (components below it are zeroed, entries then exactly zero are dropped --
 * core/data/csr.pxd:88-113; 0 drops exact zeros only).  format as qb_csr_upload;
 * format 1 keeps the device CSR as it is (the operator never visits the host). */
int qb_liouvillian_build(const void* a_data, const int32_t* a_col, const int32_t* a_rowptr, int64_t a_nnz,
                         const void* c_data, const int32_t* c_col, const int32_t* c_rowptr, int64_t c_nnz,
                         int64_t n, int64_t nstack, double tol, int format, qb_handle* out);
/* Kronecker product A (x) B of two CSR matrices assembled on the device (core/data/kron.pyx,
 * kron_csr): one warp per output row; canonical CSR when A and B are.  format as above. */
int qb_kron_build(const void* a_data, const int32_t* a_col, const int32_t* a_rowptr, int64_t a_rows,
                  int64_t a_cols, const void* b_data, const int32_t* b_col, const int32_t* b_rowptr,
                  int64_t b_rows, int64_t b_cols, int format, qb_handle* out);
/* Format conversion of a CSR-format operator (the `_data.to(Dia, CSR)` family of
 * core/data/convert.pyx:208-329, csr.pyx:710, dia.pyx:364): format 2 (diagonal-masked slices)
 * is produced on the device -- one warp per 32-row slice merges its rows by diagonal offset;
 * the column-rule / padding analysers of formats 0, 3, 5 run on the host. */
int qb_op_convert(qb_handle h, int format, qb_handle* out);
/* copy a CSR-format operator back: data[nnz] complex128, col[nnz], rowptr[rows+1] (qb_op_info sizes) */
int qb_op_csr_download(qb_handle h, void* data, int32_t* col, int32_t* rowptr);
int qb_op_info(qb_handle h, int* fmt, int64_t* rows, int64_t* cols, int64_t* nnz,
               int64_t* device_bytes);
int qb_free(qb_handle h);

/* ---- data-layer operations (Dispatcher specialisations) ----------------------------
 * qb_matmul        : matmul_csr_dense_dense / matmul_dia_dense_dense / matmul_dense
 *                    (core/data/matmul.pyx:226-347,429-507)  out += scale * A @ X
 * qb_axpy          : iadd_dense   core/data/add.pyx:203-218       y += a * x
 * qb_scal          : imul_dense   core/data/mul.pyx:73-79         x *= a
 * qb_nrm2          : frobenius_dense / l2_dense  core/data/norm.pyx:127-135
 * qb_wrms_error    : wrmn_error_dense  core/data/ode.pyx:39-64
 * qb_inner         : inner_dense  core/data/inner.pyx   <a|b> (conj on a unless !conj)
 * qb_expect_ket    : expect_csr_dense ket branch  core/data/expect.pyx:133-144
 * qb_expect_dm     : expect_csr_dense dm branch   core/data/expect.pyx:146-158  tr(A rho)
 * qb_expect_super  : expect_super_csr_dense       core/data/expect.pyx:223-236
 * qb_trace_oper_ket: trace_oper_ket_dense         core/data/trace.pyx:78-86 */
int qb_matmul(qb_handle op, qb_handle x, double scale_re, double scale_im, qb_handle out);
/* dense column-major A times a block of columns as a complex128 GEMM on the FP64 tensor
 * cores (DMMA): out += scale * A @ X   (matmul_dense zgemm branch, matmul.pyx:329-346).
 * qb_matmul routes here for dense operands with >= 8 columns. */
int qb_zgemm(qb_handle a, qb_handle x, double scale_re, double scale_im, qb_handle out);
int qb_zgemm_bench(qb_handle a, qb_handle x, qb_handle out, int iters, double* ms_total);
/* measured FP64 tensor-core (DMMA) peak of the current device in TFLOP/s (register-only chains) */
int qb_dmma_peak_bench(int iters, double* tflops);
int qb_axpy(qb_handle x, double a_re, double a_im, qb_handle y);
int qb_scal(qb_handle x, double a_re, double a_im);
int qb_copy(qb_handle src, qb_handle dst);
int qb_zero(qb_handle x);
int qb_nrm2(qb_handle x, double* out);
int qb_wrms_error(qb_handle diff, qb_handle state, double atol, double rtol, double* out);
int qb_inner(qb_handle a, qb_handle b, int conj_a, double out[2]);
int qb_expect_ket(qb_handle op, qb_handle x, double out[2]);
int qb_expect_dm(qb_handle op, qb_handle rho, double out[2]);
int qb_expect_super(qb_handle op, qb_handle vec, double out[2]);
int qb_trace_oper_ket(qb_handle vec, double out[2]);

/* ---- fused evolution engine ---------------------------------------------------------
 * A system is what QobjEvo.matmul_data sums (core/cy/qobjevo.pyx:1103-1116): elements
 * coeff_k(t) * A_k, plus for mcsolve the collapse operators c_k / n_k = c_k^dag c_k
 * (solver/mcsolve.py:468-502) and the e_ops (solver/result.py:303-339).  Coefficients are
 * stack programs (struct qb_instr, see qb_types.h QbInstr) compiled by the host from the
 * reference's Coefficient objects; prog == NULL means the constant 1. */
typedef struct { int op; int iarg; double re, im; } qb_instr;

int qb_system_create(int64_t N, int nargs, qb_handle* out);
int qb_system_add_element(qb_handle sys, qb_handle op, const qb_instr* prog, int nprog);
int qb_system_add_collapse(qb_handle sys, qb_handle c_op, const qb_instr* cprog, int ncprog,
                           qb_handle n_op, const qb_instr* nprog, int nnprog);
/* functional = 1: the operator is diag(w) and the "expectation" is sum_r w[r]*y[r]
 * (tr(E rho) on a column-stacked rho, core/data/expect.pyx:146-158) */
int qb_system_add_eop(qb_handle sys, qb_handle op, const qb_instr* prog, int nprog);
int qb_system_set_eop_functional(qb_handle sys, int functional);
/* mcsolve with a super-operator H (solver/mcsolve.py:481-490): the state is the column-stacked
 * n x n rho (system size n*n), the jump logic compares tr(rho).real with the threshold
 * (mcsolve.py:311-319), probabilities are tr(n_k rho), states are renormalised by their trace */
int qb_system_set_mc_trace(qb_handle sys, int n);
/* InterCoefficient tables (core/cy/coefficient.pyx:412-549): returns the spline id */
int qb_system_add_spline(qb_handle sys, const double* tlist, const void* poly,
                         int n, int order, double dt, int* id);

/* integrator options, defaults = qutip_integrator.py:51-59 and mcsolve.py:460-465 */
typedef struct {
    double atol, rtol;
    int nsteps;
    double first_step, min_step, max_step;
    int interpolate;
    int norm_steps;
    double norm_t_tol, norm_tol, norm_min_step, mc_corr_eps;
    int store_states;
    int max_collapses;
    int no_jump;
    double jump_prob_floor;
    int max_order;             /* adams: highest order (1..12; 0 = 12), scipy_integrator.py:27 */
    int pad_;
} qb_options;
int qb_options_default(qb_options* opt);

/* tableau: 0 vern7, 1 vern9 (solver/integrator/verner{7,9}efficient.py), 2 tsit5 (FSAL,
 * solver/integrator/tsit5.py), 3 adams: variable-order Adams-Moulton in Nordsieck form on the
 * device (counterpart of IntegratorScipyAdams, solver/integrator/scipy_integrator.py:20-196;
 * options: nsteps, max_order, first_step, min_step, max_step, atol, rtol) */
int qb_engine_create(qb_handle sys, int tableau, int nslots, const qb_options* opt,
                     qb_handle* out);

/* Batched run = Solver.run / MultiTrajSolver.run for `ntraj` independent trajectories.
 *   mode 0: mesolve-like (integrate through tlist, solver_base.py:159-226)
 *   mode 1: mcsolve (jump detection + collapse, mcsolve.py:228-414)
 * Host inputs : init_states[ninit][N] complex, init_map[ntraj] (NULL = all use state 0),
 *               tlist[nt], args[ntraj][nargs] complex (NULL if nargs == 0),
 *               draws[ntraj][ndraws] = the successive generator.random() values of each
 *               trajectory (multitraj.py:354-393); mode 1 only.
 * Host outputs: expect[ntraj][neops][nt] complex, status[ntraj] (1 ok, <0 QB_ST_*),
 *               ncol[ntraj], col_t[ntraj][max_collapses], col_which[...] (mode 1),
 *               stats[ntraj][4] = {rhs evals, accepted, rejected, passes},
 *               states[ntraj][nt][N] complex when opt.store_states,
 *               final_states[ntraj][N] (may be NULL), final_t[ntraj] (may be NULL).
 * Any output pointer may be NULL. */
int qb_engine_run(qb_handle eng, int mode, int64_t ntraj,
                  const void* init_states, int64_t ninit, const int32_t* init_map,
                  const double* tlist, int nt,
                  const void* args, const double* draws, int ndraws,
                  void* expect, int32_t* status, int32_t* ncol, double* col_t,
                  int32_t* col_which, int32_t* stats, void* states, void* final_states);
/* same, but inputs already resident and outputs left on the device (device pointers;
 * used by bench.py's HBM-resident timing and by the multi-GPU reduce) */
int qb_engine_run_device(qb_handle eng, int mode, int64_t ntraj,
                         const void* d_init_states, int64_t ninit, const int32_t* d_init_map,
                         const double* d_tlist, int nt, const void* d_args,
                         const double* d_draws, int ndraws,
                         void* d_expect, int32_t* d_status, int32_t* d_ncol, double* d_col_t,
                         int32_t* d_col_which, int32_t* d_stats, void* d_states);
/* sum over trajectories for the multi-GPU reduce (multitrajresult.py:1116-1124):
 * d_sums[2][neops][nt] complex = (sum_j e_j, sum_j e_j^2 elementwise on re/im) */
int qb_reduce_expect(const void* d_expect, int64_t ntraj, int neops, int nt, void* d_sums);
/* rounds (pass+control launch pairs) and kernel time of the last run */
int qb_engine_last_run_info(qb_handle eng, int64_t* rounds, double* gpu_ms);

/* Integrator protocol on slot 0 (Integrator ABC, solver/integrator/integrator.py:23-240;
 * IntegratorVern7 qutip_integrator.py:69-92).  `y` are host pointers to N complex. */
int qb_integ_set_state(qb_handle eng, double t, const void* y);
int qb_integ_integrate(qb_handle eng, double t, int step, double* t_out, int* status);
int qb_integ_get_state(qb_handle eng, double* t, void* y);
/* Python-callable coefficients (FunctionCoefficient, core/cy/coefficient.pyx:176) cannot run
 * on the device: elements whose program is the single instruction QB_I_HOST are evaluated
 * by the host.  qb_integ_set_state returns 3 / qb_integ_integrate reports status 3 when the
 * controller needs their values at the time given by qb_integ_pending_coef; the host calls
 * qb_integ_resume with vals[nelem] (complex) until the status is no longer 3.  The matvecs,
 * stage combinations and the step controller stay on the device. */
int qb_integ_pending_coef(qb_handle eng, double* t);
int qb_integ_resume(qb_handle eng, const void* vals, double* t_out, int* status);
int qb_integ_set_args(qb_handle eng, const void* args);
int qb_integ_stats(qb_handle eng, int64_t stats[4]);

/* ---- measurement hooks ----
 * qb_engine_rhs      : one RHS evaluation out = sum_k c_k(t) A_k x on device vectors
 *                      (QobjEvo.matmul_data, core/cy/qobjevo.pyx:1103-1116)
 * qb_engine_rhs_bench: `iters` of them back to back, timed with CUDA events
 * profiling          : when on, every pass-kernel launch of a run is bracketed by CUDA
 *                      events on the launching stream; qb_engine_profile returns their
 *                      summed duration, the launch count and the number of state-sized
 *                      vector accesses the passes had to make (algorithmic traffic) */
int qb_engine_rhs(qb_handle eng, double t, qb_handle x, qb_handle out);
/* same with the coefficient values vals[nelem] (complex) supplied by the caller */
int qb_engine_rhs_coef(qb_handle eng, const void* vals, qb_handle x, qb_handle out);
int qb_engine_rhs_bench(qb_handle eng, double t, qb_handle x, qb_handle out, int iters,
                        double* ms_total);
int qb_engine_set_profiling(qb_handle eng, int on);
int qb_engine_profile(qb_handle eng, double* pass_ms, int64_t* pass_launches,
                      double* state_vector_accesses);
/* per pass launch of the last profiled run: time (ms) and cumulative state-vector accesses */
int qb_engine_profile_rounds(qb_handle eng, double* ms, double* cum_vec, int64_t max, int64_t* n);

/* ---- multi-GPU: the one collective of the sharded workloads -----------------------------
 * mcsolve trajectories (and sweep members) are independent given their seeds
 * (solver/multitraj.py:250-256; map dispatch solver/parallel.py:541-559); the only exchange
 * is the sum _TrajectorySum.reduce_expect accumulates (solver/multitrajresult.py:1116-1124).
 * A communicator is an NCCL group (bound at run time, libnccl.so.2):
 *   qb_comm_init_all  : one process drives ndev devices (ncclCommInitAll); local member i
 *                       lives on devs[i].
 *   qb_comm_unique_id / qb_comm_init_rank : one process per device (ncclCommInitRank on the
 *                       current device); the 128-byte id of rank 0 is passed to the other
 *                       ranks by the launcher.
 *   qb_comm_allreduce_sum : bufs[i] = count doubles in host memory of local member i, summed
 *                       over the whole group in place with ONE ncclAllReduce.
 *   qb_comm_reduce_expect : engines[i] (on member i's device, NULL if that member ran
 *                       nothing) has just finished qb_engine_run; the per-trajectory expectation
 *                       values still on the devices are reduced to [2][n_e][n_t] complex sums
 *                       (sum e, sum (re^2, im^2)) per member, summed with ONE ncclAllReduce over
 *                       NVLink and returned in sums (host, 4 * n_e * n_t doubles). */
int qb_comm_nccl_version(int* version);
int qb_comm_init_all(int ndev, const int* devs, qb_handle* out);
int qb_comm_unique_id(void* id, int nbytes);
int qb_comm_init_rank(int nranks, int rank, const void* id, qb_handle* out);
int qb_comm_info(qb_handle comm, int* nranks, int* nlocal);
int qb_comm_allreduce_sum(qb_handle comm, double* const* bufs, int64_t count);
int qb_comm_allreduce_sum_device(qb_handle comm, void* const* dbufs, int64_t count);   /* device buffers */
int qb_comm_reduce_expect(qb_handle comm, const qb_handle* engines, int neops, int nt, void* sums);

#ifdef __cplusplus
}
#endif
#endif
