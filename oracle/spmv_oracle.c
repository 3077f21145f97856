/*
 * oracle/spmv_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's CPU matvec kernels for the hot path.  Used by
 * tests/ and by bench.py's cpu_baseline leg as the checker for the CUDA kernels.
 * Parity is pinned against the reference itself: tests/test_oracle.py compares these
 * functions with qutip.core.data.matmul on the golden fixtures (tests/golden/).
 *
 * Each function cites the reference routine it follows.
 *   build:  gcc -O2 -fPIC -shared -o oracle/liboracle.so oracle/spmv_oracle.c -lm
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

typedef struct { double re, im; } zc;

/* out[r] += scale * sum_p data[p] * vec[col[p]]
 * follows qutip/core/data/src/matmul_csr_vector.cpp:134-176 (scalar path):
 * the row dot product is accumulated first, then scaled, then added to out. */
void orc_csr_matvec(const zc *data, const int32_t *col, const int32_t *rowptr,
                    int64_t nrows, const zc *vec, double sre, double sim, zc *out)
{
    for (int64_t r = 0; r < nrows; r++) {
        double dre = 0.0, dim = 0.0;
        for (int32_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
            double are = data[p].re, aim = data[p].im;
            double vre = vec[col[p]].re, vim = vec[col[p]].im;
            dre += are * vre - aim * vim;
            dim += are * vim + aim * vre;
        }
        out[r].re += sre * dre - sim * dim;
        out[r].im += sre * dim + sim * dre;
    }
}

/* out[r, c] += scale * sum_p data[p] * X[col[p], c], X and out row-major (C order)
 * follows qutip/core/data/src/matmul_csr_dense.cpp:20-70. */
void orc_csr_matmat_c(const zc *data, const int32_t *col, const int32_t *rowptr,
                      int64_t nrows, int64_t ncols_x, const zc *X,
                      double sre, double sim, zc *out)
{
    for (int64_t r = 0; r < nrows; r++) {
        for (int32_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
            double are = sre * data[p].re - sim * data[p].im;
            double aim = sre * data[p].im + sim * data[p].re;
            const zc *xr = X + (size_t)col[p] * ncols_x;
            zc *o = out + (size_t)r * ncols_x;
            for (int64_t c = 0; c < ncols_x; c++) {
                o[c].re += are * xr[c].re - aim * xr[c].im;
                o[c].im += are * xr[c].im + aim * xr[c].re;
            }
        }
    }
}

/* Dia (SciPy convention: element (r, c) of diagonal k is data[k*ncols + c], c - r = off[k]).
 * follows qutip/core/data/matmul.pyx:429-507 + src/matmul_diag_vector.cpp:57-83:
 * accumulate the un-scaled product into tmp (caller passes a zeroed tmp), one
 * diagonal at a time; scaling is a second pass (orc_axpy). */
void orc_dia_matvec_noscale(const zc *data, const int32_t *offsets, int64_t ndiag,
                            int64_t nrows, int64_t ncols, const zc *vec, zc *tmp)
{
    for (int64_t d = 0; d < ndiag; d++) {
        int64_t off = offsets[d];
        int64_t start_left = off > 0 ? off : 0;
        int64_t end_left = ncols < nrows + off ? ncols : nrows + off;
        int64_t start_out = off < 0 ? -off : 0;
        int64_t end_out = nrows < ncols - off ? nrows : ncols - off;
        int64_t length = end_left - start_left;
        if (end_out - start_out < length) length = end_out - start_out;
        const zc *dl = data + d * ncols + start_left;
        const zc *vr = vec + start_left;
        zc *o = tmp + start_out;
        for (int64_t i = 0; i < length; i++) {
            o[i].re += dl[i].re * vr[i].re - dl[i].im * vr[i].im;
            o[i].im += dl[i].re * vr[i].im + dl[i].im * vr[i].re;
        }
    }
}

/* y += a * x   (zaxpy; qutip/core/data/add.pyx:203-218) */
void orc_axpy(int64_t n, double are, double aim, const zc *x, zc *y)
{
    for (int64_t i = 0; i < n; i++) {
        y[i].re += are * x[i].re - aim * x[i].im;
        y[i].im += are * x[i].im + aim * x[i].re;
    }
}

/* dense column-major (Fortran) A[nrows x ncols] times vector; zgemv semantics
 * (qutip/core/data/matmul.pyx:297-313): out += scale * A x */
void orc_dense_matvec_f(const zc *A, int64_t nrows, int64_t ncols, const zc *x,
                        double sre, double sim, zc *out)
{
    for (int64_t c = 0; c < ncols; c++) {
        double xre = sre * x[c].re - sim * x[c].im;
        double xim = sre * x[c].im + sim * x[c].re;
        const zc *a = A + (size_t)c * nrows;
        for (int64_t r = 0; r < nrows; r++) {
            out[r].re += a[r].re * xre - a[r].im * xim;
            out[r].im += a[r].re * xim + a[r].im * xre;
        }
    }
}

/* sqrt( 1/N * sum ( |diff_i| / (atol + rtol*|state_i|) )^2 )
 * follows qutip/core/data/ode.pyx:39-64 */
double orc_wrmn_error(int64_t n, const zc *diff, const zc *state, double atol, double rtol)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double q = hypot(diff[i].re, diff[i].im)
                   / (atol + rtol * hypot(state[i].re, state[i].im));
        s += q * q;
    }
    return sqrt(s / (double)n);
}
