// qb_coeff.h -- time-dependent coefficient programs evaluated on the device.
//
// Replaces the reference's Coefficient._call(t) family (core/cy/coefficient.pyx:88-988,
// generated StrCoefficient core/coefficient.py:573-576): the host compiles each
// Coefficient into a small stack program (qutip_b200/coeffs.py); the step controller
// evaluates it for the stage time of every RHS evaluation, so no host round trip is
// needed for time-dependent systems.  All arithmetic is complex128 like the reference's
// generated code (args are `double complex`).
#pragma once
#include <math.h>
#include "qb_types.h"

struct QbSpline {            // InterCoefficient (coefficient.pyx:412-539)
    int n, order, uniform, pad_;
    double dt;
    long long t_off;         // offset of tlist[n] in the spline pool (doubles)
    long long p_off;         // offset of poly[(order+1)][n] complex, stored re,im pairs
};

QB_HD qb_c128 qb_cmul(qb_c128 a, qb_c128 b) {
    qb_c128 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}
QB_HD qb_c128 qb_cadd(qb_c128 a, qb_c128 b) { qb_c128 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
QB_HD qb_c128 qb_cdiv(qb_c128 a, qb_c128 b) {
    qb_c128 r;
    if (b.im == 0.0) { r.re = a.re / b.re; r.im = a.im / b.re; return r; }
    // Smith's algorithm
    if (fabs(b.re) >= fabs(b.im)) {
        double q = b.im / b.re, d = b.re + b.im * q;
        r.re = (a.re + a.im * q) / d; r.im = (a.im - a.re * q) / d;
    } else {
        double q = b.re / b.im, d = b.re * q + b.im;
        r.re = (a.re * q + a.im) / d; r.im = (a.im * q - a.re) / d;
    }
    return r;
}
QB_HD qb_c128 qb_cexp(qb_c128 a) {
    double e = exp(a.re); qb_c128 r;
    if (a.im == 0.0) { r.re = e; r.im = 0.0; } else { r.re = e * cos(a.im); r.im = e * sin(a.im); }
    return r;
}
QB_HD qb_c128 qb_clog(qb_c128 a) {
    qb_c128 r; r.re = log(hypot(a.re, a.im)); r.im = atan2(a.im, a.re); return r;
}
QB_HD qb_c128 qb_csqrt(qb_c128 a) {
    qb_c128 r;
    if (a.im == 0.0) {
        if (a.re >= 0.0) { r.re = sqrt(a.re); r.im = 0.0; } else { r.re = 0.0; r.im = sqrt(-a.re); }
        return r;
    }
    double m = hypot(a.re, a.im);
    double s = sqrt(0.5 * (m + fabs(a.re)));
    if (a.re >= 0.0) { r.re = s; r.im = a.im / (2.0 * s); }
    else { r.re = fabs(a.im) / (2.0 * s); r.im = a.im >= 0.0 ? s : -s; }
    return r;
}
QB_HD qb_c128 qb_csin(qb_c128 a) {
    qb_c128 r;
    if (a.im == 0.0) { r.re = sin(a.re); r.im = 0.0; }
    else { r.re = sin(a.re) * cosh(a.im); r.im = cos(a.re) * sinh(a.im); }
    return r;
}
QB_HD qb_c128 qb_ccos(qb_c128 a) {
    qb_c128 r;
    if (a.im == 0.0) { r.re = cos(a.re); r.im = 0.0; }
    else { r.re = cos(a.re) * cosh(a.im); r.im = -sin(a.re) * sinh(a.im); }
    return r;
}
QB_HD qb_c128 qb_csinh(qb_c128 a) {
    qb_c128 r;
    if (a.im == 0.0) { r.re = sinh(a.re); r.im = 0.0; }
    else { r.re = sinh(a.re) * cos(a.im); r.im = cosh(a.re) * sin(a.im); }
    return r;
}
QB_HD qb_c128 qb_ccosh(qb_c128 a) {
    qb_c128 r;
    if (a.im == 0.0) { r.re = cosh(a.re); r.im = 0.0; }
    else { r.re = cosh(a.re) * cos(a.im); r.im = sinh(a.re) * sin(a.im); }
    return r;
}
QB_HD qb_c128 qb_cpow(qb_c128 a, qb_c128 b) {
    qb_c128 r;
    if (a.im == 0.0 && b.im == 0.0 && (a.re >= 0.0 || b.re == floor(b.re))) {
        r.re = pow(a.re, b.re); r.im = 0.0; return r;
    }
    if (a.re == 0.0 && a.im == 0.0) { r.re = 0.0; r.im = 0.0; return r; }
    return qb_cexp(qb_cmul(b, qb_clog(a)));
}

// InterCoefficient._call (coefficient.pyx:517-539): clamp outside the range, locate the
// interval (uniform grid or binary search), Horner in (t - tlist[idx]).
QB_HD qb_c128 qb_spline_eval(const QbSpline& s, const double* pool, double t) {
    const double* tl = pool + s.t_off;
    const double* poly = pool + s.p_off;
    qb_c128 r; r.re = 0.0; r.im = 0.0;
    int n = s.n, idx;
    if (t <= tl[0]) { const double* c = poly + 2 * ((size_t)s.order * n); r.re = c[0]; r.im = c[1]; return r; }
    if (t >= tl[n - 1]) { const double* c = poly + 2 * ((size_t)s.order * n + n - 1); r.re = c[0]; r.im = c[1]; return r; }
    if (s.uniform) {
        idx = (int)((t - tl[0]) / s.dt);
        if (idx > n - 1) idx = n - 1;
    } else {
        int lo = 0, hi = n;      // coefficient.pyx:_binary_search
        while (lo + 1 != hi) { int mid = (lo + hi) >> 1; if (t < tl[mid]) hi = mid; else lo = mid; }
        idx = lo;
    }
    double x = t - tl[idx];
    for (int k = 0; k <= s.order; k++) {
        const double* c = poly + 2 * ((size_t)k * n + idx);
        double nre = r.re * x + c[0], nim = r.im * x + c[1];
        r.re = nre; r.im = nim;
    }
    return r;
}

// Evaluate one program.  Returns 0 on success.
QB_HD int qb_eval_prog(const QbInstr* code, int len, double t, const qb_c128* args,
                       const QbSpline* splines, const double* spool, qb_c128* out) {
    qb_c128 st[16];
    int sp = 0;
    for (int pc = 0; pc < len; pc++) {
        const QbInstr in = code[pc];
        qb_c128 a, b, r;
        switch (in.op) {
        case QB_I_CONST: if (sp >= 16) return -1; st[sp].re = in.re; st[sp].im = in.im; sp++; break;
        case QB_I_T: if (sp >= 16) return -1; st[sp].re = t; st[sp].im = 0.0; sp++; break;
        case QB_I_ARG: if (sp >= 16) return -1; st[sp] = args[in.iarg]; sp++; break;
        case QB_I_SPLINE:
            if (sp < 1) return -1;
            st[sp - 1] = qb_spline_eval(splines[in.iarg], spool, st[sp - 1].re); break;
        case QB_I_ADD: case QB_I_SUB: case QB_I_MUL: case QB_I_DIV: case QB_I_POW:
        case QB_I_HEAVISIDE_GE: case QB_I_MIN_RE:
            if (sp < 2) return -1;
            b = st[--sp]; a = st[sp - 1];
            if (in.op == QB_I_ADD) { r.re = a.re + b.re; r.im = a.im + b.im; }
            else if (in.op == QB_I_SUB) { r.re = a.re - b.re; r.im = a.im - b.im; }
            else if (in.op == QB_I_MUL) r = qb_cmul(a, b);
            else if (in.op == QB_I_DIV) r = qb_cdiv(a, b);
            else if (in.op == QB_I_POW) r = qb_cpow(a, b);
            else if (in.op == QB_I_MIN_RE) { r.re = (b.re < a.re) ? b.re : a.re; r.im = 0.0; }
            else { r.re = (a.re >= b.re) ? 1.0 : 0.0; r.im = 0.0; }
            st[sp - 1] = r; break;
        default:
            if (sp < 1) return -1;
            a = st[sp - 1];
            switch (in.op) {
            case QB_I_NEG: r.re = -a.re; r.im = -a.im; break;
            case QB_I_CONJ: r.re = a.re; r.im = -a.im; break;
            case QB_I_SIN: r = qb_csin(a); break;
            case QB_I_COS: r = qb_ccos(a); break;
            case QB_I_TAN: r = qb_cdiv(qb_csin(a), qb_ccos(a)); break;
            case QB_I_EXP: r = qb_cexp(a); break;
            case QB_I_LOG: r = qb_clog(a); break;
            case QB_I_SQRT: r = qb_csqrt(a); break;
            case QB_I_ABS: r.re = hypot(a.re, a.im); r.im = 0.0; break;
            case QB_I_NORM2: r.re = a.re * a.re + a.im * a.im; r.im = 0.0; break;
            case QB_I_REAL: r.re = a.re; r.im = 0.0; break;
            case QB_I_IMAG: r.re = a.im; r.im = 0.0; break;
            case QB_I_SINH: r = qb_csinh(a); break;
            case QB_I_COSH: r = qb_ccosh(a); break;
            case QB_I_TANH: r = qb_cdiv(qb_csinh(a), qb_ccosh(a)); break;
            case QB_I_ASIN: r.re = asin(a.re); r.im = 0.0; break;
            case QB_I_ACOS: r.re = acos(a.re); r.im = 0.0; break;
            case QB_I_ATAN: r.re = atan(a.re); r.im = 0.0; break;
            case QB_I_SQRT_RE: r.re = sqrt(a.re); r.im = 0.0; break;
            default: return -1;
            }
            st[sp - 1] = r;
        }
    }
    if (sp != 1) return -1;
    *out = st[0];
    return 0;
}
