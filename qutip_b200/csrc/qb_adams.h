// qb_adams.h -- variable-order, variable-step Adams-Moulton integrator in Nordsieck form
// (orders 1..12, functional corrector iteration), the device-resident counterpart of the
// reference's method="adams" (IntegratorScipyAdams, qutip/solver/integrator/
// scipy_integrator.py:20-196, which drives SciPy's zvode with method "adams", order 12).
//
// zvode's source is not part of the reference tree (SciPy ships it compiled), and the
// reference pins its step sequence only loosely (tests/solver/test_integrator.py:71-98:
// 5e-5 against analytic results), so this is NOT a restatement of zvode: it is the classic
// fixed-leading-coefficient Nordsieck Adams method of the ODEPACK family -- corrector
// coefficients and error constants generated from the Adams-Moulton polynomials, local
// error test on the corrector change, order changed by +-1 from the three error estimates,
// step size changed by rescaling the Nordsieck array.  Its results agree with the reference
// within the integration tolerance; step counts differ.
//
// Work vectors (k-slots of the engine's pool): YH[0..12] the Nordsieck array at t_n scaled
// for step size ad_hyh, YP[0..12] the predicted array of the running attempt, two buffers
// for h*f of the corrector iterates; the iterates themselves use the tmpA / tmpB slots,
// y_prev holds the saved corrector change used for the order-increase estimate.
// Step-size changes never touch YH: the prediction applies eta^k = (h / ad_hyh)^k on the fly.
#pragma once
#include "qb_types.h"

#define QB_AD_MAXORD 12
// The local error test runs at QB_AD_TOL_SCALE x the requested atol / rtol: with the plain
// tolerances this fixed-leading-coefficient method lands up to 13x further from the converged
// solution than zvode (variable-coefficient formulas) does at the same settings on the golden
// cases (tests/golden/adams_zvode.npz); at 1/8 it is within 2.5x of zvode's own error on
// every case for 7-30 % more RHS evaluations (tests/test_adams_emul.py).
#define QB_AD_TOL_SCALE 0.125
#define QB_AD_YH(j) (j)
#define QB_AD_YP(j) (13 + (j))
#define QB_AD_SAVF(k) (26 + (k))
#define QB_AD_NVEC 28

// Corrector coefficients l_j (a[nq][j], j = 0..nq) and error constants (bi[nq][0..2]) of the
// Adams-Moulton methods of order nq = 1..12: with p(x) = (x+1)(x+2)...(x+nq-1),
//   l_0 = int_{-1}^0 p / (nq-1)!,  l_1 = 1,  l_j = p_{j-1} / (j (nq-1)!)   (p_i = coeff of x^i),
//   1 / bi[nq][1] = int_{-1}^0 x p / nq!   (local error constant of order nq),
// bi[nq][0] and bi[nq][2] the corresponding constants for the estimates at order nq-1 / nq+1.
QB_HD void qb_adams_table(QbTableau* T) {
    double (*el)[QB_MAX_STAGES] = T->a;
    double (*te)[QB_MAX_DENSE_ORDER] = T->bi;
    for (int i = 0; i < QB_MAX_STAGES; i++) {
        for (int j = 0; j < QB_MAX_STAGES; j++) el[i][j] = 0.0;
        for (int j = 0; j < QB_MAX_DENSE_ORDER; j++) te[i][j] = 0.0;
        T->b[i] = T->c[i] = T->e[i] = 0.0;
    }
    T->order = QB_AD_MAXORD; T->s = 0; T->S = QB_AD_NVEC; T->dense_order = 0; T->fsal = 0;
    T->method = 1;
    double pc[QB_AD_MAXORD + 1];
    el[1][0] = 1.0; el[1][1] = 1.0;
    te[1][0] = 0.0; te[1][1] = 2.0; te[2][0] = 1.0; te[12][2] = 0.0;
    pc[0] = 1.0;
    double rqfac = 1.0;
    for (int nq = 2; nq <= QB_AD_MAXORD; nq++) {
        const double rq1fac = rqfac;             // 1 / (nq-1)!
        rqfac /= nq;                             // 1 / nq!
        const double fnqm1 = nq - 1;
        // p(x) <- p(x) * (x + nq - 1)
        pc[nq - 1] = 0.0;
        for (int i = nq - 1; i >= 1; i--) pc[i] = pc[i - 1] + fnqm1 * pc[i];
        pc[0] = fnqm1 * pc[0];
        // integrals over [-1, 0] of p(x) and x p(x)
        double pint = pc[0], xpin = pc[0] / 2.0, tsign = 1.0;
        for (int i = 1; i < nq; i++) {
            tsign = -tsign;
            pint += tsign * pc[i] / (i + 1);
            xpin += tsign * pc[i] / (i + 2);
        }
        el[nq][0] = pint * rq1fac;
        el[nq][1] = 1.0;
        for (int i = 1; i < nq; i++) el[nq][i + 1] = rq1fac * pc[i] / (i + 1);
        const double ragq = 1.0 / (rqfac * xpin);
        te[nq][1] = ragq;
        if (nq < QB_AD_MAXORD) te[nq + 1][0] = ragq * rqfac / (nq + 1);
        te[nq - 1][2] = ragq;
    }
}
