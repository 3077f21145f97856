"""Host-side logic of the QuTiP plug-in that needs no GPU: registration with QuTiP's
extension surfaces, QobjEvo -> device-system translation, coefficient compilation from the
reference's Coefficient objects, and the refusals of host-only forms.  Runs against the
reference build in oracle/_ref (skipped when it is absent)."""
import sys
import warnings

import numpy as np
import pytest

import oracle

_ref = oracle.ref_path()
if _ref is None:
    pytest.skip("reference build (oracle/_ref) not present", allow_module_level=True)
sys.path.insert(0, _ref)
warnings.filterwarnings("ignore")
import qutip  # noqa: E402
from qutip import QobjEvo, destroy, qeye, tensor  # noqa: E402
import qutip_b200.plugin as plugin  # noqa: E402
from qutip_b200 import coeffs  # noqa: E402

pytestmark = pytest.mark.ref


def test_registration_surfaces():
    from qutip.core import data as _data
    for solver in (qutip.MESolver, qutip.SESolver, qutip.MCSolver):
        av = solver.avail_integrators()
        for key in ("b200_vern7", "b200_vern9", "b200_tsit5", "b200_adams", "b200_zvode"):
            assert key in av and issubclass(av[key], qutip.solver.integrator.Integrator)
    assert plugin.B200Dense in _data.to.dtypes and plugin.B200Operator in _data.to.dtypes
    assert _data.to.parse("b200") is plugin.B200Dense
    assert qutip.solver.parallel._maps["b200"] is plugin.b200_map
    # same option keys and defaults as the stock Verner integrators
    stock = qutip.solver.integrator.IntegratorVern7.integrator_options
    assert plugin.B200Vern7.integrator_options == stock
    assert plugin.B200Vern7.options.__doc__ and "atol" in plugin.B200Vern7.options.__doc__


def test_bind_qobjevo_merges_constants_and_compiles_coefficients():
    a = tensor(destroy(4), qeye(3)); b = tensor(qeye(4), destroy(3))
    tl = np.linspace(0, 1, 11)
    H = QobjEvo([5 * a.dag() * a, [a + a.dag(), "A*cos(w*t)"], [b + b.dag(), np.sin(tl)]],
                args={"A": 0.2, "w": 5.0}, tlist=tl)
    rhs = qutip.MESolver(H, [0.1 * a]).rhs

    class FakeSystem:
        splines = []

        def add_spline(self, t, poly, dt):
            self.splines.append((t, poly, dt))
            return len(self.splines) - 1

    fs = FakeSystem()
    els = plugin.bind_qobjevo(rhs, fs)
    assert els[-1][1] is None                              # merged constant part last
    assert sum(1 for _, p in els if p is None) == 1
    progs = [p for _, p in els if p is not None]
    assert len(progs) >= 3 and len(fs.splines) >= 1
    # the compiled string coefficient agrees with the reference's own evaluation
    ref_coeff = qutip.coefficient("A*cos(w*t)", args={"A": 0.2, "w": 5.0})
    prog = plugin.coefficient_to_program(ref_coeff)
    for t in (0.0, 0.37, 2.5):
        assert abs(coeffs.evaluate(prog, t) - ref_coeff(t)) < 1e-14
    conj = plugin.coefficient_to_program(ref_coeff.conj() * ref_coeff + ref_coeff)
    assert abs(coeffs.evaluate(conj, 0.4) - (abs(ref_coeff(0.4)) ** 2 + ref_coeff(0.4))) < 1e-14


def test_host_only_forms_are_detected():
    a = destroy(5)
    with pytest.raises(TypeError, match="device"):
        plugin.bind_qobjevo(QobjEvo(lambda t: a * np.cos(t)))
    with pytest.raises(TypeError, match="cannot be evaluated on the device"):
        plugin.bind_qobjevo(QobjEvo([a, lambda t: np.cos(t)]))
    els = plugin.bind_qobjevo(QobjEvo([a, lambda t: np.cos(t)]), allow_host=True)
    assert isinstance(els[0][1], plugin._HostCoefficient)
    assert abs(els[0][1].coeff(0.3) - np.cos(0.3)) < 1e-15
    with pytest.raises(TypeError, match="QobjEvo.matmul_data"):
        plugin.B200Vern7(lambda t, y: y, {})


def test_matrix_form_is_bound_as_superoperator():
    a = destroy(4)
    H = QobjEvo(a.dag() * a + 0.2 * (a + a.dag()))
    from qutip.core.cy.lindblad_matrix_form import LindbladMatrixForm
    lmf = LindbladMatrixForm(H, [QobjEvo(0.3 * a)])
    sup = plugin._device_qevo(lmf)
    ref = qutip.liouvillian(H(0), [0.3 * a])
    assert np.abs(sup(0).full() - ref.full()).max() < 1e-14


def test_integrators_registered_on_other_qobjevo_solvers_but_not_on_the_base_class():
    """HEOMSolver / BRSolver / FMESolver integrate a constant QobjEvo: they list the device
    methods; the Solver base class does not (its integrators must accept arbitrary callables,
    tests/solver/test_integrator.py parametrises over them)."""
    from qutip.solver.brmesolve import BRSolver
    from qutip.solver.floquet import FMESolver
    from qutip.solver.heom.bofin_solvers import HEOMSolver
    from qutip.solver.solver_base import Solver
    for cls in (BRSolver, FMESolver, HEOMSolver):
        assert "b200_vern7" in cls.avail_integrators() and "b200_adams" in cls.avail_integrators()
    assert not any(k.startswith("b200") for k in Solver.avail_integrators())


def test_nm_mcsolve_rate_coefficients_compile_to_device_programs():
    """The rate-shifted collapse operators of NonMarkovianMCSolver (solver/nm_mcsolve.py:400-440:
    sqrt(rate + shift) with shift = 2 |min(0, rates)|, solver/cy/nm_mcsolve.pyx) are bound
    without host evaluation, and the compiled programs reproduce the reference coefficients."""
    from qutip import NonMarkovianMCSolver, coefficient, sigmam, sigmap, sigmax, sigmaz
    H = 0.5 * sigmaz() + 0.2 * sigmax()
    rates = [(sigmam(), coefficient("0.25*sin(2*t) + 0.05")), (sigmap(), 0.15)]
    solver = NonMarkovianMCSolver(H, rates, options={"progress_bar": False})
    checked = 0
    for q in list(solver.rhs.c_ops) + list(solver.rhs.n_ops) + [solver.rhs.rhs]:
        for el in q.to_list():
            if isinstance(el, (list, tuple)):
                prog = plugin.coefficient_to_program(el[1])
                for t in (0.0, 0.4, 1.7, 2.3, 3.9):          # both signs of the first rate
                    want = complex(el[1](t))
                    got = complex(coeffs.evaluate(prog, t))
                    assert abs(want - got) <= 1e-13 * max(1.0, abs(want))
                    checked += 1
    assert checked >= 10


def test_python_rate_function_is_rejected_for_device_batches():
    from qutip import NonMarkovianMCSolver, coefficient, sigmam, sigmax
    solver = NonMarkovianMCSolver(sigmax(), [(sigmam(), coefficient(lambda t: -1 + t))],
                                  options={"progress_bar": False})
    el = [e for e in solver.rhs.c_ops[0].to_list() if isinstance(e, (list, tuple))][0]
    with pytest.raises(TypeError):
        plugin.coefficient_to_program(el[1])
