"""Device-resident Adams integrator (csrc/qb_adams.h) through the C ABI: the reference
fixtures at the accuracy the reference asks of its own Adams method
(tests/solver/test_integrator.py:71-98: 5e-5), the Integrator protocol, batches."""
import numpy as np
import pytest

import qutip_b200 as qb
from _golden import load
from _systems import mc_system_from_golden, me_system_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1_jc", "c2_tfim4", "c4_driven", "c5_kerr_0"])
@pytest.mark.parametrize("fmt", [qb.FMT_AUTO, qb.FMT_CSR])
def test_mesolve_adams_vs_reference_fixtures(name, fmt):
    g = load(name)
    eng = qb.Engine(me_system_from_golden(g, fmt), "adams", nslots=1, store_states=1, nsteps=2500)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert r.status[0] == 1
    scale = max(1.0, np.abs(g["states_vern7"]).max())
    assert np.abs(r.states[0] - g["states_vern7"]).max() < 5e-5 * scale
    assert np.abs(r.expect[0] - g["expect_vern7"]).max() < 5e-5 * max(1.0, np.abs(g["expect_vern7"]).max())
    nrhs, nacc, nrej, npass = r.stats[0]
    # about one RHS evaluation per step: far fewer than vern7's 10 per step
    assert nrhs < 2.0 * (nacc + nrej) + 5


def test_adams_agrees_with_host_emulation():
    """Same controller code on the CPU emulator: same number of steps and RHS evaluations
    unless a decision sits within round-off of its threshold (allow a small difference)."""
    from _emul import FMT_CSR, EmulSystem, default_options
    from _golden import op_arrays
    g = load("c1_jc")
    s = EmulSystem(len(g["y0"]), 0, FMT_CSR)
    for i in range(int(g["n_elements"])):
        s.add_element(*op_arrays(g, "el%d" % i))
    e = s.run(0, 3, g["y0"], g["tlist"], opt=default_options(store_states=1, nsteps=2500))
    eng = qb.Engine(me_system_from_golden(g, qb.FMT_CSR), "adams", nslots=1, store_states=1, nsteps=2500)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert abs(int(r.stats[0][0]) - int(e["stats"][0][0])) <= 3
    assert np.abs(r.states[0] - e["states"][0]).max() < 1e-7


def test_adams_integrator_protocol():
    g = load("c1_jc")
    eng = qb.Engine(me_system_from_golden(g), "adams", nslots=1, nsteps=2500)
    tl = g["tlist"]
    eng.set_state(tl[0], g["y0"])
    t, y = eng.get_state()
    assert t == tl[0] and np.array_equal(y, g["y0"])
    for k in (5, 10):
        t, st = eng.integrate(tl[k], False)
        assert t == tl[k] and st in (1, 2)
        assert np.abs(eng.get_state()[1] - g["states_vern7"][k]).max() < 5e-5
    # mcstep: one internal step, then interpolation back inside the last step
    t1, st = eng.integrate(tl[10] + 50.0, True)
    assert tl[10] < t1 < tl[10] + 50.0 and st == 2
    t2, st = eng.integrate(0.5 * (tl[10] + t1), True)
    assert t2 == 0.5 * (tl[10] + t1) and st == 1
    # before the last step: refused like the reference's RK integrators
    t3, st = eng.integrate(tl[2], True)
    assert st == -3


@pytest.mark.parametrize("name,nslots", [("c3_tfim6_mc", 24), ("c3_tfim6_mc", 5),
                                         ("c3_tfim4_mc_strong", 7)])
def test_mcsolve_adams_vs_reference_fixture(name, nslots):
    g = load(name)
    eng = qb.Engine(mc_system_from_golden(g), "adams", nslots=nslots, nsteps=2500)
    ntraj = int(g["ntraj"])
    r = eng.run_mcsolve(g["psi0"], g["tlist"], g["draws"], ntraj=ntraj)
    assert (r.status == 1).all()
    assert np.array_equal(r.ncol, g["col_count"])
    cc = np.concatenate([[0], np.cumsum(g["col_count"])])
    for j in range(ntraj):
        n = r.ncol[j]
        assert np.array_equal(r.col_which[j, :n], g["col_which"][cc[j]:cc[j + 1]])
        np.testing.assert_allclose(r.col_t[j, :n], g["col_times"][cc[j]:cc[j + 1]], rtol=0, atol=1e-3)
    assert np.abs(np.transpose(r.expect, (1, 0, 2)) - g["runs_expect"]).max() < 5e-4


def test_adams_max_order_and_failure_status():
    g = load("c1_jc")
    eng = qb.Engine(me_system_from_golden(g), "adams", nslots=1, store_states=1, nsteps=20000,
                    max_order=2, atol=1e-9, rtol=1e-7)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert r.status[0] == 1
    assert np.abs(r.states[0] - g["states_vern7"]).max() < 5e-5
    full = qb.Engine(me_system_from_golden(g), "adams", nslots=1, nsteps=20000, atol=1e-9, rtol=1e-7)
    rf = full.run_mesolve(g["y0"], g["tlist"])
    assert rf.stats[0][0] < r.stats[0][0]             # order 12 needs fewer RHS evaluations than order 2
    eng = qb.Engine(me_system_from_golden(g), "adams", nslots=1, nsteps=4)
    assert eng.run_mesolve(g["y0"], g["tlist"]).status[0] == -1


@pytest.mark.parametrize("name", ["c1_jc", "c2_tfim4", "c4_driven"])
def test_adams_pinned_against_zvode_golden(name):
    """The device-resident Adams method against the reference's method='adams' (SciPy zvode)
    golden runs (tests/golden/adams_zvode.npz): within 1e-8 / 1e-6 at tight tolerances; at the
    default tolerances as close to the converged solution as zvode itself (up to a factor),
    with a comparable number of RHS evaluations."""
    z = load("adams_zvode")
    g = load(name)
    tight = z[name + "_tight_expect"]
    sysm = me_system_from_golden(g, qb.FMT_AUTO)
    eng = qb.Engine(sysm, "adams", nslots=1, store_states=1, nsteps=100000, atol=1e-12, rtol=1e-10)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    assert r.status[0] == 1
    np.testing.assert_allclose(r.expect[0], tight, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(r.states[0][-1], z[name + "_tight_final"], rtol=1e-6, atol=1e-8)
    eng = qb.Engine(sysm, "adams", nslots=1, store_states=1, nsteps=2500)
    r = eng.run_mesolve(g["y0"], g["tlist"])
    zerr = max(np.abs(z[name + "_default_expect"] - tight).max(), 1e-7)
    assert np.abs(r.expect[0] - tight).max() < 4 * zerr
    assert r.stats[0][0] < 1.5 * int(z[name + "_default_nst_nfe"][1]) + 20
