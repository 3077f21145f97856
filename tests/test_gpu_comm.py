"""GPU tests of the NCCL communicator behind the C ABI (qb_comm_*): the one collective of
the sharded workloads (solver/multitrajresult.py:1116-1124 across shards)."""
import numpy as np
import pytest

import qutip_b200 as qb
from qutip_b200 import engine as E
from qutip_b200 import models, solve

pytestmark = pytest.mark.gpu


def _mc_inputs(n=6, ntraj=24):
    H, c_ops, sz = models.tfim(n)
    return [models.heff(H, c_ops)], c_ops, models.basis_state(n), np.linspace(0, 2, 11), [sz[0], sz[2]]


def test_single_member_group_is_identity_and_matches_host_sums():
    comm = E.Comm.all([0])
    assert (comm.nranks, comm.nlocal) == (1, 1)
    a = np.arange(12, dtype=np.float64) * 0.5
    b = a.copy()
    comm.allreduce_sum([b])
    np.testing.assert_array_equal(a, b)
    heff, c_ops, psi0, tl, e_ops = _mc_inputs()
    res = solve.mcsolve(heff, c_ops, psi0, tl, 24, seeds=9, e_ops=e_ops)
    s1, s2 = comm.reduce_expect([res.engine], len(e_ops), len(tl))
    h = solve.local_expect_sums(res.runs_expect)
    np.testing.assert_allclose(s1, h[0], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(s2, h[1], rtol=1e-13, atol=1e-13)
    avg, std = solve.finish_expect_sums(s1, s2, 24)
    np.testing.assert_allclose(avg, res.average_expect, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(std, res.std_expect, rtol=1e-9, atol=1e-12)


def test_unique_id_rank_group_of_one():
    uid = E.Comm.unique_id()
    assert len(uid) == E.Comm.ID_BYTES
    comm = E.Comm.rank(1, 0, uid)
    x = np.array([1.0 + 2.0j, -3.0j])
    comm.allreduce_sum([x])
    np.testing.assert_array_equal(x, [1.0 + 2.0j, -3.0j])


def test_trajectories_sharded_over_two_devices_match_one_device():
    """devices=[0, 1]: identical per-trajectory records (jump times and indices bit-exact,
    they depend only on the seed-derived thresholds) and NCCL-reduced averages."""
    if qb.device_count() < 2:
        pytest.skip("needs two GPUs")
    heff, c_ops, psi0, tl, e_ops = _mc_inputs(6, 37)
    one = solve.mcsolve(heff, c_ops, psi0, tl, 37, seeds=9, e_ops=e_ops)
    two = solve.mcsolve(heff, c_ops, psi0, tl, 37, seeds=9, e_ops=e_ops, devices=[0, 1])
    assert [list(w) for w in two.col_which] == [list(w) for w in one.col_which]
    assert [list(t) for t in two.col_times] == [list(t) for t in one.col_times]
    np.testing.assert_array_equal(two.runs_expect, one.runs_expect)
    np.testing.assert_allclose(two.average_expect, one.average_expect, rtol=1e-12, atol=1e-13)
    # std = sqrt(|<e^2> - <e>^2|): where it vanishes, rounding of the two sums enters as its square root
    np.testing.assert_allclose(two.std_expect, one.std_expect, rtol=1e-9, atol=1e-7)
