"""EVOPF env step against the UNMODIFIED reference on edge cases its own goldens do not reach: battery power beyond the
state-of-charge dependent limits (clipped, Battery.step evopf.py:74-102), charging / discharging efficiencies, states
of charge at the 0.1 / 0.8 limits, generation cost with out-of-range generation (obj_fn :509-518).  The step code the
kernel runs is compiled by g++ (tests/hostsim); the reference runs beside it in the build container."""
import numpy as np
import pytest
import torch

from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_step_soc_and_reward_follow_the_reference_on_clipped_battery_power():
    from oracle.ref_loader import make_cpu_env
    from rpo_b200 import synth
    from tests import hostsim_ops as hs
    ref = make_cpu_env("evopf")
    rng = np.random.default_rng(5)
    N = 60
    st = synth.evopf_states(N, seed=9).numpy().astype(np.float64)
    soc = rng.uniform(0.1, 0.8, size=(N, 5))
    soc[:6] = [[0.1, 0.8, 0.1000001, 0.7999999, 0.45], [0.8] * 5, [0.1] * 5, [0.3, 0.62, 0.58, 0.12, 0.78],
               [0.6, 0.6, 0.6, 0.6, 0.6], [0.28, 0.3, 0.32, 0.1, 0.8]]
    st[:, 28:33] = soc
    lo, hi = ref.box_constraint
    y = (lo + (hi - lo) * rng.uniform(-0.1, 1.1, size=(N, 43))).astype(np.float32)
    y[:, 38:] = rng.uniform(-0.35, 0.35, size=(N, 5)).astype(np.float32)      # battery power well beyond +-0.2 / 0.9
    y[0, 38:] = [0.5, 0.5, -0.5, -0.5, 0.0]
    exo = np.zeros((N, 52), np.float32)
    ns, rew, eqv, iqv = hs.evopf_step(torch.from_numpy(st).float(), torch.from_numpy(y), torch.from_numpy(exo))
    for i in range(N):
        ref.reset()
        ref.state = st[i].copy()
        ref.evs.state = st[i, 28:].copy()
        ns_r, r_r, _, info = ref.step(y[i].copy())                          # float32 action, as the agents pass it
        assert np.abs(ns[i, 28:33].numpy() - ns_r[28:33]).max() < 1e-6, i          # state of charge after the clipped step
        assert abs(float(rew[i]) - float(r_r)) <= 2e-5 * max(1.0, abs(float(r_r))), i
        np.testing.assert_allclose(iqv[i].numpy(), np.ravel(info["ineq_viol"]), rtol=0, atol=1e-6)
        np.testing.assert_allclose(eqv[i].numpy(), np.ravel(info["eq_viol"]), rtol=0, atol=3e-5)
    assert float(ns[:, 28:33].min()) >= 0.1 - 1e-6 and float(ns[:, 28:33].max()) <= 0.8 + 1e-6
