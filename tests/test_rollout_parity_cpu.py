"""f-2 / f-4 parity on CPU (host simulator ops): device replay == reference ReplayBuffer, exogenous generators ==
reference loaders, batched eval == the UNMODIFIED reference's eval() on identical data.  Bodies: tests/rollout_cases.py."""
import pytest

from oracle.ref_loader import reference_available
from tests import rollout_cases

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture()
def hs(monkeypatch):
    from tests import hostsim_ops
    hostsim_ops.install(monkeypatch)


def test_exo_sampler_matches_reference_loaders(hs):
    rollout_cases.check_exo_sampler_matches_reference_loaders("cpu")


def test_device_replay_matches_reference_buffer():
    rollout_cases.check_device_replay_matches_reference_buffer("cpu")


def test_batched_eval_matches_reference_eval(hs, capsys):
    stats, ref = rollout_cases.check_batched_eval_matches_reference_eval("cpu")
    with capsys.disabled():
        print("\n[f-4] evaluate_batched vs reference eval():", ["%.5g / %.5g" % (a, b) for a, b in zip(stats, ref)])
