"""GPU twin of tests/test_rollout_parity_cpu.py: the same bodies with the CUDA kernels and device-resident buffers."""
import pytest

from oracle.ref_loader import reference_available
from tests import rollout_cases

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="no reference tree: run oracle/make_ref_copy.sh")]


def test_exo_sampler_matches_reference_loaders_on_device():
    rollout_cases.check_exo_sampler_matches_reference_loaders("cuda:0")


def test_device_replay_matches_reference_buffer_on_device():
    rollout_cases.check_device_replay_matches_reference_buffer("cuda:0")


def test_batched_eval_matches_reference_eval_on_device(capsys):
    stats, ref = rollout_cases.check_batched_eval_matches_reference_eval("cuda:0")
    with capsys.disabled():
        print("\n[f-4, cuda] evaluate_batched vs reference eval():", ["%.5g / %.5g" % (a, b) for a, b in zip(stats, ref)])
