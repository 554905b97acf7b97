"""Learner, mixer and replay kernels checked on the CPU through the host-emulation build
(tests/emul) against the reference goldens.  GPU twin: tests/test_gpu_learner.py."""
import pytest

from tests import learner_checks as LC
from tests.helpers import emul_lib


@pytest.mark.parametrize("name", ["c1", "small"])
def test_mixer_forward_backward(name):
    LC.check_mixer_against_golden(name, "cpu", emul_lib())


@pytest.mark.parametrize("name", ["c1", "small_fastlr"])
def test_learner_train_steps(name):
    LC.check_learner_against_golden(name, "cpu", emul_lib())


def test_replay_store_sample():
    LC.check_replay_against_golden("cpu", emul_lib())


def test_replay_rollout_roundtrip():
    LC.check_rollout_store_roundtrip("cpu", emul_lib())


@pytest.mark.parametrize("n_agents", [2, 8])
def test_shared_obs_replay(n_agents):
    LC.check_shared_obs_replay("cpu", emul_lib(), n_agents=n_agents)


def test_hidden_bf16_replay():
    LC.check_hidden_bf16_replay("cpu", emul_lib())


@pytest.mark.parametrize("O,A,H,AH", [(24, 5, 128, 128), (176, 33, 128, 128), (39, 7, 64, 64)])
def test_qhead_repack_equals_full_repack(O, A, H, AH):
    """macjd_qhead_repack (one launch after the optimiser step) against the full re-pack and the host fallback."""
    LC.check_qhead_repack("cpu", emul_lib(), O=O, A=A, H=H, AH=AH)
