"""Runners + training loop on the CPU through the host-emulation build (tests/emul)."""
import pytest

from tests import runner_checks as RC
from tests.helpers import emul_lib


def test_batched_rollout_vs_oracles():
    RC.check_batched_rollout_against_oracles("cpu", emul_lib())


def test_training_loop_smoke():
    RC.check_training_loop_smoke("cpu", emul_lib())


def test_reference_protocol_runner():
    RC.check_reference_protocol_runner("cpu", emul_lib())


def test_host_buffer_api():
    RC.check_host_buffer_api("cpu", emul_lib())


@pytest.mark.parametrize("group_envs", ["0", "64", "100"])
def test_fused_host_step(monkeypatch, group_envs):
    # 0: never grouped; 64: 150 envs -> two groups, envs [0, 128) and [128, 150); 100: batch too small, one group
    monkeypatch.setenv("MACJD_HOST_GROUP_ENVS", group_envs)
    RC.check_fused_host_step("cpu", emul_lib(), n_envs=150)


def test_main_loop(tmp_path):
    RC.check_main_loop("cpu", emul_lib(), tmp_path)


def test_fast_path_chains_hidden_state():
    RC.check_fast_path_chains_hidden_state("cpu", emul_lib())


def test_rollout_step_call_equals_two_calls():
    """macjd_rollout_step without the tensor-core kernel (host emulation): the two kernels behind one call."""
    RC.check_fused_rollout_step("cpu", emul_lib(), n_envs=9, kind="c2", path=1)
