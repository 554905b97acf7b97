"""Runners + training loop on the CPU through the host-emulation build (tests/emul)."""
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
