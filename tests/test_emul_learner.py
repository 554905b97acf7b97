"""Learner, mixer and replay kernels checked on the CPU through the host-emulation build
(tests/emul) against the reference goldens.  GPU twin: tests/test_gpu_learner.py."""
import pytest

from tests import learner_checks as LC
from tests.helpers import emul_lib


@pytest.mark.parametrize("name", ["c1", "small", "c3", "c4"])
def test_mixer_forward_backward(name):
    LC.check_mixer_against_golden(name, "cpu", emul_lib())


@pytest.mark.parametrize("name", ["c1", "small_fastlr", "c3", "c4"])
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


def test_fast_sampling_draws():
    """fast_sampling=True (utils/replay_buffer.py header): distinct in-range indices, reproducible under np.random.seed,
    every slot equally likely; the default stays the reference's draw (np.random.choice, replay_buffer.py:178)."""
    import types
    import numpy as np
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from tests.helpers import emul_lib
    args = types.SimpleNamespace(buffer_size=50, episode_limit=3, n_actions=5, n_agents=2, state_shape=4, obs_shape=4,
                                 rnn_hidden_dim=8, use_cuda=False, device="cpu")
    def draws(fast, seed, n=400):
        buf = EpisodeReplayBuffer(args, device="cpu", _lib=emul_lib(), fast_sampling=fast)
        buf.current_size = 40
        np.random.seed(seed)
        return np.stack([buf._draw_indices(8) for _ in range(n)])
    a, b, c = draws(True, 3), draws(True, 3), draws(True, 4)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= 0 and a.max() < 40 and all(len(set(r)) == 8 for r in a)
    counts = np.bincount(a.reshape(-1), minlength=40)                # 3 200 draws over 40 slots: 80 expected each
    assert counts.min() > 45 and counts.max() < 120, counts
    np.random.seed(3)
    want = np.stack([np.random.choice(40, 8, replace=False) for _ in range(5)])
    assert np.array_equal(draws(False, 3, 5), want)
    buf = EpisodeReplayBuffer(args, device="cpu", _lib=emul_lib(), fast_sampling=True)
    buf.current_size = 5
    assert sorted(buf._draw_indices(8)) == [0, 1, 2, 3, 4]           # fewer episodes than asked for: all of them (with the warning)
