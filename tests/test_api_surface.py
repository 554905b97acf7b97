"""The drop-in classes expose the reference's public methods with the same positional
parameter names (tests/golden/reference_api.json, recorded from the unmodified reference)."""
import inspect
import json
import os

import pytest

from tests.helpers import GOLDEN


def classes():
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.networks import RNNAgent, QMixer
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.utils.action_selectors import EpsilonGreedyActionSelector
    return {c.__name__: c for c in (ElectromagneticEnvironment, BasicMAC, RNNAgent, QMixer, QMixLearner,
                                    EpisodeReplayBuffer, EpsilonGreedyActionSelector)}


with open(os.path.join(GOLDEN, "reference_api.json")) as f:
    REF_API = json.load(f)

# reference methods that are deliberately absent, with the reason
WAIVED = {
    ("QMixer", "forward"): "evaluated by macjd_mixer_forward inside the learner; nn.Module.forward is not the hot path",
    ("RNNAgent", "get_q_value_for_action"): "evaluated by macjd_qhead_forward inside the learner",
}


@pytest.mark.parametrize("cls_name", sorted(classes()))
def test_public_methods_match_reference(cls_name):
    cls = classes()[cls_name]
    for method, ref_params in REF_API[cls_name].items():
        if (cls_name, method) in WAIVED:
            continue
        assert hasattr(cls, method), f"{cls_name}.{method} missing"
        mine = [p.name for p in inspect.signature(getattr(cls, method)).parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        assert mine[:len(ref_params)] == ref_params, f"{cls_name}.{method}: {mine} vs reference {ref_params}"


def test_state_dict_keys_match_reference_checkpoints():
    import types
    import numpy as np
    from macjd_b200.core.networks import QMixer, RNNAgent
    args = types.SimpleNamespace(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=128,
                                 actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128)
    g = np.load(os.path.join(GOLDEN, "agent_c1.npz"))
    ref_keys = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    mine = {k: tuple(v.shape) for k, v in RNNAgent(24, args, _lib=object()).state_dict().items()}
    assert mine == ref_keys
    g = np.load(os.path.join(GOLDEN, "mixer_c1.npz"))
    ref_keys = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    assert {k: tuple(v.shape) for k, v in QMixer(args).state_dict().items()} == ref_keys


def test_seeded_init_equals_reference_init():
    """Same layer creation order as the reference -> same weights under the same seed."""
    import types
    import numpy as np
    import torch
    from macjd_b200.core.networks import RNNAgent
    args = types.SimpleNamespace(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=128,
                                 actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128)
    g = np.load(os.path.join(GOLDEN, "agent_c1.npz"))
    torch.manual_seed(42)
    agent = RNNAgent(24, args, _lib=object())
    for k, v in agent.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), g["sd." + k], err_msg=k)
