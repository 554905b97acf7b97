"""Runners + training loop on the GPU."""
import pytest

from tests import runner_checks as RC

pytestmark = pytest.mark.gpu


def test_batched_rollout_vs_oracles():
    RC.check_batched_rollout_against_oracles("cuda", None, n_envs=300)


def test_training_loop_smoke():
    RC.check_training_loop_smoke("cuda", None)


def test_reference_protocol_runner():
    RC.check_reference_protocol_runner("cuda", None)


@pytest.mark.parametrize("n_envs,pinned", [(300, True), (300, False), (6000, True)])
def test_host_buffer_api(n_envs, pinned):
    RC.check_host_buffer_api("cuda", None, n_envs=n_envs, pinned=pinned)


@pytest.mark.parametrize("n_envs,pinned,group_envs", [(300, True, "0"), (300, False, "0"), (6000, True, "0"),
                                                      (6000, True, "2048"), (6000, False, "1000"), (300, True, "100")])
def test_fused_host_step(monkeypatch, n_envs, pinned, group_envs):
    # group_envs > 0: batches of >= 2 x that many envs are stepped as two groups on two streams (opt-in)
    monkeypatch.setenv("MACJD_HOST_GROUP_ENVS", group_envs)
    RC.check_fused_host_step("cuda", None, n_envs=n_envs, pinned=pinned)


def test_episode_graph_equals_stepwise():
    RC.check_episode_graph_equals_stepwise("cuda")


def test_main_loop(tmp_path):
    RC.check_main_loop("cuda", None, tmp_path)


def test_fast_path_chains_hidden_state():
    RC.check_fast_path_chains_hidden_state("cuda", None)


@pytest.mark.parametrize("kind,n_envs,path,fused", [("c2", 300, 0, True), ("c2", 4096, 0, True), ("c2", 33, 1, None),
                                                    ("c3", 100, 0, None), ("c3", 1030, 0, None), ("j3", 200, 0, False)])
def test_fused_rollout_step_equals_two_kernels(monkeypatch, kind, n_envs, path, fused):
    # (the library fuses the 8 x 16 x 4 scenario only when asked to: 50 KB of views per CTA-step; ask)
    monkeypatch.setenv("MACJD_FUSE_MAX_VIEW_BYTES", "1000000")
    RC.check_fused_rollout_step("cuda", None, n_envs=n_envs, kind=kind, path=path, expect_fused=fused)


def test_pipelined_loop_trains_through_the_step_graph():
    """main.run(pipeline=True): the learner's phase is QMixLearner.train_sampled on its own stream -- eager first step,
    captured second, replayed from then on (a hetero scenario: episodes end at different lengths, so several shapes
    occur and the cache limit is exercised) -- and the loop's counters agree with the learner's."""
    import types
    import numpy as np
    from macjd_b200 import main as M
    from macjd_b200.simulation.scenario import hetero_spec
    n_envs, T = 16, 8
    args = RC.rl_args("cuda", episode_limit=T, batch_size=8, buffer_size=32, total_env_steps=6 * n_envs * T,
                      start_training_steps=n_envs * T, train_interval=2, log_interval=10, log_interval_seconds=0,
                      save_model=False, test_interval=0, test_nepisodes=0, device_request="cuda", rnn_hidden_dim=128,
                      actor_hidden_dim=128, target_update_interval=5)
    scalars = []
    writer = types.SimpleNamespace(add_scalar=lambda tag, v, step: scalars.append((tag, float(v), step)), close=lambda: None)
    np.random.seed(4)
    p = M.run(args, spec=hetero_spec(n_envs, seed=3, active=True, episode_limit=T), writer=writer, log=lambda s: None, pipeline=True)
    assert p["total_steps"] == 6 * n_envs * T and p["episodes"] == 6 * n_envs
    # rollouts 3 .. 6 train (the pipelined learner works from the ring as of the previous rollout), n_envs * (T // 2) steps each
    assert p["train_steps"] == 4 * n_envs * (T // 2)
    graphs = p["learner"].__dict__.get("_step_graphs", {})
    assert any(isinstance(g, dict) for g in graphs.values()) and len(graphs) <= p["learner"].MAX_STEP_GRAPHS
    assert p["learner"].train_step == p["train_steps"] and p["learner"]._opt_state["step"] == p["train_steps"]
    assert p["learner"].last_target_update_step == (p["train_steps"] // 5) * 5
    losses = [v for t, v, _ in scalars if t == "Loss/train_avg"]
    assert losses and all(np.isfinite(losses))
