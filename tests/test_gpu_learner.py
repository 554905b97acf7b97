"""GPU parity tests of the learner, mixer and replay kernels through the C ABI."""
import types

import numpy as np
import pytest
import torch

from tests import learner_checks as LC

pytestmark = pytest.mark.gpu


def lib():
    from macjd_b200 import _native as N
    return N.get_lib()


@pytest.mark.parametrize("name", ["c1", "small", "c3", "c4"])
def test_mixer_forward_backward(name):
    LC.check_mixer_against_golden(name, "cuda", lib())


@pytest.mark.parametrize("path", [1, 0])
@pytest.mark.parametrize("name", ["c1", "small_fastlr", "c3", "c4"])
def test_learner_train_steps(name, path):
    """path 0 = what bench.py runs (tcgen05 pair kernel, split unrolls on side streams)."""
    LC.check_learner_against_golden(name, "cuda", lib(), path=path)


def test_replay_store_sample():
    LC.check_replay_against_golden("cuda", lib())


def test_replay_rollout_roundtrip():
    LC.check_rollout_store_roundtrip("cuda", lib())


def test_learner_full_size_vs_oracle():
    """Default training shape (B=32, T=100, H=128, E=64): one step against the eager oracle
    (loss / grad-norm / Q_tot means), determinism of repeated identical steps."""
    from oracle import agent_oracle as AO
    from tests.agent_checks import random_agent
    from macjd_b200.core.qmix import QMixLearner
    args = types.SimpleNamespace(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=128,
                                 actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0,
                                 epsilon_finish=0.05, epsilon_anneal_time=1000, gamma=0.99, lr=5e-6, grad_norm_clip=1.0,
                                 target_update_interval=200, use_cuda=True, device="cuda", seed=0)
    B, T, Nn, A, S, H = 32, 100, 2, 5, 24, 128
    rng = np.random.default_rng(11)
    batch = {"state": rng.standard_normal((B, T + 1, S)).astype(np.float32),
             "obs": rng.standard_normal((B, T + 1, Nn, S)).astype(np.float32),
             "actions_discrete": rng.integers(0, A, size=(B, T, Nn, 1)).astype(np.int32),
             "actions_continuous": rng.random((B, T, Nn, 1)).astype(np.float32),
             "avail_actions": np.ones((B, T + 1, Nn, A), dtype=np.int64),
             "reward": rng.standard_normal((B, T, 1)).astype(np.float32),
             "terminated": np.zeros((B, T, 1), dtype=bool), "filled": np.ones((B, T, 1), dtype=bool),
             "hidden_state": (rng.standard_normal((B, T + 1, Nn, H)) * 0.5).astype(np.float32), "max_seq_len": T}
    stats = []
    for rep in range(2):
        torch.manual_seed(42)
        from macjd_b200.core.mac import BasicMAC
        mac = BasicMAC(S, args)
        L = QMixLearner(mac, args)
        if rep == 0:
            agent_sd = {k: v.detach().cpu().clone() for k, v in mac.agent.state_dict().items()}
            mixer_sd = {k: v.detach().cpu().clone() for k, v in L.eval_qmix_net.state_dict().items()}
        stats.append(L.train(batch, {}))
    assert stats[0] == stats[1], "identical steps must be bit-reproducible"
    ora = AO.LearnerOracle(agent_sd, mixer_sd, Nn, 64, 0.99, 5e-6, 1.0, 200)
    ref, _, _ = ora.train({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in batch.items()})
    for k in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
        np.testing.assert_allclose(stats[0][k], ref[k], rtol=2e-4, atol=1e-5, err_msg=k)


def test_learner_stress_dims_vs_oracle():
    """BASELINE config 4 dims (H = 256, E = 128) at a small batch: one step against the eager oracle."""
    from oracle import agent_oracle as AO
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    args = types.SimpleNamespace(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=256,
                                 actor_hidden_dim=128, mixing_embed_dim=128, hyper_hidden_dim=128, epsilon_start=1.0,
                                 epsilon_finish=0.05, epsilon_anneal_time=1000, gamma=0.99, lr=1e-4, grad_norm_clip=1.0,
                                 target_update_interval=200, use_cuda=True, device="cuda", seed=0)
    B, T, Nn, A, S, H = 12, 9, 2, 5, 24, 256
    rng = np.random.default_rng(5)
    batch = {"state": rng.standard_normal((B, T + 1, S)).astype(np.float32),
             "obs": rng.standard_normal((B, T + 1, Nn, S)).astype(np.float32),
             "actions_discrete": rng.integers(0, A, size=(B, T, Nn, 1)).astype(np.int32),
             "actions_continuous": rng.random((B, T, Nn, 1)).astype(np.float32),
             "avail_actions": np.ones((B, T + 1, Nn, A), dtype=np.int64),
             "reward": rng.standard_normal((B, T, 1)).astype(np.float32),
             "terminated": np.zeros((B, T, 1), dtype=bool), "filled": np.ones((B, T, 1), dtype=bool),
             "hidden_state": (rng.standard_normal((B, T + 1, Nn, H)) * 0.5).astype(np.float32), "max_seq_len": T}
    torch.manual_seed(1)
    mac = BasicMAC(S, args)
    L = QMixLearner(mac, args)
    agent_sd = {k: v.detach().cpu().clone() for k, v in mac.agent.state_dict().items()}
    mixer_sd = {k: v.detach().cpu().clone() for k, v in L.eval_qmix_net.state_dict().items()}
    stats = L.train(batch, {})
    ora = AO.LearnerOracle(agent_sd, mixer_sd, Nn, 128, 0.99, 1e-4, 1.0, 200)
    ref, _, _ = ora.train({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in batch.items()})
    for k in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
        np.testing.assert_allclose(stats[k], ref[k], rtol=2e-4, atol=1e-5, err_msg=k)
    for k in AO.TRAINED_AGENT_KEYS:
        w0 = agent_sd[k].numpy()
        np.testing.assert_allclose(mac.agent.state_dict()[k].cpu().numpy() - w0, ora.agent[k].numpy() - w0,
                                   rtol=1e-2, atol=1e-4 * 1e-2 + 2.4e-7 * np.abs(w0).max(), err_msg=k)


@pytest.mark.parametrize("O,A,H,AH,tc", [(24, 5, 128, 128, True), (176, 33, 128, 128, True), (24, 5, 256, 128, False), (39, 7, 64, 64, False)])
def test_qhead_repack_equals_full_repack(O, A, H, AH, tc):
    """After the learner's in-place Adam step only fc2_q_head is re-packed (six fields, the last tensor-core chunks and
    the Q-head entries of the constant block, one launch): every packed buffer must equal a full re-pack bit for bit."""
    pk = LC.check_qhead_repack("cuda", None, O=O, A=A, H=H, AH=AH)
    assert (pk.tc_buffer is not None) == tc


@pytest.mark.parametrize("n_agents", [2, 8])
def test_shared_obs_replay(n_agents):
    LC.check_shared_obs_replay("cuda", lib(), n_agents=n_agents)


def test_hidden_bf16_replay():
    LC.check_hidden_bf16_replay("cuda", lib())


def test_hidden_bf16_effect_on_a_train_step():
    """What the opt-in BF16 hidden_state ring costs the learner: one train step on the same episodes from a float32
    ring and from a bfloat16 ring.  Stated bound of the option: loss and Q_tot means within 1 % (the stored states
    carry up to 2^-9 relative rounding error; the double-DQN targets do not read them at all)."""
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    B, T, Nn, A, S, H = 32, 100, 2, 5, 24, 128
    args = types.SimpleNamespace(n_agents=Nn, n_actions=A, state_shape=S, obs_shape=S, rnn_hidden_dim=H, actor_hidden_dim=128,
                                 mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0, epsilon_finish=0.05,
                                 epsilon_anneal_time=1000, gamma=0.99, lr=5e-6, grad_norm_clip=1.0, target_update_interval=200,
                                 use_cuda=True, device="cuda", seed=0, buffer_size=B, episode_limit=T, batch_size=B)
    g = torch.Generator(device="cuda").manual_seed(5)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    traj = {"state": rn(T + 1, B, S), "obs": rn(T + 1, B, Nn, S),
            "actions_discrete": torch.randint(0, A, (T, B, Nn, 1), device="cuda", generator=g, dtype=torch.int32),
            "actions_continuous": torch.rand(T, B, Nn, 1, device="cuda", generator=g),
            "avail_actions": torch.ones(T + 1, B, Nn, A, dtype=torch.uint8, device="cuda"),
            "reward": rn(T, B, 1), "terminated": torch.zeros(T, B, 1, dtype=torch.uint8, device="cuda"),
            "hidden_state": rn(T + 1, B, Nn, H) * 0.5}
    stats = []
    for bf16 in (False, True):
        torch.manual_seed(42)
        mac = BasicMAC(S, args)
        learner = QMixLearner(mac, args)
        buf = EpisodeReplayBuffer(args, device="cuda", hidden_bf16=bf16)
        buf.store_rollout(traj)
        batch = buf.gather(np.arange(B), time_major=True)
        batch["time_major"] = True
        stats.append(learner.train(batch, {}))
    print("\nfloat32 ring:", stats[0], "\nbfloat16 ring:", stats[1])
    for k in ("loss", "eval_qtot_avg", "target_qtot_avg"):
        np.testing.assert_allclose(stats[1][k], stats[0][k], rtol=1e-2, atol=1e-3, err_msg=k)
    assert stats[1]["target_qtot_avg"] == stats[0]["target_qtot_avg"]      # the unrolls start from zeros: no stored state involved


def _sampled_learner(seed=5, n_envs=64, device="cuda:0", same_init_seed=None):
    """A learner + a filled replay ring at the reference dims (the bench's C1 learner shape, fewer envs).  same_init_seed:
    the networks' initial weights come from this seed (data-parallel ranks start equal), the env from `seed`."""
    import bench
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    rl = bench.rl_args(device, n_envs)
    rl.target_update_interval = 3                 # the hard target update falls between replays of the captured step
    torch.manual_seed(seed if same_init_seed is None else same_init_seed)
    env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=device, seed=seed)
    mac = BasicMAC(bench.OBS, rl)
    mac.cuda()
    buf = EpisodeReplayBuffer(rl, device=device)
    runner = BatchedEpisodeRunner(env, mac, buf, rl)
    learner = QMixLearner(mac, rl)
    runner.run()
    return learner, buf


def test_train_sampled_graph_equals_eager_steps(monkeypatch):
    """QMixLearner.train_sampled replays the whole step as one CUDA graph (from its second step of a shape on): same
    index draws, same kernels -- seven steps must leave the same statistics and the same parameters, target networks
    and optimiser state as sample() + train() with the graph switched off (bit for bit: nothing in the step is
    order-dependent), across two hard target updates."""
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MACJD_TRAIN_GRAPH", mode)
        learner, buf = _sampled_learner()
        np.random.seed(11)
        stats = [learner.train_sampled(buf, 16, {})["stats_tensor"] for _ in range(7)]
        torch.cuda.synchronize()
        assert ("_step_graphs" in learner.__dict__ and any(isinstance(g, dict) for g in learner._step_graphs.values())) == (mode == "1")
        out[mode] = (torch.stack(stats).cpu(),
                     {k: v.detach().cpu().clone() for k, v in learner.mac.agent.state_dict().items()},
                     {k: v.detach().cpu().clone() for k, v in learner.eval_qmix_net.state_dict().items()},
                     {k: v.detach().cpu().clone() for k, v in learner.target_qmix_net.state_dict().items()},
                     {k: v.detach().cpu().clone() for k, v in learner.target_mac.agent.state_dict().items()},
                     learner._opt_state["m"].cpu().clone(), learner._opt_state["v"].cpu().clone(),
                     learner.train_step, learner._opt_state["step"], learner.last_target_update_step)
    g, e = out["1"], out["0"]
    assert torch.isfinite(g[0]).all() and g[0].abs().sum() > 0
    assert torch.equal(g[0], e[0]), (g[0], e[0])
    for a, b in zip(g[1:5], e[1:5]):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(g[5], e[5]) and torch.equal(g[6], e[6])
    assert g[7:] == e[7:] == (7, 7, 6)


def test_dp_train_sampled_graph_equals_eager():
    """Two ranks over NCCL (tests/dp_graph_worker.py): the data-parallel train step with its all-reduce captured in the
    step graph equals the eager data-parallel step bit for bit on every rank, and the ranks hold the same networks."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dp_graph_worker.py")],
                       cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DP_GRAPH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
