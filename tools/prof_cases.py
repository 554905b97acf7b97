"""Driver for ncu captures (round 2): runs ONE configuration a few times so that `ncu -k regex:... -s ... -c ...`
can pick the launches.    python tools/prof_cases.py {learner_c1|learner_c4|agent_c3|agent_c2|env_c3|env_c3_64k|env_c2_1m|rollout_c2}"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B   # noqa: E402  (rl_args, dims)

which = sys.argv[1]
dev = "cuda:0"
if which.startswith("learner"):
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    c4 = which == "learner_c4"
    Bn, H, E, T = (1024, 256, 128, 100) if c4 else (32, 128, 64, 100)
    rl = B.rl_args(dev, Bn, rnn_hidden_dim=H, mixing_embed_dim=E, batch_size=Bn, buffer_size=Bn)
    torch.manual_seed(44)
    mac = BasicMAC(B.OBS, rl)
    mac.cuda()
    buf = EpisodeReplayBuffer(rl, device=dev)
    g = torch.Generator(device=dev).manual_seed(11)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    buf.store_rollout({
        "state": rn(T + 1, Bn, B.OBS), "obs": rn(T + 1, Bn, 2, B.OBS),
        "actions_discrete": torch.randint(0, 5, (T, Bn, 2, 1), device=dev, generator=g, dtype=torch.int32),
        "actions_continuous": torch.rand(T, Bn, 2, 1, device=dev, generator=g),
        "avail_actions": torch.ones(T + 1, Bn, 2, 5, dtype=torch.uint8, device=dev),
        "reward": rn(T, Bn, 1), "terminated": torch.zeros(T, Bn, 1, dtype=torch.uint8, device=dev),
        "hidden_state": rn(T + 1, Bn, 2, H) * 0.5})
    learner = QMixLearner(mac, rl)
    np.random.seed(0)
    for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
        torch.cuda.nvtx.range_push(f"train{i}")
        learner.train(buf.sample(Bn, time_major=True), {})
        torch.cuda.nvtx.range_pop()
elif which.startswith("agent"):
    from tests.agent_checks import random_agent
    O, A, M = (176, 33, 65536) if which == "agent_c3" else (24, 5, 8192)
    mac, _ = random_agent(0, O, A, 128, 128, 2, dev)
    obs = torch.randn(1, M, O, device=dev)
    h = torch.randn(M, 128, device=dev) * 0.3
    for _ in range(3):
        mac.agent.run(obs, h, select=True, test_mode=False, epsilon=0.3, path=3)
elif which.startswith("env"):
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec, scaled_spec
    if which in ("env_c2_1m", "env_c2"):
        n, J, R, K = (1 << 20) if which == "env_c2_1m" else 4096, 2, 2, 1
        spec = default_spec(n)
    else:
        n, J, R, K = (65536 if which == "env_c3_64k" else 8192), 8, 16, 4
        spec = scaled_spec(n, n_jammers=J, n_radars=R, n_targets=K, seed=1)
    env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=spec, device=dev)
    act_d = torch.randint(0, 2 * R + 1, (n, J), dtype=torch.int32, device=dev)
    act_p = torch.rand(n, J, device=dev)
    for _ in range(3):
        env.step_device(act_d, act_p)
elif which == "rollout_c2":
    # the bench's timed launch: macjd_rollout_steps over one episode (100 timesteps, env steps inline), 4 096 envs
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    rl = B.rl_args(dev, B.N_ENVS)
    torch.manual_seed(42)
    env = ElectromagneticEnvironment(rl, spec=default_spec(B.N_ENVS), device=dev, seed=1000)
    mac = BasicMAC(B.OBS, rl)
    mac.cuda()
    runner = BatchedEpisodeRunner(env, mac, None, rl)
    for _ in range(3):
        runner.reset()
        runner.rollout(0, 100)
torch.cuda.synchronize()
print("done", which)
