#!/usr/bin/env python
"""Benchmark of the MA-CJD hot path on B200 (contract: see the build prompt / DESIGN.md).

Workload (BASELINE.json configs[1]): the reference's default scenario (2 jammers x 2 radars
x 1 target, obs/state 24, 5 actions, GRU 128) batched to 4096 parallel envs per GPU.
A *step* is one batched environment timestep: the fused agent-act kernel over
n_envs x n_agents rows followed by the fused env-step kernel over n_envs episodes, both
writing into the rollout trajectory in HBM.  `value` = env-agent steps/s with everything
resident in HBM (CUDA events per step, L2 flushed between steps); `e2e` = the same step
through the reference-facing host API (numpy obs / avail in pinned host memory ->
select_actions -> actions back to the host -> env.step -> obs / reward / terminated back).
The learner (QMix train samples/s, B=32 x T=100) and the CPU baseline (the oracle port of
the reference path on the host cores) ride along in the same JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS, N_AGENTS, OBS, N_ACTIONS, HID = 4096, 2, 24, 5, 128
ENV_BYTES_PER_STEP = 569.0          # algorithmic bytes per env-step (C1/C2)
AGENT_FLOP_PER_ROW = 279_000        # minimal-work FLOPs per agent-step (C1/C2) -- see flop_per_row()
LEARNER_B, LEARNER_T = 32, 100


def rl_args(device, n_envs):
    a = types.SimpleNamespace(
        n_agents=N_AGENTS, n_actions=N_ACTIONS, state_shape=OBS, obs_shape=OBS, rnn_hidden_dim=HID,
        actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0, epsilon_finish=0.05,
        epsilon_anneal_time=100000, gamma=0.99, lr=5e-6, grad_norm_clip=1.0, target_update_interval=200,
        use_cuda=True, device=device, batch_size=LEARNER_B, buffer_size=2 * n_envs, episode_limit=100, seed=42,
        data_parallel=True, agent_kernel_path=0)     # 0: tcgen05 3xTF32 agent kernel where the dims allow
    a.env_info = {"state_shape": OBS, "obs_shape": OBS, "n_actions": N_ACTIONS, "n_agents": N_AGENTS, "episode_limit": 100}
    return a


def flop_per_row(O=OBS, A=N_ACTIONS, H=HID, AH=128):
    """Algorithmic FLOPs of one agent step with the shared Q-head product."""
    return 2 * (O * AH + AH * AH + AH * A) + 2 * O * H + 12 * H * H + 2 * H * H + A * H * 5


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline_run(steps, warmup, n_envs=N_ENVS):
    """The oracle port of the reference path on the host cores: NumPy float64 env step
    (vectorised over envs) + eager-PyTorch agent act (all torch threads).  One step = the
    bench workload's step (n_envs envs x n_agents agents)."""
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the baseline is meant to use every host core it can
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    from macjd_b200.simulation.scenario import default_spec
    from oracle.env_oracle import EnvOracle
    from oracle import agent_oracle as AO
    torch.manual_seed(42)
    ora = EnvOracle(default_spec(n_envs))
    import torch.nn as nn
    # random-init weights of the reference architecture (state_dict layout of RNNAgent)
    mods = {"actor.0": nn.Linear(OBS, 128), "actor.2": nn.Linear(128, 128), "actor.4": nn.Linear(128, N_ACTIONS),
            "fc1": nn.Linear(OBS, HID), "fc2_q_head.0": nn.Linear(HID + N_ACTIONS + 1, HID), "fc2_q_head.2": nn.Linear(HID, 1)}
    sd = {}
    for k, m in mods.items():
        sd[k + ".weight"], sd[k + ".bias"] = m.weight.detach(), m.bias.detach()
    gru = nn.GRUCell(HID, HID)
    sd.update({"rnn.weight_ih": gru.weight_ih.detach(), "rnn.weight_hh": gru.weight_hh.detach(),
               "rnn.bias_ih": gru.bias_ih.detach(), "rnn.bias_hh": gru.bias_hh.detach()})
    rng = np.random.default_rng(7)
    h = torch.zeros(n_envs * N_AGENTS, HID)
    avail = torch.ones(n_envs, N_AGENTS, N_ACTIONS, dtype=torch.long)
    ora.reset()

    def one_step(h):
        obs = torch.from_numpy(ora.get_obs())
        u = torch.from_numpy(rng.random((n_envs, N_AGENTS)).astype(np.float32))
        ra = torch.from_numpy(rng.integers(0, N_ACTIONS, size=(n_envs, N_AGENTS)))
        with torch.no_grad():
            a, p, h, _, _ = AO.select_actions(sd, obs, avail, h, 0.5, False, u, ra)
        ora.step(a.view(n_envs, N_AGENTS).numpy(), p.view(n_envs, N_AGENTS).numpy(), rng.random((n_envs, 4)))
        return h

    for _ in range(warmup):
        h = one_step(h)
    t0 = time.perf_counter()
    for _ in range(steps):
        h = one_step(h)
    dt = time.perf_counter() - t0
    return {"value": n_envs * N_AGENTS * steps / dt, "unit": "env-agent steps/s", "cores": int(torch.get_num_threads()),
            "kind": "port", "ms_per_step": dt / steps * 1e3,
            "sample": f"{steps} batched steps of the same workload ({n_envs} envs x {N_AGENTS} agents): "
                      f"NumPy f64 env oracle (1 thread) + eager-PyTorch agent oracle ({torch.get_num_threads()} threads); "
                      f"host has {os.cpu_count()} cores"}


def cpu_learner_run(mac, learner, buf, steps):
    """The oracle port of QMixLearner.train (eager PyTorch autograd, all host threads) on a batch drawn from
    the same replay ring: the CPU figure beside the learner's train-samples/s."""
    import torch
    from oracle import agent_oracle as AO
    agent_sd = {k: v.detach().cpu().clone() for k, v in mac.agent.state_dict().items()}
    mixer_sd = {k: v.detach().cpu().clone() for k, v in learner.eval_qmix_net.state_dict().items()}
    ora = AO.LearnerOracle(agent_sd, mixer_sd, N_AGENTS, 64, 0.99, 5e-6, 1.0, 200)
    batch = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in buf.sample(LEARNER_B).items()}
    ora.train(batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.train(batch)
    dt = (time.perf_counter() - t0) / steps
    return {"train_episodes_per_sec": LEARNER_B / dt, "train_transitions_per_sec": LEARNER_B * (LEARNER_T - 1) / dt,
            "ms_per_train_step": dt * 1e3, "cores": int(torch.get_num_threads()), "kind": "port",
            "sample": f"{steps} train steps of B = {LEARNER_B} x T = {LEARNER_T} through oracle/agent_oracle.py LearnerOracle"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 1)
    steps = min(steps, 200)
    base = cpu_baseline_run(steps, warmup)
    line = {"impl": "reference", "metric": "env_agent_steps_per_sec", "value": base["value"], "unit": "env-agent steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 env + f32 nets",
            "data": "synthetic",
            "config": {"workload": "default scenario x 4096 envs, fused env step + agent act (BASELINE.json configs[1])",
                       "n_envs": N_ENVS, "n_agents": N_AGENTS},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "env-agent steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- the B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="macjd_b200")
    ap.add_argument("--n-envs", type=int, default=N_ENVS, help="envs per GPU")
    ap.add_argument("--learner-steps", type=int, default=10)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    real_stdout = sys.stdout
    sys.stdout = sys.stderr                     # library chatter must not pollute the one JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = f"cuda:{local}"
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner

    n_envs, K, W = args.n_envs, args.steps, max(args.warmup, 3)
    rl = rl_args(device, n_envs)
    torch.manual_seed(42)                       # identical weights on every rank
    env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=device, seed=1000 + rank)
    mac = BasicMAC(OBS, rl)
    mac.cuda()
    buf = EpisodeReplayBuffer(rl, device=device)
    runner = BatchedEpisodeRunner(env, mac, buf, rl)
    learner = QMixLearner(mac, rl)
    M = n_envs * N_AGENTS
    T = runner.episode_limit
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn, n, pre=None):
        """Sum of per-step CUDA-event times; L2 flushed (outside the timed region) before each."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            if pre is not None:
                pre(i)
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) * 1e-3

    # ---- warm-up: one full episode (also fills the replay ring), then W steps
    runner.run()
    runner.reset()
    t_cur = [0]

    def rollout_step(i):
        if t_cur[0] == T:
            runner.reset()
            t_cur[0] = 0
        runner.step(t_cur[0])
        t_cur[0] += 1

    def pre_step(i):
        if t_cur[0] == T:                       # episode boundary handled outside the timed region
            runner.reset()
            t_cur[0] = 0

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        rollout_step(i)
    barrier()
    dt = timed_steps(rollout_step, K, pre=pre_step)
    barrier()
    dt = reduce_max(dt)
    value = world * M * K / dt

    # ---- roofline of the dominant kernel (agent_forward) and of the env-step kernel, each alone
    # the two launches of one timestep exactly as the timed rollout issues them (cached C structs: the
    # host side is one ctypes call, so the CUDA events bracket the kernel and not Python)
    import copy as _copy
    if getattr(runner, "_structs_for", None) != (mac.hidden_states.data_ptr(), mac.agent.path):
        runner._build_step_structs()
    lib, ctx = mac.agent.lib(), mac.agent._ctx()
    # timestep 1: a typical step (timestep 0 starts from zeros and does not read a recurrent state)
    aio, eio = runner._agent_io[1], runner._env_io[1]
    aio.epsilon, aio.rng_step, aio.test_mode = 0.3, 1, 0
    aio_simt = type(aio).from_buffer_copy(aio)
    aio_simt.path = 1
    eio_alone = type(eio).from_buffer_copy(eio)
    eio_alone.flags = 0                          # timed on its own: not behind an agent kernel (include/macjd.h)
    wts = mac.agent.packed().cstruct()
    kn = 30
    dt_agent = timed_steps(lambda i: lib.call("macjd_agent_forward", ctx, wts, aio), kn) / kn
    dt_env = timed_steps(lambda i: lib.call("macjd_env_step", ctx, env._ctab, eio_alone), kn) / kn
    clocks = None
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    fpr = flop_per_row()
    # the same launch on the FP32 SIMT kernel, for reference
    dt_simt = timed_steps(lambda i: lib.call("macjd_agent_forward", ctx, wts, aio_simt), kn) / kn
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    tf32_peak = bf16_peak / 2.0
    tr = traffic.get("agent_forward_tc2_kernel") or traffic.get("agent_forward_tc_kernel") or {}
    roofline = {"kernel": "agent_forward_tc2_kernel", "bound": "tensor", "achieved": M * fpr / dt_agent / 1e12,
                "peak": tf32_peak, "unit": "TFLOP/s", "frac": M * fpr / dt_agent / 1e12 / tf32_peak,
                "traffic": tr.get("dram_bytes_per_launch"), "flop_per_agent_step": fpr, "us_per_launch": dt_agent * 1e6,
                "executed_tensor_tflops": 3 * M * fpr / dt_agent / 1e12,
                "peak_source": ("TF32 dense = MEASURED_PEAKS.json bf16_tflops / 2 (of measured)" if "bf16_tflops" in peaks
                                else "TF32 dense = fallback 1590 / 2 (of fallback)"),
                "note": "achieved counts algorithmic FLOPs once; the 3xTF32 split executes 3x that on the tensor pipe. "
                        "CTA pairs issue M=128,N=128,K=8 cta_group::2 MMAs; inside the MMA phases they run at ~53 cycles each against the "
                        "pipe's 36.8 (refill latency of the 2-stage weight ring, tools/tc_mma_rate.py), and the MMA phases alternate "
                        "with epilogue phases along the layer dependency chain: latency-bound at 64 rows per SM, not at the roofline",
                "simt_kernel": {"kernel": "agent_forward_kernel<256>", "bound": "fp32", "us_per_launch": dt_simt * 1e6,
                                "achieved": M * fpr / dt_simt / 1e12, "peak": fp32_peak,
                                "frac": M * fpr / dt_simt / 1e12 / fp32_peak,
                                "peak_source": f"derived: {n_sm} SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz"}}
    roofline_env = {"kernel": "env_step_kernel", "bound": "hbm", "achieved": n_envs * ENV_BYTES_PER_STEP / dt_env / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": n_envs * ENV_BYTES_PER_STEP / dt_env / 1e9 / hbm_peak,
                    "traffic": (traffic.get("env_step_kernel") or {}).get("dram_bytes_per_launch_at_bench_size"),
                    "bytes_per_env_step": ENV_BYTES_PER_STEP,
                    "us_per_launch": dt_env * 1e6,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}

    # ---- e2e: the reference-facing host API (host buffers in, host buffers out) per step: what the
    # reference's runner loop does -- mac.select_actions(obs) -> env.step(actions) -- through the two
    # host-buffer C-ABI calls (copies inside the call, stream drained on return)
    hb = env.host_buffers()                                    # pinned: act_d, act_p, reward, terminated, obs
    avail_h = torch.ones(n_envs, N_AGENTS, N_ACTIONS, dtype=torch.uint8).pin_memory()
    env.reset()
    hb["obs"].copy_(env.get_obs())
    mac.init_hidden(n_envs)
    t_env = [0]

    def host_step_two_calls():
        # H2D obs + avail, fused agent step, D2H actions + power (written straight into the env's action buffers)
        mac.select_actions_host(hb["obs"], avail_h, t_env[0], actions_out=hb["act_d"], power_out=hb["act_p"])
        # H2D actions + power, fused env step, D2H reward + terminated + next obs
        env.step_host(hb)

    def host_step_fused():
        # the same iteration as ONE call and one stream drain (BatchedEpisodeRunner.step_host ->
        # macjd_rollout_step_host): H2D obs + avail, agent step, env step on the device-resident actions,
        # D2H actions + power + reward + terminated + next obs
        runner.t_env = t_env[0]
        runner.step_host(hb["obs"], avail_h, hb)

    def time_host(step_fn):
        env.reset()
        hb["obs"].copy_(env.get_obs())
        mac.init_hidden(n_envs)
        t_env[0] = 0

        def one():
            step_fn()
            t_env[0] += 1
            if t_env[0] % T == 0:
                env.reset()
                hb["obs"].copy_(env.get_obs())
                mac.init_hidden(n_envs)
        for _ in range(W):
            one()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            one()
        torch.cuda.synchronize()
        return reduce_max(time.perf_counter() - t0)

    t_env_saved = runner.t_env
    dt_two = time_host(host_step_two_calls)
    dt_e2e = time_host(host_step_fused)
    runner.t_env = t_env_saved
    obs_b, act_b = hb["obs"].numel() * 4, hb["act_d"].numel() * 4 + hb["act_p"].numel() * 4
    out_b = hb["reward"].numel() * 4 + hb["terminated"].numel()
    e2e = {"value": world * M * K / dt_e2e, "unit": "env-agent steps/s", "h2d_bytes_per_step": int(obs_b + avail_h.numel()),
           "d2h_bytes_per_step": int(act_b + obs_b + out_b), "ms_per_step": dt_e2e / K * 1e3,
           "api": "BatchedEpisodeRunner.step_host (C-ABI macjd_rollout_step_host): pinned host observations / masks in; "
                  "actions, power, reward, terminated and next observations out to pinned host buffers; one stream "
                  "drain per step",
           "two_calls": {"value": world * M * K / dt_two, "ms_per_step": dt_two / K * 1e3,
                         "h2d_bytes_per_step": int(obs_b + avail_h.numel() + act_b), "d2h_bytes_per_step": int(act_b + obs_b + out_b),
                         "api": "BasicMAC.select_actions_host + ElectromagneticEnvironment.step_host (macjd_agent_act_host + "
                                "macjd_env_step_host): the reference's two calls, two stream drains per step"}}

    # ---- learner: sample + train (B=32 episodes x T=100), all-reduce of the gradient bucket when N > 1
    np.random.seed(1 + rank)
    for _ in range(3):
        learner.train(buf.sample(LEARNER_B, time_major=True), {})
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.learner_steps):
        stats = learner.train(buf.sample(LEARNER_B, time_major=True), {})
    ev1.record()
    barrier()
    dt_l = reduce_max(ev0.elapsed_time(ev1) * 1e-3) / args.learner_steps
    clocks = sampler.stop()
    learner_rec = {"train_episodes_per_sec": world * LEARNER_B / dt_l,
                   "train_transitions_per_sec": world * LEARNER_B * (LEARNER_T - 1) / dt_l, "ms_per_train_step": dt_l * 1e3,
                   "batch_episodes_per_gpu": LEARNER_B, "episode_len": LEARNER_T, "last_loss": stats["loss"],
                   "includes": "replay sample (gather kernel) + 2 unrolls + mixers + TD + backward + clip/Adam"
                               + (" + NCCL all-reduce" if world > 1 else "")}

    # 2 N F_agent (two unrolls) + 3 N F_qhead (q_taken fwd + bwd) + 4 F_mixer per transition
    f_agent, f_qhead, f_mixer = 415_744, 34_560, 68_288
    flop_tr = 2 * N_AGENTS * f_agent + 3 * N_AGENTS * f_qhead + 4 * f_mixer
    l_tflops = LEARNER_B * (LEARNER_T - 1) * flop_tr / dt_l / 1e12
    learner_rec["roofline"] = {
        "bound": "tensor", "achieved": l_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": l_tflops / tf32_peak,
        "flop_per_transition": flop_tr,
        "note": "B = 32 episodes is 64 agent rows: both 100-step unrolls run on one CTA pair each, so the step is bound by "
                "the serial depth (2 T dependent GRU steps, SURVEY 8d), not by throughput"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_run(args.cpu_steps, 3)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["learner"] = cpu_learner_run(mac, learner, buf, 2)

    if rank == 0:
        line = {"metric": "env_agent_steps_per_sec", "value": value, "unit": "env-agent steps/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 env physics + 3xTF32 tensor-core GEMMs (f32-level) + f32 epilogues",
                "data": "synthetic",
                "config": {"workload": "default scenario x 4096 envs per GPU, fused agent act + fused env step "
                                       "(BASELINE.json configs[1]); one step = one batched timestep",
                           "n_envs_per_gpu": n_envs, "n_agents": N_AGENTS, "obs_dim": OBS, "n_actions": N_ACTIONS,
                           "rnn_hidden": HID, "l2": "flushed between timed steps (256 MiB write)",
                           "timing": "CUDA events per step on the launch stream, summed; max over ranks"},
                "e2e": e2e, "gpu_launches": 2 * K, "roofline": roofline, "roofline_env": roofline_env,
                # the metric for the env step alone and the agent act alone (same launches, timed apart)
                "env_only": {"value": world * M / dt_env, "unit": "env-agent steps/s", "us_per_launch": dt_env * 1e6},
                "act_only": {"value": world * M / dt_agent, "unit": "env-agent steps/s", "us_per_launch": dt_agent * 1e6},
                "learner": learner_rec, "clocks": clocks, "cpu_baseline": cpu}
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
