#!/usr/bin/env python
"""Benchmark of the MA-CJD hot path on B200 (contract: see the build prompt / DESIGN.md).

Headline workload (BASELINE.json configs[1], "C2"): the reference's default scenario (2 jammers x 2 radars
x 1 target, obs/state 24, 5 actions, GRU 128) batched to 4096 parallel envs per GPU.  A *step* is one batched
environment timestep: the fused agent-act kernel over n_envs x n_agents rows followed by the fused env-step
kernel over n_envs episodes, both writing into the rollout trajectory in HBM.  `value` = env-agent steps/s
with everything resident in HBM (CUDA events per step, L2 flushed between steps); `e2e` = the same step through
the reference-facing host API (numpy obs / avail in pinned host memory -> select_actions -> actions back to the
host -> env.step -> obs / reward / terminated back).

The other BASELINE.json configs ride along in the same JSON line under `configs`:
  C3  scaled scenario 8 jammers x 16 radars x 4 targets, 8192 envs per GPU (65 536 over 8): env-only, act-only, fused
  C4  learner stress: global episode batch 1024 (split over the ranks), GRU hidden 256, mixer embed 128, all-reduce
  C5  end-to-end loop (act, store, sample, train, target update) with the replay ring sized for 1 M episodes over
      8 GPUs (125 000 per GPU), pipelined actor / learner streams
each with its roofline (one FLOP convention throughout: MINIMAL-work FLOPs, i.e. the Q-head's hidden product shared
by the A actions; the as-coded SURVEY 8d figure is reported beside it) and a bounded CPU baseline with its core count.

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
LEARNER_B, LEARNER_T = 32, 100
C3 = dict(J=8, R=16, K=4, n_envs=8192, env_bytes=8625.0)
C4 = dict(B=1024, H=256, E=128, T=100)
C5 = dict(episodes_total=1_000_000, gpus_nominal=8)


def rl_args(device, n_envs, **kw):
    a = types.SimpleNamespace(
        n_agents=N_AGENTS, n_actions=N_ACTIONS, state_shape=OBS, obs_shape=OBS, rnn_hidden_dim=HID,
        actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0, epsilon_finish=0.05,
        epsilon_anneal_time=100000, gamma=0.99, lr=5e-6, grad_norm_clip=1.0, target_update_interval=200,
        use_cuda=True, device=device, batch_size=LEARNER_B, buffer_size=2 * n_envs, episode_limit=100, seed=42,
        data_parallel=True, agent_kernel_path=0)     # 0: tcgen05 3xTF32 agent kernel where the dims allow
    for k, v in kw.items():
        setattr(a, k, v)
    a.env_info = {"state_shape": a.state_shape, "obs_shape": a.obs_shape, "n_actions": a.n_actions, "n_agents": a.n_agents,
                  "episode_limit": a.episode_limit}
    return a


def flop_per_row(O=OBS, A=N_ACTIONS, H=HID, AH=128):
    """MINIMAL-work FLOPs of one agent step: the Q-head's hidden product h' W1[:, :H]^T computed once and shared by
    the A actions (SURVEY 8d "minimal-work variant")."""
    return 2 * (O * AH + AH * AH + AH * A) + 2 * O * H + 12 * H * H + 2 * H * H + A * H * 5


def flop_per_row_as_coded(O=OBS, A=N_ACTIONS, H=HID, AH=128):
    """SURVEY 8d "as coded": the reference evaluates the whole Q-head once per action (core/mac.py:112-135)."""
    return 2 * O * H + 12 * H * H + 2 * (O * AH + AH * AH + AH * A) + A * 2 * ((H + A + 1) * H + H)


def flop_per_transition(N, O, A, H, AH, S, E, HH, minimal=True):
    """SURVEY 8d learner: 2 N F_agent (two unrolls) + 3 N F_qhead (q_taken fwd + bwd) + 4 F_mixer."""
    f_agent = flop_per_row(O, A, H, AH) if minimal else flop_per_row_as_coded(O, A, H, AH)
    f_qhead = 2 * ((H + A + 1) * H + H)
    f_mixer = 2 * (S * HH + HH * N * E) + 2 * (S * HH + HH * E) + 2 * S * E + 2 * (S * E + E) + 4 * N * E + 4 * E
    return 2 * N * f_agent + 3 * N * f_qhead + 4 * f_mixer


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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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


# ----------------------------------------------------------------------------- CPU baselines
def _all_threads():
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the baseline is meant to use every host core it can
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    return int(torch.get_num_threads())


def _random_agent_sd(O, A, H, AH=128):
    """random-init weights of the reference architecture (state_dict layout of RNNAgent)"""
    import torch.nn as nn
    mods = {"actor.0": nn.Linear(O, AH), "actor.2": nn.Linear(AH, AH), "actor.4": nn.Linear(AH, A),
            "fc1": nn.Linear(O, H), "fc2_q_head.0": nn.Linear(H + A + 1, H), "fc2_q_head.2": nn.Linear(H, 1)}
    sd = {}
    for k, m in mods.items():
        sd[k + ".weight"], sd[k + ".bias"] = m.weight.detach(), m.bias.detach()
    gru = nn.GRUCell(H, H)
    sd.update({"rnn.weight_ih": gru.weight_ih.detach(), "rnn.weight_hh": gru.weight_hh.detach(),
               "rnn.bias_ih": gru.bias_ih.detach(), "rnn.bias_hh": gru.bias_hh.detach()})
    return sd


def cpu_baseline_run(steps, warmup, n_envs=N_ENVS, spec=None, label=None):
    """The oracle port of the reference path on the host cores: NumPy float64 env step (vectorised over envs) +
    eager-PyTorch agent act (all torch threads).  One step = one batched timestep of `n_envs` envs."""
    import torch
    cores = _all_threads()
    from macjd_b200.simulation.scenario import default_spec
    from oracle.env_oracle import EnvOracle
    from oracle import agent_oracle as AO
    torch.manual_seed(42)
    ora = EnvOracle(spec if spec is not None else default_spec(n_envs))
    J, R, K = ora.J, ora.R, ora.K
    S, A = R * (6 + ora.types) + 2 * J, 2 * R + 1
    sd = _random_agent_sd(S, A, HID)
    rng = np.random.default_rng(7)
    h = torch.zeros(n_envs * J, HID)
    avail = torch.ones(n_envs, J, A, dtype=torch.long)
    ora.reset()

    def one_step(h):
        obs = torch.from_numpy(ora.get_obs())
        u = torch.from_numpy(rng.random((n_envs, J)).astype(np.float32))
        ra = torch.from_numpy(rng.integers(0, A, size=(n_envs, J)))
        with torch.no_grad():
            a, p, h, _, _ = AO.select_actions(sd, obs, avail, h, 0.5, False, u, ra)
        ora.step(a.view(n_envs, J).numpy(), p.view(n_envs, J).numpy(), rng.random((n_envs, R * K + J)))
        return h

    for _ in range(warmup):
        h = one_step(h)
    t0 = time.perf_counter()
    for _ in range(steps):
        h = one_step(h)
    dt = time.perf_counter() - t0
    return {"value": n_envs * J * steps / dt, "unit": "env-agent steps/s", "cores": cores,
            "kind": "port", "ms_per_step": dt / steps * 1e3,
            "sample": f"{steps} batched steps of {label or 'the same workload'} ({n_envs} envs x {J} agents): "
                      f"NumPy f64 env oracle (1 thread) + eager-PyTorch agent oracle ({cores} threads); "
                      f"host has {os.cpu_count()} cores"}


def cpu_reference_classes_run(n_envs=32, steps=30):
    """kind = "reference": the UNMODIFIED reference classes (baseline/_ref, a git-ignored copy of /root/reference made
    by __graft_entry__.build()) -- n_envs ElectromagneticEnvironment objects stepped in a Python loop (1 core, as the
    reference does) and the reference BasicMAC.select_actions at batch n_envs (all torch threads)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "simulation")):
        return {"unavailable": "baseline/_ref (copy of the reference sources) is not present on this box"}
    import contextlib
    import io
    import torch
    cores = _all_threads()
    cwd, saved_path = os.getcwd(), list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("core", "simulation", "utils", "runners")}
    try:
        os.chdir(ref)                               # the reference resolves config/*.yaml against the cwd
        for k in saved_mods:
            del sys.modules[k]
        sys.path.insert(0, ref)
        with contextlib.redirect_stdout(io.StringIO()):
            import yaml
            from simulation.environment import ElectromagneticEnvironment as RefEnv
            from core.mac import BasicMAC as RefMAC
            with open(os.path.join(ref, "config", "default.yaml")) as f:
                cfg = types.SimpleNamespace(**yaml.safe_load(f))
            cfg.device, cfg.use_cuda = "cpu", False
            envs = [RefEnv(cfg) for _ in range(n_envs)]
            info = envs[0].get_env_info()
            cfg.n_agents, cfg.n_actions = info["n_agents"], info["n_actions"]
            cfg.state_shape, cfg.obs_shape, cfg.episode_limit = info["state_shape"], info["obs_shape"], info["episode_limit"]
            torch.manual_seed(0)
            mac = RefMAC(info["obs_shape"], cfg)
            mac.init_hidden(n_envs)
            for e in envs:
                e.reset()

            def one_step(t):
                obs = torch.as_tensor(np.array([e.get_obs() for e in envs]), dtype=torch.float32)
                avail = torch.as_tensor(np.array([e.get_avail_actions() for e in envs]), dtype=torch.long)
                with torch.no_grad():
                    a, p = mac.select_actions(obs, avail, t, test_mode=False)
                a, p = a.squeeze(-1).numpy(), p.squeeze(-1).numpy()
                for i, e in enumerate(envs):
                    e.step([(int(a[i, j]), float(p[i, j])) for j in range(cfg.n_agents)])

            for t in range(3):
                one_step(t)
            t0 = time.perf_counter()
            for t in range(steps):
                one_step(3 + t)
            dt = time.perf_counter() - t0
        return {"value": n_envs * cfg.n_agents * steps / dt, "unit": "env-agent steps/s", "cores": cores, "kind": "reference",
                "ms_per_step": dt / steps * 1e3,
                "sample": f"{steps} timesteps of {n_envs} unmodified reference ElectromagneticEnvironment objects in a Python "
                          f"loop (1 core) + the reference BasicMAC.select_actions at batch {n_envs} ({cores} torch threads), "
                          f"default scenario; env-agent steps/s does not depend on n_envs for the scalar env loop"}
    except Exception as e:                         # the leg is optional: say why instead of failing the bench
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        os.chdir(cwd)
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in ("core", "simulation", "utils", "runners")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def _ref_env_worker(conn, RefEnv, cfg, n_local, n_agents, seed):
    """Child of cpu_reference_parallel_run (forked: the reference modules are already imported): steps its share of
    unmodified reference env objects on one core; an env that terminates is reset, as the reference runner does."""
    import contextlib
    import io
    try:
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            envs = [RefEnv(cfg) for _ in range(n_local)]
            for e in envs:
                e.reset()
        conn.send("ready")
        while True:
            msg = conn.recv()
            if msg is None:
                break
            with contextlib.redirect_stdout(io.StringIO()):
                if isinstance(msg, tuple):
                    a, p = msg
                    for i, e in enumerate(envs):
                        _, _, terminated, _ = e.step([(int(a[i, j]), float(p[i, j])) for j in range(n_agents)])
                        if terminated:
                            e.reset()
                obs = np.array([e.get_obs() for e in envs], dtype=np.float32)
                avail = np.array([e.get_avail_actions() for e in envs], dtype=np.int64)
            conn.send((obs, avail))
    except Exception as e:                          # the parent reports it and falls back
        try:
            conn.send(RuntimeError(f"{type(e).__name__}: {e}"))
        except Exception:
            pass
    finally:
        conn.close()


def cpu_reference_parallel_run(n_envs, steps, warmup=2):
    """kind = "reference" on ALL host cores: n_envs unmodified reference
    ElectromagneticEnvironment objects (baseline/_ref), split over one forked worker process per core and stepped in
    each worker's Python loop, + the reference BasicMAC.select_actions at batch n_envs on all torch threads --
    the loop body of the reference's runners/episode_runner.py:49-119 for the SAME workload the B200 arm steps."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "simulation")):
        return {"unavailable": "baseline/_ref (copy of the reference sources) is not present on this box"}
    import contextlib
    import io
    import multiprocessing as mp
    import torch
    try:
        n_cores = max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        n_cores = max(1, os.cpu_count() or 1)
    n_workers = max(1, min(n_cores, n_envs))
    cwd, saved_path = os.getcwd(), list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("core", "simulation", "utils", "runners")}
    procs, conns = [], []
    try:
        os.chdir(ref)                               # the reference resolves config/*.yaml against the cwd
        for k in saved_mods:
            del sys.modules[k]
        sys.path.insert(0, ref)
        with contextlib.redirect_stdout(io.StringIO()):
            import yaml
            from simulation.environment import ElectromagneticEnvironment as RefEnv
            from core.mac import BasicMAC as RefMAC
            with open(os.path.join(ref, "config", "default.yaml")) as f:
                cfg = types.SimpleNamespace(**yaml.safe_load(f))
            cfg.device, cfg.use_cuda = "cpu", False
            info = RefEnv(cfg).get_env_info()
        cfg.n_agents, cfg.n_actions = info["n_agents"], info["n_actions"]
        cfg.state_shape, cfg.obs_shape, cfg.episode_limit = info["state_shape"], info["obs_shape"], info["episode_limit"]
        # workers first (fork), before this process runs any multi-threaded torch work
        ctx = mp.get_context("fork")
        share = [n_envs // n_workers + (1 if i < n_envs % n_workers else 0) for i in range(n_workers)]
        for i, n_local in enumerate(share):
            parent, child = ctx.Pipe()
            pr = ctx.Process(target=_ref_env_worker, args=(child, RefEnv, cfg, n_local, cfg.n_agents, 1000 + i), daemon=True)
            pr.start()
            child.close()
            procs.append(pr)
            conns.append(parent)

        def recv(c):
            if not c.poll(300.0):
                raise RuntimeError("a reference env worker did not answer within 300 s")
            m = c.recv()
            if isinstance(m, Exception):
                raise m
            return m

        for c in conns:
            recv(c)                                 # "ready": envs constructed and reset
        cores = _all_threads()
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            mac = RefMAC(info["obs_shape"], cfg)
            mac.init_hidden(n_envs)
        offs = np.cumsum([0] + share)

        def gather(send):
            for i, c in enumerate(conns):
                c.send(send(i))
            parts = [recv(c) for c in conns]
            return (torch.from_numpy(np.concatenate([p[0] for p in parts])),
                    torch.from_numpy(np.concatenate([p[1] for p in parts])))

        obs, avail = gather(lambda i: "views")
        t_mac = 0.0

        def one_step(t, obs, avail):
            nonlocal t_mac
            t0 = time.perf_counter()
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                a, p = mac.select_actions(obs, avail, t, test_mode=False)
            a, p = a.squeeze(-1).numpy(), p.squeeze(-1).numpy()
            t_mac += time.perf_counter() - t0
            return gather(lambda i: (a[offs[i]:offs[i + 1]], p[offs[i]:offs[i + 1]]))

        for t in range(warmup):
            obs, avail = one_step(t, obs, avail)
        t_mac = 0.0
        t0 = time.perf_counter()
        for t in range(steps):
            obs, avail = one_step(warmup + t, obs, avail)
        dt = time.perf_counter() - t0
        return {"value": n_envs * cfg.n_agents * steps / dt, "unit": "env-agent steps/s", "cores": n_workers, "kind": "reference",
                "ms_per_step": dt / steps * 1e3, "ms_per_step_select_actions": t_mac / steps * 1e3,
                "sample": f"{steps} timesteps of the same workload ({n_envs} envs x {cfg.n_agents} agents): {n_envs} unmodified reference "
                          f"ElectromagneticEnvironment objects in {n_workers} worker processes (one per host core, {share[0]} envs "
                          f"each, stepped in the worker's Python loop; terminated envs reset) + the reference BasicMAC.select_actions "
                          f"at batch {n_envs} ({cores} torch threads); host has {os.cpu_count()} cores"}
    except Exception as e:                         # say why instead of failing the bench
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        for c in conns:
            try:
                c.send(None)
                c.close()
            except Exception:
                pass
        for pr in procs:
            pr.join(timeout=5.0)
            if pr.is_alive():
                pr.terminate()                     # (exactly the process this function started)
        os.chdir(cwd)
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in ("core", "simulation", "utils", "runners")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def cpu_learner_run(agent_sd, mixer_sd, batch, n_agents, embed, steps, B, T, label):
    """The oracle port of QMixLearner.train (eager PyTorch autograd, all host threads)."""
    import torch
    from oracle import agent_oracle as AO
    cores = _all_threads()
    ora = AO.LearnerOracle(agent_sd, mixer_sd, n_agents, embed, 0.99, 5e-6, 1.0, 200)
    ora.train(batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.train(batch)
    dt = (time.perf_counter() - t0) / steps
    return {"train_episodes_per_sec": B / dt, "train_transitions_per_sec": B * (T - 1) / dt,
            "ms_per_train_step": dt * 1e3, "cores": cores, "kind": "port",
            "sample": f"{steps} train steps of B = {B} x T = {T} ({label}) through oracle/agent_oracle.py LearnerOracle"}


def run_reference(args):
    """--impl reference: the reference path on the host cores for the SAME config (N GPUs x 4096 envs per step,
    bounded number of steps).  Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    n_total = args.n_envs * world
    steps, warmup = min(args.steps, max(4, 200 // world)), max(min(args.warmup, 3), 1)
    # headline: the UNMODIFIED reference classes on every host core (forked first: no torch thread pool exists yet);
    # beside it the oracle port (a much stronger CPU baseline: the env step vectorised over the envs in NumPy) and the
    # reference's own single-process loop
    ref_all = cpu_reference_parallel_run(n_total, steps, warmup)
    port = cpu_baseline_run(steps, warmup, n_envs=n_total)
    port = {k: port[k] for k in ("value", "unit", "cores", "kind", "ms_per_step", "sample")}
    ref_one = cpu_reference_classes_run()
    head = ref_all if "value" in ref_all else port
    line = {"impl": "reference", "metric": "env_agent_steps_per_sec", "value": head["value"], "unit": "env-agent steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 env + f32 nets",
            "data": "synthetic",
            "config": bench_config(args.n_envs),
            "cpu_baseline": {k: head[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "port": port,
            "reference_all_cores": ref_all,
            "reference_classes": ref_one,
            "e2e": {"value": head["value"], "unit": "env-agent steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_config(n_envs):
    """`config` of the JSON line -- identical keys for both arms."""
    return {"workload": "default scenario x 4096 envs per GPU, fused agent act + fused env step "
                        "(BASELINE.json configs[1]); one step = one batched timestep",
            "n_envs_per_gpu": n_envs, "n_agents": N_AGENTS, "obs_dim": OBS, "n_actions": N_ACTIONS,
            "rnn_hidden": HID, "l2": "flushed between timed steps (256 MiB write)",
            "timing": "CUDA events per step on the launch stream, summed; max over ranks"}


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
    ap.add_argument("--only", default="", help="comma list of {c2,c3,c4,c5}: run a subset (development)")
    ap.add_argument("--c5-episodes-per-gpu", type=int, default=C5["episodes_total"] // C5["gpus_nominal"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    only = set(filter(None, args.only.split(","))) or {"c2", "c3", "c4", "c5"}

    import torch
    import torch.distributed as dist
    # library chatter must not pollute the one JSON line -- including what C libraries write to file descriptor 1
    # (NCCL prints its version banner there when NCCL_DEBUG is set): fd 1 points at stderr until the line is written
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(json_fd, "w")
    sys.stdout = sys.stderr
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = f"cuda:{local}"
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec, scaled_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    from macjd_b200 import main as loop

    n_envs, K, W = args.n_envs, args.steps, max(args.warmup, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
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
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    tf32_peak = bf16_peak / 2.0
    tf32_src = ("TF32 dense = MEASURED_PEAKS.json bf16_tflops / 2 (of measured)" if "bf16_tflops" in peaks
                else "TF32 dense = fallback 1590 / 2 (of fallback)")
    fp32_src = f"derived: {n_sm} SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz (no measured FP32 figure in MEASURED_PEAKS.json)"
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass

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

    def agent_kernel_name(mac):
        return "agent_forward_tc2_kernel" if mac.agent._pair_kernel_ok(mac.agent.packed()) and mac.agent.path in (0, 3) \
            else "agent_forward_kernel (FP32 SIMT)"

    def rollout_bench(env, mac, runner, K_, W_, with_e2e):
        """value / roofline / (e2e) of one rollout workload; shared by C2 (headline) and C3."""
        T = runner.episode_limit
        M = env.n_envs * env.num_jammers
        runner.run(store=runner.buffer is not None)       # warm-up: one full episode (also fills the replay ring)
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

        for i in range(W_):
            rollout_step(i)
        barrier()
        dt = timed_steps(rollout_step, K_, pre=pre_step)
        barrier()
        dt = reduce_max(dt)
        res = {"value": world * M * K_ / dt, "ms_per_step": dt / K_ * 1e3, "M": M, "launch": "one library call per timestep"}
        from macjd_b200 import _native as N_
        fn = N_.get_lib().lib.macjd_rollout_fused_supported
        fn.restype = N_.C.c_int
        res["fused"] = bool(runner.fused_step and fn(N_.C.byref(mac.agent.packed().cstruct()), N_.C.byref(env._ctab)))
        if res["fused"]:
            # The same K timesteps as launches of up to one episode each (macjd_rollout_steps: the CTA pairs loop over the
            # timesteps, recurrent state in shared memory, their envs' steps inline): K steps = ceil(K / T) launches, L2
            # flushed before each launch, CUDA events around each launch.
            def chunks(total):
                out, t = [], 0
                while total > 0:
                    n_ = min(total, T - t)
                    out.append((t, n_))
                    total -= n_
                    t = (t + n_) % T
                return out
            runner.reset()
            for t0, n_ in chunks(max(W_, 3)):
                if t0 == 0:
                    runner.reset()
                runner.rollout(t0, n_)
            plan = chunks(K_)
            runner.reset()
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plan]
            for i, (t0, n_) in enumerate(plan):
                if t0 == 0:
                    runner.reset()                         # episode boundary: outside the timed region
                flush.fill_(i & 0xFF)
                ev[i][0].record()
                runner.rollout(t0, n_)
                ev[i][1].record()
            torch.cuda.synchronize()
            barrier()
            dt_m = reduce_max(sum(a.elapsed_time(b) for a, b in ev) * 1e-3)
            res["per_step_launch"] = {"value": res["value"], "ms_per_step": res["ms_per_step"],
                                      "what": "the same K timesteps as one macjd_rollout_step launch each, L2 flushed before every step"}
            res["dt_persistent_per_step"] = dt_m / K_
            res.update({"value": world * M * K_ / dt_m, "ms_per_step": dt_m / K_ * 1e3, "launches": len(plan),
                        "launch": f"macjd_rollout_steps: {len(plan)} launch(es) of up to {T} timesteps (one episode) each; L2 flushed before each launch"})

        # ---- the two launches of one timestep, each alone (cached C structs: one ctypes call per launch)
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
        res["dt_agent"] = timed_steps(lambda i: lib.call("macjd_agent_forward", ctx, wts, aio), kn) / kn
        res["dt_env"] = timed_steps(lambda i: lib.call("macjd_env_step", ctx, env._ctab, eio_alone), kn) / kn
        res["dt_simt"] = timed_steps(lambda i: lib.call("macjd_agent_forward", ctx, wts, aio_simt), kn) / kn

        # ---- parity of the benchmarked (tensor-core) launch against the FP32 SIMT kernel on this step's real inputs
        if agent_kernel_name(mac).startswith("agent_forward_tc2"):
            obs = runner.traj["obs"][1].reshape(1, M, -1)
            h0 = runner.traj["hidden_state"][0].reshape(M, -1).clone()
            a = mac.agent.run(obs, h0.clone(), select=True, test_mode=True, want_q=True, want_hidden_seq=True, path=3)
            b = mac.agent.run(obs, h0.clone(), select=True, test_mode=True, want_q=True, want_hidden_seq=True, path=1)
            scale = max(1.0, float(b["q_all"].abs().max()))
            srt = torch.sort(b["q_all"][0], dim=-1).values
            margin = srt[:, -1] - srt[:, -2]
            differ = a["actions"][0] != b["actions"][0]
            res["parity"] = {
                "rows": M, "greedy_actions_differing_from_fp32_kernel": int(differ.sum()),
                "fraction": float(differ.float().mean()),
                "largest_fp32_margin_among_differing_rows": float(margin[differ].max()) if bool(differ.any()) else 0.0,
                "max_abs_q_diff_over_scale": float((a["q_all"] - b["q_all"]).abs().max()) / scale,
                "max_abs_hidden_diff": float((a["hidden_seq"] - b["hidden_seq"]).abs().max()),
                "stated_bound": "Q, h within 2e-4 of scale; identical actions where the FP32 margin exceeds 1e-4 x scale "
                                "(tests/test_gpu_fullsize.py holds both kernels to the float64 truth)"}
        if not with_e2e:
            return res

        # ---- e2e: the reference-facing host API (host buffers in, host buffers out) per step
        n = env.n_envs
        hb = env.host_buffers()                                    # pinned: act_d, act_p, reward, terminated, obs
        avail_h = torch.ones(n, env.num_jammers, mac.agent.n_actions, dtype=torch.uint8).pin_memory()
        t_env = [0]

        def host_step_two_calls():
            mac.select_actions_host(hb["obs"], avail_h, t_env[0], actions_out=hb["act_d"], power_out=hb["act_p"])
            env.step_host(hb)

        def host_step_fused():
            runner.t_env = t_env[0]
            runner.step_host(hb["obs"], avail_h, hb)

        # the same call with ONE observation row per env: in this environment every jammer observes the global state
        # (environment.py:512-522: get_obs() returns the get_state() array once per jammer), so a host-side caller need
        # not ship J copies of it: state [n, S] in, next state [n, S] out
        hb_s = {k: v for k, v in hb.items() if k != "obs"}
        hb_s["state"] = torch.zeros(n, env.state_dim, dtype=torch.float32).pin_memory()

        def host_step_state():
            runner.t_env = t_env[0]
            runner.step_host(hb_s["state"], avail_h, hb_s)

        def time_host(step_fn):
            def restart():
                env.reset()
                hb["obs"].copy_(env.get_obs())
                hb_s["state"].copy_(env.get_state())
                mac.init_hidden(n)
            restart()
            t_env[0] = 0

            def one():
                step_fn()
                t_env[0] += n
                if (t_env[0] // n) % T == 0:
                    restart()
            for _ in range(W_):
                one()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K_):
                one()
            torch.cuda.synchronize()
            return reduce_max(time.perf_counter() - t0)

        t_env_saved = runner.t_env
        dt_two = time_host(host_step_two_calls)
        dt_rep = time_host(host_step_fused)
        dt_e2e = time_host(host_step_state)
        runner.t_env = t_env_saved
        obs_b, act_b = hb["obs"].numel() * 4, hb["act_d"].numel() * 4 + hb["act_p"].numel() * 4
        st_b = hb_s["state"].numel() * 4
        out_b = hb["reward"].numel() * 4 + hb["terminated"].numel()
        res["e2e"] = {"value": world * M * K_ / dt_e2e, "unit": "env-agent steps/s", "h2d_bytes_per_step": int(st_b + avail_h.numel()),
                      "d2h_bytes_per_step": int(act_b + st_b + out_b), "ms_per_step": dt_e2e / K_ * 1e3,
                      "api": "BatchedEpisodeRunner.step_host (C-ABI macjd_rollout_step_host, one launch): pinned host state [n, S] (the one "
                             "observation every jammer of an env shares, environment.py:512-522) and masks in; actions, power, reward, "
                             "terminated and the next state out to pinned host buffers; one stream drain per step",
                      "replicated_obs": {"value": world * M * K_ / dt_rep, "ms_per_step": dt_rep / K_ * 1e3,
                                         "h2d_bytes_per_step": int(obs_b + avail_h.numel()), "d2h_bytes_per_step": int(act_b + obs_b + out_b),
                                         "api": "the same call with the reference's per-jammer copies of the observation "
                                                "([n, J, S] in and out)"},
                      "two_calls": {"value": world * M * K_ / dt_two, "ms_per_step": dt_two / K_ * 1e3,
                                    "h2d_bytes_per_step": int(obs_b + avail_h.numel() + act_b), "d2h_bytes_per_step": int(act_b + obs_b + out_b),
                                    "api": "BasicMAC.select_actions_host + ElectromagneticEnvironment.step_host (macjd_agent_act_host + "
                                           "macjd_env_step_host): the reference's two calls, two stream drains per step"}}
        return res

    def agent_roofline(mac, M, dt_agent, dt_simt, O, A, H, dt_persistent=None, traffic_key="agent_forward_tc2_kernel"):
        """dt_agent: one single-step launch of the agent kernel alone.  dt_persistent: per-timestep time of the multi-step
        launches the timed region consists of (the same kernel looping over an episode, env steps inline): when given,
        `achieved` / `frac` describe THAT launch (algorithmic FLOPs of its timesteps / its duration; the env work adds
        time but no FLOPs) and the single-step figures move to `single_step_launch`."""
        fpr, fpr_coded = flop_per_row(O, A, H), flop_per_row_as_coded(O, A, H)
        single = None
        if dt_persistent is not None:
            single = {"us_per_launch": dt_agent * 1e6, "achieved": M * fpr / dt_agent / 1e12,
                      "what": "one single-timestep launch of the agent kernel alone (no env step), L2 flushed"}
            dt_agent = dt_persistent
        name = agent_kernel_name(mac)
        tc = name.startswith("agent_forward_tc2")
        peak = tf32_peak if tc else fp32_peak
        tr = traffic.get(traffic_key) or {}
        r = {"kernel": name, "bound": "tensor" if tc else "fp32", "achieved": M * fpr / dt_agent / 1e12, "peak": peak,
             "unit": "TFLOP/s", "frac": M * fpr / dt_agent / 1e12 / peak, "traffic": tr.get("dram_bytes_per_launch") if tc else None,
             "flop_per_agent_step": fpr, "flop_convention": "minimal work (Q-head hidden product shared by the actions)",
             "flop_per_agent_step_as_coded": fpr_coded, "achieved_as_coded": M * fpr_coded / dt_agent / 1e12,
             "us_per_launch": dt_agent * 1e6, "peak_source": tf32_src if tc else fp32_src,
             "simt_kernel": {"kernel": "agent_forward_kernel<256>", "bound": "fp32", "us_per_launch": dt_simt * 1e6,
                             "achieved": M * fpr / dt_simt / 1e12, "peak": fp32_peak, "frac": M * fpr / dt_simt / 1e12 / fp32_peak,
                             "peak_source": fp32_src}}
        if single is not None:
            tr_r = traffic.get(traffic_key + "_rollout") or {}
            r["traffic"] = tr_r.get("dram_bytes_per_launch")
            r["traffic_note"] = (f"DRAM bytes of one launch of {tr_r.get('timesteps_per_launch')} timesteps (ncu --set full); "
                                 f"{tr_r.get('dram_bytes_per_timestep')} per timestep") if tr_r else None
            single["traffic"] = tr.get("dram_bytes_per_launch")
            single["frac"] = single["achieved"] / peak
            r["single_step_launch"] = single
            r["us_per_launch"] = "see us_per_timestep_in_launch x timesteps per launch (config.timed_as)"
            r["us_per_timestep_in_launch"] = dt_agent * 1e6
            r["kernel"] = name + "<whole step, env fused> looping over the timesteps of an episode (macjd_rollout_steps)"
        if tc:
            r["executed_tensor_tflops"] = 3 * M * fpr / dt_agent / 1e12
            r["note"] = ("achieved counts algorithmic FLOPs once; the 3xTF32 split executes 3x that on the tensor pipe. CTA pairs issue "
                         "M=128,N=128,K=8 cta_group::2 MMAs; MMA phases alternate with epilogue phases along the layer dependency "
                         "chain: latency-bound at 64 rows per SM, not at the roofline")
        return r

    def env_roofline(n, dt_env, bytes_per_step, key):
        return {"kernel": "env_step_kernel", "bound": "hbm", "achieved": n * bytes_per_step / dt_env / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": n * bytes_per_step / dt_env / 1e9 / hbm_peak,
                "traffic": (traffic.get("env_step_kernel") or {}).get(key), "bytes_per_env_step": bytes_per_step,
                "us_per_launch": dt_env * 1e6, "peak_source": hbm_src}

    # rank 0 samples its GPU (one nvidia-smi poller per box: every query takes driver locks that the other ranks' launch
    # paths share, and eight pollers next to eight host-bound e2e loops are measurable)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    line = {"metric": "env_agent_steps_per_sec", "unit": "env-agent steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 env physics + 3xTF32 tensor-core GEMMs (f32-level) + f32 epilogues", "data": "synthetic",
            "config": bench_config(n_envs), "configs": {}}
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    # =============================================================================== C2 (headline) + C1-size learner
    rl = rl_args(device, n_envs)
    torch.manual_seed(42)                       # identical weights on every rank
    env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=device, seed=1000 + rank)
    mac = BasicMAC(OBS, rl)
    mac.cuda()
    buf = EpisodeReplayBuffer(rl, device=device)
    runner = BatchedEpisodeRunner(env, mac, buf, rl)
    learner = QMixLearner(mac, rl)
    r2 = rollout_bench(env, mac, runner, K, W, with_e2e=True)
    M = r2["M"]
    line["config"]["launches_per_step"] = ("<= 1: the CTA-pair agent kernel also runs the env step of its rows' envs, and loops over the timesteps of an episode"
                                           if r2["fused"] else "2: agent_forward kernel + env_step kernel (programmatic dependent launch)")
    line["config"]["timed_as"] = r2["launch"]
    line["per_step_launch"] = r2.get("per_step_launch")
    line.update({"value": r2["value"], "ms_per_step": r2["ms_per_step"], "e2e": r2["e2e"],
                 "gpu_launches": r2.get("launches", (1 if r2["fused"] else 2) * K),
                 "roofline": agent_roofline(mac, M, r2["dt_agent"], r2["dt_simt"], OBS, N_ACTIONS, HID, r2.get("dt_persistent_per_step")),
                 "roofline_env": env_roofline(n_envs, r2["dt_env"], ENV_BYTES_PER_STEP, "dram_bytes_per_launch_at_bench_size"),
                 "parity": r2.get("parity"),
                 # the metric for the env step alone and the agent act alone (same launches, timed apart)
                 "env_only": {"value": world * M / r2["dt_env"], "unit": "env-agent steps/s", "us_per_launch": r2["dt_env"] * 1e6},
                 "act_only": {"value": world * M / r2["dt_agent"], "unit": "env-agent steps/s", "us_per_launch": r2["dt_agent"] * 1e6}})

    def learner_bench(learner, buf, B, T, n_steps, dims, label, sampled=False):
        """sample + train, all-reduce when N > 1; B = episodes per GPU.  sampled: through QMixLearner.train_sampled, the call
        the training loop makes (one GPU: the whole step as a replayed CUDA graph; data parallel: the same eager step)."""
        np.random.seed(1 + rank)
        for _ in range(3):
            learner.train(buf.sample(B, time_major=True), {})
        if sampled:
            for _ in range(3):               # eager (sizes the workspaces), capture + replay, replay
                learner.train_sampled(buf, B, {})
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n_steps):
            if sampled:
                stats = learner.train_sampled(buf, B, {})
            else:
                stats = learner.train(buf.sample(B, time_major=True), {}, lazy_stats=True, check_actions=False)
        ev1.record()
        barrier()
        dt_l = reduce_max(ev0.elapsed_time(ev1) * 1e-3) / n_steps
        last_loss = float(stats["stats_tensor"][0])
        N_, O_, A_, H_, S_, E_, HH_ = dims
        ftr, ftr_coded = flop_per_transition(N_, O_, A_, H_, 128, S_, E_, HH_), flop_per_transition(N_, O_, A_, H_, 128, S_, E_, HH_, False)
        tflops = B * (T - 1) * ftr / dt_l / 1e12
        ag = learner.mac.agent
        pair = ag._pair_kernel_ok(ag.packed()) and ag.path in (0, 3)
        tc = ag.path in (0, 3)                   # pair kernel, or batched layers on the tcgen05 GEMM (macjd_agent_unroll)
        peak = tf32_peak if tc else fp32_peak
        unroll_kernel = (("agent_forward_tc2_kernel (parts 3 + 2) around gru_rec_rows_kernel (the recurrence of few rows: FP32, weight_hh on chip)"
                          if B * N_ <= 1184 else "agent_forward_tc2_kernel (parts 3 + 4 + 2)") if pair else
                         "macjd_agent_unroll: batched layers on tc_gemm_kernel (tcgen05 3xTF32) + the recurrence as one gru_recurrence_tc2_kernel launch (CTA pairs, h in shared memory)" if tc else
                         "agent_forward_kernel (FP32 SIMT)")
        return {"train_episodes_per_sec": world * B / dt_l, "train_transitions_per_sec": world * B * (T - 1) / dt_l,
                "ms_per_train_step": dt_l * 1e3, "batch_episodes_per_gpu": B, "episode_len": T, "last_loss": last_loss,
                "includes": "replay sample (gather kernel) + 2 unrolls + mixers + TD + backward + clip/Adam"
                            + (" + NCCL all-reduce" if world > 1 else "")
                            + ("; issued as one replayed CUDA graph per step (QMixLearner.train_sampled)"
                               if sampled and learner._graphable() else ""),
                "roofline": {"bound": "tensor" if tc else "fp32", "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
                             "frac": tflops / peak, "flop_per_transition": ftr, "flop_convention": "minimal work",
                             "flop_per_transition_as_coded": ftr_coded, "achieved_as_coded": B * (T - 1) * ftr_coded / dt_l / 1e12,
                             "peak_source": tf32_src if tc else fp32_src,
                             "agent_unroll_kernel": unroll_kernel,
                             "executed_tensor_tflops": 3 * tflops if tc else None,
                             "note": label}}

    line["learner"] = learner_bench(
        learner, buf, LEARNER_B, LEARNER_T, args.learner_steps, (N_AGENTS, OBS, N_ACTIONS, HID, OBS, 64, 128),
        "B = 32 episodes is 64 agent rows: the step is bound by its serial depth (2 x T dependent GRU steps, which run on the "
        "row-split FP32 recurrence kernel with weight_hh on chip, and ~90 short dependent launches), not by throughput",
        sampled=True)
    cpu = None
    if want_cpu:
        cpu = cpu_baseline_run(args.cpu_steps, 3)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        agent_sd = {k: v.detach().cpu().clone() for k, v in mac.agent.state_dict().items()}
        mixer_sd = {k: v.detach().cpu().clone() for k, v in learner.eval_qmix_net.state_dict().items()}
        batch = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in buf.sample(LEARNER_B).items()}
        cpu["learner"] = cpu_learner_run(agent_sd, mixer_sd, batch, N_AGENTS, 64, 2, LEARNER_B, LEARNER_T, "C1 dims")
        cpu["reference_classes"] = cpu_reference_classes_run()
        # the unmodified reference classes on every host core, in a child interpreter (its worker processes are forked,
        # which this process -- CUDA context, NCCL, poller thread -- should not do)
        try:
            env_ = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
            env_["CUDA_VISIBLE_DEVICES"] = ""
            out = subprocess.run([sys.executable, "-c", "import json, bench; print('REF_ALL ' + json.dumps("
                                  f"bench.cpu_reference_parallel_run({n_envs}, 12, 2)))"], cwd=ROOT, env=env_,
                                 capture_output=True, text=True, timeout=900)
            got = [ln[8:] for ln in out.stdout.splitlines() if ln.startswith("REF_ALL ")]
            cpu["reference_all_cores"] = json.loads(got[-1]) if got else {"unavailable": (out.stderr or "no output")[-300:]}
        except Exception as e:
            cpu["reference_all_cores"] = {"unavailable": f"{type(e).__name__}: {e}"}
    line["cpu_baseline"] = cpu
    del runner, learner, buf, env, mac
    torch.cuda.empty_cache()

    # =============================================================================== C3: scaled scenario
    if "c3" in only:
        J, R, Kt, n3 = C3["J"], C3["R"], C3["K"], C3["n_envs"]
        S3, A3 = R * 10 + 2 * J, 2 * R + 1
        rl3 = rl_args(device, n3, n_agents=J, n_actions=A3, state_shape=S3, obs_shape=S3, buffer_size=n3)
        torch.manual_seed(43)
        spec3 = scaled_spec(n3, n_jammers=J, n_radars=R, n_targets=Kt, seed=1234 + rank)
        env3 = ElectromagneticEnvironment(rl3, spec=spec3, device=device, seed=2000 + rank)
        mac3 = BasicMAC(S3, rl3)
        mac3.cuda()
        runner3 = BatchedEpisodeRunner(env3, mac3, None, rl3)
        k3 = max(10, min(K, 50))
        r3 = rollout_bench(env3, mac3, runner3, k3, 3, with_e2e=False)
        M3 = r3["M"]
        blk = {"workload": f"scaled scenario {J} jammers x {R} radars x {Kt} targets (obs {S3}, {A3} actions), {n3} envs per GPU "
                           f"({n3 * world} over {world} GPU{'s' if world > 1 else ''}; BASELINE.json configs[2] = 65 536 over 8), "
                           "per-env scenario tables; L2 flushed between steps",
               "value": r3["value"], "unit": "env-agent steps/s", "ms_per_step": r3["ms_per_step"], "steps": k3, "scaling": "weak",
               "launches_per_step": 1 if r3["fused"] else 2,
               "env_only": {"value": world * M3 / r3["dt_env"], "us_per_launch": r3["dt_env"] * 1e6},
               "act_only": {"value": world * M3 / r3["dt_agent"], "us_per_launch": r3["dt_agent"] * 1e6},
               "roofline": agent_roofline(mac3, M3, r3["dt_agent"], r3["dt_simt"], S3, A3, HID, r3.get("dt_persistent_per_step"),
                                          traffic_key="agent_forward_tc2_kernel_c3"),
               "roofline_env": env_roofline(n3, r3["dt_env"], C3["env_bytes"], "dram_bytes_per_launch_c3"),
               "parity": r3.get("parity"), "cpu_baseline": None}
        if want_cpu:
            c = cpu_baseline_run(3, 1, n_envs=1024, spec=scaled_spec(1024, n_jammers=J, n_radars=R, n_targets=Kt, seed=1234),
                                 label="the C3 scenario at 1 024 envs")
            blk["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["configs"]["C3"] = blk
        del runner3, env3, mac3
        torch.cuda.empty_cache()

    # =============================================================================== C4: learner stress
    if "c4" in only:
        B4, H4, E4, T4 = C4["B"], C4["H"], C4["E"], C4["T"]
        Bg = max(1, B4 // world)                                   # global batch 1024 split over the ranks
        rl4 = rl_args(device, Bg, rnn_hidden_dim=H4, mixing_embed_dim=E4, batch_size=Bg, buffer_size=Bg)
        torch.manual_seed(44)
        mac4 = BasicMAC(OBS, rl4)
        mac4.cuda()
        buf4 = EpisodeReplayBuffer(rl4, device=device)
        g = torch.Generator(device=device).manual_seed(11 + rank)
        rn = lambda *s: torch.randn(*s, device=device, generator=g)
        buf4.store_rollout({                                       # SURVEY 8d learner batch: synthetic N(0,1) episodes
            "state": rn(T4 + 1, Bg, OBS), "obs": rn(T4 + 1, Bg, N_AGENTS, OBS),
            "actions_discrete": torch.randint(0, N_ACTIONS, (T4, Bg, N_AGENTS, 1), device=device, generator=g, dtype=torch.int32),
            "actions_continuous": torch.rand(T4, Bg, N_AGENTS, 1, device=device, generator=g),
            "avail_actions": torch.ones(T4 + 1, Bg, N_AGENTS, N_ACTIONS, dtype=torch.uint8, device=device),
            "reward": rn(T4, Bg, 1), "terminated": torch.zeros(T4, Bg, 1, dtype=torch.uint8, device=device),
            "hidden_state": rn(T4 + 1, Bg, N_AGENTS, H4) * 0.5})
        learner4 = QMixLearner(mac4, rl4)
        blk = learner_bench(learner4, buf4, Bg, T4, max(3, args.learner_steps // 2), (N_AGENTS, OBS, N_ACTIONS, H4, OBS, E4, 128),
                            f"global batch {B4} episodes = {Bg} per GPU x {world} GPU(s), GRU hidden {H4}, mixer embed {E4}")
        blk["workload"] = (f"QMix learner stress (BASELINE.json configs[3]): episode batch {B4} in total ({Bg} per GPU), T = {T4}, "
                           f"GRU hidden {H4}, mixer embed {E4}" + (", NCCL gradient all-reduce" if world > 1 else ""))
        blk["scaling"] = "strong (global batch fixed at 1024 episodes)"
        blk["cpu_baseline"] = None
        if want_cpu:
            agent_sd = {k: v.detach().cpu().clone() for k, v in mac4.agent.state_dict().items()}
            mixer_sd = {k: v.detach().cpu().clone() for k, v in learner4.eval_qmix_net.state_dict().items()}
            np.random.seed(3)
            batch = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in buf4.sample(16).items()}
            blk["cpu_baseline"] = cpu_learner_run(agent_sd, mixer_sd, batch, N_AGENTS, E4, 1, 16, T4, "C4 dims, 16 of the 1024 episodes")
        line["configs"]["C4"] = blk
        del learner4, buf4, mac4
        torch.cuda.empty_cache()

    # =============================================================================== C5: end-to-end loop, big replay ring
    if "c5" in only:
        cap = int(args.c5_episodes_per_gpu)
        n5, rollouts, tpr = n_envs, max(34, cap // n_envs + 4), 16
        cfg = loop.default_config(buffer_size=cap, total_env_steps=rollouts * n5 * 100, start_training_steps=0,
                                  train_steps_per_rollout=tpr, save_model=False, test_nepisodes=0, test_interval=0,
                                  log_interval_seconds=1e9, seed=42 + rank, data_parallel=True, agent_kernel_path=0,
                                  device=device, replay_fast_sampling=os.environ.get("MACJD_BENCH_C5_REFERENCE_DRAWS", "0") == "0")
        barrier()
        t0 = time.perf_counter()
        out = loop.run(cfg, spec=default_spec(n5), writer=False, log=lambda *_: None, pipeline=True,
                       use_graph=os.environ.get("MACJD_BENCH_C5_ROLLOUT_GRAPH", "0") != "0")
        torch.cuda.synchronize()
        dt5 = reduce_max(time.perf_counter() - t0)
        loop_s = reduce_max(out["time_s"])             # main.run's own clock: the loop alone (device drained at its end)
        held = len(out["buffer"])
        bpe = out["buffer"].bytes_per_episode()
        line["configs"]["C5"] = {
            "workload": f"end-to-end training loop (macjd_b200.main.run, pipeline=True: act with a frozen copy on one stream while the "
                        f"learner trains on another, store, sample, train, target update) with the replay ring sized for "
                        f"{C5['episodes_total']} episodes over {C5['gpus_nominal']} GPUs = {cap} per GPU (BASELINE.json configs[4]); "
                        f"{rollouts} rollouts of {n5} episodes x 100 steps, {tpr} train steps of B = 32 per rollout; batches drawn with "
                        + ("the ring's O(batch) sampler (replay_fast_sampling: uniform without replacement, Floyd)"
                           if cfg.replay_fast_sampling else "the reference's np.random.choice (a full permutation per draw)"),
            "ring_capacity_episodes_per_gpu": cap, "ring_bytes_per_episode": bpe, "ring_gb_per_gpu": cap * bpe / 1e9,
            "episodes_held_per_gpu": held, "ring_wrapped": bool(rollouts * n5 > cap),
            "env_agent_steps_per_sec": world * out["total_steps"] * N_AGENTS / dt5,
            "train_episodes_per_sec": world * out["train_steps"] * 32 / dt5,
            "train_transitions_per_sec": world * out["train_steps"] * 32 * 99 / dt5,
            "train_steps": out["train_steps"], "rollouts": rollouts, "wall_s": dt5, "scaling": "weak",
            "timing": "host wall clock around the whole loop incl. construction of env / ring / networks, max over ranks",
            "loop_only": {"wall_s": loop_s, "env_agent_steps_per_sec": world * out["total_steps"] * N_AGENTS / loop_s,
                          "train_episodes_per_sec": world * out["train_steps"] * 32 / loop_s,
                          "what": "the same run on main.run's own clock (time_s): from the first rollout to the drained "
                                  "device, without the construction of env / ring / networks; still includes the first "
                                  "eager train step and the capture of the step graph"},
            "last_loss": out["last_logged"].get("avg_loss")}
        del out
        torch.cuda.empty_cache()

    line["clocks"] = sampler.stop()
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
