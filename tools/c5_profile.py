"""Where does the host spend the end-to-end loop (bench C5: 34 rollouts of 4 096 episodes, 16 train steps each, 125 000-episode
ring)?  cProfile of macjd_b200.main.run, top entries by own time -- a blocked read-back shows up as `tolist` / `item` / `cpu`.
   python tools/c5_profile.py [rollouts]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from macjd_b200 import main as loop   # noqa: E402
from macjd_b200.simulation.scenario import default_spec   # noqa: E402

os.environ.setdefault("MACJD_LOOP_TIMING", "1")
rollouts = int(sys.argv[1]) if len(sys.argv) > 1 else 34
cfg = loop.default_config(buffer_size=125000, total_env_steps=rollouts * 4096 * 100, start_training_steps=0, train_steps_per_rollout=16,
                          save_model=False, test_nepisodes=0, test_interval=0, log_interval_seconds=1e9, seed=42,
                          agent_kernel_path=0, device="cuda:0", replay_fast_sampling=True)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out = loop.run(cfg, spec=default_spec(4096), writer=False, log=lambda *_: None, pipeline=True)
torch.cuda.synchronize()
pr.disable()
print(f"wall {time.perf_counter() - t0:.3f} s, loop {out['time_s']:.3f} s, {out['train_steps']} train steps, {rollouts} rollouts, device ms per phase (median): {out['phase_ms']}")
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
