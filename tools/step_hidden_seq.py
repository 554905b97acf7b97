"""What the second hidden-state store costs: flushed rollout step with and without hidden_seq (the per-step
hidden state the replay keeps) next to the in-place `hidden` update.  Measured before the runner chained the
state through the h_t records (41.7 vs 39.9 us); with the chaining the second variant only drops the record
store of steps whose successor reads it, so its numbers are a timing experiment, not a valid rollout."""
import sys, torch
sys.path.insert(0, ".")
import bench
from tools.microbench import flush_l2
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
dev = "cuda:0"; n_envs = 4096
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl); mac.cuda()
runner = BatchedEpisodeRunner(env, mac, EpisodeReplayBuffer(rl, device=dev), rl)
runner.reset()
runner.step(0)
def timed(label):
    ts = []
    for rep in range(3):
        tot = 0.0
        for t in range(60):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); runner.step(t % runner.episode_limit); e1.record()
            torch.cuda.synchronize()
            if t >= 10: tot += e0.elapsed_time(e1)
        ts.append(tot / 50 * 1e3)
    print(f"{label}: " + ", ".join(f"{x:.2f}" for x in ts) + " us per flushed step")
timed("hidden + hidden_seq")
for io in runner._agent_io: io.hidden_seq = None
timed("hidden only")
