"""Worker of test_gpu_learner.py::test_dp_train_sampled_graph_equals_eager (one process per GPU, torch.distributed.run):
the data-parallel QMix train step (core/qmix.py:129-205 per rank + one gradient all-reduce) replayed as a CUDA graph --
NCCL's kernel captured with the step -- must leave every rank with the same statistics, parameters, target networks and
optimiser state, bit for bit, as the eager data-parallel step; and the ranks must agree with each other."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tests.test_gpu_learner import _sampled_learner
    out = {}
    for mode in ("1", "0"):
        os.environ["MACJD_TRAIN_GRAPH_DP"] = mode
        learner, buf = _sampled_learner(seed=5 + rank, device=f"cuda:{local}", same_init_seed=5)
        assert learner._graphable() == (mode == "1")
        np.random.seed(11 + rank)
        stats = [learner.train_sampled(buf, 16, {})["stats_tensor"] for _ in range(7)]
        torch.cuda.synchronize()
        captured = "_step_graphs" in learner.__dict__ and any(isinstance(g, dict) for g in learner._step_graphs.values())
        assert captured == (mode == "1"), (mode, captured)
        flat = torch.cat([v.detach().reshape(-1).float() for sd in (learner.mac.agent.state_dict(), learner.eval_qmix_net.state_dict(),
                                                                     learner.target_qmix_net.state_dict(),
                                                                     learner.target_mac.agent.state_dict())
                          for v in sd.values()] + [learner._opt_state["m"].reshape(-1), learner._opt_state["v"].reshape(-1)])
        out[mode] = (torch.stack(stats).clone(), flat.clone(), (learner.train_step, learner._opt_state["step"],
                                                                 learner.last_target_update_step))
    g, e = out["1"], out["0"]
    assert torch.isfinite(g[0]).all() and g[0].abs().sum() > 0
    assert torch.equal(g[0], e[0]), (g[0], e[0])
    assert torch.equal(g[1], e[1]), int((g[1] != e[1]).sum())
    assert g[2] == e[2] == (7, 7, 6), (g[2], e[2])
    # the ranks trained on different episodes but hold the same networks (and saw the same global statistics)
    both = [torch.empty_like(g[1]) for _ in range(dist.get_world_size())]
    dist.all_gather(both, g[1])
    assert all(torch.equal(both[0], b) for b in both[1:])
    st = [torch.empty_like(g[0]) for _ in range(dist.get_world_size())]
    dist.all_gather(st, g[0])
    assert all(torch.equal(st[0][:, :2], s[:, :2]) for s in st[1:])          # loss and gradient norm are global
    dist.barrier()
    if rank == 0:
        print("DP_GRAPH_OK", g[0][-1].tolist())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
