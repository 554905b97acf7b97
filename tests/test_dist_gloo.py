"""Data-parallel learner on 2 ranks (gloo, CPU, host-emulated kernels): every rank trains on its
own half of the episodes, one flat bucket (gradients + loss sums) is all-reduced, and both ranks
must end with the same weights as a single process trained on the whole batch."""
import json
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _load():
    g = np.load(os.path.join(GOLDEN, "learner_small_fastlr.npz"))
    args = types.SimpleNamespace(**json.loads(str(g["args_json"])))
    agent0 = {k[len("agent0."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("agent0.")}
    mixer0 = {k[len("mixer0."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("mixer0.")}
    batch = {k[len("step0.batch."):]: (int(g[k]) if g[k].ndim == 0 else g[k]) for k in g.files if k.startswith("step0.batch.")}
    return args, agent0, mixer0, batch


def _train(args, agent0, mixer0, batches, group=None):
    from tests.helpers import emul_lib
    from tests.learner_checks import make_learner
    L = make_learner(args, agent0, mixer0, "cpu", emul_lib())
    L.process_group = group
    stats = [L.train(b, {}) for b in batches]
    return L, stats


def _slice(batch, sl):
    return {k: (v[sl] if isinstance(v, np.ndarray) else v) for k, v in batch.items()}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        args, agent0, mixer0, batch = _load()
        B = batch["state"].shape[0]
        per = B // world                       # ragged on purpose: the last rank also takes the remainder
        sl = slice(rank * per, B if rank == world - 1 else (rank + 1) * per)
        L, stats = _train(args, agent0, mixer0, [_slice(batch, sl)] * 2, group=dist.group.WORLD)
        sd = {"agent." + k: v.detach().clone() for k, v in L.mac.agent.state_dict().items()}
        sd.update({"mixer." + k: v.detach().clone() for k, v in L.eval_qmix_net.state_dict().items()})
        torch.save({"sd": sd, "stats": stats}, os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_equals_single_process(tmp_path):
    from tests.emul.build_emul import build
    build()                                     # compile once in the parent
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    for k in r0["sd"]:
        assert torch.equal(r0["sd"][k], r1["sd"][k]), f"ranks diverged on {k}"
    assert r0["stats"] == r1["stats"]
    args, agent0, mixer0, batch = _load()
    L, stats = _train(args, agent0, mixer0, [batch] * 2)
    for name, ref in (("loss", stats), ):
        for s_dp, s_1 in zip(r0["stats"], stats):
            for key in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
                np.testing.assert_allclose(s_dp[key], s_1[key], rtol=2e-5, err_msg=key)
    single = {"agent." + k: v for k, v in L.mac.agent.state_dict().items()}
    single.update({"mixer." + k: v for k, v in L.eval_qmix_net.state_dict().items()})
    for k, v in single.items():
        np.testing.assert_allclose(r0["sd"][k].numpy(), v.detach().numpy(), rtol=1e-4, atol=1e-6, err_msg=k)
