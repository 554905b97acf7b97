"""Learner / mixer / replay parity checks shared by the host-emulation (CPU) and GPU tests."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import agent_oracle as AO
from tests.helpers import GOLDEN
from macjd_b200 import _native as N


def _load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return g, types.SimpleNamespace(**json.loads(str(g["args_json"])))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _ctx(dev):
    return N.torch_ctx(dev) if torch.device(dev).type == "cuda" else N.Ctx(device=0, reserved=0, stream=None)


def check_mixer_against_golden(name, device, lib):
    """QMixer forward and every parameter gradient against the reference module's autograd."""
    from macjd_b200.core.networks import QMixer
    g, args = _load("mixer_" + name)
    net = QMixer(args)
    net.load_state_dict(_sd(g, "sd."))
    net.to(device)
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])
    dev = torch.device(device)
    q = torch.from_numpy(g["q"]).to(dev)
    s = torch.from_numpy(g["s"]).to(dev)
    R = q.shape[0]
    dims = N.MixerDims(n_rows=R, state_dim=args.state_shape, n_agents=args.n_agents, embed_dim=args.mixing_embed_dim,
                       hyper_hidden=args.hyper_hidden_dim, reserved=0)
    wsf = lib.lib.macjd_mixer_workspace_floats(dims)
    ws = torch.empty(wsf, dtype=torch.float32, device=dev)
    sd = dict(net.named_parameters())
    w = N.MixerParams(**{f: sd[k].data_ptr() for f, k in zip(N.MIXER_FIELDS, N.MIXER_KEYS)})
    y = torch.empty(R, dtype=torch.float32, device=dev)
    lib.callv("macjd_mixer_forward", _ctx(dev), dims, w, q, s, y, ws, wsf)
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=1e-5, atol=2e-6)
    grads = {k: torch.full_like(p, 7.0) for k, p in sd.items()}      # poisoned: must be overwritten
    gw = N.MixerParams(**{f: grads[k].data_ptr() for f, k in zip(N.MIXER_FIELDS, N.MIXER_KEYS)})
    dq = torch.empty(R, args.n_agents, dtype=torch.float32, device=dev)
    dy = torch.from_numpy(g["w"]).to(dev)
    lib.callv("macjd_mixer_backward", _ctx(dev), dims, w, q, dy, ws, wsf, gw, dq)
    np.testing.assert_allclose(dq.cpu().numpy(), g["dq"], rtol=1e-4, atol=1e-6)
    for k in sd:
        ref = g["grad." + k]
        np.testing.assert_allclose(grads[k].cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(ref).max()), err_msg=k)


def make_learner(args, agent_sd, mixer_sd, device, lib):
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    args.use_cuda = torch.device(device).type == "cuda"
    args.device = device
    mac = BasicMAC(args.obs_shape, args, _lib=lib)
    mac.agent.load_state_dict(agent_sd)
    learner = QMixLearner(mac, args, _lib=lib)
    learner.eval_qmix_net.load_state_dict(mixer_sd)
    learner._update_targets()
    return learner


def _check_adam_delta(new, ref_new, w0, ref_grad, lr, k_adam, tiny_seen, key, step):
    """Post-Adam weights as deltas from the initial weights, against the reference's.  Adam moves an element by
    lr * m / (sqrt(v) + 1e-8): where the (clipped) gradient is of the order of that epsilon the move is a fraction of lr
    that hangs on the gradient's last bits, so elements outside the bound must be such elements (|reference gradient|
    < 1e-6 at this or an earlier step -- the deltas are cumulative), rare (<= 0.1 % of the tensor) and within one Adam
    step per train step."""
    tiny = tiny_seen[key] = tiny_seen.get(key, False) | (np.abs(ref_grad) < 1e-6)
    err = np.abs((new - w0) - (ref_new - w0))
    bad = err > lr * 5e-3 * k_adam + 2.4e-7 * np.abs(w0).max() + 5e-3 * k_adam * np.abs(ref_new - w0)
    if bad.any():
        assert bad.sum() <= max(2, 1e-3 * bad.size) and tiny[bad].all() and (err[bad] <= lr * (step + 1)).all(), \
            (key, step, int(bad.sum()), bad.size, float(err.max()), int((bad & ~tiny).sum()))


def check_learner_against_golden(name, device, lib, path=1):
    """Whole train steps against the unmodified reference learner: stats, (clipped) gradients
    of every trained tensor, post-Adam weights, target sync; untrained tensors stay frozen.
    path = agent_kernel_path of the unrolls: 1 FP32 SIMT (bounds: stats 5e-5, gradients 5e-4, Adam deltas 5e-3),
    0 the benchmarked tcgen05 3xTF32 pair kernel with split unrolls where the dims allow it (stated looser
    bounds: stats 2e-4, gradients 2e-3, Adam deltas 1e-2 -- the tensor-core Q's only enter through the
    double-DQN arg-max and the target network's Q at it)."""
    g, args = _load("learner_" + name)
    args.agent_kernel_path = path
    k_stats, k_grad, k_adam = (1.0, 1.0, 1.0) if path == 1 else (4.0, 4.0, 2.0)
    agent0, mixer0 = _sd(g, "agent0."), _sd(g, "mixer0.")
    L = make_learner(args, agent0, mixer0, device, lib)
    tiny_seen = {}
    for step in range(int(g["n_steps"])):
        pre = f"step{step}."
        batch = {}
        for k in g.files:
            if k.startswith(pre + "batch."):
                v = g[k]
                batch[k[len(pre + "batch."):]] = int(v) if v.ndim == 0 else v
        stats = L.train(batch, {}, return_debug=True)
        dbg = stats.pop("debug")
        ref = g[pre + "stats"]
        np.testing.assert_allclose([stats["loss"], stats["grad_norm"], stats["eval_qtot_avg"], stats["target_qtot_avg"]],
                                   ref, rtol=5e-5 * k_stats, atol=5e-7 * k_stats, err_msg=f"stats step {step}")   # (atol: means that cancel)
        coef = min(1.0, args.grad_norm_clip / (stats["grad_norm"] + 1e-6))
        scale = coef / float(dbg["sums"][1])
        off = 0
        for nm, n in zip(dbg["names"], dbg["sizes"]):
            kind, key = nm.split(".", 1)
            ref_g = g[pre + ("agent_grad." if kind == "agent" else "mixer_grad.") + key]
            mine = dbg["grad"][off:off + n].cpu().numpy().reshape(ref_g.shape) * scale
            np.testing.assert_allclose(mine, ref_g, rtol=5e-4 * k_grad, atol=5e-6 * k_grad * max(1e-3, np.abs(ref_g).max()),
                                       err_msg=f"{nm} step {step}")
            off += n
        sd_now = {k: v.detach().cpu().numpy() for k, v in L.mac.agent.state_dict().items()}
        for k in agent0:
            if k in AO.TRAINED_AGENT_KEYS:
                _check_adam_delta(sd_now[k], g[pre + "agent." + k], agent0[k].numpy(), g[pre + "agent_grad." + k], args.lr, k_adam,
                                  tiny_seen, "agent." + k, step)
            else:
                np.testing.assert_array_equal(sd_now[k], agent0[k].numpy(), err_msg=f"{k} must stay frozen")
        tgt_now = L.target_mac.agent.state_dict()
        for k in AO.TRAINED_AGENT_KEYS:
            np.testing.assert_allclose(tgt_now[k].cpu().numpy(), g[pre + "tgt_agent." + k], rtol=1e-5, atol=1e-6)
        mix_now = L.eval_qmix_net.state_dict()
        tmix_now = L.target_qmix_net.state_dict()
        for k in mixer0:
            _check_adam_delta(mix_now[k].cpu().numpy(), g[pre + "mixer." + k], mixer0[k].numpy(), g[pre + "mixer_grad." + k], args.lr,
                              k_adam, tiny_seen, "mixer." + k, step)
            np.testing.assert_allclose(tmix_now[k].cpu().numpy(), g[pre + "tgt_mixer." + k], rtol=1e-5, atol=1e-6)
    return L


def check_replay_against_golden(device, lib, name="replay"):
    """store_episode / sample against the reference ring buffer: ring index arithmetic, padding
    rules, trimming to the longest sampled episode, dtypes of the returned dict.  Recordings with a `np_seed` per sample
    (the live differential test) also hold the public sample() to the reference's own index draws."""
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    g, args = _load(name)
    buf = EpisodeReplayBuffer(args, device=device, _lib=lib)
    keys = ("state", "obs", "actions_discrete", "actions_continuous", "avail_actions", "reward", "terminated", "hidden_state")
    for i in range(int(g["n_eps"])):
        buf.store_episode({k: [g[f"ep{i}.{k}"]] for k in keys})
        assert buf.current_index == int(g[f"after{i}.current_index"])
        assert buf.current_size == int(g[f"after{i}.current_size"]) == len(buf)
        if f"sample{i}.indices" in g.files:
            out = buf.gather(g[f"sample{i}.indices"])
            assert out["max_seq_len"] == int(g[f"sample{i}.max_seq_len"])
            for k in keys + ("filled",):
                ref = g[f"sample{i}.{k}"]
                mine = out[k].cpu().numpy()
                assert mine.shape == ref.shape, k
                np.testing.assert_array_equal(mine.astype(ref.dtype), ref, err_msg=k)
            if f"sample{i}.np_seed" in g.files:
                np.random.seed(int(g[f"sample{i}.np_seed"]))
                s = buf.sample(len(g[f"sample{i}.indices"]))
                assert s["max_seq_len"] == int(g[f"sample{i}.max_seq_len"])
                for k in keys + ("filled",):
                    ref = g[f"sample{i}.{k}"]
                    np.testing.assert_array_equal(s[k].cpu().numpy().astype(ref.dtype), ref, err_msg=f"seeded sample: {k}")
    # the public sample(): reference dtypes for masks / avail, no replacement
    np.random.seed(0)
    s = buf.sample(4)
    assert s["terminated"].dtype == torch.bool and s["filled"].dtype == torch.bool and s["avail_actions"].dtype == torch.int64
    assert s["state"].shape[0] == min(4, len(buf))
    np.random.seed(0)
    s2 = buf.sample(4, time_major=True)
    assert torch.equal(s2["state"].transpose(0, 1), s["state"])
    assert buf.sample(0) is None
    return buf


def check_rollout_store_roundtrip(device, lib, n=7, seed=0):
    """store_rollout (time-major device buffers of many episodes) -> gather returns them."""
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    args = types.SimpleNamespace(buffer_size=10, episode_limit=5, n_actions=5, n_agents=2, state_shape=8, obs_shape=8,
                                 rnn_hidden_dim=64, use_cuda=True, device=device)
    buf = EpisodeReplayBuffer(args, device=device, _lib=lib)
    gen = torch.Generator().manual_seed(seed)
    T, Nn = 5, 2
    def mk(n_eps):
        return {"state": torch.randn(T + 1, n_eps, 8, generator=gen), "obs": torch.randn(T + 1, n_eps, Nn, 8, generator=gen),
                "actions_discrete": torch.randint(0, 5, (T, n_eps, Nn, 1), generator=gen, dtype=torch.int32),
                "actions_continuous": torch.rand(T, n_eps, Nn, 1, generator=gen),
                "avail_actions": torch.randint(0, 2, (T + 1, n_eps, Nn, 5), generator=gen, dtype=torch.uint8),
                "reward": torch.randn(T, n_eps, 1, generator=gen),
                "terminated": torch.randint(0, 2, (T, n_eps, 1), generator=gen, dtype=torch.uint8),
                "hidden_state": torch.randn(T + 1, n_eps, Nn, 64, generator=gen)}
    first, second = mk(n), mk(n)
    buf.store_rollout({k: v.to(device) for k, v in first.items()})
    assert (buf.current_index, buf.current_size) == (7, 7)
    buf.store_rollout({k: v.to(device) for k, v in second.items()})          # wraps: slots 7,8,9,0,1,2,3
    assert (buf.current_index, buf.current_size) == (4, 10)
    out = buf.gather(np.array([8, 0, 4, 6]))
    src = [(second, 1), (second, 3), (first, 4), (first, 6)]
    for k in first:
        for b, (tr, e) in enumerate(src):
            assert torch.equal(out[k][b].cpu(), tr[k][:, e]), (k, b)
    assert bool(out["filled"].all())


def check_shared_obs_replay(device, lib, n_agents=2, seed=1):
    """shared_obs=True (the ring keeps no obs; sample() rebuilds it from the state inside the gather launch)
    returns exactly what the full-format ring returns when obs is the replicated state -- both layouts,
    both store paths, more agents than one launch has descriptors for."""
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    T, Nn, S, A, H = 5, n_agents, 12, 5, 64
    args = types.SimpleNamespace(buffer_size=9, episode_limit=T, n_actions=A, n_agents=Nn, state_shape=S, obs_shape=S,
                                 rnn_hidden_dim=H, use_cuda=True, device=device)
    full = EpisodeReplayBuffer(args, device=device, _lib=lib)
    lean = EpisodeReplayBuffer(args, device=device, _lib=lib, shared_obs=True)
    assert "obs" not in lean.buffers and full.bytes_per_episode() - lean.bytes_per_episode() == (T + 1) * Nn * S * 4
    gen = torch.Generator().manual_seed(seed)
    def mk(n_eps):
        st = torch.randn(T + 1, n_eps, S, generator=gen)
        return {"state": st, "obs": st[:, :, None, :].expand(T + 1, n_eps, Nn, S).contiguous(),
                "actions_discrete": torch.randint(0, A, (T, n_eps, Nn, 1), generator=gen, dtype=torch.int32),
                "actions_continuous": torch.rand(T, n_eps, Nn, 1, generator=gen),
                "avail_actions": torch.randint(0, 2, (T + 1, n_eps, Nn, A), generator=gen, dtype=torch.uint8),
                "reward": torch.randn(T, n_eps, 1, generator=gen),
                "terminated": torch.randint(0, 2, (T, n_eps, 1), generator=gen, dtype=torch.uint8),
                "hidden_state": torch.randn(T + 1, n_eps, Nn, H, generator=gen)}
    for n_eps in (6, 5):                                   # the second store wraps the ring
        tr = {k: v.to(device) for k, v in mk(n_eps).items()}
        full.store_rollout(tr)
        lean.store_rollout(tr)
    # one short host-side episode through the reference's store path (padding rules)
    L = 3
    ep = mk(1)
    host = {k: [v[: (L + 1 if k in ("state", "obs", "avail_actions", "hidden_state") else L), 0].numpy()] for k, v in ep.items()}
    full.store_episode(host)
    lean.store_episode(host)
    assert (full.current_index, full.current_size) == (lean.current_index, lean.current_size)
    for time_major in (False, True):
        for idx in (np.array([2, 0, 8, 3]), np.array([3])):      # slot 3 holds the short episode
            a, b = full.gather(idx, time_major=time_major), lean.gather(idx, time_major=time_major)
            assert list(a) == list(b)
            for k in a:
                assert (a[k] == b[k]) if k == "max_seq_len" else (a[k].shape == b[k].shape and torch.equal(a[k], b[k])), k
    np.random.seed(4)
    sa = full.sample(4)
    np.random.seed(4)
    sb = lean.sample(4)
    assert sb["obs"].dtype == torch.float32 and torch.equal(sa["obs"], sb["obs"]) and sb["avail_actions"].dtype == torch.int64
    bad = types.SimpleNamespace(**{**vars(args), "obs_shape": S + 1})
    with pytest.raises(ValueError):
        EpisodeReplayBuffer(bad, device=device, _lib=lib, shared_obs=True)


def check_hidden_bf16_replay(device, lib, seed=2):
    """hidden_bf16=True: the ring keeps hidden_state as bfloat16 (the store / gather launches convert, round to nearest
    even); sample() still returns float32 with every other key bit-identical to the full-format ring, and the states
    within the stated bound of the option: what torch's own float32 -> bfloat16 cast gives (relative error <= 2^-9).
    Both store paths, both layouts, a width that is not a multiple of four (scalar conversion path)."""
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    for H in (64, 67):
        T, Nn, S, A = 5, 2, 12, 5
        args = types.SimpleNamespace(buffer_size=9, episode_limit=T, n_actions=A, n_agents=Nn, state_shape=S, obs_shape=S,
                                     rnn_hidden_dim=H, use_cuda=True, device=device)
        full = EpisodeReplayBuffer(args, device=device, _lib=lib)
        lean = EpisodeReplayBuffer(args, device=device, _lib=lib, hidden_bf16=True, shared_obs=True)
        assert lean.buffers["hidden_state"].dtype == torch.bfloat16
        assert full.bytes_per_episode() - lean.bytes_per_episode() == (T + 1) * Nn * (H * 2 + S * 4)
        gen = torch.Generator().manual_seed(seed)
        def mk(n_eps):
            st = torch.randn(T + 1, n_eps, S, generator=gen)
            hs = torch.randn(T + 1, n_eps, Nn, H, generator=gen) * 3.0
            hs[0, 0, 0, :4] = torch.tensor([0.0, -0.0, 1e-30, 65504.0])
            return {"state": st, "obs": st[:, :, None, :].expand(T + 1, n_eps, Nn, S).contiguous(),
                    "actions_discrete": torch.randint(0, A, (T, n_eps, Nn, 1), generator=gen, dtype=torch.int32),
                    "actions_continuous": torch.rand(T, n_eps, Nn, 1, generator=gen),
                    "avail_actions": torch.randint(0, 2, (T + 1, n_eps, Nn, A), generator=gen, dtype=torch.uint8),
                    "reward": torch.randn(T, n_eps, 1, generator=gen),
                    "terminated": torch.randint(0, 2, (T, n_eps, 1), generator=gen, dtype=torch.uint8),
                    "hidden_state": hs}
        for n_eps in (6, 5):                                   # the second store wraps the ring
            tr = {k: v.to(device) for k, v in mk(n_eps).items()}
            full.store_rollout(tr)
            lean.store_rollout(tr)
        L = 3
        ep = mk(1)
        host = {k: [v[: (L + 1 if k in ("state", "obs", "avail_actions", "hidden_state") else L), 0].numpy()] for k, v in ep.items()}
        full.store_episode(host)
        lean.store_episode(host)
        for time_major in (False, True):
            for idx in (np.array([2, 0, 8, 3]), np.array([3])):
                a, b = full.gather(idx, time_major=time_major), lean.gather(idx, time_major=time_major)
                assert list(a) == list(b)
                for k in a:
                    if k == "max_seq_len":
                        assert a[k] == b[k]
                    elif k == "hidden_state":
                        assert b[k].dtype == torch.float32 and b[k].shape == a[k].shape
                        assert torch.equal(b[k], a[k].to(torch.bfloat16).to(torch.float32)), "round to nearest even, as torch casts"
                        assert bool(((b[k] - a[k]).abs() <= a[k].abs() * 2.0 ** -8 + 1e-37).all())
                    else:
                        assert torch.equal(a[k], b[k]), k


def check_qhead_repack(device, lib=None, O=24, A=5, H=128, AH=128):
    """After the learner's in-place Adam step only fc2_q_head is re-packed, by ONE launch (macjd_qhead_repack): the packed
    FP32 buffer, the tensor-core chunk buffer and its constant block must equal a full re-pack bit for bit, and so must
    the host-side fallback of the same re-layout."""
    import types
    from macjd_b200.core.mac import BasicMAC
    args = types.SimpleNamespace(n_agents=2, n_actions=A, rnn_hidden_dim=H, actor_hidden_dim=AH, epsilon_start=1.0,
                                 epsilon_finish=0.05, epsilon_anneal_time=1000, seed=0, agent_kernel_path=0)
    torch.manual_seed(3)
    mac = BasicMAC(O, args, _lib=lib)
    if device != "cpu":
        mac.cuda()
    pk = mac.agent.packed()
    snap = lambda: (pk.buffer.clone(), None if pk.tc_flat is None else pk.tc_flat.clone())
    with torch.no_grad():                       # what the clip+Adam kernel does: raw in-place writes, no version bump
        for p in mac.agent.fc2_q_head.parameters():
            p.data.add_(torch.randn_like(p) * 0.01)
    before = snap()
    mac.agent.packed_qhead()
    part = snap()
    assert not torch.equal(before[0], part[0])
    pk._refresh_qhead_host(mac.agent)
    host = snap()
    mac.agent.packed(force=True)
    full = snap()
    for got, name in ((part, "kernel"), (host, "host fallback")):
        assert torch.equal(got[0], full[0]), name
        assert (got[1] is None) == (full[1] is None)
        if full[1] is not None:
            assert torch.equal(got[1], full[1]), name
    return pk
