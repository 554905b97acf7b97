"""Agent / controller parity checks shared by the host-emulation (CPU) and the GPU tests."""
import json
import os
import types

import numpy as np
import torch

from oracle import agent_oracle as AO
from tests.helpers import GOLDEN

Q_RTOL, Q_ATOL = 1e-5, 2e-6     # FP32 SIMT GEMMs vs eager torch: summation-order noise only


def load_agent_golden(name):
    g = np.load(os.path.join(GOLDEN, f"agent_{name}.npz"))
    args = types.SimpleNamespace(**json.loads(str(g["args_json"])))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    return g, args, sd


def make_mac(args, sd, device, lib=None):
    from macjd_b200.core.mac import BasicMAC
    mac = BasicMAC(args.obs_shape, args, _lib=lib)
    mac.agent.load_state_dict(sd)
    if device != "cpu":
        mac.cuda()
    return mac


def assert_as_accurate(actual, ref32, ref64, what, k=4.0, rtol=1e-5, atol=2e-6):
    """FP32 results depend on summation order when the pre-activations are large (real
    observations carry values of several hundred).  The bar: the kernel is within rtol of
    the float64 truth, or no further from it than k x the eager-FP32 oracle's own worst
    rounding error on the same input."""
    ref64 = np.asarray(ref64, dtype=np.float64)
    err_ref = np.abs(np.asarray(ref32, dtype=np.float64) - ref64).max()
    err = np.abs(np.asarray(actual, dtype=np.float64) - ref64)
    bound = np.maximum(rtol * np.abs(ref64) + atol, k * err_ref)
    bad = err > bound
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} elements off; max err {err.max():.3e}, "
                           f"oracle-fp32 max err {err_ref:.3e}")


def argmax_margin(q):
    """Gap between the best and the second best finite Q per row."""
    s = np.sort(np.where(np.isfinite(q), q, -1e30), axis=-1)
    return s[..., -1] - s[..., -2]


def check_mac_against_golden(name, device, lib=None):
    """select_actions driven by the reference's recorded inputs and injected selector draws:
    actions bit-exact, hidden / power within FP32 tolerance."""
    g, args, sd = load_agent_golden(name)
    mac = make_mac(args, sd, device, lib)
    assert sum(p.numel() for p in mac.parameters()) == int(g["n_params"])
    dev = mac.device
    steps, B, Nn, O = g["obs"].shape
    mac.hidden_states = torch.from_numpy(g["h0"].copy()).to(dev)
    for t in range(steps):
        obs = torch.from_numpy(g["obs"][t]).to(dev)
        avail = torch.from_numpy(g["avail"][t]).to(dev)
        h_before = mac.hidden_states.clone()
        kw = dict(u_eps=torch.from_numpy(g["u"][t]), rand_actions=torch.from_numpy(g["rand_actions"][t]))
        a_test, p_test = mac.select_actions(obs, avail, int(g["t_env"][t]), test_mode=True, **kw)
        mac.hidden_states.copy_(h_before)
        a, p = mac.select_actions(obs, avail, int(g["t_env"][t]), test_mode=False, **kw)
        assert a.dtype == torch.int64 and a.shape == (B, Nn, 1) and p.shape == (B, Nn, 1)
        assert mac.action_selector.epsilon == float(g["eps"][t])
        # the reference's own Q margins show which argmaxes are numerically decidable
        qm = np.where(g["avail"][t].reshape(B * Nn, -1) != 0, g["q"][t], -np.inf)
        decidable = (argmax_margin(qm) > 1e-5).reshape(B, Nn, 1)
        assert decidable.mean() > 0.9
        picked_random = (g["u"][t] < np.float32(g["eps"][t]))[..., None]
        np.testing.assert_array_equal(a.cpu().numpy()[decidable | picked_random], g["actions"][t][decidable | picked_random])
        np.testing.assert_array_equal(a_test.cpu().numpy()[decidable], g["actions_test"][t][decidable])
        same = a.cpu().numpy() == g["actions"][t]
        np.testing.assert_allclose(p.cpu().numpy()[same], g["power"][t][same], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(mac.hidden_states.cpu().numpy(), g["hidden"][t], rtol=1e-5, atol=2e-6)
        # keep both sides on the reference trajectory
        mac.hidden_states.copy_(torch.from_numpy(g["hidden"][t]))


def check_agent_outputs_against_golden(name, device, lib=None):
    """Q for every action, actor parameters and hidden state against the reference."""
    g, args, sd = load_agent_golden(name)
    mac = make_mac(args, sd, device, lib)
    dev = mac.device
    steps, B, Nn, O = g["obs"].shape
    h = torch.from_numpy(g["h0"].copy()).to(dev)
    for t in range(steps):
        obs = torch.from_numpy(g["obs"][t]).reshape(B * Nn, O).to(dev)
        out = mac.agent.run(obs, h, want_q=True, want_params=True, want_greedy=True)
        np.testing.assert_allclose(out["q_all"][0].cpu().numpy(), g["q"][t], rtol=Q_RTOL, atol=Q_ATOL)
        np.testing.assert_allclose(out["params_all"][0].cpu().numpy(), g["params"][t], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(h.cpu().numpy(), g["hidden"][t], rtol=1e-5, atol=2e-6)
        dec = argmax_margin(g["q"][t]) > 1e-5
        np.testing.assert_array_equal(out["greedy"][0].cpu().numpy()[dec], g["q"][t].argmax(-1)[dec])
        h.copy_(torch.from_numpy(g["hidden"][t]))
        # the reference-API methods
        h_in = torch.from_numpy(g["h0"] if t == 0 else g["hidden"][t - 1]).to(dev)
        h_out, params = mac.forward(obs, h_in)
        np.testing.assert_allclose(h_out.cpu().numpy(), g["hidden"][t], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(params.cpu().numpy(), g["params"][t], rtol=1e-5, atol=1e-6)


def random_agent(seed, O, A, H, AH, Nn, device, lib=None):
    from macjd_b200.core.mac import BasicMAC
    args = types.SimpleNamespace(n_agents=Nn, n_actions=A, rnn_hidden_dim=H, actor_hidden_dim=AH, obs_shape=O,
                                 epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=1000, seed=seed)
    torch.manual_seed(seed)
    mac = BasicMAC(O, args, _lib=lib)
    if device != "cpu":
        mac.cuda()
    return mac, args


def check_unroll_against_oracle(device, lib=None, O=24, A=5, H=128, AH=128, Nn=2, B=19, T=5, seed=3, tile_rows=0, path=None,
                                k=4.0, rtol=1e-5):
    """T-step unroll inside one launch (learner mode): all-action Q, unmasked argmax, gather of
    given actions and hidden sequence against the eager oracle; ragged row count."""
    mac, args = random_agent(seed, O, A, H, AH, Nn, device, lib)
    dev = mac.device
    rng = np.random.default_rng(seed)
    M = B * Nn
    obs = (rng.standard_normal((T, M, O)) * rng.choice([1.0, 30.0], size=O)).astype(np.float32)
    sel = rng.integers(0, A, size=(T, M)).astype(np.int32)
    sd = {k: v.detach().cpu() for k, v in mac.agent.state_dict().items()}
    out = mac.agent.run(torch.from_numpy(obs).to(dev), None, n_steps=T, zero_init=True, want_q=True, want_greedy=True,
                        sel_actions=torch.from_numpy(sel), want_hidden_seq=True, want_params=True, tile_rows=tile_rows, path=path)
    sd64 = AO.cast_sd(sd, torch.float64)
    h = torch.zeros(M, H)
    h64 = torch.zeros(M, H, dtype=torch.float64)
    for t in range(T):
        x = torch.from_numpy(obs[t])
        h = AO.agent_hidden(sd, x, h)
        params = AO.actor_params(sd, x)
        q = AO.q_all_actions(sd, h, params).numpy()
        h64 = AO.agent_hidden(sd64, x.double(), h64)
        params64 = AO.actor_params(sd64, x.double())
        q64 = AO.q_all_actions(sd64, h64, params64).numpy()
        assert_as_accurate(out["hidden_seq"][t].cpu().numpy(), h.numpy(), h64.numpy(), f"h t={t}", k=k, rtol=rtol)
        assert_as_accurate(out["q_all"][t].cpu().numpy(), q, q64, f"q t={t}", k=k, rtol=rtol)
        assert_as_accurate(out["params_all"][t].cpu().numpy(), params.numpy(), params64.numpy(), f"P t={t}", atol=1e-6)
        dec = argmax_margin(q64) > 1e-4
        assert dec.mean() > 0.9
        np.testing.assert_array_equal(out["greedy"][t].cpu().numpy()[dec], q64.argmax(-1)[dec])
        assert_as_accurate(out["q_sel"][t].cpu().numpy(), q[np.arange(M), sel[t]], q64[np.arange(M), sel[t]], f"q_sel t={t}", k=k, rtol=rtol)
    assert_as_accurate(out["hidden"].cpu().numpy(), h.numpy(), h64.numpy(), "final hidden", k=k, rtol=rtol)


def check_recurrence_rows_against_float64(device, lib=None, M=11, T=7, seed=5, path=0, with_initial_state=False, rtol=1e-5):
    """The recurrence launch of a time-unrolled pass (macjd_agent_forward, io.part == 4) on given input products:
    gate_x [T, M, 3, H] in, every step's hidden state out, against a float64 GRU (core/networks.py:101-129 unrolled
    as core/qmix.py:129-147 does).  Few rows run on csrc/gru_rec_rows.cuh (rows split over CTAs, FP32 SIMT); the
    bound is the FP32 one."""
    from macjd_b200 import _native as N
    O, A, H = 24, 5, 128
    mac, args = random_agent(seed, O, A, H, 128, 2, device, lib)
    agent, dev = mac.agent, mac.device
    rng = np.random.default_rng(seed)
    gx = (rng.standard_normal((T, M, 3, H)) * 1.5).astype(np.float32)
    h0 = (rng.standard_normal((M, H)) * 0.5).astype(np.float32) if with_initial_state else np.zeros((M, H), np.float32)
    gx_t = torch.from_numpy(gx).to(dev)
    hidden = torch.from_numpy(h0.copy()).to(dev)
    hs = torch.full((T, M, H), float("nan"), dtype=torch.float32, device=dev)
    io = N.AgentIO(n_rows=M, n_steps=T, hidden=N.ptr(hidden), hidden_zero_init=0 if with_initial_state else 1,
                   hidden_seq=N.ptr(hs), gate_x=N.ptr(gx_t), part=4, path=path, test_mode=1, tile_rows=0)
    agent.lib().call("macjd_agent_forward", agent._ctx(), agent.packed().cstruct(), io)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    sd = {k: v.detach().cpu().double().numpy() for k, v in agent.state_dict().items()}
    w_hh, b_ih, b_hh = sd["rnn.weight_hh"], sd["rnn.bias_ih"], sd["rnn.bias_hh"]
    h = h0.astype(np.float64)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    for t in range(T):
        gh = h @ w_hh.T + b_hh
        g = gx[t].astype(np.float64)
        r = sig(g[:, 0] + b_ih[:H] + gh[:, :H])
        z = sig(g[:, 1] + b_ih[H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(g[:, 2] + b_ih[2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        np.testing.assert_allclose(hs[t].cpu().numpy(), h, rtol=rtol, atol=2e-6 * (t + 1), err_msg=f"h t={t}")
    np.testing.assert_array_equal(hidden.cpu().numpy(), hs[T - 1].cpu().numpy())


def check_device_rng_selection(device, lib=None):
    """Without injected draws the kernel's Philox stream decides: reproduce it with the
    oracle's Philox and check the epsilon test and the uniform-over-available draw."""
    from oracle.env_oracle import philox4x32_10, u01_from_u32
    mac, args = random_agent(11, 24, 5, 64, 64, 2, device, lib)
    dev = mac.device
    rng = np.random.default_rng(0)
    B = 40
    obs = torch.from_numpy(rng.standard_normal((B, 2, 24)).astype(np.float32)).to(dev)
    avail = (rng.random((B, 2, 5)) < 0.6).astype(np.int64)
    avail[..., 2] = 1
    a_greedy, _ = mac.select_actions(obs, torch.from_numpy(avail).to(dev), 0, test_mode=True)
    mac.init_hidden(B)
    mac._rng_step = 0
    a, _ = mac.select_actions(obs, torch.from_numpy(avail).to(dev), 500, test_mode=False)   # eps = 0.525
    eps = np.float32(mac.action_selector.epsilon)
    key = (args.seed & 0xFFFFFFFF, args.seed >> 32)
    for m in range(B * 2):
        u = u01_from_u32(philox4x32_10((0, 1, m, 2), key)[0])
        if u < eps:
            u2 = u01_from_u32(philox4x32_10((0, 1, m, 3), key)[0])
            av = np.flatnonzero(avail.reshape(-1, 5)[m])
            kth = min(len(av) - 1, int(np.float32(u2) * np.float32(len(av))))
            assert int(a.reshape(-1)[m]) == av[kth], m
        else:
            assert int(a.reshape(-1)[m]) == int(a_greedy.reshape(-1)[m]), m
    assert 0.3 < float((a != a_greedy).float().mean()) < 0.7
