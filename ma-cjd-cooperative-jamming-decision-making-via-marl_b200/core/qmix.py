"""QMix learner on CUDA kernels (drop-in for core/qmix.py:25-334).

``train(batch, train_info)`` reproduces the reference step, quirks included (SURVEY facts
8 and 9): the eval / target unrolls start from a zero hidden state and only feed the
double-DQN argmax / gather (no availability mask, qmix.py:141-143); ``q_taken`` is computed
from the *stored* hidden states, so only ``fc2_q_head`` and the eval mixer ever receive
gradients; the loss is normalised by ``mask[:, :-1].sum()``; ``grad_norm`` is the pre-clip
norm; Adam runs with torch's defaults; targets are hard-copied every
``target_update_interval`` train steps.

Kernel sequence per step (all stream-ordered, no host sync until the stats are read):
  agent unroll (eval) -> agent unroll (target, gathers Q at the eval argmax)   csrc/agent_act.cuh
  target mixer fwd -> Q-head fwd on stored h -> eval mixer fwd                 csrc/learner.cuh
  TD target + masked loss -> mixer bwd -> Q-head bwd -> [NCCL all-reduce]      csrc/learner.cuh
  grad-norm + clip + Adam (in place on the nn.Parameters)                      csrc/learner_kernels.cuh
Data parallel: every rank trains on its own episodes; one flat FP32 bucket (gradients + the
loss / mask sums) is all-reduced, then every rank applies the identical update.
"""
from __future__ import annotations

import copy
import math
import os

import numpy as np
import torch
import torch.optim as optim

from .. import _native as N
from .networks import QMixer

TRAINED_AGENT_KEYS = ("fc2_q_head.0.weight", "fc2_q_head.0.bias", "fc2_q_head.2.weight", "fc2_q_head.2.bias")
BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8


class QMixLearner:
    def __init__(self, mac, args, _lib=None, process_group=None):
        self.args = args
        self.mac = mac
        self.n_agents = args.n_agents
        self.n_actions = args.n_actions
        self.state_shape = args.state_shape
        self.obs_shape = args.obs_shape
        self._lib = _lib if _lib is not None else getattr(mac, "_lib", None)
        self.process_group = process_group
        use_cuda = bool(getattr(args, "use_cuda", True)) and torch.cuda.is_available()
        self.device = torch.device(getattr(args, "device", "cuda") if use_cuda else "cpu")
        self.eval_qmix_net = QMixer(args)
        self.target_mac = copy.deepcopy(mac)
        self.target_qmix_net = QMixer(args)
        self.target_qmix_net.load_state_dict(self.eval_qmix_net.state_dict())
        if use_cuda:
            self.cuda()
        elif self._lib is None:
            raise N.MacjdError("QMixLearner runs on CUDA only (no CPU fallback)")
        self.agent_params = list(self.mac.parameters())
        self.qmix_params = list(self.eval_qmix_net.parameters())
        self.params = self.agent_params + self.qmix_params
        # kept for checkpoint-format compatibility (optimizer.pth); the update itself is a kernel
        self.optimizer = optim.Adam(params=self.params, lr=args.lr)
        self.last_target_update_step = 0
        self.train_step = 0
        self._opt_state = None
        self._ws = {}
        self._side_stream = None
        print(f"QMix Learner Initialized on device: {self.device}")

    # ------------------------------------------------------------------ plumbing
    def lib(self):
        return self._lib if self._lib is not None else N.get_lib()

    def _ctx(self):
        if self.device.type == "cuda":
            return N.torch_ctx(self.device)
        return N.Ctx(device=0, reserved=0, stream=None)

    @staticmethod
    def _param_slots(module):
        """(owner module, parameter key, qualified name) of every parameter: collected once -- walking
        named_parameters() costs ~0.1 ms per call, and a train step needs the trained tensors' addresses four
        times; looking the Parameters up through their owner modules still sees replaced ones."""
        slots = []
        for prefix, m in module.named_modules():
            for k, p_ in m._parameters.items():
                if p_ is not None:
                    slots.append((m, k, (prefix + "." if prefix else "") + k))
        return slots

    def _trainable(self):
        """(tensor, name) in flat-bucket order: the Q-head, then every mixer parameter."""
        sl = self.__dict__.get("_train_slots")
        if sl is None or sl[0] is not self.mac.agent or sl[1] is not self.eval_qmix_net:
            agent_slots = {n: (m, k) for m, k, n in self._param_slots(self.mac.agent)}
            order = [agent_slots[k] + ("agent." + k,) for k in TRAINED_AGENT_KEYS]
            order += [(m, k, "mixer." + n) for m, k, n in self._param_slots(self.eval_qmix_net)]
            sl = self._train_slots = (self.mac.agent, self.eval_qmix_net, order)
        return [(m._parameters[k], n) for m, k, n in sl[2]]

    def _ensure_opt_state(self):
        tr = self._trainable()
        dev = tr[0][0].device
        if self._opt_state is None or self._opt_state["grad"].device != dev:
            sizes = [p.numel() for p, _ in tr]
            total = sum(sizes)
            z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
            old = self._opt_state
            self._opt_state = {"sizes": sizes, "total": total, "grad": z(total + 8), "m": z(total), "v": z(total),
                               "step": 0, "scal": z(4), "qhead_offsets": [int(x) for x in np.cumsum([0] + sizes[:3])],
                               "scratch": z(self.lib().lib.macjd_opt_scratch_floats())}
            if old is not None:
                self._opt_state["m"].copy_(old["m"]); self._opt_state["v"].copy_(old["v"]); self._opt_state["step"] = old["step"]
        return self._opt_state

    def _mixer_struct(self, net=None, flat=None, offset=0):
        """macjd_mixer_params over a QMixer's parameters, or over views of a flat buffer."""
        cache = self.__dict__.setdefault("_mixer_struct_cache", {})
        if net is not None:
            ent = cache.get(id(net))
            if ent is None or ent[0] is not net:
                by_name = {n: (m, k) for m, k, n in self._param_slots(net)}
                ent = cache[id(net)] = (net, [by_name[k] for k in N.MIXER_KEYS], None, None)
            ptrs = tuple([m._parameters[k].data_ptr() for m, k in ent[1]])
            if ptrs != ent[2]:
                ent = cache[id(net)] = (net, ent[1], ptrs, N.MixerParams(**dict(zip(N.MIXER_FIELDS, ptrs))))
            return ent[3]
        key = ("flat", flat.data_ptr(), offset)
        st = cache.get(key)
        if st is None:
            sizes = {n: m._parameters[k].numel() for m, k, n in self._param_slots(self.eval_qmix_net)}
            offs, o = {}, offset
            for k in sizes:                                   # (named_parameters order: the flat bucket's order)
                offs[k] = o
                o += sizes[k]
            st = cache[key] = N.MixerParams(**{f: flat.data_ptr() + 4 * offs[k] for f, k in zip(N.MIXER_FIELDS, N.MIXER_KEYS)})
        return st

    def _workspace(self, name, n_floats, dev):
        t = self._ws.get(name)
        if t is None or t.numel() < n_floats or t.device != dev:
            t = torch.empty(max(int(n_floats), 4), dtype=torch.float32, device=dev)
            self._ws[name] = t
        return t

    # ------------------------------------------------------------------ batch handling
    def _time_major(self, batch):
        """Reference-layout sampled batch (numpy / torch, [B, T(+1), ...]) or a time-major
        batch from EpisodeReplayBuffer.sample(time_major=True) -> time-major device tensors."""
        dev = self.device
        T = int(batch["max_seq_len"])
        tm = bool(batch.get("time_major", False))

        def get(key, dtype, n_t):
            x = batch[key]
            x = torch.as_tensor(x) if not torch.is_tensor(x) else x
            x = x.to(device=dev)
            x = x[:n_t] if tm else x[:, :n_t].transpose(0, 1)
            return x.to(dtype).contiguous()

        B = (batch["state"].shape[1] if tm else batch["state"].shape[0])
        return {
            "T": T, "B": int(B),
            "state": get("state", torch.float32, T), "obs": get("obs", torch.float32, T),
            "hidden": get("hidden_state", torch.float32, max(T - 1, 0)),
            "a_d": get("actions_discrete", torch.int32, max(T - 1, 0)),
            "a_c": get("actions_continuous", torch.float32, max(T - 1, 0)),
            "reward": get("reward", torch.float32, max(T - 1, 0)),
            "terminated": get("terminated", torch.uint8, max(T - 1, 0)),
            "filled": get("filled", torch.uint8, max(T - 1, 0)),
        }

    # ------------------------------------------------------------------ the train step
    @torch.no_grad()
    def train(self, batch, train_info=None, *, lazy_stats=False, return_debug=False, check_actions=True, _bias_corr=None):
        """core/qmix.py:76-215.  Returns {loss, grad_norm, eval_qtot_avg, target_qtot_avg}.
        ``lazy_stats=True`` returns the four statistics as one device tensor (``stats_tensor``) instead of
        reading them back, and ``check_actions=False`` skips the action-range read-back (for batches whose
        actions this library's own selection kernel produced): together they leave the step without any
        host synchronisation, which is what the pipelined training loop (main.py) needs."""
        self.train_step += 1
        L, ctx, dev = self.lib(), self._ctx(), self.device
        tb = self._time_major(batch)
        T, B, Nn, A = tb["T"], tb["B"], self.n_agents, self.n_actions
        if T < 2:
            raise ValueError("QMixLearner.train needs episodes of at least 2 steps")
        M, R = B * Nn, B * (T - 1)
        S = tb["state"].shape[-1]
        H = self.args.rnn_hidden_dim
        agent, tgt_agent = self.mac.agent, self.target_mac.agent
        obs = tb["obs"].view(T, M, -1)
        # networks.py:157-158 raises on an action index outside [0, A).  Checked here, before the long
        # unrolls are queued: the read-back then only waits for the batch preparation (checked later it
        # would block the host until both unrolls finish and serialise the rest of the step behind them).
        if check_actions and int(tb["a_d"].numel()):
            lo_hi = torch.stack([tb["a_d"].min(), tb["a_d"].max()]).tolist()
            if lo_hi[0] < 0 or lo_hi[1] >= A:
                raise IndexError(f"Action index out of bounds, n_actions: {A}")

        # 1-2. unrolls from a zero hidden state (qmix.py:129-147).  The two networks are independent
        # until the double-DQN gather, and neither feeds steps 4-5 (Q of the taken actions from the
        # STORED hidden states, eval mixer), so both unrolls go to side streams -- each occupies one
        # CTA pair -- and the main stream computes 4-5 underneath them.
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        # (re-)pack both networks' weights on THIS stream before the fork: a stale pack (first step, after
        # load_models / load_state_dict) would otherwise be rebuilt on a side stream while the main stream's
        # Q-head launch below reads the same buffer
        agent.packed()
        tgt_agent.packed()
        if dev.type == "cuda":
            cur = torch.cuda.current_stream(dev)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=dev)
                self._side_stream2 = torch.cuda.Stream(device=dev)
            side, side2 = self._side_stream, self._side_stream2
            side.wait_stream(cur)
            side2.wait_stream(cur)
            with torch.cuda.stream(side):
                tg = tgt_agent.run(obs, None, n_steps=T, zero_init=True, want_q=True)
            with torch.cuda.stream(side2):
                ev = agent.run(obs, None, n_steps=T, zero_init=True, want_greedy=True)
        else:
            tg = tgt_agent.run(obs, None, n_steps=T, zero_init=True, want_q=True)
            ev = agent.run(obs, None, n_steps=T, zero_init=True, want_greedy=True)

        dims = N.MixerDims(n_rows=R, state_dim=S, n_agents=Nn, embed_dim=self.args.mixing_embed_dim,
                           hyper_hidden=self.args.hyper_hidden_dim, reserved=0)
        ws_floats = L.lib.macjd_mixer_workspace_floats(dims)
        ws = self._workspace("mixer", ws_floats, dev)              # eval mixer: intermediates kept for the backward
        ws_tgt = self._workspace("mixer_tgt", ws_floats, dev)
        tq_tot, q_tot, dq_tot, targets = f32(R), f32(R), f32(R), f32(R)
        states_next = tb["state"][1:T].reshape(R, S)
        states_cur = tb["state"][0:T - 1].reshape(R, S)

        # 4. Q(s_t, a_t, P_t) from the stored hidden states (qmix.py:161-184)
        qd = N.QheadDims(n_rows=R * Nn, hidden=H, n_actions=A, reserved=0)
        pk = agent.packed().cstruct()
        hidden = tb["hidden"].view(R * Nn, H)
        a_d, a_c = tb["a_d"].view(-1), tb["a_c"].view(-1)
        q_taken = f32(R * Nn)
        hid = self._workspace("qhead_hid", R * Nn * H, dev)
        qs = self._workspace("qhead_scratch", L.lib.macjd_qhead_scratch_floats(qd), dev)
        L.callv("macjd_qhead_forward_ws", ctx, qd, pk, hidden, a_d, a_c, q_taken, hid, qs, qs.numel())

        # 5. eval mixer (qmix.py:187); leaves its intermediates in the workspace
        eval_struct = self._mixer_struct(self.eval_qmix_net)
        L.callv("macjd_mixer_forward", ctx, dims, eval_struct, q_taken, states_cur, q_tot, ws, ws_floats)

        # 3. double-DQN gather and target mixer (qmix.py:143-151): first use of the unrolls
        if dev.type == "cuda":
            cur.wait_stream(side)
            cur.wait_stream(side2)
            tg["q_all"].record_stream(cur)
            ev["greedy"].record_stream(cur)
        tq_taken = f32(R, Nn)
        L.callv("macjd_gather_q", ctx, R * Nn, A, tg["q_all"][1:], ev["greedy"][1:], tq_taken)
        L.callv("macjd_mixer_forward", ctx, dims, self._mixer_struct(self.target_qmix_net), tq_taken, states_next,
                tq_tot, ws_tgt, ws_floats)

        # 6. TD targets and masked loss sums (qmix.py:155,191-194)
        opt = self._ensure_opt_state()
        grad, total = opt["grad"], opt["total"]
        sums = grad[total:total + 8]                       # rides in the all-reduce bucket
        sums[4:].fill_(float(R))                           # local row count (ragged shards)
        td_ws = self._workspace("td", L.lib.macjd_td_scratch_floats(R), dev)
        L.callv("macjd_td_loss", ctx, R, q_tot, tq_tot, tb["reward"].view(-1), tb["terminated"].view(-1),
                tb["filled"].view(-1), float(self.args.gamma), dq_tot, targets, sums, td_ws, td_ws.numel())

        # 7-8. backward: mixer, then the Q-head (the only agent tensors with gradients)
        qh = sum(opt["sizes"][:4])
        dq = f32(R * Nn)
        L.callv("macjd_mixer_backward", ctx, dims, eval_struct, q_taken, dq_tot, ws, ws_floats,
                self._mixer_struct(flat=grad, offset=qh), dq)
        o0, o1, o2, o3 = opt["qhead_offsets"]
        L.callv("macjd_qhead_backward", ctx, qd, pk, hidden, a_d, a_c, hid, dq,
                grad[o0:], grad[o1:], grad[o2:], grad[o3:], qs, qs.numel())

        # 9. data-parallel exchange: gradients + {sum td^2, sum mask, sum q_tot, sum targets}
        if self.process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                              and getattr(self.args, "data_parallel", False)):
            torch.distributed.all_reduce(grad, group=self.process_group)

        # 10. clip + Adam on the Q-head and the mixer (qmix.py:197-200)
        opt["step"] += 1
        tr = self._trainable()
        ptrs = tuple([p.data_ptr() for p, _ in tr])
        if opt.get("tensors_for") != ptrs:
            tensors = N.OptTensors(count=len(tr), reserved=0)
            for i, (p, _) in enumerate(tr):
                tensors.param[i] = p.data_ptr()
                tensors.numel[i] = p.numel()
            opt["tensors_for"], opt["tensors"] = ptrs, tensors
        tensors = opt["tensors"]
        debug = None
        if return_debug:
            debug = {"grad": grad[:total].clone(), "names": [n for _, n in tr], "sizes": list(opt["sizes"]),
                     "q_tot": q_tot.clone(), "targets": targets.clone(), "q_taken": q_taken.clone(),
                     "tq_taken": tq_taken.clone(), "next_actions": ev["greedy"][1:].clone(), "sums": sums.clone()}
        if _bias_corr is None:
            L.callv("macjd_clip_adam", ctx, tensors, grad, opt["m"], opt["v"], sums, float(self.args.grad_norm_clip),
                    float(self.args.lr), BETA1, BETA2, ADAM_EPS, int(opt["step"]), opt["scal"], opt["scratch"],
                    opt["scratch"].numel())
        else:      # captured step (train_sampled): the step's bias corrections are read from device memory
            L.callv("macjd_clip_adam_dev", ctx, tensors, grad, opt["m"], opt["v"], sums, float(self.args.grad_norm_clip),
                    float(self.args.lr), BETA1, BETA2, ADAM_EPS, _bias_corr, opt["scal"], opt["scratch"],
                    opt["scratch"].numel())
        agent.packed_qhead()              # the Q-head changed under the packed copy (nothing else is trained)

        # 11. hard target sync (qmix.py:203-205); a captured step leaves it to its replayer
        if _bias_corr is None and (self.train_step - self.last_target_update_step) >= self.args.target_update_interval:
            self._update_targets()
            self.last_target_update_step = self.train_step

        # 12. stats (qmix.py:209-215)
        stats_dev = torch.stack([opt["scal"][2], opt["scal"][0], sums[2] / sums[4], sums[3] / sums[4]])
        if lazy_stats:
            stats = {"stats_tensor": stats_dev}
        else:
            s = stats_dev.tolist()
            stats = {"loss": s[0], "grad_norm": s[1], "eval_qtot_avg": s[2], "target_qtot_avg": s[3]}
        if return_debug:
            stats["debug"] = debug
        return stats

    # ------------------------------------------------------------------ sample + train as one replayed graph
    MAX_STEP_GRAPHS = 4

    def _graphable(self):
        if self.device.type != "cuda" or os.environ.get("MACJD_TRAIN_GRAPH", "1") == "0":
            return False
        dist = torch.distributed
        if self.process_group is not None or (dist.is_available() and dist.is_initialized()
                                              and getattr(self.args, "data_parallel", False)):
            # data-parallel: the gradient all-reduce is captured with the step (NCCL records its kernel into the graph;
            # every rank issues one all-reduce per step whether it replays or runs eagerly, so ranks may mix the two)
            return (dist.get_backend(self.process_group) == "nccl"
                    and os.environ.get("MACJD_TRAIN_GRAPH_DP", "1") != "0")
        return True

    def train_sampled(self, buffer, batch_size, train_info=None):
        """``train(buffer.sample(batch_size, time_major=True), lazy_stats=True, check_actions=False)`` -- the learner's half
        of main.py:212-228 -- as ONE call.  Same index draws, same kernels, same results; on a single GPU the whole step
        (replay gather, both unrolls, mixers, TD loss, backward, clip + Adam, Q-head re-pack: ~90 launches on three
        streams, ~0.6 ms of host work against ~0.7 ms of GPU work at the reference batch) is captured once per
        (batch, episode length) as a CUDA graph and replayed: per step the host then draws the indices, refreshes them and
        the two Adam bias corrections in device memory, and launches the graph.  The first step of a shape runs eagerly
        (it sizes the workspaces), the second is captured; the hard target update stays on the host between replays.
        Data-parallel learners over NCCL capture the gradient all-reduce with the step (MACJD_TRAIN_GRAPH_DP=0: eager);
        other backends and MACJD_TRAIN_GRAPH=0 take the eager path.  Returns
        {"stats_tensor": [loss, grad_norm, eval_qtot_avg, target_qtot_avg]} or None when the ring is empty."""
        indices = buffer._draw_indices(batch_size)
        if indices is None:
            return None
        eager = lambda: self.train(dict(buffer.gather(indices, time_major=True), time_major=True), train_info,
                                   lazy_stats=True, check_actions=False)
        if not self._graphable():
            return eager()
        B, max_len = len(indices), int(buffer.ep_len[indices].max())
        graphs = self.__dict__.setdefault("_step_graphs", {})
        key = (id(buffer), B, max_len, self.mac.agent.path)
        g = graphs.get(key)
        if g is None:
            if max_len < 2 or len(graphs) >= self.MAX_STEP_GRAPHS:
                return eager()
            graphs[key] = "warm"             # this step sizes every workspace; the next one of this shape is captured
            return eager()
        if g == "warm":
            g = graphs[key] = self._capture_step(buffer, B, max_len)
        return self._replay_step(g, indices)

    def _capture_step(self, buffer, B, max_len):
        dev = self.device
        g = {"buffer": buffer, "idx": torch.zeros(B, dtype=torch.int32, device=dev),
             "bias_corr": torch.ones(2, dtype=torch.float32, device=dev), "graph": torch.cuda.CUDAGraph(),
             "packs": (self.mac.agent.packed().buffer.data_ptr(), self.target_mac.agent.packed().buffer.data_ptr())}
        step0, opt_step0 = self.train_step, self._ensure_opt_state()["step"]
        torch.cuda.synchronize(dev)
        # (thread-local capture mode: the NCCL watchdog thread polls its events while this thread captures)
        with torch.cuda.graph(g["graph"], capture_error_mode="thread_local"):
            batch = dict(buffer.gather(None, time_major=True, idx_dev=g["idx"], max_len=max_len), time_major=True)
            g["stats"] = self.train(batch, None, lazy_stats=True, check_actions=False, _bias_corr=g["bias_corr"])["stats_tensor"]
        # capturing recorded the launches without running them: the counters go back, the replay below is the step
        self.train_step, self._opt_state["step"] = step0, opt_step0
        g["keep"] = (dict(self._ws), batch)      # a later, larger shape may replace a workspace: this graph keeps its own alive
        return g

    def _replay_step(self, g, indices):
        self.train_step += 1
        opt = self._opt_state
        opt["step"] += 1
        step = opt["step"]
        # (copies from pageable host memory are staged by the driver before they return: the arrays may die right away)
        g["idx"].copy_(torch.from_numpy(indices.astype(np.int32)), non_blocking=True)
        b1, b2 = float(np.float32(BETA1)), float(np.float32(BETA2))      # (macjd_clip_adam takes them as floats)
        g["bias_corr"].copy_(torch.tensor([1.0 - b1 ** step, math.sqrt(1.0 - b2 ** step)], dtype=torch.float64).float(),
                             non_blocking=True)
        g["graph"].replay()
        if (self.train_step - self.last_target_update_step) >= self.args.target_update_interval:
            self._update_targets()
            self.target_mac.agent.packed()              # same buffers: the captured launches read the new weights
            self.last_target_update_step = self.train_step
        return {"stats_tensor": g["stats"].clone()}

    # ------------------------------------------------------------------ reference API
    def _update_targets(self):
        """qmix.py:282-290"""
        self.target_mac.load_state(self.mac.state_dict())
        self.target_qmix_net.load_state_dict(self.eval_qmix_net.state_dict())

    def cuda(self):
        self.mac.cuda()
        self.target_mac.cuda()
        self.eval_qmix_net.cuda()
        self.target_qmix_net.cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _export_optimizer_state(self):
        """Fill the torch.optim.Adam object with the kernel's moments so optimizer.pth keeps
        the reference's format (state only for the tensors that ever had a gradient)."""
        if self._opt_state is None or self._opt_state["step"] == 0:
            return
        o, off = self._opt_state, 0
        for (p, _), n in zip(self._trainable(), o["sizes"]):
            self.optimizer.state[p] = {"step": torch.tensor(float(o["step"])),
                                       "exp_avg": o["m"][off:off + n].view_as(p).clone(),
                                       "exp_avg_sq": o["v"][off:off + n].view_as(p).clone()}
            off += n

    def save_models(self, path):
        """qmix.py:300-315: agent.pth, qmix_net.pth, optimizer.pth"""
        os.makedirs(path, exist_ok=True)
        self.mac.save_models(path)
        torch.save(self.eval_qmix_net.state_dict(), f"{path}/qmix_net.pth")
        self._export_optimizer_state()
        torch.save(self.optimizer.state_dict(), f"{path}/optimizer.pth")

    def load_models(self, path):
        """qmix.py:317-334 (the optimizer state is not restored, as in the reference)."""
        self.mac.load_models(path)
        self.eval_qmix_net.load_state_dict(torch.load(f"{path}/qmix_net.pth", map_location=lambda storage, loc: storage))
        self._update_targets()
