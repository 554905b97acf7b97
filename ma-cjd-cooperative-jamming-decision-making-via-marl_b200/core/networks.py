"""Parameter containers of the MP-DQN agent and the QMix mixer, evaluated by CUDA kernels.

Drop-in for ``core/networks.py`` of the reference: ``RNNAgent`` (:16-180) and ``QMixer``
(:182-316) keep the reference's constructor signatures, sub-module layout and therefore
its ``state_dict`` key names, shapes and -- because the layers are created in the same
order -- its random initialisation under a given ``torch.manual_seed``.  Checkpoints
(``agent.pth`` / ``qmix_net.pth``) load both ways.

The forward passes do not run through ``torch.nn``: they call the fused kernels of
libmacjd_b200.so on a packed K-major copy of the weights (``pack_agent_weights``), which is
rebuilt whenever a parameter's version counter changes.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N


def _ceil_to(x, m):
    return (x + m - 1) // m * m


class PackedAgentWeights:
    """K-major packed copy of an RNNAgent's parameters (layout: include/macjd.h,
    macjd_agent_weights).  One flat float32 buffer; every segment starts 16-byte aligned."""

    FIELDS = ("wa1t", "ba1", "wa2t", "ba2", "wa3t", "ba3", "wfc1t", "bfc1", "wrzt", "brz", "wint", "bin",
              "whnt", "bhn", "wqt", "bq1", "w1a", "w1p", "w2", "bq2", "wiht", "whht", "bgx")

    def __init__(self, agent: "RNNAgent"):
        self.O, self.H, self.AH, self.A = agent.input_shape, agent.rnn_hidden_dim, agent.actor_hidden_dim, agent.n_actions
        self.Op = _ceil_to(self.O, 32)
        if self.H % 64 or self.AH % 64 or self.H > 256 or self.AH > 256 or self.A > 64:
            raise N.MacjdError(f"unsupported agent dims H={self.H} AH={self.AH} A={self.A} "
                               "(kernels need H, AH multiples of 64 and <= 256, A <= 64)")
        O, Op, H, AH, A = self.O, self.Op, self.H, self.AH, self.A
        self.shapes = {"wa1t": (Op, AH), "ba1": (AH,), "wa2t": (AH, AH), "ba2": (AH,), "wa3t": (AH, A), "ba3": (A,),
                       "wfc1t": (Op, H), "bfc1": (H,), "wrzt": (2 * H, 2 * H), "brz": (2 * H,), "wint": (H, H),
                       "bin": (H,), "whnt": (H, H), "bhn": (H,), "wqt": (H, H), "bq1": (H,), "w1a": (A, H),
                       "w1p": (H,), "w2": (H,), "bq2": (1,), "wiht": (H, 3 * H), "whht": (H, 3 * H), "bgx": (3 * H,)}
        self.offsets, off = {}, 0
        for f in self.FIELDS:
            self.offsets[f] = off
            off += _ceil_to(int(np.prod(self.shapes[f])), 4)
        self.size = off
        self.buffer = None
        self.versions = None
        # tensor-core copy (csrc/agent_act_tc.cuh): reference width only
        self.tc_ok = (self.H == 128 and self.AH == 128)
        self.tc_buffer = None
        self.tc_flat = None
        # rnn.weight_hh packed for the one-launch recurrence of macjd_agent_unroll (csrc/gru_rec_tc2.cuh): the widths
        # the fused CTA-pair kernel does not take (rnn_hidden_dim = 256, BASELINE config 4)
        self.rec_ok = self.H in (128, 256) and not self.tc_ok
        self.rec_buffer = None
        self._cstruct = None
        self._cstruct_for = None

    def __deepcopy__(self, memo):
        # target networks deep-copy the agent: the copy re-packs from its own parameters
        new = PackedAgentWeights.__new__(PackedAgentWeights)
        new.__dict__.update({k: v for k, v in self.__dict__.items()
                             if k not in ("buffer", "tc_buffer", "tc_flat", "rec_buffer", "versions", "_cstruct", "_cstruct_for", "_slots")})
        new.buffer = new.tc_buffer = new.tc_flat = new.rec_buffer = new.versions = new._cstruct = new._cstruct_for = None
        return new

    def view(self, f):
        n = int(np.prod(self.shapes[f]))
        return self.buffer[self.offsets[f]:self.offsets[f] + n].view(*self.shapes[f])

    def refresh(self, agent, force=False):
        """Re-pack if any parameter changed since the last pack (or moved device)."""
        # the per-act fast path: no context manager, no tensor ops
        slots = self.__dict__.get("_slots")
        if slots is not None and not force and self.buffer is not None:
            v = self.versions
            i = 0
            for m, k in slots:
                p_ = m._parameters[k]
                if p_._version != v[i][0]:
                    break
                i += 1
            else:
                p0 = slots[0][0]._parameters[slots[0][1]]
                if p0.data_ptr() == v[0][1] and p0.device == self.buffer.device:
                    return self
        with torch.no_grad():
            return self._repack(agent, force)

    def _repack(self, agent, force):
        # (module, name) slots are collected once: walking agent.parameters() costs ~25 us per call, and this
        # check runs on every act; looking the Parameters up through their modules still sees replaced ones
        slots = self.__dict__.get("_slots")
        if slots is None or force:
            slots = self._slots = [(m, k) for m in agent.modules() for k in m._parameters if m._parameters[k] is not None]
        params = [m._parameters[k] for m, k in slots]
        dev = params[0].device
        versions = tuple([(p._version, p.data_ptr()) for p in params])
        if not force and self.buffer is not None and self.buffer.device == dev and versions == self.versions:
            return self
        if self.buffer is None or self.buffer.device != dev:
            self.buffer = torch.zeros(self.size, dtype=torch.float32, device=dev)
        O, H, A = self.O, self.H, self.A
        sd = {k: v.detach().float() for k, v in agent.state_dict().items()}
        self.view("wa1t")[:O].copy_(sd["actor.0.weight"].t())
        self.view("ba1").copy_(sd["actor.0.bias"])
        self.view("wa2t").copy_(sd["actor.2.weight"].t())
        self.view("ba2").copy_(sd["actor.2.bias"])
        self.view("wa3t").copy_(sd["actor.4.weight"].t())
        self.view("ba3").copy_(sd["actor.4.bias"])
        self.view("wfc1t")[:O].copy_(sd["fc1.weight"].t())
        self.view("bfc1").copy_(sd["fc1.bias"])
        wih, whh = sd["rnn.weight_ih"], sd["rnn.weight_hh"]          # [3H, H], gate blocks r, z, n
        wrz = self.view("wrzt")
        wrz[:H].copy_(wih[:2 * H].t())
        wrz[H:].copy_(whh[:2 * H].t())
        self.view("brz").copy_(sd["rnn.bias_ih"][:2 * H] + sd["rnn.bias_hh"][:2 * H])
        self.view("wint").copy_(wih[2 * H:].t())
        self.view("bin").copy_(sd["rnn.bias_ih"][2 * H:])
        self.view("whnt").copy_(whh[2 * H:].t())
        self.view("wiht").copy_(wih.t())                             # [H][3H]: one product per GRU side (macjd_agent_unroll)
        self.view("whht").copy_(whh.t())
        self.view("bhn").copy_(sd["rnn.bias_hh"][2 * H:])
        self.view("bgx")[:2 * H].copy_(self.view("brz"))             # input-side biases as one vector (macjd_agent_unroll)
        self.view("bgx")[2 * H:].copy_(self.view("bin"))
        w1 = sd["fc2_q_head.0.weight"]                               # [H, H + A + 1]
        self.view("wqt").copy_(w1[:, :H].t())
        self.view("bq1").copy_(sd["fc2_q_head.0.bias"])
        self.view("w1a").copy_(w1[:, H:H + A].t())
        self.view("w1p").copy_(w1[:, H + A])
        self.view("w2").copy_(sd["fc2_q_head.2.weight"].reshape(-1))
        self.view("bq2").copy_(sd["fc2_q_head.2.bias"])
        if self.tc_ok and dev.type == "cuda":
            kc = int(agent.lib().lib.macjd_agent_tc_chunk_k())
            if kc > 0:
                self._pack_tc(sd, kc)
        if self.rec_ok and dev.type == "cuda":
            self._pack_rec(sd["rnn.weight_hh"])
        self.versions = versions
        return self

    def _pack_rec(self, whh, kc=32):
        """include/macjd.h: macjd_agent_weights.rec_chunks -- per 128-unit block of the hidden state, per gate (r, z, n),
        the H / 32 chunks of weight_hh rows [g H + 128 b, + 128), each as its TF32 hi part then its lo part."""
        H = self.H
        blocks = [self._umma_chunks(whh[g * H + 128 * b:g * H + 128 * (b + 1)].contiguous(), kc)
                  for b in range(H // 128) for g in range(3)]
        full = torch.cat(blocks, dim=0).contiguous()                  # [3 (H / 128) (H / kc), 128 * kc]
        hi = (full.view(torch.int32) & -8192).view(torch.float32)
        new = torch.stack([hi, full - hi], dim=1)
        if self.rec_buffer is None or self.rec_buffer.shape != new.shape or self.rec_buffer.device != new.device:
            self.rec_buffer = torch.empty_like(new)                   # same address on later re-packs
        self.rec_buffer.copy_(new)

    @torch.no_grad()
    def refresh_qhead(self, agent):
        """Re-pack only what the learner trains (fc2_q_head, core/qmix.py:178): six small fields and the
        last four tensor-core chunks, in place.  The full re-pack is ~60 tensor ops per train step."""
        if self.buffer is None or self.buffer.device != agent.fc1.weight.device:
            return self.refresh(agent, force=True)
        H, A = self.H, self.A
        q0, q2 = agent.fc2_q_head[0], agent.fc2_q_head[2]
        if all(p_.dtype == torch.float32 and p_.is_contiguous() for p_ in (q0.weight, q0.bias, q2.weight, q2.bias)):
            # ONE launch (include/macjd.h: macjd_qhead_repack) instead of ~20 tensor operations per train step
            tc_chunks = tc_q_c = tc_w1a = None
            kc = stride = 0
            if self.tc_buffer is not None:
                kc = self.tc_buffer.shape[-1] // 128
                tc_chunks = self.tc_buffer[self.tc_buffer.shape[0] - H // kc:]
                nb = self._tc_big_floats()
                c = self.tc_flat[-(self.TC_CONST_FLOATS + nb):]
                tc_q_c = c[512:1024]
                if nb:
                    stride = _ceil_to(A, 8)
                    tc_w1a = c[self.TC_CONST_FLOATS + 128 * stride:]
                else:
                    stride, tc_w1a = 8, c[2432:3456]
            agent.lib().callv("macjd_qhead_repack", agent._ctx(), self.cstruct(), q0.weight.detach(), q0.bias.detach(),
                              q2.weight.detach(), q2.bias.detach(), tc_chunks, kc, tc_q_c, tc_w1a, stride)
        else:
            self._refresh_qhead_host(agent)
        slots = self.__dict__.get("_slots")
        if slots is not None:
            self.versions = tuple([(m._parameters[k]._version, m._parameters[k].data_ptr()) for m, k in slots])
        return self

    def _refresh_qhead_host(self, agent):
        """The same re-layout as tensor operations (parameters that are not contiguous float32)."""
        H, A = self.H, self.A
        w1 = agent.fc2_q_head[0].weight.detach().float()             # [H, H + A + 1]
        self.view("wqt").copy_(w1[:, :H].t())
        self.view("bq1").copy_(agent.fc2_q_head[0].bias.detach())
        self.view("w1a").copy_(w1[:, H:H + A].t())
        self.view("w1p").copy_(w1[:, H + A])
        self.view("w2").copy_(agent.fc2_q_head[2].weight.detach().reshape(-1))
        self.view("bq2").copy_(agent.fc2_q_head[2].bias.detach())
        if self.tc_buffer is not None:
            kc = self.tc_buffer.shape[-1] // 128
            full = self._umma_chunks(w1[:, :H].contiguous(), kc)     # [H / kc, 128 * kc]: the chunks the buffer ends with
            hi = (full.view(torch.int32) & -8192).view(torch.float32)
            self.tc_buffer[-full.shape[0]:, 0].copy_(hi)
            self.tc_buffer[-full.shape[0]:, 1].copy_(full - hi)
            self._pack_tc_const(qhead_only=True)

    @staticmethod
    def _umma_chunks(w, kc=16):
        """[128, K] (out, in) -> [K / kc] chunks in the K-major no-swizzle UMMA layout
        (8 x 16-byte core matrices: n/8, k/4, n%8, k%4), each followed by nothing (hi / lo are
        interleaved by the caller)."""
        n, K = w.shape
        x = w.reshape(n // 8, 8, K // kc, kc // 4, 4)          # n/8, n%8, chunk, k/4, k%4
        return x.permute(2, 0, 3, 1, 4).contiguous().reshape(K // kc, n * kc)

    def _pack_tc(self, sd, kc=16):
        """Weight chunks in the order agent_forward_tc_kernel consumes them; TF32 hi part
        (low 13 mantissa bits cleared) followed by the lo remainder, per chunk."""
        O, Op, H, A = self.O, self.Op, self.H, self.A
        dev = sd["fc1.weight"].device
        pad = lambda w: torch.cat([w, torch.zeros(w.shape[0], Op - O, device=dev)], dim=1)
        wih, whh = sd["rnn.weight_ih"], sd["rnn.weight_hh"]
        uc = lambda w: self._umma_chunks(w, kc)
        ca1, cfc1 = uc(pad(sd["actor.0.weight"])), uc(pad(sd["fc1.weight"]))
        seq = []
        for i in range(Op // kc):
            seq += [ca1[i:i + 1], cfc1[i:i + 1]]
        seq.append(uc(sd["actor.2.weight"]))
        for g in range(2):                                      # r, z: W_i{g} on xf then W_h{g} on h
            seq.append(uc(wih[g * H:(g + 1) * H]))
            seq.append(uc(whh[g * H:(g + 1) * H]))
        seq.append(uc(wih[2 * H:]))                             # W_in on xf
        seq.append(uc(whh[2 * H:]))                             # W_hn on h
        seq.append(uc(sd["fc2_q_head.0.weight"][:, :H]))
        full = torch.cat(seq, dim=0).contiguous()               # [n_chunks, 128 * kc]
        hi = (full.view(torch.int32) & -8192).view(torch.float32)
        lo = full - hi
        new = torch.stack([hi, lo], dim=1)                            # [n_chunks, 2, 128 * kc]
        n = new.numel()
        tail = self.TC_CONST_FLOATS + self._tc_big_floats()
        if self.tc_flat is None or self.tc_flat.numel() != n + tail or self.tc_flat.device != new.device:
            self.tc_flat = torch.zeros(n + tail, dtype=torch.float32, device=new.device)
        self.tc_buffer = self.tc_flat[:n].view(new.shape)             # same address on every re-pack: captured launches stay valid
        self.tc_buffer.copy_(new)
        self._pack_tc_const()

    # Per-layer vectors in the layout of csrc/agent_act_tc.cuh: TcConst, appended to the chunk buffer; the
    # CTA-pair kernel fetches the block with one bulk copy instead of ~26 scalar loads per thread.
    TC_CONST_FLOATS = 4 * 128 + 4 * 128 + 3 * 128 + 8 * 128 + 8 * 128 + 8

    def _tc_big_floats(self):
        """More than 8 actions: the per-action tables do not fit the fixed block; they follow it as
        [unit][A8] actor.4.weight^T, [unit][A8] fc2_q_head.0.weight[:, H + a], [A8] actor.4.bias (A8 = A rounded up
        to a multiple of 8, zero padded) and the kernel reads them through L1 (csrc/agent_act_tc2.cuh, kBigA)."""
        if self.A <= 8:
            return 0
        a8 = _ceil_to(self.A, 8)
        return 2 * 128 * a8 + a8

    def _pack_tc_const(self, qhead_only=False):
        H, A = self.H, self.A
        nb = self._tc_big_floats()
        c = self.tc_flat[-(self.TC_CONST_FLOATS + nb):]
        gate_b, q_c = c[:512].view(128, 4), c[512:1024].view(128, 4)
        q_c[:, 0].copy_(self.view("bq1")); q_c[:, 1].copy_(self.view("w1p")); q_c[:, 2].copy_(self.view("w2"))
        if nb:
            a8 = _ceil_to(A, 8)
            big = c[self.TC_CONST_FLOATS:]
            w3, w1a, b3 = big[:128 * a8].view(128, a8), big[128 * a8:2 * 128 * a8].view(128, a8), big[2 * 128 * a8:]
        else:
            w3, w1a, b3 = c[1408:2432].view(128, 8), c[2432:3456].view(128, 8), c[3456:3464]
        w1a[:, :A].copy_(self.view("w1a").t())                        # [unit][action] = fc2_q_head.0.weight[unit, H + action]
        if qhead_only:
            return
        brz = self.view("brz")
        gate_b[:, 0].copy_(brz[:H]); gate_b[:, 1].copy_(brz[H:]); gate_b[:, 2].copy_(self.view("bin")); gate_b[:, 3].copy_(self.view("bhn"))
        c[1024:1152].copy_(self.view("ba1")); c[1152:1280].copy_(self.view("ba2")); c[1280:1408].copy_(self.view("bfc1"))
        w3[:, :A].copy_(self.view("wa3t"))                            # [unit][action] = actor.4.weight^T
        b3[:A].copy_(self.view("ba3"))

    def cstruct(self):
        key = (self.buffer.data_ptr(), id(self.tc_buffer), id(self.rec_buffer))
        if self._cstruct is not None and self._cstruct_for == key:
            return self._cstruct
        self._cstruct = self._make_cstruct()
        self._cstruct_for = key
        return self._cstruct

    def _make_cstruct(self):
        base = self.buffer.data_ptr()
        kw = {f: base + 4 * self.offsets[f] for f in self.FIELDS}
        tc = self.tc_buffer.data_ptr() if self.tc_buffer is not None else None
        rec = self.rec_buffer.data_ptr() if self.rec_buffer is not None else None
        return N.AgentWeights(obs_dim=self.O, obs_pad=self.Op, hidden=self.H, actor_hidden=self.AH,
                              n_actions=self.A, tc_format=1 if tc else 0, tc_chunks=tc, rec_chunks=rec, **kw)


class RNNAgent(nn.Module):
    """core/networks.py:16-180.  Sub-modules exist to own the parameters (names, shapes,
    init); every forward goes through the fused CUDA kernel (csrc/agent_act.cuh)."""

    def __init__(self, input_shape, args, _lib=None):
        super().__init__()
        self.args = args
        self.n_actions = args.n_actions
        self.input_shape = int(input_shape)
        self.rnn_hidden_dim = args.rnn_hidden_dim
        self.actor_hidden_dim = args.actor_hidden_dim
        # creation order == reference order (networks.py:54-79) -> identical seeded init
        self.actor = nn.Sequential(
            nn.Linear(self.input_shape, self.actor_hidden_dim), nn.ReLU(),
            nn.Linear(self.actor_hidden_dim, self.actor_hidden_dim), nn.ReLU(),
            nn.Linear(self.actor_hidden_dim, self.n_actions), nn.Sigmoid())
        self.fc1 = nn.Linear(self.input_shape, self.rnn_hidden_dim)
        self.rnn = nn.GRUCell(self.rnn_hidden_dim, self.rnn_hidden_dim)
        self.fc2_q_head = nn.Sequential(
            nn.Linear(self.rnn_hidden_dim + self.n_actions + 1, self.rnn_hidden_dim), nn.ReLU(),
            nn.Linear(self.rnn_hidden_dim, 1))
        self._packed = PackedAgentWeights(self)
        self._lib = _lib
        # kernel path: 1 = FP32 SIMT (default: the parity-exact path), 2 / 3 = tcgen05 3xTF32 with one
        # CTA per 64 rows / with CTA pairs (needs hidden = actor_hidden = 128), 0 = let the library
        # pick a tensor-core kernel when it can
        self.path = int(getattr(args, "agent_kernel_path", 1))

    # ---- native plumbing
    def lib(self):
        return self._lib if self._lib is not None else N.get_lib()

    def packed(self, force=False):
        return self._packed.refresh(self, force=force)

    def packed_qhead(self):
        """After an in-place update of fc2_q_head alone (the learner's Adam step)."""
        return self._packed.refresh_qhead(self)

    def _ctx(self):
        dev = self.fc1.weight.device
        if dev.type == "cuda":
            return N.torch_ctx(dev)
        if self._lib is None:
            raise N.MacjdError("RNNAgent runs on CUDA only (no CPU fallback); call .cuda() first")
        return N.Ctx(device=0, reserved=0, stream=None)

    @torch.no_grad()
    def run(self, obs, hidden=None, *, n_steps=1, zero_init=False, avail=None, epsilon=0.0, test_mode=True,
            u_eps=None, rand_actions=None, seed=0, rng_step=0, select=False, want_q=False, want_params=False,
            want_greedy=False, sel_actions=None, want_hidden_seq=False, tile_rows=0, out=None, path=None,
            split_unroll=True):
        """One fused launch (two for a time-unrolled call on the CTA-pair kernel, see below).  obs float32 [T, M, O] (or [M, O]); hidden float32 [M, H] updated in
        place.  Returns a dict with the requested outputs."""
        dev = self.fc1.weight.device
        if obs.device != dev or obs.dtype != torch.float32:
            obs = obs.to(device=dev, dtype=torch.float32)
        if obs.dim() == 2:
            obs = obs.unsqueeze(0)
        obs = obs.contiguous()
        T, M, O = obs.shape
        assert O == self.input_shape and T == n_steps
        A, H = self.n_actions, self.rnn_hidden_dim
        pk = self.packed()
        given = out or {}     # caller-provided output buffers (e.g. slices of a trajectory)
        out = {}

        def mk(*shape, dtype=torch.float32, name=None):
            t = given.get(name)
            if t is not None:
                assert t.is_contiguous() and t.dtype == dtype and t.numel() == int(np.prod(shape)), name
                return t
            return torch.empty(*shape, dtype=dtype, device=dev)

        if hidden is None:
            hidden = torch.zeros(M, H, dtype=torch.float32, device=dev)
            zero_init = True
        assert hidden.is_contiguous() and hidden.shape == (M, H) and hidden.dtype == torch.float32
        out["hidden"] = hidden
        if want_hidden_seq:
            out["hidden_seq"] = mk(T, M, H, name="hidden_seq")
        if want_q:
            out["q_all"] = mk(T, M, A, name="q_all")
        if want_params:
            out["params_all"] = mk(T, M, A, name="params_all")
        if want_greedy:
            out["greedy"] = mk(T, M, dtype=torch.int32, name="greedy")
        if sel_actions is not None:
            sel_actions = sel_actions.to(device=dev, dtype=torch.int32).contiguous()
            out["q_sel"] = mk(T, M, name="q_sel")
        if select:
            out["actions"] = mk(T, M, dtype=torch.int32, name="actions")
            out["power"] = mk(T, M, name="power")
            out["q_chosen"] = mk(T, M, name="q_chosen")
        if avail is not None:
            if avail.dtype != torch.uint8 or avail.device != dev:
                avail = avail.to(device=dev).ne(0).to(torch.uint8)
            avail = avail.reshape(T, M, A).contiguous()
        if u_eps is not None:
            u_eps = u_eps.to(device=dev, dtype=torch.float32).reshape(T, M).contiguous()
        if rand_actions is not None:
            rand_actions = rand_actions.to(device=dev, dtype=torch.int32).reshape(T, M).contiguous()
        p = N.ptr
        path = int(self.path if path is None else path)
        common = dict(test_mode=int(test_mode), tile_rows=int(tile_rows), path=path, epsilon=float(epsilon),
                      rng_step=int(rng_step) & 0xFFFFFFFF, seed=int(seed) & 0xFFFFFFFFFFFFFFFF, rng_row_offset=0)
        heads = dict(q_all=p(out.get("q_all")), params_all=p(out.get("params_all")), greedy=p(out.get("greedy")),
                     sel_actions=p(sel_actions), q_sel=p(out.get("q_sel")), avail=p(avail), u_eps=p(u_eps),
                     rand_actions=p(rand_actions), actions=p(out.get("actions")), power=p(out.get("power")),
                     q_chosen=p(out.get("q_chosen")))
        wants_heads = want_q or want_params or want_greedy or sel_actions is not None or select
        if split_unroll and T > 1 and path in (0, 3) and (not select or test_mode) and self._pair_kernel_ok(pk):
            # Time-unrolled use (core/qmix.py:217-280): only h_t -> h_t+1 is sequential.  On the CTA-pair
            # kernel the call becomes three launches: (3) the GRU input products W_i* relu(fc1 obs) for all
            # T x M rows at once, (4) the recurrence proper -- per step only W_h* h and the gates --, (2) actor +
            # Q-head + arg-max for all T x M rows on the stored hidden states.  split_unroll="exact" keeps the
            # input products inside the per-step launch (part 1) and is bit-identical to the fused step; the
            # default differs from it by FP32 rounding (x and h products summed in the epilogue).
            hs = out.get("hidden_seq")
            if hs is None:
                hs = torch.empty(T, M, H, dtype=torch.float32, device=dev)
            call = lambda io: self.lib().call("macjd_agent_forward", self._ctx(), pk.cstruct(), io)
            if split_unroll == "exact":
                call(N.AgentIO(n_rows=M, n_steps=T, obs=p(obs), hidden=p(hidden), hidden_zero_init=int(zero_init),
                               hidden_seq=p(hs), part=1, **common))
            else:
                gx = torch.empty(T * M, 3 * H, dtype=torch.float32, device=dev)
                call(N.AgentIO(n_rows=T * M, n_steps=1, obs=p(obs), hidden_zero_init=1, gate_x=p(gx), part=3, **common))
                call(N.AgentIO(n_rows=M, n_steps=T, hidden=p(hidden), hidden_zero_init=int(zero_init), hidden_seq=p(hs),
                               gate_x=p(gx), part=4, **common))
            if wants_heads:
                call(N.AgentIO(n_rows=T * M, n_steps=1, obs=p(obs), hidden=p(hs), hidden_zero_init=0, part=2,
                               **common, **heads))
            return out
        io = N.AgentIO(n_rows=M, n_steps=T, obs=p(obs), hidden=p(hidden), hidden_zero_init=int(zero_init),
                       hidden_seq=p(out.get("hidden_seq")), part=0, **common, **heads)
        if split_unroll and T > 1 and path in (0, 3) and not select and not self._pair_kernel_ok(pk):
            # Widths the fused CTA-pair kernel does not take (rnn_hidden_dim = 256): the time-unrolled pass as batched
            # layers -- every layer but the recurrence is one dense product over all T x M rows -- on the tensor
            # cores where they fill its tiles (include/macjd.h: macjd_agent_unroll).  FP32-level accuracy (3xTF32).
            L = self.lib()
            n_ws = int(L.lib.macjd_agent_unroll_workspace_floats(N.C.byref(pk.cstruct()), M, T))
            ws = self.__dict__.get("_unroll_ws")
            if ws is None or ws.numel() < n_ws or ws.device != dev:
                ws = self.__dict__["_unroll_ws"] = torch.empty(max(n_ws, 4), dtype=torch.float32, device=dev)
            L.callv("macjd_agent_unroll", self._ctx(), pk.cstruct(), io, ws, ws.numel())
            return out
        self.lib().call("macjd_agent_forward", self._ctx(), pk.cstruct(), io)
        return out

    def _pair_kernel_ok(self, pk):
        ok = getattr(pk, "_pair_ok", None)
        if ok is None:
            fn = self.lib().lib.macjd_agent_pair_supported
            fn.restype = N.C.c_int
            ok = pk._pair_ok = bool(fn(N.C.byref(pk.cstruct())))
        return ok

    # ---- reference API
    def init_hidden(self):
        """networks.py:81-86"""
        return self.fc1.weight.new(1, self.rnn_hidden_dim).zero_()

    def forward(self, agent_inputs, h_in):
        """networks.py:88-114 -> new hidden state [M, H] (h_in is not modified)."""
        h = h_in.detach().to(dtype=torch.float32).contiguous().clone()
        return self.run(agent_inputs, h)["hidden"]

    def actor_forward(self, inputs):
        """networks.py:116-129 -> P for every discrete action [M, A]."""
        return self.run(inputs, None, want_params=True)["params_all"][0]


class QMixer(nn.Module):
    """core/networks.py:182-316 parameter container (LayerNorm + 4 hypernetworks).  The
    forward / backward passes are the learner kernels (core/qmix.py of this package)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.n_agents = args.n_agents
        self.state_dim = int(np.prod(args.state_shape))
        self.embed_dim = args.mixing_embed_dim
        self.hyper_hidden_dim = args.hyper_hidden_dim
        # creation order == reference order (networks.py:215-248)
        self.state_norm = nn.LayerNorm(self.state_dim)
        self.hyper_w_1 = nn.Sequential(nn.Linear(self.state_dim, self.hyper_hidden_dim), nn.ReLU(),
                                       nn.Linear(self.hyper_hidden_dim, self.n_agents * self.embed_dim))
        self.hyper_w_final = nn.Sequential(nn.Linear(self.state_dim, self.hyper_hidden_dim), nn.ReLU(),
                                           nn.Linear(self.hyper_hidden_dim, self.embed_dim))
        self.hyper_b_1 = nn.Linear(self.state_dim, self.embed_dim)
        self.V = nn.Sequential(nn.Linear(self.state_dim, self.embed_dim), nn.ReLU(), nn.Linear(self.embed_dim, 1))
