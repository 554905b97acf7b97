"""Episode replay ring buffer in HBM (drop-in for utils/replay_buffer.py:12-257).

Whole episodes live in device memory, one tensor per key with the reference's key names
and shapes (``[capacity, T(+1), ...]``); ``store_episode`` takes the runner's host-side
episode dict, ``store_rollout`` takes time-major device buffers of many episodes at once,
``sample`` gathers a batch with one strided-copy kernel launch (csrc/replay.cuh).
Indices are drawn on the host exactly like the reference does
(``np.random.choice(current_size, B, replace=False)``, replay_buffer.py:178) so a seeded run
samples the same episodes.  That draw is a full permutation of the ring's population -- 0.2 ms at 4 096 episodes, 2.7 ms at
125 000, against a 0.7 ms train step -- so ``fast_sampling=True`` (opt-in; ``args.replay_fast_sampling``) draws the
same distribution (uniform, without replacement) with Floyd's algorithm in ~12 us from a ``numpy.random.Generator``
seeded off the legacy global state (``np.random.seed`` still fixes the run; the index stream differs from the reference's).

Storage format: masks and ``avail_actions`` are bytes (the reference keeps int64 masks,
replay_buffer.py:59); with ``shared_obs=True`` the ring does not hold ``obs`` at all -- in this environment
every agent observes the global state (environment.py:512-522: ``get_obs`` replicates ``get_state``), so
``sample`` rebuilds ``obs[b, t, i, :] = state[b, t, :]`` inside the same gather launch.  ``sample`` returns
the reference's keys, shapes and dtypes either way; an episode shrinks from 142.8 KB to 123.4 KB at the
default dims.  ``hidden_bf16=True`` (opt-in; ``args.replay_hidden_bf16``) stores ``hidden_state`` -- 72 % of an
episode (replay_buffer.py:66) -- as bfloat16: the store and gather launches convert (round to nearest even),
``sample`` still returns float32.  The learner computes Q of the taken actions from the stored states
(core/qmix.py:161-184), so this option changes its numbers: every state carries a relative rounding error of up
to 2^-9 (stated bound of the option; ``tests/learner_checks.py:check_hidden_bf16_replay``); 71.5 KB per episode
with both options at the default dims.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from .. import _native as N

T_PLUS_1 = ("state", "obs", "avail_actions", "hidden_state")


class EpisodeReplayBuffer:
    KEY_ORDER = ("state", "obs", "actions_discrete", "actions_continuous", "avail_actions", "reward", "terminated",
                 "filled", "hidden_state")

    def __init__(self, args, device=None, _lib=None, shared_obs=None, hidden_bf16=None, fast_sampling=None):
        self.args = args
        self.fast_sampling = bool(getattr(args, "replay_fast_sampling", False) if fast_sampling is None else fast_sampling)
        self._rng = None
        self.hidden_bf16 = bool(getattr(args, "replay_hidden_bf16", False) if hidden_bf16 is None else hidden_bf16)
        # obs is state replicated per agent: keep only the state (caller vouches; needs obs_shape == state_shape)
        self.shared_obs = bool(getattr(args, "replay_shared_obs", False) if shared_obs is None else shared_obs)
        self._lib = _lib if _lib is not None else N.get_lib()
        self.buffer_size = args.buffer_size
        self.episode_limit = args.episode_limit
        self.n_actions = args.n_actions
        self.n_agents = args.n_agents
        self.state_shape = int(np.prod(args.state_shape)) if isinstance(args.state_shape, tuple) else args.state_shape
        self.obs_shape = int(np.prod(args.obs_shape)) if isinstance(args.obs_shape, tuple) else args.obs_shape
        if device is None:
            device = getattr(args, "device", "cuda") if getattr(args, "use_cuda", True) else "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda" and _lib is None:
            raise N.MacjdError("EpisodeReplayBuffer lives in GPU memory (no CPU fallback)")
        C_, T, Nn = self.buffer_size, self.episode_limit, self.n_agents
        dev = self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        # same keys / shapes as replay_buffer.py:49-68; masks are bytes, avail is a byte mask
        self.buffers = {
            "state": z((C_, T + 1, self.state_shape), torch.float32),
            "obs": z((C_, T + 1, Nn, self.obs_shape), torch.float32),
            "actions_discrete": z((C_, T, Nn, 1), torch.int32),
            "actions_continuous": z((C_, T, Nn, 1), torch.float32),
            "avail_actions": z((C_, T + 1, Nn, self.n_actions), torch.uint8),
            "reward": z((C_, T, 1), torch.float32),
            "terminated": z((C_, T, 1), torch.uint8),
            "filled": z((C_, T, 1), torch.uint8),
            "hidden_state": z((C_, T + 1, Nn, args.rnn_hidden_dim), torch.bfloat16 if self.hidden_bf16 else torch.float32),
        }
        if self.shared_obs:
            if self.obs_shape != self.state_shape:
                raise ValueError("shared_obs needs obs_shape == state_shape (obs must be the replicated state)")
            del self.buffers["obs"]
        self.ep_len = np.zeros(C_, dtype=np.int32)        # host mirror of sum(filled)
        self.current_index = 0
        self.current_size = 0
        self.lock = threading.Lock()

    # ------------------------------------------------------------------ helpers
    def _ctx(self):
        if self.device.type == "cuda":
            return N.torch_ctx(self.device)
        return N.Ctx(device=0, reserved=0, stream=None)

    def bytes_per_episode(self):
        return sum(v[0].numel() * v.element_size() for v in self.buffers.values())

    def _get_storage_idx(self, inc=None):
        """replay_buffer.py:216-251 (ring index arithmetic)."""
        inc = inc or 1
        if inc > self.buffer_size:
            raise ValueError("Attempting to store more episodes than the buffer capacity in a single call.")
        idx = (self.current_index + np.arange(inc)) % self.buffer_size
        self.current_index = int((self.current_index + inc) % self.buffer_size) if self.current_index + inc > self.buffer_size \
            else self.current_index + inc
        self.current_size = min(self.current_size + inc, self.buffer_size)
        return idx

    def _copy(self, descs, idx_dev, n_eps, index_on_src):
        arr = (N.CopyDesc * len(descs))(*descs)
        self._lib.callv("macjd_replay_copy", self._ctx(), arr, len(descs), idx_dev, n_eps, int(index_on_src))

    # ------------------------------------------------------------------ store
    def store_episode(self, episode_batch):
        """replay_buffer.py:78-151: one host-side episode (lists holding one array per key),
        padded to the episode limit with the reference's padding values."""
        batch_size = len(episode_batch["state"])
        if batch_size != 1:
            print("Warning: EpisodeReplayBuffer expects batch_size=1 from runner")
        with self.lock:
            idx = int(self._get_storage_idx(inc=1)[0])
            ep = {k: np.asarray(v[0]) for k, v in episode_batch.items()}
            L = ep["reward"].shape[0]
            T = self.episode_limit
            for key, buf in self.buffers.items():       # (with shared_obs the episode's "obs" entry is not kept)
                n_t = buf.shape[1]
                host = np.zeros(tuple(buf.shape[1:]), dtype={torch.float32: np.float32, torch.int32: np.int32,
                                                              torch.uint8: np.uint8, torch.bfloat16: np.float32}[buf.dtype])
                if key == "filled":
                    host[:L] = 1
                elif key == "hidden_state" and key not in ep:
                    print("Warning: 'hidden_state' key not found in episode_batch data during buffer storage.")
                else:
                    src = ep[key]
                    n = L + 1 if key in T_PLUS_1 else L
                    host[:n] = src.reshape((n,) + host.shape[1:])
                    if key == "terminated":
                        host[L:] = 1                      # replay_buffer.py:150
                buf[idx].copy_(torch.from_numpy(host), non_blocking=False)      # (casts to bfloat16, round to nearest even)
            self.ep_len[idx] = L

    def store_rollout(self, traj, n_eps=None):
        """Store ``n`` full-length episodes produced on the device.  ``traj`` holds time-major
        tensors: state [T+1,n,S], obs [T+1,n,N,O], actions_discrete int32 [T,n,N],
        actions_continuous [T,n,N], avail_actions uint8 [T+1,n,N,A], reward [T,n],
        terminated uint8 [T,n], hidden_state [T+1,n,N,H]; optional filled uint8 [T,n]."""
        n = int(n_eps if n_eps is not None else traj["reward"].shape[1])
        T = self.episode_limit
        with self.lock:
            slots = self._get_storage_idx(inc=n)
            idx_dev = torch.from_numpy(slots.astype(np.int32)).to(self.device)
            descs, keep = [], []
            for key, buf in self.buffers.items():
                if key == "filled" and key not in traj:
                    buf[torch.from_numpy(slots).to(self.device)] = 1
                    continue
                src = traj[key]
                to_bf16 = buf.dtype == torch.bfloat16 and src.dtype == torch.float32      # converted by the copy launch
                assert src.is_contiguous() and (src.dtype == buf.dtype or to_bf16), key
                inner = buf[0, 0].numel() * buf.element_size()                  # ring side
                s_inner = buf[0, 0].numel() * src.element_size()                # rollout side
                n_t = buf.shape[1]
                assert src.shape[0] == n_t and src.numel() * src.element_size() == n_t * src.shape[1] * s_inner, key
                descs.append(N.CopyDesc(src=src.data_ptr(), dst=buf.data_ptr(), src_ep_stride=s_inner,
                                        src_t_stride=src.shape[1] * s_inner, dst_ep_stride=n_t * inner,
                                        dst_t_stride=inner, n_t=n_t, inner_bytes=s_inner, convert=1 if to_bf16 else 0))
                keep.append(src)
            self._copy(descs, idx_dev, n, index_on_src=False)
            self.ep_len[slots] = T

    # ------------------------------------------------------------------ sample
    def _draw_indices(self, batch_size):
        if self.current_size < batch_size:
            print(f"Warning: Sampling {batch_size} but buffer only contains {self.current_size} episodes. "
                  f"Sampling {self.current_size}.")
            batch_size = self.current_size
        if batch_size <= 0:
            print("Error: Cannot sample 0 or negative episodes.")
            return None
        if self.fast_sampling:
            if self._rng is None:        # seeded off the legacy global stream: np.random.seed() still determines the run
                self._rng = np.random.Generator(np.random.PCG64(int(np.random.randint(0, 2 ** 31 - 1))))
            return self._rng.choice(self.current_size, batch_size, replace=False)
        return np.random.choice(self.current_size, batch_size, replace=False)

    def gather(self, indices, time_major=False, *, idx_dev=None, max_len=None):
        """Gather episodes ``indices`` (host array of ring slots) into fresh device tensors.
        Reference layout [B, T(+1), ...] or, for the learner kernels, time-major [T(+1), B, ...];
        trimmed to the longest episode of the batch (replay_buffer.py:183-209).
        ``idx_dev`` (int32 device tensor) + ``max_len`` instead of ``indices``: the slots are read from device memory at
        run time and nothing is taken from the host -- the form a captured launch (CUDA graph) needs."""
        if idx_dev is None:
            indices = np.asarray(indices)
            B = len(indices)
            max_len = int(self.ep_len[indices].max()) if B else 0
            idx_dev = torch.from_numpy(indices.astype(np.int32)).to(self.device)
        else:
            B, max_len = int(idx_dev.numel()), int(max_len)
        out, descs = {}, []
        for key, buf in self.buffers.items():
            n_t = max_len + 1 if key in T_PLUS_1 else max_len
            inner_shape = tuple(buf.shape[2:])
            from_bf16 = buf.dtype == torch.bfloat16                          # the gather launch converts back to float32
            s_inner = int(np.prod(inner_shape)) * buf.element_size()         # ring side
            shape = (n_t, B) + inner_shape if time_major else (B, n_t) + inner_shape
            dst = torch.empty(shape, dtype=torch.float32 if from_bf16 else buf.dtype, device=self.device)
            inner = int(np.prod(inner_shape)) * dst.element_size()           # batch side
            out[key] = dst
            if B == 0 or n_t == 0:
                continue
            descs.append(N.CopyDesc(src=buf.data_ptr(), dst=dst.data_ptr(), src_ep_stride=buf.shape[1] * s_inner,
                                    src_t_stride=s_inner, dst_ep_stride=inner if time_major else n_t * inner,
                                    dst_t_stride=B * inner if time_major else inner, n_t=n_t, inner_bytes=s_inner,
                                    convert=2 if from_bf16 else 0))
        if self.shared_obs:
            # obs[b, t, i, :] <- state[b, t, :] for every agent i: one more strided copy per agent in the same launch
            n_t, Nn, sbuf = max_len + 1, self.n_agents, self.buffers["state"]
            row = self.state_shape * sbuf.element_size()
            inner = Nn * row
            shape = (n_t, B, Nn, self.obs_shape) if time_major else (B, n_t, Nn, self.obs_shape)
            obs = torch.empty(shape, dtype=sbuf.dtype, device=self.device)
            if B and n_t:
                for i in range(Nn):
                    descs.append(N.CopyDesc(src=sbuf.data_ptr(), dst=obs.data_ptr() + i * row, src_ep_stride=sbuf.shape[1] * row,
                                            src_t_stride=row, dst_ep_stride=inner if time_major else n_t * inner,
                                            dst_t_stride=B * inner if time_major else inner, n_t=n_t, inner_bytes=row))
            out = {k: (out[k] if k != "obs" else obs) for k in self.KEY_ORDER if k == "obs" or k in out}
        for i in range(0, len(descs), N.MAX_COPY_KEYS):
            self._copy(descs[i:i + N.MAX_COPY_KEYS], idx_dev, B, index_on_src=True)
        out["max_seq_len"] = max_len
        return out

    def sample(self, batch_size, time_major=False):
        """replay_buffer.py:153-214.  Returns the reference's dict (device tensors; masks as
        bool, avail_actions as int64 like the reference's arrays) plus ``max_seq_len``.
        ``time_major=True`` returns the raw time-major batch the learner consumes directly."""
        indices = self._draw_indices(batch_size)
        if indices is None:
            return None
        out = self.gather(indices, time_major=time_major)
        if time_major:
            out["time_major"] = True
            return out
        out["terminated"] = out["terminated"].bool()
        out["filled"] = out["filled"].bool()
        out["avail_actions"] = out["avail_actions"].to(torch.int64)
        return out

    def __len__(self):
        with self.lock:
            return self.current_size
