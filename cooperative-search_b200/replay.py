"""Device-resident replay ring and replay statistics.

* ``DeviceReplayBuffer`` -- common/replay_buffer.py:5-100 with the episode arrays kept in HBM: ``store_episode`` takes
  the padded episode batch ``generate_episodes`` produces on the device (no host staging), ``sample`` /
  ``sample_latest`` return device batches with the reference's keys and shapes.  The ring index arithmetic is the
  reference's ``_get_storage_idx`` (:81-99), including its wrap-around rules; rows move with the library's row
  gather / scatter kernel (cs_rows_copy).
* ``collect_replay_stats`` -- runner.py:139-172 over rollout.py:143-204 (generate_replay): every env is one replay
  (reset(init=True), greedy policy or the uniform-random one, epsilon 0), the fraction of targets found is recorded after
  every step and padded with 1.0 after the episode ends; returns what the reference prints and saves as
  ``average_res_*.npy`` -- with the env count as the number of replays instead of 100.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CoopSearchError

EPISODE_KEYS = ("o", "u", "s", "r", "o_next", "s_next", "avail_u", "avail_u_next", "u_onehot", "padded", "terminated")


def _rows_copy(lib, dst, src, dst_idx, src_idx, stream):
    """dst[dst_idx[i]] = src[src_idx[i]] over the leading axis (None = identity), in one kernel launch."""
    count = int(dst_idx.numel() if dst_idx is not None else (src_idx.numel() if src_idx is not None else src.shape[0]))
    if count == 0:
        return
    row_bytes = int(src[0].numel() * src.element_size())
    if row_bytes != int(dst[0].numel() * dst.element_size()) or not (src.is_contiguous() and dst.is_contiguous()):
        raise CoopSearchError("rows_copy needs contiguous tensors with equal row sizes")
    _lib.check(lib.cs_rows_copy(C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), C.c_uint64(row_bytes),
                                C.c_void_p(dst_idx.data_ptr()) if dst_idx is not None else None,
                                C.c_void_p(src_idx.data_ptr()) if src_idx is not None else None, count, stream), "cs_rows_copy")


class DeviceReplayBuffer:
    """ReplayBuffer(args, buffer_size) of the reference with device storage.  ``args`` needs n_actions, n_agents,
    state_shape, obs_shape, episode_limit (main.py:114-118) and optionally conv / map_size (replay_buffer.py:18-21)."""

    def __init__(self, args, buffer_size, device=None, dtypes=None):
        if not torch.cuda.is_available():
            raise CoopSearchError("coopsearch_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.args = args
        self.n_actions, self.n_agents = int(args.n_actions), int(args.n_agents)
        self.state_shape, self.obs_shape = int(args.state_shape), int(args.obs_shape)
        self.size, self.episode_limit = int(buffer_size), int(args.episode_limit)
        self.current_idx = 0
        self.current_size = 0
        obs_shape = self.obs_shape + (int(args.map_size) ** 2 if getattr(args, "conv", False) else 0)      # replay_buffer.py:18-21
        S, T, n, A = self.size, self.episode_limit, self.n_agents, self.n_actions
        shapes = {"o": (S, T, n, obs_shape), "u": (S, T, n, 1), "s": (S, T, self.state_shape), "r": (S, T, 1),
                  "o_next": (S, T, n, obs_shape), "s_next": (S, T, self.state_shape), "avail_u": (S, T, n, A),
                  "avail_u_next": (S, T, n, A), "u_onehot": (S, T, n, A), "padded": (S, T, 1), "terminated": (S, T, 1)}
        # float32 where the reference holds real values, uint8 for the 0/1 and action arrays (what generate_episodes writes)
        kinds = {"o": torch.float32, "s": torch.float32, "r": torch.float32, "o_next": torch.float32, "s_next": torch.float32}
        kinds.update(dtypes or {})
        self.buffers = {k: torch.empty(shapes[k], dtype=kinds.get(k, torch.uint8), device=self.device) for k in EPISODE_KEYS}
        print('Init ReplayBuffer({})'.format(self.size))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _get_storage_idx(self, inc=None):
        """replay_buffer.py:81-99, index arithmetic restated (always returns an array)."""
        inc = inc or 1
        if inc > self.size:
            raise CoopSearchError("episode batch of %d does not fit a buffer of %d" % (inc, self.size))
        if self.current_idx + inc <= self.size:
            idx = np.arange(self.current_idx, self.current_idx + inc)
            self.current_idx += inc
        elif self.current_idx < self.size:
            overflow = inc - (self.size - self.current_idx)
            idx = np.concatenate([np.arange(self.current_idx, self.size), np.arange(0, overflow)])
            self.current_idx = overflow
        else:
            idx = np.arange(0, inc)
            self.current_idx = inc
        self.current_size = min(self.size, self.current_size + inc)
        return idx

    def store_episode(self, episode_batch):
        """episode_batch: dict with the 11 keys, leading axis = episodes (generate_episodes' output, device tensors; numpy
        arrays of the reference's RolloutWorker are accepted and uploaded)."""
        batch = int(episode_batch["o"].shape[0])
        idxs = torch.as_tensor(self._get_storage_idx(inc=batch), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            for k in EPISODE_KEYS:
                src = torch.as_tensor(episode_batch[k], device=self.device).to(self.buffers[k].dtype).reshape(
                    (batch,) + tuple(self.buffers[k].shape[1:])).contiguous()
                _rows_copy(self.lib, self.buffers[k], src, idxs, None, self._stream())

    def can_sample(self, batch_size):
        return self.current_size >= batch_size

    def _gather(self, idx):
        out = {}
        with torch.cuda.device(self.device):
            for k in EPISODE_KEYS:
                dst = torch.empty((idx.numel(),) + tuple(self.buffers[k].shape[1:]), dtype=self.buffers[k].dtype, device=self.device)
                _rows_copy(self.lib, dst, self.buffers[k], None, idx, self._stream())
                out[k] = dst
        return out

    def sample(self, batch_size, generator=None):
        """replay_buffer.py:63-68: batch_size episodes drawn uniformly WITH replacement from the filled part."""
        idx = torch.randint(0, self.current_size, (int(batch_size),), device=self.device, generator=generator, dtype=torch.int64)
        return self._gather(idx)

    def sample_latest(self, batch_size):
        """replay_buffer.py:70-79."""
        assert self.can_sample(batch_size)
        if self.current_idx >= batch_size:
            idx = list(range(self.current_idx - batch_size, self.current_idx))
        else:
            left = batch_size - self.current_idx
            idx = list(range(self.current_size - left, self.current_size)) + list(range(self.current_idx))
        return self._gather(torch.as_tensor(idx, dtype=torch.int64, device=self.device))


def collect_replay_stats(env, agents=None, generator=None):
    """runner.collect_experiment_data (runner.py:139-172) with every env of ``env`` as one replay
    (rollout.generate_replay, rollout.py:143-204): reset(init=True), greedy actions of ``agents`` (a BatchedRNNAgents;
    None = the uniform-random policy of alg=random), per step the fraction of targets found, padded with 1.0 once the
    env has terminated.  ``env`` must not auto-reset.  Returns a dict: average_tgt_find, average_rew, average_step,
    average_res (the [episode_limit] curve in percent = average_res_*.npy) and replays."""
    if env.auto_reset:
        raise CoopSearchError("collect_replay_stats needs auto_reset=False (a finished replay is not stepped again)")
    E, T, m, dev = env.num_envs, env.time_limit, env.target_num, env.device
    env.reset(init=True)
    if agents is not None:
        agents.init_hidden()
    alive = torch.ones(E, dtype=torch.bool, device=dev)
    reward_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    steps = torch.zeros(E, dtype=torch.int64, device=dev)
    curve = torch.empty(T, dtype=torch.float64, device=dev)
    is_flight = getattr(env, "VARIANT", 0) == 1
    for t in range(T):
        if agents is None:
            actions = torch.randint(0, env.n_actions, (E, env.n_agents), dtype=torch.uint8, device=dev, generator=generator)
        elif is_flight:
            actions = agents.choose_actions(env.get_obs(full=False), epsilon=0.0, evaluate=True, env=env if agents.conv else None)
        else:
            actions = agents.choose_actions(env.get_obs(), epsilon=0.0, evaluate=True)
        reward, terminated, _ = env.step(actions)                      # masked no-op on finished envs
        reward_sum += torch.where(alive, reward.to(torch.float64), torch.zeros((), dtype=torch.float64, device=dev))
        steps += alive.to(torch.int64)
        frac = env.target_find.to(torch.float64) / float(m)
        curve[t] = torch.where(alive, frac, torch.ones((), dtype=torch.float64, device=dev)).mean()   # rollout.py:190-198
        alive = alive & (terminated == 0)
    return {"average_tgt_find": float(env.target_find.to(torch.float64).mean()), "average_rew": float(reward_sum.mean()),
            "average_step": float(steps.to(torch.float64).mean()), "average_res": (curve * 100.0).cpu().numpy(), "replays": E}
