"""Batched simple_spread behind the reference's env protocol (env/simple_spread.py: SimpleSpreadEnv).

``get_env_info, reset, step, get_obs, get_state, get_avail_agent_actions, close, render`` with a leading ``num_envs`` axis
and torch CUDA tensors.  All arithmetic happens in csrc/spread.cu behind the C ABI; there is no CPU path.  Differences to
the reference, as for the other envs: a finished env is a masked no-op until it is reset (or ``auto_reset=True``), and the
reset placement comes from a keyed Philox stream (shard invariant) unless a layout is injected."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CoopSearchError
from .vec_flight import _wrap


class _Handle:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().cs_spread_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class VecSimpleSpreadEnv:
    """num_envs x SimpleSpreadEnv.  ``args`` needs map_size, target_num, n_agents (simple_spread.py:19-23)."""
    ENV_NAME = "simple_spread"

    def __init__(self, args, num_envs=1, device=None, seed=0, env_id_base=0, auto_reset=False, reset=True):
        if not torch.cuda.is_available():
            raise CoopSearchError("coopsearch_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        self.args = args
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.map_size = int(args.map_size)
        self.target_num = int(args.target_num)
        self.n_agents = int(args.n_agents)
        self.target_radius, self.agent_radius = 1, 6                         # simple_spread.py:22,24
        self.time_limit = 100                                                # :25
        self.n_actions = 5                                                   # :26
        self.state_shape = self.n_agents * 2 + self.target_num * 2           # :27
        self.obs_shape = 2 + (self.n_agents - 1) * 2 + self.target_num * 4   # :28
        self.seed, self.env_id_base, self.auto_reset = int(seed), int(env_id_base), bool(auto_reset)
        cfg = _lib.SpreadCfg(struct_size=C.sizeof(_lib.SpreadCfg), num_envs=self.num_envs, n_agents=self.n_agents,
                             target_num=self.target_num, map_size=self.map_size, time_limit=self.time_limit,
                             auto_reset=int(self.auto_reset), device=self.device.index, seed=self.seed & 0xFFFFFFFF,
                             env_id_base=self.env_id_base & 0xFFFFFFFF)
        hp = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_create(C.byref(cfg), C.byref(hp)), "cs_spread_create")
        self._h = _Handle(hp.value)
        b = _lib.SpreadBuffers()
        _lib.check(self.lib.cs_spread_buffers_get(self._h.ptr, C.byref(b)), "cs_spread_buffers_get")
        E, n, m, dev, own = self.num_envs, self.n_agents, self.target_num, self.device, self._h
        pos = _wrap(b.pos, (E, n + m, 2), "<f8", dev, own)
        self.agent_xy, self.tgt_xy = pos[:, :n], pos[:, n:]                  # [E,n,2], [E,m,2] f64 views
        self._meta = _wrap(b.meta, (E, 4), "<i4", dev, own)
        self._obs = _wrap(b.obs, (E, n, int(b.obs_dim)), "<f4", dev, own)
        self._state = _wrap(b.state, (E, int(b.state_dim)), "<f4", dev, own)
        self._reward = _wrap(b.reward, (E,), "<f4", dev, own)
        self.reward64 = _wrap(b.reward64, (E,), "<f8", dev, own)
        self._terminated = _wrap(b.terminated, (E,), "|u1", dev, own)
        self.occupied = _wrap(b.occupied, (E, m), "|u1", dev, own)
        self.total_reward = _wrap(b.episode_reward, (E,), "<f8", dev, own)
        self._avail = torch.ones((E, n, self.n_actions), dtype=torch.float32, device=dev)
        self._false = torch.zeros(E, dtype=torch.uint8, device=dev)
        print('Init Env ' + getattr(args, "env", self.ENV_NAME) + ' {}a{}t x{} envs on {}'.format(self.n_agents, self.target_num, self.num_envs, self.device))
        if reset:
            self.reset()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def get_env_info(self):
        """simple_spread.py:41-46 (+ n_envs)."""
        out = (C.c_int32 * 4)()
        _lib.check(self.lib.cs_spread_env_info(self._h.ptr, out), "cs_spread_env_info")
        return {"n_actions": out[0], "state_shape": out[1], "obs_shape": out[2], "episode_limit": out[3], "n_envs": self.num_envs}

    def reset(self, init=False, mask=None, targets=None, agents=None):
        """reset(init) (simple_spread.py:48-70).  mask: [E] envs to reset (None = all).  targets [E,m,2] / agents [E,n,2]
        (float64) inject a layout instead of the keyed draw (both must be given)."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        flags = 0
        if (targets is None) != (agents is None):
            raise CoopSearchError("reset(targets=..., agents=...): give both or neither")
        if targets is not None:
            self.tgt_xy.copy_(torch.as_tensor(np.asarray(targets, dtype=np.float64), device=self.device))
            self.agent_xy.copy_(torch.as_tensor(np.asarray(agents, dtype=np.float64), device=self.device))
            flags |= _lib.CS_RESET_KEEP_TARGETS
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_reset(self._h.ptr, C.c_void_p(m.data_ptr()) if m is not None else None, flags, self._stream()),
                       "cs_spread_reset")

    def step(self, actions):
        """step(act_list) (:169-180) -> (reward [E] f32, terminated [E] u8, win [E] u8 -- always 0, the reference returns False)."""
        if torch.is_tensor(actions):
            a = actions.to(device=self.device, dtype=torch.uint8)
        else:
            h = np.asarray(actions)
            if h.size and (h.min() < 0 or h.max() >= self.n_actions):
                raise IndexError('list index out of range')                  # dpos[act] (:135-138)
            a = torch.as_tensor(h.astype(np.uint8), device=self.device)
        if a.dim() == 1 and self.num_envs == 1:
            a = a.unsqueeze(0)
        if a.dim() != 2 or a.shape[0] != self.num_envs or a.shape[1] != self.n_agents:
            raise CoopSearchError('Act num mismatch agent')                  # :131-132
        a = a.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_step(self._h.ptr, C.c_void_p(a.data_ptr()), self._stream()), "cs_spread_step")
        return self._reward, self._terminated, self._false

    def step_random(self, k=1):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_step_random(self._h.ptr, int(k), self._stream()), "cs_spread_step_random")
        return self._reward, self._terminated, self._false

    def step_host(self, actions, out):
        """HOST actions in, HOST results out: ``out`` = dict of pinned tensors reward [E] f32, terminated [E] u8, obs
        [E,n,obs_shape] f32, state [E,state_shape] f32 (any may be missing)."""
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        ptr = lambda k: C.c_void_p(out[k].data_ptr()) if out.get(k) is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_step_host(self._h.ptr, C.c_void_p(a.ctypes.data), ptr("reward"), ptr("terminated"), ptr("obs"), ptr("state"),
                                                    self._stream()), "cs_spread_step_host")
        return out

    def get_obs(self):
        return self._obs

    def get_state(self):
        return self._state

    def get_avail_agent_actions(self, agent_id):
        if agent_id >= self.n_agents:
            raise CoopSearchError('Agent id out of range')                   # :72-76
        return self._avail[:, agent_id]

    def get_avail_actions(self):
        return self._avail

    @property
    def time_step(self):
        return self._meta[:, 0]

    def stats(self):
        out = (C.c_double * _lib.CS_NUM_STATS)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_spread_stats(self._h.ptr, out, self._stream()), "cs_spread_stats")
        return dict(zip(_lib.STAT_NAMES, [float(x) for x in out]))

    def close(self):
        pass

    def render(self):
        raise CoopSearchError("render() is not part of the batched env (matplotlib rendering is out of scope)")
