"""Batched flight_easy / flight environments behind the reference's SMAC-style env protocol.

Mirrors the interface that common/rollout.py:24-201 and main.py:114,134 of the reference call on
``FlightSearchEnvEasy`` (env/flight_env_easy.py) and ``FlightSearchEnv`` (env/flight_env.py):
``get_env_info, reset, step, get_obs, get_state, get_avail_agent_actions, target_find, close,
render`` -- same names, same argument meaning, same error behaviour -- with a leading ``num_envs``
axis on everything and torch CUDA tensors instead of numpy arrays.  All arithmetic happens in
hand-written sm_100a kernels behind the C ABI of include/coopsearch.h; torch is used for device
memory views and streams only.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CoopSearchError


class _DevView:
    """Exposes a library-owned device buffer through __cuda_array_interface__ (zero-copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}
        self._owner = owner


def _wrap(ptr, shape, typestr, device, owner):
    t = torch.as_tensor(_DevView(ptr, shape, typestr, owner), device=device)
    t._cs_owner = owner           # keep the handle alive as long as any view is
    return t


def load_targets(filename):
    """circle_dict of main.py:19-32: skip the header line, whitespace-split rows of
    ``x y deter priority dx dy`` (blank lines ignored)."""
    cols = {"x": [], "y": [], "deter": [], "priority": [], "dx": [], "dy": []}
    with open(filename, "r") as fh:
        rows = fh.readlines()[1:]
    for row in rows:
        tok = row.split()
        if not tok:
            continue
        cols["x"].append(float(tok[0])); cols["y"].append(float(tok[1])); cols["deter"].append(tok[2])
        cols["priority"].append(int(tok[3])); cols["dx"].append(float(tok[4])); cols["dy"].append(float(tok[5]))
    return cols


class _Handle:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().cs_flight_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class _VecFlightBase:
    VARIANT = 0
    ENV_NAME = "flight_easy"

    def __init__(self, args, circle_dict=None, num_envs=1, device=None, seed=0, env_id_base=0,
                 auto_reset=False, lanes_per_env=0, count_touched=False, reset=True, map_overlap=False):
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
        self.target_mode = int(args.target_mode)
        self.agent_mode = int(args.agent_mode)
        self.n_agents = int(args.n_agents)
        self.view_range = int(args.view_range)
        self.time_limit = int(args.time_limit)
        self.detect_prob = float(args.detect_prob)
        self.safe_dist = float(args.safe_dist)
        self.velocity = float(args.agent_velocity)
        self.force_dist = float(args.force_dist)
        # stored-but-unused by the dynamics, as in the reference (flight_env_easy.py:26,28)
        self.turn_limit = getattr(args, "turn_limit", None)
        self.wrong_alarm_prob = getattr(args, "wrong_alarm_prob", None)
        self.circle_dict = circle_dict
        self.n_actions = 3
        self.state_shape = self.n_agents * 4 + self.target_num * 3
        self.obs_shape = 4
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self.auto_reset = bool(auto_reset)

        cfg = _lib.FlightCfg(
            struct_size=C.sizeof(_lib.FlightCfg), num_envs=self.num_envs, n_agents=self.n_agents,
            target_num=self.target_num, map_size=self.map_size, view_range=self.view_range,
            time_limit=self.time_limit, agent_mode=self.agent_mode, target_mode=self.target_mode,
            variant=self.VARIANT, auto_reset=int(self.auto_reset), count_touched=int(count_touched),
            lanes_per_env=int(lanes_per_env), device=self.device.index,
            velocity=self.velocity, detect_prob=self.detect_prob, safe_dist=self.safe_dist,
            force_dist=self.force_dist, seed=self.seed & 0xFFFFFFFF, env_id_base=self.env_id_base & 0xFFFFFFFF,
            map_overlap=int(bool(map_overlap) and self.VARIANT == 1))
        self.map_overlap = bool(map_overlap) and self.VARIANT == 1
        hp = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_create(C.byref(cfg), C.byref(hp)), "cs_flight_create")
        self._h = _Handle(hp.value)
        if self.target_mode == 0 and circle_dict is not None:
            self._set_template(circle_dict)
        b = _lib.FlightBuffers()
        _lib.check(self.lib.cs_flight_buffers_get(self._h.ptr, C.byref(b)), "cs_flight_buffers_get")
        self._b = b
        E, n, m, M = self.num_envs, self.n_agents, self.target_num, self.map_size
        dev, own = self.device, self._h
        # the state layout is described by strides (include/coopsearch.h: record per env, or structure of arrays for
        # large handles): these are strided zero-copy views
        rows = b.dyn_doubles
        flat = _wrap(b.dyn, (rows * E,), "<f8", dev, own)
        self._dyn = flat.as_strided((E, rows), (b.dyn_env_stride, b.dyn_row_stride))              # [E,rows] f64 view
        self.agent_xy = self._dyn[:, :2 * n].unflatten(1, (n, 2))            # [E,n,2] f64 view
        self.agent_yaw = self._dyn[:, b.yaw_off:b.yaw_off + n]               # [E,n]   f64 view
        flat32 = _wrap(b.dyn, (rows * E * 2,), "<i4", dev, own)
        self._meta3 = flat32.as_strided((E, _lib.CS_META_WORDS // 2, 2), (2 * b.dyn_env_stride, 2 * b.dyn_row_stride, 1),
                                        2 * b.meta_off * b.dyn_row_stride)                        # [E,4,2] i32 view
        self.tgt_xy = _wrap(b.tgt, (2 * m * E,), "<f8", dev, own).as_strided(
            (E, m, 2), (b.tgt_env_stride, 2 * b.tgt_row_stride, b.tgt_row_stride))               # [E,m,2] f64 view
        self._obs = _wrap(b.obs, (E, n, 4), "<f4", dev, own)
        self._state = _wrap(b.state, (E, b.state_stride), "<f4", dev, own)[:, :b.state_len]   # rows padded to 16 B
        self._reward = _wrap(b.reward, (E,), "<f4", dev, own)
        self._terminated = _wrap(b.terminated, (E,), "|u1", dev, own)
        self._win = _wrap(b.win, (E,), "|u1", dev, own)
        self._target_find = _wrap(b.target_find, (E,), "<i4", dev, own)
        self._stats = _wrap(b.stats, (_lib.CS_NUM_STATS,), "<f8", dev, own)
        self._slab = _wrap(b.slab, (int(b.slab_bytes),), "|u1", dev, own)      # all step outputs (checkpointing)
        # the belief map lives on the device as 4x4-cell tiles (include/coopsearch.h): zero-copy tiled view
        self._map_tiles = _wrap(b.prob_map, (E, b.map_tiles, b.map_tiles, 4, 4), "<f4", dev, own) if b.prob_map else None
        self._avail = torch.ones((E, n, self.n_actions), dtype=torch.float32, device=dev)
        self._host = None
        self._host_c = None
        print('Init Env ' + getattr(args, "env", self.ENV_NAME) + ' {}a{}t(agent mode:{}, target mode:{}) x{} envs on {}'.format(
            self.n_agents, self.target_num, self.agent_mode, self.target_mode, self.num_envs, self.device))
        if reset:
            self.reset(init=True)     # the reference ctor ends with reset(init=True) (flight_env_easy.py:68)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _set_template(self, cd):
        m = self.target_num
        if len(cd["x"]) < m:
            raise CoopSearchError("target file has %d rows, target_num is %d" % (len(cd["x"]), m))
        rows = (C.c_double * (5 * m))()
        for j in range(m):
            rows[5 * j + 0] = cd["x"][j]; rows[5 * j + 1] = cd["y"][j]
            rows[5 * j + 2] = cd["dx"][j]; rows[5 * j + 3] = cd["dy"][j]
            rows[5 * j + 4] = 1.0 if cd["deter"][j] == "f" else 0.0
        _lib.check(self.lib.cs_flight_set_target_template(self._h.ptr, rows, m), "cs_flight_set_target_template")

    @property
    def lanes_per_env(self):
        return int(self.lib.cs_flight_lanes_per_env(self._h.ptr))

    # ------------------------------------------------------------ reference API
    def get_env_info(self):
        """flight_env_easy.py:71-77 (+ n_envs)."""
        out = (C.c_int32 * 4)()
        _lib.check(self.lib.cs_flight_env_info(self._h.ptr, out), "cs_flight_env_info")
        return {"n_actions": out[0], "state_shape": out[1], "obs_shape": out[2], "episode_limit": out[3],
                "n_envs": self.num_envs}

    def reset(self, init=False, mask=None, targets=None, keep_episode=False):
        """reset(init) (flight_env_easy.py:79-182) of every env, or of those with mask[e] != 0.

        targets: optional [E,m,2] float64 tensor/array of target coordinates to inject (parity tests:
        the reference's own MT19937-drawn layout); otherwise targets are redrawn on the device."""
        flags = _lib.CS_RESET_INIT if init else 0
        if keep_episode:
            flags |= _lib.CS_RESET_KEEP_EPISODE
        if targets is not None:
            t = torch.as_tensor(np.asarray(targets) if not torch.is_tensor(targets) else targets,
                                dtype=torch.float64, device=self.device)
            if tuple(t.shape) != (self.num_envs, self.target_num, 2):
                raise CoopSearchError("targets must have shape (num_envs, target_num, 2)")
            self.tgt_xy.copy_(t)
            flags |= _lib.CS_RESET_KEEP_TARGETS
        mptr = None
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if mask.numel() != self.num_envs:
                raise CoopSearchError("mask must have num_envs elements")
            mptr = C.c_void_p(mask.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_reset(self._h.ptr, mptr, flags, self._stream()), "cs_flight_reset")

    def _as_actions(self, actions):
        if torch.is_tensor(actions):
            a = actions.to(device=self.device, dtype=torch.uint8)
        else:
            h = np.asarray(actions)
            if h.size and (h.min() < 0 or h.max() >= self.n_actions):
                # the reference indexes dyaw[act] (flight_env_easy.py:259-262): IndexError.  Host actions are checked here;
                # device tensors are not (that would cost a synchronisation): the kernels treat a byte > 2 as "no turn"
                raise IndexError('list index out of range')
            a = torch.as_tensor(h.astype(np.uint8), device=self.device)
        if a.dim() == 1 and self.num_envs == 1:
            a = a.unsqueeze(0)
        if a.dim() != 2 or a.shape[0] != self.num_envs or a.shape[1] != self.n_agents:
            raise CoopSearchError('Act num mismatch agent')          # flight_env_easy.py:256-257
        return a.contiguous()

    def step(self, actions):
        """step(act_list) (flight_env_easy.py:303-314) -> (reward [E] f32, terminated [E] u8, win [E] u8).
        Returned tensors are views of library-owned buffers, overwritten by the next step."""
        a = self._as_actions(actions)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_step(self._h.ptr, C.c_void_p(a.data_ptr()), self._stream()), "cs_flight_step")
        return self._reward, self._terminated, self._win

    def step_random(self, k=1):
        """k steps under the uniform-random policy drawn inside the kernel (alg=random, agent/agent.py:34-36)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_step_random(self._h.ptr, int(k), self._stream()), "cs_flight_step_random")
        return self._reward, self._terminated, self._win

    def get_obs(self):
        """[E,n,4] (flight_env_easy.py:218-221)."""
        return self._obs

    def get_state(self):
        """[E,4n+3m] (flight_env_easy.py:190-216)."""
        return self._state

    def get_avail_agent_actions(self, agent_id):
        if agent_id >= self.n_agents:
            raise CoopSearchError('Agent id out of range')            # flight_env_easy.py:185-186
        return self._avail[:, agent_id]

    def get_avail_actions(self):
        return self._avail

    @property
    def target_find(self):
        return self._target_find

    @property
    def win_flag(self):
        return self._win

    def _meta_word(self, w):
        return self._meta3[:, w >> 1, w & 1]                                 # [E] int32 view of meta word w

    @property
    def meta(self):
        """[E, CS_META_WORDS] int32 (a copy; the words live in four [E][2] device rows)."""
        return self._meta3.reshape(self.num_envs, _lib.CS_META_WORDS)

    @property
    def time_step(self):
        return self._meta_word(_lib.META_TIME)

    @property
    def found_mask(self):
        return self._meta_word(_lib.META_FOUND)

    @property
    def out_mask(self):
        return self._meta_word(_lib.META_OUT)

    def close(self):
        pass

    def render(self):
        raise CoopSearchError("render() is not part of the batched hot path")

    # ------------------------------------------------------------------ extras
    def stats(self):
        """Running sums over finished episodes on this GPU (the vector all-reduced by dist.allreduce_stats)."""
        out = (C.c_double * _lib.CS_NUM_STATS)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_stats(self._h.ptr, out, self._stream()), "cs_flight_stats")
        return dict(zip(_lib.STAT_NAMES, list(out)))

    @property
    def stats_tensor(self):
        return self._stats

    def sync_map(self):
        """map_overlap=True: the current stream waits for the latest belief-map kernel (which runs on the handle's own
        stream).  The library does this itself wherever it touches the map; call it before reading ``prob_map_tiles``
        on another stream's schedule, before a CUDA-graph capture of step() calls begins, and before it ends."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_map_sync(self._h.ptr, self._stream()), "cs_flight_map_sync")

    @property
    def prob_map_tiles(self):
        """Zero-copy view of the belief map in its tiled device layout [E, tiles, tiles, 4, 4]: cell (i, j) of the
        reference's prob_map is element [e, i // 4, j // 4, i % 4, j % 4] (joined with the map stream first)."""
        if self._map_tiles is not None and self.map_overlap:
            self.sync_map()
        return self._map_tiles

    @property
    def prob_map(self):
        """self.prob_map of the reference (flight_env.py:53): [E,M,M] float32, prob_map[e,i,j] with i <-> x -- a fresh
        row-major copy of the tiled device map (cs_flight_map_export); None for flight_easy.  Assigning
        (``env.prob_map = t``) converts back into the tiled map."""
        if self.prob_map_tiles is None:
            return None
        out = torch.empty((self.num_envs, self.map_size, self.map_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_map_export(self._h.ptr, C.c_void_p(out.data_ptr()), self._stream()), "cs_flight_map_export")
        return out

    @prob_map.setter
    def prob_map(self, value):
        if self.prob_map_tiles is None:
            raise CoopSearchError("flight_easy has no probability map")
        t = torch.as_tensor(value, dtype=torch.float32, device=self.device).contiguous()
        if tuple(t.shape) != (self.num_envs, self.map_size, self.map_size):
            raise CoopSearchError("prob_map must have shape (num_envs, map_size, map_size)")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_map_import(self._h.ptr, C.c_void_p(t.data_ptr()), self._stream()), "cs_flight_map_import")

    def get_state_dict(self):
        """Checkpoint of the full env state (the reference never checkpoints env state): the
        dynamic state, the targets, the output buffers (so that get_obs / get_state / reward / target_find after a
        restore are those of the checkpointed step) and, for the flight variant, the belief map."""
        d = {"dyn": self._dyn.clone(), "tgt": self.tgt_xy.clone(), "outputs": self._slab.clone()}
        if self.prob_map_tiles is not None:
            d["prob_map_tiles"] = self.prob_map_tiles.clone()
        return d

    def set_state_dict(self, d):
        self._dyn.copy_(d["dyn"])
        self.tgt_xy.copy_(d["tgt"])
        if "outputs" in d:
            self._slab.copy_(d["outputs"])
        if self.prob_map_tiles is not None:
            if "prob_map_tiles" in d:
                self.prob_map_tiles.copy_(d["prob_map_tiles"])
                if self.map_overlap:          # later map kernels run on the handle's own stream: finish the copy first
                    torch.cuda.current_stream(self.device).synchronize()
            elif "prob_map" in d:
                self.prob_map = d["prob_map"]

    def host_buffers(self, compact=True):
        """Host-side results of step_host (allocated once).

        compact=True (default): the library's compact host path -- per step the device sends 16 + 16n bytes per env (a
        16-byte record, and the agent rows scattered in place by the copy engine) and the library keeps the
        reference-shaped rows up to date on the host (cs_flight_host_compact_begin);
        the returned tensors are views of library-owned host arrays, updated in place by every step_host.
        compact=False: a pinned mirror of the whole device output slab, fetched with one D2H copy per step."""
        key = "_host_c" if compact else "_host"
        if getattr(self, key, None) is None:
            E, n = self.num_envs, self.n_agents
            if compact:
                v = _lib.FlightHostViews()
                _lib.check(self.lib.cs_flight_host_compact_begin(self._h.ptr, C.byref(v)), "cs_flight_host_compact_begin")
                self._host_c = self._compact_views(v, torch.empty((E, n), dtype=torch.uint8).pin_memory(), self._h)
            else:
                lay = (C.c_uint64 * 8)()
                _lib.check(self.lib.cs_flight_slab_layout(self._h.ptr, lay), "cs_flight_slab_layout")
                total, o_rew, o_tf, o_term, o_win, o_obs, o_state, pitch = [int(x) for x in lay]
                slab = torch.empty(total, dtype=torch.uint8).pin_memory()
                stride = pitch // 4
                view = lambda off, nbytes, dtype: slab[off:off + nbytes].view(dtype)
                self._host = {
                    "actions": torch.empty((E, n), dtype=torch.uint8).pin_memory(),
                    "slab": slab,
                    "reward": view(o_rew, 4 * E, torch.float32),
                    "target_find": view(o_tf, 4 * E, torch.int32),
                    "terminated": view(o_term, E, torch.uint8),
                    "win": view(o_win, E, torch.uint8),
                    "state": view(o_state, pitch * E, torch.float32).view(E, stride)[:, :self.state_shape],
                    # get_obs rows are the agent part of the state rows (flight_env_easy.py:192-193): not sent twice
                    "obs": view(o_state, pitch * E, torch.float32).view(E, stride)[:, :4 * n].unflatten(1, (n, 4)),
                    "d2h_bytes": total,
                }
        return getattr(self, key)

    def _compact_views(self, v, actions, owner):
        """Tensor views of the library-owned host arrays described by a cs_flight_host_views."""
        E, n, stride = self.num_envs, self.n_agents, int(v.state_stride)

        def arr(ptr, ctype, count, dtype):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,))
            t = torch.from_numpy(a.view(dtype))
            t._cs_owner = owner
            return t
        state = arr(v.state, C.c_float, E * stride, np.float32).view(E, stride)
        return {
            "actions": actions,
            "reward": arr(v.reward, C.c_float, E, np.float32),
            "target_find": arr(v.target_find, C.c_int32, E, np.int32),
            "terminated": arr(v.terminated, C.c_uint8, E, np.uint8),
            "win": arr(v.win, C.c_uint8, E, np.uint8),
            "state": state[:, :self.state_shape],
            # get_obs rows are the agent part of the state rows (flight_env_easy.py:192-193)
            "obs": state[:, :4 * n].unflatten(1, (n, 4)),
            "d2h_bytes": int(v.d2h_bytes_per_step),
        }

    def step_host(self, actions, want_obs=True, want_state=True, sync=True, compact=True):
        """The call a CPU-side rollout makes: HOST actions in, HOST results out (numpy views of host buffers that the
        next step_host overwrites).  H2D + kernel(s) + D2H happen inside the library.  sync=False only enqueues
        (pipelining several env batches on different streams); synchronise the stream -- and, on the compact path, call
        host_expand() -- before reading the results."""
        hb = self.host_buffers(compact)
        a = np.asarray(actions, dtype=np.uint8)
        if a.shape != (self.num_envs, self.n_agents):
            raise CoopSearchError('Act num mismatch agent')
        hb["actions"].numpy()[...] = a
        with torch.cuda.device(self.device):
            if compact:
                _lib.check(self.lib.cs_flight_step_host_compact(self._h.ptr, C.c_void_p(hb["actions"].data_ptr()),
                                                                0 if sync else _lib.CS_HOST_NO_SYNC, self._stream()),
                           "cs_flight_step_host_compact")
            else:
                io = _lib.FlightHostIO(actions=hb["actions"].data_ptr(), slab=hb["slab"].data_ptr(),
                                       flags=0 if sync else _lib.CS_HOST_NO_SYNC)
                _lib.check(self.lib.cs_flight_step_host(self._h.ptr, C.byref(io), self._stream()), "cs_flight_step_host")
        return (hb["reward"].numpy(), hb["terminated"].numpy(), hb["win"].numpy(),
                hb["obs"].numpy() if want_obs else None, hb["state"].numpy() if want_state else None)

    def host_expand(self, sync=True):
        """Compact path, after step_host(sync=False): wait for the stream (sync=True) and rebuild the host rows."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_host_expand(self._h.ptr, self._stream(), 1 if sync else 0), "cs_flight_host_expand")


def generate_episodes(env, actions=None, targets=None, generator=None):
    """RolloutWorker.generate_episode (common/rollout.py:22-141) for every env of ``env`` at once: reset, then up to
    episode_limit steps, into the padded episode arrays of :118-132 with a leading num_envs axis (device tensors;
    float32 for o/s/r/o_next/s_next, uint8 for the rest).  ``actions``: [T,E,n] integers, or None for the
    uniform-random policy (alg=random, agent/agent.py:34-36) drawn with ``generator``.  ``targets``: optional
    [E,m,2] target layout to inject at the reset.  The env must not auto-reset: a finished env is a masked no-op and
    keeps its padding rows.  Returns (episode dict, episode_reward [E], win_tag [E], targets_find [E])."""
    if env.auto_reset:
        raise CoopSearchError("generate_episodes needs auto_reset=False (rollout.py:43: a finished env is not stepped again)")
    E, n, S, T, dev = env.num_envs, env.n_agents, env.state_shape, env.time_limit, env.device
    A = env.n_actions
    if actions is None:
        actions = torch.randint(0, A, (T, E, n), dtype=torch.uint8, device=dev, generator=generator)
    else:
        actions = torch.as_tensor(np.asarray(actions) if not torch.is_tensor(actions) else actions).to(device=dev, dtype=torch.uint8)
        if tuple(actions.shape) != (T, E, n):
            raise CoopSearchError("actions must have shape (episode_limit, num_envs, n_agents)")
    actions = actions.contiguous()
    env.reset(targets=targets)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    u8 = lambda *shape: torch.empty(shape, dtype=torch.uint8, device=dev)
    ep = {"o": f32(E, T, n, 4), "s": f32(E, T, S), "u": u8(E, T, n, 1), "r": f32(E, T, 1), "avail_u": u8(E, T, n, A),
          "o_next": f32(E, T, n, 4), "s_next": f32(E, T, S), "avail_u_next": u8(E, T, n, A), "u_onehot": u8(E, T, n, A),
          "padded": u8(E, T, 1), "terminated": u8(E, T, 1)}
    summary = {"episode_reward": f32(E), "win_tag": u8(E), "targets_find": torch.empty(E, dtype=torch.int32, device=dev),
               "length": torch.empty(E, dtype=torch.int32, device=dev)}
    bufs = _lib.EpisodeBuffers(**{k: v.data_ptr() for k, v in {**ep, **summary}.items()})
    with torch.cuda.device(dev):
        _lib.check(env.lib.cs_flight_record_begin(env._h.ptr, C.byref(bufs), T, env._stream()), "cs_flight_record_begin")
        for t in range(T):
            a = actions[t]
            _lib.check(env.lib.cs_flight_step(env._h.ptr, C.c_void_p(a.data_ptr()), env._stream()), "cs_flight_step")
            _lib.check(env.lib.cs_flight_record(env._h.ptr, C.byref(bufs), t, T, C.c_void_p(a.data_ptr()), env._stream()),
                       "cs_flight_record")
    ep["length"] = summary["length"]
    return ep, summary["episode_reward"], summary["win_tag"], summary["targets_find"]


class DeviceStepper:
    """One env-step of many independent flight_easy env batches (rollout workers) in ONE kernel launch
    (cs_flight_group_step).  The envs must share n_agents, lanes_per_env and device; results land in each env's own
    buffers exactly as calling ``env.step`` on each would leave them."""

    def __init__(self, envs):
        self.envs = list(envs)
        self.lib = self.envs[0].lib
        self.device = self.envs[0].device
        n = len(self.envs)
        handles = (C.c_void_p * n)(*[e._h.ptr for e in self.envs])
        g = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_group_create(handles, n, C.byref(g)), "cs_flight_group_create")
        self._g = g
        self._ptrs = (C.c_void_p * n)()

    def step(self, actions):
        """actions: one [E,n] integer tensor per env."""
        keep = []
        for i, (e, a) in enumerate(zip(self.envs, actions)):
            if not (torch.is_tensor(a) and a.dtype is torch.uint8 and a.is_cuda and a.is_contiguous()
                    and tuple(a.shape) == (e.num_envs, e.n_agents)):
                a = e._as_actions(a)              # any integer tensor / array -> device u8 [E,n]
            keep.append(a)
            self._ptrs[i] = a.data_ptr()
        if len(keep) != len(self.envs):
            raise CoopSearchError("DeviceStepper.step needs one action tensor per env")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_group_step(self._g, self._ptrs, self.envs[0]._stream()), "cs_flight_group_step")

    def __del__(self):
        try:
            if self._g:
                self.lib.cs_flight_group_destroy(self._g)
                self._g = None
        except Exception:
            pass


class _Pool:
    """Owner of a cs_flight_host_pool: freed with the last reference (the stepper, or a view of the pooled arrays).  The
    library detaches envs that are destroyed first."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.cs_flight_host_pool_destroy(C.c_void_p(self.ptr))
        except Exception:  # pragma: no cover
            pass
        self.ptr = None


class HostStepper:
    """Host-buffer steps of many independent env batches (rollout workers) with ONE library call per step; results land
    in each env's host_buffers(compact).

    compact=True (default), pooled (the default whenever the envs have no compact host buffers yet): the batches share
    pooled buffers (cs_flight_host_pool_*): one H2D copy of all actions, one grouped step launch, one pack launch, one
    flat D2H copy of the 16-byte records and one strided D2H copy that writes every agent row in place into the
    reference-shaped host rows; host threads then update result arrays and find flags.  ``self.actions`` is the pinned
    uint8 [sum E, n] action buffer of all batches (each env's host_buffers()["actions"] is its segment); step(actions=t)
    takes the actions from another pinned tensor of that shape instead.
    compact=True, not pooled: cs_flight_step_host_compact_many, batch i on stream i % len(streams).
    compact=False: cs_flight_step_host_many -- the whole output slab of every batch comes back.
    ``actions`` (optional, at construction; not pooled): pinned uint8 [E,n] tensors, one per env, that hold the actions
    of every step.  graph=True captures the device side of the step into one CUDA graph after a first eager step, so
    that a step costs one graph launch on the host."""

    def __init__(self, envs, streams, actions=None, graph=False, compact=True, pooled=None):
        self.envs = list(envs)
        self.lib = self.envs[0].lib
        self.device = self.envs[0].device
        self.streams = list(streams)
        self.compact = bool(compact)
        n = len(self.envs)
        self._handles = (C.c_void_p * n)(*[e._h.ptr for e in self.envs])
        self._streams = (C.c_void_p * len(self.streams))(*[st.cuda_stream for st in self.streams])
        self._pool = None
        self.actions = None
        can_pool = self.compact and actions is None and all(getattr(e, "_host_c", None) is None for e in self.envs)
        if pooled and not can_pool:
            raise CoopSearchError("HostStepper(pooled=True) needs compact=True, no per-env action tensors and envs without compact host buffers")
        if can_pool and pooled is not False:
            self._make_pool()
        self._ios = (_lib.FlightHostIO * n)()
        self._acts = (C.c_void_p * n)()
        self._keep = actions
        for i, e in enumerate(self.envs):
            hb = e.host_buffers(self.compact)
            self._acts[i] = (actions[i] if actions is not None else hb["actions"]).data_ptr()
            self._ios[i].actions = self._acts[i]
            if not self.compact:
                self._ios[i].slab = hb["slab"].data_ptr()
        self._want_graph = bool(graph)
        self._graph = None
        self._graph_actions = None
        self._cap = torch.cuda.Stream(device=self.device) if graph else None

    def _make_pool(self):
        hp = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.cs_flight_host_pool_create(self._handles, len(self.envs), C.byref(hp))
        if rc != 0:
            return                       # batches that do not pool keep their own buffers
        self._pool = _Pool(self.lib, hp.value)
        total = sum(e.num_envs for e in self.envs)
        n_agents = self.envs[0].n_agents
        first = None
        for i, e in enumerate(self.envs):
            v, ap = _lib.FlightHostViews(), C.c_void_p()
            _lib.check(self.lib.cs_flight_host_pool_views(hp, i, C.byref(v), C.byref(ap)), "cs_flight_host_pool_views")
            if i == 0:
                first = ap.value
            a = np.ctypeslib.as_array(C.cast(ap, C.POINTER(C.c_uint8)), shape=(e.num_envs * n_agents,))
            act = torch.from_numpy(a).view(e.num_envs, n_agents)
            act._cs_owner = self._pool
            e._host_c = e._compact_views(v, act, self._pool)
        a = np.ctypeslib.as_array(C.cast(C.c_void_p(first), C.POINTER(C.c_uint8)), shape=(total * n_agents,))
        self.actions = torch.from_numpy(a).view(total, n_agents)
        self.actions._cs_owner = self._pool

    def _call(self, flags, actions=None):
        n = len(self.envs)
        if self._pool is not None:
            ptr = C.c_void_p(actions.data_ptr()) if actions is not None else None
            st = torch.cuda.current_stream(self.device).cuda_stream if self._want_graph else self.streams[0].cuda_stream
            _lib.check(self.lib.cs_flight_host_pool_step(C.c_void_p(self._pool.ptr), ptr, flags, C.c_void_p(st)), "cs_flight_host_pool_step")
            return
        if self.compact:
            _lib.check(self.lib.cs_flight_step_host_compact_many(self._handles, self._acts, n, self._streams, len(self.streams), flags),
                       "cs_flight_step_host_compact_many")
            return
        for i in range(n):
            self._ios[i].flags = flags
        _lib.check(self.lib.cs_flight_step_host_many(self._handles, self._ios, n, self._streams, len(self.streams)),
                   "cs_flight_step_host_many")

    def _expand(self, sync):
        if self._pool is not None:
            st = self._cap.cuda_stream if self._want_graph else self.streams[0].cuda_stream
            _lib.check(self.lib.cs_flight_host_pool_expand(C.c_void_p(self._pool.ptr), C.c_void_p(st), 1 if sync else 0), "cs_flight_host_pool_expand")
        elif self.compact:
            _lib.check(self.lib.cs_flight_host_expand_many(self._handles, len(self.envs), self._streams, len(self.streams), 1 if sync else 0),
                       "cs_flight_host_expand_many")

    def step(self, actions=None):
        """One env-step of every batch; returns when all results are in host memory.  actions (pooled only): a pinned
        uint8 [sum E, n] tensor to take this step's actions from instead of ``self.actions``."""
        if actions is not None and self._pool is None:
            raise CoopSearchError("HostStepper.step(actions=...) needs the pooled form")
        if not self._want_graph:
            with torch.cuda.device(self.device):
                self._call(0, actions)
            return
        key = None if actions is None else actions.data_ptr()
        if self._graph is None or key not in self._graph:
            if self._graph is None:
                self._graph = {}
                self._cap.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self._cap):
                    self._call(0, actions)               # eager first step (warm-up outside the capture)
                return
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._cap):
                for st in self.streams:
                    st.wait_stream(self._cap)
                self._call(_lib.CS_HOST_NO_SYNC, actions)
                for st in self.streams:
                    self._cap.wait_stream(st)
            self._graph[key] = g                 # (capturing records the launches, it does not run them; one graph per action buffer)
            self._keep_actions = getattr(self, "_keep_actions", []) + [actions]
        self._cap.wait_stream(torch.cuda.current_stream(self.device))       # e.g. a reset issued on the caller's stream
        with torch.cuda.stream(self._cap):
            self._graph[key].replay()
        self._cap.synchronize()
        self._expand(sync=False)


class VecFlightEasyEnv(_VecFlightBase):
    """num_envs x FlightSearchEnvEasy (env/flight_env_easy.py)."""
    VARIANT = 0
    ENV_NAME = "flight_easy"


class VecFlightEnv(_VecFlightBase):
    """num_envs x FlightSearchEnv (env/flight_env.py): flight_easy dynamics with the '>=' wall test
    (:328) plus the per-env probability map updated by every sensing call (:275-303)."""
    VARIANT = 1
    ENV_NAME = "flight"

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._obs_full = None

    def set_obs_kernel(self, kind="tma"):
        """How the de-tiling observation kernel writes the reference-shaped rows: "tma" (bulk async stores from shared
        memory, default) or "plain" (ordinary stores); results are identical, the switch exists for A/B measurement."""
        _lib.check(self.lib.cs_debug_flight_obs_path(self._h.ptr, {"tma": 0, "plain": 1}[kind]), "cs_debug_flight_obs_path")

    def get_obs(self, full=True):
        """Reference-shaped [E,n,M*M+4] = prob_map.ravel() || 4 features (flight_env.py:223-230),
        materialised by a streaming kernel.  full=False returns the [E,n,4] features only; the map
        itself is available zero-copy in its tiled device layout as ``self.prob_map_tiles`` and as a row-major
        copy as ``self.prob_map`` ([E,M,M])."""
        if not full:
            return self._obs
        if self._obs_full is None:
            self._obs_full = torch.empty((self.num_envs, self.n_agents, self.map_size ** 2 + 4),
                                         dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_flight_obs_full(self._h.ptr, C.c_void_p(self._obs_full.data_ptr()), self._stream()),
                       "cs_flight_obs_full")
        return self._obs_full
