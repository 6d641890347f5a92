"""Scalar Python restatement of the reference env-step hot path -- TEST INFRASTRUCTURE.

This module is the *oracle*: a CPU restatement of the algorithms of
  env/flight_env_easy.py   (FlightOracle, variant="easy")
  env/flight_env.py        (FlightOracle, variant="probmap")
  env/search_env.py        (SearchOracle)
of WZN1ng/Cooperative-Search, float64 scalar arithmetic in the same operation
order as the reference, with the reference's global ``np.random`` detection draw
replaced by the keyed Philox draw of oracle/philox.py.

PINNING: the reference ships no tests or golden vectors ("parity unpinned" by
its own tests).  This oracle is pinned instead against the
reference *itself*, executed in the build container: tests/golden/*.npz are
trajectories produced by the unmodified reference classes under
oracle/refharness.py (generator: tests/golden/make_golden.py), and
tests/test_oracle_golden.py requires this restatement to reproduce them
(integers bit-exact, float64 state to 1e-12).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import philox

FIND_ONE_TGT = 10          # env/flight_env_easy.py:54
FIND_ALL_TGT = 100         # env/flight_env_easy.py:55
OUT_PUNISH = -1            # env/flight_env_easy.py:61
MOVE_COST = -1             # env/flight_env_easy.py:62
FORCE_FACTOR = 0.8         # env/flight_env_easy.py:65


@dataclass
class FlightSpec:
    n_agents: int = 3
    target_num: int = 15
    map_size: int = 50
    view_range: int = 7
    velocity: float = 1
    time_limit: int = 200
    detect_prob: float = 0.9
    safe_dist: float = 1
    force_dist: float = 3
    agent_mode: int = 0
    target_mode: int = 0
    variant: str = "easy"          # "easy" -> flight_env_easy.py, "probmap" -> flight_env.py

    @property
    def state_shape(self):
        return 4 * self.n_agents + 3 * self.target_num


def keyed_target_layout(spec, template, seed, env_id, episode):
    """Target placement used by the *device-side* reset (no reference counterpart
    for the random stream: the reference draws np.random.randn()/rand() from the
    global MT19937 stream, env/flight_env_easy.py:107-108,124).  Same formulae,
    normals from Box-Muller over Philox words.  Returns list of [x, y]."""
    M = spec.map_size
    out = []
    for j in range(spec.target_num):
        w = philox.philox4x32(env_id, (episode & 0xFFFF) << 16, j, 0, seed, philox.stream_key(philox.STREAM_TARGET, episode))
        u1, u2 = philox.u53(w[0], w[1]), philox.u53(w[2], w[3])
        if spec.target_mode == 0:
            a = M / 10
            x = a * template["x"][j]
            y = a * template["y"][j]
            if template["deter"][j] == "f":
                rad = math.sqrt(-2.0 * math.log(u1))
                ang = 2.0 * math.pi * u2
                z0, z1 = rad * math.cos(ang), rad * math.sin(ang)
                x += (a * template["dx"][j]) * 2 * (z0 - 0.5)
                y += (a * template["dy"][j]) * 2 * (z1 - 0.5)
        elif spec.target_mode == 1:
            x, y = M * u1, M * u2
        else:
            raise ValueError("No such target mode")
        out.append([x, y])
    return out


class FlightOracle:
    """One env instance.  Restates FlightSearchEnvEasy / FlightSearchEnv."""

    def __init__(self, spec, template=None, seed=0, env_id=0):
        self.spec = spec
        self.template = template
        self.seed = seed
        self.env_id = env_id
        self.episode = -1
        self.pos = []            # [[x, y]] per agent
        self.yaw = []
        self.out = []
        self.tgt = []            # [[x, y]] per target
        self.found = []
        self.new_found = []      # indices found by the latest sensing call
        self.target_find = 0
        self.win = False
        self.time_step = 0
        self.reward = 0
        self.total_reward = 0
        self.feat = []           # un-normalised [x, y, cos, sin] per agent (flight_env_easy.py:252)
        self.n_touched = 0       # cells with percent > 0 in the latest map update
        if spec.variant == "probmap":
            self.prob_map = 0.5 * np.ones((spec.map_size, spec.map_size))   # flight_env.py:53

    # ------------------------------------------------------------------ reset
    def _start_agents(self):
        """Agent start layouts, env/flight_env_easy.py:139-180."""
        s = self.spec
        n, M = s.n_agents, s.map_size
        lin = [i * M / (n - 1) for i in range(n)] if n != 1 else [M / 2]
        self.pos, self.yaw, self.out = [], [], []
        for i in range(n):
            if s.agent_mode == 0:
                p, h = [lin[i], 0], np.pi / 2
            elif s.agent_mode == 1:
                p, h = [lin[i], M / 2], np.pi / 2
            elif s.agent_mode == 2:
                p, h = [0, lin[i]], 0
            elif s.agent_mode == 3:
                p, h = [M, lin[i]], np.pi
            else:
                raise ValueError("No such agent mode")
            self.pos.append(p); self.yaw.append(h); self.out.append(0)

    def reset(self, targets=None, init=False, episode=None):
        """env/flight_env_easy.py:79-182 / env/flight_env.py:83-191.
        `targets`: injected target coordinates (parity tests); None -> keyed layout."""
        s = self.spec
        self.episode = self.episode + 1 if episode is None else episode
        if s.variant == "probmap" and init:
            self.prob_map = 0.5 * np.ones((s.map_size, s.map_size))      # flight_env.py:84-86
        self.time_step = 0
        self.target_find = 0
        self.total_reward = 0
        self.win = False
        if targets is None:
            targets = keyed_target_layout(s, self.template, self.seed, self.env_id, self.episode)
        self.tgt = [[t[0], t[1]] for t in targets]
        self.found = [False] * s.target_num
        self._start_agents()
        self._sense(0)                                                    # flight_env_easy.py:182

    # ---------------------------------------------------------------- sensing
    def _sense(self, t):
        """_update_obs: env/flight_env_easy.py:223-253, env/flight_env.py:232-266."""
        s = self.spec
        self.feat = []
        self.new_found = []
        rew = MOVE_COST
        thr = philox.detect_threshold(s.detect_prob)
        for i in range(s.n_agents):
            x, y = self.pos[i]
            for j in range(s.target_num):
                tx, ty = self.tgt[j]
                if (tx - x) ** 2 + (ty - y) ** 2 <= s.view_range ** 2:
                    r = philox.detect_draw(self.seed, self.env_id, self.episode, t, i, j)
                    if (not self.found[j]) and r <= thr:
                        self.found[j] = True
                        rew += FIND_ONE_TGT
                        self.target_find += 1
                        self.new_found.append(j)
                        if self.target_find == s.target_num and not self.win:
                            rew += FIND_ALL_TGT
                            self.win = True
            if self.out[i]:
                rew += OUT_PUNISH
            h = self.yaw[i]
            self.feat.append(np.array([x, y, np.cos(h), np.sin(h)]))
        self.reward = rew
        if s.variant == "probmap":
            self._belief_update(self.new_found)

    # ------------------------------------------------------------- kinematics
    def _repulsion(self, k):
        """_potential_energy_force: env/flight_env_easy.py:293-301 (agent k's OWN stored
        position, the CURRENT list for everybody else -> Gauss-Seidel)."""
        s = self.spec
        x, y = self.pos[k]
        fx, fy = 0, 0
        for q in range(s.n_agents):
            xa, ya = self.pos[q]
            if q != k and (xa - x) ** 2 + (ya - y) ** 2 < s.force_dist ** 2:
                if xa != x or ya != y:
                    fx += s.safe_dist * FORCE_FACTOR * s.velocity * (x - xa) / ((x - xa) ** 2 + (y - ya) ** 2)
                    fy += s.safe_dist * FORCE_FACTOR * s.velocity * (y - ya) / ((x - xa) ** 2 + (y - ya) ** 2)
        return fx, fy

    def _move(self, actions):
        """_agent_step: env/flight_env_easy.py:255-291, env/flight_env.py:305-345."""
        s = self.spec
        if len(actions) != s.n_agents:
            raise ValueError("Act num mismatch agent")
        M = s.map_size
        turn = [0, np.pi / 18, -np.pi / 18]
        for k in range(s.n_agents):
            x, y = self.pos[k]
            h = self.yaw[k] + turn[int(actions[k])]
            if h > 2 * np.pi:
                h -= 2 * np.pi
            elif h < 0:
                h += 2 * np.pi
            x += s.velocity * np.cos(h)
            y += s.velocity * np.sin(h)
            fx, fy = self._repulsion(k)
            x += fx
            y += fy
            if s.variant == "easy":
                outside = x < 0 or x > M or y < 0 or y > M          # flight_env_easy.py:278
            else:
                outside = x < 0 or x >= M or y < 0 or y >= M        # flight_env.py:328
            if outside:
                x = min(max(x, 0), M)
                y = min(max(y, 0), M)
                h = np.pi - h if h <= np.pi else 3 * np.pi - h
                self.out[k] = 1
            else:
                self.out[k] = 0
            self.pos[k] = [x, y]
            self.yaw[k] = h

    def step(self, actions):
        """env/flight_env_easy.py:303-314."""
        s = self.spec
        self._move(actions)
        self._sense(self.time_step + 1)
        self.total_reward += self.reward
        self.time_step += 1
        terminated = self.target_find >= s.target_num or self.time_step >= s.time_limit
        return self.reward, terminated, self.win

    # ---------------------------------------------------------- observations
    def get_obs(self):
        """env/flight_env_easy.py:218-221; env/flight_env.py:223-230 for the map variant."""
        s = self.spec
        o = np.array(self.feat)
        o[:, :2] = (o[:, :2] - 0.5 * s.map_size) / (s.map_size / 2)
        if s.variant == "probmap":
            pm = np.repeat(self.prob_map.reshape(1, s.map_size * s.map_size), s.n_agents, axis=0)
            o = np.concatenate((pm, o), axis=1)
        return o

    def get_state(self):
        """env/flight_env_easy.py:190-216."""
        s = self.spec
        a = np.array(self.feat)
        a[:, :2] = (a[:, :2] - 0.5 * s.map_size) / (s.map_size / 2)
        t = np.array([[p[0], p[1], 1.0 if f else 0.0] for p, f in zip(self.tgt, self.found)])
        t[:, :2] = (t[:, :2] - 0.5 * s.map_size) / (s.map_size / 2)
        return np.hstack([a.reshape(4 * s.n_agents), t.reshape(3 * s.target_num)])

    def get_avail_agent_actions(self, agent_id):
        if agent_id >= self.spec.n_agents:
            raise ValueError("Agent id out of range")
        return np.ones(3)

    def found_mask(self):
        return sum(1 << j for j, f in enumerate(self.found) if f)

    # ------------------------------------------------------------ belief map
    def _corner_fraction(self, i, j):
        """_percent_in_agent_viewrange: env/flight_env.py:294-303."""
        R2 = self.spec.view_range ** 2
        inside = 0
        for cx, cy in ((i, j), (i + 1, j), (i, j + 1), (i + 1, j + 1)):
            for ax, ay in self.pos:
                if (cx - ax) ** 2 + (cy - ay) ** 2 < R2:
                    inside += 1
                    break
        return inside / 4

    def _belief_update(self, found_now):
        """_update_prob_map: env/flight_env.py:275-292."""
        s = self.spec
        M, d = s.map_size, s.detect_prob
        hit = [[min(int(self.tgt[j][0]), M - 1), min(int(self.tgt[j][1]), M - 1)] for j in found_now]
        touched = 0
        # bounding rows/cols that can hold a corner inside some disc; outside of it the
        # fraction is 0 and the reference leaves the cell untouched (flight_env.py:285-286)
        R = s.view_range
        lo_i = max(0, int(math.floor(min(p[0] for p in self.pos) - R)) - 1)
        hi_i = min(M - 1, int(math.ceil(max(p[0] for p in self.pos) + R)))
        lo_j = max(0, int(math.floor(min(p[1] for p in self.pos) - R)) - 1)
        hi_j = min(M - 1, int(math.ceil(max(p[1] for p in self.pos) + R)))
        for i in range(lo_i, hi_i + 1):
            for j in range(lo_j, hi_j + 1):
                frac = self._corner_fraction(i, j)
                if frac == 0:
                    continue
                touched += 1
                if [i, j] in hit:
                    self.prob_map[i, j] = 1
                else:
                    p = self.prob_map[i, j]
                    self.prob_map[i, j] = frac * (1 - d) * p / ((1 - d) * p + (1 - p))
        self.n_touched = touched


# =============================================================================
#  search_env
# =============================================================================
@dataclass
class SearchSpec:
    n_agents: int = 3
    target_num: int = 15
    map_size: int = 50
    view_range: int = 7
    agent_mode: int = 0
    target_mode: int = 0
    circle_dict: dict = field(default_factory=dict)

    @property
    def obs_size(self):
        return 2 * self.view_range - 1

    @property
    def obs_shape(self):
        return self.obs_size ** 2 + 2

    @property
    def state_shape(self):
        return 2 * self.map_size ** 2


def keyed_search_cells(spec, seed, env_id, episode):
    """Device-side target placement for search_env target_mode 0/1 (reference:
    np.random.randint rejection sampling, env/search_env.py:86-104).  Candidate k for
    the env is Philox(env, episode, k) -> (x, y) = (w0 % M, w1 % M); candidates are
    consumed in order k = 0, 1, ... and rejected when occupied (mode 1: or not in the
    edge band), exactly the reference's accept rule."""
    M = spec.map_size
    lo, hi = M // 4, 3 * M // 4
    cells, taken, k = [], set(), 0
    while len(cells) < spec.target_num:
        w = philox.philox4x32(env_id, (episode & 0xFFFF) << 16, k, 0, seed, philox.stream_key(philox.STREAM_SEARCH, episode))
        k += 1
        x, y = w[0] % M, w[1] % M
        if (x, y) in taken:
            continue
        if spec.target_mode == 1 and not (x <= lo or x >= hi or y <= lo or y >= hi):
            continue
        taken.add((x, y))
        cells.append([x, y])
    return cells


class SearchOracle:
    """One discrete-grid env instance.  Restates SearchEnv (env/search_env.py)."""

    MOVE_COST = -1       # search_env.py:46
    REWARD_FIND = 10     # search_env.py:45

    def __init__(self, spec, seed=0, env_id=0):
        self.spec = spec
        self.seed, self.env_id = seed, env_id
        self.episode = -1
        M = spec.map_size
        self.freq = np.zeros((M, M))                 # never cleared: search_env.py:39 vs :70-80
        self.target_map = np.zeros((M, M))
        self.state = np.zeros((M, M, 2))
        self.pos = []
        self.cells = []
        self.found = []
        self.target_find = 0
        self.time_step = 0
        self.win = []
        self.illegal = False

    def _start_agents(self):
        """search_env.py:146-180."""
        s = self.spec
        n, M = s.n_agents, s.map_size
        self.pos = []
        if s.agent_mode == 0:
            L = int(np.ceil(np.sqrt(n)))
            b = (M - L) // 2
            for i in range(b, b + L):
                for j in range(b, b + L):
                    if len(self.pos) < n:
                        self.pos.append([i, j])
        elif s.agent_mode == 1:
            L = int(np.ceil(np.sqrt(n)))
            for i in range(M - 1, M - 1 - L, -1):
                for j in range(0, L):
                    if len(self.pos) < n:
                        self.pos.append([i, j])
        elif s.agent_mode == 2:
            gap = (M - 1) // (n - 1)
            for i in range(n):
                self.pos.append([M - 1, i * gap])
        else:
            raise ValueError("Unknown agent mode")
        for x, y in self.pos:
            self.freq[x, y] += 1

    def reset(self, cells=None, episode=None):
        """search_env.py:69-183 with init=False semantics (everything but freq cleared)."""
        s = self.spec
        M = s.map_size
        self.episode = self.episode + 1 if episode is None else episode
        self.target_map = np.zeros((M, M))
        self.state = np.zeros((M, M, 2))
        self.time_step = 0
        self.target_find = 0
        self.illegal = False
        if cells is None:
            cells = keyed_search_cells(s, self.seed, self.env_id, self.episode)
        self.cells = [[int(c[0]), int(c[1])] for c in cells]
        self.found = [False] * len(self.cells)
        for x, y in self.cells:
            self.target_map[x, y] = 1
            self.state[x, y, 0] = 1
        self._start_agents()
        self._windows()
        self._planes()

    def _planes(self):
        """_update_state/_clear_agent_state: search_env.py:190-200."""
        M = self.spec.map_size
        for x, y in self.pos:
            for dx, dy in ((0, 0), (-1, 0), (1, 0), (0, 1), (0, -1)):
                if 0 <= x + dx < M and 0 <= y + dy < M:
                    self.state[x + dx, y + dy, 1] = 0
        for x, y in self.pos:
            self.state[x, y, 1] = 1

    def _windows(self):
        """_update_obs: search_env.py:212-227."""
        s = self.spec
        S, R, M = s.obs_size, s.view_range, s.map_size
        self.win = []
        for x, y in self.pos:
            w = np.zeros((S, S))
            for i in range(S):
                for j in range(S):
                    gx, gy = i + x - R + 1, j + y - R + 1
                    if 0 <= gx < M and 0 <= gy < M:
                        if (R - 1 - i) ** 2 + (R - 1 - j) ** 2 > R ** 2:
                            w[i, j] = 0.5
                        elif self.target_map[gx, gy] == 1:
                            w[i, j] = 1
                    else:
                        w[i, j] = 0.5
            self.win.append(w)

    def get_obs(self):
        """search_env.py:203-210 (without the prints)."""
        S = self.spec.obs_size
        o = np.array(self.win).reshape(-1, S * S)
        return np.concatenate((o, np.array(self.pos)), axis=1)

    def get_state(self):
        return self.state.reshape(self.spec.state_shape)

    def get_avail_agent_actions(self, agent_id):
        """search_env.py:230-243."""
        if agent_id >= self.spec.n_agents:
            raise ValueError("Agent id out of range")
        M = self.spec.map_size
        x, y = self.pos[agent_id]
        return np.array([float(x > 0), float(y > 0), float(x < M - 1), float(y < M - 1)])

    def step(self, actions):
        """search_env.py:246-296.  An illegal move raises in the reference (:293); here it
        is recorded in `self.illegal` and the agent stays (the batched API's contract)."""
        s = self.spec
        M = s.map_size
        if len(actions) != s.n_agents:
            raise ValueError("Act num mismatch agent")
        rew = self.MOVE_COST
        self.time_step += 1
        for i in range(s.n_agents):
            x, y = self.pos[i]
            a = int(actions[i])
            if a == 0 and x > 0:
                self.pos[i][0] -= 1
            elif a == 1 and y > 0:
                self.pos[i][1] -= 1
            elif a == 2 and x < M - 1:
                self.pos[i][0] += 1
            elif a == 3 and y < M - 1:
                self.pos[i][1] += 1
            else:
                self.illegal = True
                continue
            self.freq[self.pos[i][0], self.pos[i][1]] += 1
        for xa, ya in self.pos:
            rew += 1 / self.freq[xa, ya]
            for k, (xt, yt) in enumerate(self.cells):
                if (xt - xa) ** 2 + (yt - ya) ** 2 <= s.view_range ** 2:
                    if not self.found[k]:
                        rew += self.REWARD_FIND
                        self.found[k] = True
                        self.target_find += 1
        terminated = self.target_find >= s.target_num
        self._windows()
        self._planes()
        return rew, terminated, ""


# =============================================================================
#  simple_spread  (env/simple_spread.py)
# =============================================================================
@dataclass
class SpreadSpec:
    n_agents: int = 3
    target_num: int = 3
    map_size: int = 50
    time_limit: int = 100          # hard-coded in the reference (simple_spread.py:25)
    agent_radius: int = 6          # (:24)

    @property
    def n_actions(self):
        return 5                   # (:26)

    @property
    def state_shape(self):
        return self.n_agents * 2 + self.target_num * 2                                   # (:27)

    @property
    def obs_shape(self):
        return 2 + (self.n_agents - 1) * 2 + self.target_num * 4                         # (:28)


def keyed_spread_layout(spec, seed, env_id, episode):
    """Device-side reset placement (reference: map_size * np.random.rand() twice per entity, targets first, then agents,
    simple_spread.py:57-68).  Entity k (targets 0..m-1, then agents) takes one Philox block: x = M * u53(w0, w1),
    y = M * u53(w2, w3)."""
    M = spec.map_size
    out = []
    for k in range(spec.target_num + spec.n_agents):
        w = philox.philox4x32(env_id, (episode & 0xFFFF) << 16, k, 0, seed, philox.stream_key(philox.STREAM_SPREAD, episode))
        out.append([M * philox.u53(w[0], w[1]), M * philox.u53(w[2], w[3])])
    return out[:spec.target_num], out[spec.target_num:]


class SimpleSpreadOracle:
    """Scalar restatement of SimpleSpreadEnv in the reference's operation order (Python floats, `**2`, np.sqrt)."""
    DPOS = [[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]]                                     # (:135)

    def __init__(self, spec, seed=0, env_id=0):
        self.spec, self.seed, self.env_id = spec, seed, env_id
        self.episode = -1
        self.time_step = 0
        self.total_reward = 0.0
        self.tgt, self.agents, self.occupied = [], [], []

    def reset(self, targets=None, agents=None):
        """(:48-70); targets / agents inject a layout (e.g. the reference's own MT19937 draws)."""
        self.episode += 1
        self.time_step = 0
        self.total_reward = 0.0
        if targets is None or agents is None:
            kt, ka = keyed_spread_layout(self.spec, self.seed, self.env_id, self.episode)
            targets = kt if targets is None else targets
            agents = ka if agents is None else agents
        self.tgt = [[float(p[0]), float(p[1])] for p in targets]
        self.agents = [[float(p[0]), float(p[1])] for p in agents]
        self._update_occupied()

    def _update_occupied(self):
        """(:96-110)"""
        self.occupied = []
        for t in self.tgt:
            occ = 0
            for a in self.agents:
                dis = 0
                for k in range(2):
                    dis += (t[k] - a[k]) ** 2
                if np.sqrt(dis) < self.spec.agent_radius:
                    occ = 1
                    break
            self.occupied.append(occ)

    def get_obs(self):
        """(:78-101,112-114): own position, targets relative, other agents relative, then targets absolute."""
        h = 0.5 * self.spec.map_size
        obs = []
        for i, a in enumerate(self.agents):
            row = [a[0] - h, a[1] - h]
            for t in self.tgt:
                row += [t[0] - a[0], t[1] - a[1]]
            for j, o in enumerate(self.agents):
                if i != j:
                    row += [o[0] - a[0], o[1] - a[1]]
            obs.append(row)
        for t in self.tgt:
            for row in obs:
                row += [t[0] - h, t[1] - h]
        return np.array(obs)

    def get_state(self):
        """(:116-128)"""
        h = 0.5 * self.spec.map_size
        s = []
        for a in self.agents:
            s += [a[0] - h, a[1] - h]
        for t in self.tgt:
            s += [t[0] - h, t[1] - h]
        return np.array(s)

    def reward(self):
        """(:141-153): minus the distance from every target to its nearest agent, summed in target order."""
        r = 0
        for t in self.tgt:
            dis = []
            for a in self.agents:
                d = 0
                for k in range(2):
                    d += (a[k] - t[k]) ** 2
                dis.append(np.sqrt(d))
            r -= min(dis)
        return r

    def step(self, act_list):
        """(:130-139,169-180)"""
        if len(act_list) != self.spec.n_agents:
            raise Exception('Act num mismatch agent')
        M = self.spec.map_size
        for i, a in enumerate(self.agents):
            for k in range(2):
                a[k] += self.DPOS[act_list[i]][k]
                a[k] = min(max(0, a[k]), M)
        self._update_occupied()
        r = self.reward()
        self.total_reward += r
        self.time_step += 1
        return r, self.time_step >= self.spec.time_limit, False


class SpreadBatch:
    """numpy-vectorised form of SimpleSpreadOracle for the configuration-size GPU tests (squares by multiplication; the
    reference's `**2` is libm pow, which differs in the last bit for ~0.08 % of arguments -- only the float64 reward can see
    it, by one ulp)."""

    def __init__(self, spec, seed, env_id_base, E, auto_reset=False):
        self.spec, self.seed, self.base, self.E, self.auto_reset = spec, seed, env_id_base, E, auto_reset
        self.tgt = np.zeros((E, spec.target_num, 2))
        self.agents = np.zeros((E, spec.n_agents, 2))
        self.time_step = np.zeros(E, np.int32)
        self.episode = np.full(E, -1, np.int64)
        self.done = np.zeros(E, bool)

    def reset(self, mask=None):
        sel = np.ones(self.E, bool) if mask is None else np.asarray(mask, bool)
        for e in np.nonzero(sel)[0]:
            self.episode[e] += 1
            t, a = keyed_spread_layout(self.spec, self.seed, self.base + int(e), int(self.episode[e]))
            self.tgt[e], self.agents[e] = np.array(t), np.array(a)
        self.time_step[sel] = 0
        self.done[sel] = False

    def obs_state(self):
        sp, h = self.spec, 0.5 * self.spec.map_size
        n, m, E = sp.n_agents, sp.target_num, self.E
        obs = np.zeros((E, n, sp.obs_shape))
        obs[:, :, 0:2] = self.agents - h
        rel_t = self.tgt[:, None, :, :] - self.agents[:, :, None, :]                      # [E,n,m,2]
        obs[:, :, 2:2 + 2 * m] = rel_t.reshape(E, n, 2 * m)
        for i in range(n):
            others = [j for j in range(n) if j != i]
            rel_a = self.agents[:, others, :] - self.agents[:, i:i + 1, :]
            obs[:, i, 2 + 2 * m:2 + 2 * m + 2 * (n - 1)] = rel_a.reshape(E, 2 * (n - 1))
        obs[:, :, 2 + 2 * m + 2 * (n - 1):] = np.broadcast_to((self.tgt - h).reshape(E, 1, 2 * m), (E, n, 2 * m))
        state = np.concatenate([(self.agents - h).reshape(E, 2 * n), (self.tgt - h).reshape(E, 2 * m)], axis=1)
        return obs, state

    def step(self, actions):
        sp = self.spec
        dpos = np.array(SimpleSpreadOracle.DPOS, np.float64)
        live = ~self.done
        act = np.asarray(actions)
        moved = np.minimum(np.maximum(0.0, self.agents + dpos[act]), float(sp.map_size))
        self.agents[live] = moved[live]
        d = self.agents[:, None, :, :] - self.tgt[:, :, None, :]                          # [E,m,n,2]
        dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
        mins = dist.min(axis=2)                                                           # [E,m]
        r = np.zeros(self.E)
        for j in range(sp.target_num):                                                    # the reference's summation order
            r = r - mins[:, j]
        r[~live] = 0.0
        self.time_step[live] += 1
        term = self.time_step >= sp.time_limit
        term[~live] = True
        self.done = term.copy()
        if self.auto_reset:
            self.reset(mask=term & live)
        return r, term
