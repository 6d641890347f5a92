"""Edge cases of the flight / search CUDA paths against the C oracle: ragged env counts, maximum agent / target
counts, degenerate limits, odd and large maps (the belief map's scalar and wide fallbacks), wide view ranges, coincident
agents (the repulsion guard of flight_env_easy.py:298), detect_prob 0 and 1, targets outside the map."""
import types

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import c_oracle
from oracle.py_envs import FlightSpec, SearchSpec
from test_gpu_flight_easy import make_args, cpu
from test_gpu_search import make_args as search_args

pytestmark = pytest.mark.gpu


def template_for(m, seed=0):
    """A synthetic target file with m rows (half of them randomised) in the reference's units (0..10)."""
    rng = np.random.default_rng(seed)
    return {"x": rng.uniform(0.5, 9.5, m).round(2).tolist(), "y": rng.uniform(0.5, 9.5, m).round(2).tolist(),
            "deter": ["f" if k % 2 == 0 else "t" for k in range(m)], "priority": [1] * m,
            "dx": rng.uniform(0.1, 0.35, m).round(2).tolist(), "dy": rng.uniform(0.1, 0.35, m).round(2).tolist()}


def run_against_oracle(spec, E, T, template, seed=5, base=1234, auto_reset=False, map_rtol=1e-5):
    import coopsearch_b200 as cs
    cls = cs.VecFlightEasyEnv if spec.variant == "easy" else cs.VecFlightEnv
    env = cls(make_args(dict(spec.__dict__)), template, num_envs=E, seed=seed, env_id_base=base, auto_reset=auto_reset,
              count_touched=(spec.variant == "probmap"), reset=False)
    orc = c_oracle.FlightBatch(spec, template, seed, base, E, auto_reset=auto_reset)
    orc.reset(init=True)
    env.reset(init=True, targets=orc.tgt)          # same float64 target coordinates on both sides
    actions = np.random.default_rng(11).integers(0, 3, size=(T, E, spec.n_agents), dtype=np.uint8)
    for t in range(T):
        r, term, win = env.step(actions[t])
        orr, ot, ow = orc.step(actions[t])
        where = "step %d" % t
        meta = cpu(env.meta).astype(np.uint32)
        assert np.array_equal(meta[:, 0], orc.found), where
        assert np.array_equal(meta[:, 2], orc.out), where
        assert np.array_equal(meta[:, 3], orc.time_step), where
        assert np.array_equal(cpu(r), orr.astype(np.float32)), where
        assert np.array_equal(cpu(term), ot) and np.array_equal(cpu(win), ow), where
        if auto_reset:
            # targets redrawn on the device (Box-Muller with CUDA log/sincos) agree to ~1e-15; keep both sides identical
            np.testing.assert_allclose(cpu(env.tgt_xy), orc.tgt, rtol=0, atol=1e-9, err_msg=where)
            env.tgt_xy.copy_(torch.from_numpy(orc.tgt).cuda())
    np.testing.assert_allclose(cpu(env.agent_xy), orc.xy, rtol=0, atol=1e-9)
    obs, state = orc.obs_state()
    np.testing.assert_allclose(cpu(env.get_obs(full=False) if spec.variant == "probmap" else env.get_obs()), obs, rtol=1e-5, atol=1e-6)
    if not auto_reset:
        np.testing.assert_allclose(cpu(env.get_state()), state, rtol=1e-5, atol=1e-6)
    if spec.variant == "probmap":
        np.testing.assert_allclose(cpu(env.prob_map), orc.map.astype(np.float32), rtol=map_rtol, atol=1e-37)
        assert env.stats()["map_cells_touched"] == float(orc.touched[0])
    return env, orc


@pytest.mark.parametrize("E", [1, 2, 7, 33, 4097])
def test_ragged_env_counts(E):
    run_against_oracle(FlightSpec(n_agents=3, time_limit=60), E, 70, gu.TEMPLATE, auto_reset=True)


@pytest.mark.parametrize("n,m", [(32, 32), (1, 1), (17, 5), (2, 32), (32, 3)])
def test_extreme_agent_and_target_counts(n, m):
    spec = FlightSpec(n_agents=n, target_num=m, agent_mode=1, target_mode=0, time_limit=80)
    run_against_oracle(spec, 24, 80, template_for(m))


def test_crowded_agents_repulsion_and_walls(oracle_squares_by_multiplication):
    """16 agents on a 10x10 map: the sequential repulsion path and the wall reflection fire on every step."""
    spec = FlightSpec(n_agents=16, target_num=8, map_size=10, view_range=2, agent_mode=1, target_mode=1, time_limit=150)
    run_against_oracle(spec, 40, 150, None)


def test_episode_counter_beyond_16_bits():
    """Episodes 65536 + k use their own Philox streams (cs_stream_key): device and oracle agree there too, and the
    draws differ from episode k's."""
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=3, time_limit=40)
    E = 64
    env = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=3, auto_reset=True)
    low = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=3, auto_reset=True)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, 3, 0, E, auto_reset=True)
    env._meta_word(4).fill_(65536 + 2)                # the next reset opens episode 65539
    low._meta_word(4).fill_(2)
    orc.meta[:, 4] = 65536 + 2
    env.reset(); low.reset(); orc.reset()
    assert np.array_equal(cpu(env.tgt_xy), orc.tgt) or np.allclose(cpu(env.tgt_xy), orc.tgt, rtol=0, atol=1e-9)
    assert not torch.equal(env.tgt_xy, low.tgt_xy)
    acts = np.random.default_rng(0).integers(0, 3, size=(90, E, 3), dtype=np.uint8)
    for t in range(90):
        r, term, _ = env.step(acts[t])
        orr, ot, _ = orc.step(acts[t])
        assert np.array_equal(cpu(env.found_mask).astype(np.uint32), orc.found), t
        assert np.array_equal(cpu(r), orr.astype(np.float32)) and np.array_equal(cpu(term), ot), t


def test_coincident_agents_guard(oracle_squares_by_multiplication):
    """Agents at exactly the same point exert no force on each other (flight_env_easy.py:298)."""
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=4, agent_mode=0, time_limit=30)
    env = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=8, seed=1)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, 1, 0, 8)
    orc.reset(init=True)
    env.reset(init=True, targets=orc.tgt)
    xy = np.tile(np.array([[20.0, 20.0], [20.0, 20.0], [20.5, 20.0], [20.0, 20.0]]), (8, 1, 1))
    env.agent_xy.copy_(torch.from_numpy(xy).cuda())
    orc.xy[...] = xy
    acts = np.zeros((8, 4), np.uint8)
    for t in range(20):
        env.step(acts); orc.step(acts)
        assert np.array_equal(cpu(env.agent_xy), orc.xy), "step %d" % t      # bit for bit, incl. the divisions


@pytest.mark.parametrize("d", [0.0, 1.0, 0.5])
def test_detect_prob_limits(d):
    spec = FlightSpec(n_agents=3, detect_prob=d, time_limit=60)
    env, orc = run_against_oracle(spec, 64, 60, gu.TEMPLATE)
    if d == 0.0:
        assert int(env.target_find.sum().item()) == 0


def test_time_limit_one_and_auto_reset_every_step():
    spec = FlightSpec(n_agents=2, time_limit=1)
    env, orc = run_against_oracle(spec, 16, 12, gu.TEMPLATE, auto_reset=True)
    assert env.stats()["episodes"] == 16 * 12


@pytest.mark.parametrize("M,R,n", [(51, 7, 3), (50, 12, 2), (63, 7, 3), (64, 7, 3), (100, 9, 2), (20, 15, 2), (6, 2, 3)])
def test_probability_map_geometries(M, R, n):
    """Odd map (scalar sweep), view range wider than 16 columns (column-chunk loop), the 63/64 boundary of the
    bit-row path, a map larger than 64 (per-cell fallback), a disc larger than the map, a tiny map."""
    spec = FlightSpec(n_agents=n, target_num=6, map_size=M, view_range=R, agent_mode=1, target_mode=1, time_limit=50,
                      variant="probmap")
    run_against_oracle(spec, 12, 50, None)


def test_targets_outside_the_map_and_on_cell_borders():
    """Targets may lie outside [0, M] (file targets are not clamped, flight_env_easy.py:107-113); the cell forced to 1
    uses int() truncation and the min(.., M-1) clamp (flight_env.py:279); negative indices never match."""
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=2, target_num=6, map_size=20, view_range=6, agent_mode=1, time_limit=40, variant="probmap",
                      detect_prob=1.0)
    E = 4
    tg = np.tile(np.array([[-0.5, 10.0], [-1.5, 10.2], [20.0, 10.0], [23.0, 9.5], [10.0, 10.0], [0.0, 10.0]]), (E, 1, 1))
    env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), None, num_envs=E, seed=3, reset=False, count_touched=True)
    orc = c_oracle.FlightBatch(spec, None, 3, 0, E)
    orc.reset(targets=tg, init=True)
    env.reset(init=True, targets=tg)
    acts = np.random.default_rng(2).integers(0, 3, size=(40, E, 2), dtype=np.uint8)
    for t in range(40):
        env.step(acts[t]); orc.step(acts[t])
        assert np.array_equal(cpu(env.found_mask).astype(np.uint32), orc.found)
    got, want = cpu(env.prob_map), orc.map.astype(np.float32)
    finite = np.isfinite(want)                       # detect_prob = 1 makes 0/0 on p = 1 cells, like the reference
    assert np.array_equal(np.isfinite(got), finite)
    np.testing.assert_allclose(got[finite], want[finite], rtol=1e-5, atol=1e-37)


@pytest.mark.parametrize("n,m,M,R,am,tm", [(1, 1, 8, 2, 0, 0), (2, 64, 8, 1, 2, 0), (100, 300, 33, 3, 0, 1), (7, 50, 127, 9, 1, 0),
                                           (5, 40, 65, 16, 0, 0)])
def test_search_geometries(n, m, M, R, am, tm):
    """Single agent, a full grid of targets, >64 agents, multi-word bit rows (M > 64: direct observation path), odd M,
    a window wider than 32 columns."""
    import coopsearch_b200 as cs
    E, T, seed, base = 9, 60, 4, 50
    spec = SearchSpec(n_agents=n, target_num=m, map_size=M, view_range=R, agent_mode=am, target_mode=tm)
    env = cs.VecSearchEnv(search_args(n, m, M, R, am, tm), num_envs=E, seed=seed, env_id_base=base)
    orc = c_oracle.SearchBatch(spec, seed, base, E)
    orc.reset()
    for t in range(T):
        r, term, _ = env.step_random(1)
        orr, ot = orc.step(None)
        assert np.array_equal(cpu(env.agent_pos), orc.pos), "step %d" % t
        np.testing.assert_allclose(cpu(r), orr.astype(np.float32), rtol=1e-5, atol=1e-6)
        assert np.array_equal(cpu(term), ot)
    obs, state, avail = orc.views()
    assert np.array_equal(cpu(env.get_obs()), obs)
    assert np.array_equal(cpu(env.get_state()), state)
    assert np.array_equal(cpu(env.get_avail_actions()), avail)
    assert np.array_equal(cpu(env.freq_map), orc.freq)


def test_two_live_search_handles_of_different_size_keep_working():
    """Dynamic shared-memory limits are per kernel: a second, smaller handle must not break the first (> 48 KB) one."""
    import coopsearch_b200 as cs
    big = cs.VecSearchEnv(search_args(64, 1000, 64, 7, 0, 0), num_envs=8, seed=1)
    small = cs.VecSearchEnv(search_args(2, 5, 10, 2, 0, 0), num_envs=4, seed=1)
    big.step_random(3); small.step_random(3); big.step_random(3)
    import torch
    torch.cuda.synchronize()


def test_search_target_modes_2_and_3(tmp_path):
    """target_mode 2 (circle dictionary, main.py:13-15) and 3 (cell file, search_env.py:127-138)."""
    import coopsearch_b200 as cs
    circle = {'circle_center': [[10, 32], [12, 15]], 'circle_radius': [5, 7], 'target_num': [7, 8]}
    env = cs.VecSearchEnv(search_args(3, 15, 50, 7, 0, 2), circle_dict=circle, num_envs=5, seed=8)
    tb = cpu(env.target_bits).astype(np.uint32)
    dense = ((tb[:, :, :, None] >> np.arange(32)[None, None, None, :]) & 1).reshape(5, 50, -1)[:, :, :50]
    assert (dense.sum(axis=(1, 2)) == 15).all()
    xs, ys = np.nonzero(dense[0])
    in0 = (np.abs(xs - 10) <= 5) & (np.abs(ys - 32) <= 5)
    in1 = (np.abs(xs - 12) <= 7) & (np.abs(ys - 15) <= 7)
    assert (in0 | in1).all()
    path = tmp_path / "cells.txt"
    cells = [[1, 2], [3, 4], [10, 10], [49, 0]]
    path.write_text("".join("%d %d\n" % (x, y) for x, y in cells))
    env3 = cs.VecSearchEnv(search_args(2, 4, 50, 7, 1, 3), targets_filename=str(path), num_envs=3)
    tb = cpu(env3.target_bits).astype(np.uint32)
    dense = ((tb[:, :, :, None] >> np.arange(32)[None, None, None, :]) & 1).reshape(3, 50, -1)[:, :, :50]
    for e in range(3):
        assert sorted(map(list, np.argwhere(dense[e]))) == sorted(cells)
    # every reset of a mode-2 env draws a fresh layout (search_env.py:106-123); mode 2 / 3 refuse the in-kernel auto-reset
    before = cpu(env.target_bits).copy()
    env.reset()
    assert not np.array_equal(before, cpu(env.target_bits))
    with pytest.raises(Exception, match="auto_reset=False"):
        cs.VecSearchEnv(search_args(3, 15, 50, 7, 0, 2), circle_dict=circle, num_envs=2, auto_reset=True)
    with pytest.raises(Exception, match="No circle dictionary"):
        cs.VecSearchEnv(search_args(3, 15, 50, 7, 0, 2), num_envs=1)
    with pytest.raises(Exception, match="No target file"):
        cs.VecSearchEnv(search_args(3, 15, 50, 7, 0, 3), num_envs=1)
