"""Pins the C oracle (oracle/coopsearch_oracle.c) against the reference goldens and the Python oracle."""
import numpy as np
import pytest

import golden_util as gu
from oracle import c_oracle
from oracle.py_envs import FlightOracle, FlightSpec, SearchOracle, SearchSpec


@pytest.mark.parametrize("name", gu.EASY_FIXTURES + gu.FLIGHT_FIXTURES)
def test_c_oracle_flight_matches_reference_golden(name):
    variant = "easy" if name.startswith("easy") else "probmap"
    g = gu.load(name)
    kw, base, seed = gu.flight_spec_kwargs(g, variant)
    spec = FlightSpec(**kw)
    T, E = g["reward"].shape
    thin = int(g["thin"][0])
    b = c_oracle.FlightBatch(spec, None, seed, base, E)
    b.reset(targets=g["tgt_xy"], init=True)
    assert np.array_equal(b.found, g["init_found"])
    assert np.array_equal(b.xy, g["init_xy"]) and np.array_equal(b.yaw, g["init_yaw"])
    obs, state = b.obs_state()
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(state, g["init_state"])
    for t in range(T):
        r, term, win = b.step(g["actions"][t])
        where = (name, t)
        assert np.array_equal(b.found, g["found"][t]), where
        assert np.array_equal(r, g["reward"][t]), where
        assert np.array_equal(term, g["terminated"][t]) and np.array_equal(win, g["win"][t]), where
        assert np.array_equal(b.time_step, g["time_step"][t]), where
        # float64 state is bit-identical: glibc cos/sin/pow are what numpy calls for Python floats
        assert np.array_equal(b.xy, g["xy"][t]) and np.array_equal(b.yaw, g["yaw"][t]), where
        out = (b.out[:, None] >> np.arange(spec.n_agents)[None, :]) & 1
        assert np.array_equal(out.astype(np.uint8), g["out"][t]), where
        if (t + 1) % thin == 0:
            obs, state = b.obs_state()
            k = (t + 1) // thin - 1
            assert np.array_equal(obs, g["obs"][k]) and np.array_equal(state, g["state"][k]), where
        if variant == "probmap" and (t + 1) in g["map_steps"]:
            assert np.array_equal(b.map, g["maps"][list(g["map_steps"]).index(t + 1)]), where
    if variant == "probmap":
        b.meta[:, 5] &= ~np.uint32(2)          # the reference keeps stepping after `terminated`; so does this replay
        b.reset(targets=g["ep2_tgt_xy"], init=False)
        for t in range(g["ep2_actions"].shape[0]):
            r, _, _ = b.step(g["ep2_actions"][t])
            assert np.array_equal(b.found, g["ep2_found"][t]) and np.array_equal(r, g["ep2_reward"][t])
        assert np.array_equal(b.map, g["ep2_map"])


@pytest.mark.parametrize("name", gu.SEARCH_FIXTURES)
def test_c_oracle_search_matches_reference_golden(name):
    g = gu.load(name)
    n, m, M, R, am, tm, base = [int(v) for v in g["meta"]]
    spec = SearchSpec(n_agents=n, target_num=m, map_size=M, view_range=R, agent_mode=am, target_mode=tm)
    T, E = g["reward"].shape
    thin = int(g["thin"][0])
    b = c_oracle.SearchBatch(spec, 0, base, E)
    b.reset(cells=g["cells"])
    assert np.array_equal(b.pos, g["init_pos"]) and np.array_equal(b.freq, g["init_freq"])
    obs, state, avail = b.views()
    assert np.array_equal(obs, g["init_obs"].astype(np.float32)) and np.array_equal(state, g["init_state"].astype(np.float32))
    for t in range(T):
        _, _, avail = b.views(want_obs=False, want_state=False)
        assert np.array_equal(avail, g["avail"][t]), (name, t)
        r, term = b.step(g["actions"][t])
        assert np.array_equal(r, g["reward"][t]), (name, t)
        assert np.array_equal(term, g["terminated"][t]), (name, t)
        assert np.array_equal(b.pos, g["pos"][t]) and np.array_equal(b.found, g["found"][t]), (name, t)
        assert np.array_equal(b.counters[:, 0], g["target_find"][t])
        if (t + 1) % thin == 0:
            obs, state, _ = b.views()
            k = (t + 1) // thin - 1
            assert np.array_equal(obs, g["obs"][k]) and np.array_equal(state, g["state"][k].astype(np.float32)), (name, t)
    assert np.array_equal(b.freq, g["freq"])


@pytest.mark.parametrize("variant,n,am,tm", [("easy", 5, 2, 0), ("easy", 3, 0, 1), ("probmap", 3, 0, 0), ("probmap", 2, 3, 1)])
def test_c_oracle_equals_python_oracle_with_keyed_reset(variant, n, am, tm):
    """Device-style keyed target placement + random-policy stream + auto-reset: C vs Python, bit for bit."""
    E, T, seed, base = 6, 70, 17, 300
    spec = FlightSpec(n_agents=n, agent_mode=am, target_mode=tm, time_limit=50, variant=variant,
                      map_size=30 if variant == "probmap" else 50, view_range=5 if variant == "probmap" else 7)
    c_oracle.set_threads(3)
    b = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E, auto_reset=True)
    b.reset(init=True)
    pys = [FlightOracle(spec, gu.TEMPLATE, seed, base + e) for e in range(E)]
    for o in pys:
        o.reset(init=True)
    assert np.array_equal(b.tgt, np.array([o.tgt for o in pys]))
    actions = np.random.default_rng(0).integers(0, 3, size=(T, E, n), dtype=np.uint8)
    for t in range(T):
        r, term, win = b.step(actions[t])
        for e, o in enumerate(pys):
            pr, pterm, pwin = o.step(actions[t, e])
            assert (pr, int(pterm), int(pwin)) == (r[e], term[e], win[e]), (t, e)
            if pterm:
                o.reset(init=False)                      # what auto_reset does inside the same call
            assert o.found_mask() == b.found[e] and o.time_step == b.time_step[e], (t, e)
            assert np.array_equal(np.array(o.pos, float), b.xy[e]), (t, e)
            if variant == "probmap":
                assert np.array_equal(o.prob_map, b.map[e]), (t, e)
    c_oracle.set_threads(1)


def test_c_oracle_search_equals_python_oracle_keyed_placement():
    spec = SearchSpec(n_agents=5, target_num=40, map_size=24, view_range=4, agent_mode=0, target_mode=1)
    b = c_oracle.SearchBatch(spec, 5, 70, 3)
    b.reset()
    for e in range(3):
        o = SearchOracle(spec, 5, 70 + e)
        o.reset()
        assert np.array_equal(np.array(o.cells), b.cells[e])
        assert np.array_equal(o.target_map.astype(np.uint8), b.tmap[e])
        assert np.array_equal(np.array(o.pos), b.pos[e])


def test_c_oracle_philox_kat():
    import ctypes as C
    out = np.zeros(4, np.uint32)
    inp = np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0], np.uint32)
    c_oracle.lib().of_philox(inp.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
