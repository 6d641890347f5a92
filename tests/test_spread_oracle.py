"""simple_spread (env/simple_spread.py): the oracle restatements against golden trajectories of the unmodified reference
(tests/golden/make_golden.py: run_spread) -- bit for bit in float64 for the scalar oracle, which uses the reference's own
arithmetic; the numpy-batched oracle (squares by multiplication) to one ulp of the float64 reward."""
import os

import numpy as np
import pytest

from oracle.py_envs import SimpleSpreadOracle, SpreadBatch, SpreadSpec, keyed_spread_layout

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, m, M, base = [int(v) for v in g["meta"]]
    return g, SpreadSpec(n_agents=n, target_num=m, map_size=M), base


@pytest.mark.parametrize("name", ["spread_3a3t", "spread_5a7t_small"])
def test_scalar_oracle_equals_reference_bit_for_bit(name):
    g, spec, base = load(name)
    assert tuple(int(v) for v in g["info"]) == (spec.n_actions, spec.state_shape, spec.obs_shape, spec.time_limit)
    T, E = g["actions"].shape[:2]
    for e in range(E):
        env = SimpleSpreadOracle(spec, seed=0, env_id=base + e)
        env.reset(targets=g["tgt"][e], agents=g["agents0"][e])
        assert np.array_equal(env.get_obs(), g["init_obs"][e]) and np.array_equal(env.get_state(), g["init_state"][e])
        for t in range(T):
            r, term, win = env.step([int(a) for a in g["actions"][t, e]])
            where = (name, e, t)
            assert r == g["reward"][t, e] and int(term) == g["terminated"][t, e] and win is False, where
            assert np.array_equal(env.get_obs(), g["obs"][t, e]) and np.array_equal(env.get_state(), g["state"][t, e]), where
            assert np.array_equal(np.array(env.agents), g["agents"][t, e]), where
            assert np.array_equal(np.array(env.occupied, np.uint8), g["occupied"][t, e]), where
        assert g["terminated"][-1, e] == 1 and g["terminated"][-2, e] == 0          # time_limit = 100 (:25)


@pytest.mark.parametrize("name", ["spread_3a3t", "spread_5a7t_small"])
def test_batched_oracle_equals_reference(name):
    g, spec, base = load(name)
    T, E = g["actions"].shape[:2]
    b = SpreadBatch(spec, 0, base, E)
    b.reset()
    b.tgt[...] = g["tgt"]
    b.agents[...] = g["agents0"]
    obs, state = b.obs_state()
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(state, g["init_state"])
    for t in range(T):
        r, term = b.step(g["actions"][t])
        obs, state = b.obs_state()
        assert np.array_equal(obs, g["obs"][t]) and np.array_equal(state, g["state"][t]), t      # differences only: exact
        np.testing.assert_allclose(r, g["reward"][t], rtol=4e-16, atol=0)                        # pow(d, 2) against d * d
        assert np.array_equal(term.astype(np.uint8), g["terminated"][t]), t


def test_keyed_layout_is_inside_the_map_and_keyed():
    spec = SpreadSpec(n_agents=4, target_num=5, map_size=30)
    t0, a0 = keyed_spread_layout(spec, 7, 11, 0)
    t1, a1 = keyed_spread_layout(spec, 7, 11, 1)
    t2, _ = keyed_spread_layout(spec, 7, 12, 0)
    assert len(t0) == 5 and len(a0) == 4 and all(0 <= v < 30 for p in t0 + a0 for v in p)
    assert t0 != t1 and t0 != t2 and (t0, a0) == keyed_spread_layout(spec, 7, 11, 0)
