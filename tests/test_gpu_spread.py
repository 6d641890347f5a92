"""simple_spread on the GPU (csrc/spread.cu) against the golden trajectories of the unmodified reference and against the
oracle at size.  Observations and states are float32 casts of float64 differences (exact: compared with array_equal); the
reward is a float64 sum of square roots -- the reference squares through libm pow, the kernel multiplies: tolerance 4e-16
relative on the float64 reward, the float32 reward equal after the cast up to one float32 ulp."""
import os

import numpy as np
import pytest
import torch

from oracle.py_envs import SpreadBatch, SpreadSpec
from oracle.refharness import make_args

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
cpu = lambda t: t.detach().cpu().numpy()


def spread_args(spec):
    return make_args("simple_spread", n_agents=spec.n_agents, target_num=spec.target_num, map_size=spec.map_size)


@pytest.mark.parametrize("name", ["spread_3a3t", "spread_5a7t_small"])
def test_golden_trajectories_of_the_reference(name):
    import coopsearch_b200 as cs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, m, M, base = [int(v) for v in g["meta"]]
    spec = SpreadSpec(n_agents=n, target_num=m, map_size=M)
    T, E = g["actions"].shape[:2]
    env = cs.VecSimpleSpreadEnv(spread_args(spec), num_envs=E, seed=0, env_id_base=base, reset=False)
    info = env.get_env_info()
    assert tuple(int(v) for v in g["info"]) == (info["n_actions"], info["state_shape"], info["obs_shape"], info["episode_limit"])
    env.reset(targets=g["tgt"], agents=g["agents0"])
    assert np.array_equal(cpu(env.get_obs()), g["init_obs"].astype(np.float32))
    assert np.array_equal(cpu(env.get_state()), g["init_state"].astype(np.float32))
    for t in range(T):
        r, term, win = env.step(g["actions"][t])
        where = (name, t)
        assert np.array_equal(cpu(env.get_obs()), g["obs"][t].astype(np.float32)), where
        assert np.array_equal(cpu(env.get_state()), g["state"][t].astype(np.float32)), where
        assert np.array_equal(cpu(env.agent_xy), g["agents"][t]), where                           # float64, bit for bit
        np.testing.assert_allclose(cpu(env.reward64), g["reward"][t], rtol=4e-16, atol=0, err_msg=str(where))
        np.testing.assert_allclose(cpu(r), g["reward"][t].astype(np.float32), rtol=1.2e-7, atol=0, err_msg=str(where))
        assert np.array_equal(cpu(term), g["terminated"][t]) and not bool(win.any()), where
        assert np.array_equal(cpu(env.occupied), g["occupied"][t]), where
    # a finished env is a masked no-op until it is reset
    before = cpu(env.agent_xy).copy()
    r, term, _ = env.step(g["actions"][0])
    assert np.array_equal(cpu(env.agent_xy), before) and bool((cpu(term) == 1).all()) and bool((cpu(r) == 0).all())


@pytest.mark.parametrize("n,m,M,E,auto_reset", [(3, 3, 50, 4096, True), (1, 1, 5, 33, True), (16, 32, 40, 257, False), (5, 7, 12, 1000, True)])
def test_device_reset_and_steps_match_the_oracle(n, m, M, E, auto_reset):
    """Keyed reset placement, 230 steps through two episode ends (in-call auto-reset where enabled), device actions."""
    import coopsearch_b200 as cs
    spec = SpreadSpec(n_agents=n, target_num=m, map_size=M)
    env = cs.VecSimpleSpreadEnv(spread_args(spec), num_envs=E, seed=9, env_id_base=700, auto_reset=auto_reset)
    orc = SpreadBatch(spec, 9, 700, E, auto_reset=auto_reset)
    orc.reset()
    assert np.array_equal(cpu(env.tgt_xy), orc.tgt) and np.array_equal(cpu(env.agent_xy), orc.agents)      # M * u53: exact
    rng = np.random.default_rng(5)
    for t in range(230):
        act = rng.integers(0, 5, size=(E, n), dtype=np.uint8)
        r, term, _ = env.step(torch.from_numpy(act).cuda())
        orr, ot = orc.step(act)
        where = "step %d" % t
        np.testing.assert_allclose(cpu(env.reward64), orr, rtol=1e-15, atol=0, err_msg=where)
        assert np.array_equal(cpu(term).astype(bool), ot), where
        assert np.array_equal(cpu(env.agent_xy), orc.agents) and np.array_equal(cpu(env.tgt_xy), orc.tgt), where
        if t % 23 == 0 or t in (99, 100, 199, 200):
            obs, state = orc.obs_state()
            assert np.array_equal(cpu(env.get_obs()), obs.astype(np.float32)), where
            assert np.array_equal(cpu(env.get_state()), state.astype(np.float32)), where
        assert np.array_equal(cpu(env.time_step), orc.time_step), where
    st = env.stats()
    assert st["episodes"] == (2 * E if auto_reset else E) and st["episode_len_sum"] == 100 * st["episodes"]


def test_shards_protocol_errors_and_host_step():
    import coopsearch_b200 as cs
    spec = SpreadSpec(n_agents=3, target_num=3, map_size=50)
    args = spread_args(spec)
    whole = cs.VecSimpleSpreadEnv(args, num_envs=64, seed=3, env_id_base=1000, auto_reset=True)
    lo = cs.VecSimpleSpreadEnv(args, num_envs=32, seed=3, env_id_base=1000, auto_reset=True)
    hi = cs.VecSimpleSpreadEnv(args, num_envs=32, seed=3, env_id_base=1032, auto_reset=True)
    for t in range(120):                                   # the in-kernel random policy is keyed by the global env id
        whole.step_random(1); lo.step_random(1); hi.step_random(1)
    assert torch.equal(whole.get_obs(), torch.cat([lo.get_obs(), hi.get_obs()])) and torch.equal(whole.agent_xy, torch.cat([lo.agent_xy, hi.agent_xy]))
    with pytest.raises(Exception, match="Act num mismatch agent"):
        lo.step(np.zeros((32, 2), np.uint8))
    with pytest.raises(Exception, match="Agent id out of range"):
        lo.get_avail_agent_actions(3)
    with pytest.raises(IndexError):
        lo.step(np.full((32, 3), 5))
    assert lo.get_avail_agent_actions(0).shape == (32, 5) and bool((lo.get_avail_actions() == 1).all())
    # host-buffer step == device step
    a = cs.VecSimpleSpreadEnv(args, num_envs=50, seed=1)
    b = cs.VecSimpleSpreadEnv(args, num_envs=50, seed=1)
    out = {"reward": torch.empty(50).pin_memory(), "terminated": torch.empty(50, dtype=torch.uint8).pin_memory(),
           "obs": torch.empty(50, 3, a.obs_shape).pin_memory(), "state": torch.empty(50, a.state_shape).pin_memory()}
    rng = np.random.default_rng(2)
    for t in range(10):
        act = rng.integers(0, 5, size=(50, 3), dtype=np.uint8)
        r, term, _ = a.step(act)
        b.step_host(act, out)
        assert np.array_equal(cpu(r), out["reward"].numpy()) and np.array_equal(cpu(a.get_obs()), out["obs"].numpy())
        assert np.array_equal(cpu(a.get_state()), out["state"].numpy()) and np.array_equal(cpu(term), out["terminated"].numpy())
