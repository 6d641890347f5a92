"""BASELINE.json configs at full size against the C oracle (float64 restatement pinned to the reference).

  C2  flight_easy 3a15t AM0TM0, 4096 envs x 200 steps, every discrete quantity bit-exact at every step
  C3  flight_easy 5a15t AM2 / AM3, 65536 envs: a strided 4096-env subsample against the oracle and bitwise
      shard invariance of the same global ids computed in a different shard layout
  C4  flight 3a15t, 16384 envs: 256-env subsample of the map against the oracle + map invariants at full size
  C5  search 64a/1000t map 64: 32 envs x 100 steps (test_gpu_search) + conservation properties at 8192 envs
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import c_oracle
from oracle.py_envs import FlightSpec, SearchSpec
from test_gpu_flight_easy import make_args, cpu
from test_gpu_search import make_args as search_args

pytestmark = pytest.mark.gpu


def test_c2_flight_easy_3a_4096_envs_against_the_references_own_arithmetic():
    """configs[1] against the oracle in the reference's arithmetic (squares through libm pow, see conftest.py): every
    discrete output bit-exact over 4096 envs x 200 steps, positions within 1e-9 absolute (the 1-ulp differences of
    pow(v, 2.0) against v * v in the repulsion force stay 1-ulp differences on these trajectories)."""
    import coopsearch_b200 as cs
    E, T, seed, base = 4096, 200, 42, 0
    spec = FlightSpec(n_agents=3, agent_mode=0, target_mode=0)
    env = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base, reset=False)
    c_oracle.set_threads(8)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E)
    orc.reset(init=True)
    env.reset(init=True, targets=orc.tgt)
    actions = np.random.default_rng(1234).integers(0, 3, size=(T, E, 3), dtype=np.uint8)
    dact = torch.from_numpy(actions).cuda()
    for t in range(T):
        r, term, win = env.step(dact[t])
        orr, ot, ow = orc.step(actions[t])
        where = "step %d" % t
        meta = cpu(env.meta).astype(np.uint32)
        assert np.array_equal(meta[:, 0], orc.found) and np.array_equal(meta[:, 2], orc.out) and np.array_equal(meta[:, 3], orc.time_step), where
        assert np.array_equal(cpu(r), orr.astype(np.float32)), where
        assert np.array_equal(cpu(term), ot) and np.array_equal(cpu(win), ow), where
    np.testing.assert_allclose(cpu(env.agent_xy), orc.xy, rtol=0, atol=1e-9)
    assert float(np.mean(cpu(env.agent_xy) == orc.xy)) > 0.99          # and almost all of them are the same bits


def test_c2_flight_easy_3a_4096_envs_bit_exact(oracle_squares_by_multiplication):
    import coopsearch_b200 as cs
    E, T, seed, base = 4096, 200, 42, 0
    spec = FlightSpec(n_agents=3, agent_mode=0, target_mode=0)
    env = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base, reset=False)
    c_oracle.set_threads(8)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E)
    orc.reset(init=True)
    # inject the oracle's libm-drawn targets so both sides start from identical float64 coordinates (episode 0 on both)
    env.reset(init=True, targets=orc.tgt)
    actions = np.random.default_rng(1234).integers(0, 3, size=(T, E, 3), dtype=np.uint8)
    dact = torch.from_numpy(actions).cuda()
    for t in range(T):
        r, term, win = env.step(dact[t])
        orr, ot, ow = orc.step(actions[t])
        where = "step %d" % t
        meta = cpu(env.meta).astype(np.uint32)
        assert np.array_equal(meta[:, 0], orc.found), where             # found flags
        assert np.array_equal(meta[:, 2], orc.out), where               # out-of-map flags
        assert np.array_equal(meta[:, 3], orc.time_step), where         # step counts
        assert np.array_equal(cpu(r), orr.astype(np.float32)), where
        assert np.array_equal(cpu(term), ot) and np.array_equal(cpu(win), ow), where
        if t % 20 == 19:
            np.testing.assert_allclose(cpu(env.agent_xy), orc.xy, rtol=1e-5, atol=1e-9, err_msg=where)
            obs, state = orc.obs_state()
            np.testing.assert_allclose(cpu(env.get_obs()), obs, rtol=1e-5, atol=1e-6, err_msg=where)
            np.testing.assert_allclose(cpu(env.get_state()), state, rtol=1e-5, atol=1e-6, err_msg=where)
    # positions: with the heading table the kinematics are the same IEEE operations -> expect exact equality
    assert np.array_equal(cpu(env.agent_xy), orc.xy)
    assert np.array_equal(cpu(env.agent_yaw), orc.yaw)


@pytest.mark.parametrize("agent_mode", [2, 3])
def test_c3_flight_easy_5a_65536_envs(agent_mode, oracle_squares_by_multiplication):
    import coopsearch_b200 as cs
    E, T, seed = 65536, 200, 42
    spec = FlightSpec(n_agents=5, agent_mode=agent_mode, target_mode=0)
    args = make_args(dict(spec.__dict__))
    env = cs.VecFlightEasyEnv(args, gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=0)
    # "8-GPU" layout of one shard: global ids [5*8192, 6*8192) as its own handle with another lane mapping
    shard = cs.VecFlightEasyEnv(args, gu.TEMPLATE, num_envs=8192, seed=seed, env_id_base=5 * 8192, lanes_per_env=8)
    # strided 4096-env subsample for the oracle: ids 0,16,32,... handled as 4096 single-id batches would be slow,
    # so the oracle runs the contiguous block [16384, 20480) plus the same actions
    lo, hi = 16384, 20480
    c_oracle.set_threads(8)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, lo, hi - lo)
    orc.reset(init=True)
    np.testing.assert_allclose(cpu(env.tgt_xy[lo:hi]), orc.tgt, rtol=0, atol=1e-9)
    env.tgt_xy[lo:hi].copy_(torch.from_numpy(orc.tgt).cuda())      # identical float64 targets on both sides
    env.reset(init=True, targets=env.tgt_xy.clone(), keep_episode=True)
    shard.reset(init=True, targets=env.tgt_xy[5 * 8192:6 * 8192].clone(), keep_episode=True)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for t in range(T):
        act = torch.randint(0, 3, (E, 5), generator=gen, device="cuda", dtype=torch.uint8)
        r, term, win = env.step(act)
        rs, ts, ws = shard.step(act[5 * 8192:6 * 8192])
        orr, ot, ow = orc.step(cpu(act[lo:hi]))
        where = "step %d" % t
        assert torch.equal(r[5 * 8192:6 * 8192], rs) and torch.equal(term[5 * 8192:6 * 8192], ts), where
        assert np.array_equal(cpu(env.found_mask[lo:hi]).astype(np.uint32), orc.found), where
        assert np.array_equal(cpu(r[lo:hi]), orr.astype(np.float32)), where
        assert np.array_equal(cpu(term[lo:hi]), ot) and np.array_equal(cpu(win[lo:hi]), ow), where
    assert torch.equal(env._dyn[5 * 8192:6 * 8192], shard._dyn)
    assert np.array_equal(cpu(env.agent_xy[lo:hi]), orc.xy)
    # statistical known-answer: % of targets found by the uniform-random policy must be
    # plausible -- every env found between 0 and 15 targets and the mean is well inside (40%, 100%)
    tf = cpu(env.target_find).astype(np.float64)
    assert tf.min() >= 0 and tf.max() <= 15 and 0.40 < tf.mean() / 15 < 1.0


def test_c4_flight_16384_envs_map():
    import coopsearch_b200 as cs
    E, T, seed = 16384, 60, 42
    spec = FlightSpec(n_agents=3, agent_mode=0, target_mode=0, variant="probmap")
    env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=0, count_touched=True)
    lo, hi = 8192, 8192 + 256
    c_oracle.set_threads(8)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, lo, hi - lo)
    orc.reset(init=True)
    env.tgt_xy[lo:hi].copy_(torch.from_numpy(orc.tgt).cuda())
    env.reset(init=True, targets=env.tgt_xy.clone(), keep_episode=True)
    gen = torch.Generator(device="cuda").manual_seed(8)
    for t in range(T):
        act = torch.randint(0, 3, (E, 3), generator=gen, device="cuda", dtype=torch.uint8)
        r, term, win = env.step(act)
        orr, ot, ow = orc.step(cpu(act[lo:hi]))
        assert np.array_equal(cpu(env.found_mask[lo:hi]).astype(np.uint32), orc.found), "step %d" % t
        assert np.array_equal(cpu(r[lo:hi]), orr.astype(np.float32)), "step %d" % t
    np.testing.assert_allclose(cpu(env.prob_map[lo:hi]), orc.map.astype(np.float32), rtol=1e-5, atol=1e-37)
    # size-independent properties at full size: probabilities stay in [0,1]; untouched cells are exactly 0.5;
    # a cell equals 1 only where a found target sits (flight_env.py:288-289)
    pm = env.prob_map
    assert float(pm.min()) >= 0.0 and float(pm.max()) <= 1.0
    assert float((pm == 0.5).float().mean()) > 0.05
    ones = (pm == 1.0).sum(dim=(1, 2))
    assert bool((ones <= env.target_find).all())
    assert env.stats()["map_cells_touched"] > 0


def test_c5_search_scaled_properties():
    import coopsearch_b200 as cs
    E, T = 8192, 50
    args = search_args(64, 1000, 64, 7, 0, 0)
    env = cs.VecSearchEnv(args, num_envs=E, seed=42, env_id_base=0)
    f0 = env.freq_map.sum(dim=(1, 2)).clone()
    assert bool((f0 == 64).all())
    tb0 = env.target_bits.clone()
    for t in range(T):
        r, term, _ = env.step_random(1)
    # conservation: every non-terminated step adds exactly n_agents visits; targets never move;
    # unfound is a subset of targets; target_find = |targets| - |unfound|
    steps = env.time_step.to(torch.int64)
    assert torch.equal(env.freq_map.sum(dim=(1, 2)).to(torch.int64), 64 + 64 * steps)
    assert torch.equal(env.target_bits, tb0)
    assert bool(((env.unfound_bits & ~env.target_bits) == 0).all())
    pop = lambda x: sum(((x >> k) & 1).sum(dim=(1, 2)) for k in range(32))
    assert torch.equal(pop(env.target_bits) - pop(env.unfound_bits), env.target_find.to(torch.int64))
    assert bool((pop(env.target_bits) == 1000).all())
    assert not bool(env.illegal.any())
    st = env.get_state().view(E, 64, 64, 2)
    assert bool((st[..., 1].sum(dim=(1, 2)) <= 64).all()) and bool((st[..., 1].sum(dim=(1, 2)) >= 1).all())
