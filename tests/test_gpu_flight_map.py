"""GPU parity of the flight (probability-map) variant: reference goldens + C oracle.

Map bar: rtol 1e-5, atol 1e-37 against the float64 reference cast to float32 (north_star tolerance;
SURVEY.md section 7 hard part 2).  Everything discrete is bit-exact."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import c_oracle
from oracle.py_envs import FlightSpec
from test_gpu_flight_easy import make_args, cpu

pytestmark = pytest.mark.gpu
MAP_RTOL, MAP_ATOL = 1e-5, 1e-37


def assert_map_close(got, want, msg=""):
    want32 = want.astype(np.float32)
    np.testing.assert_allclose(got, want32, rtol=MAP_RTOL, atol=MAP_ATOL, err_msg=msg)


# lanes 0: thread-per-env step kernel + tiled map kernel (the default); lanes 8: both fused in one kernel (8 lanes per
# env); lanes 16: lane-per-agent step kernel followed by the generic per-cell map kernel
@pytest.mark.parametrize("lanes", [0, 8, 16])
@pytest.mark.parametrize("name", gu.FLIGHT_FIXTURES)
def test_flight_matches_reference_golden(name, lanes):
    import coopsearch_b200 as cs
    touched = True
    g = gu.load(name)
    kw, base, seed = gu.flight_spec_kwargs(g, "probmap")
    T, E = g["reward"].shape
    n, M = kw["n_agents"], kw["map_size"]
    env = cs.VecFlightEnv(make_args(kw), None, num_envs=E, seed=seed, env_id_base=base, count_touched=touched, reset=False,
                          lanes_per_env=lanes)
    assert env.lanes_per_env == {0: 4, 8: 8, 16: 16}[lanes]
    env.reset(init=True, targets=g["tgt_xy"])
    assert_map_close(cpu(env.prob_map), g["init_map"], "init map")
    assert np.array_equal(cpu(env.found_mask).astype(np.uint32), g["init_found"])
    for t in range(T):
        r, term, win = env.step(g["actions"][t])
        where = "%s step %d" % (name, t)
        assert np.array_equal(cpu(env.found_mask).astype(np.uint32), g["found"][t]), where
        assert np.array_equal(cpu(r), g["reward"][t].astype(np.float32)), where
        assert np.array_equal(cpu(term), g["terminated"][t]), where
        assert np.array_equal(cpu(win), g["win"][t]), where
        np.testing.assert_allclose(cpu(env.agent_xy), g["xy"][t], rtol=0, atol=1e-9, err_msg=where)
        if (t + 1) in g["map_steps"]:
            k = list(g["map_steps"]).index(t + 1)
            assert_map_close(cpu(env.prob_map), g["maps"][k], where)
    assert_map_close(cpu(env.prob_map), g["final_map"], "final map")
    # reference-shaped observation: map replicated per agent || 4 features (flight_env.py:223-230)
    full = cpu(env.get_obs())
    assert full.shape == (E, n, M * M + 4)
    for a in range(n):
        assert np.array_equal(full[:, a, :M * M], cpu(env.prob_map).reshape(E, -1))
    assert np.array_equal(full[:, :, M * M:], cpu(env.get_obs(full=False)))
    # second episode: reset() without init keeps the map (flight_env.py:84-86)
    env.reset(init=False, targets=g["ep2_tgt_xy"])
    for t in range(g["ep2_actions"].shape[0]):
        r, _, _ = env.step(g["ep2_actions"][t])
        assert np.array_equal(cpu(env.found_mask).astype(np.uint32), g["ep2_found"][t])
        assert np.array_equal(cpu(r), g["ep2_reward"][t].astype(np.float32))
    assert_map_close(cpu(env.prob_map), g["ep2_map"], "second-episode map")


@pytest.mark.parametrize("lanes", [0, 8, 16])
@pytest.mark.parametrize("n_agents,agent_mode,map_size,view_range", [(3, 0, 50, 7), (5, 1, 30, 5), (2, 3, 64, 9), (4, 2, 17, 3),
                                                                     (5, 0, 16, 7), (3, 2, 62, 6), (8, 1, 63, 7), (1, 0, 5, 2)])
def test_flight_matches_c_oracle(n_agents, agent_mode, map_size, view_range, lanes):
    """Fresh seeded inputs, device-drawn targets, two episodes with auto-reset, touched-cell count."""
    import coopsearch_b200 as cs
    touched = True
    E, T, seed, base = 48, 130, 21, 9000
    spec = FlightSpec(n_agents=n_agents, agent_mode=agent_mode, map_size=map_size, view_range=view_range,
                      time_limit=100, variant="probmap")
    env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base,
                          auto_reset=True, count_touched=touched, lanes_per_env=lanes)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E, auto_reset=True)
    orc.reset(init=True)
    actions = np.random.default_rng(3).integers(0, 3, size=(T, E, n_agents), dtype=np.uint8)
    assert_map_close(cpu(env.prob_map), orc.map, "after reset")
    for t in range(T):
        r, term, win = env.step(actions[t])
        orr, ot, ow = orc.step(actions[t])
        where = "step %d" % t
        assert np.array_equal(cpu(env.found_mask).astype(np.uint32), orc.found), where
        assert np.array_equal(cpu(r), orr.astype(np.float32)), where
        assert np.array_equal(cpu(term), ot) and np.array_equal(cpu(win), ow), where
        assert np.array_equal(cpu(env.time_step), orc.time_step.astype(np.int32)), where
        if t % 10 == 9 or t == T - 1:
            assert_map_close(cpu(env.prob_map), orc.map, where)
            np.testing.assert_allclose(cpu(env.agent_xy), orc.xy, rtol=0, atol=1e-9, err_msg=where)
    if touched:
        assert env.stats()["map_cells_touched"] == float(orc.touched[0])


@pytest.mark.parametrize("n_agents,map_size,view_range,time_limit", [(3, 50, 7, 25), (5, 24, 7, 12), (8, 16, 3, 9)])
def test_map_kernels_and_reset_paths_agree_bitwise(n_agents, map_size, view_range, time_limit):
    """The tiled map kernel and its fused form (corner-row intervals, tile list, packed float4 sweep), the generic kernel
    (per-cell fp64 corner tests, scalar update), the map kernel on its own stream (map_overlap) and both ways of ending an
    episode -- in-call auto-reset (two sensing calls in one launch) and step + reset(mask=terminated) -- must leave
    bit-identical maps."""
    import coopsearch_b200 as cs
    E, T = 512, 60
    spec = FlightSpec(n_agents=n_agents, map_size=map_size, view_range=view_range, time_limit=time_limit, variant="probmap",
                      target_mode=1)
    args = make_args(dict(spec.__dict__))
    fused_auto = cs.VecFlightEnv(args, None, num_envs=E, seed=4, auto_reset=True, lanes_per_env=8)
    tiled_auto = cs.VecFlightEnv(args, None, num_envs=E, seed=4, auto_reset=True)
    tiled_async = cs.VecFlightEnv(args, None, num_envs=E, seed=4, auto_reset=True, map_overlap=True)
    generic_auto = cs.VecFlightEnv(args, None, num_envs=E, seed=4, auto_reset=True, lanes_per_env=16)
    fused_manual = cs.VecFlightEnv(args, None, num_envs=E, seed=4, auto_reset=False, lanes_per_env=8)
    assert fused_auto.lanes_per_env == 8 and generic_auto.lanes_per_env >= 16 and tiled_auto.lanes_per_env in (1, 4)
    actions = torch.from_numpy(np.random.default_rng(8).integers(0, 3, size=(T, E, n_agents), dtype=np.uint8)).cuda()
    resets = 0
    for t in range(T):
        fused_auto.step(actions[t])
        tiled_auto.step(actions[t])
        tiled_async.step(actions[t])
        generic_auto.step(actions[t])
        _, term, _ = fused_manual.step(actions[t])
        if bool(term.any()):
            resets += int(term.sum())
            fused_manual.reset(mask=term.clone())
        assert torch.equal(fused_auto.prob_map, generic_auto.prob_map), "fused vs generic, step %d" % t
        assert torch.equal(fused_auto.prob_map, tiled_auto.prob_map), "fused vs tiled, step %d" % t
        if t % 7 == 6 or t == T - 1:          # in between, the overlapped map kernels run ahead of / behind the steps
            assert torch.equal(tiled_auto.prob_map, tiled_async.prob_map), "map on its own stream, step %d" % t
        assert torch.equal(fused_auto.prob_map, fused_manual.prob_map), "auto-reset vs manual reset, step %d" % t
        assert torch.equal(fused_auto.found_mask, fused_manual.found_mask)
    assert resets >= E


def test_map_overlap_inside_a_cuda_graph_equals_plain_steps():
    """map_overlap=True captured into a CUDA graph (step kernels on the capturing stream, map kernels forked onto the
    handle's stream, joined by sync_map) against plain eager steps: same maps, same state, through auto-resets."""
    import coopsearch_b200 as cs
    E, K = 700, 9
    spec = FlightSpec(n_agents=3, variant="probmap", time_limit=30)
    args = make_args(dict(spec.__dict__))
    plain = cs.VecFlightEnv(args, gu.TEMPLATE, num_envs=E, seed=6, auto_reset=True)
    over = cs.VecFlightEnv(args, gu.TEMPLATE, num_envs=E, seed=6, auto_reset=True, map_overlap=True)
    acts = torch.from_numpy(np.random.default_rng(1).integers(0, 3, size=(K, E, 3), dtype=np.uint8)).cuda()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        over.step(acts[0]); over.sync_map()                # warm-up outside the capture; joined before it begins
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for k in range(K):
                over.step(acts[k])
            over.sync_map()
        for _ in range(8):
            g.replay()
    torch.cuda.synchronize()
    plain.step(acts[0])
    for _ in range(8):
        for k in range(K):
            plain.step(acts[k])
    assert torch.equal(plain.prob_map, over.prob_map)
    assert torch.equal(plain.get_state(), over.get_state()) and torch.equal(plain._dyn, over._dyn)
    assert plain.stats() == over.stats()
    over.step(acts[1]); plain.step(acts[1])                # eager again after the graph
    assert torch.equal(plain.prob_map, over.prob_map)


def test_state_dict_round_trip_restores_outputs_and_map():
    """get_state_dict / set_state_dict: after a restore, get_obs / get_state / reward / target_find / prob_map are those
    of the checkpointed step, and stepping on equals an uninterrupted run."""
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=3, variant="probmap", time_limit=40)
    mk = lambda ov: cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=200, seed=12, auto_reset=True, map_overlap=ov)
    a, b = mk(False), mk(True)
    actions = torch.from_numpy(np.random.default_rng(5).integers(0, 3, size=(90, 200, 3), dtype=np.uint8)).cuda()
    for t in range(30):
        a.step(actions[t])
    ckpt = a.get_state_dict()
    want = {"obs": a.get_obs(full=False).clone(), "state": a.get_state().clone(), "tf": a.target_find.clone(), "map": a.prob_map}
    for t in range(30, 90):
        a.step(actions[t])
    for t in range(17):                       # b is somewhere else entirely (other episode, other outputs)
        b.step(actions[60 + t])
    b.set_state_dict(ckpt)
    assert torch.equal(b.get_obs(full=False), want["obs"]) and torch.equal(b.get_state(), want["state"])
    assert torch.equal(b.target_find, want["tf"]) and torch.equal(b.prob_map, want["map"])
    for t in range(30, 90):
        b.step(actions[t])
    assert torch.equal(a.get_state(), b.get_state()) and torch.equal(a.prob_map, b.prob_map) and torch.equal(a._dyn, b._dyn)


def test_prob_map_assignment_round_trips_through_the_tiled_layout():
    import coopsearch_b200 as cs
    for M in (50, 17, 5):
        spec = FlightSpec(n_agents=2, map_size=M, view_range=2, variant="probmap", target_mode=1)
        env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), None, num_envs=9, seed=1)
        ref = torch.rand(9, M, M, device="cuda")
        env.prob_map = ref
        assert torch.equal(env.prob_map, ref)
        t = env.prob_map_tiles                                       # [E, tiles, tiles, 4, 4]: cell (i, j) -> [i//4, j//4, i%4, j%4]
        assert float(t[3, 1, 0, 2, 1]) == float(ref[3, 6, 1]) if M > 6 else True


def test_two_live_handles_of_different_size_keep_working():
    """Dynamic shared-memory limits are per kernel: creating a second, smaller handle must not break the first."""
    import coopsearch_b200 as cs
    big = FlightSpec(n_agents=8, map_size=63, view_range=7, variant="probmap", target_mode=1)
    small = FlightSpec(n_agents=2, map_size=16, view_range=3, variant="probmap", target_mode=1)
    for lanes in (0, 8, 32):
        e1 = cs.VecFlightEnv(make_args(dict(big.__dict__)), None, num_envs=64, seed=1, lanes_per_env=lanes)
        e2 = cs.VecFlightEnv(make_args(dict(small.__dict__)), None, num_envs=8, seed=1, lanes_per_env=lanes)
        e1.step_random(3); e2.step_random(3); e1.step_random(3)
        e1.get_obs(); e2.get_obs(); e1.get_obs()
        torch.cuda.synchronize()


def test_partial_reset_keeps_other_maps():
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=3, variant="probmap")
    env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=6, seed=5)
    env.step_random(20)
    before = env.prob_map.clone()
    mask = torch.tensor([0, 1, 0, 0, 1, 0], dtype=torch.uint8, device="cuda")
    env.reset(init=True, mask=mask)
    after = env.prob_map
    for e in (0, 2, 3, 5):
        assert torch.equal(after[e], before[e])
    for e in (1, 4):
        assert not torch.equal(after[e], before[e])
        assert float(after[e].max()) <= 1.0 and float((after[e] == 0.5).float().mean()) > 0.5


@pytest.mark.parametrize("n_agents,map_size,E", [(3, 50, 1000), (5, 20, 777), (1, 62, 64), (2, 51, 33)])
def test_map_observation_tma_equals_plain_copy(n_agents, map_size, E):
    """get_obs() of the flight variant (flight_env.py:223-230): the de-tiling kernel with TMA bulk stores and with plain
    stores -- both equal prob_map.ravel() || features, row by row (odd map sizes always take plain stores)."""
    import coopsearch_b200 as cs
    spec = FlightSpec(n_agents=n_agents, map_size=map_size, view_range=max(2, map_size // 8), variant="probmap", target_mode=1)
    env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), None, num_envs=E, seed=9)
    env.step_random(7)
    env.set_obs_kernel("tma")
    a = env.get_obs().clone()
    env.set_obs_kernel("plain")
    b = env.get_obs().clone()
    assert torch.equal(a, b)
    M2 = map_size * map_size
    assert torch.equal(a[:, :, :M2], env.prob_map.reshape(E, 1, M2).expand(E, n_agents, M2))
    assert torch.equal(a[:, :, M2:], env.get_obs(full=False))
