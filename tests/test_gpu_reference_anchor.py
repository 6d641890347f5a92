"""The CUDA path against hundreds of trajectories of the UNMODIFIED reference directly (no oracle in between): the compact
fixtures tests/golden/anchor_*.npz hold the reference's own reset layouts, the action sequences, every discrete per-step
output of 384 (3 agents, AM0) and 192 (5 agents, AM2) envs x 200 steps, float64 checkpoints and float32 observation / state
checkpoints.  Discrete outputs must be equal; float64 positions within 1e-9 (the reference squares through libm pow, the
kernels multiply: DESIGN.md section 3) and almost all of them the same bits.  A CPU test pins the C oracle to the same files."""
import os

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import c_oracle
from oracle.py_envs import FlightSpec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
cpu = lambda t: t.detach().cpu().numpy()


def load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, m, M, R, T, am, tm, base, seed = [int(v) for v in g["meta"]]
    vel, d, safe, fd = [float(v) for v in g["fmeta"]]
    as_num = lambda v: int(v) if float(v).is_integer() else v
    spec = FlightSpec(n_agents=n, target_num=m, map_size=M, view_range=R, time_limit=T, agent_mode=am, target_mode=tm,
                      velocity=as_num(vel), detect_prob=d, safe_dist=as_num(safe), force_dist=as_num(fd), variant="easy")
    return g, spec, base, seed


@pytest.mark.parametrize("name", ["anchor_easy_3a_384", "anchor_easy_5a_am2_192"])
def test_c_oracle_equals_the_reference_on_the_anchor_fixtures(name):
    g, spec, base, seed = load(name)
    T, E = g["actions"].shape[:2]
    b = c_oracle.FlightBatch(spec, None, seed, base, E)
    b.reset(targets=g["tgt_xy"], init=True)
    assert np.array_equal(b.found, g["init_found"]) and np.array_equal(b.xy, g["init_xy"])
    for t in range(T):
        r, term, win = b.step(g["actions"][t])
        assert np.array_equal(b.found, g["found"][t]) and np.array_equal(r.astype(np.float32), g["reward"][t]), t
        assert np.array_equal(term, g["terminated"][t]) and np.array_equal(win, g["win"][t]), t
        if t in g["chk_steps"]:
            assert np.array_equal(b.xy, g["chk_xy"][list(g["chk_steps"]).index(t)]), t                 # float64, bit for bit
    assert np.array_equal(b.xy, g["final_xy"]) and np.array_equal(b.yaw, g["final_yaw"])


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [0, 1])
@pytest.mark.parametrize("name", ["anchor_easy_3a_384", "anchor_easy_5a_am2_192"])
def test_cuda_path_equals_the_reference_on_the_anchor_fixtures(name, lanes):
    import coopsearch_b200 as cs
    from test_gpu_flight_easy import make_args
    g, spec, base, seed = load(name)
    T, E = g["actions"].shape[:2]
    env = cs.VecFlightEasyEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base, lanes_per_env=lanes, reset=False)
    env.reset(init=True, targets=g["tgt_xy"])
    assert np.array_equal(cpu(env.found_mask).astype(np.uint32), g["init_found"]) and np.array_equal(cpu(env.agent_xy), g["init_xy"])
    acts = torch.from_numpy(g["actions"]).cuda()
    same_bits, total = 0, 0
    for t in range(T):
        r, term, win = env.step(acts[t])
        meta = cpu(env.meta).astype(np.uint32)
        where = (name, t)
        assert np.array_equal(meta[:, 0], g["found"][t]), where
        assert np.array_equal(meta[:, 3].astype(np.int32), g["time_step"][t]), where
        assert np.array_equal(cpu(r), g["reward"][t]), where
        assert np.array_equal(cpu(term), g["terminated"][t]) and np.array_equal(cpu(win), g["win"][t]), where
        out = ((meta[:, 2][:, None] >> np.arange(spec.n_agents)) & 1).astype(np.uint8)
        live = g["n_steps"] > t
        assert np.array_equal(out[live], g["out"][t][live]), where
        if t in g["chk_steps"]:
            k = list(g["chk_steps"]).index(t)
            np.testing.assert_allclose(cpu(env.agent_xy), g["chk_xy"][k], rtol=0, atol=1e-9, err_msg=str(where))
            np.testing.assert_allclose(cpu(env.get_obs()), g["chk_obs"][k], rtol=1e-6, atol=1e-6, err_msg=str(where))
            np.testing.assert_allclose(cpu(env.get_state()), g["chk_state"][k], rtol=1e-6, atol=1e-6, err_msg=str(where))
            same_bits += int((cpu(env.agent_xy) == g["chk_xy"][k]).sum()); total += g["chk_xy"][k].size
    np.testing.assert_allclose(cpu(env.agent_xy), g["final_xy"], rtol=0, atol=1e-9)
    assert same_bits > 0.99 * total
