"""GPU parity of the batched episode writer (generate_episodes) with the reference's own RolloutWorker.generate_episode
(common/rollout.py:22-141): tests/golden/rollout_easy_*.npz were produced by the unmodified RolloutWorker + Agents
(alg=random) on FlightSearchEnvEasy with keyed detection draws (tests/golden/make_golden.py: run_rollout)."""
import types

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu
EP_KEYS = ("o", "s", "u", "r", "avail_u", "o_next", "s_next", "avail_u_next", "u_onehot", "padded", "terminated")


def cpu(t):
    return t.detach().cpu().numpy()


def args_of(g):
    n, m, M, R, T, am = [int(v) for v in g["meta"][:6]]
    v, d, sd, fd = [float(x) for x in g["fmeta"]]
    return types.SimpleNamespace(env="flight_easy", map_size=M, target_num=m, target_mode=0, agent_mode=am, n_agents=n,
                                 view_range=R, time_limit=T, detect_prob=d, safe_dist=sd, agent_velocity=v, force_dist=fd)


@pytest.mark.parametrize("name", ["rollout_easy_3a", "rollout_easy_5a_am3"])
def test_episode_batch_equals_reference_rollout_worker(name):
    import coopsearch_b200 as cs
    g = gu.load(name)
    args = args_of(g)
    env_id, seed = int(g["meta"][7]), int(g["meta"][8])
    episodes = g["u"].shape[0]
    env = cs.VecFlightEasyEnv(args, gu.TEMPLATE, num_envs=1, seed=seed, env_id_base=env_id)     # ctor reset = episode 0
    for k in range(episodes):                                                                   # rollouts = episodes 1, 2, ...
        actions = torch.from_numpy(g["u"][k][:, None, :, 0].copy())                              # [T,1,n]
        ep, rew, win, tf = cs.generate_episodes(env, actions=actions, targets=g["tgt_xy"][k][None])
        where = "%s episode %d" % (name, k)
        for key in EP_KEYS:
            got, want = cpu(ep[key])[0], g[key][k]
            assert got.shape == want.shape, (where, key, got.shape, want.shape)
            if key in ("o", "s", "o_next", "s_next"):
                np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg="%s %s" % (where, key))
            else:
                assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), (where, key)
        assert float(rew[0]) == float(g["episode_reward"][k]), where
        assert int(win[0]) == int(g["win_tag"][k]) and int(tf[0]) == int(g["targets_find"][k]), where
        assert int(ep["length"][0]) == int((g["padded"][k] == 0).sum()), where


def test_episode_batch_many_envs_is_consistent():
    """4096 envs, random policy: every env's arrays obey the layout rules of rollout.py:66-116."""
    import coopsearch_b200 as cs
    args = types.SimpleNamespace(env="flight_easy", map_size=50, target_num=15, target_mode=0, agent_mode=0, n_agents=3,
                                 view_range=7, time_limit=60, detect_prob=0.9, safe_dist=1, agent_velocity=1, force_dist=3)
    env = cs.VecFlightEasyEnv(args, gu.TEMPLATE, num_envs=4096, seed=3)
    ep, rew, win, tf = cs.generate_episodes(env, generator=torch.Generator(device="cuda").manual_seed(5))
    L = cpu(ep["length"])
    T = 60
    assert L.min() >= 1 and L.max() == T
    live = np.arange(T)[None, :] < L[:, None]
    assert np.array_equal(cpu(ep["padded"])[:, :, 0] == 0, live)
    term = cpu(ep["terminated"])[:, :, 0]
    assert np.all(term[~live] == 1) and np.all(term[np.arange(4096), L - 1] == 1)
    assert np.all(term[live & (np.arange(T)[None, :] < (L - 1)[:, None])] == 0)
    r = cpu(ep["r"])[:, :, 0]
    assert np.all(r[~live] == 0) and np.allclose(r.sum(axis=1), cpu(rew))
    o, o_next = cpu(ep["o"]), cpu(ep["o_next"])
    assert np.array_equal(o[:, 1:][live[:, 1:]], o_next[:, :-1][live[:, 1:]])           # o[t+1] = o_next[t] inside an episode
    assert np.all(o[~live] == 0) and np.all(o_next[~live] == 0)
    assert np.array_equal(cpu(ep["u_onehot"]).argmax(-1)[live], cpu(ep["u"])[..., 0][live])
    assert np.all(cpu(ep["avail_u"])[live] == 1) and np.all(cpu(ep["avail_u"])[~live] == 0)
    assert np.array_equal(cpu(tf), cpu(env.target_find))
