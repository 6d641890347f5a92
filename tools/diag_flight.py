"""GPU-vs-C-oracle diagnosis: prints the first state divergence per configuration in full detail."""
import sys, os, types
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
import coopsearch_b200 as cs
from oracle import c_oracle
from oracle.py_envs import FlightSpec


def args_of(spec):
    return types.SimpleNamespace(env="x", map_size=spec.map_size, target_num=spec.target_num, target_mode=spec.target_mode,
                                 agent_mode=spec.agent_mode, n_agents=spec.n_agents, view_range=spec.view_range,
                                 time_limit=spec.time_limit, detect_prob=spec.detect_prob, safe_dist=spec.safe_dist,
                                 agent_velocity=spec.velocity, force_dist=spec.force_dist)


def run(n, am, tm, lpe, E=96, T=200, seed=7, base=5000, variant="easy"):
    spec = FlightSpec(n_agents=n, agent_mode=am, target_mode=tm, variant=variant)
    cls = cs.VecFlightEasyEnv if variant == "easy" else cs.VecFlightEnv
    env = cls(args_of(spec), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base, lanes_per_env=lpe)
    orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E)
    orc.reset(init=True)
    actions = np.random.default_rng(99).integers(0, 3, size=(T, E, n), dtype=np.uint8)
    print("config n=%d am=%d tm=%d lpe=%d(%d) tgt maxdiff %.3g" % (n, am, tm, lpe, env.lanes_per_env,
          np.abs(env.tgt_xy.cpu().numpy() - orc.tgt).max()))
    for t in range(T):
        pre_xy = env.agent_xy.cpu().numpy().copy(); pre_yaw = env.agent_yaw.cpu().numpy().copy()
        o_pre_xy = orc.xy.copy(); o_pre_yaw = orc.yaw.copy()
        env.step(actions[t]); orc.step(actions[t])
        meta = env.meta.cpu().numpy().astype(np.uint32)
        bad = np.nonzero((meta[:, 2] != orc.out) | (meta[:, 0] != orc.found))[0]
        if len(bad):
            e = bad[0]
            print("  DIVERGE step %d env %d: gpu out %s found %s | oracle out %s found %s" % (t, e, bin(meta[e, 2]), bin(meta[e, 0]), bin(orc.out[e]), bin(orc.found[e])))
            print("  actions", actions[t, e])
            for a in range(n):
                print("   agent %d pre gpu xy=(%r,%r) yaw=%r | pre orc xy=(%r,%r) yaw=%r" % (a, pre_xy[e, a, 0], pre_xy[e, a, 1], pre_yaw[e, a], o_pre_xy[e, a, 0], o_pre_xy[e, a, 1], o_pre_yaw[e, a]))
                g = env.agent_xy[e, a].cpu().numpy(); gy = env.agent_yaw[e, a].item()
                print("           post gpu xy=(%r,%r) yaw=%r | post orc xy=(%r,%r) yaw=%r" % (g[0], g[1], gy, orc.xy[e, a, 0], orc.xy[e, a, 1], orc.yaw[e, a]))
            return False
    dxy = np.abs(env.agent_xy.cpu().numpy() - orc.xy).max()
    print("  OK all %d steps; final max|dxy| %.3g" % (T, dxy))
    return True


if __name__ == "__main__":
    for cfg in [(5, 2, 0, 8), (5, 2, 0, 1), (5, 2, 0, 32), (5, 3, 1, 1), (1, 1, 0, 2), (8, 0, 1, 16), (3, 0, 0, 0)]:
        run(*cfg)
