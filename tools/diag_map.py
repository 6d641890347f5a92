"""GPU-vs-C-oracle diagnosis of the probability map: first diverging cell, with geometry."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
import coopsearch_b200 as cs
from oracle import c_oracle
from oracle.py_envs import FlightSpec
from test_gpu_flight_easy import make_args

n_agents, agent_mode, map_size, view_range = 4, 2, 17, 3
E, T, seed, base = 48, 130, 21, 9000
spec = FlightSpec(n_agents=n_agents, agent_mode=agent_mode, map_size=map_size, view_range=view_range, time_limit=100, variant="probmap")
env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=E, seed=seed, env_id_base=base, auto_reset=True)
orc = c_oracle.FlightBatch(spec, gu.TEMPLATE, seed, base, E, auto_reset=True)
orc.reset(init=True)
actions = np.random.default_rng(3).integers(0, 3, size=(T, E, n_agents), dtype=np.uint8)
for t in range(T):
    prev_g = env.prob_map.cpu().numpy().copy(); prev_o = orc.map.copy()
    pre_xy = orc.xy.copy()
    env.step(actions[t]); orc.step(actions[t])
    g = env.prob_map.cpu().numpy(); o = orc.map.astype(np.float32)
    bad = ~np.isclose(g, o, rtol=1e-5, atol=1e-37)
    if bad.any():
        e, i, j = [int(v[0]) for v in np.nonzero(bad)]
        print("step", t, "env", e, "cell", (i, j), "gpu", g[e, i, j], "orc", orc.map[e, i, j], "prev gpu", prev_g[e, i, j], "prev orc", prev_o[e, i, j])
        print(" time_step gpu", env.time_step[e].item(), "orc", orc.time_step[e], "found gpu", bin(env.found_mask[e].item()), "orc", bin(orc.found[e]), "newfound", bin(orc.meta[e, 1]))
        print(" agents gpu", env.agent_xy[e].cpu().numpy().tolist())
        print(" agents orc", orc.xy[e].tolist(), "pre", pre_xy[e].tolist())
        print(" targets", orc.tgt[e].tolist())
        for (cx, cy) in ((i, j), (i + 1, j), (i, j + 1), (i + 1, j + 1)):
            for a in range(n_agents):
                ax, ay = orc.xy[e, a]
                print("   corner", (cx, cy), "agent", a, "d2 mul %r pow %r" % ((cx - ax) * (cx - ax) + (cy - ay) * (cy - ay), (cx - ax) ** 2 + (cy - ay) ** 2))
        break
else:
    print("no divergence")
