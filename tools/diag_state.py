import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
import coopsearch_b200 as cs
from test_gpu_flight_easy import make_args
g = gu.load("easy_1a_am0")
kw, base, seed = gu.flight_spec_kwargs(g, "easy")
T, E = g["reward"].shape
for lpe in (1, 2, 32):
    env = cs.VecFlightEasyEnv(make_args(kw), None, num_envs=E, seed=seed, env_id_base=base, lanes_per_env=lpe, reset=False)
    env.reset(init=True, targets=g["tgt_xy"])
    torch.cuda.synchronize()
    st = env.get_state().cpu().numpy()
    bad = np.argwhere(~np.isclose(st, g["init_state"], rtol=1e-5, atol=1e-6))
    print("lpe", lpe, "bad positions", bad.tolist())
    for e, k in bad[:12]:
        print("   env", e, "idx", k, "got", st[e, k], "want", g["init_state"][e, k])
    print("   found", env.found_mask.cpu().numpy(), g["init_found"])
    full = env._state.cpu().numpy()
