"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import coopsearch_b200 as cs
import golden_util as gu

def fargs(n, M=50, R=7, T=25, am=0):
    return types.SimpleNamespace(env="x", map_size=M, target_num=15, target_mode=0, agent_mode=am, n_agents=n, view_range=R,
                                 time_limit=T, detect_prob=0.9, safe_dist=1, agent_velocity=1, force_dist=3)
# (class, agents, map size, lanes_per_env: 0 = thread-per-env kernel (flight variant: fused step + map kernel),
#  16 = lane-per-agent kernel (+ generic map kernel))
for cls, n, M, lpe in ((cs.VecFlightEasyEnv, 3, 50, 0), (cs.VecFlightEasyEnv, 5, 12, 0), (cs.VecFlightEasyEnv, 3, 50, 16),
                       (cs.VecFlightEnv, 3, 50, 0), (cs.VecFlightEnv, 8, 63, 0), (cs.VecFlightEnv, 5, 50, 16),
                       (cs.VecFlightEnv, 2, 51, 0), (cs.VecFlightEnv, 2, 70, 0)):
    if os.environ.get("CS_SAN_K"):
        os.environ["CS_TPE_K"] = os.environ["CS_SAN_K"]
    env = cls(fargs(n, M=M, R=min(7, M // 3)), gu.TEMPLATE, num_envs=37, seed=1, auto_reset=True, count_touched=True, lanes_per_env=lpe)
    env.step_random(40)
    acts = torch.randint(0, 3, (37, n), dtype=torch.uint8, device="cuda")
    env.step(acts)
    env.step_host(acts.cpu().numpy())
    if isinstance(env, cs.VecFlightEnv):
        env.get_obs()
        env.set_obs_kernel("plain"); env.get_obs()
        env.prob_map = env.prob_map
    env.stats()
sargs = types.SimpleNamespace(env="search", map_size=33, target_num=40, target_mode=1, target_dir="", agent_mode=0, n_agents=9, view_range=5)
s = cs.VecSearchEnv(sargs, num_envs=21, seed=2, auto_reset=True)
s.step_random(60)
s.stats()
torch.cuda.synchronize()
print("sanitize_run done, launches:", cs.load_library().cs_launch_count())
