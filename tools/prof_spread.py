"""ncu target: a few launches of the simple_spread step kernel on 1M envs (python tools/prof_spread.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, coopsearch_b200 as cs
import types
env = bench.silence(cs.VecSimpleSpreadEnv, types.SimpleNamespace(env="simple_spread", n_agents=3, target_num=3, map_size=50), num_envs=1048576, seed=42, auto_reset=True)
acts = torch.randint(0, 5, (1048576, 3), device="cuda", dtype=torch.uint8)
for _ in range(20):
    env.step(acts)
torch.cuda.synchronize()
print("done")
