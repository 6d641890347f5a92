import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
import coopsearch_b200 as cs
import golden_util as gu
from oracle.py_envs import FlightSpec
from test_gpu_flight_easy import make_args
spec = FlightSpec(n_agents=3, variant="probmap")
env = cs.VecFlightEnv(make_args(dict(spec.__dict__)), gu.TEMPLATE, num_envs=4, seed=5)
torch.cuda.synchronize()
print("reset ok", float(env.prob_map.min()), float(env.prob_map.max()))
env.step_random(3)
torch.cuda.synchronize()
print("step ok", float(env.prob_map.min()))
