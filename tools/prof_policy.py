"""ncu target: a few launches of the batched agent network, tensor-core and fp32 kernels (python tools/prof_policy.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, coopsearch_b200 as cs
print(bench.measure_policy(cs, torch, torch.device("cuda", 0), iters=4))
