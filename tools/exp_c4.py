"""c4 (flight + probability map) experiments: python tools/exp_c4.py [steps] [envs]
env: CS_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity before the first allocation.
Prints us per env-step launch (CUDA events over plain stream launches, actions resident)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import coopsearch_b200 as cs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = dict(bench.WORKLOADS["c4"])
if len(sys.argv) > 2:
    w["envs"] = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.init()
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
if os.environ.get("CS_L2_FETCH"):
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["CS_L2_FETCH"])))   # cudaLimitMaxL2FetchGranularity
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("L2 fetch granularity rc", rc, "now", v.value)
split = int(os.environ.get("CS_C4_SPLIT", "1"))       # the same envs as `split` independent batches on `split` streams
w["batches"] = split
w["envs"] //= split
envs = bench.silence(bench.make_envs, cs, w, dev, 0)
gen = torch.Generator(device=dev).manual_seed(1)
acts = [torch.randint(0, 3, (w["envs"], w["n"]), generator=gen, device=dev, dtype=torch.uint8) for _ in range(8)]
streams = [torch.cuda.Stream(device=dev) for _ in envs]
torch.cuda.synchronize()


def step(k):
    for e, st in zip(envs, streams):
        with torch.cuda.stream(st):
            e.step(acts[k % 8])


for k in range(30):
    step(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for st in streams:
    st.wait_stream(torch.cuda.current_stream())
for k in range(steps):
    step(k)
for st in streams:
    torch.cuda.current_stream().wait_stream(st)
e1.record()
torch.cuda.synchronize()
us = 1000.0 * e0.elapsed_time(e1) / steps
tot = w["envs"] * split
print("c4 envs", tot, "split", split, "steps", steps, "us_per_step %.2f" % us, "env-steps/s %.3e" % (tot / us * 1e6))
