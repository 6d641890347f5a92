"""Small driver for ncu captures: python tools/profile_run.py <workload> [steps] [lanes_per_env]
Runs plain stream launches (no CUDA graph) of one bench workload so that `ncu -k regex:... -s N -c M` sees them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import coopsearch_b200 as cs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
w = dict(bench.WORKLOADS[name])
w["batches"] = min(w["batches"], int(os.environ.get("CS_PROFILE_BATCHES", "4")))
dev = torch.device("cuda", 0)
envs = bench.silence(bench.make_envs, cs, w, dev, 0)
gen = torch.Generator(device=dev).manual_seed(1)
acts = [torch.randint(0, 3, (w["envs"], w["n"]), generator=gen, device=dev, dtype=torch.uint8) for _ in envs]
grouped = cs.DeviceStepper(envs) if (w.get("grouped") and len(envs) > 1) else None
for k in range(steps):
    if grouped is not None:
        grouped.step(acts)
        continue
    for b, e in enumerate(envs):
        if w["kind"] == "search":
            e.step_random(1)
        else:
            e.step(acts[b])
torch.cuda.synchronize()
print("done", name, steps, "lanes_per_env", getattr(envs[0], "lanes_per_env", None))
