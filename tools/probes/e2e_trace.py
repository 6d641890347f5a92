"""Device-side timeline of one pooled compact host-buffer step (torch profiler / CUPTI): kernels and copies with durations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import bench  # noqa: E402
import coopsearch_b200 as cs  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = dict(bench.WORKLOADS["c2"])
envs = bench.silence(bench.make_envs, cs, w, dev, 0)
hs = cs.HostStepper(envs, [torch.cuda.Stream(device=dev)], graph=False, compact=True)
hs.actions.random_(0, 3)
for _ in range(5):
    hs.step()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        hs.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print("%9.1f us  +%8.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
