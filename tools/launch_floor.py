"""Per-launch floor of back-to-back tiny kernels in a CUDA graph on this GPU (context for the small-config numbers)."""
import torch, time
dev = torch.device("cuda", 0)
xs = [torch.zeros(4096, device=dev) for _ in range(64)]
def step():
    for x in xs:
        x.add_(1.0)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    step(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
torch.cuda.synchronize()
for _ in range(20): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph of 64 tiny kernels: %.2f us per launch" % (1000 * e0.elapsed_time(e1) / (200 * 64)))
e0.record()
for _ in range(200): step()
e1.record(); torch.cuda.synchronize()
print("stream launches of tiny kernels: %.2f us per launch" % (1000 * e0.elapsed_time(e1) / (200 * 64)))
