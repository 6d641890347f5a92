"""c4 timing experiment: python tools/exp_c4.py [lanes_per_env] [map_overlap] [envs]
K steps of the flight variant captured into one CUDA graph, replayed until >= 150 ms; prints us per env-step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import coopsearch_b200 as cs

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
overlap = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
E = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
dev = torch.device("cuda", 0)
env = bench.silence(cs.VecFlightEnv, bench.flight_args("flight", 3, 0), bench.TEMPLATE, num_envs=E, device=dev, seed=42, auto_reset=True,
                    lanes_per_env=lanes, map_overlap=overlap)
gen = torch.Generator(device=dev).manual_seed(1)
K = 48
acts = [torch.randint(0, 3, (E, 3), generator=gen, device=dev, dtype=torch.uint8) for _ in range(8)]
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    for k in range(30):
        env.step(acts[k % 8])
    env.sync_map()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for k in range(K):
            env.step(acts[k % 8])
        env.sync_map()
    torch.cuda.synchronize()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    reps = 80
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        g.replay()
        evs[r + 1].record()
    torch.cuda.synchronize()
ms = sorted(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))
print("lib %s lanes %d (%d) overlap %d E %d: median %.2f us/step  min %.2f  max %.2f" % (
    os.path.basename(os.environ.get("COOPSEARCH_LIB", "default")), lanes, env.lanes_per_env, overlap, E,
    1000 * ms[reps // 2] / K, 1000 * ms[0] / K, 1000 * ms[-1] / K), flush=True)
