"""Where a pooled compact host-buffer step spends its time (eager, one stream): every stage timed with events / the clock."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import coopsearch_b200 as cs  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = dict(bench.WORKLOADS["c2"])
envs = bench.silence(bench.make_envs, cs, w, dev, 0)
streams = [torch.cuda.Stream(device=dev)]
for graph in (False, True):
    if graph:
        envs = bench.silence(bench.make_envs, cs, w, dev, 0)
    hs = cs.HostStepper(envs, streams, graph=graph, compact=True)
    hs.actions.random_(0, 3)
    for _ in range(5):
        hs.step()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        hs.step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("pooled compact graph=%s: %.3f ms per step (median of 30), envs %d" % (graph, ts[15] * 1e3, sum(e.num_envs for e in envs)), flush=True)
    if not graph:
        # stage by stage on the stream, with events
        lib, pool = hs.lib, hs._pool.ptr
        st = streams[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        from importlib import import_module
        _lib = import_module("coopsearch_b200")._lib
        for rep in range(3):
            t0 = time.perf_counter()
            with torch.cuda.stream(st):
                ev[0].record()
                lib.cs_flight_host_pool_step(C.c_void_p(pool), None, _lib.CS_HOST_NO_SYNC, C.c_void_p(st.cuda_stream))
                ev[1].record()
            t1 = time.perf_counter()
            st.synchronize()
            t2 = time.perf_counter()
            lib.cs_flight_host_pool_expand(C.c_void_p(pool), C.c_void_p(st.cuda_stream), 0)
            t3 = time.perf_counter()
            print("  enqueue %.3f ms, device %.3f ms (events), wait %.3f ms, expand %.3f ms" % ((t1 - t0) * 1e3, ev[0].elapsed_time(ev[1]), (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
