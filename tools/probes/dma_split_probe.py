"""Can the strided D2H copy of the agent rows be split over several streams (copy engines)?"""
import ctypes as C
import glob
import os
import time
import torch

path = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(path[0] if path else "libcudart.so")
E = 262144
dev = torch.device("cuda", 0)
print("asyncEngineCount", torch.cuda.get_device_properties(0).async_engine_count if hasattr(torch.cuda.get_device_properties(0), "async_engine_count") else "?")
src = torch.zeros(E * 48, dtype=torch.uint8, device=dev)
dst = torch.zeros(E * 240, dtype=torch.uint8).pin_memory()
for parts in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    rows = E // parts
    ts = []
    for rep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, s in enumerate(streams):
            rt.cudaMemcpy2DAsync(C.c_void_p(dst.data_ptr() + k * rows * 240), C.c_size_t(240), C.c_void_p(src.data_ptr() + k * rows * 48), C.c_size_t(48),
                                 C.c_size_t(48), C.c_size_t(rows), 2, C.c_void_p(s.cuda_stream))
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print("%d stream(s): %.3f ms" % (parts, min(ts[1:]) * 1e3), flush=True)
# the same bytes through a kernel writing mapped pinned memory (zero-copy stores), for comparison
