"""Is a D2H copy slower when the host cores have just read (or written) the destination?  Flat 4.2 MB copy and the
strided 48-of-240-byte copy of 262144 rows, untouched / after the CPU read the buffer / with write-combined memory."""
import ctypes as C
import glob
import os
import time
import numpy as np
import torch

path = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(path[0] if path else "libcudart.so")
E = 262144
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))


def timed(fn, touch):
    ts = []
    for rep in range(8):
        touch()
        t0 = time.perf_counter()
        fn()
        st.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts[1:]) * 1e3, sorted(ts[1:])[len(ts[1:]) // 2] * 1e3


for name, flags in (("default", 0), ("write-combined", 4)):
    src = torch.zeros(E * 48, dtype=torch.uint8, device=dev)
    src_rec = torch.zeros(E * 16, dtype=torch.uint8, device=dev)
    p_rows, rows = host_alloc(E * 240, flags)
    p_rec, rec = host_alloc(E * 16, flags)
    flat = lambda: rt.cudaMemcpyAsync(p_rec, C.c_void_p(src_rec.data_ptr()), C.c_size_t(E * 16), 2, C.c_void_p(st.cuda_stream))
    two_d = lambda: rt.cudaMemcpy2DAsync(p_rows, C.c_size_t(240), C.c_void_p(src.data_ptr()), C.c_size_t(48), C.c_size_t(48), C.c_size_t(E), 2, C.c_void_p(st.cuda_stream))
    nothing = lambda: None
    print("%-15s flat 4.2 MB:    untouched min/med %.3f / %.3f ms" % ((name,) + timed(flat, nothing)), flush=True)
    if flags == 0:
        print("%-15s flat 4.2 MB:    after a CPU read   %.3f / %.3f ms" % ((name,) + timed(flat, lambda: rec.sum())), flush=True)
        print("%-15s flat 4.2 MB:    after a CPU write  %.3f / %.3f ms" % ((name,) + timed(flat, lambda: rec.fill(1))), flush=True)
    print("%-15s strided 12.6 MB: untouched          %.3f / %.3f ms" % ((name,) + timed(two_d, nothing)), flush=True)
    if flags == 0:
        print("%-15s strided 12.6 MB: after a CPU read   %.3f / %.3f ms" % ((name,) + timed(two_d, lambda: rows.sum())), flush=True)
