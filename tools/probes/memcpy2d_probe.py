"""How fast does the copy engine scatter small rows into pinned host memory?  cudaMemcpy2DAsync D2H of E rows of `width`
bytes (device pitch = width) into host rows `pitch` bytes apart, against one flat copy of the same payload."""
import ctypes as C
import glob
import os
import time
import torch

path = [p for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))]
rt = C.CDLL(path[0] if path else "libcudart.so")
E = 262144
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
for width, pitch in [(48, 240), (48, 256), (64, 256), (64, 64), (228, 240)]:
    src = torch.zeros(E * width, dtype=torch.uint8, device=dev)
    dst = torch.zeros(E * pitch, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    for flat in (False, True):
        ts = []
        for rep in range(6):
            t0 = time.perf_counter()
            if flat:
                rc = rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), C.c_size_t(E * width), 2, C.c_void_p(st.cuda_stream))
            else:
                rc = rt.cudaMemcpy2DAsync(C.c_void_p(dst.data_ptr()), C.c_size_t(pitch), C.c_void_p(src.data_ptr()), C.c_size_t(width),
                                          C.c_size_t(width), C.c_size_t(E), 2, C.c_void_p(st.cuda_stream))
            assert rc == 0, rc
            st.synchronize()
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        print("width %3d pitch %3d %s: %.3f ms  %.1f GB/s payload" % (width, pitch, "flat" if flat else "2D  ", t * 1e3, E * width / t / 1e9), flush=True)

# 64 copies of 4096 rows (one per env batch), on 1 and 4 streams
E1, B = 4096, 64
srcs = [torch.zeros(E1 * 48, dtype=torch.uint8, device=dev) for _ in range(B)]
dsts = [torch.zeros(E1 * 240, dtype=torch.uint8).pin_memory() for _ in range(B)]
streams = [torch.cuda.Stream() for _ in range(4)]
for ns in (1, 4):
    for flat in (False, True):
        ts = []
        for rep in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for b in range(B):
                s = streams[b % ns]
                if flat:
                    rt.cudaMemcpyAsync(C.c_void_p(dsts[b].data_ptr()), C.c_void_p(srcs[b].data_ptr()), C.c_size_t(E1 * 48), 2, C.c_void_p(s.cuda_stream))
                else:
                    rt.cudaMemcpy2DAsync(C.c_void_p(dsts[b].data_ptr()), C.c_size_t(240), C.c_void_p(srcs[b].data_ptr()), C.c_size_t(48),
                                         C.c_size_t(48), C.c_size_t(E1), 2, C.c_void_p(s.cuda_stream))
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print("64 x 4096 rows, %d stream(s), %s: %.3f ms" % (ns, "flat" if flat else "2D  ", min(ts[1:]) * 1e3), flush=True)
