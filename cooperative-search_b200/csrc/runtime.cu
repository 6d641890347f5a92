// Library-wide runtime pieces: version, thread-local error string, launch counter, pinned host memory.
#include <atomic>
#include <stdarg.h>
#include "cs_common.cuh"
#include "cs_philox.cuh"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void cs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void cs_count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" {
int cs_version(void) { return CS_ABI_VERSION; }
const char* cs_last_error(void) { return g_err; }
uint64_t cs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int cs_host_alloc(void** out, uint64_t bytes) {
    CS_REQUIRE(out != nullptr, "cs_host_alloc: null out");
    CS_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return CS_OK;
}
int cs_host_free(void* p) {
    CS_CUDA(cudaFreeHost(p));
    return CS_OK;
}

// test hook: one Philox block evaluated on the device (tests/test_gpu_philox.py)
__global__ void cs_philox_probe_kernel(const uint32_t* in6, uint32_t* out4, int count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t* c = in6 + 6 * k;
    cs_u4 w = cs_philox4x32_10(c[0], c[1], c[2], c[3], c[4], c[5]);
    out4[4 * k + 0] = w.x; out4[4 * k + 1] = w.y; out4[4 * k + 2] = w.z; out4[4 * k + 3] = w.w;
}
int cs_debug_philox(const uint32_t* d_in6, uint32_t* d_out4, int32_t count, void* stream) {
    if (count <= 0) return CS_OK;
    cs_philox_probe_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_in6, d_out4, count);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
}
