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

// dst row dst_idx[i] <- src row src_idx[i] (a null index array = identity), rows of row_bytes bytes: the row gather /
// scatter of the device replay ring (common/replay_buffer.py:36-79).  16-byte words where the rows allow it.
template <typename W>
__global__ void __launch_bounds__(256) cs_rows_copy_kernel(W* __restrict__ dst, const W* __restrict__ src, uint64_t row_words,
                                                           const long long* __restrict__ dst_idx, const long long* __restrict__ src_idx,
                                                           long long count) {
    const uint64_t total = (uint64_t)count * row_words;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_words, k = i - r * row_words;
        const uint64_t d = dst_idx ? (uint64_t)dst_idx[r] : r, s = src_idx ? (uint64_t)src_idx[r] : r;
        dst[d * row_words + k] = src[s * row_words + k];
    }
}
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

int cs_rows_copy(void* d_dst, const void* d_src, uint64_t row_bytes, const int64_t* d_dst_idx, const int64_t* d_src_idx, int64_t count,
                 void* stream) {
    CS_REQUIRE(d_dst && d_src && row_bytes > 0 && count >= 0, "cs_rows_copy: bad argument");
    if (count == 0) return CS_OK;
    const bool wide = row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(d_dst) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_src) % 16 == 0);
    const uint64_t words = wide ? row_bytes / 16 : row_bytes;
    const uint64_t total = (uint64_t)count * words;
    const int grid = (int)((total + 255) / 256 < (uint64_t)CS_NUM_SMS_B200 * 16 ? (total + 255) / 256 : (uint64_t)CS_NUM_SMS_B200 * 16);
    if (wide)
        cs_rows_copy_kernel<uint4><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uint4*>(d_dst), static_cast<const uint4*>(d_src), words,
                                                                          reinterpret_cast<const long long*>(d_dst_idx), reinterpret_cast<const long long*>(d_src_idx), count);
    else
        cs_rows_copy_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uint8_t*>(d_dst), static_cast<const uint8_t*>(d_src), words,
                                                                            reinterpret_cast<const long long*>(d_dst_idx), reinterpret_cast<const long long*>(d_src_idx), count);
    cs_count_launch(1);
    CS_CUDA(cudaGetLastError());
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
