// TMA probe: which inner-coordinate alignments does cp.async.bulk.tensor.3d accept for a (2M, M/2, E) fp32 view?
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probes/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, float* out, int do_store) {
    __shared__ __align__(128) float tile[8 * 16];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512u) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        for (int i = 0; i < 128; ++i) out[i] = tile[i];
        if (do_store) {
            for (int i = 0; i < 128; ++i) tile[i] += 1000.f;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(tile)), "r"(c0), "r"(c1), "r"(c2) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

int main() {
    const int M = 50, E = 4;
    std::vector<float> h((size_t)E * M * M);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 128 * 4);
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {2 * M, M / 2, E};
    const cuuint64_t gstr[2] = {2 * M * 4, M * M * 4};
    const cuuint32_t bdim[3] = {16, 8, 1}, estr[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d (query %d)\n", (int)r, (int)q);
    const int tests[][4] = {{0, 0, 0, 0}, {4, 1, 1, 0}, {2, 1, 1, 0}, {1, 1, 1, 0}, {50, 2, 3, 0}, {52, 20, 3, 1}, {90, 22, 2, 1}, {34, 0, 0, 1}, {6, 3, 1, 1}};
    for (auto& t : tests) {
        cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        probe<<<1, 32>>>(tm, t[0], t[1], t[2], out, t[3]);
        cudaError_t e = cudaDeviceSynchronize();
        float o[128];
        cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
        const float want0 = (float)((size_t)t[2] * M * M + (size_t)t[1] * 2 * M + t[0]);
        printf("coords (%d,%d,%d) store %d: %s  tile[0]=%.0f (want %.0f) tile[16]=%.0f (want %.0f) tile[15]=%.0f\n", t[0], t[1], t[2], t[3],
               cudaGetErrorString(e), o[0], want0, o[16], want0 + 2 * M, o[15]);
        if (e != cudaSuccess) return 1;
        if (t[3]) {
            std::vector<float> back(h.size());
            cudaMemcpy(back.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
            size_t changed = 0, bad = 0;
            for (size_t i = 0; i < h.size(); ++i)
                if (back[i] != h[i]) { ++changed; if (back[i] != h[i] + 1000.f) ++bad; }
            printf("   store: %zu cells changed, %zu unexpected\n", changed, bad);
        }
    }
    return 0;
}
