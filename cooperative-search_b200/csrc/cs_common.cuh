// Shared host/device helpers of libcoopsearch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/coopsearch.h"

#define CS_NUM_SMS_B200 148

void cs_set_error(const char* fmt, ...);
void cs_count_launch(uint64_t n);

#define CS_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            cs_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CS_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define CS_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            cs_set_error(__VA_ARGS__); \
            return CS_ERR_INVALID;     \
        }                              \
    } while (0)

#ifdef __CUDACC__
// ---- sub-warp "group" of LPE lanes working on one env instance -----------------------------
template <int LPE>
struct Group {
    static_assert(LPE == 1 || LPE == 2 || LPE == 4 || LPE == 8 || LPE == 16 || LPE == 32, "LPE");
    __device__ __forceinline__ static unsigned mask() {
        if constexpr (LPE == 32) {
            return 0xffffffffu;
        } else {
            const unsigned lane = threadIdx.x & 31u;
            return ((1u << LPE) - 1u) << (lane & ~(unsigned)(LPE - 1));
        }
    }
    __device__ __forceinline__ static void sync() {
        if (LPE > 1) __syncwarp(mask());
    }
    __device__ __forceinline__ static uint32_t reduce_or(uint32_t v) {
        if (LPE == 1) return v;
        return __reduce_or_sync(mask(), v);
    }
    __device__ __forceinline__ static uint32_t reduce_add(uint32_t v) {
        if (LPE == 1) return v;
        return __reduce_add_sync(mask(), v);
    }
    __device__ __forceinline__ static bool any(bool p) {
        if (LPE == 1) return p;
        return __any_sync(mask(), p) != 0;
    }
};
#endif
