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
