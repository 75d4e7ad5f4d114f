// rlca_common.cuh — error plumbing shared by the translation units of librlca.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/rlca.h"

extern thread_local char rlca_g_err[512];

inline int rlca_set_err(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(rlca_g_err, sizeof(rlca_g_err), fmt, a, b);
    return code;
}

#define RLCA_CUDA_TRY(expr)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return rlca_set_err(RLCA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
