// Shared helpers for the emd_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define EMD_OK 0
#define EMD_ERR_BAD_ARG -1
#define EMD_ERR_ALIGN -2
#define EMD_ERR_WORKSPACE -3
#define EMD_ERR_CUDA -4
#define EMD_ERR_UNSUPPORTED -5

// thread-local last error text (no global mutable state shared across threads)
void emd_set_error(const char* fmt, ...);

#define EMD_CHECK_ARG(cond, ...)        \
    do {                                \
        if (!(cond)) {                  \
            emd_set_error(__VA_ARGS__); \
            return EMD_ERR_BAD_ARG;     \
        }                               \
    } while (0)

#define EMD_CHECK_LAUNCH(name)                                                   \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            emd_set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__)); \
            return EMD_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

static inline bool emd_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static inline int64_t emd_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int EMD_TILE = 16;          // raster tile edge (pixels)
constexpr int EMD_NUM_SMS = 148;      // B200

#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
#endif
