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

// ---- launch accounting / optional per-kernel CUDA-event timing (api.cu) ----------------------
enum EmdKernelId {
    EK_PROJ_FWD = 0, EK_PROJ_BWD, EK_SCAN, EK_ISECT_EMIT, EK_SORT_HIST, EK_SORT_SCATTER, EK_ISECT_OFFSETS,
    EK_RASTER_PACK, EK_RASTER_FWD, EK_RASTER_BWD, EK_RASTER_GATHER, EK_SH_FWD, EK_SH_BWD, EK_ACT_FWD, EK_ACT_BWD,
    EK_RIGID_FWD, EK_RIGID_BWD, EK_SMPL_FWD, EK_SMPL_BWD, EK_MLP_FWD, EK_MLP_BWD, EK_DG_PREP_FWD, EK_DG_PREP_BWD,
    EK_HEX_FWD, EK_HEX_BWD, EK_ADAM, EK_LOSS_FWD, EK_LOSS_BWD, EK_VOX_FWD, EK_VOX_BWD, EK_DENSE_FWD, EK_DENSE_BWD,
    EK_DEFORM_IN, EK_MISC, EK_COUNT
};
void emd_prof_begin(int id, cudaStream_t stream);
void emd_prof_end(int id, cudaStream_t stream);
// every kernel launch of the library goes through this: counts it, and times it when profiling is on
#define EMD_LAUNCH(ID, STREAM, ...)  \
    do {                             \
        emd_prof_begin(ID, STREAM);  \
        __VA_ARGS__;                 \
        emd_prof_end(ID, STREAM);    \
    } while (0)

constexpr int EMD_TILE = 16;          // raster tile edge (pixels)
constexpr int EMD_NUM_SMS = 148;      // B200

#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
#endif
