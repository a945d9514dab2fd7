// Measured FP32-pipe ceilings of the device the library runs on (bench.py's roofline denominator for the
// compositing kernels: MEASURED_PEAKS.json carries HBM and bf16 tensor figures only).  Each probe is a
// register-only kernel -- persistent grid of 148 x 8 CTAs of 256 threads, 8 independent dependency chains per
// thread, no memory traffic until the final store -- timed by the caller with CUDA events on the launch stream.
//   kind 0: FFMA      (fma.rn.f32)        2 flop / lane-instruction
//   kind 1: FFMA2     (fma.rn.f32x2)      4 flop / lane-instruction   (sm_100 packed fp32)
//   kind 2: FADD      (add.rn.f32)        1 flop
//   kind 3: FADD2     (add.rn.f32x2)      2 flop
//   kind 4: SHFL.BFLY (shfl.sync.bfly)    0 flop -- warp-shuffle issue rate, what the backward's reduce-scatter pays
//   kind 5: MUFU.EX2  (ex2.approx.ftz)    1 "flop" -- the compositor's exponential
#include "common.cuh"

namespace {

constexpr int PROBE_CHAINS = 8;
constexpr int PROBE_UNROLL = 32;   // instructions per chain per loop iteration

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}

template <int KIND>
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float seed, float* __restrict__ out) {
    float a = seed * 1e-3f + 1.0f, b = 1.0f - seed * 1e-6f;
    float acc = 0.f;
    if (KIND == 0 || KIND == 2 || KIND == 5) {
        float v[PROBE_CHAINS];
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) v[c] = (float)(threadIdx.x + c) * 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < PROBE_UNROLL; ++u) {
#pragma unroll
                for (int c = 0; c < PROBE_CHAINS; ++c) {
                    if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(b), "f"(a));
                    if (KIND == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a));
                    if (KIND == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) acc += v[c];
    } else if (KIND == 1 || KIND == 3) {
        unsigned long long v[PROBE_CHAINS];
        const unsigned long long a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) v[c] = pack2((float)(threadIdx.x + c) * 1e-3f, (float)c);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < PROBE_UNROLL; ++u) {
#pragma unroll
                for (int c = 0; c < PROBE_CHAINS; ++c) {
                    if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(b2), "l"(a2));
                    if (KIND == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(a2));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) acc += __uint_as_float((unsigned)(v[c] & 0xffffffffull)) + __uint_as_float((unsigned)(v[c] >> 32));
    } else {
        float v[PROBE_CHAINS];
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) v[c] = (float)(threadIdx.x + c);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < PROBE_UNROLL; ++u) {
#pragma unroll
                for (int c = 0; c < PROBE_CHAINS; ++c)
                    asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+f"(v[c]));
            }
        }
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) acc += v[c];
    }
    if (acc == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;   // never true in practice: keeps the chains live
}

}  // namespace

// lane-level instructions one launch executes (caller divides by the event-timed duration); 0 for an unknown kind
extern "C" int64_t emd_fp32_probe_lane_instructions(int kind, int iters) {
    if (kind < 0 || kind > 5) return 0;
    return (int64_t)EMD_NUM_SMS * 8 * 256 * PROBE_CHAINS * PROBE_UNROLL * (int64_t)iters;
}

// out: >= 148 * 8 * 256 floats (never written in practice)
extern "C" int emd_fp32_probe(int kind, int iters, float* out, cudaStream_t stream) {
    EMD_CHECK_ARG(kind >= 0 && kind <= 5 && iters >= 1, "fp32_probe: kind 0..5, iters >= 1");
    const dim3 grid(EMD_NUM_SMS * 8), block(256);
    switch (kind) {
        case 0: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<0><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
        case 1: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<1><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
        case 2: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<2><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
        case 3: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<3><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
        case 4: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<4><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
        default: EMD_LAUNCH(EK_MISC, stream, fp32_probe_kernel<5><<<grid, block, 0, stream>>>(iters, 1.0f, out)); break;
    }
    EMD_CHECK_LAUNCH("fp32_probe");
    return EMD_OK;
}
