// Fused multi-tensor Adam step: the optimizer pass that follows the render hot path.
//
// Replaces torch.optim.Adam.step() over the reference's ~10-40 parameter groups (OmniRe/models/trainers/base.py:226,
// S3Gaussian/scene/gaussian_model.py:200).  One launch updates up to ADAM_MAX_TENSORS tensors: the tensor table
// travels in the kernel parameter space (the library never allocates), every CTA takes 4096-element chunks of the
// concatenated chunk list (grid = a multiple of the 148 SMs, grid-stride), 128-bit loads/stores when all four
// pointers of a tensor are 16-byte aligned.  HBM-bound: 16 B read + 12 B written per parameter.
#include "adam_math.cuh"
#include "common.cuh"

constexpr int ADAM_MAX_TENSORS = 32;
constexpr int ADAM_CHUNK = 4096;
constexpr int ADAM_THREADS = 256;

struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long numel[ADAM_MAX_TENSORS];
    long long chunk_start[ADAM_MAX_TENSORS + 1];   // prefix sum of ceil(numel / ADAM_CHUNK)
    AdamScalars s[ADAM_MAX_TENSORS];
    int n;
};

__global__ void __launch_bounds__(ADAM_THREADS) adam_step_kernel(const __grid_constant__ AdamTable T) {
    const long long total = T.chunk_start[T.n];
    for (long long c = blockIdx.x; c < total; c += gridDim.x) {
        int lo = 0, hi = T.n - 1;               // tensor holding chunk c
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (T.chunk_start[mid] <= c) lo = mid; else hi = mid - 1;
        }
        const long long base = (c - T.chunk_start[lo]) * ADAM_CHUNK;
        const long long left = T.numel[lo] - base;
        const int cnt = left < ADAM_CHUNK ? (int)left : ADAM_CHUNK;
        float* p = T.p[lo] + base;
        const float* g = T.g[lo] + base;
        float* m = T.m[lo] + base;
        float* v = T.v[lo] + base;
        const AdamScalars s = T.s[lo];
        const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
        if (vec) {
            const int n4 = cnt >> 2;
            for (int i = threadIdx.x; i < n4; i += ADAM_THREADS) {
                float4 P = reinterpret_cast<float4*>(p)[i];
                const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
                float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
                adam_update(P.x, G.x, M.x, V.x, s);
                adam_update(P.y, G.y, M.y, V.y, s);
                adam_update(P.z, G.z, M.z, V.z, s);
                adam_update(P.w, G.w, M.w, V.w, s);
                reinterpret_cast<float4*>(p)[i] = P;
                reinterpret_cast<float4*>(m)[i] = M;
                reinterpret_cast<float4*>(v)[i] = V;
            }
            for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += ADAM_THREADS) adam_update(p[i], g[i], m[i], v[i], s);
        } else {
            for (int i = threadIdx.x; i < cnt; i += ADAM_THREADS) adam_update(p[i], g[i], m[i], v[i], s);
        }
    }
}

extern "C" int emd_adam_max_tensors() { return ADAM_MAX_TENSORS; }

// One Adam step over n_tensors <= emd_adam_max_tensors() fp32 tensors.  All arrays are HOST arrays of length n_tensors;
// params / grads / exp_avg / exp_avg_sq hold DEVICE pointers.  step[i] is the 1-based step count AFTER this update
// (torch increments state["step"] before using it).  grad_scale multiplies every gradient first (1/world_size after a
// sum all-reduce).  Tensors with numel 0 are skipped.
extern "C" int emd_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                             const int64_t* numel, const double* lr, const double* beta1, const double* beta2,
                             const double* eps, const double* weight_decay, const int64_t* step, int n_tensors,
                             double grad_scale, cudaStream_t stream) {
    EMD_CHECK_ARG(n_tensors >= 0 && n_tensors <= ADAM_MAX_TENSORS, "emd_adam_step: n_tensors=%d outside [0,%d]", n_tensors,
                  ADAM_MAX_TENSORS);
    if (n_tensors == 0) return EMD_OK;
    EMD_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && lr && beta1 && beta2 && eps && weight_decay && step,
                  "emd_adam_step: null argument");
    AdamTable T;
    T.n = 0;
    long long chunks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        EMD_CHECK_ARG(numel[i] >= 0, "emd_adam_step: numel[%d]=%lld", i, (long long)numel[i]);
        if (numel[i] == 0) continue;
        EMD_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "emd_adam_step: tensor %d has a null pointer", i);
        EMD_CHECK_ARG(step[i] >= 1, "emd_adam_step: step[%d]=%lld must be >= 1", i, (long long)step[i]);
        EMD_CHECK_ARG(beta1[i] >= 0 && beta1[i] < 1 && beta2[i] >= 0 && beta2[i] < 1, "emd_adam_step: betas[%d] outside [0,1)", i);
        const int k = T.n++;
        T.p[k] = params[i]; T.g[k] = grads[i]; T.m[k] = exp_avg[i]; T.v[k] = exp_avg_sq[i];
        T.numel[k] = numel[i];
        T.chunk_start[k] = chunks;
        chunks += emd_cdiv(numel[i], ADAM_CHUNK);
        T.s[k] = adam_scalars(lr[i], beta1[i], beta2[i], eps[i], weight_decay[i], step[i], grad_scale);
    }
    if (T.n == 0) return EMD_OK;
    T.chunk_start[T.n] = chunks;
    const long long want = chunks < (long long)EMD_NUM_SMS * 16 ? chunks : (long long)EMD_NUM_SMS * 16;
    EMD_LAUNCH(EK_ADAM, stream, (adam_step_kernel<<<(unsigned)want, ADAM_THREADS, 0, stream>>>(T)));
    EMD_CHECK_LAUNCH("emd_adam_step");
    return EMD_OK;
}
