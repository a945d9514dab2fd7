// Densification statistics of a training step (second half): what BasicTrainer.postprocess_per_train_step
// (OmniRe/models/trainers/base.py:279-297) and VanillaGaussians.after_train (OmniRe/models/gaussians/vanilla.py:163-191)
// keep per Gaussian from the rasterizer's radii and screen-space (abs)gradients -- the running sum of the gradient
// norms, the visibility count and the largest screen radius.  The reference spends ~15 ATen launches and several
// boolean-mask gathers (each a host sync) per Gaussian class and step; here ONE launch covers all classes (the classes
// are slices of the concatenated Gaussian list) and the C cameras of the step, applied in camera order exactly as C
// successive reference steps would apply them.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) densify_stats_kernel(
    const int32_t* __restrict__ radii, const float* __restrict__ grads2d, int64_t N, int C, float sx, float sy,
    float inv_last_size, int first, float* __restrict__ xys_grad_norm, float* __restrict__ vis_counts,
    float* __restrict__ max_2Dsize) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float gn = xys_grad_norm[i], vc = vis_counts[i], ms = max_2Dsize[i];
    for (int c = 0; c < C; ++c) {
        const int r = radii[(int64_t)c * N + i];
        const float2 g = __ldg(reinterpret_cast<const float2*>(grads2d) + (int64_t)c * N + i);
        const float gx = g.x * sx, gy = g.y * sy;
        const float norm = sqrtf(gx * gx + gy * gy);
        if (first && c == 0) {
            // vanilla.py:177-180: the very first call stores the norms of ALL points and starts every count at one
            gn = norm;
            vc = 1.0f;
        } else if (r > 0) {
            gn += norm;
            vc += 1.0f;
        }
        if (r > 0) ms = fmaxf(ms, (float)r * inv_last_size);
    }
    xys_grad_norm[i] = gn;
    vis_counts[i] = vc;
    max_2Dsize[i] = ms;
}

}  // namespace

// radii [C,N] int32, grads2d [C,N,2] (means2d.absgrad or .grad); sx = width / 2 * batch_size, sy = height / 2 * batch_size
// (base.py:285-286); last_size = max(width, height).  first != 0 on the first call after the statistics were reset (they
// must then hold zeros).  State arrays [N] float32, updated in place.
extern "C" int emd_densify_stats(const int32_t* radii, const float* grads2d, int64_t N, int64_t C, float sx, float sy,
                                 float last_size, int first, float* xys_grad_norm, float* vis_counts, float* max_2Dsize,
                                 cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && C >= 1 && C <= 64, "densify_stats: bad sizes");
    EMD_CHECK_ARG(last_size > 0.f, "densify_stats: last_size must be positive");
    if (!emd_aligned(grads2d, 8)) {
        emd_set_error("densify_stats: grads2d must be 8-B aligned");
        return EMD_ERR_ALIGN;
    }
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_MISC, stream, densify_stats_kernel<<<(unsigned)emd_cdiv(N, 256), 256, 0, stream>>>(
        radii, grads2d, N, (int)C, sx, sy, 1.0f / last_size, first, xys_grad_norm, vis_counts, max_2Dsize));
    EMD_CHECK_LAUNCH("densify_stats");
    return EMD_OK;
}
