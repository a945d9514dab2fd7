// Residual application of the S3Gaussian EMD deformation (Deformation.apply_deform sums of
// S3Gaussian/scene/deformation.py:439-481 as deform_network.forward returns them, :484-527) fused with
//   * GaussianModel.get_features' concatenation of the DC and higher-order SH blocks
//     (S3Gaussian/scene/gaussian_model.py: torch.cat((features_dc, features_rest), dim=1)), and
//   * the sums the trainer's deformation regularisers need (S3Gaussian/train.py:240-305: mean |dx|, |do|, |dshs| of the
//     coarse and the fine branch).
// The reference (and this package's first version) spends ~25 ATen launches and ~5 GB of traffic per step at 1 M
// Gaussians on these element-wise steps; here one forward and one backward kernel, HBM-bound:
//   forward   read 624 B, write 208 B per Gaussian;   backward   read 624 B, write 608 B per Gaussian.
#include "common.cuh"

namespace {

constexpr int SA_THREADS = 256;
constexpr int SA_SUMS = 6;   // |dx_c| |dx_f| |do_c| |do_f| |dshs_c| |dshs_f|

// one thread per (Gaussian, SH coefficient triple k in 0..15): 16 threads per Gaussian; thread k == 0 also handles the
// mean and the opacity.  shs_out[n, k, :] = (k == 0 ? dc[n, 0, :] : rest[n, k-1, :]) + dshs_c[n, k, :] + dshs_f[n, k, :]
__global__ void __launch_bounds__(SA_THREADS) s3g_apply_fwd_kernel(
    const float* __restrict__ point, const float* __restrict__ opacity, const float* __restrict__ dc,
    const float* __restrict__ rest, const float* __restrict__ dx_c, const float* __restrict__ dx_f,
    const float* __restrict__ do_c, const float* __restrict__ do_f, const float* __restrict__ dshs_c,
    const float* __restrict__ dshs_f, int64_t N, float* __restrict__ means, float* __restrict__ opac,
    float* __restrict__ shs, float* __restrict__ partial) {
    __shared__ float s_red[SA_THREADS / 32][SA_SUMS];
    const int64_t gid = (int64_t)blockIdx.x * SA_THREADS + threadIdx.x;
    const int64_t n = gid >> 4;
    const int k = (int)(gid & 15);
    float s[SA_SUMS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        const float* base = k == 0 ? dc + n * 3 : rest + (n * 15 + (k - 1)) * 3;
        const int64_t o = (n * 16 + k) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = __ldg(dshs_c + o + c), b = __ldg(dshs_f + o + c);
            shs[o + c] = __ldg(base + c) + a + b;
            s[4] += fabsf(a);
            s[5] += fabsf(b);
        }
        if (k == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = __ldg(dx_c + n * 3 + c), b = __ldg(dx_f + n * 3 + c);
                means[n * 3 + c] = __ldg(point + n * 3 + c) + a + b;
                s[0] += fabsf(a);
                s[1] += fabsf(b);
            }
            const float a = __ldg(do_c + n), b = __ldg(do_f + n);
            opac[n] = __ldg(opacity + n) + a + b;
            s[2] = fabsf(a);
            s[3] = fabsf(b);
        }
    }
    // fixed-order block reduction -> one partial row per block (summed by the caller in double: bit-reproducible)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < SA_SUMS; ++j) {
        float v = s[j];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < SA_SUMS) {
        float v = 0.f;
        for (int w = 0; w < SA_THREADS / 32; ++w) v += s_red[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * SA_SUMS + threadIdx.x] = v;
    }
}

// coef[j] = dL/d(sum_j): the regulariser's weight x 1/numel, scaled by the loss cotangent (a DEVICE array of 6 floats)
__global__ void __launch_bounds__(SA_THREADS) s3g_apply_bwd_kernel(
    const float* __restrict__ v_means, const float* __restrict__ v_opac, const float* __restrict__ v_shs,
    const float* __restrict__ dx_c, const float* __restrict__ dx_f, const float* __restrict__ do_c,
    const float* __restrict__ do_f, const float* __restrict__ dshs_c, const float* __restrict__ dshs_f,
    const float* __restrict__ coef, int64_t N, float* __restrict__ v_dc, float* __restrict__ v_rest,
    float* __restrict__ v_dx_c, float* __restrict__ v_dx_f, float* __restrict__ v_do_c, float* __restrict__ v_do_f,
    float* __restrict__ v_dshs_c, float* __restrict__ v_dshs_f) {
    const int64_t gid = (int64_t)blockIdx.x * SA_THREADS + threadIdx.x;
    const int64_t n = gid >> 4;
    const int k = (int)(gid & 15);
    if (n >= N) return;
    const float c0 = __ldg(coef + 0), c1 = __ldg(coef + 1), c2 = __ldg(coef + 2), c3 = __ldg(coef + 3),
                c4 = __ldg(coef + 4), c5 = __ldg(coef + 5);
    auto sgn = [](float x) { return x > 0.f ? 1.0f : (x < 0.f ? -1.0f : 0.0f); };   // d|x|/dx as torch.abs' backward (0 at 0)
    const int64_t o = (n * 16 + k) * 3;
    float* vb = k == 0 ? v_dc + n * 3 : v_rest + (n * 15 + (k - 1)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float g = v_shs ? __ldg(v_shs + o + c) : 0.f;
        vb[c] = g;
        v_dshs_c[o + c] = g + c4 * sgn(__ldg(dshs_c + o + c));
        v_dshs_f[o + c] = g + c5 * sgn(__ldg(dshs_f + o + c));
    }
    if (k == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float g = v_means ? __ldg(v_means + n * 3 + c) : 0.f;
            v_dx_c[n * 3 + c] = g + c0 * sgn(__ldg(dx_c + n * 3 + c));
            v_dx_f[n * 3 + c] = g + c1 * sgn(__ldg(dx_f + n * 3 + c));
        }
        const float g = v_opac ? __ldg(v_opac + n) : 0.f;
        v_do_c[n] = g + c2 * sgn(__ldg(do_c + n));
        v_do_f[n] = g + c3 * sgn(__ldg(do_f + n));
    }
}

}  // namespace

extern "C" int64_t emd_s3g_apply_blocks(int64_t N) { return emd_cdiv(N * 16, SA_THREADS); }

// point[N,3] opacity[N] dc[N,1,3] rest[N,15,3] dx_*[N,3] do_*[N] dshs_*[N,16,3]  ->  means[N,3] opac[N] shs[N,16,3],
// partial[emd_s3g_apply_blocks(N)][6] = per-block sums of |dx_c| |dx_f| |do_c| |do_f| |dshs_c| |dshs_f|
extern "C" int emd_s3g_apply_fwd(const float* point, const float* opacity, const float* dc, const float* rest,
                                 const float* dx_c, const float* dx_f, const float* do_c, const float* do_f,
                                 const float* dshs_c, const float* dshs_f, int64_t N, float* means, float* opac,
                                 float* shs, float* partial, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0, "s3g_apply_fwd: negative N");
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_MISC, stream, s3g_apply_fwd_kernel<<<(unsigned)emd_s3g_apply_blocks(N), SA_THREADS, 0, stream>>>(
        point, opacity, dc, rest, dx_c, dx_f, do_c, do_f, dshs_c, dshs_f, N, means, opac, shs, partial));
    EMD_CHECK_LAUNCH("s3g_apply_fwd");
    return EMD_OK;
}

// VJP of emd_s3g_apply_fwd.  v_means / v_opac / v_shs may be NULL (zero cotangent); the gradient w.r.t. point is v_means
// and w.r.t. opacity v_opac themselves (not written).  coef: DEVICE float[6], dL/d(sum_j).
extern "C" int emd_s3g_apply_bwd(const float* v_means, const float* v_opac, const float* v_shs, const float* dx_c,
                                 const float* dx_f, const float* do_c, const float* do_f, const float* dshs_c,
                                 const float* dshs_f, const float* coef, int64_t N, float* v_dc, float* v_rest,
                                 float* v_dx_c, float* v_dx_f, float* v_do_c, float* v_do_f, float* v_dshs_c,
                                 float* v_dshs_f, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0, "s3g_apply_bwd: negative N");
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_MISC, stream, s3g_apply_bwd_kernel<<<(unsigned)emd_s3g_apply_blocks(N), SA_THREADS, 0, stream>>>(
        v_means, v_opac, v_shs, dx_c, dx_f, do_c, do_f, dshs_c, dshs_f, coef, N, v_dc, v_rest, v_dx_c, v_dx_f, v_do_c,
        v_do_f, v_dshs_c, v_dshs_f));
    EMD_CHECK_LAUNCH("s3g_apply_bwd");
    return EMD_OK;
}
