// K1f -- voxel LBS-weight lookup of the SMPL nodes, forward and VJP.
//
// Replaces VoxelDeformer.forward (OmniRe/models/modules.py:612-625: normalize -> F.grid_sample on the 5-D
// [B, 24, D, H, W] weight volume -> squeeze / permute) and the full-volume `lbs_voxel_base + voxel_w_correction`
// add of get_voxel_weight (:575-582) that the reference performs every step; call site
// SMPLTemplate.forward (OmniRe/models/human_body.py:174-179), config `use_voxel_deformer: true`.
//
// Layout: the reference stores the volume channel-major, so the 24 bone weights of one voxel corner sit in 24
// different sectors.  Here base and correction live channel-LAST, [B][D][H][W][J]: a corner is 96 contiguous
// bytes (J = 24).  Eight lanes own one point (J/4 of them carry a float4 of channels), a warp owns four
// points; a lane issues its 16 independent 16-byte loads (8 corners x {base, correction}) before the first use,
// so the base + correction sum is formed only at the ~55 k x 8 corners that are read, not over B x 24 x 64 k voxels.
// HBM/L2-bound gather: 8 corners x 2 x 4J B read + 12 B of coordinates + 4J B written per point.
//
// Backward: the correction gradient is scattered with 16-byte vector reductions into a caller-zeroed buffer of
// the same layout (as the reference's grid_sample backward does); the gradient w.r.t. the canonical point
// (the Gaussian mean) is reduced over the eight lanes by shuffles and written once per point.
#include "common.cuh"
#include "voxel_math.cuh"

constexpr int VOX_THREADS = 256;
constexpr int VOX_LPP = 8;                        // lanes per point
constexpr int VOX_PPB = VOX_THREADS / VOX_LPP;    // points per block
constexpr int VOX_JMAX = VOX_LPP * 4;

struct VoxArgs {
    VoxGeom G;
    const float *base, *corr;       // [B,D,H,W,J]; corr may be NULL
    const float *offset, *scale;    // [B,3], [B]
    const float* xc;                // [B,V,3]
    int B;
    int64_t V;
};

__device__ __forceinline__ float4 vox_f4_fma(float w, float4 g, float4 a) {
    return make_float4(fmaf(w, g.x, a.x), fmaf(w, g.y, a.y), fmaf(w, g.z, a.z), fmaf(w, g.w, a.w));
}

__device__ __forceinline__ void vox_point_tap(const VoxArgs& A, int64_t n, int& b, float& scale, VoxTap& T) {
    b = (int)(n / A.V);
    const float x[3] = {__ldg(A.xc + n * 3 + 0), __ldg(A.xc + n * 3 + 1), __ldg(A.xc + n * 3 + 2)};
    const float off[3] = {__ldg(A.offset + b * 3 + 0), __ldg(A.offset + b * 3 + 1), __ldg(A.offset + b * 3 + 2)};
    scale = __ldg(A.scale + b);
    vox_tap(A.G, x, off, scale, T);
}

__global__ void __launch_bounds__(VOX_THREADS) voxel_lbs_fwd_kernel(const VoxArgs A, float* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * VOX_PPB + (threadIdx.x / VOX_LPP);
    const int q = threadIdx.x % VOX_LPP;
    const int J4 = A.G.J / 4;
    if (n >= (int64_t)A.B * A.V || q >= J4) return;
    int b;
    float scale;
    VoxTap T;
    vox_point_tap(A, n, b, scale, T);
    float4 g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t e = vox_corner_index(A.G, T, b, k) * J4 + q;
        g[k] = __ldg(reinterpret_cast<const float4*>(A.base) + e);
        if (A.corr) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(A.corr) + e);
            g[k] = make_float4(g[k].x + c.x, g[k].y + c.y, g[k].z + c.z, g[k].w + c.w);
        }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = vox_f4_fma(vox_corner_weight(T, k), g[k], acc);
    reinterpret_cast<float4*>(out)[n * J4 + q] = acc;
}

__global__ void __launch_bounds__(VOX_THREADS) voxel_lbs_bwd_kernel(const VoxArgs A, const float* __restrict__ v_out,
                                                                    float* v_corr, float* __restrict__ v_xc) {
    const int64_t n = (int64_t)blockIdx.x * VOX_PPB + (threadIdx.x / VOX_LPP);
    const int q = threadIdx.x % VOX_LPP;
    const int J4 = A.G.J / 4;
    const bool point = n < (int64_t)A.B * A.V;
    float gx[3] = {0.f, 0.f, 0.f};
    float chain[3] = {0.f, 0.f, 0.f};
    if (point) {
        int b;
        float scale;
        VoxTap T;
        vox_point_tap(A, n, b, scale, T);
#pragma unroll
        for (int a = 0; a < 3; ++a) chain[a] = vox_coord_chain(A.G, T, a, scale);
        if (q < J4) {
            const float4 go = __ldg(reinterpret_cast<const float4*>(v_out) + n * J4 + q);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t e = vox_corner_index(A.G, T, b, k) * J4 + q;
                const float w = vox_corner_weight(T, k);
                if (v_corr && w != 0.f)
                    atomicAdd(reinterpret_cast<float4*>(v_corr) + e, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
                if (v_xc) {
                    float4 g = __ldg(reinterpret_cast<const float4*>(A.base) + e);
                    if (A.corr) {
                        const float4 c = __ldg(reinterpret_cast<const float4*>(A.corr) + e);
                        g = make_float4(g.x + c.x, g.y + c.y, g.z + c.z, g.w + c.w);
                    }
                    const float dot = go.x * g.x + go.y * g.y + go.z * g.z + go.w * g.w;
#pragma unroll
                    for (int a = 0; a < 3; ++a) gx[a] = fmaf(vox_corner_dweight(T, a, k), dot, gx[a]);
                }
            }
        }
    }
    if (v_xc == nullptr) return;   // uniform
    // sum the channel lanes of a point (the eight lanes of a group are contiguous in the warp)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        gx[a] += __shfl_xor_sync(0xffffffffu, gx[a], 1);
        gx[a] += __shfl_xor_sync(0xffffffffu, gx[a], 2);
        gx[a] += __shfl_xor_sync(0xffffffffu, gx[a], 4);
    }
    if (point && q == 0) {
        v_xc[n * 3 + 0] = gx[0] * chain[0];
        v_xc[n * 3 + 1] = gx[1] * chain[1];
        v_xc[n * 3 + 2] = gx[2] * chain[2];
    }
}

static int vox_fill(VoxArgs& A, const char* fn, const float* base, const float* corr, const float* offset, const float* scale,
                    float ratio, int ratio_dim, int B, int D, int H, int W, int J, const float* xc, int64_t V) {
    EMD_CHECK_ARG(B >= 0 && V >= 0 && D >= 1 && H >= 1 && W >= 1, "%s: B=%d V=%lld grid %dx%dx%d", fn, B, (long long)V, D, H, W);
    EMD_CHECK_ARG(J >= 4 && J % 4 == 0 && J <= VOX_JMAX, "%s: J=%d must be a multiple of 4 in [4,%d]", fn, J, VOX_JMAX);
    EMD_CHECK_ARG(ratio_dim >= 0 && ratio_dim < 3, "%s: ratio_dim=%d (0 = x, 1 = y, 2 = z)", fn, ratio_dim);
    if ((int64_t)B * V == 0) return EMD_OK;
    EMD_CHECK_ARG(base && offset && scale && xc, "%s: null argument", fn);
    if (!emd_aligned(base, 16) || (corr && !emd_aligned(corr, 16))) {
        emd_set_error("%s: base / corr must be 16-byte aligned", fn);
        return EMD_ERR_ALIGN;
    }
    A.G.D = D; A.G.H = H; A.G.W = W; A.G.J = J; A.G.ratio = ratio; A.G.ratio_dim = ratio_dim;
    A.base = base; A.corr = corr; A.offset = offset; A.scale = scale; A.xc = xc; A.B = B; A.V = V;
    return EMD_OK;
}

// out[B,V,J] = VoxelDeformer(xc).  base / corr: DEVICE [B,D,H,W,J] channel-last (corr may be NULL: no correction enabled);
// offset[B,3], scale[B]: DEVICE (the module's buffers); ratio / ratio_dim: `normalize`'s stretch of the short axis.
extern "C" int emd_voxel_lbs_fwd(const float* base, const float* corr, const float* offset, const float* scale, float ratio,
                                 int ratio_dim, int B, int D, int H, int W, int J, const float* xc, int64_t V, float* out,
                                 cudaStream_t stream) {
    VoxArgs A;
    int rc = vox_fill(A, "emd_voxel_lbs_fwd", base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V);
    if (rc != EMD_OK) return rc;
    const int64_t N = (int64_t)B * V;
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(out != nullptr, "emd_voxel_lbs_fwd: null output");
    if (!emd_aligned(out, 16)) {
        emd_set_error("emd_voxel_lbs_fwd: out must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    const unsigned grid = (unsigned)emd_cdiv(N, VOX_PPB);
    EMD_LAUNCH(EK_VOX_FWD, stream, (voxel_lbs_fwd_kernel<<<grid, VOX_THREADS, 0, stream>>>(A, out)));
    EMD_CHECK_LAUNCH("emd_voxel_lbs_fwd");
    return EMD_OK;
}

// VJP of emd_voxel_lbs_fwd.  v_corr (layout of corr) is ADDED into: the caller zero-fills it once per step; may be
// NULL.  v_xc[B,V,3] is written; may be NULL.
extern "C" int emd_voxel_lbs_bwd(const float* base, const float* corr, const float* offset, const float* scale, float ratio,
                                 int ratio_dim, int B, int D, int H, int W, int J, const float* xc, int64_t V,
                                 const float* v_out, float* v_corr, float* v_xc, cudaStream_t stream) {
    VoxArgs A;
    int rc = vox_fill(A, "emd_voxel_lbs_bwd", base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V);
    if (rc != EMD_OK) return rc;
    const int64_t N = (int64_t)B * V;
    if (N == 0 || (v_corr == nullptr && v_xc == nullptr)) return EMD_OK;
    EMD_CHECK_ARG(v_out != nullptr, "emd_voxel_lbs_bwd: null v_out");
    if (!emd_aligned(v_out, 16) || (v_corr && !emd_aligned(v_corr, 16))) {
        emd_set_error("emd_voxel_lbs_bwd: v_out / v_corr must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    const unsigned grid = (unsigned)emd_cdiv(N, VOX_PPB);
    EMD_LAUNCH(EK_VOX_BWD, stream, (voxel_lbs_bwd_kernel<<<grid, VOX_THREADS, 0, stream>>>(A, v_out, v_corr, v_xc)));
    EMD_CHECK_LAUNCH("emd_voxel_lbs_bwd");
    return EMD_OK;
}
