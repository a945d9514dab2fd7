// K1b: spherical-harmonics colour (+ the OmniRe node activations fused around it).
//
//  * emd_sh_fwd/bwd        -- gsplat.cuda._wrapper.spherical_harmonics(deg, dirs, coeffs)
//                             (reference: OmniRe/models/gaussians/basics.py:16; calls at
//                             vanilla.py:388, rigid.py:584, smpl.py:555)
//  * emd_activate_fwd/bwd  -- everything VanillaGaussians/RigidNodes/SMPLNodes.get_gaussians do
//                             after the deformation (vanilla.py:378-414, rigid.py:578-603):
//                             viewdir, SH -> clamp(+0.5, 0, 1), sigmoid(opacity) * frame-valid mask,
//                             exp(scale), normalize(quat) -- one pass, dc/rest read in place (no cat).
//
// HBM-bound: deg 3 reads 192 B of coefficients per Gaussian.  A warp owns 32
// consecutive Gaussians; their coefficient rows are contiguous in memory, so the
// warp streams them with fully coalesced requests into an odd-stride shared
// tile and each lane then reads its own row conflict-free.
#include "common.cuh"

namespace {

constexpr int SH_WARPS = 4;
constexpr int SH_THREADS = SH_WARPS * 32;
constexpr int SH_MAX_ROW = 48;           // 16 bases x 3 channels
constexpr int SH_STRIDE = SH_MAX_ROW + 1;  // odd => conflict-free per-lane rows

__device__ __forceinline__ void sh_bases(int deg, float x, float y, float z, float* b) {
    b[0] = 0.2820947917738781f;
    if (deg < 1) return;
    b[1] = -0.48860251190292f * y;
    b[2] = 0.48860251190292f * z;
    b[3] = -0.48860251190292f * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
    b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// warp streams rows [row0, row0+32) x L floats of src into tile (stride SH_STRIDE)
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ src, int64_t row0, int64_t nrows, int L,
                                               float* tile, int col0, int lane) {
    const int64_t rows = min((int64_t)32, nrows - row0);
    if (rows <= 0) return;
    const float* p = src + row0 * L;
    const int total = (int)rows * L;
    for (int i = lane; i < total; i += 32) {
        const int r = i / L, c = i - r * L;
        tile[r * SH_STRIDE + col0 + c] = __ldg(p + i);
    }
}

__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst, int64_t row0, int64_t nrows, int L,
                                                const float* tile, int col0, int lane) {
    const int64_t rows = min((int64_t)32, nrows - row0);
    if (rows <= 0) return;
    float* p = dst + row0 * L;
    const int total = (int)rows * L;
    for (int i = lane; i < total; i += 32) {
        const int r = i / L, c = i - r * L;
        p[i] = tile[r * SH_STRIDE + col0 + c];
    }
}

// ------------------------------------------------------------------ plain SH
__global__ void __launch_bounds__(SH_THREADS) sh_fwd_kernel(int deg, const float* __restrict__ dirs,
                                                            const float* __restrict__ coeffs, int64_t N, int K,
                                                            float* __restrict__ out) {
    __shared__ float s_tile[SH_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * SH_WARPS + warp) * 32;
    if (row0 >= N) return;
    const int nb = (deg + 1) * (deg + 1);
    float* tile = s_tile[warp];
    warp_load_rows(coeffs, row0, N, K * 3, tile, 0, lane);
    __syncwarp();
    const int64_t n = row0 + lane;
    float col[3] = {0.f, 0.f, 0.f};
    if (n < N) {
        float x = dirs[n * 3 + 0], y = dirs[n * 3 + 1], z = dirs[n * 3 + 2];
        const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
        x *= inv; y *= inv; z *= inv;
        float b[16];
        sh_bases(deg, x, y, z, b);
        const float* row = tile + lane * SH_STRIDE;
        for (int k = 0; k < nb; ++k) {
            col[0] += b[k] * row[k * 3 + 0];
            col[1] += b[k] * row[k * 3 + 1];
            col[2] += b[k] * row[k * 3 + 2];
        }
    }
    __syncwarp();
    if (n < N) { tile[lane * SH_STRIDE + 0] = col[0]; tile[lane * SH_STRIDE + 1] = col[1]; tile[lane * SH_STRIDE + 2] = col[2]; }
    __syncwarp();
    warp_store_rows(out, row0, N, 3, tile, 0, lane);
}

__global__ void __launch_bounds__(SH_THREADS) sh_bwd_kernel(int deg, const float* __restrict__ dirs, int64_t N, int K,
                                                            const float* __restrict__ v_out,
                                                            float* __restrict__ v_coeffs) {
    __shared__ float s_tile[SH_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * SH_WARPS + warp) * 32;
    if (row0 >= N) return;
    const int nb = (deg + 1) * (deg + 1);
    float* tile = s_tile[warp];
    const int64_t n = row0 + lane;
    if (n < N) {
        float x = dirs[n * 3 + 0], y = dirs[n * 3 + 1], z = dirs[n * 3 + 2];
        const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
        x *= inv; y *= inv; z *= inv;
        float b[16];
        sh_bases(deg, x, y, z, b);
        const float v0 = v_out[n * 3 + 0], v1 = v_out[n * 3 + 1], v2 = v_out[n * 3 + 2];
        float* row = tile + lane * SH_STRIDE;
        for (int k = 0; k < K; ++k) {
            const float bk = k < nb ? b[k] : 0.f;
            row[k * 3 + 0] = bk * v0; row[k * 3 + 1] = bk * v1; row[k * 3 + 2] = bk * v2;
        }
    }
    __syncwarp();
    warp_store_rows(v_coeffs, row0, N, K * 3, tile, 0, lane);
}

// ------------------------------------------------------ fused node activation
struct ActArgs {
    const float* means;      // [N,3] world means (viewdir source; no gradient flows through it)
    const float* dc;         // [N,3]
    const float* rest;       // [N,K-1,3]
    const float* opac_logit; // [N]
    const float* log_scales; // [N,3]
    const float* quats;      // [N,4]
    const int64_t* point_ids;  // [N] or null
    const uint8_t* inst_valid; // [I] (frame-valid flag per instance) or null
    float cam[8][3];       // camera centres: colours are produced per camera, coefficients read once
    int C;
    int64_t N;
    int K;          // total bases stored (1 + rest rows)
    int deg;        // degree to use
    int parts;      // bit 0: colours (dc/rest/means -> rgbs), bit 1: geometry (opacity, scales, quats); the host
                    // mirror runs them as two autograd nodes so the SH-coefficient gradient (75 % of the bytes the
                    // all-reduce moves) is ready before the projection backward runs
};
constexpr int ACT_COLORS = 1, ACT_GEOM = 2;

__global__ void __launch_bounds__(SH_THREADS) activate_fwd_kernel(ActArgs a, float* __restrict__ rgbs,
                                                                  float* __restrict__ opac, float* __restrict__ scales,
                                                                  float* __restrict__ quats_n,
                                                                  uint8_t* __restrict__ clamp_pass) {
    __shared__ float s_tile[SH_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * SH_WARPS + warp) * 32;
    if (row0 >= a.N) return;
    const int nb = (a.deg + 1) * (a.deg + 1);
    float* tile = s_tile[warp];
    const bool do_col = a.parts & ACT_COLORS, do_geom = a.parts & ACT_GEOM;
    if (do_col) {
        warp_load_rows(a.dc, row0, a.N, 3, tile, 0, lane);
        if (a.K > 1 && nb > 1) warp_load_rows(a.rest, row0, a.N, (a.K - 1) * 3, tile, 3, lane);
    }
    __syncwarp();
    const int64_t n = row0 + lane;
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (n < a.N && do_col) { mx = a.means[n * 3 + 0]; my = a.means[n * 3 + 1]; mz = a.means[n * 3 + 2]; }
    if (n < a.N && do_geom) {
        float valid = 1.f;
        if (a.inst_valid) valid = a.inst_valid[a.point_ids[n]] ? 1.f : 0.f;
        opac[n] = valid / (1.0f + expf(-a.opac_logit[n]));
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.quats) + n);
        const float qi = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        reinterpret_cast<float4*>(quats_n)[n] = make_float4(q.x * qi, q.y * qi, q.z * qi, q.w * qi);
    }
    // colours: one pass per camera over the staged coefficients (read from HBM once); each camera's
    // colours leave through a small second tile so the stores stay coalesced.
    __shared__ float s_rgb[SH_WARPS][32 * 3];
    for (int c = 0; c < (do_col ? a.C : 0); ++c) {
        float col[3] = {0.f, 0.f, 0.f};
        if (n < a.N) {
            float x = mx - a.cam[c][0], y = my - a.cam[c][1], z = mz - a.cam[c][2];
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inv; y *= inv; z *= inv;
            float b[16];
            sh_bases(a.deg, x, y, z, b);
            const float* row = tile + lane * SH_STRIDE;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < nb) {
                    col[0] += b[k] * row[k * 3 + 0];
                    col[1] += b[k] * row[k * 3 + 1];
                    col[2] += b[k] * row[k * 3 + 2];
                }
            }
            // torch.clamp passes the gradient on the closed interval [0,1]; remember which channels do
            uint32_t pass = 0;
            for (int ch = 0; ch < 3; ++ch) {
                const float pre = col[ch] + 0.5f;
                if (pre >= 0.f && pre <= 1.f) pass |= 1u << ch;
                col[ch] = fminf(fmaxf(pre, 0.f), 1.f);
            }
            clamp_pass[(int64_t)c * a.N + n] = (uint8_t)pass;
        }
        __syncwarp();
        s_rgb[warp][lane * 3 + 0] = col[0]; s_rgb[warp][lane * 3 + 1] = col[1]; s_rgb[warp][lane * 3 + 2] = col[2];
        __syncwarp();
        {
            const int64_t rows = min((int64_t)32, a.N - row0);
            float* dst = rgbs + ((int64_t)c * a.N + row0) * 3;
            for (int i = lane; i < (int)rows * 3; i += 32) dst[i] = s_rgb[warp][i];
        }
    }
    __syncwarp();
    if (!do_geom) return;
    if (n < a.N) {
        float* row = tile + lane * SH_STRIDE;
        row[3] = expf(a.log_scales[n * 3 + 0]); row[4] = expf(a.log_scales[n * 3 + 1]); row[5] = expf(a.log_scales[n * 3 + 2]);
    }
    __syncwarp();
    warp_store_rows(scales, row0, a.N, 3, tile, 3, lane);
}

__global__ void __launch_bounds__(SH_THREADS) activate_bwd_kernel(
    ActArgs a, const uint8_t* __restrict__ clamp_pass, const float* __restrict__ scales,
    const float* __restrict__ v_rgbs, const float* __restrict__ v_opac, const float* __restrict__ v_scales,
    const float* __restrict__ v_quats_n, float* __restrict__ v_dc, float* __restrict__ v_rest,
    float* __restrict__ v_opac_logit, float* __restrict__ v_log_scales, float* __restrict__ v_quats) {
    __shared__ float s_tile[SH_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * SH_WARPS + warp) * 32;
    if (row0 >= a.N) return;
    const int nb = (a.deg + 1) * (a.deg + 1);
    float* tile = s_tile[warp];
    const int64_t n = row0 + lane;
    const bool do_col = a.parts & ACT_COLORS, do_geom = a.parts & ACT_GEOM;
    if (n < a.N && do_col) {
        float* row = tile + lane * SH_STRIDE;
        for (int k = 0; k < a.K * 3; ++k) row[k] = 0.f;
        const float mx = a.means[n * 3 + 0], my = a.means[n * 3 + 1], mz = a.means[n * 3 + 2];
        for (int c = 0; c < a.C; ++c) {
            float x = mx - a.cam[c][0], y = my - a.cam[c][1], z = mz - a.cam[c][2];
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inv; y *= inv; z *= inv;
            float b[16];
            sh_bases(a.deg, x, y, z, b);
            float v[3];
            const uint32_t pass = clamp_pass[(int64_t)c * a.N + n];
            for (int ch = 0; ch < 3; ++ch) v[ch] = ((pass >> ch) & 1u) ? v_rgbs[((int64_t)c * a.N + n) * 3 + ch] : 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < nb) {
                    row[k * 3 + 0] += b[k] * v[0]; row[k * 3 + 1] += b[k] * v[1]; row[k * 3 + 2] += b[k] * v[2];
                }
            }
        }
    }
    if (n < a.N && do_geom) {
        // sigmoid * mask
        float valid = 1.f;
        if (a.inst_valid) valid = a.inst_valid[a.point_ids[n]] ? 1.f : 0.f;
        const float s = 1.0f / (1.0f + expf(-a.opac_logit[n]));
        v_opac_logit[n] = v_opac[n] * valid * s * (1.0f - s);
        // normalize(q)
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.quats) + n);
        const float qi = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float4 g = __ldg(reinterpret_cast<const float4*>(v_quats_n) + n);
        const float qn[4] = {q.x * qi, q.y * qi, q.z * qi, q.w * qi};
        const float d = g.x * qn[0] + g.y * qn[1] + g.z * qn[2] + g.w * qn[3];
        reinterpret_cast<float4*>(v_quats)[n] =
            make_float4((g.x - d * qn[0]) * qi, (g.y - d * qn[1]) * qi, (g.z - d * qn[2]) * qi, (g.w - d * qn[3]) * qi);
    }
    __syncwarp();
    if (do_col) {
        warp_store_rows(v_dc, row0, a.N, 3, tile, 0, lane);
        if (a.K > 1) warp_store_rows(v_rest, row0, a.N, (a.K - 1) * 3, tile, 3, lane);
    }
    if (!do_geom) return;
    __syncwarp();
    if (n < a.N) {
        float* row = tile + lane * SH_STRIDE;
        for (int c = 0; c < 3; ++c) row[c] = v_scales[n * 3 + c] * scales[n * 3 + c];
    }
    __syncwarp();
    warp_store_rows(v_log_scales, row0, a.N, 3, tile, 0, lane);
}

}  // namespace

extern "C" int emd_sh_fwd(int degree, const float* dirs, const float* coeffs, int64_t N, int K, float* out,
                          cudaStream_t stream) {
    EMD_CHECK_ARG(degree >= 0 && degree <= 3, "sh_fwd: degree must be 0..3");
    EMD_CHECK_ARG(K >= (degree + 1) * (degree + 1) && K <= 16, "sh_fwd: K=%d too small for degree %d or > 16", K, degree);
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_SH_FWD, stream, sh_fwd_kernel<<<(unsigned)emd_cdiv(N, SH_WARPS * 32), SH_THREADS, 0, stream>>>(degree, dirs, coeffs, N, K, out));
    EMD_CHECK_LAUNCH("sh_fwd");
    return EMD_OK;
}

extern "C" int emd_sh_bwd(int degree, const float* dirs, int64_t N, int K, const float* v_out, float* v_coeffs,
                          cudaStream_t stream) {
    EMD_CHECK_ARG(degree >= 0 && degree <= 3, "sh_bwd: degree must be 0..3");
    EMD_CHECK_ARG(K >= (degree + 1) * (degree + 1) && K <= 16, "sh_bwd: bad K");
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_SH_BWD, stream, sh_bwd_kernel<<<(unsigned)emd_cdiv(N, SH_WARPS * 32), SH_THREADS, 0, stream>>>(degree, dirs, N, K, v_out, v_coeffs));
    EMD_CHECK_LAUNCH("sh_bwd");
    return EMD_OK;
}

static int make_act_args(ActArgs& a, const float* means, const float* dc, const float* rest, const float* opac_logit,
                         const float* log_scales, const float* quats, const int64_t* point_ids,
                         const uint8_t* inst_valid, const float* cam_pos_host, int C, int64_t N, int K, int degree) {
    EMD_CHECK_ARG(degree >= 0 && degree <= 3, "activate: degree must be 0..3");
    EMD_CHECK_ARG(K >= 1 && K <= 16 && K >= (degree + 1) * (degree + 1), "activate: bad K=%d for degree %d", K, degree);
    EMD_CHECK_ARG((point_ids == nullptr) == (inst_valid == nullptr), "activate: point_ids and inst_valid go together");
    if (!emd_aligned(quats, 16)) { emd_set_error("activate: quats must be 16-B aligned"); return EMD_ERR_ALIGN; }
    a.means = means; a.dc = dc; a.rest = rest; a.opac_logit = opac_logit; a.log_scales = log_scales; a.quats = quats;
    a.point_ids = point_ids; a.inst_valid = inst_valid;
    EMD_CHECK_ARG(C >= 1 && C <= 8, "activate: 1..8 cameras per call (got %d)", C);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < 3; ++k) a.cam[c][k] = cam_pos_host[c * 3 + k];
    a.C = C; a.N = N; a.K = K; a.deg = degree; a.parts = ACT_COLORS | ACT_GEOM;
    return EMD_OK;
}

// cam_pos_host is a HOST pointer to C x 3 floats (per-call scalars, like the image size).  rgbs / clamp_pass are [C,N,.].
extern "C" int emd_activate_fwd(const float* means, const float* dc, const float* rest, const float* opac_logit,
                                const float* log_scales, const float* quats, const int64_t* point_ids,
                                const uint8_t* inst_valid, const float* cam_pos_host, int C, int64_t N, int K, int degree,
                                float* rgbs, float* opac, float* scales, float* quats_n, uint8_t* clamp_pass,
                                cudaStream_t stream) {
    ActArgs a;
    int rc = make_act_args(a, means, dc, rest, opac_logit, log_scales, quats, point_ids, inst_valid, cam_pos_host, C, N, K, degree);
    if (rc != EMD_OK) return rc;
    // a null output group selects the other part alone: rgbs == null -> geometry only; opac == null -> colours only
    a.parts = (rgbs ? ACT_COLORS : 0) | (opac ? ACT_GEOM : 0);
    EMD_CHECK_ARG(a.parts != 0, "activate_fwd: nothing to compute (rgbs and opac both null)");
    EMD_CHECK_ARG(!(a.parts & ACT_COLORS) || clamp_pass, "activate_fwd: colours need clamp_pass");
    EMD_CHECK_ARG(!(a.parts & ACT_GEOM) || (scales && quats_n), "activate_fwd: geometry needs opac, scales and quats_n");
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_ACT_FWD, stream, activate_fwd_kernel<<<(unsigned)emd_cdiv(N, SH_WARPS * 32), SH_THREADS, 0, stream>>>(a, rgbs, opac, scales, quats_n, clamp_pass));
    EMD_CHECK_LAUNCH("activate_fwd");
    return EMD_OK;
}

extern "C" int emd_activate_bwd(const float* means, const float* dc, const float* rest, const float* opac_logit,
                                const float* log_scales, const float* quats, const int64_t* point_ids,
                                const uint8_t* inst_valid, const float* cam_pos_host, int C, int64_t N, int K, int degree,
                                const uint8_t* clamp_pass, const float* scales, const float* v_rgbs,
                                const float* v_opac, const float* v_scales, const float* v_quats_n, float* v_dc,
                                float* v_rest, float* v_opac_logit, float* v_log_scales, float* v_quats,
                                cudaStream_t stream) {
    ActArgs a;
    int rc = make_act_args(a, means, dc, rest, opac_logit, log_scales, quats, point_ids, inst_valid, cam_pos_host, C, N, K, degree);
    if (rc != EMD_OK) return rc;
    // v_rgbs == null -> geometry only; v_opac == null -> colours only (same rule as the forward)
    a.parts = (v_rgbs ? ACT_COLORS : 0) | (v_opac ? ACT_GEOM : 0);
    EMD_CHECK_ARG(a.parts != 0, "activate_bwd: nothing to compute (v_rgbs and v_opac both null)");
    EMD_CHECK_ARG(!(a.parts & ACT_COLORS) || (clamp_pass && v_dc), "activate_bwd: colours need clamp_pass and v_dc");
    EMD_CHECK_ARG(!(a.parts & ACT_GEOM) || (scales && v_scales && v_quats_n && v_opac_logit && v_log_scales && v_quats),
                  "activate_bwd: geometry needs scales, v_scales, v_quats_n and the three outputs");
    if (N == 0) return EMD_OK;
    if ((a.parts & ACT_GEOM) && (!emd_aligned(v_quats, 16) || !emd_aligned(v_quats_n, 16))) { emd_set_error("activate_bwd: quats grads must be 16-B aligned"); return EMD_ERR_ALIGN; }
    EMD_LAUNCH(EK_ACT_BWD, stream, activate_bwd_kernel<<<(unsigned)emd_cdiv(N, SH_WARPS * 32), SH_THREADS, 0, stream>>>(
        a, clamp_pass, scales, v_rgbs, v_opac, v_scales, v_quats_n, v_dc, v_rest, v_opac_logit, v_log_scales, v_quats));
    EMD_CHECK_LAUNCH("activate_bwd");
    return EMD_OK;
}
