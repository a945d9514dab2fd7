// K2': diff_gauss (Inria-derived) preprocess, forward + VJP -- the second front end onto the
// shared sort / tile-range / compositing kernels.  Replaces diff_gauss' preprocessCUDA (+ its
// backward) as called from S3Gaussian/gaussian_renderer/__init__.py:145.  One thread per
// Gaussian; HBM-bound (44 B read + 232 B of SH when colours come from SH; 40 B written).
#include "common.cuh"
#include "dg_math.cuh"

namespace {

constexpr int DG_THREADS = 256;

struct DgArgs {
    DgCam cam;
    float campos[3];
    int sh_degree;   // degree to evaluate
    int K;           // SH bases stored per Gaussian (0 => colours are precomputed)
    int64_t N;
};

__device__ __forceinline__ void load_world(const float* __restrict__ means, const float* __restrict__ scales,
                                           const float* __restrict__ rots, int64_t n, float mod, float p[3],
                                           float s_mod[3], float q[4], float R[9], float M[9], float S[6]) {
    for (int k = 0; k < 3; ++k) { p[k] = means[n * 3 + k]; s_mod[k] = c_mul(mod, scales[n * 3 + k]); }
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(rots) + n);
    q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
    dg_quat_to_rotmat_c(q, R);
    covar_world_c(R, s_mod, M, S);
}

__global__ void __launch_bounds__(DG_THREADS) dg_preprocess_fwd_kernel(
    DgArgs a, const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ rots,
    const float* __restrict__ shs, int32_t* __restrict__ radii, float* __restrict__ means2d,
    float* __restrict__ depths, float* __restrict__ conics, int32_t* __restrict__ tiles_touched,
    float* __restrict__ rgb, uint8_t* __restrict__ clamped) {
    const int64_t n = (int64_t)blockIdx.x * DG_THREADS + threadIdx.x;
    if (n >= a.N) return;
    float p[3], s_mod[3], q[4], R[9], M[9], S[6];
    load_world(means, scales, rots, n, a.cam.mod, p, s_mod, q, R, M, S);
    DgFwd o;
    dg_project_c(p, S, a.cam, o);
    const bool vis = o.f.radius > 0;
    radii[n] = vis ? o.f.radius : 0;
    tiles_touched[n] = vis ? (o.x1 - o.x0) * (o.y1 - o.y0) : 0;
    depths[n] = vis ? o.f.z : 0.f;
    reinterpret_cast<float2*>(means2d)[n] = vis ? make_float2(o.f.m2x, o.f.m2y) : make_float2(0.f, 0.f);
    conics[n * 3 + 0] = vis ? o.f.conic_a : 0.f;
    conics[n * 3 + 1] = vis ? o.f.conic_b : 0.f;
    conics[n * 3 + 2] = vis ? o.f.conic_c : 0.f;
    if (a.K > 0) {
        float col[3] = {0.f, 0.f, 0.f};
        uint32_t cl = 0;
        if (vis) {
            float x = p[0] - a.campos[0], y = p[1] - a.campos[1], z = p[2] - a.campos[2];
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inv; y *= inv; z *= inv;
            float b[16];
            dg_sh_bases(a.sh_degree, x, y, z, b);
            const int nb = (a.sh_degree + 1) * (a.sh_degree + 1);
            const float* c = shs + n * a.K * 3;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k < nb) { col[0] += b[k] * c[k * 3]; col[1] += b[k] * c[k * 3 + 1]; col[2] += b[k] * c[k * 3 + 2]; }
            for (int ch = 0; ch < 3; ++ch) {
                col[ch] += 0.5f;
                if (col[ch] < 0.f) { cl |= 1u << ch; col[ch] = 0.f; }
            }
        }
        rgb[n * 3] = col[0]; rgb[n * 3 + 1] = col[1]; rgb[n * 3 + 2] = col[2];
        clamped[n] = (uint8_t)cl;
    }
}

__global__ void __launch_bounds__(DG_THREADS) dg_preprocess_bwd_kernel(
    DgArgs a, const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ rots,
    const float* __restrict__ shs, const int32_t* __restrict__ radii, const uint8_t* __restrict__ clamped,
    const float* __restrict__ v_means2d, const float* __restrict__ v_depths, const float* __restrict__ v_conics,
    const float* __restrict__ v_rgb, float* __restrict__ v_means, float* __restrict__ v_scales,
    float* __restrict__ v_rots, float* __restrict__ v_shs) {
    const int64_t n = (int64_t)blockIdx.x * DG_THREADS + threadIdx.x;
    if (n >= a.N) return;
    float v_p[3] = {0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f}, v_q[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vis = radii[n] > 0;
    if (vis) {
        float p[3], s_mod[3], q[4], R[9], M[9], S[6];
        load_world(means, scales, rots, n, a.cam.mod, p, s_mod, q, R, M, S);
        DgFwd o;
        dg_project_c(p, S, a.cam, o);
        float v_S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float2 vm = __ldg(reinterpret_cast<const float2*>(v_means2d) + n);
        dg_project_vjp(o, a.cam, p, vm.x, vm.y, v_depths ? v_depths[n] : 0.f, v_conics[n * 3], v_conics[n * 3 + 1],
                       v_conics[n * 3 + 2], v_p, v_S);
        dg_covar_vjp(q, R, M, s_mod, a.cam.mod, v_S, v_q, v_s);
        if (a.K > 0) {
            float x = p[0] - a.campos[0], y = p[1] - a.campos[1], z = p[2] - a.campos[2];
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inv; y *= inv; z *= inv;
            float b[16], bx[16], by[16], bz[16];
            dg_sh_bases(a.sh_degree, x, y, z, b);
            dg_sh_bases_grad(a.sh_degree, x, y, z, bx, by, bz);
            const int nb = (a.sh_degree + 1) * (a.sh_degree + 1);
            const uint32_t cl = clamped[n];
            float g[3];
            for (int ch = 0; ch < 3; ++ch) g[ch] = ((cl >> ch) & 1u) ? 0.f : v_rgb[n * 3 + ch];
            const float* c = shs + n * a.K * 3;
            float* vc = v_shs + n * a.K * 3;
            float dx = 0.f, dy = 0.f, dz = 0.f;
            for (int k = 0; k < a.K; ++k) {
                const float bk = k < nb ? b[k] : 0.f;
                vc[k * 3] = bk * g[0]; vc[k * 3 + 1] = bk * g[1]; vc[k * 3 + 2] = bk * g[2];
                if (k < nb) {
                    const float cg = c[k * 3] * g[0] + c[k * 3 + 1] * g[1] + c[k * 3 + 2] * g[2];
                    dx += bx[k] * cg; dy += by[k] * cg; dz += bz[k] * cg;
                }
            }
            const float d = dx * x + dy * y + dz * z;
            v_p[0] += (dx - d * x) * inv; v_p[1] += (dy - d * y) * inv; v_p[2] += (dz - d * z) * inv;
        }
    } else if (a.K > 0) {
        float* vc = v_shs + n * a.K * 3;
        for (int k = 0; k < a.K * 3; ++k) vc[k] = 0.f;
    }
    for (int k = 0; k < 3; ++k) { v_means[n * 3 + k] = v_p[k]; v_scales[n * 3 + k] = v_s[k]; }
    reinterpret_cast<float4*>(v_rots)[n] = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
}

__global__ void dg_isect_emit_kernel(const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                                     const float* __restrict__ depths, const int64_t* __restrict__ cum_tiles,
                                     int64_t N, int tile_w, int tile_h, int64_t* __restrict__ keys,
                                     int32_t* __restrict__ vals) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int r = radii[n];
    if (r <= 0) return;
    const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + n);
    int x0, y0, x1, y1;
    tile_rect_dg(m.x, m.y, r, tile_w, tile_h, x0, y0, x1, y1);
    const int64_t lo = (int64_t)(uint32_t)__float_as_int(depths[n]);
    int64_t cur = n == 0 ? 0 : cum_tiles[n - 1];
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            keys[cur] = (((int64_t)ty * tile_w + tx) << 32) | lo;
            vals[cur] = (int32_t)n;
            ++cur;
        }
}

int fill(DgArgs& a, const float* viewmatrix_host, const float* projmatrix_host, const float* campos_host, float tanfovx,
         float tanfovy, int W, int H, float scale_modifier, int sh_degree, int K, int64_t N) {
    EMD_CHECK_ARG(W > 0 && H > 0 && tanfovx > 0 && tanfovy > 0, "dg_preprocess: bad camera");
    EMD_CHECK_ARG(K == 0 || (K >= (sh_degree + 1) * (sh_degree + 1) && K <= 16 && sh_degree >= 0 && sh_degree <= 3),
                  "dg_preprocess: bad SH layout K=%d degree=%d", K, sh_degree);
    make_dg_cam(viewmatrix_host, projmatrix_host, tanfovx, tanfovy, W, H, scale_modifier, a.cam);
    for (int k = 0; k < 3; ++k) a.campos[k] = campos_host[k];
    a.sh_degree = sh_degree; a.K = K; a.N = N;
    return EMD_OK;
}

}  // namespace

// viewmatrix / projmatrix / campos are HOST pointers (16, 16, 3 floats): per-call camera scalars exactly like
// GaussianRasterizationSettings carries them (S3Gaussian/gaussian_renderer/__init__.py:49-62).
extern "C" int emd_dg_preprocess_fwd(const float* means3D, const float* scales, const float* rotations,
                                     const float* shs, const float* viewmatrix_host, const float* projmatrix_host,
                                     const float* campos_host, float tanfovx, float tanfovy, int width, int height,
                                     float scale_modifier, int sh_degree, int K, int64_t N, int32_t* radii,
                                     float* means2d, float* depths, float* conics, int32_t* tiles_touched, float* rgb,
                                     uint8_t* clamped, cudaStream_t stream) {
    DgArgs a;
    int rc = fill(a, viewmatrix_host, projmatrix_host, campos_host, tanfovx, tanfovy, width, height, scale_modifier, sh_degree, K, N);
    if (rc != EMD_OK) return rc;
    if (N == 0) return EMD_OK;
    if (!emd_aligned(rotations, 16) || !emd_aligned(means2d, 8)) { emd_set_error("dg_preprocess_fwd: rotations 16-B / means2d 8-B alignment"); return EMD_ERR_ALIGN; }
    EMD_LAUNCH(EK_DG_PREP_FWD, stream, dg_preprocess_fwd_kernel<<<(unsigned)emd_cdiv(N, DG_THREADS), DG_THREADS, 0, stream>>>(
        a, means3D, scales, rotations, shs, radii, means2d, depths, conics, tiles_touched, rgb, clamped));
    EMD_CHECK_LAUNCH("dg_preprocess_fwd");
    return EMD_OK;
}

extern "C" int emd_dg_preprocess_bwd(const float* means3D, const float* scales, const float* rotations,
                                     const float* shs, const float* viewmatrix_host, const float* projmatrix_host,
                                     const float* campos_host, float tanfovx, float tanfovy, int width, int height,
                                     float scale_modifier, int sh_degree, int K, int64_t N, const int32_t* radii,
                                     const uint8_t* clamped, const float* v_means2d, const float* v_depths,
                                     const float* v_conics, const float* v_rgb, float* v_means3D, float* v_scales,
                                     float* v_rotations, float* v_shs, cudaStream_t stream) {
    DgArgs a;
    int rc = fill(a, viewmatrix_host, projmatrix_host, campos_host, tanfovx, tanfovy, width, height, scale_modifier, sh_degree, K, N);
    if (rc != EMD_OK) return rc;
    if (N == 0) return EMD_OK;
    if (!emd_aligned(rotations, 16) || !emd_aligned(v_rotations, 16) || !emd_aligned(v_means2d, 8)) { emd_set_error("dg_preprocess_bwd: alignment"); return EMD_ERR_ALIGN; }
    EMD_LAUNCH(EK_DG_PREP_BWD, stream, dg_preprocess_bwd_kernel<<<(unsigned)emd_cdiv(N, DG_THREADS), DG_THREADS, 0, stream>>>(
        a, means3D, scales, rotations, shs, radii, clamped, v_means2d, v_depths, v_conics, v_rgb, v_means3D, v_scales, v_rotations, v_shs));
    EMD_CHECK_LAUNCH("dg_preprocess_bwd");
    return EMD_OK;
}

// diff_gauss duplicateWithKeys: key = tile << 32 | float_bits(view depth), value = Gaussian index
extern "C" int emd_dg_isect_emit(const float* means2d, const int32_t* radii, const float* depths,
                                 const int64_t* cum_tiles, int64_t N, int tile_w, int tile_h, int64_t* keys,
                                 int32_t* vals, cudaStream_t stream) {
    if (N == 0) return EMD_OK;
    EMD_LAUNCH(EK_ISECT_EMIT, stream, dg_isect_emit_kernel<<<(unsigned)emd_cdiv(N, 256), 256, 0, stream>>>(
        means2d, radii, depths, cum_tiles, N, tile_w, tile_h, keys, vals));
    EMD_CHECK_LAUNCH("dg_isect_emit");
    return EMD_OK;
}
