// K2: fused 3D->2D projection (forward + VJP), gsplat pinhole flavour.
//
// One thread per Gaussian; the world covariance is built once and reused for
// every camera of the call.  HBM-bound: 40 B read per Gaussian + 32 B written
// per (camera, Gaussian) forward (radii 4, means2d 8, depth 4, conic 12,
// tiles-per-Gaussian 4).  [N,3] streams are staged through shared memory with
// 128-bit loads so every global transaction is a full, aligned sector run.
#include "common.cuh"
#include "proj_math.cuh"

namespace {

constexpr int PROJ_THREADS = 256;

// Cooperative load of rows [base, base+PROJ_THREADS) of a dense [N,3] fp32
// array into smem (float4 path when the block's slice is 16-B aligned).
__device__ __forceinline__ void stage_vec3(const float* __restrict__ src, int64_t base, int64_t N, float* smem) {
    const int64_t first = base * 3;
    const int64_t count = min((int64_t)PROJ_THREADS, N - base) * 3;
    if (count <= 0) return;
    const float* p = src + first;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const int n4 = (int)(count >> 2);
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (int i = threadIdx.x; i < n4; i += PROJ_THREADS) reinterpret_cast<float4*>(smem)[i] = __ldg(p4 + i);
        for (int i = (n4 << 2) + threadIdx.x; i < count; i += PROJ_THREADS) smem[i] = __ldg(p + i);
    } else {
        for (int i = threadIdx.x; i < count; i += PROJ_THREADS) smem[i] = __ldg(p + i);
    }
}

__device__ __forceinline__ void unstage_vec3(float* __restrict__ dst, int64_t base, int64_t N, const float* smem) {
    const int64_t first = base * 3;
    const int64_t count = min((int64_t)PROJ_THREADS, N - base) * 3;
    if (count <= 0) return;
    float* p = dst + first;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const int n4 = (int)(count >> 2);
        for (int i = threadIdx.x; i < n4; i += PROJ_THREADS) reinterpret_cast<float4*>(p)[i] = reinterpret_cast<const float4*>(smem)[i];
        for (int i = (n4 << 2) + threadIdx.x; i < count; i += PROJ_THREADS) p[i] = smem[i];
    } else {
        for (int i = threadIdx.x; i < count; i += PROJ_THREADS) p[i] = smem[i];
    }
}

__global__ void __launch_bounds__(PROJ_THREADS) projection_fwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int64_t N, int C, int width, int height,
    float eps2d, float near_plane, float far_plane, float radius_clip, int tile_w, int tile_h,
    int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths,
    float* __restrict__ conics, float* __restrict__ comps, int32_t* __restrict__ tiles_per_gauss) {
    __shared__ __align__(16) float s_mean[PROJ_THREADS * 3];
    __shared__ __align__(16) float s_scale[PROJ_THREADS * 3];
    __shared__ __align__(16) float s_out[PROJ_THREADS * 3];
    __shared__ CamConst s_cam;

    const int64_t base = (int64_t)blockIdx.x * PROJ_THREADS;
    const int64_t i = base + threadIdx.x;
    stage_vec3(means, base, N, s_mean);
    stage_vec3(scales, base, N, s_scale);
    __syncthreads();

    float R[9], M[9], S[6], qn[4], m[3], s[3];
    const bool live = i < N;
    if (live) {
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + i);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        quat_to_rotmat_c(q, R, qn);
        for (int k = 0; k < 3; ++k) { m[k] = s_mean[threadIdx.x * 3 + k]; s[k] = s_scale[threadIdx.x * 3 + k]; }
        covar_world_c(R, s, M, S);
    }
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        if (threadIdx.x == 0) make_cam_const(viewmats + c * 16, Ks + c * 9, width, height, s_cam);
        __syncthreads();
        ProjFwd o;
        o.radius = 0;
        if (live) project_gaussian_c(m, S, s_cam, width, height, eps2d, near_plane, far_plane, radius_clip, o);
        const bool vis = live && o.radius > 0;
        const int64_t ci = (int64_t)c * N + i;
        int tiles = 0;
        if (vis) {
            int x0, y0, x1, y1;
            tile_rect_c(o.m2x, o.m2y, o.radius, tile_w, tile_h, x0, y0, x1, y1);
            tiles = (x1 - x0) * (y1 - y0);
        }
        if (live) {
            radii[ci] = vis ? o.radius : 0;
            tiles_per_gauss[ci] = tiles;
            depths[ci] = vis ? o.z : 0.0f;
            reinterpret_cast<float2*>(means2d)[ci] = vis ? make_float2(o.m2x, o.m2y) : make_float2(0.f, 0.f);
            if (comps) comps[ci] = vis ? o.comp : 0.0f;
        }
        // conics [C,N,3] through smem for vectorised stores
        s_out[threadIdx.x * 3 + 0] = vis ? o.conic_a : 0.0f;
        s_out[threadIdx.x * 3 + 1] = vis ? o.conic_b : 0.0f;
        s_out[threadIdx.x * 3 + 2] = vis ? o.conic_c : 0.0f;
        __syncthreads();
        unstage_vec3(conics + (int64_t)c * N * 3, base, N, s_out);
    }
}

constexpr int PROJ_BWD_CAMS = 8;   // camera constants kept in shared memory per pass over the cameras

// One thread per Gaussian.  The per-camera constants of up to 8 cameras are built by 8 threads at once (not by thread 0
// camera after camera), and the camera loop holds no barrier: the cotangents of a (camera, Gaussian) pair are read
// straight from global memory (consecutive threads read consecutive 8 / 12-byte items, every line is consumed whole by
// the warp), so the loads of the next camera can be in flight under the arithmetic of the current one.
__global__ void __launch_bounds__(PROJ_THREADS, 3) projection_bwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int64_t N, int C, int width, int height,
    float eps2d, float near_plane, float far_plane, float radius_clip, const int32_t* __restrict__ radii,
    const float* __restrict__ v_means2d, const float* __restrict__ v_depths, const float* __restrict__ v_conics,
    float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales) {
    __shared__ __align__(16) float s_a[PROJ_THREADS * 3];
    __shared__ __align__(16) float s_b[PROJ_THREADS * 3];
    __shared__ CamConst s_cam[PROJ_BWD_CAMS];

    const int64_t base = (int64_t)blockIdx.x * PROJ_THREADS;
    const int64_t i = base + threadIdx.x;
    stage_vec3(means, base, N, s_a);
    stage_vec3(scales, base, N, s_b);
    if (threadIdx.x < min(C, PROJ_BWD_CAMS))
        make_cam_const(viewmats + threadIdx.x * 16, Ks + threadIdx.x * 9, width, height, s_cam[threadIdx.x]);
    __syncthreads();
    const bool live = i < N;
    float R[9], M[9], S[6], qn[4], m[3], s[3];
    float inv_norm = 0.f;
    if (live) {
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + i);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        inv_norm = quat_to_rotmat_c(q, R, qn);
        for (int k = 0; k < 3; ++k) { m[k] = s_a[threadIdx.x * 3 + k]; s[k] = s_b[threadIdx.x * 3 + k]; }
        covar_world_c(R, s, M, S);
    }
    float v_mean[3] = {0.f, 0.f, 0.f};
    float v_S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool any = false;
    for (int c0 = 0; c0 < C; c0 += PROJ_BWD_CAMS) {
        if (c0 > 0) {   // more than 8 cameras: next group of constants
            __syncthreads();
            if (threadIdx.x < min(C - c0, PROJ_BWD_CAMS))
                make_cam_const(viewmats + (c0 + threadIdx.x) * 16, Ks + (c0 + threadIdx.x) * 9, width, height, s_cam[threadIdx.x]);
            __syncthreads();
        }
        if (!live) continue;
        for (int c = c0; c < min(C, c0 + PROJ_BWD_CAMS); ++c) {
            const int64_t ci = (int64_t)c * N + i;
            if (radii[ci] <= 0) continue;
            const float2 vm = __ldg(reinterpret_cast<const float2*>(v_means2d) + ci);
            const float vz = v_depths ? __ldg(v_depths + ci) : 0.0f;
            const float vca = __ldg(v_conics + ci * 3 + 0), vcb = __ldg(v_conics + ci * 3 + 1), vcc = __ldg(v_conics + ci * 3 + 2);
            ProjFwd o;
            project_gaussian_c(m, S, s_cam[c - c0], width, height, eps2d, near_plane, far_plane, radius_clip, o);
            if (o.radius <= 0) continue;  // cannot happen: same arithmetic as forward
            project_gaussian_vjp(o, s_cam[c - c0], vm.x, vm.y, vz, vca, vcb, vcc, v_mean, v_S);
            any = true;
        }
    }
    float v_q[4] = {0.f, 0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f};
    if (live && any) covar_world_vjp(qn, inv_norm, R, M, s, v_S, v_q, v_s);
    __syncthreads();
    for (int k = 0; k < 3; ++k) { s_a[threadIdx.x * 3 + k] = v_mean[k]; s_b[threadIdx.x * 3 + k] = v_s[k]; }
    __syncthreads();
    unstage_vec3(v_means, base, N, s_a);
    unstage_vec3(v_scales, base, N, s_b);
    if (live) reinterpret_cast<float4*>(v_quats)[i] = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
}

}  // namespace

extern "C" int emd_projection_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                                  const float* Ks, int64_t N, int64_t C, int width, int height, float eps2d,
                                  float near_plane, float far_plane, float radius_clip, int tile_w, int tile_h,
                                  int32_t* radii, float* means2d, float* depths, float* conics, float* comps,
                                  int32_t* tiles_per_gauss, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && C >= 1 && C <= 4096, "projection_fwd: bad N=%lld C=%lld", (long long)N, (long long)C);
    EMD_CHECK_ARG(width > 0 && height > 0, "projection_fwd: bad image size %dx%d", width, height);
    EMD_CHECK_ARG(C * N < (int64_t)2147483647, "projection_fwd: C*N must fit int32");
    if (N == 0) return EMD_OK;
    if (!emd_aligned(quats, 16) || !emd_aligned(means2d, 8)) {
        emd_set_error("projection_fwd: quats must be 16-B and means2d 8-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t blocks = emd_cdiv(N, PROJ_THREADS);
    EMD_LAUNCH(EK_PROJ_FWD, stream, projection_fwd_kernel<<<(unsigned)blocks, PROJ_THREADS, 0, stream>>>(
        means, quats, scales, viewmats, Ks, N, (int)C, width, height, eps2d, near_plane, far_plane, radius_clip,
        tile_w, tile_h, radii, means2d, depths, conics, comps, tiles_per_gauss));
    EMD_CHECK_LAUNCH("projection_fwd");
    return EMD_OK;
}

extern "C" int emd_projection_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                                  const float* Ks, int64_t N, int64_t C, int width, int height, float eps2d,
                                  float near_plane, float far_plane, float radius_clip, const int32_t* radii,
                                  const float* v_means2d, const float* v_depths, const float* v_conics,
                                  float* v_means, float* v_quats, float* v_scales, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && C >= 1, "projection_bwd: bad N=%lld C=%lld", (long long)N, (long long)C);
    if (N == 0) return EMD_OK;
    if (!emd_aligned(quats, 16) || !emd_aligned(v_quats, 16) || !emd_aligned(v_means2d, 8)) {
        emd_set_error("projection_bwd: quats/v_quats must be 16-B and v_means2d 8-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t blocks = emd_cdiv(N, PROJ_THREADS);
    EMD_LAUNCH(EK_PROJ_BWD, stream, projection_bwd_kernel<<<(unsigned)blocks, PROJ_THREADS, 0, stream>>>(
        means, quats, scales, viewmats, Ks, N, (int)C, width, height, eps2d, near_plane, far_plane, radius_clip,
        radii, v_means2d, v_depths, v_conics, v_means, v_quats, v_scales));
    EMD_CHECK_LAUNCH("projection_bwd");
    return EMD_OK;
}
