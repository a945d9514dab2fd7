// K1e -- HexPlane feature gather feeding the S3Gaussian EMD deformation MLP, forward and VJP.
//
// Replaces HexPlaneField.forward -> interpolate_ms_features -> 24 x F.grid_sample
// (S3Gaussian/scene/hexplane.py:73-106, 165-187; call site deformation.py:187-199).
//
// Layout: the reference stores a plane as [1, F, H, W] (feature-major), so one bilinear tap of F = 32
// features touches 32 different 32-byte sectors.  Here every plane lives feature-LAST, [H][W][F], in one
// flat fp32 buffer: a tap is 128 contiguous bytes.  Eight lanes own one Gaussian (one float4 of features
// each), a warp owns four Gaussians; per scale a lane issues its 24 independent 16-byte loads (6 planes x
// 4 corners) before the first use.  HBM/L2-bound gather: algorithmic traffic per Gaussian is
// 24 planes x 4 corners x 128 B = 12 288 B read + 16 B of coordinates + S*F*4 B written.
//
// Backward: plane gradients are scattered with 16-byte vector reductions (REDG.E.ADD.F32x4) into a
// caller-zeroed buffer of the same layout -- the one place the library uses floating-point atomics, as the
// reference's grid_sample backward does; the coordinate gradients (points, time) are reduced over the eight
// lanes by shuffles and written once per Gaussian; a shared (scalar) time is reduced over the grid in a
// fixed order (per-block partials + one finishing block).
#include "common.cuh"
#include "hexplane_math.cuh"

struct HexParams {
    const float* planes;
    long long off[HEX_MAX_SCALES][HEX_PLANES];   // float offset of plane (s, p) in `planes`
    int reso[HEX_MAX_SCALES][4];                 // grid size per coordinate (x, y, z, t) at scale s
    int S;
    float a0[3], k[3];                           // aabb[0] and 2 / (aabb[1] - aabb[0])
};

constexpr int HEX_THREADS = 256;
constexpr int HEX_LPG = HEX_F / 4;               // lanes per Gaussian
constexpr int HEX_GPB = HEX_THREADS / HEX_LPG;   // Gaussians per block

__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_lerp(float4 a, float4 b, float w) {   // a + w (b - a)
    return make_float4(fmaf(w, b.x - a.x, a.x), fmaf(w, b.y - a.y, a.y), fmaf(w, b.z - a.z, a.z), fmaf(w, b.w - a.w, a.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__device__ __forceinline__ void hex_coords(const HexParams& P, const float* __restrict__ pts, const float* __restrict__ t,
                                           int t_stride, int64_t n, float u[4]) {
    u[0] = hex_normalize(pts[n * 3 + 0], P.a0[0], P.k[0]);
    u[1] = hex_normalize(pts[n * 3 + 1], P.a0[1], P.k[1]);
    u[2] = hex_normalize(pts[n * 3 + 2], P.a0[2], P.k[2]);
    u[3] = t[n * t_stride];
}

__global__ void __launch_bounds__(HEX_THREADS) hexplane_fwd_kernel(const HexParams P, const float* __restrict__ pts,
                                                                   const float* __restrict__ t, int t_stride, int64_t N,
                                                                   float* __restrict__ feat, int64_t ld) {
    const int64_t n = (int64_t)blockIdx.x * HEX_GPB + (threadIdx.x / HEX_LPG);
    const int q = threadIdx.x % HEX_LPG;
    if (n >= N) return;
    float u[4];
    hex_coords(P, pts, t, t_stride, n, u);
    float4* out = reinterpret_cast<float4*>(feat + n * ld) + q;
    for (int s = 0; s < P.S; ++s) {
        HexAxis ax[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ax[c] = hex_axis(u[c], P.reso[s][c]);
        float4 tap[HEX_PLANES][4];
#pragma unroll
        for (int p = 0; p < HEX_PLANES; ++p) {
            const HexAxis X = ax[HEX_AX(p)], Y = ax[HEX_AY(p)];
            const int W = P.reso[s][HEX_AX(p)];
            const float4* base = reinterpret_cast<const float4*>(P.planes + P.off[s][p]) + q;
            const int64_t r0 = (int64_t)Y.i0 * W, r1 = (int64_t)Y.i1 * W;
            tap[p][0] = __ldg(base + (r0 + X.i0) * HEX_LPG);
            tap[p][1] = __ldg(base + (r0 + X.i1) * HEX_LPG);
            tap[p][2] = __ldg(base + (r1 + X.i0) * HEX_LPG);
            tap[p][3] = __ldg(base + (r1 + X.i1) * HEX_LPG);
        }
        float4 prod = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
        for (int p = 0; p < HEX_PLANES; ++p) {
            const float wx = ax[HEX_AX(p)].w1, wy = ax[HEX_AY(p)].w1;
            const float4 top = f4_lerp(tap[p][0], tap[p][1], wx), bot = f4_lerp(tap[p][2], tap[p][3], wx);
            prod = f4_mul(prod, f4_lerp(top, bot, wy));
        }
        out[s * HEX_LPG] = prod;
    }
}

__global__ void __launch_bounds__(HEX_THREADS, 2) hexplane_bwd_kernel(const HexParams P, const float* __restrict__ pts,
                                                                   const float* __restrict__ t, int t_stride, int64_t N,
                                                                   const float* __restrict__ v_feat, int64_t ld, float* v_planes,
                                                                   float* __restrict__ v_pts, float* __restrict__ v_t,
                                                                   float* __restrict__ t_partial) {
    __shared__ float s_t[HEX_THREADS / 32];
    const int64_t n = (int64_t)blockIdx.x * HEX_GPB + (threadIdx.x / HEX_LPG);
    const int q = threadIdx.x % HEX_LPG;
    float g[4] = {0.f, 0.f, 0.f, 0.f};   // d loss / d (ix-space coordinate * dmul) accumulated per coordinate
    if (n < N) {
        float u[4];
        hex_coords(P, pts, t, t_stride, n, u);
        const float4* vf = reinterpret_cast<const float4*>(v_feat + n * ld) + q;
        for (int s = 0; s < P.S; ++s) {
            HexAxis ax[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) ax[c] = hex_axis(u[c], P.reso[s][c]);
            float4 val[HEX_PLANES], dvx[HEX_PLANES], dvy[HEX_PLANES];
#pragma unroll
            for (int p = 0; p < HEX_PLANES; ++p) {
                const HexAxis X = ax[HEX_AX(p)], Y = ax[HEX_AY(p)];
                const int W = P.reso[s][HEX_AX(p)];
                const float4* base = reinterpret_cast<const float4*>(P.planes + P.off[s][p]) + q;
                const int64_t r0 = (int64_t)Y.i0 * W, r1 = (int64_t)Y.i1 * W;
                const float4 nw = __ldg(base + (r0 + X.i0) * HEX_LPG), ne = __ldg(base + (r0 + X.i1) * HEX_LPG);
                const float4 sw = __ldg(base + (r1 + X.i0) * HEX_LPG), se = __ldg(base + (r1 + X.i1) * HEX_LPG);
                const float4 top = f4_lerp(nw, ne, X.w1), bot = f4_lerp(sw, se, X.w1);
                val[p] = f4_lerp(top, bot, Y.w1);
                dvx[p] = f4_lerp(f4_sub(ne, nw), f4_sub(se, sw), Y.w1);   // d val / d ix
                dvy[p] = f4_sub(bot, top);                                // d val / d iy
            }
            const float4 go = vf[s * HEX_LPG];
            // exclusive products over the planes, per feature
            float4 ex[HEX_PLANES];
            float4 run = go;
#pragma unroll
            for (int p = 0; p < HEX_PLANES; ++p) {
                ex[p] = run;
                run = f4_mul(run, val[p]);
            }
            run = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
            for (int p = HEX_PLANES - 1; p >= 0; --p) {
                ex[p] = f4_mul(ex[p], run);   // = go * prod_{q != p} val[q]  = d loss / d val[p]
                run = f4_mul(run, val[p]);
            }
#pragma unroll
            for (int p = 0; p < HEX_PLANES; ++p) {
                const HexAxis X = ax[HEX_AX(p)], Y = ax[HEX_AY(p)];
                const int W = P.reso[s][HEX_AX(p)];
                float4* gb = reinterpret_cast<float4*>(v_planes + P.off[s][p]) + q;
                const int64_t r0 = (int64_t)Y.i0 * W, r1 = (int64_t)Y.i1 * W;
                const float wx1 = X.w1, wx0 = 1.f - X.w1, wy1 = Y.w1, wy0 = 1.f - Y.w1;
                atomicAdd(gb + (r0 + X.i0) * HEX_LPG, f4_scale(ex[p], wx0 * wy0));
                if (wx1 != 0.f) atomicAdd(gb + (r0 + X.i1) * HEX_LPG, f4_scale(ex[p], wx1 * wy0));
                if (wy1 != 0.f) {
                    atomicAdd(gb + (r1 + X.i0) * HEX_LPG, f4_scale(ex[p], wx0 * wy1));
                    if (wx1 != 0.f) atomicAdd(gb + (r1 + X.i1) * HEX_LPG, f4_scale(ex[p], wx1 * wy1));
                }
                g[HEX_AX(p)] = fmaf(X.dmul, f4_dot(ex[p], dvx[p]), g[HEX_AX(p)]);
                g[HEX_AY(p)] = fmaf(Y.dmul, f4_dot(ex[p], dvy[p]), g[HEX_AY(p)]);
            }
        }
    }
    // sum the eight feature lanes of a Gaussian (lanes of one group are contiguous in the warp)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        g[c] += __shfl_xor_sync(0xffffffffu, g[c], 1);
        g[c] += __shfl_xor_sync(0xffffffffu, g[c], 2);
        g[c] += __shfl_xor_sync(0xffffffffu, g[c], 4);
    }
    if (n < N && q == 0) {
        if (v_pts) {
            v_pts[n * 3 + 0] = g[0] * P.k[0];
            v_pts[n * 3 + 1] = g[1] * P.k[1];
            v_pts[n * 3 + 2] = g[2] * P.k[2];
        }
        if (t_stride != 0 && v_t) v_t[n] = g[3];
    }
    if (t_stride == 0 && t_partial) {
        // shared time: fixed-order block sum of the per-Gaussian values (one per group of 8 lanes)
        float gt = (n < N && q == 0) ? g[3] : 0.f;
        gt += __shfl_xor_sync(0xffffffffu, gt, 8);
        gt += __shfl_xor_sync(0xffffffffu, gt, 16);
        if ((threadIdx.x & 31) == 0) s_t[threadIdx.x >> 5] = gt;
        __syncthreads();
        if (threadIdx.x == 0) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < HEX_THREADS / 32; ++w) acc += s_t[w];
            t_partial[blockIdx.x] = acc;
        }
    }
}

// v_t[0] += sum of the per-block partials, fixed order (one block)
__global__ void __launch_bounds__(256) hexplane_tsum_kernel(const float* __restrict__ partial, int64_t n, float* __restrict__ v_t) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += (double)partial[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) v_t[0] += (float)s[0];
}

static int hex_fill_params(HexParams& P, const char* fn, const float* planes, const int64_t* plane_offsets, const int* reso,
                           int S, int F, const float* aabb) {
    EMD_CHECK_ARG(S >= 1 && S <= HEX_MAX_SCALES, "%s: S=%d outside [1,%d]", fn, S, HEX_MAX_SCALES);
    EMD_CHECK_ARG(F == HEX_F, "%s: F=%d; this library is built for %d features per plane", fn, F, HEX_F);
    EMD_CHECK_ARG(planes && plane_offsets && reso && aabb, "%s: null argument", fn);
    if (!emd_aligned(planes, 16)) {
        emd_set_error("%s: planes must be 16-byte aligned", fn);
        return EMD_ERR_ALIGN;
    }
    P.planes = planes;
    P.S = S;
    for (int s = 0; s < S; ++s) {
        for (int c = 0; c < 4; ++c) {
            P.reso[s][c] = reso[s * 4 + c];
            EMD_CHECK_ARG(P.reso[s][c] >= 1, "%s: reso[%d][%d]=%d", fn, s, c, P.reso[s][c]);
        }
        for (int p = 0; p < HEX_PLANES; ++p) {
            P.off[s][p] = plane_offsets[s * HEX_PLANES + p];
            EMD_CHECK_ARG(P.off[s][p] >= 0 && P.off[s][p] % 4 == 0, "%s: plane offset %lld not a multiple of 4 floats", fn,
                          (long long)P.off[s][p]);
        }
    }
    for (int a = 0; a < 3; ++a) {
        P.a0[a] = aabb[a];
        P.k[a] = 2.0f / (aabb[3 + a] - aabb[a]);
    }
    return EMD_OK;
}

// feat[N, S*F] = HexPlaneField(pts, t).  planes: flat device buffer, plane (s, p) feature-last [H][W][F] at float
// offset plane_offsets[s*6+p] (HOST array), H = reso[s*4 + j], W = reso[s*4 + i] for the coordinate pair (i, j) of p
// (HOST array reso[S*4]); aabb: HOST float[6] = {aabb[0], aabb[1]}; t: DEVICE, one value (t_stride 0) or N (t_stride 1).
// emd_hexplane_fwd_ld: the same with a row pitch -- row n of the features starts at feat + n * ld floats (ld >= S*F, a
// multiple of 4), so the gather can write straight into the left columns of the deformation MLP's input [N, S*F + E]
// (deformation.py:205 concatenates the per-Gaussian embedding behind the features) without a concatenation pass.
extern "C" int emd_hexplane_fwd_ld(const float* planes, const int64_t* plane_offsets, const int* reso, int S, int F,
                                   const float* aabb, const float* pts, const float* t, int t_stride, int64_t N, float* feat,
                                   int64_t ld, cudaStream_t stream) {
    HexParams P;
    int rc = hex_fill_params(P, "emd_hexplane_fwd", planes, plane_offsets, reso, S, F, aabb);
    if (rc != EMD_OK) return rc;
    EMD_CHECK_ARG(N >= 0 && (t_stride == 0 || t_stride == 1), "emd_hexplane_fwd: N=%lld t_stride=%d", (long long)N, t_stride);
    EMD_CHECK_ARG(ld >= (int64_t)S * F && ld % 4 == 0, "emd_hexplane_fwd: row pitch %lld (need >= %d and a multiple of 4)",
                  (long long)ld, S * F);
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(pts && t && feat, "emd_hexplane_fwd: null argument");
    if (!emd_aligned(feat, 16)) {
        emd_set_error("emd_hexplane_fwd: feat must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    const unsigned grid = (unsigned)emd_cdiv(N, HEX_GPB);
    EMD_LAUNCH(EK_HEX_FWD, stream, (hexplane_fwd_kernel<<<grid, HEX_THREADS, 0, stream>>>(P, pts, t, t_stride, N, feat, ld)));
    EMD_CHECK_LAUNCH("emd_hexplane_fwd");
    return EMD_OK;
}

extern "C" int emd_hexplane_fwd(const float* planes, const int64_t* plane_offsets, const int* reso, int S, int F,
                                const float* aabb, const float* pts, const float* t, int t_stride, int64_t N, float* feat,
                                cudaStream_t stream) {
    return emd_hexplane_fwd_ld(planes, plane_offsets, reso, S, F, aabb, pts, t, t_stride, N, feat, (int64_t)S * F, stream);
}

extern "C" size_t emd_hexplane_bwd_workspace_bytes(int64_t N) { return (size_t)(emd_cdiv(N > 0 ? N : 1, HEX_GPB)) * sizeof(float); }

// VJP of emd_hexplane_fwd.  v_planes (same layout as planes) is ADDED into: the caller zero-fills it once per step.
// v_pts[N,3] may be NULL.  v_t: N values written (t_stride 1) or one value ADDED into (t_stride 0); may be NULL.
// emd_hexplane_bwd_ld: v_feat with a row pitch of ld floats (see emd_hexplane_fwd_ld).
extern "C" int emd_hexplane_bwd_ld(const float* planes, const int64_t* plane_offsets, const int* reso, int S, int F,
                                   const float* aabb, const float* pts, const float* t, int t_stride, int64_t N,
                                   const float* v_feat, int64_t ld, float* v_planes, float* v_pts, float* v_t,
                                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    HexParams P;
    int rc = hex_fill_params(P, "emd_hexplane_bwd", planes, plane_offsets, reso, S, F, aabb);
    if (rc != EMD_OK) return rc;
    EMD_CHECK_ARG(N >= 0 && (t_stride == 0 || t_stride == 1), "emd_hexplane_bwd: N=%lld t_stride=%d", (long long)N, t_stride);
    EMD_CHECK_ARG(ld >= (int64_t)S * F && ld % 4 == 0, "emd_hexplane_bwd: row pitch %lld (need >= %d and a multiple of 4)",
                  (long long)ld, S * F);
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(pts && t && v_feat && v_planes, "emd_hexplane_bwd: null argument");
    if (!emd_aligned(v_feat, 16) || !emd_aligned(v_planes, 16)) {
        emd_set_error("emd_hexplane_bwd: v_feat / v_planes must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    const unsigned grid = (unsigned)emd_cdiv(N, HEX_GPB);
    float* partial = nullptr;
    if (t_stride == 0 && v_t) {
        if (!workspace || workspace_bytes < emd_hexplane_bwd_workspace_bytes(N)) {
            emd_set_error("emd_hexplane_bwd: workspace too small (%zu < %zu)", workspace_bytes, emd_hexplane_bwd_workspace_bytes(N));
            return EMD_ERR_WORKSPACE;
        }
        partial = static_cast<float*>(workspace);
    }
    EMD_LAUNCH(EK_HEX_BWD, stream,
               (hexplane_bwd_kernel<<<grid, HEX_THREADS, 0, stream>>>(P, pts, t, t_stride, N, v_feat, ld, v_planes, v_pts, v_t, partial)));
    EMD_CHECK_LAUNCH("emd_hexplane_bwd");
    if (partial) {
        EMD_LAUNCH(EK_MISC, stream, (hexplane_tsum_kernel<<<1, 256, 0, stream>>>(partial, (int64_t)grid, v_t)));
        EMD_CHECK_LAUNCH("emd_hexplane_bwd(tsum)");
    }
    return EMD_OK;
}

extern "C" int emd_hexplane_bwd(const float* planes, const int64_t* plane_offsets, const int* reso, int S, int F,
                                const float* aabb, const float* pts, const float* t, int t_stride, int64_t N,
                                const float* v_feat, float* v_planes, float* v_pts, float* v_t, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
    return emd_hexplane_bwd_ld(planes, plane_offsets, reso, S, F, aabb, pts, t, t_stride, N, v_feat, (int64_t)S * F, v_planes,
                               v_pts, v_t, workspace, workspace_bytes, stream);
}
