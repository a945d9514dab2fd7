// Host-only build of proj_math.cuh (g++ -ffp-contract=off) exposing the product's
// projection math to the CPU test-suite: lets `pytest -m "not gpu"` check the
// canonical op order (bit-exact radii / tile rects / depth bits) and the VJP
// against the oracle without a GPU.  Test support; never loaded by the product path.
#include <stdint.h>
#include <string.h>

#include "proj_math.cuh"

extern "C" void emd_host_projection_fwd(const float* means, const float* quats, const float* scales,
                                        const float* viewmats, const float* Ks, int64_t N, int64_t C, int width,
                                        int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                                        int tile_w, int tile_h, int32_t* radii, float* means2d, float* depths,
                                        float* conics, float* comps, int32_t* tiles_per_gauss, int32_t* rects) {
    for (int64_t c = 0; c < C; ++c) {
        CamConst cam;
        make_cam_const(viewmats + c * 16, Ks + c * 9, width, height, cam);
        for (int64_t i = 0; i < N; ++i) {
            float R[9], M[9], S[6], qn[4];
            quat_to_rotmat_c(quats + i * 4, R, qn);
            covar_world_c(R, scales + i * 3, M, S);
            ProjFwd o;
            memset(&o, 0, sizeof(o));
            project_gaussian_c(means + i * 3, S, cam, width, height, eps2d, near_plane, far_plane, radius_clip, o);
            const int64_t ci = c * N + i;
            const bool vis = o.radius > 0;
            radii[ci] = vis ? o.radius : 0;
            means2d[ci * 2 + 0] = vis ? o.m2x : 0.f;
            means2d[ci * 2 + 1] = vis ? o.m2y : 0.f;
            depths[ci] = vis ? o.z : 0.f;
            conics[ci * 3 + 0] = vis ? o.conic_a : 0.f;
            conics[ci * 3 + 1] = vis ? o.conic_b : 0.f;
            conics[ci * 3 + 2] = vis ? o.conic_c : 0.f;
            comps[ci] = vis ? o.comp : 0.f;
            int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
            if (vis) tile_rect_c(o.m2x, o.m2y, o.radius, tile_w, tile_h, x0, y0, x1, y1);
            tiles_per_gauss[ci] = (x1 - x0) * (y1 - y0);
            rects[ci * 4 + 0] = x0; rects[ci * 4 + 1] = y0; rects[ci * 4 + 2] = x1; rects[ci * 4 + 3] = y1;
        }
    }
}

extern "C" void emd_host_projection_bwd(const float* means, const float* quats, const float* scales,
                                        const float* viewmats, const float* Ks, int64_t N, int64_t C, int width,
                                        int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                                        const float* v_means2d, const float* v_depths, const float* v_conics,
                                        float* v_means, float* v_quats, float* v_scales) {
    for (int64_t i = 0; i < N; ++i) {
        float R[9], M[9], S[6], qn[4];
        const float inv_norm = quat_to_rotmat_c(quats + i * 4, R, qn);
        covar_world_c(R, scales + i * 3, M, S);
        float v_mean[3] = {0, 0, 0}, v_S[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t c = 0; c < C; ++c) {
            CamConst cam;
            make_cam_const(viewmats + c * 16, Ks + c * 9, width, height, cam);
            ProjFwd o;
            memset(&o, 0, sizeof(o));
            project_gaussian_c(means + i * 3, S, cam, width, height, eps2d, near_plane, far_plane, radius_clip, o);
            if (o.radius <= 0) continue;
            const int64_t ci = c * N + i;
            project_gaussian_vjp(o, cam, v_means2d[ci * 2], v_means2d[ci * 2 + 1], v_depths[ci], v_conics[ci * 3],
                                 v_conics[ci * 3 + 1], v_conics[ci * 3 + 2], v_mean, v_S);
        }
        float v_q[4], v_s[3];
        covar_world_vjp(qn, inv_norm, R, M, scales + i * 3, v_S, v_q, v_s);
        for (int k = 0; k < 3; ++k) { v_means[i * 3 + k] = v_mean[k]; v_scales[i * 3 + k] = v_s[k]; }
        for (int k = 0; k < 4; ++k) v_quats[i * 4 + k] = v_q[k];
    }
}

// ---- per-instance EMD heads (emd_math.cuh) -------------------------------------------------
#include "emd_math.cuh"

extern "C" int emd_host_rigid_param_count(int d, int g) { return rigid_param_count(d + g); }

// heads: 8 host pointers in the order documented at emd_rigid_deform_fwd
extern "C" void emd_host_rigid_instance_fwd(const float* table, int I, int E, int d, int g, const float* mean_emb,
                                            float t, int cur_c, int cur_f, const float* const* heads,
                                            const float* pose_q_means, const float* pose_q_quats, const float* pose_t,
                                            float* inst_out /*[I][16]*/) {
    RigidHeads H{heads[0], heads[1], heads[2], heads[3], heads[4], heads[5], heads[6], heads[7]};
    for (int i = 0; i < I; ++i) {
        RigidInstOut o;
        rigid_instance_fwd(table + (int64_t)i * E * d, E, d, g, mean_emb + i * g, t, cur_c, cur_f, H,
                           pose_q_means + i * 4, pose_q_quats + i * 4, pose_t + i * 3, o);
        float* out = inst_out + i * 16;
        for (int k = 0; k < 9; ++k) out[k] = o.R[k];
        for (int k = 0; k < 3; ++k) out[9 + k] = o.t[k];
        for (int k = 0; k < 4; ++k) out[12 + k] = o.Q[k];
    }
}

extern "C" void emd_host_rigid_instance_bwd(const float* table, int I, int E, int d, int g, const float* mean_emb,
                                            float t, int cur_c, int cur_f, const float* const* heads,
                                            const float* pose_q_means, const float* pose_q_quats, const float* pose_t,
                                            const float* v_inst /*[I][16]*/, float* v_pose_q_means,
                                            float* v_pose_q_quats, float* v_pose_t, float* v_params /*[pc], summed*/,
                                            float* v_table /*[I][E][d] zeroed*/, float* v_mean_emb) {
    RigidHeads H{heads[0], heads[1], heads[2], heads[3], heads[4], heads[5], heads[6], heads[7]};
    const int pc = rigid_param_count(d + g);
    float* part = new float[pc];
    for (int k = 0; k < pc; ++k) v_params[k] = 0.f;
    for (int i = 0; i < I; ++i) {
        const float* v = v_inst + i * 16;
        rigid_instance_bwd(table + (int64_t)i * E * d, E, d, g, mean_emb + i * g, t, cur_c, cur_f, H,
                           pose_q_means + i * 4, pose_q_quats + i * 4, pose_t + i * 3, v, v + 9, v + 12,
                           v_pose_q_means + i * 4, v_pose_q_quats + i * 4, v_pose_t + i * 3, part,
                           v_table + (int64_t)i * E * d, v_mean_emb + i * g);
        for (int k = 0; k < pc; ++k) v_params[k] += part[k];
    }
    delete[] part;
}

// ---- diff_gauss preprocess (dg_math.cuh) ---------------------------------------------------------
#include "dg_math.cuh"

extern "C" void emd_host_dg_preprocess_fwd(const float* means, const float* scales, const float* rots,
                                           const float* viewmatrix, const float* projmatrix, float tanfovx,
                                           float tanfovy, int W, int H, float mod, int64_t N, int32_t* radii,
                                           float* means2d, float* depths, float* conics, int32_t* rects) {
    DgCam cam;
    make_dg_cam(viewmatrix, projmatrix, tanfovx, tanfovy, W, H, mod, cam);
    for (int64_t n = 0; n < N; ++n) {
        float s_mod[3], R[9], M[9], S[6];
        for (int k = 0; k < 3; ++k) s_mod[k] = c_mul(mod, scales[n * 3 + k]);
        dg_quat_to_rotmat_c(rots + n * 4, R);
        covar_world_c(R, s_mod, M, S);
        DgFwd o;
        memset(&o, 0, sizeof(o));
        dg_project_c(means + n * 3, S, cam, o);
        const bool vis = o.f.radius > 0;
        radii[n] = vis ? o.f.radius : 0;
        means2d[n * 2] = vis ? o.f.m2x : 0.f; means2d[n * 2 + 1] = vis ? o.f.m2y : 0.f;
        depths[n] = vis ? o.f.z : 0.f;
        conics[n * 3] = vis ? o.f.conic_a : 0.f; conics[n * 3 + 1] = vis ? o.f.conic_b : 0.f; conics[n * 3 + 2] = vis ? o.f.conic_c : 0.f;
        rects[n * 4] = vis ? o.x0 : 0; rects[n * 4 + 1] = vis ? o.y0 : 0; rects[n * 4 + 2] = vis ? o.x1 : 0; rects[n * 4 + 3] = vis ? o.y1 : 0;
    }
}

// ---- HexPlane (hexplane_math.cuh): scalar host loop over the same tap arithmetic the kernels use ------------
#include "hexplane_math.cuh"

struct HostHexGeom {
    const float* planes;
    const int64_t* off;
    const int* reso;
    float a0[3], k[3];
};

static void host_hex_eval(const HostHexGeom& G, int s, const float u[4], int ch, float val[HEX_PLANES], float dvx[HEX_PLANES],
                          float dvy[HEX_PLANES], HexAxis ax[4]) {
    for (int c = 0; c < 4; ++c) ax[c] = hex_axis(u[c], G.reso[s * 4 + c]);
    for (int p = 0; p < HEX_PLANES; ++p) {
        const HexAxis X = ax[HEX_AX(p)], Y = ax[HEX_AY(p)];
        const int W = G.reso[s * 4 + HEX_AX(p)];
        const float* base = G.planes + G.off[s * HEX_PLANES + p] + ch;
        const float nw = base[((int64_t)Y.i0 * W + X.i0) * HEX_F], ne = base[((int64_t)Y.i0 * W + X.i1) * HEX_F];
        const float sw = base[((int64_t)Y.i1 * W + X.i0) * HEX_F], se = base[((int64_t)Y.i1 * W + X.i1) * HEX_F];
        const float top = nw + X.w1 * (ne - nw), bot = sw + X.w1 * (se - sw);
        val[p] = top + Y.w1 * (bot - top);
        dvx[p] = (ne - nw) + Y.w1 * ((se - sw) - (ne - nw));
        dvy[p] = bot - top;
    }
}

extern "C" void emd_host_hexplane(const float* planes, const int64_t* plane_offsets, const int* reso, int S,
                                  const float* aabb, const float* pts, const float* t, int t_stride, int64_t N,
                                  float* feat, const float* v_feat, float* v_planes, float* v_pts, float* v_t) {
    HostHexGeom G;
    G.planes = planes; G.off = plane_offsets; G.reso = reso;
    for (int a = 0; a < 3; ++a) { G.a0[a] = aabb[a]; G.k[a] = 2.0f / (aabb[3 + a] - aabb[a]); }
    for (int64_t n = 0; n < N; ++n) {
        float u[4];
        for (int a = 0; a < 3; ++a) u[a] = hex_normalize(pts[n * 3 + a], G.a0[a], G.k[a]);
        u[3] = t[n * t_stride];
        float g[4] = {0, 0, 0, 0};
        for (int s = 0; s < S; ++s) {
            for (int ch = 0; ch < HEX_F; ++ch) {
                float val[HEX_PLANES], dvx[HEX_PLANES], dvy[HEX_PLANES], ex[HEX_PLANES];
                HexAxis ax[4];
                host_hex_eval(G, s, u, ch, val, dvx, dvy, ax);
                float prod = 1.0f;
                for (int p = 0; p < HEX_PLANES; ++p) prod *= val[p];
                feat[n * (int64_t)(S * HEX_F) + s * HEX_F + ch] = prod;
                if (!v_feat) continue;
                const float go = v_feat[n * (int64_t)(S * HEX_F) + s * HEX_F + ch];
                hex_excl_products(val, ex);
                for (int p = 0; p < HEX_PLANES; ++p) {
                    const float e = go * ex[p];
                    const HexAxis X = ax[HEX_AX(p)], Y = ax[HEX_AY(p)];
                    const int W = reso[s * 4 + HEX_AX(p)];
                    float* gb = v_planes + plane_offsets[s * HEX_PLANES + p] + ch;
                    gb[((int64_t)Y.i0 * W + X.i0) * HEX_F] += e * ((1.f - X.w1) * (1.f - Y.w1));
                    gb[((int64_t)Y.i0 * W + X.i1) * HEX_F] += e * (X.w1 * (1.f - Y.w1));
                    gb[((int64_t)Y.i1 * W + X.i0) * HEX_F] += e * ((1.f - X.w1) * Y.w1);
                    gb[((int64_t)Y.i1 * W + X.i1) * HEX_F] += e * (X.w1 * Y.w1);
                    g[HEX_AX(p)] += X.dmul * (e * dvx[p]);
                    g[HEX_AY(p)] += Y.dmul * (e * dvy[p]);
                }
            }
        }
        if (v_feat) {
            for (int a = 0; a < 3; ++a) v_pts[n * 3 + a] = g[a] * G.k[a];
            if (t_stride) v_t[n] = g[3]; else v_t[0] += g[3];
        }
    }
}

// ---- Adam (adam_math.cuh) -----------------------------------------------------------------------------------
#include "adam_math.cuh"

extern "C" void emd_host_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, double lr, double beta1,
                                   double beta2, double eps, double weight_decay, int64_t step, double grad_scale) {
    const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
    for (int64_t i = 0; i < numel; ++i) adam_update(p[i], g[i], m[i], v[i], s);
}
