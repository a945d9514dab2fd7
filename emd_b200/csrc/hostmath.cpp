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
