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

// ---- Voxel LBS weights (voxel_math.cuh): scalar host loop over the same tap arithmetic the kernels use -----------
#include "voxel_math.cuh"

// grids channel-last [B,D,H,W,J]; out[B,V,J]; when v_out != NULL also the VJP: v_corr (added into), v_xc[B,V,3]
extern "C" void emd_host_voxel_lbs(const float* base, const float* corr, const float* offset, const float* scale, float ratio,
                                   int ratio_dim, int B, int D, int H, int W, int J, const float* xc, int64_t V, float* out,
                                   const float* v_out, float* v_corr, float* v_xc) {
    VoxGeom G;
    G.D = D; G.H = H; G.W = W; G.J = J; G.ratio = ratio; G.ratio_dim = ratio_dim;
    for (int64_t n = 0; n < (int64_t)B * V; ++n) {
        const int b = (int)(n / V);
        VoxTap T;
        vox_tap(G, xc + n * 3, offset + b * 3, scale[b], T);
        float gx[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < J; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 8; ++k) {
                const int64_t e = vox_corner_index(G, T, b, k) * J + c;
                const float g = base[e] + (corr ? corr[e] : 0.f);
                const float w = vox_corner_weight(T, k);
                acc += w * g;
                if (v_out) {
                    const float go = v_out[n * J + c];
                    if (v_corr) v_corr[e] += go * w;
                    for (int a = 0; a < 3; ++a) gx[a] += vox_corner_dweight(T, a, k) * (go * g);
                }
            }
            out[n * J + c] = acc;
        }
        if (v_out && v_xc)
            for (int a = 0; a < 3; ++a) v_xc[n * 3 + a] = gx[a] * vox_coord_chain(G, T, a, scale[b]);
    }
}

// ---- strided GEMM of the DeformableNodes network (dense_math.cuh): the kernel's tile logic run thread by thread ------
#include <vector>
#include "dense_math.cuh"

template <bool TA, bool TB>
static void host_sgemm(const GemmArgs& g, int splits) {
    static float As[DG_BK][DG_PITCH], Bs[DG_BK][DG_PITCH];
    static float ra[DG_THREADS][8], rb[DG_THREADS][8], acc[DG_THREADS][8][8];
    for (int64_t bz = 0; bz < splits; ++bz)
        for (int64_t by = 0; by < dg_cdiv(g.N, DG_BN); ++by)
            for (int64_t bx = 0; bx < dg_cdiv(g.M, DG_BM); ++bx) {
                const int64_t m0 = bx * DG_BM, n0 = by * DG_BN;
                const int64_t kbeg = bz * g.k_per_split;
                const int64_t kend = g.K < kbeg + g.k_per_split ? g.K : kbeg + g.k_per_split;
                float* C = g.C + bz * g.split_stride;
                for (int t = 0; t < DG_THREADS; ++t)
                    for (int i = 0; i < 8; ++i)
                        for (int j = 0; j < 8; ++j) acc[t][i][j] = 0.f;
                static GemmLoadState S[DG_THREADS];
                for (int t = 0; t < DG_THREADS; ++t) gemm_prepare<TA, TB>(g, t, m0, n0, kbeg, S[t]);
                if (kbeg < kend)
                    for (int t = 0; t < DG_THREADS; ++t) gemm_load_tile<TA, TB>(S[t], kbeg, kend, ra[t], rb[t]);
                for (int64_t k0 = kbeg; k0 < kend; k0 += DG_BK) {
                    for (int t = 0; t < DG_THREADS; ++t) gemm_store<TA, TB>(t, ra[t], rb[t], As, Bs);
                    if (k0 + DG_BK < kend)
                        for (int t = 0; t < DG_THREADS; ++t) {
                            gemm_advance(S[t]);
                            gemm_load_tile<TA, TB>(S[t], k0 + DG_BK, kend, ra[t], rb[t]);
                        }
                    for (int t = 0; t < DG_THREADS; ++t) gemm_compute(t, As, Bs, acc[t]);
                }
                for (int t = 0; t < DG_THREADS; ++t) gemm_epilogue(g, t, m0, n0, C, acc[t]);
            }
}

extern "C" void emd_host_dense_fwd(const float* X, int64_t ldx, const float* W, const float* b, int64_t M, int K, int Nout,
                                   int relu_out, float* Y, int64_t ldy) {
    host_sgemm<false, true>(dense_fwd_args(X, ldx, W, b, M, K, Nout, relu_out, Y, ldy), 1);
}

extern "C" void emd_host_dense_bwd(const float* X, int64_t ldx, const float* W, const float* dZ, int64_t lddz, int64_t M, int K,
                                   int Nout, float* dX, int64_t lddx, int col0, int ncols, const float* mask, int64_t ldmask,
                                   float* dW, float* db) {
    if (dX) host_sgemm<false, false>(dense_dgrad_args(W, dZ, lddz, M, K, Nout, dX, lddx, col0, ncols, mask, ldmask), 1);
    const DenseSplit s = dense_split(M, K, Nout);
    if (dW) {
        std::vector<float> wpart((size_t)s.splits * Nout * K + 4, -777.0f);   // poisoned: every element must be written
        float* wp = wpart.data();
        while (!dg_aligned16(wp)) ++wp;
        host_sgemm<true, false>(dense_wgrad_args(X, ldx, dZ, lddz, M, K, Nout, s, wp), s.splits);
        const int64_t n = (int64_t)Nout * K;
        for (int64_t i = 0; i < n; ++i) {
            float a = 0.f;
            for (int z = 0; z < s.splits; ++z) a += wp[(int64_t)z * n + i];
            dW[i] = a;
        }
    }
    if (db) {
        std::vector<float> bpart((size_t)s.col_chunks * Nout);
        for (int ch = 0; ch < s.col_chunks; ++ch) {
            const int64_t r0 = ch * s.rows_per_chunk, r1 = M < r0 + s.rows_per_chunk ? M : r0 + s.rows_per_chunk;
            for (int c = 0; c < Nout; ++c) {
                float a = 0.f;
                for (int64_t r = r0; r < r1; ++r) a += dZ[r * lddz + c];
                bpart[(size_t)ch * Nout + c] = a;
            }
        }
        for (int c = 0; c < Nout; ++c) {
            float a = 0.f;
            for (int ch = 0; ch < s.col_chunks; ++ch) a += bpart[(size_t)ch * Nout + c];
            db[c] = a;
        }
    }
}

// ---- SMPL LBS-weight gradient (emd_math.cuh: smpl_point_weight_grad), one point per row ---------------------------------
extern "C" void emd_host_smpl_weight_grad(const float* Wn, const float* A, const float* x, const float* q, const float* g,
                                          const float* vg, int64_t N, float* v_W) {
    for (int64_t n = 0; n < N; ++n)
        smpl_point_weight_grad(Wn + n * SMPL_J, A, x + n * 3, q + n * 4, g + n * 3, vg + n * 4, v_W + n * SMPL_J);
}

// ---- tensor-core staging maps (tc_stage_math.cuh): replay every thread's shared-memory writes of one chunk, then read
//      the operand back the way the UMMA descriptor walks it (canonical K-major layout) -------------------------------------
#include <string.h>
#include "tc_stage_math.cuh"

// which = 0: A chunk [128 x 32] from src[128][32]; 1: B chunk forward from src[n][32] (n < npad); 2: B chunk dgrad from
// src[32][npad]; 3: A chunk of the weight gradient from src[32][128] (operand rows contiguous in memory).  out[rows][32] = the operand as k-step descriptors (start = kstep offset, LBO, SBO = 128) expose it;
// returns the number of bytes written more than once (must be 0) and fills *untouched with the count of operand bytes
// of the [rows x 32] region no thread wrote.
extern "C" int emd_host_tc_stage_replay(int which, int npad, const float* src, float* out, int* untouched) {
    const int lbo = (which == 0 || which == 3) ? DTS_A_LBO : DTS_B_LBO;
    const int rows = (which == 0 || which == 3) ? DTS_ROWS : npad;
    const int bytes = (DTS_KC / 4) * lbo;
    std::vector<unsigned char> mem(bytes, 0), hit(bytes, 0);
    int twice = 0;
    auto put = [&](int off, const float* v, int n) {
        for (int b = 0; b < 4 * n; ++b) { twice += hit[off + b]; hit[off + b] = 1; }
        memcpy(mem.data() + off, v, 4 * n);
    };
    for (int tid = 0; tid < DTS_THREADS; ++tid) {
        if (which == 0) {
            for (int i = 0; i < 4; ++i) {
                int r, cj;
                dts_a_elem(tid, i, r, cj);
                put(dts_a_store_offset(r, cj), src + r * DTS_KC + 4 * cj, 4);
            }
        } else if (which == 1) {
            for (int i = 0; i < 8; ++i) {
                int n, j;
                dts_b_elem_fwd(tid, i, n, j);
                if (n < npad) put(dts_b_store_offset_fwd(n, j), src + n * DTS_KC + 4 * j, 4);
            }
        } else if (which == 3) {
            for (int i = 0; i < 4; ++i) {
                int k, m;
                dts_at_elem(tid, i, k, m);
                for (int q = 0; q < 4; ++q) put(dts_at_store_offset(k, m + q), src + k * DTS_ROWS + m + q, 1);
            }
        } else {
            for (int i = 0; i < 8; ++i) {
                int k, n;
                dts_b_elem_dgrad(tid, i, k, n);
                if (n < npad)
                    for (int q = 0; q < 4; ++q) put(dts_b_store_offset_dgrad(k, n + q), src + k * npad + n + q, 1);
            }
        }
    }
    int miss = 0;
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < DTS_KC; ++k) {
            // k-step sl = k / 8 starts at dts_kstep_offset(sl); inside it the element is column k % 8
            const int off = dts_kstep_offset(k / 8, lbo) + dts_canonical_offset(r, k % 8, lbo);
            for (int b = 0; b < 4; ++b) miss += !hit[off + b];
            memcpy(out + r * DTS_KC + k, mem.data() + off, 4);
        }
    *untouched = miss;
    return twice;
}

// ---- Adam (adam_math.cuh) -----------------------------------------------------------------------------------
#include "adam_math.cuh"

extern "C" void emd_host_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, double lr, double beta1,
                                   double beta2, double eps, double weight_decay, int64_t step, double grad_scale) {
    const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
    for (int64_t i = 0; i < numel; ++i) adam_update(p[i], g[i], m[i], v[i], s);
}

// ---- fused image losses (loss_math.cuh) ----------------------------------------------------------------------
// Same per-pixel functions, same formulas and normalisers as image_loss.cu, as plain loops (windowed sums taken directly,
// no tiling): terms[C][EMD_LOSS_TERMS] and, when v_terms is given, the cotangents.
#include <vector>
#include "loss_math.cuh"

namespace {
struct HostLoss {
    const float *rgb, *depth, *alpha, *sky, *gt, *valid_mask, *sky_mask, *lidar;
    int C, H, W;
    EmdImageLossConfig g;
    const float* win;
    bool in(int y, int x) const { return y >= 0 && y < H && x >= 0 && x < W; }
    bool in_map(int y, int x) const {
        if (!in(y, x)) return false;
        return g.ssim_pad ? true : (y >= SSIM_R && y < H - SSIM_R && x >= SSIM_R && x < W - SSIM_R);
    }
    float valid(int c, int y, int x) const { return valid_mask ? valid_mask[((int64_t)c * H + y) * W + x] : 1.0f; }
    LossBlend blend(int c, int y, int x, int ch) const {
        const int64_t pix = (int64_t)y * W + x;
        const float rg = rgb[c * g.rgb_vs + pix * g.rgb_ps + ch * g.rgb_cs];
        const float al = sky ? alpha[(int64_t)c * H * W + pix] : 0.0f;
        const float sk = sky ? sky[c * g.sky_vs + pix * g.sky_ps + ch * g.sky_cs] : 0.0f;
        return loss_blend(rg, al, sk, sky != nullptr, g.blend);
    }
    void pg(int c, int y, int x, int ch, float& p, float& q) const {
        p = q = 0.0f;
        if (!in(y, x)) return;
        const float v = valid(c, y, x);
        p = blend(c, y, x, ch).p * v;
        q = gt[c * g.gt_vs + ((int64_t)y * W + x) * g.gt_ps + ch * g.gt_cs] * v;
    }
    void gt3(int c, int y, int x, float o[3]) const {
        for (int ch = 0; ch < 3; ++ch) o[ch] = gt[c * g.gt_vs + ((int64_t)y * W + x) * g.gt_ps + ch * g.gt_cs];
    }
    float dep(int c, int y, int x) const { return depth[c * g.depth_vs + ((int64_t)y * W + x) * g.depth_ps]; }
    float hit(int c, int y, int x, float li) const {
        const int64_t vpix = ((int64_t)c * H + y) * W + x;
        if (g.depth_mask_mode == 0) return (li > 0.0f ? 1.0f : 0.0f) * valid(c, y, x);
        return sky_mask ? 1.0f - sky_mask[vpix] : 1.0f;
    }
};
}  // namespace

extern "C" void emd_host_image_loss(const float* rgb, const float* depth, const float* alpha, const float* sky,
                                    const float* gt, const float* valid_mask, const float* sky_mask, const float* lidar,
                                    int C, int H, int W, const EmdImageLossConfig* cfg, const float* window, float* terms,
                                    const float* v_terms, float* v_rgb, float* v_depth, float* v_alpha, float* v_sky) {
    HostLoss A{rgb, depth, alpha, sky, gt, valid_mask, sky_mask, lidar, C, H, W, *cfg, window};
    const EmdImageLossConfig& g = A.g;
    const double n_l1 = 3.0 * H * W, n_hw = (double)H * W, n_sx = (double)H * (W - 1), n_sy = (double)(H - 1) * W;
    const double n_ss = g.ssim_pad ? 3.0 * H * W : 3.0 * (H - 2 * SSIM_R) * (W - 2 * SSIM_R);
    const int64_t HW = (int64_t)H * W;
    for (int c = 0; c < C; ++c) {
        double S[EMD_LOSS_SUMS] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<float> maps((size_t)9 * HW, 0.0f);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int64_t pix = (int64_t)y * W + x, vpix = (int64_t)c * HW + pix;
                for (int ch = 0; ch < 3; ++ch) {
                    float p, q;
                    A.pg(c, y, x, ch, p, q);
                    S[0] += fabsf(q - p);
                    if (!A.in_map(y, x)) continue;
                    float m[5] = {0, 0, 0, 0, 0};
                    for (int j = 0; j < EMD_SSIM_TAPS; ++j) {
                        float h[5] = {0, 0, 0, 0, 0};
                        for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                            float pp, qq;
                            A.pg(c, y + j - SSIM_R, x + k - SSIM_R, ch, pp, qq);
                            const float w = window[k];
                            h[0] += w * pp; h[1] += w * qq; h[2] += w * pp * pp; h[3] += w * qq * qq; h[4] += w * pp * qq;
                        }
                        for (int i = 0; i < 5; ++i) m[i] += window[j] * h[i];
                    }
                    const SsimPoint s = ssim_point(m[0], m[1], m[2], m[3], m[4]);
                    S[1] += s.m;
                    maps[(ch * 3 + 0) * HW + pix] = s.d_mu;
                    maps[(ch * 3 + 1) * HW + pix] = s.d_pp;
                    maps[(ch * 3 + 2) * HW + pix] = s.d_pg;
                }
                const float v = A.valid(c, y, x), al = alpha[vpix];
                float l, d;
                if (sky_mask) {
                    loss_opacity(al * v, (1.0f - sky_mask[vpix]) * v, g.opacity_loss, g.bce_limit, l, d);
                    S[2] += l;
                }
                loss_entropy(al, l, d);
                S[5] += l;
                if (depth) {
                    const float de = A.dep(c, y, x);
                    if (lidar) {
                        const float li = lidar[vpix], h = A.hit(c, y, x, li);
                        float e, dd;
                        if (loss_depth(de * h, li * h, g, e, dd)) { S[3] += e; S[4] += 1.0; }
                    }
                    const float id = loss_inv_depth(de);
                    float g0[3], g1[3];
                    A.gt3(c, y, x, g0);
                    if (x + 1 < W) { A.gt3(c, y, x + 1, g1); S[6] += fabsf(id - loss_inv_depth(A.dep(c, y, x + 1))) * loss_edge_weight(g0, g1); }
                    if (y + 1 < H) { A.gt3(c, y + 1, x, g1); S[7] += fabsf(id - loss_inv_depth(A.dep(c, y + 1, x))) * loss_edge_weight(g0, g1); }
                }
            }
        float* T = terms + c * EMD_LOSS_TERMS;
        T[0] = (float)(g.w_l1 * (S[0] / n_l1));
        T[1] = (float)(g.w_ssim * (1.0 - S[1] / n_ss));
        T[2] = sky_mask ? (float)(g.w_opacity * (S[2] / n_hw)) : 0.f;
        T[3] = (depth && lidar && g.w_depth != 0.f) ? (float)(g.w_depth * (S[3] / S[4])) : 0.f;
        T[4] = (float)(g.w_entropy * (S[5] / n_hw));
        T[5] = depth ? (float)(g.w_smooth * (S[6] / n_sx + S[7] / n_sy)) : 0.f;
        if (!v_terms) continue;
        const float* vt = v_terms + c * EMD_LOSS_TERMS;
        const float g_l1 = (float)(vt[0] * g.w_l1 / n_l1), g_ss = (float)(-(double)vt[1] * g.w_ssim / n_ss);
        const float g_op = (float)(vt[2] * g.w_opacity / n_hw), g_dp = S[4] > 0 ? (float)(vt[3] * g.w_depth / S[4]) : 0.f;
        const float g_en = (float)(vt[4] * g.w_entropy / n_hw);
        const float g_sx = (float)(vt[5] * g.w_smooth / n_sx), g_sy = (float)(vt[5] * g.w_smooth / n_sy);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int64_t pix = (int64_t)y * W + x, vpix = (int64_t)c * HW + pix;
                const float v = A.valid(c, y, x), al = alpha[vpix];
                float d_alpha = 0.0f;
                for (int ch = 0; ch < 3; ++ch) {
                    float s[3] = {0, 0, 0};
                    for (int j = 0; j < EMD_SSIM_TAPS; ++j) {
                        float h[3] = {0, 0, 0};
                        for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                            const int yy = y + j - SSIM_R, xx = x + k - SSIM_R;
                            if (!A.in(yy, xx)) continue;
                            for (int i = 0; i < 3; ++i) h[i] += window[k] * maps[(ch * 3 + i) * HW + (int64_t)yy * W + xx];
                        }
                        for (int i = 0; i < 3; ++i) s[i] += window[j] * h[i];
                    }
                    const LossBlend b = A.blend(c, y, x, ch);
                    const float p = b.p * v, q = gt[c * g.gt_vs + pix * g.gt_ps + ch * g.gt_cs] * v;
                    const float dp = g_l1 * loss_sign(p - q) + g_ss * (s[0] + 2.0f * p * s[1] + q * s[2]);
                    const float db = dp * v;
                    v_rgb[c * g.rgb_vs + pix * g.rgb_ps + ch * g.rgb_cs] = db * b.d_rgb;
                    d_alpha += db * b.d_alpha;
                    if (v_sky) v_sky[c * g.sky_vs + pix * g.sky_ps + ch * g.sky_cs] = db * b.d_sky;
                }
                float l, d;
                if (sky_mask) {
                    loss_opacity(al * v, (1.0f - sky_mask[vpix]) * v, g.opacity_loss, g.bce_limit, l, d);
                    d_alpha += g_op * d * v;
                }
                loss_entropy(al, l, d);
                d_alpha += g_en * d;
                v_alpha[vpix] = d_alpha;
                if (!depth) continue;
                const float de = A.dep(c, y, x);
                float vd = 0.0f;
                if (lidar) {
                    const float li = lidar[vpix], h = A.hit(c, y, x, li);
                    float e, dd;
                    if (loss_depth(de * h, li * h, g, e, dd)) vd += g_dp * dd * h;
                }
                const float id = loss_inv_depth(de);
                float g0[3], g1[3], did = 0.0f;
                A.gt3(c, y, x, g0);
                if (x + 1 < W) { A.gt3(c, y, x + 1, g1); did += g_sx * loss_sign(id - loss_inv_depth(A.dep(c, y, x + 1))) * loss_edge_weight(g0, g1); }
                if (x > 0) { A.gt3(c, y, x - 1, g1); did -= g_sx * loss_sign(loss_inv_depth(A.dep(c, y, x - 1)) - id) * loss_edge_weight(g1, g0); }
                if (y + 1 < H) { A.gt3(c, y + 1, x, g1); did += g_sy * loss_sign(id - loss_inv_depth(A.dep(c, y + 1, x))) * loss_edge_weight(g0, g1); }
                if (y > 0) { A.gt3(c, y - 1, x, g1); did -= g_sy * loss_sign(loss_inv_depth(A.dep(c, y - 1, x)) - id) * loss_edge_weight(g1, g0); }
                vd += did * (-id * id);
                v_depth[c * g.depth_vs + pix * g.depth_ps] = vd;
            }
    }
}
