// Per-instance EMD arithmetic (motion embeddings -> pose offsets), forward and VJP.
//
// Restates, for one instance, OmniRe/models/nodes/rigid.py:150-246 (temporal
// embedding resample + sample, feature concat, track_* Linear heads, yaw-offset
// quaternion) and smpl.py:401-436 (24 joint offsets).  Host/device: tests
// compile it with g++ to check it against the oracle on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef EMD_HD
#ifdef __CUDACC__
#define EMD_HD __host__ __device__ __forceinline__
#else
#define EMD_HD static inline
#endif
#endif

constexpr int EMD_TDIM_MAX = 64;   // temporal embedding width (reference: 32)
constexpr int EMD_GDIM_MAX = 16;   // per-Gaussian embedding width (reference: 4)

// emb[j] = sum_k w[k] * table[row[k]][j]  (bilinear row-resize to `cur` rows, then
// bilinear sample at time t with reflection padding; both align_corners=True)
struct TembTaps {
    int row[4];
    float w[4];
};

EMD_HD void temb_resize_taps(int r, int cur, int E, int& r0, int& r1, float& l0, float& l1) {
    const float scale = cur > 1 ? (float)(E - 1) / (float)(cur - 1) : 0.0f;
    const float src = scale * (float)r;
    r0 = (int)src;
    if (r0 > E - 1) r0 = E - 1;
    r1 = r0 + (r0 < E - 1 ? 1 : 0);
    l1 = src - (float)r0;
    l0 = 1.0f - l1;
}

EMD_HD void temb_taps(float t, int cur, int E, TembTaps& taps) {
    const float y_norm = (t - 0.5f) * 2.0f;
    float iy = ((y_norm + 1.0f) / 2.0f) * (float)(cur - 1);
    // reflection padding over [0, cur-1] (ATen reflect_coordinates with align_corners), then clip
    const float span = (float)(cur - 1);
    if (span <= 0.0f) {
        iy = 0.0f;
    } else {
        float a = fabsf(iy);
        const float extra = fmodf(a, span);
        const int flips = (int)floorf(a / span);
        iy = (flips % 2 == 0) ? extra : span - extra;
        iy = fminf(span, fmaxf(iy, 0.0f));
    }
    const int y0 = (int)floorf(iy);
    const float wy1 = iy - (float)y0, wy0 = 1.0f - wy1;
    int r0, r1; float l0, l1;
    temb_resize_taps(y0, cur, E, r0, r1, l0, l1);
    taps.row[0] = r0; taps.w[0] = wy0 * l0;
    taps.row[1] = r1; taps.w[1] = wy0 * l1;
    if (y0 + 1 <= cur - 1) {
        temb_resize_taps(y0 + 1, cur, E, r0, r1, l0, l1);
        taps.row[2] = r0; taps.w[2] = wy1 * l0;
        taps.row[3] = r1; taps.w[3] = wy1 * l1;
    } else {
        taps.row[2] = 0; taps.w[2] = 0.0f;
        taps.row[3] = 0; taps.w[3] = 0.0f;
    }
}

EMD_HD void temb_eval(const float* table /*[E][d]*/, int d, const TembTaps& taps, float* emb) {
    for (int j = 0; j < d; ++j) {
        float s = 0.0f;
        for (int k = 0; k < 4; ++k) s += taps.w[k] * table[taps.row[k] * d + j];
        emb[j] = s;
    }
}

EMD_HD void temb_vjp(float* v_table /*[E][d], accumulated*/, int d, const TembTaps& taps, const float* v_emb) {
    for (int k = 0; k < 4; ++k) {
        if (taps.w[k] == 0.0f) continue;
        for (int j = 0; j < d; ++j) v_table[taps.row[k] * d + j] += taps.w[k] * v_emb[j];
    }
}

// y[o] = W[o][:] . h + b[o]
EMD_HD void linear_fwd(const float* W, const float* b, int out, int in, const float* h, float* y) {
    for (int o = 0; o < out; ++o) {
        float s = b[o];
        for (int k = 0; k < in; ++k) s += W[o * in + k] * h[k];
        y[o] = s;
    }
}

// accumulates v_W, v_b (per-instance partial buffers) and v_h
EMD_HD void linear_vjp(const float* W, int out, int in, const float* h, const float* v_y, float* v_W, float* v_b,
                       float* v_h) {
    for (int o = 0; o < out; ++o) {
        v_b[o] += v_y[o];
        for (int k = 0; k < in; ++k) {
            v_W[o * in + k] += v_y[o] * h[k];
            v_h[k] += W[o * in + k] * v_y[o];
        }
    }
}

// Hamilton product (w first) and its VJPs
EMD_HD void qmul(const float* a, const float* b, float* o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
EMD_HD void qconj(const float* a, float* o) { o[0] = a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = -a[3]; }
// p = a (x) b :  v_a = v_p (x) conj(b),  v_b = conj(a) (x) v_p
EMD_HD void qmul_vjp(const float* a, const float* b, const float* v_p, float* v_a, float* v_b) {
    float t[4];
    if (v_a) { qconj(b, t); qmul(v_p, t, v_a); }
    if (v_b) { qconj(a, t); qmul(t, v_p, v_b); }
}
EMD_HD float qnormalize(const float* q, float* qn) {
    const float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) qn[k] = q[k] * inv;
    return inv;
}
EMD_HD void qnormalize_vjp(const float* qn, float inv, const float* v_qn, float* v_q) {
    const float d = v_qn[0] * qn[0] + v_qn[1] * qn[1] + v_qn[2] * qn[2] + v_qn[3] * qn[3];
    for (int k = 0; k < 4; ++k) v_q[k] = (v_qn[k] - d * qn[k]) * inv;
}
// unit quaternion -> row-major rotation (basics.py:30-49)
EMD_HD void qrot(const float* q, float* R) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
EMD_HD void qrot_vjp(const float* q, const float* vR, float* vq) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    vq[0] = 2.0f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    vq[1] = 2.0f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.0f * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] - 2.0f * x * vR[8]);
    vq[2] = 2.0f * (-2.0f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] - 2.0f * y * vR[8]);
    vq[3] = 2.0f * (-2.0f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.0f * z * vR[4] + y * vR[5] + x * vR[6] + y * vR[7]);
}

EMD_HD bool any_nan(const float* v, int n) {
    for (int k = 0; k < n; ++k)
        if (v[k] != v[k]) return true;
    return false;
}

// --------------------------------------------------------------------------
// Rigid node: per-instance pose with EMD offsets (rigid.py:203-246, 499-530, 547-566)
// --------------------------------------------------------------------------
struct RigidHeads {        // track_{rot,trans}_{c,f}: nn.Linear(d+g, 1|3)
    const float *rot_c_w, *rot_c_b, *rot_f_w, *rot_f_b;
    const float *trans_c_w, *trans_c_b, *trans_f_w, *trans_f_b;
};
// parameter-gradient layout of one instance's partial (floats): see rigid_param_count()
EMD_HD int rigid_param_count(int in) { return 2 * (in + 1) + 2 * (3 * in + 3); }

struct RigidInstOut {
    float R[9];   // rotmat(normalize(pose quat used for means))
    float t[3];   // pose trans + EMD translation offset
    float Q[4];   // normalize(pose quat (x) EMD yaw-offset quat)
};

// mean_emb: per-instance mean of the Gaussian embeddings (NaN when the instance owns no point)
EMD_HD void rigid_instance_fwd(const float* table, int E, int d, int g, const float* mean_emb, float t, int cur_c,
                               int cur_f, const RigidHeads& H, const float* pose_q_means, const float* pose_q_quats,
                               const float* pose_t, RigidInstOut& o) {
    float hc[EMD_TDIM_MAX + EMD_GDIM_MAX], hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    TembTaps tc, tf;
    temb_taps(t, cur_c, E, tc);
    temb_taps(t, cur_f, E, tf);
    temb_eval(table, d, tc, hc);
    temb_eval(table, d, tf, hf);
    for (int k = 0; k < g; ++k) { hc[d + k] = mean_emb[k]; hf[d + k] = mean_emb[k]; }
    const int in = d + g;
    float dtc[3], dtf[3], ac, af;
    linear_fwd(H.trans_c_w, H.trans_c_b, 3, in, hc, dtc);
    linear_fwd(H.trans_f_w, H.trans_f_b, 3, in, hf, dtf);
    linear_fwd(H.rot_c_w, H.rot_c_b, 1, in, hc, &ac);
    linear_fwd(H.rot_f_w, H.rot_f_b, 1, in, hf, &af);
    float dt[3] = {dtc[0] + dtf[0], dtc[1] + dtf[1], dtc[2] + dtf[2]};
    const float qc[4] = {cosf(ac), 0.f, 0.f, sinf(ac)}, qf[4] = {cosf(af), 0.f, 0.f, sinf(af)};
    float qoff[4];
    qmul(qc, qf, qoff);
    // means path: un-offset pose rotation, offset translation (skipped on NaN: rigid.py:528)
    float qn[4];
    qnormalize(pose_q_means, qn);
    qrot(qn, o.R);
    const bool skip_t = any_nan(dt, 3);
    for (int k = 0; k < 3; ++k) o.t[k] = pose_t[k] + (skip_t ? 0.0f : dt[k]);
    // quats path (skipped on NaN: rigid.py:559)
    const bool skip_q = any_nan(qoff, 4);
    float Qg[4];
    if (skip_q) { for (int k = 0; k < 4; ++k) Qg[k] = pose_q_quats[k]; }
    else qmul(pose_q_quats, qoff, Qg);
    qnormalize(Qg, o.Q);
}

// v_params: this instance's partial [rigid_param_count(in)] in the order
//   rot_c_w[in] rot_c_b[1] rot_f_w[in] rot_f_b[1] trans_c_w[3*in] trans_c_b[3] trans_f_w[3*in] trans_f_b[3]
// v_table: this instance's [E][d] slab (accumulated; caller zeroes it)
EMD_HD void rigid_instance_bwd(const float* table, int E, int d, int g, const float* mean_emb, float t, int cur_c,
                               int cur_f, const RigidHeads& H, const float* pose_q_means, const float* pose_q_quats,
                               const float* pose_t, const float* v_R, const float* v_t, const float* v_Q,
                               float* v_pose_q_means, float* v_pose_q_quats, float* v_pose_t, float* v_params,
                               float* v_table, float* v_mean_emb) {
    const int in = d + g;
    float hc[EMD_TDIM_MAX + EMD_GDIM_MAX], hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    TembTaps tc, tf;
    temb_taps(t, cur_c, E, tc);
    temb_taps(t, cur_f, E, tf);
    temb_eval(table, d, tc, hc);
    temb_eval(table, d, tf, hf);
    for (int k = 0; k < g; ++k) { hc[d + k] = mean_emb[k]; hf[d + k] = mean_emb[k]; }
    float dtc[3], dtf[3], ac, af;
    linear_fwd(H.trans_c_w, H.trans_c_b, 3, in, hc, dtc);
    linear_fwd(H.trans_f_w, H.trans_f_b, 3, in, hf, dtf);
    linear_fwd(H.rot_c_w, H.rot_c_b, 1, in, hc, &ac);
    linear_fwd(H.rot_f_w, H.rot_f_b, 1, in, hf, &af);
    const float dt[3] = {dtc[0] + dtf[0], dtc[1] + dtf[1], dtc[2] + dtf[2]};
    const float qc[4] = {cosf(ac), 0.f, 0.f, sinf(ac)}, qf[4] = {cosf(af), 0.f, 0.f, sinf(af)};
    float qoff[4];
    qmul(qc, qf, qoff);
    const bool skip_t = any_nan(dt, 3), skip_q = any_nan(qoff, 4);

    float* p = v_params;
    float* v_rot_c_w = p; p += in; float* v_rot_c_b = p; p += 1;
    float* v_rot_f_w = p; p += in; float* v_rot_f_b = p; p += 1;
    float* v_trans_c_w = p; p += 3 * in; float* v_trans_c_b = p; p += 3;
    float* v_trans_f_w = p; p += 3 * in; float* v_trans_f_b = p;
    for (int k = 0; k < rigid_param_count(in); ++k) v_params[k] = 0.0f;
    for (int k = 0; k < g; ++k) v_mean_emb[k] = 0.0f;

    // means path
    float qn[4], vqn[4];
    const float inv = qnormalize(pose_q_means, qn);
    qrot_vjp(qn, v_R, vqn);
    qnormalize_vjp(qn, inv, vqn, v_pose_q_means);
    for (int k = 0; k < 3; ++k) v_pose_t[k] = v_t[k];
    float v_hc[EMD_TDIM_MAX + EMD_GDIM_MAX], v_hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    for (int k = 0; k < in; ++k) { v_hc[k] = 0.0f; v_hf[k] = 0.0f; }
    if (!skip_t) {
        linear_vjp(H.trans_c_w, 3, in, hc, v_t, v_trans_c_w, v_trans_c_b, v_hc);
        linear_vjp(H.trans_f_w, 3, in, hf, v_t, v_trans_f_w, v_trans_f_b, v_hf);
    }
    // quats path
    float Qg[4], Qn[4], v_Qg[4];
    if (skip_q) { for (int k = 0; k < 4; ++k) Qg[k] = pose_q_quats[k]; }
    else qmul(pose_q_quats, qoff, Qg);
    const float invQ = qnormalize(Qg, Qn);
    qnormalize_vjp(Qn, invQ, v_Q, v_Qg);
    if (skip_q) {
        for (int k = 0; k < 4; ++k) v_pose_q_quats[k] = v_Qg[k];
    } else {
        float v_qoff[4];
        qmul_vjp(pose_q_quats, qoff, v_Qg, v_pose_q_quats, v_qoff);
        // qoff = qc (x) qf = (cos(ac+af), 0, 0, sin(ac+af)):  d/d ac = d/d af = (-qoff.z, 0, 0, qoff.w)
        const float v_a = -qoff[3] * v_qoff[0] + qoff[0] * v_qoff[3];
        linear_vjp(H.rot_c_w, 1, in, hc, &v_a, v_rot_c_w, v_rot_c_b, v_hc);
        linear_vjp(H.rot_f_w, 1, in, hf, &v_a, v_rot_f_w, v_rot_f_b, v_hf);
    }
    temb_vjp(v_table, d, tc, v_hc);
    temb_vjp(v_table, d, tf, v_hf);
    for (int k = 0; k < g; ++k) v_mean_emb[k] = v_hc[d + k] + v_hf[d + k];
}

// --------------------------------------------------------------------------
// SMPL node: per-instance joint transforms with EMD joint-yaw offsets
// (smpl.py:401-436, 459-489; human_body.py:158-172; smplx batch_rigid_transform)
// --------------------------------------------------------------------------
constexpr int SMPL_J = 24;
EMD_HD int smpl_param_count(int in) { return 2 * (SMPL_J * in + SMPL_J); }

struct SmplHeads { const float *c_w, *c_b, *f_w, *f_b; };  // track_smpl_{c,f}: nn.Linear(d+g, 24)

// 3x4 helpers: T = [R(9) | t(3)], stored R row-major then t
EMD_HD void mat3_mul(const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
EMD_HD void mat3_vec(const float* A, const float* v, float* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// theta: [24][4] pose quaternions (global orient + 23 joints); J: [24][3]; A0inv: [24][16] row-major 4x4;
// parents: [24]; out A: [24][12] (R row-major, then t).  Saves offsets qoff[24][4] for nothing: recomputed in bwd.
EMD_HD void smpl_instance_fwd(const float* table, int E, int d, int g, const float* mean_emb, float t, int cur_c,
                              int cur_f, const SmplHeads& H, const float* theta, const float* J, const float* A0inv,
                              const int* parents, float* A /*[24][12]*/) {
    float hc[EMD_TDIM_MAX + EMD_GDIM_MAX], hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    TembTaps tc, tf;
    temb_taps(t, cur_c, E, tc);
    temb_taps(t, cur_f, E, tf);
    temb_eval(table, d, tc, hc);
    temb_eval(table, d, tf, hf);
    for (int k = 0; k < g; ++k) { hc[d + k] = mean_emb[k]; hf[d + k] = mean_emb[k]; }
    const int in = d + g;
    float ac[SMPL_J], af[SMPL_J];
    linear_fwd(H.c_w, H.c_b, SMPL_J, in, hc, ac);
    linear_fwd(H.f_w, H.f_b, SMPL_J, in, hf, af);
    const bool skip = any_nan(ac, SMPL_J) || any_nan(af, SMPL_J);
    float G[SMPL_J][12];
    for (int j = 0; j < SMPL_J; ++j) {
        float th[4];
        if (skip) { for (int k = 0; k < 4; ++k) th[k] = theta[j * 4 + k]; }
        else {
            const float qc[4] = {cosf(ac[j]), 0.f, 0.f, sinf(ac[j])}, qf[4] = {cosf(af[j]), 0.f, 0.f, sinf(af[j])};
            float qo[4];
            qmul(qc, qf, qo);
            qmul(theta + j * 4, qo, th);
        }
        float thn[4], R[9];
        qnormalize(th, thn);
        qrot(thn, R);
        const int p = parents[j];
        if (p < 0) {
            for (int k = 0; k < 9; ++k) G[j][k] = R[k];
            for (int k = 0; k < 3; ++k) G[j][9 + k] = J[j * 3 + k];
        } else {
            const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
            mat3_mul(G[p], R, G[j]);
            float tr[3];
            mat3_vec(G[p], rel, tr);
            for (int k = 0; k < 3; ++k) G[j][9 + k] = tr[k] + G[p][9 + k];
        }
        // A' = [G.R | G.t - G.R J];  A = A' * A0inv
        float RJ[3];
        mat3_vec(G[j], J + j * 3, RJ);
        const float tp[3] = {G[j][9] - RJ[0], G[j][10] - RJ[1], G[j][11] - RJ[2]};
        const float* I4 = A0inv + j * 16;
        const float Ri[9] = {I4[0], I4[1], I4[2], I4[4], I4[5], I4[6], I4[8], I4[9], I4[10]};
        const float ti[3] = {I4[3], I4[7], I4[11]};
        mat3_mul(G[j], Ri, A + j * 12);
        float rt[3];
        mat3_vec(G[j], ti, rt);
        for (int k = 0; k < 3; ++k) A[j * 12 + 9 + k] = rt[k] + tp[k];
    }
}

// v_A: [24][12].  Outputs: v_theta[24][4], v_params (this instance's partial, zeroed here),
// v_table (accumulated), v_mean_emb[g].
EMD_HD void smpl_instance_bwd(const float* table, int E, int d, int g, const float* mean_emb, float t, int cur_c,
                              int cur_f, const SmplHeads& H, const float* theta, const float* J, const float* A0inv,
                              const int* parents, const float* v_A, float* v_theta, float* v_params, float* v_table,
                              float* v_mean_emb) {
    const int in = d + g;
    float hc[EMD_TDIM_MAX + EMD_GDIM_MAX], hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    TembTaps tc, tf;
    temb_taps(t, cur_c, E, tc);
    temb_taps(t, cur_f, E, tf);
    temb_eval(table, d, tc, hc);
    temb_eval(table, d, tf, hf);
    for (int k = 0; k < g; ++k) { hc[d + k] = mean_emb[k]; hf[d + k] = mean_emb[k]; }
    float ac[SMPL_J], af[SMPL_J];
    linear_fwd(H.c_w, H.c_b, SMPL_J, in, hc, ac);
    linear_fwd(H.f_w, H.f_b, SMPL_J, in, hf, af);
    const bool skip = any_nan(ac, SMPL_J) || any_nan(af, SMPL_J);
    // forward replay
    float G[SMPL_J][12], Rj[SMPL_J][9], thn[SMPL_J][4], thinv[SMPL_J], qo[SMPL_J][4];
    for (int j = 0; j < SMPL_J; ++j) {
        float th[4];
        if (skip) { for (int k = 0; k < 4; ++k) th[k] = theta[j * 4 + k]; }
        else {
            const float qc[4] = {cosf(ac[j]), 0.f, 0.f, sinf(ac[j])}, qf[4] = {cosf(af[j]), 0.f, 0.f, sinf(af[j])};
            qmul(qc, qf, qo[j]);
            qmul(theta + j * 4, qo[j], th);
        }
        thinv[j] = qnormalize(th, thn[j]);
        qrot(thn[j], Rj[j]);
        const int p = parents[j];
        if (p < 0) {
            for (int k = 0; k < 9; ++k) G[j][k] = Rj[j][k];
            for (int k = 0; k < 3; ++k) G[j][9 + k] = J[j * 3 + k];
        } else {
            const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
            mat3_mul(G[p], Rj[j], G[j]);
            float tr[3];
            mat3_vec(G[p], rel, tr);
            for (int k = 0; k < 3; ++k) G[j][9 + k] = tr[k] + G[p][9 + k];
        }
    }
    // v_G from v_A
    float vG[SMPL_J][12];
    for (int j = 0; j < SMPL_J; ++j) {
        const float* I4 = A0inv + j * 16;
        const float Ri[9] = {I4[0], I4[1], I4[2], I4[4], I4[5], I4[6], I4[8], I4[9], I4[10]};
        const float ti[3] = {I4[3], I4[7], I4[11]};
        const float* vAR = v_A + j * 12;
        const float* vAt = v_A + j * 12 + 9;
        // A.R = G.R Ri ; A.t = G.R ti + (G.t - G.R J)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                float s = vAR[r * 3] * Ri[c * 3] + vAR[r * 3 + 1] * Ri[c * 3 + 1] + vAR[r * 3 + 2] * Ri[c * 3 + 2];  // vAR * Ri^T
                s += vAt[r] * (ti[c] - J[j * 3 + c]);
                vG[j][r * 3 + c] = s;
            }
        for (int k = 0; k < 3; ++k) vG[j][9 + k] = vAt[k];
    }
    float* p_ = v_params;
    float* v_c_w = p_; p_ += SMPL_J * in; float* v_c_b = p_; p_ += SMPL_J;
    float* v_f_w = p_; p_ += SMPL_J * in; float* v_f_b = p_;
    for (int k = 0; k < smpl_param_count(in); ++k) v_params[k] = 0.0f;
    float v_ac[SMPL_J];
    for (int j = SMPL_J - 1; j >= 0; --j) {
        const int p = parents[j];
        float vR[9];
        if (p < 0) {
            for (int k = 0; k < 9; ++k) vR[k] = vG[j][k];
        } else {
            const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
            // G_j.R = G_p.R R_j ; G_j.t = G_p.R rel + G_p.t
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    // v_Gp.R += vG_j.R R_j^T + vG_j.t (x) rel
                    vG[p][r * 3 + c] += vG[j][r * 3] * Rj[j][c * 3] + vG[j][r * 3 + 1] * Rj[j][c * 3 + 1] +
                                        vG[j][r * 3 + 2] * Rj[j][c * 3 + 2] + vG[j][9 + r] * rel[c];
                    // v_R_j = G_p.R^T vG_j.R
                    vR[r * 3 + c] = G[p][0 * 3 + r] * vG[j][0 * 3 + c] + G[p][1 * 3 + r] * vG[j][1 * 3 + c] +
                                    G[p][2 * 3 + r] * vG[j][2 * 3 + c];
                }
            for (int k = 0; k < 3; ++k) vG[p][9 + k] += vG[j][9 + k];
        }
        float vthn[4], vth[4];
        qrot_vjp(thn[j], vR, vthn);
        qnormalize_vjp(thn[j], thinv[j], vthn, vth);
        if (skip) {
            for (int k = 0; k < 4; ++k) v_theta[j * 4 + k] = vth[k];
            v_ac[j] = 0.0f;
        } else {
            float vqo[4];
            qmul_vjp(theta + j * 4, qo[j], vth, v_theta + j * 4, vqo);
            v_ac[j] = -qo[j][3] * vqo[0] + qo[j][0] * vqo[3];
        }
    }
    float v_hc[EMD_TDIM_MAX + EMD_GDIM_MAX], v_hf[EMD_TDIM_MAX + EMD_GDIM_MAX];
    for (int k = 0; k < in; ++k) { v_hc[k] = 0.0f; v_hf[k] = 0.0f; }
    if (!skip) {
        linear_vjp(H.c_w, SMPL_J, in, hc, v_ac, v_c_w, v_c_b, v_hc);
        linear_vjp(H.f_w, SMPL_J, in, hf, v_ac, v_f_w, v_f_b, v_hf);
    }
    temb_vjp(v_table, d, tc, v_hc);
    temb_vjp(v_table, d, tf, v_hf);
    for (int k = 0; k < g; ++k) v_mean_emb[k] = v_hc[d + k] + v_hf[d + k];
}

// pytorch3d.transforms.matrix_to_quaternion (+ w >= 0 standardisation), row-major m[9] -> q[4]; returns branch | sign<<2
EMD_HD int mat_to_quat(const float* m, float* q) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    const float s[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    float qa[4];
    int best = 0;
    for (int k = 0; k < 4; ++k) { qa[k] = s[k] > 0.0f ? sqrtf(s[k]) : 0.0f; if (qa[k] > qa[best]) best = k; }
    float num[4];
    if (best == 0) { num[0] = qa[0] * qa[0]; num[1] = m21 - m12; num[2] = m02 - m20; num[3] = m10 - m01; }
    else if (best == 1) { num[0] = m21 - m12; num[1] = qa[1] * qa[1]; num[2] = m10 + m01; num[3] = m02 + m20; }
    else if (best == 2) { num[0] = m02 - m20; num[1] = m10 + m01; num[2] = qa[2] * qa[2]; num[3] = m12 + m21; }
    else { num[0] = m10 - m01; num[1] = m20 + m02; num[2] = m21 + m12; num[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
    int flip = num[0] / den < 0.0f ? 1 : 0;
    for (int k = 0; k < 4; ++k) q[k] = (flip ? -num[k] : num[k]) / den;
    return best | (flip << 2);
}

EMD_HD void mat_to_quat_vjp(const float* m, const float* v_q_in, float* v_m) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    const float s[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    float qa[4];
    int best = 0;
    for (int k = 0; k < 4; ++k) { qa[k] = s[k] > 0.0f ? sqrtf(s[k]) : 0.0f; if (qa[k] > qa[best]) best = k; }
    float num[4];
    if (best == 0) { num[0] = qa[0] * qa[0]; num[1] = m21 - m12; num[2] = m02 - m20; num[3] = m10 - m01; }
    else if (best == 1) { num[0] = m21 - m12; num[1] = qa[1] * qa[1]; num[2] = m10 + m01; num[3] = m02 + m20; }
    else if (best == 2) { num[0] = m02 - m20; num[1] = m10 + m01; num[2] = qa[2] * qa[2]; num[3] = m12 + m21; }
    else { num[0] = m10 - m01; num[1] = m20 + m02; num[2] = m21 + m12; num[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
    const bool flip = num[0] / den < 0.0f;
    float v_q[4];
    for (int k = 0; k < 4; ++k) v_q[k] = flip ? -v_q_in[k] : v_q_in[k];
    float v_num[4], v_den = 0.0f;
    for (int k = 0; k < 4; ++k) { v_num[k] = v_q[k] / den; v_den -= v_q[k] * num[k] / (den * den); }
    // den = 2 max(qa, .1); num[best] = qa^2 ; qa = sqrt(s_best)
    float v_qa = (qa[best] > 0.1f ? 2.0f * v_den : 0.0f) + 2.0f * qa[best] * v_num[best];
    const float v_s = qa[best] > 0.0f ? v_qa / (2.0f * qa[best]) : 0.0f;
    for (int k = 0; k < 9; ++k) v_m[k] = 0.0f;
    const float sg[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    v_m[0] += sg[best][0] * v_s; v_m[4] += sg[best][1] * v_s; v_m[8] += sg[best][2] * v_s;
    // off-diagonal combos: (index in m, sign) per numerator slot
    if (best == 0) { v_m[7] += v_num[1]; v_m[5] -= v_num[1]; v_m[2] += v_num[2]; v_m[6] -= v_num[2]; v_m[3] += v_num[3]; v_m[1] -= v_num[3]; }
    else if (best == 1) { v_m[7] += v_num[0]; v_m[5] -= v_num[0]; v_m[3] += v_num[2]; v_m[1] += v_num[2]; v_m[2] += v_num[3]; v_m[6] += v_num[3]; }
    else if (best == 2) { v_m[2] += v_num[0]; v_m[6] -= v_num[0]; v_m[3] += v_num[1]; v_m[1] += v_num[1]; v_m[5] += v_num[3]; v_m[7] += v_num[3]; }
    else { v_m[3] += v_num[0]; v_m[1] -= v_num[0]; v_m[6] += v_num[1]; v_m[2] += v_num[1]; v_m[7] += v_num[2]; v_m[5] += v_num[2]; }
}

// d loss / d W[n, :] of one skinned point (needed when the LBS weights come from the trainable voxel deformer,
// human_body.py:174-179): T = sum_j W_j A_j, x' = T.R x + T.t, q' = normalize(mat2quat(T.R)) (x) normalize(q).
// v_T is formed exactly as smpl_points_bwd_kernel forms it; v_W[j] = <v_T, A_j>.  A_j = [R row-major (9) | t (3)].
EMD_HD void smpl_point_weight_grad(const float* Wn, const float* Ab, const float* x, const float* q, const float* g,
                                   const float* vg, float* v_W) {
    float T[12];
    for (int k = 0; k < 12; ++k) T[k] = 0.f;
    for (int j = 0; j < SMPL_J; ++j) {
        const float w = Wn[j];
        if (w == 0.f) continue;
        for (int k = 0; k < 12; ++k) T[k] += w * Ab[j * 12 + k];
    }
    float vT[12];
    vT[0] = g[0] * x[0]; vT[1] = g[0] * x[1]; vT[2] = g[0] * x[2];
    vT[3] = g[1] * x[0]; vT[4] = g[1] * x[1]; vT[5] = g[1] * x[2];
    vT[6] = g[2] * x[0]; vT[7] = g[2] * x[1]; vT[8] = g[2] * x[2];
    vT[9] = g[0]; vT[10] = g[1]; vT[11] = g[2];
    float qR[4], qRn[4], qn[4], v_qRn[4], v_qR[4], v_m[9];
    mat_to_quat(T, qR);
    const float invR = qnormalize(qR, qRn);
    qnormalize(q, qn);
    qmul_vjp(qRn, qn, vg, v_qRn, nullptr);
    qnormalize_vjp(qRn, invR, v_qRn, v_qR);
    mat_to_quat_vjp(T, v_qR, v_m);
    for (int k = 0; k < 9; ++k) vT[k] += v_m[k];
    for (int j = 0; j < SMPL_J; ++j) {
        float s = 0.f;
        for (int k = 0; k < 12; ++k) s += vT[k] * Ab[j * 12 + k];
        v_W[j] = s;
    }
}
