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
