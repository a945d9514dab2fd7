// Per-Gaussian projection math (gsplat-1.x pinhole flavour), forward and VJP.
//
// Replaces gsplat's fully_fused_projection as reached from
// OmniRe/models/trainers/base.py:393 (reference tree; the CUDA source itself is
// an un-vendored dependency).
//
// CANONICAL OP ORDER.  Every quantity that feeds an integer artefact (radius,
// tile rectangle, depth bits of the sort key) is evaluated as a fixed sequence
// of IEEE-754 round-to-nearest fp32 operations with NO fused multiply-add:
// the c_mul/c_add/c_sub/c_rcp/c_sqrt wrappers map to __f*_rn intrinsics on the
// device (ptxas never contracts those) and to plain operators on the host
// (tests compile this header with -ffp-contract=off).  DESIGN.md section 4 lists
// the sequence; the CPU oracle evaluates the same sequence with torch
// elementwise ops, which is what makes radii / tile ids / keys bit-exact.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define EMD_HD __host__ __device__ __forceinline__
#else
#define EMD_HD static inline
#endif

#ifdef __CUDA_ARCH__
EMD_HD float c_mul(float a, float b) { return __fmul_rn(a, b); }
EMD_HD float c_add(float a, float b) { return __fadd_rn(a, b); }
EMD_HD float c_sub(float a, float b) { return __fsub_rn(a, b); }
EMD_HD float c_rcp(float a) { return __frcp_rn(a); }
EMD_HD float c_div(float a, float b) { return __fdiv_rn(a, b); }
EMD_HD float c_sqrt(float a) { return __fsqrt_rn(a); }
#else
EMD_HD float c_mul(float a, float b) { return a * b; }
EMD_HD float c_add(float a, float b) { return a + b; }
EMD_HD float c_sub(float a, float b) { return a - b; }
EMD_HD float c_rcp(float a) { return 1.0f / a; }
EMD_HD float c_div(float a, float b) { return a / b; }
EMD_HD float c_sqrt(float a) { return sqrtf(a); }
#endif

EMD_HD float c_dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return c_add(c_add(c_mul(a0, b0), c_mul(a1, b1)), c_mul(a2, b2));
}

// Per-camera constants (computed once per block, canonical order).
struct CamConst {
    float V[12];  // rows 0..2 of the row-major world->camera matrix
    float fx, fy, cx, cy;
    float lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
};

EMD_HD void make_cam_const(const float* viewmat, const float* K, int width, int height, CamConst& c) {
    for (int i = 0; i < 12; ++i) c.V[i] = viewmat[i];
    c.fx = K[0]; c.cx = K[2]; c.fy = K[4]; c.cy = K[5];
    const float Wf = (float)width, Hf = (float)height;
    const float tanx = c_div(c_mul(0.5f, Wf), c.fx);
    const float tany = c_div(c_mul(0.5f, Hf), c.fy);
    c.lim_x_pos = c_add(c_div(c_sub(Wf, c.cx), c.fx), c_mul(0.3f, tanx));
    c.lim_x_neg = c_add(c_div(c.cx, c.fx), c_mul(0.3f, tanx));
    c.lim_y_pos = c_add(c_div(c_sub(Hf, c.cy), c.fy), c_mul(0.3f, tany));
    c.lim_y_neg = c_add(c_div(c.cy, c.fy), c_mul(0.3f, tany));
}

// normalised quaternion -> rotation (row-major R[9]); returns 1/|q|
EMD_HD float quat_to_rotmat_c(const float q[4], float R[9], float qn[4]) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    const float n2 = c_add(c_add(c_add(c_mul(w, w), c_mul(x, x)), c_mul(y, y)), c_mul(z, z));
    const float inv = c_div(1.0f, c_sqrt(n2));
    w = c_mul(w, inv); x = c_mul(x, inv); y = c_mul(y, inv); z = c_mul(z, inv);
    qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
    const float x2 = c_mul(x, x), y2 = c_mul(y, y), z2 = c_mul(z, z);
    const float xy = c_mul(x, y), xz = c_mul(x, z), yz = c_mul(y, z);
    const float wx = c_mul(w, x), wy = c_mul(w, y), wz = c_mul(w, z);
    R[0] = c_sub(1.0f, c_mul(2.0f, c_add(y2, z2)));
    R[1] = c_mul(2.0f, c_sub(xy, wz));
    R[2] = c_mul(2.0f, c_add(xz, wy));
    R[3] = c_mul(2.0f, c_add(xy, wz));
    R[4] = c_sub(1.0f, c_mul(2.0f, c_add(x2, z2)));
    R[5] = c_mul(2.0f, c_sub(yz, wx));
    R[6] = c_mul(2.0f, c_sub(xz, wy));
    R[7] = c_mul(2.0f, c_add(yz, wx));
    R[8] = c_sub(1.0f, c_mul(2.0f, c_add(x2, y2)));
    return inv;
}

// world covariance S (6 unique: 00 01 02 11 12 22) from R and scale; also M = R diag(s)
EMD_HD void covar_world_c(const float R[9], const float s[3], float M[9], float S[6]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = c_mul(R[i * 3 + j], s[j]);
    S[0] = c_dot3(M[0], M[0], M[1], M[1], M[2], M[2]);
    S[1] = c_dot3(M[0], M[3], M[1], M[4], M[2], M[5]);
    S[2] = c_dot3(M[0], M[6], M[1], M[7], M[2], M[8]);
    S[3] = c_dot3(M[3], M[3], M[4], M[4], M[5], M[5]);
    S[4] = c_dot3(M[3], M[6], M[4], M[7], M[5], M[8]);
    S[5] = c_dot3(M[6], M[6], M[7], M[7], M[8], M[8]);
}

struct ProjFwd {
    float m2x, m2y, z;
    float conic_a, conic_b, conic_c;
    float comp;
    int radius;  // 0 => culled
    // intermediates the VJP re-derives from (kept in registers only)
    float x, y, rz, tx, ty, J00, J02, J11, J12;
    float Sc[6];
    float c00, c01, c11;  // blurred 2-D covariance
};

EMD_HD void project_gaussian_c(const float m[3], const float S[6], const CamConst& cam, int width, int height,
                               float eps2d, float near_plane, float far_plane, float radius_clip, ProjFwd& o) {
    const float* V = cam.V;
    o.radius = 0;
    const float x = c_add(c_add(c_add(c_mul(V[0], m[0]), c_mul(V[1], m[1])), c_mul(V[2], m[2])), V[3]);
    const float y = c_add(c_add(c_add(c_mul(V[4], m[0]), c_mul(V[5], m[1])), c_mul(V[6], m[2])), V[7]);
    const float z = c_add(c_add(c_add(c_mul(V[8], m[0]), c_mul(V[9], m[1])), c_mul(V[10], m[2])), V[11]);
    o.x = x; o.y = y; o.z = z;
    if (!(z >= near_plane && z <= far_plane)) return;
    // T = W S, Sc = T W^T   (W = V[:3,:3])
    const float Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    float T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            T[i * 3 + j] = c_dot3(V[i * 4 + 0], Sf[0 * 3 + j], V[i * 4 + 1], Sf[1 * 3 + j], V[i * 4 + 2], Sf[2 * 3 + j]);
#define EMD_SC(i, j) c_dot3(T[i * 3 + 0], V[j * 4 + 0], T[i * 3 + 1], V[j * 4 + 1], T[i * 3 + 2], V[j * 4 + 2])
    const float Sc00 = EMD_SC(0, 0), Sc01 = EMD_SC(0, 1), Sc02 = EMD_SC(0, 2);
    const float Sc11 = EMD_SC(1, 1), Sc12 = EMD_SC(1, 2), Sc22 = EMD_SC(2, 2);
#undef EMD_SC
    o.Sc[0] = Sc00; o.Sc[1] = Sc01; o.Sc[2] = Sc02; o.Sc[3] = Sc11; o.Sc[4] = Sc12; o.Sc[5] = Sc22;

    const float rz = c_rcp(z);
    const float rz2 = c_mul(rz, rz);
    const float tx = c_mul(z, fminf(cam.lim_x_pos, fmaxf(-cam.lim_x_neg, c_mul(x, rz))));
    const float ty = c_mul(z, fminf(cam.lim_y_pos, fmaxf(-cam.lim_y_neg, c_mul(y, rz))));
    const float J00 = c_mul(cam.fx, rz);
    const float J02 = -c_mul(c_mul(cam.fx, tx), rz2);
    const float J11 = c_mul(cam.fy, rz);
    const float J12 = -c_mul(c_mul(cam.fy, ty), rz2);
    o.rz = rz; o.tx = tx; o.ty = ty; o.J00 = J00; o.J02 = J02; o.J11 = J11; o.J12 = J12;
    const float A0 = c_add(c_mul(J00, Sc00), c_mul(J02, Sc02));
    const float A1 = c_add(c_mul(J00, Sc01), c_mul(J02, Sc12));
    const float A2 = c_add(c_mul(J00, Sc02), c_mul(J02, Sc22));
    const float B1 = c_add(c_mul(J11, Sc11), c_mul(J12, Sc12));
    const float B2 = c_add(c_mul(J11, Sc12), c_mul(J12, Sc22));
    float c00 = c_add(c_mul(A0, J00), c_mul(A2, J02));
    const float c01 = c_add(c_mul(A1, J11), c_mul(A2, J12));
    float c11 = c_add(c_mul(B1, J11), c_mul(B2, J12));
    const float m2x = c_add(c_mul(c_mul(cam.fx, x), rz), cam.cx);
    const float m2y = c_add(c_mul(c_mul(cam.fy, y), rz), cam.cy);

    const float det_orig = c_sub(c_mul(c00, c11), c_mul(c01, c01));
    c00 = c_add(c00, eps2d);
    c11 = c_add(c11, eps2d);
    const float det = c_sub(c_mul(c00, c11), c_mul(c01, c01));
    if (!(det > 0.0f)) return;
    const float inv_det = c_rcp(det);
    const float b = c_mul(0.5f, c_add(c00, c11));
    const float v1 = c_add(b, c_sqrt(fmaxf(0.01f, c_sub(c_mul(b, b), det))));
    const float radius = ceilf(c_mul(3.0f, c_sqrt(v1)));
    if (!(radius > radius_clip)) return;
    if (!(radius <= 3.0e38f)) return;  // inf / NaN
    if (c_add(m2x, radius) <= 0.0f || c_sub(m2x, radius) >= (float)width || c_add(m2y, radius) <= 0.0f ||
        c_sub(m2y, radius) >= (float)height)
        return;
    o.m2x = m2x; o.m2y = m2y;
    o.conic_a = c_mul(c11, inv_det);
    o.conic_b = -c_mul(c01, inv_det);
    o.conic_c = c_mul(c00, inv_det);
    o.c00 = c00; o.c01 = c01; o.c11 = c11;
    o.comp = c_sqrt(fmaxf(0.0f, c_div(det_orig, det)));
    // radii beyond int32 cannot occur for finite on-screen means; clamp defensively
    o.radius = radius < 2.0e9f ? (int)radius : 2000000000;
}

// Tile rectangle [x0,x1) x [y0,y1) of a projected Gaussian (gsplat isect_tiles).
// Shared by the counting pass, the emitting pass and the backward's slot lookup,
// so all three agree bit for bit.
EMD_HD void tile_rect_c(float m2x, float m2y, int radius, int tile_w, int tile_h, int& x0, int& y0, int& x1, int& y1) {
    const float inv_ts = 1.0f / 16.0f;  // exact power of two: multiply == divide
    const float tr = c_mul((float)radius, inv_ts);
    const float tx = c_mul(m2x, inv_ts);
    const float ty = c_mul(m2y, inv_ts);
    const float fx0 = floorf(c_sub(tx, tr)), fy0 = floorf(c_sub(ty, tr));
    const float fx1 = ceilf(c_add(tx, tr)), fy1 = ceilf(c_add(ty, tr));
    x0 = (int)fminf(fmaxf(fx0, 0.0f), (float)tile_w);
    y0 = (int)fminf(fmaxf(fy0, 0.0f), (float)tile_h);
    x1 = (int)fminf(fmaxf(fx1, 0.0f), (float)tile_w);
    y1 = (int)fminf(fmaxf(fy1, 0.0f), (float)tile_h);
}

// ---------------------------------------------------------------------------
// VJP.  Not bit-critical (gradient tolerance is 1e-3 relative): plain arithmetic,
// the compiler is free to contract.
// ---------------------------------------------------------------------------
// Accumulates v_mean (world) and v_S (world covariance, 6 unique, "symmetric
// full-matrix" convention: off-diagonals hold the gradient of BOTH mirrored
// entries) for one camera.
EMD_HD void project_gaussian_vjp(const ProjFwd& f, const CamConst& cam, float v_m2x, float v_m2y, float v_z,
                                 float v_ca, float v_cb, float v_cc, float v_mean[3], float v_S[6]) {
    const float* V = cam.V;
    // conic = inverse(cov2d'):  v_cov = -C * Vm * C, Vm = [[va, vb/2],[vb/2, vc]]
    const float a = f.conic_a, b = f.conic_b, c = f.conic_c;
    const float hb = 0.5f * v_cb;
    // X = C * Vm
    const float X00 = a * v_ca + b * hb, X01 = a * hb + b * v_cc;
    const float X10 = b * v_ca + c * hb, X11 = b * hb + c * v_cc;
    // v_cov = -(X * C)
    const float g00 = -(X00 * a + X01 * b);
    const float g01 = -(X00 * b + X01 * c);
    const float g10 = -(X10 * a + X11 * b);
    const float g11 = -(X10 * b + X11 * c);
    const float gs01 = 0.5f * (g01 + g10);  // symmetrised
    // cov2d = J Sc J^T,  J = [[J00,0,J02],[0,J11,J12]]
    const float J00 = f.J00, J02 = f.J02, J11 = f.J11, J12 = f.J12;
    const float Sc00 = f.Sc[0], Sc01 = f.Sc[1], Sc02 = f.Sc[2], Sc11 = f.Sc[3], Sc12 = f.Sc[4], Sc22 = f.Sc[5];
    // v_Sc = J^T G J   (3x3 symmetric), G = [[g00, gs01],[gs01, g11]]
    // rows of (G J): r0 = [g00*J00, gs01*J11, g00*J02 + gs01*J12], r1 = [gs01*J00, g11*J11, gs01*J02 + g11*J12]
    const float r00 = g00 * J00, r01 = gs01 * J11, r02 = g00 * J02 + gs01 * J12;
    const float r10 = gs01 * J00, r11 = g11 * J11, r12 = gs01 * J02 + g11 * J12;
    float vSc[9];
    vSc[0] = J00 * r00;            vSc[1] = J00 * r01;            vSc[2] = J00 * r02;
    vSc[3] = J11 * r10;            vSc[4] = J11 * r11;            vSc[5] = J11 * r12;
    vSc[6] = J02 * r00 + J12 * r10; vSc[7] = J02 * r01 + J12 * r11; vSc[8] = J02 * r02 + J12 * r12;
    // v_J = 2 G J Sc  (2x3);  (G J) rows are r0, r1
    const float vJ00 = 2.0f * (r00 * Sc00 + r01 * Sc01 + r02 * Sc02);
    const float vJ02 = 2.0f * (r00 * Sc02 + r01 * Sc12 + r02 * Sc22);
    const float vJ11 = 2.0f * (r10 * Sc01 + r11 * Sc11 + r12 * Sc12);
    const float vJ12 = 2.0f * (r10 * Sc02 + r11 * Sc12 + r12 * Sc22);
    // camera-space mean gradient
    const float rz = f.rz, rz2 = rz * rz, rz3 = rz2 * rz;
    const float fx = cam.fx, fy = cam.fy;
    float vx = fx * rz * v_m2x;
    float vy = fy * rz * v_m2y;
    float vz = v_z - (fx * f.x * v_m2x + fy * f.y * v_m2y) * rz2;
    // J00 = fx/z ; J11 = fy/z
    vz += -fx * rz2 * vJ00 - fy * rz2 * vJ11;
    // J02 = -fx*tx/z^2 with tx = z*clamp(x/z)
    const float xr = f.x * rz, yr = f.y * rz;
    const bool x_in = (xr <= cam.lim_x_pos) && (xr >= -cam.lim_x_neg);
    const bool y_in = (yr <= cam.lim_y_pos) && (yr >= -cam.lim_y_neg);
    if (x_in) { vx += -fx * rz2 * vJ02; vz += 2.0f * fx * f.tx * rz3 * vJ02; }
    else      { vz += fx * f.tx * rz3 * vJ02; }
    if (y_in) { vy += -fy * rz2 * vJ12; vz += 2.0f * fy * f.ty * rz3 * vJ12; }
    else      { vz += fy * f.ty * rz3 * vJ12; }
    // world mean: m_c = W m + t
    v_mean[0] += V[0] * vx + V[4] * vy + V[8] * vz;
    v_mean[1] += V[1] * vx + V[5] * vy + V[9] * vz;
    v_mean[2] += V[2] * vx + V[6] * vy + V[10] * vz;
    // world covariance: Sc = W S W^T  =>  v_S = W^T v_Sc W
    float t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            t[i * 3 + j] = vSc[i * 3 + 0] * V[0 * 4 + j] + vSc[i * 3 + 1] * V[1 * 4 + j] + vSc[i * 3 + 2] * V[2 * 4 + j];
    float vS[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            vS[i * 3 + j] = V[0 * 4 + i] * t[0 * 3 + j] + V[1 * 4 + i] * t[1 * 3 + j] + V[2 * 4 + i] * t[2 * 3 + j];
    v_S[0] += vS[0];
    v_S[1] += vS[1] + vS[3];
    v_S[2] += vS[2] + vS[6];
    v_S[3] += vS[4];
    v_S[4] += vS[5] + vS[7];
    v_S[5] += vS[8];
}

// S = M M^T, M = R diag(s), R = rot(normalize(q)).  v_S uses the convention above.
EMD_HD void covar_world_vjp(const float qn[4], float inv_norm, const float R[9], const float M[9], const float s[3],
                            const float v_S[6], float v_q[4], float v_s[3]) {
    // full symmetric gradient matrix Gs with Gs_ij = dL/dS_ij treating S_ij and S_ji as one variable:
    // dL = sum_{i<=j} v_S[ij] dS_ij ;  dS = dM M^T + M dM^T  =>  v_M = (Gf + Gf^T) M with Gf upper-triangular.
    const float G[9] = {2.0f * v_S[0], v_S[1], v_S[2], v_S[1], 2.0f * v_S[3], v_S[4], v_S[2], v_S[4], 2.0f * v_S[5]};
    float vM[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) vM[i * 3 + j] = G[i * 3 + 0] * M[0 * 3 + j] + G[i * 3 + 1] * M[1 * 3 + j] + G[i * 3 + 2] * M[2 * 3 + j];
    float vR[9];
    for (int j = 0; j < 3; ++j) {
        v_s[j] = R[0 * 3 + j] * vM[0 * 3 + j] + R[1 * 3 + j] * vM[1 * 3 + j] + R[2 * 3 + j] * vM[2 * 3 + j];
        for (int i = 0; i < 3; ++i) vR[i * 3 + j] = vM[i * 3 + j] * s[j];
    }
    const float w = qn[0], x = qn[1], y = qn[2], z = qn[3];
    float vqn[4];
    vqn[0] = 2.0f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    vqn[1] = 2.0f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.0f * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] - 2.0f * x * vR[8]);
    vqn[2] = 2.0f * (-2.0f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] - 2.0f * y * vR[8]);
    vqn[3] = 2.0f * (-2.0f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.0f * z * vR[4] + y * vR[5] + x * vR[6] + y * vR[7]);
    const float d = vqn[0] * w + vqn[1] * x + vqn[2] * y + vqn[3] * z;
    v_q[0] = (vqn[0] - d * w) * inv_norm;
    v_q[1] = (vqn[1] - d * x) * inv_norm;
    v_q[2] = (vqn[2] - d * y) * inv_norm;
    v_q[3] = (vqn[3] - d * z) * inv_norm;
}
