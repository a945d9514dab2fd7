// Per-Gaussian preprocess math of the diff_gauss (Inria-derived) front end, forward and VJP.
//
// Replaces diff_gauss' preprocessCUDA as reached from
// S3Gaussian/gaussian_renderer/__init__.py:145 (the CUDA source is an un-vendored
// dependency; semantics restated in oracle/diff_gauss_ref.py).  Same canonical-op-order
// discipline as proj_math.cuh for everything that feeds radii / tile rects / depth keys.
// Matrices arrive in the reference's row-vector convention (S3Gaussian/scene/cameras.py:55-66):
// the flattened tensor holds the TRANSPOSE of the usual column-vector matrix.
#pragma once
#include "proj_math.cuh"

struct DgCam {
    float V[16];   // viewmatrix (transposed storage): view_i = sum_k V[k*4+i] p_k + V[12+i]
    float Pm[16];  // full projection, same storage
    float fx, fy, limx, limy;
    float mod;     // scale_modifier
    int W, H, tile_w, tile_h;
};

EMD_HD void make_dg_cam(const float* viewmatrix, const float* projmatrix, float tanfovx, float tanfovy, int W, int H,
                        float scale_modifier, DgCam& c) {
    for (int i = 0; i < 16; ++i) { c.V[i] = viewmatrix[i]; c.Pm[i] = projmatrix[i]; }
    c.fx = c_div((float)W, c_mul(2.0f, tanfovx));
    c.fy = c_div((float)H, c_mul(2.0f, tanfovy));
    c.limx = c_mul(1.3f, tanfovx);
    c.limy = c_mul(1.3f, tanfovy);
    c.mod = scale_modifier;
    c.W = W; c.H = H; c.tile_w = (W + 15) / 16; c.tile_h = (H + 15) / 16;
}

EMD_HD float dg_tp(const float* M, int col, const float p[3]) {
    return c_add(c_add(c_add(c_mul(M[0 * 4 + col], p[0]), c_mul(M[1 * 4 + col], p[1])), c_mul(M[2 * 4 + col], p[2])), M[3 * 4 + col]);
}

// rotation from the quaternion AS GIVEN (Inria computeCov3D does not renormalise)
EMD_HD void dg_quat_to_rotmat_c(const float q[4], float R[9]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = c_sub(1.0f, c_mul(2.0f, c_add(c_mul(y, y), c_mul(z, z))));
    R[1] = c_mul(2.0f, c_sub(c_mul(x, y), c_mul(r, z)));
    R[2] = c_mul(2.0f, c_add(c_mul(x, z), c_mul(r, y)));
    R[3] = c_mul(2.0f, c_add(c_mul(x, y), c_mul(r, z)));
    R[4] = c_sub(1.0f, c_mul(2.0f, c_add(c_mul(x, x), c_mul(z, z))));
    R[5] = c_mul(2.0f, c_sub(c_mul(y, z), c_mul(r, x)));
    R[6] = c_mul(2.0f, c_sub(c_mul(x, z), c_mul(r, y)));
    R[7] = c_mul(2.0f, c_add(c_mul(y, z), c_mul(r, x)));
    R[8] = c_sub(1.0f, c_mul(2.0f, c_add(c_mul(x, x), c_mul(y, y))));
}

// getRect: [x0,x1) x [y0,y1) with C-style truncation
EMD_HD void tile_rect_dg(float m2x, float m2y, int radius, int tile_w, int tile_h, int& x0, int& y0, int& x1, int& y1) {
    const float inv = 1.0f / 16.0f, r = (float)radius;
    const float ax0 = truncf(c_mul(c_sub(m2x, r), inv)), ay0 = truncf(c_mul(c_sub(m2y, r), inv));
    const float ax1 = truncf(c_mul(c_add(c_add(m2x, r), 15.0f), inv)), ay1 = truncf(c_mul(c_add(c_add(m2y, r), 15.0f), inv));
    x0 = (int)fminf(fmaxf(ax0, 0.0f), (float)tile_w);
    y0 = (int)fminf(fmaxf(ay0, 0.0f), (float)tile_h);
    x1 = (int)fminf(fmaxf(ax1, 0.0f), (float)tile_w);
    y1 = (int)fminf(fmaxf(ay1, 0.0f), (float)tile_h);
}

struct DgFwd {
    ProjFwd f;        // reuses the gsplat-flavour intermediates (x,y,z = view coords; J; Sc; conic; radius)
    float hx, hy, p_w;
    int x0, y0, x1, y1;
};

// S: world covariance (6 unique) from covar_world_c(R, mod*scale).
EMD_HD void dg_project_c(const float p[3], const float S[6], const DgCam& cam, DgFwd& o) {
    ProjFwd& f = o.f;
    f.radius = 0;
    const float tx = dg_tp(cam.V, 0, p), ty = dg_tp(cam.V, 1, p), tz = dg_tp(cam.V, 2, p);
    f.x = tx; f.y = ty; f.z = tz;
    if (!(tz > 0.2f)) return;
    o.hx = dg_tp(cam.Pm, 0, p); o.hy = dg_tp(cam.Pm, 1, p);
    const float hw = dg_tp(cam.Pm, 3, p);
    o.p_w = c_rcp(c_add(hw, 0.0000001f));
    const float ndc_x = c_mul(o.hx, o.p_w), ndc_y = c_mul(o.hy, o.p_w);
    // Wr[i][k] = V[k*4+i]
    const float* V = cam.V;
    const float Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    float T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            T[i * 3 + j] = c_dot3(V[0 * 4 + i], Sf[0 * 3 + j], V[1 * 4 + i], Sf[1 * 3 + j], V[2 * 4 + i], Sf[2 * 3 + j]);
#define EMD_SC(i, j) c_dot3(T[i * 3 + 0], V[0 * 4 + j], T[i * 3 + 1], V[1 * 4 + j], T[i * 3 + 2], V[2 * 4 + j])
    const float Sc00 = EMD_SC(0, 0), Sc01 = EMD_SC(0, 1), Sc02 = EMD_SC(0, 2);
    const float Sc11 = EMD_SC(1, 1), Sc12 = EMD_SC(1, 2), Sc22 = EMD_SC(2, 2);
#undef EMD_SC
    f.Sc[0] = Sc00; f.Sc[1] = Sc01; f.Sc[2] = Sc02; f.Sc[3] = Sc11; f.Sc[4] = Sc12; f.Sc[5] = Sc22;
    const float rz = c_rcp(tz), rz2 = c_mul(rz, rz);
    const float cx = c_mul(tz, fminf(cam.limx, fmaxf(-cam.limx, c_mul(tx, rz))));
    const float cy = c_mul(tz, fminf(cam.limy, fmaxf(-cam.limy, c_mul(ty, rz))));
    const float J00 = c_mul(cam.fx, rz), J02 = -c_mul(c_mul(cam.fx, cx), rz2);
    const float J11 = c_mul(cam.fy, rz), J12 = -c_mul(c_mul(cam.fy, cy), rz2);
    f.rz = rz; f.tx = cx; f.ty = cy; f.J00 = J00; f.J02 = J02; f.J11 = J11; f.J12 = J12;
    const float A0 = c_add(c_mul(J00, Sc00), c_mul(J02, Sc02));
    const float A1 = c_add(c_mul(J00, Sc01), c_mul(J02, Sc12));
    const float A2 = c_add(c_mul(J00, Sc02), c_mul(J02, Sc22));
    const float B1 = c_add(c_mul(J11, Sc11), c_mul(J12, Sc12));
    const float B2 = c_add(c_mul(J11, Sc12), c_mul(J12, Sc22));
    const float c00 = c_add(c_add(c_mul(A0, J00), c_mul(A2, J02)), 0.3f);
    const float c01 = c_add(c_mul(A1, J11), c_mul(A2, J12));
    const float c11 = c_add(c_add(c_mul(B1, J11), c_mul(B2, J12)), 0.3f);
    const float det = c_sub(c_mul(c00, c11), c_mul(c01, c01));
    if (det == 0.0f || !(det == det)) return;
    const float det_inv = c_rcp(det);
    const float mid = c_mul(0.5f, c_add(c00, c11));
    const float root = c_sqrt(fmaxf(0.1f, c_sub(c_mul(mid, mid), det)));
    const float lam = fmaxf(c_add(mid, root), c_sub(mid, root));
    const float radius = ceilf(c_mul(3.0f, c_sqrt(lam)));
    if (!(radius <= 3.0e38f)) return;
    const float m2x = c_mul(c_sub(c_mul(c_add(ndc_x, 1.0f), (float)cam.W), 1.0f), 0.5f);
    const float m2y = c_mul(c_sub(c_mul(c_add(ndc_y, 1.0f), (float)cam.H), 1.0f), 0.5f);
    const int ri = radius < 2.0e9f ? (int)radius : 2000000000;
    tile_rect_dg(m2x, m2y, ri, cam.tile_w, cam.tile_h, o.x0, o.y0, o.x1, o.y1);
    if ((o.x1 - o.x0) * (o.y1 - o.y0) <= 0) return;
    f.m2x = m2x; f.m2y = m2y;
    f.conic_a = c_mul(c11, det_inv); f.conic_b = -c_mul(c01, det_inv); f.conic_c = c_mul(c00, det_inv);
    f.c00 = c00; f.c01 = c01; f.c11 = c11;
    f.comp = 1.0f;
    f.radius = ri;
}

// v_m2d is the gradient w.r.t. the PIXEL mean.  Accumulates v_p (world mean) and v_S.
EMD_HD void dg_project_vjp(const DgFwd& o, const DgCam& cam, const float p[3], float v_m2x, float v_m2y, float v_z,
                           float v_ca, float v_cb, float v_cc, float v_p[3], float v_S[6]) {
    // covariance + depth path through the shared VJP: build a CamConst whose rows are Wr | view translation
    CamConst cc;
    for (int i = 0; i < 3; ++i) {
        for (int k = 0; k < 3; ++k) cc.V[i * 4 + k] = cam.V[k * 4 + i];
        cc.V[i * 4 + 3] = cam.V[12 + i];
    }
    cc.fx = cam.fx; cc.fy = cam.fy; cc.cx = 0.f; cc.cy = 0.f;
    cc.lim_x_pos = cam.limx; cc.lim_x_neg = cam.limx; cc.lim_y_pos = cam.limy; cc.lim_y_neg = cam.limy;
    project_gaussian_vjp(o.f, cc, 0.0f, 0.0f, v_z, v_ca, v_cb, v_cc, v_p, v_S);
    // screen-position path: m2x = ((hx*p_w + 1) W - 1)/2,  p_w = 1/(hw + 1e-7)
    const float v_ndc_x = 0.5f * (float)cam.W * v_m2x, v_ndc_y = 0.5f * (float)cam.H * v_m2y;
    const float v_hx = v_ndc_x * o.p_w, v_hy = v_ndc_y * o.p_w;
    const float v_hw = -(v_ndc_x * o.hx + v_ndc_y * o.hy) * o.p_w * o.p_w;
    for (int k = 0; k < 3; ++k) v_p[k] += cam.Pm[k * 4 + 0] * v_hx + cam.Pm[k * 4 + 1] * v_hy + cam.Pm[k * 4 + 3] * v_hw;
}

// S = M M^T, M = R(q) diag(mod*s) with R the un-normalised polynomial in q
EMD_HD void dg_covar_vjp(const float q[4], const float R[9], const float M[9], const float s_mod[3], float mod,
                         const float v_S[6], float v_q[4], float v_s[3]) {
    const float G[9] = {2.0f * v_S[0], v_S[1], v_S[2], v_S[1], 2.0f * v_S[3], v_S[4], v_S[2], v_S[4], 2.0f * v_S[5]};
    float vM[9], vR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) vM[i * 3 + j] = G[i * 3 + 0] * M[0 * 3 + j] + G[i * 3 + 1] * M[1 * 3 + j] + G[i * 3 + 2] * M[2 * 3 + j];
    for (int j = 0; j < 3; ++j) {
        v_s[j] = mod * (R[0 * 3 + j] * vM[0 * 3 + j] + R[1 * 3 + j] * vM[1 * 3 + j] + R[2 * 3 + j] * vM[2 * 3 + j]);
        for (int i = 0; i < 3; ++i) vR[i * 3 + j] = vM[i * 3 + j] * s_mod[j];
    }
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    v_q[0] = 2.0f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    v_q[1] = 2.0f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.0f * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] - 2.0f * x * vR[8]);
    v_q[2] = 2.0f * (-2.0f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] - 2.0f * y * vR[8]);
    v_q[3] = 2.0f * (-2.0f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.0f * z * vR[4] + y * vR[5] + x * vR[6] + y * vR[7]);
}

// ---- SH colour with the view-direction gradient (computeColorFromSH fwd/bwd) ------------------
EMD_HD void dg_sh_bases(int deg, float x, float y, float z, float* b) {
    b[0] = 0.28209479177387814f;
    if (deg < 1) return;
    b[1] = -0.4886025119029199f * y; b[2] = 0.4886025119029199f * z; b[3] = -0.4886025119029199f * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy; b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz; b[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[9] = -0.5900435899266435f * y * (3.0f * xx - yy); b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy); b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// d(bases)/d(x,y,z) of the unit direction
EMD_HD void dg_sh_bases_grad(int deg, float x, float y, float z, float* bx, float* by, float* bz) {
    for (int k = 0; k < 16; ++k) { bx[k] = 0.f; by[k] = 0.f; bz[k] = 0.f; }
    if (deg < 1) return;
    const float C1 = 0.4886025119029199f;
    by[1] = -C1; bz[2] = C1; bx[3] = -C1;
    if (deg < 2) return;
    const float a0 = 1.0925484305920792f, a2 = 0.31539156525252005f, a4 = 0.5462742152960396f;
    bx[4] = a0 * y; by[4] = a0 * x;
    by[5] = -a0 * z; bz[5] = -a0 * y;
    bx[6] = -2.0f * a2 * x; by[6] = -2.0f * a2 * y; bz[6] = 4.0f * a2 * z;
    bx[7] = -a0 * z; bz[7] = -a0 * x;
    bx[8] = 2.0f * a4 * x; by[8] = -2.0f * a4 * y;
    if (deg < 3) return;
    const float c0 = -0.5900435899266435f, c1 = 2.890611442640554f, c2 = -0.4570457994644658f,
                c3 = 0.3731763325901154f, c5 = 1.445305721320277f;
    const float xx = x * x, yy = y * y, zz = z * z;
    bx[9] = c0 * 6.0f * x * y;          by[9] = c0 * (3.0f * xx - 3.0f * yy);
    bx[10] = c1 * y * z;                by[10] = c1 * x * z;                     bz[10] = c1 * x * y;
    bx[11] = c2 * (-2.0f * x * y);      by[11] = c2 * (4.0f * zz - xx - 3.0f * yy); bz[11] = c2 * 8.0f * y * z;
    bx[12] = c3 * (-6.0f * x * z);      by[12] = c3 * (-6.0f * y * z);           bz[12] = c3 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
    bx[13] = c2 * (4.0f * zz - 3.0f * xx - yy); by[13] = c2 * (-2.0f * x * y);   bz[13] = c2 * 8.0f * x * z;
    bx[14] = c5 * 2.0f * x * z;         by[14] = c5 * (-2.0f * y * z);           bz[14] = c5 * (xx - yy);
    bx[15] = c0 * (3.0f * xx - 3.0f * yy); by[15] = c0 * (-6.0f * x * y);
}
