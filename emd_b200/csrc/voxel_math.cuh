// Voxel LBS-weight lookup arithmetic (K1f), host/device.
//
// Restates VoxelDeformer.normalize + F.grid_sample on a 5-D input for the one configuration the reference
// uses (OmniRe/models/modules.py:612-632: mode='bilinear' i.e. trilinear, padding_mode='border',
// align_corners=True).  The CUDA kernels (voxel_lbs.cu) and the host build (hostmath.cpp, checked against
// the oracle by `pytest -m "not gpu"`) both include this file.  One axis of a tap is hex_axis() of
// hexplane_math.cuh (same index / weight / clip rules as the 2-D sampler).
#pragma once
#include "hexplane_math.cuh"

struct VoxGeom {
    int D, H, W;        // voxel grid (z, y, x)
    int J;              // channels per voxel (24 bones + optional extra), multiple of 4
    float ratio;        // resolution[long] / resolution[short]  (modules.py:481-486)
    int ratio_dim;      // coordinate (0 = x, 1 = y, 2 = z) that `normalize` multiplies by ratio
};

// The eight corners of one trilinear tap: per-axis indices / weights, and d(ix)/d(u) per axis.
struct VoxTap {
    HexAxis ax[3];      // x (along W), y (along H), z (along D)
};

// modules.py:627-632: x_n = (x - offset) / scale ; x_n[ratio_dim] *= ratio        (two / three rounded operations)
#ifdef __CUDA_ARCH__
EMD_HD float vox_normalize(float x, float off, float scale, float mul) { return __fmul_rn(__fdiv_rn(__fsub_rn(x, off), scale), mul); }
#else
EMD_HD float vox_normalize(float x, float off, float scale, float mul) { return ((x - off) / scale) * mul; }
#endif

EMD_HD void vox_tap(const VoxGeom& G, const float x[3], const float off[3], float scale, VoxTap& T) {
    const int size[3] = {G.W, G.H, G.D};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float u = vox_normalize(x[a], off[a], scale, a == G.ratio_dim ? G.ratio : 1.0f);
        T.ax[a] = hex_axis(u, size[a]);
    }
}

// voxel index (channel-less) of corner k = (kz << 2 | ky << 1 | kx) of instance b
EMD_HD int64_t vox_corner_index(const VoxGeom& G, const VoxTap& T, int b, int k) {
    const int x = (k & 1) ? T.ax[0].i1 : T.ax[0].i0;
    const int y = (k & 2) ? T.ax[1].i1 : T.ax[1].i0;
    const int z = (k & 4) ? T.ax[2].i1 : T.ax[2].i0;
    return (((int64_t)b * G.D + z) * G.H + y) * G.W + x;
}

// per-axis weight of corner k along axis a, and its derivative sign w.r.t. the un-normalised coordinate
EMD_HD float vox_axis_weight(const VoxTap& T, int a, int k) {
    const bool hi = (k >> a) & 1;
    return hi ? T.ax[a].w1 : 1.0f - T.ax[a].w1;
}
EMD_HD float vox_corner_weight(const VoxTap& T, int k) {
    return vox_axis_weight(T, 0, k) * vox_axis_weight(T, 1, k) * vox_axis_weight(T, 2, k);
}
// d (corner weight) / d (ix along axis a)  (before the dmul chain factor)
EMD_HD float vox_corner_dweight(const VoxTap& T, int a, int k) {
    const float s = ((k >> a) & 1) ? 1.0f : -1.0f;
    const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    return s * vox_axis_weight(T, a1, k) * vox_axis_weight(T, a2, k);
}
// d (ix along axis a) / d (x_a): dmul / scale (* ratio on the stretched axis); 0 where the border clip is active
EMD_HD float vox_coord_chain(const VoxGeom& G, const VoxTap& T, int a, float scale) {
    return T.ax[a].dmul / scale * (a == G.ratio_dim ? G.ratio : 1.0f);
}
