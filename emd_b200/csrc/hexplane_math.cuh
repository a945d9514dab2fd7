// HexPlane bilinear-tap arithmetic (K1e), host/device.
//
// Restates ATen's grid_sampler_2d for the one configuration the reference uses
// (S3Gaussian/scene/hexplane.py:39-43: mode='bilinear', padding_mode='border', align_corners=True)
// and the aabb normalisation (hexplane.py:19-20).  The CUDA kernels (hexplane.cu) and the host
// build (hostmath.cpp, checked against the oracle by `pytest -m "not gpu"`) both include this file.
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

constexpr int HEX_PLANES = 6;        // C(4,2) coordinate pairs of (x, y, z, t)
constexpr int HEX_MAX_SCALES = 8;
constexpr int HEX_F = 32;            // features per plane (reference: output_coordinate_dim = 32)

// coordinate pair of plane p: itertools.combinations(range(4), 2)  (hexplane.py:82-84)
#define HEX_AX(p) ((p) < 3 ? 0 : ((p) < 5 ? 1 : 2))
#define HEX_AY(p) ((p) == 0 ? 1 : ((p) == 1 || (p) == 3 ? 2 : 3))

// One axis of a bilinear tap: cell index, its neighbour (clamped: the out-of-range corner has weight 0),
// fractional weight and d(ix)/d(u) (0 where the border clip is active, boundaries included).
struct HexAxis {
    int i0, i1;
    float w1;      // weight of i1; weight of i0 is 1 - w1
    float dmul;    // (size-1)/2 inside, 0 where clipped
};

EMD_HD HexAxis hex_axis(float u, int size) {
    HexAxis a;
    const float hi = (float)(size - 1);
    float ix = ((u + 1.0f) * 0.5f) * hi;
    const bool inside = (ix > 0.0f) && (ix < hi);     // NaN -> clipped to 0 like ATen's min/max
    ix = ix > 0.0f ? ix : 0.0f;
    ix = ix < hi ? ix : hi;
    const float f = floorf(ix);
    a.i0 = (int)f;
    a.i1 = a.i0 + 1 < size ? a.i0 + 1 : size - 1;
    a.w1 = ix - f;
    a.dmul = inside ? 0.5f * hi : 0.0f;
    return a;
}

// p -> u = (p - a0) * (2 / (a1 - a0)) - 1 ; k = 2 / (a1 - a0) precomputed in fp32 on the host
// (two rounded operations like the reference's `(pts - aabb[0]) * k - 1.0`: a contracted FMA moves the tap coordinate by
// up to half an ulp of (p - a0) * k, which the fine planes (512 texels) amplify to ~2e-6 in the features)
#ifdef __CUDA_ARCH__
EMD_HD float hex_normalize(float p, float a0, float k) { return __fsub_rn(__fmul_rn(__fsub_rn(p, a0), k), 1.0f); }
#else
EMD_HD float hex_normalize(float p, float a0, float k) { return (p - a0) * k - 1.0f; }
#endif

// out[p] = product of v[q] over q != p (prefix/suffix products: planes may hold exact zeros)
EMD_HD void hex_excl_products(const float v[HEX_PLANES], float out[HEX_PLANES]) {
    float pre = 1.0f;
#pragma unroll
    for (int p = 0; p < HEX_PLANES; ++p) {
        out[p] = pre;
        pre *= v[p];
    }
    float suf = 1.0f;
#pragma unroll
    for (int p = HEX_PLANES - 1; p >= 0; --p) {
        out[p] *= suf;
        suf *= v[p];
    }
}
