// Per-pixel arithmetic of the fused image losses, host/device.
//
// Restates, per pixel, what the reference evaluates with ~40 full-image ATen launches between the rasterizer forward
// and backward:
//   OmniRe      sky blend + clamp           OmniRe/models/trainers/base.py:415, 486-493
//               compute_losses              OmniRe/models/trainers/base.py:518-587
//               DepthLoss / BCE / SafeBCE   OmniRe/models/losses.py:33-83, 91-172
//               SSIM                        pytorch_msssim.SSIM(data_range=1, channel=3) (base.py:114; third-party, absent)
//               inverse-depth smoothness    kornia.losses.inverse_depth_smoothness_loss (base.py:579; third-party, absent)
//   S3Gaussian  sky blend                   S3Gaussian/gaussian_renderer/__init__.py:299-300
//               l1 / ssim / compute_depth   S3Gaussian/utils/loss_utils.py:24-96, train.py:226, 348-363
// The CUDA kernels (image_loss.cu) and the host build (hostmath.cpp, checked against the oracle by
// `pytest -m "not gpu"`) both include this file.
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

// >>> ABI types (copied verbatim into include/emd_b200.h by tools/gen_header.py)
#define EMD_LOSS_TERMS 6      /* terms[c][k]: 0 l1, 1 ssim, 2 opacity (sky mask), 3 depth, 4 opacity entropy, 5 inverse-depth smoothness */
#define EMD_LOSS_SUMS 8       /* sums[c][k]: l1, ssim map, opacity, depth error, depth valid count, entropy, smooth-x, smooth-y */
#define EMD_SSIM_TAPS 11

typedef struct EmdImageLossConfig {
    /* weight of every term; 0 disables it (its sums are still reported) */
    float w_l1, w_ssim, w_opacity, w_depth, w_entropy, w_smooth;
    /* predicted colour: 0  min(rgb_g, 1) + sky * (1 - alpha)   (OmniRe base.py:415, 491)
     *                   1  rgb_g * alpha + sky * (1 - alpha)    (S3Gaussian gaussian_renderer/__init__.py:300)
     * (sky == NULL: 0 -> min(rgb_g, 1), 1 -> rgb_g) */
    int blend;
    /* SSIM window placement: 0 valid (pytorch_msssim: map is (H-10) x (W-10)), 1 zero 'same' padding (loss_utils.py:77) */
    int ssim_pad;
    /* opacity loss: 0 F.binary_cross_entropy (log clamped at -100), 1 SafeBCE(limit = bce_limit) (losses.py:33-75),
     *               2 S3Gaussian sky loss: x clamped to [1e-6, 1 - 1e-6] (train.py:361-362) */
    int opacity_loss;
    float bce_limit;
    /* depth loss: type 0 l1, 1 l2, 2 smooth_l1(beta 1); the switches of DepthLoss (losses.py:91-146) and compute_depth */
    int depth_type, depth_inverse, depth_normalize;
    int depth_pred_gate;     /* 1: a valid pixel also needs pred > 1e-4 (losses.py:124); 0: S3Gaussian */
    float depth_norm_lo;     /* lower clamp of the normalised depth: 1e-6 (safe_normalize_depth) or 0 (S3Gaussian) */
    float depth_max;         /* 80 */
    /* depth hit mask: 0 (lidar > 0) * valid_mask (base.py:555); 1 (1 - sky_mask) (train.py:350) */
    int depth_mask_mode;
    /* strides in floats; rows are dense (row stride = W * pixel stride).  _ps pixel, _cs channel, _vs view */
    int64_t rgb_ps, rgb_cs, rgb_vs;      /* rendered colour and its gradient */
    int64_t depth_ps, depth_vs;          /* rendered depth and its gradient */
    int64_t gt_ps, gt_cs, gt_vs;         /* ground-truth pixels */
    int64_t sky_ps, sky_cs, sky_vs;      /* sky colour and its gradient */
} EmdImageLossConfig;
// <<< ABI types

constexpr int SSIM_R = EMD_SSIM_TAPS / 2;

// ---- colour blend ---------------------------------------------------------------------------------------
// p = blend(rgb_g, alpha, sky); returns p and the partials dp/d rgb_g, dp/d alpha, dp/d sky
struct LossBlend {
    float p, d_rgb, d_alpha, d_sky;
};

EMD_HD LossBlend loss_blend(float rgb_g, float alpha, float sky, int has_sky, int mode) {
    LossBlend b;
    if (mode == 0) {
        const bool pass = rgb_g <= 1.0f;              // torch.clamp(max=1): the gradient passes at equality
        const float c = pass ? rgb_g : 1.0f;          // NaN stays NaN like torch.clamp
        b.p = has_sky ? c + sky * (1.0f - alpha) : c;
        b.d_rgb = pass ? 1.0f : 0.0f;
        b.d_alpha = has_sky ? -sky : 0.0f;
        b.d_sky = has_sky ? 1.0f - alpha : 0.0f;
    } else {
        b.p = has_sky ? rgb_g * alpha + sky * (1.0f - alpha) : rgb_g;
        b.d_rgb = has_sky ? alpha : 1.0f;
        b.d_alpha = has_sky ? rgb_g - sky : 0.0f;
        b.d_sky = has_sky ? 1.0f - alpha : 0.0f;
    }
    return b;
}

EMD_HD float loss_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// ---- SSIM -----------------------------------------------------------------------------------------------
// From the five windowed moments of (p, g): the map value and its partials w.r.t. E[p], E[pp], E[pg].
struct SsimPoint {
    float m, d_mu, d_pp, d_pg;
};

EMD_HD SsimPoint ssim_point(float mu_p, float mu_g, float e_pp, float e_gg, float e_pg) {
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float mpp = mu_p * mu_p, mgg = mu_g * mu_g, mpg = mu_p * mu_g;
    const float s_pp = e_pp - mpp, s_gg = e_gg - mgg, s_pg = e_pg - mpg;
    const float A = 2.0f * mpg + C1, B = 2.0f * s_pg + C2;
    const float Cc = mpp + mgg + C1, D = s_pp + s_gg + C2;
    const float inv = 1.0f / (Cc * D);
    SsimPoint r;
    r.m = A * B * inv;
    // dA/dmu_p = 2 mu_g, dB/dmu_p = -2 mu_g, dC/dmu_p = 2 mu_p, dD/dmu_p = -2 mu_p
    r.d_mu = 2.0f * inv * (mu_g * (B - A) - r.m * mu_p * (D - Cc));
    r.d_pp = -r.m / D;
    r.d_pg = 2.0f * A * inv;
    return r;
}

// ---- opacity (sky mask) loss ------------------------------------------------------------------------------
// x = predicted occupancy, t = target occupancy; returns the loss and d loss / d x
EMD_HD void loss_opacity(float x, float t, int kind, float limit, float& loss, float& d_x) {
    if (kind == 0) {          // F.binary_cross_entropy: -(t max(log x, -100) + (1-t) max(log(1-x), -100))
        const float lx = fmaxf(logf(x), -100.0f), l1x = fmaxf(log1pf(-x), -100.0f);
        loss = -(t * lx + (1.0f - t) * l1x);
        d_x = (x - t) / fmaxf((1.0f - x) * x, 1e-12f);
    } else if (kind == 1) {   // SafeBCE (losses.py:41-75)
        const float ln_limit = logf(limit);
        x = fminf(fmaxf(x, 0.0f), 1.0f);
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        const bool t0 = t == 0.0f;
        loss = -(t0 ? fmaxf(logf(1.0f - x), ln_limit) : fmaxf(logf(x), ln_limit));
        const float xc = t0 ? fminf(x, 1.0f - limit) : fmaxf(x, limit);
        d_x = (x == t) ? 0.0f : (t0 ? 1.0f / (1.0f - xc) : -1.0f / xc);
    } else {                  // S3Gaussian train.py:361-362; t = 1 - sky_mask
        const float lo = 1e-6f, hi = 1.0f - 1e-6f;
        const bool pass = x >= lo && x <= hi;
        const float w = fminf(fmaxf(x, lo), hi);
        const bool sky = t == 0.0f;
        loss = sky ? -logf(1.0f - w) : -logf(w);
        d_x = pass ? (sky ? 1.0f / (1.0f - w) : -1.0f / w) : 0.0f;
    }
}

// ---- depth loss -------------------------------------------------------------------------------------------
// pred / gt already multiplied by the hit mask.  Returns validity; err = per-pixel error, d_pred = d err / d pred.
EMD_HD bool loss_depth(float pred, float gt, const EmdImageLossConfig& c, float& err, float& d_pred) {
    err = 0.0f;
    d_pred = 0.0f;
    bool valid = gt > 0.01f && gt < c.depth_max;
    if (c.depth_pred_gate) valid = valid && pred > 0.0001f;
    if (!valid) return false;
    float p = pred, g = gt, dp = 1.0f;
    if (c.depth_normalize) {
        const float q = pred / c.depth_max;
        const bool pass = q >= c.depth_norm_lo && q <= 1.0f;
        p = fminf(fmaxf(q, c.depth_norm_lo), 1.0f);
        g = fminf(fmaxf(gt / c.depth_max, c.depth_norm_lo), 1.0f);
        dp = pass ? 1.0f / c.depth_max : 0.0f;
    }
    if (c.depth_inverse) {
        const float ip = 1.0f / p;
        dp *= -ip * ip;
        p = ip;
        g = 1.0f / g;
    }
    const float d = p - g;
    if (c.depth_type == 0) {
        err = fabsf(d);
        d_pred = loss_sign(d) * dp;
    } else if (c.depth_type == 1) {
        err = d * d;
        d_pred = 2.0f * d * dp;
    } else {
        const float a = fabsf(d);
        err = a < 1.0f ? 0.5f * d * d : a - 0.5f;
        d_pred = (a < 1.0f ? d : loss_sign(d)) * dp;
    }
    return true;
}

// ---- opacity entropy (base.py:568-573) ---------------------------------------------------------------------
EMD_HD void loss_entropy(float alpha, float& loss, float& d_alpha) {
    const float lo = 1e-6f, hi = 1.0f - 1e-6f;
    const bool pass = alpha >= lo && alpha <= hi;
    const float o = fminf(fmaxf(alpha, lo), hi);
    const float l = logf(o);
    loss = -o * l;
    d_alpha = pass ? -(l + 1.0f) : 0.0f;
}

// ---- inverse-depth smoothness (kornia formula; base.py:576-585) ---------------------------------------------
EMD_HD float loss_inv_depth(float depth) { return 1.0f / (depth + 1e-5f); }
// edge weight between two ground-truth pixels: exp(-mean_c |a_c - b_c|)
EMD_HD float loss_edge_weight(const float a[3], const float b[3]) {
    return expf(-(fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2])) / 3.0f);
}
