// Fused image losses between the rasterizer forward and backward.
//
// The reference evaluates its photometric / SSIM / sky-opacity / lidar-depth / regularisation terms with ~40 full-image
// ATen launches per view and lets autograd replay as many to obtain the cotangents of the render
// (OmniRe/models/trainers/base.py:486-493, 518-587; S3Gaussian/train.py:226, 348-363 + utils/loss_utils.py:66-96).
// Here: ONE forward kernel (sky blend + every per-pixel term + the separable 11x11 SSIM moments in shared memory, per-CTA
// partial sums), a one-CTA-per-view finalize (fixed-order, double accumulation -> bit-reproducible), and ONE backward
// kernel that convolves the three SSIM partial-derivative maps and writes the cotangents of the rendered colour, depth,
// opacity and sky colour in the layout the rasterizer backward consumes.  HBM-bound streaming: ~60 B/pixel read +
// 36 B/pixel written forward, ~100 B/pixel read + 32 B/pixel written backward.
#include "common.cuh"
#include "loss_math.cuh"

constexpr int LTX = 32, LTY = 16;                       // tile (pixels); 512 threads, one pixel each
constexpr int LHX = LTX + 2 * SSIM_R, LHY = LTY + 2 * SSIM_R;   // 42 x 26 halo tile
constexpr int LPITCH = LHX + 1;                         // smem row pitch of the halo tiles
constexpr int LOSS_THREADS = LTX * LTY;

struct LossArgs {
    const float *rgb, *depth, *alpha, *sky, *gt, *valid_mask, *sky_mask, *lidar;
    int C, H, W, tiles_x, tiles_y;
    int packed4;             // renders are [.,.,.,4] with the depth in channel 3: one 16-byte load per pixel
    EmdImageLossConfig cfg;
    float win[EMD_SSIM_TAPS];
    // forward outputs / backward inputs
    float* ssim_maps;        // [C][3 channels][3 maps: d/dE[p], d/dE[pp], d/dE[pg]][H][W]
    float* partials;         // [C][tiles][EMD_LOSS_SUMS]
    float* sums;             // [C][EMD_LOSS_SUMS]
    float* terms;            // [C][EMD_LOSS_TERMS]
    // backward
    const float* v_terms;    // [C][EMD_LOSS_TERMS]
    float *v_rgb, *v_depth, *v_alpha, *v_sky;
};

// one pixel's inputs of the colour path
struct LossPixel {
    float rgb[3], sky[3], gt[3], alpha, valid, depth;
};

__device__ __forceinline__ void loss_load_pixel(const LossArgs& A, int c, int64_t pix, LossPixel& P) {
    const int64_t vpix = (int64_t)c * A.H * A.W + pix;
    P.valid = A.valid_mask ? __ldg(A.valid_mask + vpix) : 1.0f;
    P.alpha = __ldg(A.alpha + vpix);
    P.depth = 0.0f;
    if (A.packed4) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(A.rgb + c * A.cfg.rgb_vs) + pix);
        P.rgb[0] = r.x; P.rgb[1] = r.y; P.rgb[2] = r.z; P.depth = r.w;
    } else {
        const float* r = A.rgb + c * A.cfg.rgb_vs + pix * A.cfg.rgb_ps;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) P.rgb[ch] = __ldg(r + ch * A.cfg.rgb_cs);
        if (A.depth) P.depth = __ldg(A.depth + c * A.cfg.depth_vs + pix * A.cfg.depth_ps);
    }
    const float* g = A.gt + c * A.cfg.gt_vs + pix * A.cfg.gt_ps;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) P.gt[ch] = __ldg(g + ch * A.cfg.gt_cs);
    if (A.sky) {
        const float* k = A.sky + c * A.cfg.sky_vs + pix * A.cfg.sky_ps;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) P.sky[ch] = __ldg(k + ch * A.cfg.sky_cs);
    } else {
        P.sky[0] = P.sky[1] = P.sky[2] = 0.0f;
    }
}

__device__ __forceinline__ void loss_load_gt3(const LossArgs& A, int c, int y, int x, float out[3]) {
    const int64_t o = c * A.cfg.gt_vs + ((int64_t)y * A.W + x) * A.cfg.gt_ps;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[ch] = __ldg(A.gt + o + ch * A.cfg.gt_cs);
}

__device__ __forceinline__ float loss_load_depth(const LossArgs& A, int c, int y, int x) {
    return __ldg(A.depth + c * A.cfg.depth_vs + ((int64_t)y * A.W + x) * A.cfg.depth_ps);
}

// hit mask of the depth loss at a pixel
__device__ __forceinline__ float loss_hit(const LossArgs& A, int64_t vpix, float lidar, float valid) {
    if (A.cfg.depth_mask_mode == 0) return (lidar > 0.0f ? 1.0f : 0.0f) * valid;
    return A.sky_mask ? 1.0f - __ldg(A.sky_mask + vpix) : 1.0f;
}

// whether pixel (y, x) carries an SSIM map value
__device__ __forceinline__ bool ssim_in_map(const LossArgs& A, int y, int x) {
    if (y < 0 || y >= A.H || x < 0 || x >= A.W) return false;
    if (A.cfg.ssim_pad) return true;
    return y >= SSIM_R && y < A.H - SSIM_R && x >= SSIM_R && x < A.W - SSIM_R;
}

__global__ void __launch_bounds__(LOSS_THREADS, 2) image_loss_fwd_kernel(const __grid_constant__ LossArgs A) {
    __shared__ float sP[3][LHY][LPITCH], sG[3][LHY][LPITCH];     // predicted / ground-truth colour of the halo tile
    __shared__ float sH[5][LHY][LTX];                            // row-filtered moments of one channel
    __shared__ float sRed[LOSS_THREADS / 32][EMD_LOSS_SUMS];
    const int tid = threadIdx.x, tx = tid % LTX, ty = tid / LTX;
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LTX, y0 = blockIdx.y * LTY;
    const int x = x0 + tx, y = y0 + ty;
    const bool inside = x < A.W && y < A.H;
    const int64_t pix = (int64_t)y * A.W + x;
    const int64_t vpix = (int64_t)c * A.H * A.W + pix;
    const bool has_sky = A.sky != nullptr;
    float acc[EMD_LOSS_SUMS];
#pragma unroll
    for (int k = 0; k < EMD_LOSS_SUMS; ++k) acc[k] = 0.0f;

    // ---- halo tile: blended + masked prediction and masked ground truth, all three channels, one visit per pixel ----
    for (int i = tid; i < LHY * LHX; i += LOSS_THREADS) {
        const int r = i / LHX, q = i % LHX;
        const int yy = y0 + r - SSIM_R, xx = x0 + q - SSIM_R;
        float p[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f};
        if (yy >= 0 && yy < A.H && xx >= 0 && xx < A.W) {
            LossPixel P;
            loss_load_pixel(A, c, (int64_t)yy * A.W + xx, P);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                p[ch] = loss_blend(P.rgb[ch], P.alpha, P.sky[ch], has_sky, A.cfg.blend).p * P.valid;
                g[ch] = P.gt[ch] * P.valid;
            }
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            sP[ch][r][q] = p[ch];
            sG[ch][r][q] = g[ch];
        }
    }
    __syncthreads();

    // ---- SSIM: separable 11-tap moments, one channel at a time --------------------------------------------------------
    const bool in_map = ssim_in_map(A, y, x);
    for (int ch = 0; ch < 3; ++ch) {
        for (int i = tid; i < LHY * LTX; i += LOSS_THREADS) {
            const int r = i / LTX, q = i % LTX;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
            for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                const float w = A.win[k], p = sP[ch][r][q + k], g = sG[ch][r][q + k];
                const float wp = w * p, wg = w * g;
                a0 += wp;
                a1 += wg;
                a2 += wp * p;
                a3 += wg * g;
                a4 += wp * g;
            }
            sH[0][r][q] = a0; sH[1][r][q] = a1; sH[2][r][q] = a2; sH[3][r][q] = a3; sH[4][r][q] = a4;
        }
        __syncthreads();
        if (inside) {
            acc[0] += fabsf(sG[ch][ty + SSIM_R][tx + SSIM_R] - sP[ch][ty + SSIM_R][tx + SSIM_R]);
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                const float w = A.win[k];
#pragma unroll
                for (int j = 0; j < 5; ++j) m[j] += w * sH[j][ty + k][tx];
            }
            SsimPoint s = ssim_point(m[0], m[1], m[2], m[3], m[4]);
            if (!in_map) s.m = s.d_mu = s.d_pp = s.d_pg = 0.0f;
            acc[1] += s.m;
            float* o = A.ssim_maps + (((int64_t)c * 3 + ch) * 3) * A.H * A.W + pix;
            o[0] = s.d_mu;
            o[(int64_t)A.H * A.W] = s.d_pp;
            o[2 * (int64_t)A.H * A.W] = s.d_pg;
        }
        __syncthreads();
    }

    // ---- per-pixel terms -------------------------------------------------------------------------------------
    if (inside) {
        const float valid = A.valid_mask ? __ldg(A.valid_mask + vpix) : 1.0f;
        const float al = __ldg(A.alpha + vpix);
        if (A.sky_mask) {
            const float t = (1.0f - __ldg(A.sky_mask + vpix)) * valid;
            float l, d;
            loss_opacity(al * valid, t, A.cfg.opacity_loss, A.cfg.bce_limit, l, d);
            acc[2] += l;
        }
        {
            float l, d;
            loss_entropy(al, l, d);
            acc[5] += l;
        }
        if (A.depth) {
            const float dep = loss_load_depth(A, c, y, x);
            if (A.lidar) {
                const float li = __ldg(A.lidar + vpix);
                const float hit = loss_hit(A, vpix, li, valid);
                float e, d;
                if (loss_depth(dep * hit, li * hit, A.cfg, e, d)) {
                    acc[3] += e;
                    acc[4] += 1.0f;
                }
            }
            const float id = loss_inv_depth(dep);
            float g0[3];
            loss_load_gt3(A, c, y, x, g0);
            if (x + 1 < A.W) {
                float g1[3];
                loss_load_gt3(A, c, y, x + 1, g1);
                acc[6] += fabsf(id - loss_inv_depth(loss_load_depth(A, c, y, x + 1))) * loss_edge_weight(g0, g1);
            }
            if (y + 1 < A.H) {
                float g1[3];
                loss_load_gt3(A, c, y + 1, x, g1);
                acc[7] += fabsf(id - loss_inv_depth(loss_load_depth(A, c, y + 1, x))) * loss_edge_weight(g0, g1);
            }
        }
    }

    // ---- fixed-order block reduction ---------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < EMD_LOSS_SUMS; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) sRed[tid >> 5][k] = v;
    }
    __syncthreads();
    if (tid < EMD_LOSS_SUMS) {
        float v = 0.0f;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) v += sRed[w][tid];
        const int64_t cta = ((int64_t)c * A.tiles_y + blockIdx.y) * A.tiles_x + blockIdx.x;
        A.partials[cta * EMD_LOSS_SUMS + tid] = v;
    }
}

// normalisers shared by the finalize and the backward
struct LossNorm {
    double l1, ssim, hw, sx, sy;
};
__device__ __forceinline__ LossNorm loss_norm(const LossArgs& A) {
    LossNorm n;
    const double H = A.H, W = A.W;
    n.l1 = 3.0 * H * W;
    n.ssim = A.cfg.ssim_pad ? 3.0 * H * W : 3.0 * (H - 2 * SSIM_R) * (W - 2 * SSIM_R);
    n.hw = H * W;
    n.sx = H * (W - 1);
    n.sy = (H - 1) * W;
    return n;
}

__global__ void __launch_bounds__(256) image_loss_finalize_kernel(const __grid_constant__ LossArgs A) {
    __shared__ double sS[256][EMD_LOSS_SUMS];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int tiles = A.tiles_x * A.tiles_y;
    double acc[EMD_LOSS_SUMS];
#pragma unroll
    for (int k = 0; k < EMD_LOSS_SUMS; ++k) acc[k] = 0.0;
    for (int t = tid; t < tiles; t += 256) {
        const float* p = A.partials + ((int64_t)c * tiles + t) * EMD_LOSS_SUMS;
#pragma unroll
        for (int k = 0; k < EMD_LOSS_SUMS; ++k) acc[k] += (double)p[k];
    }
#pragma unroll
    for (int k = 0; k < EMD_LOSS_SUMS; ++k) sS[tid][k] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) {
#pragma unroll
            for (int k = 0; k < EMD_LOSS_SUMS; ++k) sS[tid][k] += sS[tid + s][k];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double* S = sS[0];
        for (int k = 0; k < EMD_LOSS_SUMS; ++k) A.sums[c * EMD_LOSS_SUMS + k] = (float)S[k];
        const LossNorm n = loss_norm(A);
        const EmdImageLossConfig& g = A.cfg;
        float* T = A.terms + c * EMD_LOSS_TERMS;
        T[0] = g.w_l1 != 0.f ? (float)(g.w_l1 * (S[0] / n.l1)) : 0.f;
        T[1] = g.w_ssim != 0.f ? (float)(g.w_ssim * (1.0 - S[1] / n.ssim)) : 0.f;
        T[2] = (g.w_opacity != 0.f && A.sky_mask) ? (float)(g.w_opacity * (S[2] / n.hw)) : 0.f;
        T[3] = (g.w_depth != 0.f && A.depth && A.lidar) ? (float)(g.w_depth * (S[3] / S[4])) : 0.f;   // no valid pixel: NaN, as torch's empty mean
        T[4] = g.w_entropy != 0.f ? (float)(g.w_entropy * (S[5] / n.hw)) : 0.f;
        T[5] = (g.w_smooth != 0.f && A.depth) ? (float)(g.w_smooth * (S[6] / n.sx + S[7] / n.sy)) : 0.f;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS, 2) image_loss_bwd_kernel(const __grid_constant__ LossArgs A) {
    __shared__ float sD[3][LHY][LPITCH];
    __shared__ float sH[3][LHY][LTX];
    const int tid = threadIdx.x, tx = tid % LTX, ty = tid / LTX;
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LTX, y0 = blockIdx.y * LTY;
    const int x = x0 + tx, y = y0 + ty;
    const bool inside = x < A.W && y < A.H;
    const int64_t HW = (int64_t)A.H * A.W;
    const int64_t pix = (int64_t)y * A.W + x;
    const int64_t vpix = (int64_t)c * HW + pix;
    const LossNorm n = loss_norm(A);
    const EmdImageLossConfig& g = A.cfg;
    const bool has_sky = A.sky != nullptr;
    const float* vt = A.v_terms + c * EMD_LOSS_TERMS;
    const float g_l1 = (float)(__ldg(vt + 0) * g.w_l1 / n.l1);
    const float g_ss = (float)(-(double)__ldg(vt + 1) * g.w_ssim / n.ssim);
    const float g_op = (float)(__ldg(vt + 2) * g.w_opacity / n.hw);
    const float cnt = __ldg(A.sums + c * EMD_LOSS_SUMS + 4);
    const float g_dp = cnt > 0.f ? __ldg(vt + 3) * g.w_depth / cnt : 0.f;
    const float g_en = (float)(__ldg(vt + 4) * g.w_entropy / n.hw);
    const float g_sx = (float)(__ldg(vt + 5) * g.w_smooth / n.sx);
    const float g_sy = (float)(__ldg(vt + 5) * g.w_smooth / n.sy);

    LossPixel P;
    P.valid = 1.0f; P.alpha = 0.0f; P.depth = 0.0f;
    if (inside) loss_load_pixel(A, c, pix, P);
    float d_alpha = 0.0f, v_rgb[3] = {0.f, 0.f, 0.f};
    for (int ch = 0; ch < 3; ++ch) {
        const float* maps = A.ssim_maps + (((int64_t)c * 3 + ch) * 3) * HW;
        for (int i = tid; i < LHY * LHX; i += LOSS_THREADS) {
            const int r = i / LHX, q = i % LHX;
            const int yy = y0 + r - SSIM_R, xx = x0 + q - SSIM_R;
            const bool ok = yy >= 0 && yy < A.H && xx >= 0 && xx < A.W;      // maps are zero outside the SSIM map region
            const int64_t o = (int64_t)yy * A.W + xx;
            sD[0][r][q] = ok ? __ldg(maps + o) : 0.0f;
            sD[1][r][q] = ok ? __ldg(maps + HW + o) : 0.0f;
            sD[2][r][q] = ok ? __ldg(maps + 2 * HW + o) : 0.0f;
        }
        __syncthreads();
        for (int i = tid; i < LHY * LTX; i += LOSS_THREADS) {
            const int r = i / LTX, q = i % LTX;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                const float w = A.win[k];
                a0 += w * sD[0][r][q + k];
                a1 += w * sD[1][r][q + k];
                a2 += w * sD[2][r][q + k];
            }
            sH[0][r][q] = a0; sH[1][r][q] = a1; sH[2][r][q] = a2;
        }
        __syncthreads();
        if (inside) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < EMD_SSIM_TAPS; ++k) {
                const float w = A.win[k];
                s0 += w * sH[0][ty + k][tx];
                s1 += w * sH[1][ty + k][tx];
                s2 += w * sH[2][ty + k][tx];
            }
            const LossBlend b = loss_blend(P.rgb[ch], P.alpha, P.sky[ch], has_sky, g.blend);
            const float p = b.p * P.valid;
            const float gt = P.gt[ch] * P.valid;
            const float dp = g_l1 * loss_sign(p - gt) + g_ss * (s0 + 2.0f * p * s1 + gt * s2);
            const float db = dp * P.valid;
            v_rgb[ch] = db * b.d_rgb;
            d_alpha += db * b.d_alpha;
            if (A.v_sky) A.v_sky[c * g.sky_vs + pix * g.sky_ps + ch * g.sky_cs] = db * b.d_sky;
        }
        // the next channel's loads overwrite sD only: sH is rewritten after the next barrier
    }
    if (!inside) return;
    const float valid = P.valid, al = P.alpha;
    if (A.sky_mask && g.w_opacity != 0.f) {
        const float t = (1.0f - __ldg(A.sky_mask + vpix)) * valid;
        float l, d;
        loss_opacity(al * valid, t, g.opacity_loss, g.bce_limit, l, d);
        d_alpha += g_op * d * valid;
    }
    if (g.w_entropy != 0.f) {
        float l, d;
        loss_entropy(al, l, d);
        d_alpha += g_en * d;
    }
    A.v_alpha[vpix] = d_alpha;
    float v = 0.0f;
    if (A.depth) {
        const float dep = P.depth;
        if (A.lidar && g.w_depth != 0.f) {
            const float li = __ldg(A.lidar + vpix);
            const float hit = loss_hit(A, vpix, li, valid);
            float e, d;
            if (loss_depth(dep * hit, li * hit, g, e, d)) v += g_dp * d * hit;
        }
        if (g.w_smooth != 0.f) {
            const float id = loss_inv_depth(dep);
            float g1[3];
            float did = 0.0f;
            if (x + 1 < A.W) {
                loss_load_gt3(A, c, y, x + 1, g1);
                did += g_sx * loss_sign(id - loss_inv_depth(loss_load_depth(A, c, y, x + 1))) * loss_edge_weight(P.gt, g1);
            }
            if (x > 0) {
                loss_load_gt3(A, c, y, x - 1, g1);
                did -= g_sx * loss_sign(loss_inv_depth(loss_load_depth(A, c, y, x - 1)) - id) * loss_edge_weight(g1, P.gt);
            }
            if (y + 1 < A.H) {
                loss_load_gt3(A, c, y + 1, x, g1);
                did += g_sy * loss_sign(id - loss_inv_depth(loss_load_depth(A, c, y + 1, x))) * loss_edge_weight(P.gt, g1);
            }
            if (y > 0) {
                loss_load_gt3(A, c, y - 1, x, g1);
                did -= g_sy * loss_sign(loss_inv_depth(loss_load_depth(A, c, y - 1, x)) - id) * loss_edge_weight(g1, P.gt);
            }
            v += did * (-id * id);
        }
    }
    if (A.packed4) {
        reinterpret_cast<float4*>(A.v_rgb + c * g.rgb_vs)[pix] = make_float4(v_rgb[0], v_rgb[1], v_rgb[2], v);
    } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) A.v_rgb[c * g.rgb_vs + pix * g.rgb_ps + ch * g.rgb_cs] = v_rgb[ch];
        if (A.depth) A.v_depth[c * g.depth_vs + pix * g.depth_ps] = v;
    }
}

static int loss_fill_args(LossArgs& A, const char* what, const float* rgb, const float* depth, const float* alpha,
                          const float* sky, const float* gt, const float* valid_mask, const float* sky_mask,
                          const float* lidar, int C, int H, int W, const EmdImageLossConfig* cfg, const float* window) {
    EMD_CHECK_ARG(cfg && window, "%s: null config / window", what);
    EMD_CHECK_ARG(C >= 1 && H >= 1 && W >= 1 && C <= 65535, "%s: bad sizes C=%d H=%d W=%d", what, C, H, W);
    EMD_CHECK_ARG(rgb && alpha && gt, "%s: rgb, alpha and gt are required", what);
    EMD_CHECK_ARG(cfg->ssim_pad || (H > 2 * SSIM_R && W > 2 * SSIM_R), "%s: valid-window SSIM needs H, W > %d (got %d x %d)",
                  what, 2 * SSIM_R, H, W);
    EMD_CHECK_ARG(cfg->blend == 0 || cfg->blend == 1, "%s: blend=%d", what, cfg->blend);
    EMD_CHECK_ARG(cfg->opacity_loss >= 0 && cfg->opacity_loss <= 2, "%s: opacity_loss=%d", what, cfg->opacity_loss);
    EMD_CHECK_ARG(cfg->depth_type >= 0 && cfg->depth_type <= 2, "%s: depth_type=%d", what, cfg->depth_type);
    EMD_CHECK_ARG(cfg->opacity_loss != 1 || (cfg->bce_limit > 0.f && cfg->bce_limit < 1.f), "%s: bce_limit=%g", what,
                  (double)cfg->bce_limit);
    A.rgb = rgb; A.depth = depth; A.alpha = alpha; A.sky = sky; A.gt = gt;
    A.valid_mask = valid_mask; A.sky_mask = sky_mask; A.lidar = lidar;
    A.C = C; A.H = H; A.W = W;
    A.tiles_x = (int)emd_cdiv(W, LTX);
    A.tiles_y = (int)emd_cdiv(H, LTY);
    A.packed4 = depth && depth == rgb + 3 && cfg->rgb_ps == 4 && cfg->rgb_cs == 1 && cfg->depth_ps == 4 &&
                cfg->depth_vs == cfg->rgb_vs && cfg->rgb_vs % 4 == 0 && emd_aligned(rgb, 16);
    A.cfg = *cfg;
    for (int k = 0; k < EMD_SSIM_TAPS; ++k) A.win[k] = window[k];
    A.ssim_maps = nullptr; A.partials = nullptr; A.sums = nullptr; A.terms = nullptr;
    A.v_terms = nullptr; A.v_rgb = A.v_depth = A.v_alpha = A.v_sky = nullptr;
    return EMD_OK;
}

// floats of the `partials` scratch buffer
extern "C" int64_t emd_image_loss_partials_floats(int C, int H, int W) {
    return (int64_t)C * emd_cdiv(W, LTX) * emd_cdiv(H, LTY) * EMD_LOSS_SUMS;
}

// Forward.  rgb / depth / gt / sky are addressed through the strides in cfg; alpha, valid_mask, sky_mask, lidar are dense
// [C,H,W].  depth, sky, valid_mask, sky_mask, lidar may be NULL (their terms are then 0).  window = HOST array of the 11
// normalised Gaussian taps.  Outputs: ssim_maps [C,3,3,H,W] (kept for the backward), partials (scratch), sums
// [C,EMD_LOSS_SUMS] (kept for the backward), terms [C,EMD_LOSS_TERMS] = weighted loss terms per view.
extern "C" int emd_image_loss_fwd(const float* rgb, const float* depth, const float* alpha, const float* sky, const float* gt,
                                  const float* valid_mask, const float* sky_mask, const float* lidar, int C, int H, int W,
                                  const EmdImageLossConfig* cfg, const float* window, float* ssim_maps, float* partials,
                                  float* sums, float* terms, cudaStream_t stream) {
    LossArgs A;
    const int rc = loss_fill_args(A, "emd_image_loss_fwd", rgb, depth, alpha, sky, gt, valid_mask, sky_mask, lidar, C, H, W,
                                  cfg, window);
    if (rc != EMD_OK) return rc;
    EMD_CHECK_ARG(ssim_maps && partials && sums && terms, "emd_image_loss_fwd: null output");
    A.ssim_maps = ssim_maps; A.partials = partials; A.sums = sums; A.terms = terms;
    const dim3 grid(A.tiles_x, A.tiles_y, C);
    EMD_LAUNCH(EK_LOSS_FWD, stream, (image_loss_fwd_kernel<<<grid, LOSS_THREADS, 0, stream>>>(A)));
    EMD_CHECK_LAUNCH("emd_image_loss_fwd");
    EMD_LAUNCH(EK_MISC, stream, (image_loss_finalize_kernel<<<C, 256, 0, stream>>>(A)));
    EMD_CHECK_LAUNCH("emd_image_loss_fwd(finalize)");
    return EMD_OK;
}

// Backward: cotangents of the rendered colour (rgb strides), depth (depth strides; NULL iff depth is NULL), opacity
// [C,H,W] and sky colour (sky strides; may be NULL) given v_terms [C,EMD_LOSS_TERMS] (DEVICE; no host read).  Every
// pixel of every output is written.
extern "C" int emd_image_loss_bwd(const float* rgb, const float* depth, const float* alpha, const float* sky, const float* gt,
                                  const float* valid_mask, const float* sky_mask, const float* lidar, int C, int H, int W,
                                  const EmdImageLossConfig* cfg, const float* window, const float* ssim_maps,
                                  const float* sums, const float* v_terms, float* v_rgb, float* v_depth, float* v_alpha,
                                  float* v_sky, cudaStream_t stream) {
    LossArgs A;
    const int rc = loss_fill_args(A, "emd_image_loss_bwd", rgb, depth, alpha, sky, gt, valid_mask, sky_mask, lidar, C, H, W,
                                  cfg, window);
    if (rc != EMD_OK) return rc;
    EMD_CHECK_ARG(ssim_maps && sums && v_terms && v_rgb && v_alpha, "emd_image_loss_bwd: null argument");
    EMD_CHECK_ARG((depth == nullptr) == (v_depth == nullptr), "emd_image_loss_bwd: v_depth must be given iff depth is");
    EMD_CHECK_ARG(!v_sky || sky, "emd_image_loss_bwd: v_sky without sky");
    A.ssim_maps = const_cast<float*>(ssim_maps);
    A.sums = const_cast<float*>(sums);
    A.v_terms = v_terms;
    A.v_rgb = v_rgb; A.v_depth = v_depth; A.v_alpha = v_alpha; A.v_sky = v_sky;
    A.packed4 = A.packed4 && v_depth == v_rgb + 3 && emd_aligned(v_rgb, 16);
    const dim3 grid(A.tiles_x, A.tiles_y, C);
    EMD_LAUNCH(EK_LOSS_BWD, stream, (image_loss_bwd_kernel<<<grid, LOSS_THREADS, 0, stream>>>(A)));
    EMD_CHECK_LAUNCH("emd_image_loss_bwd");
    return EMD_OK;
}
