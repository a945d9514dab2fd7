// K6 / K7: front-to-back alpha compositing of RGB(+depth) and opacity per 16x16
// tile, forward and backward (gsplat rasterize_to_pixels semantics; the
// reference reaches it through OmniRe/models/trainers/base.py:393).
//
// Forward: one CTA per (camera, tile); the tile's depth-sorted Gaussians are
// staged through shared memory in batches of 256 packed records (48 B each),
// every thread composites one pixel, the CTA leaves as soon as all 256 pixels
// have saturated (barrier-count early termination).
//
// Backward: same CTA shape, back-to-front replay.  Per-pair gradients are
// reduced across the warp with a 16-value butterfly (16 shuffles instead of
// 12x5), across the CTA's 8 warps through fixed-order shared-memory slabs, and
// written ONCE per (tile, Gaussian) pair into that pair's private slot of an
// n_isects-long buffer (slot = the pair's index in emission order, which makes
// every Gaussian's slots contiguous).  A second kernel sums each Gaussian's
// contiguous run.  No global float atomics anywhere; results are bit-reproducible.
#include "common.cuh"
#include "proj_math.cuh"
#include "dg_math.cuh"

// flavour of the compositing rule: gsplat (pixel centre +0.5, alpha cap 0.999, stop at T <= 1e-4, gsplat tile
// rect) or diff_gauss / Inria (integer pixel coordinates, cap 0.99, stop at T < 1e-4, getRect)
struct RasterCfg {
    float px_off;
    float alpha_max;
    int strict_stop;
    int dg_rect;
};

namespace {

constexpr int RB = 256;  // threads per CTA == pixels per tile == Gaussians per staged batch
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr int NPART = 12;  // floats per (tile, Gaussian) gradient slot
constexpr int BSUB = 64;   // staged Gaussians per cross-warp reduction step of the backward
#ifndef EMD_SEG_BATCHES
#define EMD_SEG_BATCHES 4
#endif
constexpr int SEG_BATCHES = EMD_SEG_BATCHES;  // a segment = this many staged batches of 256 Gaussians of a tile's list
constexpr int SEG = SEG_BATCHES * RB;
constexpr int CKPT_FLOATS = 5 * RB;       // per segment boundary: T and acc[4] of the 256 pixels
constexpr int SEGOUT_FLOATS = 6 * RB;     // per segment of a multi-segment tile: T_end, local acc[4], last/stop code

// ---------------------------------------------------------------------------
// pack: gather the per-(camera,Gaussian) fields the compositor reads into three
// aligned float4 records so a staged Gaussian costs three 128-bit loads.
//   rec0 = (mean_x, mean_y, opacity, conic_a)
//   rec1 = (conic_b, conic_c, ch0, ch1)
//   rec2 = (ch2, ch3, hx, hy)
// Channels: the D_color colour channels, then (optionally) the camera depth.
// (hx, hy): half extents of the axis-aligned box around the region where the
// Gaussian can pass the compositor's alpha test, opacity * exp(-sigma) >= 1/255
// <=> sigma <= ln(255 opacity): |dx| <= sqrt(2 tau c / det), |dy| <= sqrt(2 tau a / det).
// Conservative (tau, det and the result carry safety margins far above the
// rounding of __expf / the conic), so skipping a (pixel block, Gaussian) pair
// outside the box never changes a result.  NaN = "never passes" (opacity < 1/255),
// +inf = "no bound" (degenerate conic).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void alpha_extent(float opac, float ca, float cb, float cc, float& hx, float& hy) {
    const float tau = __logf(255.0f * opac) + 0.02f;   // NaN / negative when the Gaussian can never reach 1/255
    const float ac = ca * cc;
    const float det = ac - cb * cb - 4e-6f * fabsf(ac);  // lower bound of the conic determinant under fp32 rounding
    if (!(det > 0.0f) || !(ca > 0.0f) || !(cc > 0.0f)) {
        hx = hy = (tau >= 0.0f) ? __int_as_float(0x7f800000) : __int_as_float(0x7fc00000);
        return;
    }
    const float k = 2.0f * tau / det;                   // negative tau -> sqrt(negative) = NaN -> culled everywhere
    hx = sqrtf(k * cc) * 1.001f + 0.02f;
    hy = sqrtf(k * ca) * 1.001f + 0.02f;
}

__global__ void raster_pack_kernel(const float* __restrict__ means2d, const float* __restrict__ conics,
                                   const float* __restrict__ opacities, int opac_per_cam,
                                   const float* __restrict__ colors, int colors_per_cam, int d_color,
                                   const float* __restrict__ depths, int with_depth,
                                   const int32_t* __restrict__ radii, int64_t N, int64_t CN,
                                   float4* __restrict__ recs) {
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= CN) return;
    if (radii[ci] <= 0) return;  // never gathered
    const int64_t i = ci % N;
    const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + ci);
    const float ca = __ldg(conics + ci * 3 + 0), cb = __ldg(conics + ci * 3 + 1), cc = __ldg(conics + ci * 3 + 2);
    const float op = __ldg(opacities + (opac_per_cam ? ci : i));
    float ch[4] = {0.f, 0.f, 0.f, 0.f};
    const float* cp = colors + (colors_per_cam ? ci : i) * d_color;
    for (int k = 0; k < d_color; ++k) ch[k] = __ldg(cp + k);
    if (with_depth) ch[d_color] = __ldg(depths + ci);
    float hx, hy;
    alpha_extent(op, ca, cb, cc, hx, hy);
    recs[ci * 3 + 0] = make_float4(m.x, m.y, op, ca);
    recs[ci * 3 + 1] = make_float4(cb, cc, ch[0], ch[1]);
    recs[ci * 3 + 2] = make_float4(ch[2], ch[3], hx, hy);
}

// A CTA's 8 warps own the 8 blocks of 8x4 pixels of a 16x16 tile (2 across, 4 down): compact footprints
// keep the per-warp "does any of my pixels see this Gaussian" rate low.
//   warp w -> block (bx = w & 1, by = w >> 1); lane l -> (lx = l & 7, ly = l >> 3)
// block_mask: bit w set iff the alpha box of a staged Gaussian reaches a pixel centre of block w.
// (cx0, cy0) = centre of the tile's first pixel.  Comparisons with NaN are false -> mask 0.
__device__ __forceinline__ uint32_t block_mask(float mx, float my, float hx, float hy, float cx0, float cy0) {
    const float xlo = mx - hx, xhi = mx + hx, ylo = my - hy, yhi = my + hy;
    uint32_t cols = 0, mask = 0;
    if (xhi >= cx0 && xlo <= cx0 + 7.0f) cols |= 1u;
    if (xhi >= cx0 + 8.0f && xlo <= cx0 + 15.0f) cols |= 2u;
#pragma unroll
    for (int by = 0; by < 4; ++by)
        if (yhi >= cy0 + 4.0f * by && ylo <= cy0 + 4.0f * by + 3.0f) mask |= cols << (2 * by);
    return mask;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp(-sigma); underflow flushes to zero (the alpha test rejects those anyway)
__device__ __forceinline__ float exp_neg(float sigma) { return ex2_approx(sigma * -1.4426950408889634f); }

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
// A tile whose sorted list is longer than SEG (1024) is composited by one CTA per 1024-Gaussian segment, all
// segments in parallel, so the horizon tiles (tens of thousands of Gaussians) no longer run as one sequential
// CTA at the tail of the grid:
//   pass T (raster_fwd_seg_kernel<false>): pure transmittance product of every non-final segment, prod(1 - alpha)
//   pass C (raster_fwd_seg_kernel<true>) : T at the segment's start = product of the segments in front of it (a
//           pixel whose product has fallen to the stop threshold has stopped in an earlier segment); exact
//           sequential compositing inside the segment; single-segment tiles finish here
//   combine (raster_fwd_combine_kernel)  : per pixel, in segment order: colour sums, last contributor, the
//           checkpoints (T, colour in front of every segment boundary) the segment-parallel backward starts from
struct FwdTile {
    int tile_id, cam, tile_x, tile_y, seg, nseg;
};
__device__ __forceinline__ bool decode_segment(const int32_t* __restrict__ seg_prefix, const int32_t* __restrict__ tile_order,
                                               int n_cam_tiles, int tile_w, int tile_h, FwdTile& ft) {
    if ((int)blockIdx.x >= seg_prefix[n_cam_tiles - 1]) return false;
    int lo_r = 0, hi_r = n_cam_tiles - 1;
    while (lo_r < hi_r) {
        const int mid = (lo_r + hi_r) >> 1;
        if (seg_prefix[mid] > (int)blockIdx.x) hi_r = mid; else lo_r = mid + 1;
    }
    const int before = lo_r > 0 ? seg_prefix[lo_r - 1] : 0;
    ft.seg = (int)blockIdx.x - before;
    ft.nseg = seg_prefix[lo_r] - before;
    ft.tile_id = tile_order[lo_r];
    ft.cam = ft.tile_id / (tile_w * tile_h);
    ft.tile_y = (ft.tile_id - ft.cam * tile_w * tile_h) / tile_w;
    ft.tile_x = ft.tile_id - (ft.cam * tile_h + ft.tile_y) * tile_w;
    return true;
}

template <bool COMPOSITE>
__global__ void __launch_bounds__(RB, 4) raster_fwd_seg_kernel(
    const float4* __restrict__ recs, const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
    const int32_t* __restrict__ tile_order, const int32_t* __restrict__ seg_prefix, const int32_t* __restrict__ ckpt_base,
    int64_t P, int C, int width, int height, int tile_w, int tile_h, int CH,
    int ed_mode, RasterCfg cfg, const float* __restrict__ backgrounds,
    float* __restrict__ ckpt, float* __restrict__ seg_out, float* __restrict__ out_colors, float* __restrict__ out_alphas,
    int32_t* __restrict__ last_ids) {
    __shared__ float4 s_g0[RB];    // mean x, mean y, opacity, A       (exp(-sigma) = exp2(A dx^2 + B dx dy + Cc dy^2))
    __shared__ float2 s_g1[RB];    // B, Cc
    __shared__ float4 s_col[RB];   // the four composited channels
    __shared__ uint32_t s_mask[RB];
    __shared__ __align__(4) uint8_t s_list[RB / 32][RB];   // per warp: staged slots whose alpha box reaches its block

    FwdTile ft;
    if (!decode_segment(seg_prefix, tile_order, C * tile_w * tile_h, tile_w, tile_h, ft)) return;
    if (!COMPOSITE && ft.seg == ft.nseg - 1) return;   // the product of a tile's last segment is never needed
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = ft.tile_y * EMD_TILE + (warp >> 1) * 4 + (lane >> 3);
    const int j = ft.tile_x * EMD_TILE + (warp & 1) * 8 + (lane & 7);
    const float px = (float)j + cfg.px_off, py = (float)i + cfg.px_off;
    const float cx0 = (float)(ft.tile_x * EMD_TILE) + cfg.px_off, cy0 = (float)(ft.tile_y * EMD_TILE) + cfg.px_off;
    const bool inside = i < height && j < width;
    bool done = !inside;

    // P < 2^31 (checked by the host wrapper): sorted indices fit 32 bits
    const int tile_start = tile_offsets[ft.tile_id];
    const int tile_end = (ft.tile_id == C * tile_h * tile_w - 1) ? (int)P : tile_offsets[ft.tile_id + 1];
    const int range_start = tile_start + ft.seg * SEG;
    const int range_end = min(tile_end, range_start + SEG);
    const int num_batches = (range_end - range_start + RB - 1) / RB;
    const int64_t slot0 = ft.nseg > 1 ? (int64_t)ckpt_base[ft.tile_id] : 0;

    float T = 1.0f;
    bool stopped = false;
    int32_t cur_idx = -1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (COMPOSITE && ft.seg > 0) {
        // transmittance in front of this segment: the same left-to-right product in every CTA that needs it
        for (int sp = 0; sp < ft.seg; ++sp) T *= ckpt[(slot0 + sp) * CKPT_FLOATS + tr];
        if (cfg.strict_stop ? (T < 1e-4f) : (T <= 1e-4f)) done = true;   // stopped in an earlier segment
    }

    for (int b = 0; b < num_batches; ++b) {
        if (COMPOSITE && __syncthreads_count(done) >= RB) break;
        if (!COMPOSITE) __syncthreads();
        const int batch_start = range_start + RB * b;
        const int idx = batch_start + tr;
        uint32_t mask = 0;
        if (idx < range_end) {
            const int64_t g = flatten_ids[idx];
            const float4 r0 = __ldg(recs + g * 3 + 0);
            const float4 r2 = __ldg(recs + g * 3 + 2);
            const float4 r1 = __ldg(recs + g * 3 + 1);
            // conic pre-scaled so that exp(-sigma) = exp2(A dx^2 + B dx dy + Cc dy^2)
            constexpr float L2E = 1.4426950408889634f;
            s_g0[tr] = make_float4(r0.x, r0.y, r0.z, -0.5f * L2E * r0.w);
            s_g1[tr] = make_float2(-L2E * r1.x, -0.5f * L2E * r1.y);
            if (COMPOSITE) s_col[tr] = make_float4(r1.z, r1.w, r2.x, r2.y);
            mask = block_mask(r0.x, r0.y, r2.z, r2.w, cx0, cy0);
        }
        s_mask[tr] = mask;
        __syncthreads();
        if (__all_sync(0xffffffffu, done)) continue;  // every pixel of this warp's block has saturated
        // This warp's candidate list: the staged Gaussians whose alpha box reaches its 8x4 block, compacted in order.
        int cnt = 0;
#pragma unroll
        for (int chunk = 0; chunk < RB / 32; ++chunk) {
            const bool c = (s_mask[chunk * 32 + lane] >> warp) & 1u;
            const uint32_t word = __ballot_sync(0xffffffffu, c);
            if (c) s_list[warp][cnt + __popc(word & ((1u << lane) - 1u))] = (uint8_t)(chunk * 32 + lane);
            cnt += __popc(word);
        }
        __syncwarp();
        // Groups of 4: the alphas do not depend on the running transmittance, so four are evaluated with full
        // ILP before the short sequential T / accumulate chain.  (Slots past cnt hold stale in-range indices.)
        for (int q = 0; q < cnt; q += 4) {
            const uchar4 t4 = *reinterpret_cast<const uchar4*>(&s_list[warp][q]);
            const int tt[4] = {t4.x, t4.y, t4.z, t4.w};
            float alpha[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 g0 = s_g0[tt[u]];
                const float2 g1 = s_g1[tt[u]];
                const float dx = g0.x - px, dy = g0.y - py;
                const float pw = fmaf(g1.y * dy, dy, fmaf(g1.x, dy, g0.w * dx) * dx);   // -sigma * log2(e)
                alpha[u] = fminf(cfg.alpha_max, g0.z * ex2_approx(pw));
                ok[u] = (q + u < cnt) && !(pw > 0.f || alpha[u] < ALPHA_MIN);
            }
            if (COMPOSITE) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!ok[u] || done) continue;
                    const float next_T = T * (1.0f - alpha[u]);
                    if (cfg.strict_stop ? (next_T < 1e-4f) : (next_T <= 1e-4f)) { done = true; stopped = true; continue; }
                    const float vis = alpha[u] * T;
                    const float4 col = s_col[tt[u]];
                    acc[0] += col.x * vis; acc[1] += col.y * vis; acc[2] += col.z * vis; acc[3] += col.w * vis;
                    cur_idx = batch_start + tt[u];
                    T = next_T;
                }
                if (__all_sync(0xffffffffu, done)) break;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u]) T *= 1.0f - alpha[u];
            }
        }
    }
    if (!COMPOSITE) {
        ckpt[(slot0 + ft.seg) * CKPT_FLOATS + tr] = T;   // staged in the checkpoint's T plane until the combine pass
        return;
    }
    if (ft.nseg > 1) {
        float* o = seg_out + (slot0 + ft.seg) * SEGOUT_FLOATS;
        o[tr] = T; o[RB + tr] = acc[0]; o[2 * RB + tr] = acc[1]; o[3 * RB + tr] = acc[2]; o[4 * RB + tr] = acc[3];
        // code: bit 31 = the pixel stopped inside this segment; low bits = 1 + sorted index of the last Gaussian
        // it blended here (0 = none)
        reinterpret_cast<uint32_t*>(o)[5 * RB + tr] = (uint32_t)(cur_idx + 1) | (stopped ? 0x80000000u : 0u);
        return;
    }
    if (inside) {
        const int64_t pix = ((int64_t)ft.cam * height + i) * width + j;
        const float alpha_out = 1.0f - T;
        out_alphas[pix] = alpha_out;
        if (backgrounds) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < CH) acc[k] += T * backgrounds[ft.cam * CH + k];
        }
        if (ed_mode) {
            const float inv = 1.0f / fmaxf(alpha_out, 1e-10f);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k == CH - 1) acc[k] *= inv;
        }
        float* o = out_colors + pix * CH;
        if (CH == 4) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < CH) o[k] = acc[k];
        }
        last_ids[pix] = max(cur_idx, 0);
    }
}

// One CTA per multi-segment tile (grid = tiles in heavy-first order; single-segment tiles leave at once).
__global__ void __launch_bounds__(RB) raster_fwd_combine_kernel(
    const int32_t* __restrict__ tile_order, const int32_t* __restrict__ seg_prefix, const int32_t* __restrict__ ckpt_base,
    int C, int width, int height, int tile_w, int tile_h, int CH, int ed_mode, RasterCfg cfg,
    const float* __restrict__ backgrounds, float* __restrict__ ckpt, const float* __restrict__ seg_out,
    float* __restrict__ out_colors, float* __restrict__ out_alphas, int32_t* __restrict__ last_ids) {
    const int r = blockIdx.x;
    const int nseg = seg_prefix[r] - (r > 0 ? seg_prefix[r - 1] : 0);
    if (nseg <= 1) return;
    const int tile_id = tile_order[r];
    const int cam = tile_id / (tile_w * tile_h);
    const int tile_y = (tile_id - cam * tile_w * tile_h) / tile_w;
    const int tile_x = tile_id - (cam * tile_h + tile_y) * tile_w;
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = tile_y * EMD_TILE + (warp >> 1) * 4 + (lane >> 3);
    const int j = tile_x * EMD_TILE + (warp & 1) * 8 + (lane & 7);
    const bool inside = i < height && j < width;
    const int64_t slot0 = ckpt_base[tile_id];
    float T = 1.0f;        // the pixel's transmittance so far
    float Tprod = 1.0f;    // product of the pure segment products (what pass C started each segment from)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int32_t last = 0;
    bool live = inside;
    for (int sgm = 0; sgm < nseg; ++sgm) {
        if (live && !(cfg.strict_stop ? (Tprod < 1e-4f) : (Tprod <= 1e-4f))) {
            const float* o = seg_out + (slot0 + sgm) * SEGOUT_FLOATS;
            T = o[tr];
            acc[0] += o[RB + tr]; acc[1] += o[2 * RB + tr]; acc[2] += o[3 * RB + tr]; acc[3] += o[4 * RB + tr];
            const uint32_t code = reinterpret_cast<const uint32_t*>(o)[5 * RB + tr];
            if (code & 0x7fffffffu) last = (int32_t)(code & 0x7fffffffu) - 1;
            if (code & 0x80000000u) live = false;
        } else {
            live = false;
        }
        if (sgm < nseg - 1) {
            // checkpoint at the boundary behind segment sgm: T before the next segment (as pass C used it) and the
            // colour accumulated in front of it.  Only read back for pixels that blended beyond the boundary.
            float* c = ckpt + (slot0 + sgm) * CKPT_FLOATS;
            Tprod *= c[tr];
            c[tr] = Tprod; c[RB + tr] = acc[0]; c[2 * RB + tr] = acc[1]; c[3 * RB + tr] = acc[2]; c[4 * RB + tr] = acc[3];
        }
    }
    if (inside) {
        const int64_t pix = ((int64_t)cam * height + i) * width + j;
        const float alpha_out = 1.0f - T;
        out_alphas[pix] = alpha_out;
        if (backgrounds) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < CH) acc[k] += T * backgrounds[cam * CH + k];
        }
        if (ed_mode) {
            const float inv = 1.0f / fmaxf(alpha_out, 1e-10f);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k == CH - 1) acc[k] *= inv;
        }
        float* o = out_colors + pix * CH;
        if (CH == 4) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < CH) o[k] = acc[k];
        }
        last_ids[pix] = last;
    }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
// Reduce-scatter of 12 per-lane values across the warp in 13 shuffles (6 + 3 + 2 + 1 + 1): every step halves the
// values a lane is responsible for.  On return the lanes with (lane & 1) == 0 and not ((lane & 4) && (lane & 2))
// hold, in the return value, the warp-wide sum of component
//   6 * bit4 + 3 * bit3 + (bit2 ? 2 : bit1)          (bitN = (lane >> N) & 1)
// Fixed association order -> bit-reproducible.
__device__ __forceinline__ float butterfly12(float (&v)[12], int lane) {
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float send = up ? v[k] : v[k + 6];
            const float keep = up ? v[k + 6] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float send = up ? v[k] : v[k + 3];
            const float keep = up ? v[k + 3] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0;
    // 3 -> (2 | 1): lanes with bit2 clear take components 0 and 1, lanes with bit2 set take component 2
    const float r0 = __shfl_xor_sync(0xffffffffu, up4 ? v[0] : v[2], 4);
    const float r1 = __shfl_xor_sync(0xffffffffu, v[1], 4);
    const float a0 = (up4 ? v[2] : v[0]) + r0;
    const float a1 = v[1] + r1;   // meaningful on the bit2-clear lanes only
    // (2 | 1) -> 1
    const float send = up4 ? a0 : (up2 ? a0 : a1);
    const float keep = up4 ? a0 : (up2 ? a1 : a0);
    const float c = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    return c + __shfl_xor_sync(0xffffffffu, c, 1);
}

// COUNT: measurement build of the same kernel (bench.py / tools only): counters[0] += (warp, candidate) evaluations,
// [1] += evaluations in which at least one lane blended, [2] += blended (pixel, Gaussian) pairs, [3] += staged
// (tile, Gaussian) pairs.  The product path launches COUNT = false.
template <bool COUNT>
__global__ void __launch_bounds__(RB) raster_bwd_kernel(
    const float4* __restrict__ recs, const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
    const int32_t* __restrict__ tile_order, const int32_t* __restrict__ radii, const int64_t* __restrict__ cum_tiles,
    int64_t P, int C, int width, int height, int tile_w, int tile_h, int CH, int ed_mode, RasterCfg cfg,
    const float* __restrict__ backgrounds, const int32_t* __restrict__ seg_prefix,
    const int32_t* __restrict__ ckpt_base, const float* __restrict__ ckpt,
    const float* __restrict__ out_colors, const float* __restrict__ out_alphas, const int32_t* __restrict__ last_ids,
    const float* __restrict__ v_out_colors, const float* __restrict__ v_out_alphas, float* __restrict__ partials,
    uint8_t* __restrict__ touched, unsigned long long* __restrict__ counters) {
    unsigned cnt_eval = 0, cnt_hit = 0, cnt_pairs = 0, cnt_staged = 0;
    __shared__ float4 s_r0[RB];
    __shared__ float4 s_r1[RB];
    __shared__ float2 s_r2[RB];
    __shared__ uint32_t s_slot[RB];
    __shared__ uint32_t s_mask[RB];
    __shared__ float s_slab[RB / 32][BSUB][NPART];
    __shared__ uint32_t s_tmask[RB / 32][BSUB / 32];
    __shared__ int s_red[RB / 32];

    // CTA -> (tile, segment): segments of all tiles are laid out heavy-tile-first; seg_prefix[r] is the inclusive
    // number of segments of the first r+1 tiles of that order (binary search; block-uniform)
    const int n_cam_tiles = C * tile_w * tile_h;
    if ((int)blockIdx.x >= seg_prefix[n_cam_tiles - 1]) return;
    int lo_r = 0, hi_r = n_cam_tiles - 1;
    while (lo_r < hi_r) {
        const int mid = (lo_r + hi_r) >> 1;
        if (seg_prefix[mid] > (int)blockIdx.x) hi_r = mid; else lo_r = mid + 1;
    }
    const int seg = (int)blockIdx.x - (lo_r > 0 ? seg_prefix[lo_r - 1] : 0);
    const int tile_id = tile_order[lo_r];
    const int cam = tile_id / (tile_w * tile_h);
    const int tile_y = (tile_id - cam * tile_w * tile_h) / tile_w;
    const int tile_x = tile_id - (cam * tile_h + tile_y) * tile_w;
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = tile_y * EMD_TILE + (warp >> 1) * 4 + (lane >> 3);   // same pixel <-> thread map as the forward
    const int j = tile_x * EMD_TILE + (warp & 1) * 8 + (lane & 7);
    const float px = (float)j + cfg.px_off, py = (float)i + cfg.px_off;
    const float cx0 = (float)(tile_x * EMD_TILE) + cfg.px_off, cy0 = (float)(tile_y * EMD_TILE) + cfg.px_off;
    const bool inside = i < height && j < width;
    const int64_t pix = ((int64_t)cam * height + min(i, height - 1)) * width + min(j, width - 1);

    const int64_t tile_start = tile_offsets[tile_id];
    const int64_t tile_end = (tile_id == C * tile_h * tile_w - 1) ? P : (int64_t)tile_offsets[tile_id + 1];
    // this CTA's slice of the tile's sorted list
    const int64_t range_start = tile_start + (int64_t)seg * SEG;
    const int64_t range_end = min(tile_end, range_start + SEG);
    if (range_end <= range_start) return;

    // per-pixel state
    const float alpha_out = inside ? out_alphas[pix] : 0.f;
    const float T_final = 1.0f - alpha_out;
    float T = T_final;
    const int64_t bin_final = inside ? (int64_t)last_ids[pix] : -1;
    float v_c[4] = {0.f, 0.f, 0.f, 0.f};
    float v_a = 0.f;
    if (inside) {
        for (int k = 0; k < CH; ++k) v_c[k] = v_out_colors[pix * CH + k];
        v_a = v_out_alphas[pix];
        if (ed_mode) {
            // out = D / max(alpha, 1e-10):  v_D = v_out / a_c ;  v_alpha += -v_out * out / a_c  (when alpha > 1e-10)
            const float ac = fmaxf(alpha_out, 1e-10f);
            const float v_ed = v_c[CH - 1];
            if (alpha_out > 1e-10f) v_a += -v_ed * out_colors[pix * CH + CH - 1] / ac;
            v_c[CH - 1] = v_ed / ac;
        }
    }
    float bg_dot = 0.f;
    if (backgrounds) {
        for (int k = 0; k < CH; ++k) bg_dot += backgrounds[cam * CH + k] * v_c[k];
    }
    const float v_a_eff = T_final * (v_a - bg_dot);   // d(out)/d(alpha_i) term shared by every Gaussian of the pixel
    float buf[4] = {0.f, 0.f, 0.f, 0.f};
    if (inside && bin_final >= range_end) {
        // the pixel blended Gaussians beyond this segment: start from the forward checkpoint taken at the
        // segment's end (T before sorted index range_end; colour accumulated in front of it)
        const float* c = ckpt + ((int64_t)ckpt_base[tile_id] + seg) * CKPT_FLOATS;
        T = c[tr];
        float raw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < CH) {
                float o = out_colors[pix * CH + k];
                if (ed_mode && k == CH - 1) o *= fmaxf(alpha_out, 1e-10f);   // undo the expected-depth normalisation
                if (backgrounds) o -= T_final * backgrounds[cam * CH + k];   // undo the background term
                raw[k] = o;
            }
            buf[k] = raw[k] - c[(k + 1) * RB + tr];  // colour accumulated BEHIND the segment
        }
    }

    // last sorted index any pixel of this warp / this tile blended
    int wmax = (int)max(bin_final, (int64_t)-1);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_red[warp] = wmax;
    __syncthreads();
    int64_t tile_last = -1;
    for (int w = 0; w < RB / 32; ++w) tile_last = max(tile_last, (int64_t)s_red[w]);
    if (tile_last < range_start) return;  // nothing was blended in this tile
    const int64_t hi_end = min(range_end, tile_last + 1);  // exclusive

    const int num_batches = (int)((hi_end - range_start + RB - 1) / RB);
    for (int b = 0; b < num_batches; ++b) {
        // batch covers sorted indices [batch_lo, batch_hi), walked from the back
        const int64_t batch_hi = hi_end - (int64_t)RB * b;
        const int64_t batch_lo = max(range_start, batch_hi - RB);
        const int batch_size = (int)(batch_hi - batch_lo);
        __syncthreads();
        uint32_t mask = 0;
        if (tr < batch_size) {
            // slot tr holds sorted index batch_hi-1-tr  (slot 0 = farthest)
            const int64_t idx = batch_hi - 1 - tr;
            const int64_t g = flatten_ids[idx];
            const float4 r0 = __ldg(recs + g * 3 + 0);
            const float4 r2 = __ldg(recs + g * 3 + 2);
            s_r0[tr] = r0;
            s_r1[tr] = __ldg(recs + g * 3 + 1);
            s_r2[tr] = make_float2(r2.x, r2.y);
            mask = block_mask(r0.x, r0.y, r2.z, r2.w, cx0, cy0);
            int x0, y0, x1, y1;
            if (cfg.dg_rect) tile_rect_dg(r0.x, r0.y, radii[g], tile_w, tile_h, x0, y0, x1, y1);
            else tile_rect_c(r0.x, r0.y, radii[g], tile_w, tile_h, x0, y0, x1, y1);
            const int64_t base = g == 0 ? 0 : cum_tiles[g - 1];
            s_slot[tr] = (uint32_t)(base + (int64_t)(tile_y - y0) * (x1 - x0) + (tile_x - x0));
        }
        s_mask[tr] = mask;
        if (COUNT) cnt_staged += tr < batch_size;
        __syncthreads();
        for (int sub = 0; sub < batch_size; sub += BSUB) {
            const int sub_n = min(BSUB, batch_size - sub);
            uint32_t tmask[BSUB / 32];
#pragma unroll
            for (int h = 0; h < BSUB / 32; ++h) {
                tmask[h] = 0;
                if (sub + 32 * h >= batch_size) continue;   // block-uniform
                // this warp's candidates of 32 staged Gaussians: alpha box reaches its 8x4 block, and the
                // Gaussian is not behind everything the block's pixels blended
                const int64_t idx_l = batch_hi - 1 - (sub + 32 * h + lane);
                uint32_t word = __ballot_sync(0xffffffffu, ((s_mask[sub + 32 * h + lane] >> warp) & 1u) && idx_l <= (int64_t)wmax);
                while (word) {
                    const int u = __ffs(word) - 1;
                    word &= word - 1;
                    const int t = sub + 32 * h + u;
                    const int64_t idx = batch_hi - 1 - t;
                    bool valid = inside && idx <= bin_final;
                    const float4 r0 = s_r0[t];
                    const float4 r1 = s_r1[t];
                    const float dx = r0.x - px, dy = r0.y - py;
                    const float sigma = 0.5f * (r0.w * dx * dx + r1.y * dy * dy) + r1.x * dx * dy;
                    const float vis = exp_neg(sigma);
                    const float alpha = fminf(cfg.alpha_max, r0.z * vis);
                    if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
                    if (COUNT) { cnt_eval += lane == 0; cnt_pairs += valid; }
                    if (!__any_sync(0xffffffffu, valid)) continue;
                    if (COUNT) cnt_hit += lane == 0;
                    float v[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) v[k] = 0.f;
                    if (valid) {
                        const float2 r2 = s_r2[t];
                        const float col[4] = {r1.z, r1.w, r2.x, r2.y};
                        const float ra = __fdividef(1.0f, 1.0f - alpha);   // 1 - alpha in [1e-3, 1]: the fast reciprocal is safe
                        T *= ra;
                        const float fac = alpha * T;
                        float v_alpha = 0.f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            v[k] = fac * v_c[k];
                            v_alpha += (col[k] * T - buf[k] * ra) * v_c[k];
                        }
                        v_alpha += ra * v_a_eff;
                        const float opac = r0.z;
                        if (opac * vis <= cfg.alpha_max) {
                            const float v_sigma = -opac * vis * v_alpha;
                            v[4] = 0.5f * v_sigma * dx * dx;
                            v[5] = v_sigma * dx * dy;
                            v[6] = 0.5f * v_sigma * dy * dy;
                            const float gx = v_sigma * (r0.w * dx + r1.x * dy);
                            const float gy = v_sigma * (r1.x * dx + r1.y * dy);
                            v[7] = gx; v[8] = gy;
                            v[9] = fabsf(gx); v[10] = fabsf(gy);
                            v[11] = vis * v_alpha;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) buf[k] += col[k] * fac;
                    }
                    const float tot = butterfly12(v, lane);
                    const int vidx = 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1) + ((lane & 4) ? 2 : ((lane >> 1) & 1));
                    if ((lane & 1) == 0 && !((lane & 4) && (lane & 2))) s_slab[warp][32 * h + u][vidx] = tot;
                    tmask[h] |= 1u << u;
                }
            }
            bool any_t = false;
#pragma unroll
            for (int h = 0; h < BSUB / 32; ++h) {
                if (lane == 0) s_tmask[warp][h] = tmask[h];
                any_t |= tmask[h] != 0;
            }
            // barrier + block-wide "did any warp produce a partial for these Gaussians"
            if (__syncthreads_or(any_t)) {
                // fixed-order cross-warp sum, one writer per (Gaussian, component)
                for (int q = tr; q < sub_n * NPART; q += RB) {
                    const int u = q / NPART, k = q - u * NPART;
                    float sum = 0.f;
                    bool any = false;
#pragma unroll
                    for (int w = 0; w < RB / 32; ++w) {
                        if (s_tmask[w][u >> 5] & (1u << (u & 31))) { sum += s_slab[w][u][k]; any = true; }
                    }
                    if (any) {
                        const uint32_t slot = s_slot[sub + u];
                        partials[(int64_t)slot * NPART + k] = sum;
                        if (k == 0) touched[slot] = 1;
                    }
                }
                __syncthreads();
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            cnt_eval += __shfl_xor_sync(0xffffffffu, cnt_eval, o);
            cnt_hit += __shfl_xor_sync(0xffffffffu, cnt_hit, o);
            cnt_pairs += __shfl_xor_sync(0xffffffffu, cnt_pairs, o);
            cnt_staged += __shfl_xor_sync(0xffffffffu, cnt_staged, o);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, (unsigned long long)cnt_eval);
            atomicAdd(counters + 1, (unsigned long long)cnt_hit);
            atomicAdd(counters + 2, (unsigned long long)cnt_pairs);
            atomicAdd(counters + 3, (unsigned long long)cnt_staged);
        }
    }
}

// Heavy-first tile order + segment bookkeeping.  One CTA.
//   order[r]      : tile ids, bucketed by floor(log2(list length)), longest bucket first (coarse LPT; the order
//                   within a bucket is arbitrary -- it only affects scheduling, never results)
//   seg_prefix[r] : inclusive count of 1024-Gaussian segments of tiles order[0..r] (>= 1 per tile)
//   ckpt_base[t]  : first checkpoint / segment-output slot of TILE t (a tile with n > 1 segments owns n slots,
//                   a single-segment tile none)
__global__ void __launch_bounds__(1024) tile_order_kernel(const int32_t* __restrict__ tile_offsets, int64_t P,
                                                          int n_cam_tiles, int32_t* __restrict__ order,
                                                          int32_t* __restrict__ seg_prefix,
                                                          int32_t* __restrict__ ckpt_base) {
    __shared__ int s_cnt[33];
    __shared__ int s_base[33];
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x < 33) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n_cam_tiles; t += blockDim.x) {
        const int64_t e = t == n_cam_tiles - 1 ? P : (int64_t)tile_offsets[t + 1];
        const int len = (int)(e - tile_offsets[t]);
        atomicAdd(&s_cnt[len > 0 ? 32 - __clz(len) : 0], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 32; b >= 0; --b) { s_base[b] = run; run += s_cnt[b]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_cam_tiles; t += blockDim.x) {
        const int64_t e = t == n_cam_tiles - 1 ? P : (int64_t)tile_offsets[t + 1];
        const int len = (int)(e - tile_offsets[t]);
        order[atomicAdd(&s_base[len > 0 ? 32 - __clz(len) : 0], 1)] = t;
    }
    __syncthreads();  // order[] (global) is visible to the whole CTA
    if (seg_prefix == nullptr) return;
    // scan of (segments, single-segment tiles) packed as (low, high) 32-bit halves of one 64-bit value
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r0 = 0; r0 < n_cam_tiles; r0 += blockDim.x) {
        const int r = r0 + threadIdx.x;
        int nseg = 0, t = 0;
        if (r < n_cam_tiles) {
            t = order[r];
            const int64_t e = t == n_cam_tiles - 1 ? P : (int64_t)tile_offsets[t + 1];
            const int len = (int)(e - tile_offsets[t]);
            nseg = len > 0 ? (len + SEG - 1) / SEG : 1;
        }
        const long long mine = (long long)nseg | ((long long)(nseg == 1 ? 1 : 0) << 32);
        long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long incl = carry + inc + (warp > 0 ? s_warp[warp - 1] : 0ll);
        if (r < n_cam_tiles) {
            const long long excl = incl - mine;
            seg_prefix[r] = (int)(incl & 0xffffffffll);
            // slots of the multi-segment tiles before this one (single-segment tiles own none)
            ckpt_base[t] = (int)(excl & 0xffffffffll) - (int)(excl >> 32);
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = incl;
        __syncthreads();
    }
}

// Sum each Gaussian's contiguous run of slots -> dense per-(camera,Gaussian) grads.
__global__ void raster_gather_kernel(const float* __restrict__ partials, const uint8_t* __restrict__ touched,
                                     const int64_t* __restrict__ cum_tiles, int64_t CN, int d_color, int with_depth,
                                     float* __restrict__ v_means2d, float* __restrict__ v_means2d_abs,
                                     float* __restrict__ v_conics, float* __restrict__ v_colors,
                                     float* __restrict__ v_depths, float* __restrict__ v_opacities) {
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= CN) return;
    const int64_t lo = ci == 0 ? 0 : cum_tiles[ci - 1];
    const int64_t hi = cum_tiles[ci];
    float a[NPART];
#pragma unroll
    for (int k = 0; k < NPART; ++k) a[k] = 0.f;
    for (int64_t e = lo; e < hi; ++e) {
        if (!touched[e]) continue;
        const float4* p = reinterpret_cast<const float4*>(partials + e * NPART);
        const float4 p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2);
        a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w;
        a[4] += p1.x; a[5] += p1.y; a[6] += p1.z; a[7] += p1.w;
        a[8] += p2.x; a[9] += p2.y; a[10] += p2.z; a[11] += p2.w;
    }
    for (int k = 0; k < d_color; ++k) v_colors[ci * d_color + k] = a[k];
    if (with_depth) v_depths[ci] = a[d_color];
    v_conics[ci * 3 + 0] = a[4]; v_conics[ci * 3 + 1] = a[5]; v_conics[ci * 3 + 2] = a[6];
    reinterpret_cast<float2*>(v_means2d)[ci] = make_float2(a[7], a[8]);
    if (v_means2d_abs) reinterpret_cast<float2*>(v_means2d_abs)[ci] = make_float2(a[9], a[10]);
    v_opacities[ci] = a[11];
}

}  // namespace

extern "C" int emd_raster_pack(const float* means2d, const float* conics, const float* opacities, int opac_per_cam,
                               const float* colors, int colors_per_cam, int d_color, const float* depths,
                               int with_depth, const int32_t* radii, int64_t N, int64_t C, float* recs,
                               cudaStream_t stream) {
    EMD_CHECK_ARG(d_color >= 0 && d_color + (with_depth ? 1 : 0) <= 4 && d_color + (with_depth ? 1 : 0) >= 1,
                  "raster_pack: need 1..4 channels (got %d colour + %d depth)", d_color, with_depth);
    if (!emd_aligned(recs, 16) || !emd_aligned(means2d, 8)) {
        emd_set_error("raster_pack: recs must be 16-B, means2d 8-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t CN = C * N;
    if (CN == 0) return EMD_OK;
    EMD_LAUNCH(EK_RASTER_PACK, stream, raster_pack_kernel<<<(unsigned)emd_cdiv(CN, 256), 256, 0, stream>>>(
        means2d, conics, opacities, opac_per_cam, colors, colors_per_cam, d_color, depths, with_depth, radii, N, CN,
        reinterpret_cast<float4*>(recs)));
    EMD_CHECK_LAUNCH("raster_pack");
    return EMD_OK;
}

// Measurement aid: while a non-NULL device pointer to 8 uint64 counters is set, emd_rasterize_bwd launches the counting
// build of its kernel (see raster_bwd_kernel<COUNT>).  Process-global; not for concurrent use.
static unsigned long long* g_raster_counters = nullptr;
extern "C" void emd_raster_set_counters(void* counters_u64x8) { g_raster_counters = reinterpret_cast<unsigned long long*>(counters_u64x8); }

extern "C" int emd_raster_segment_size() { return SEG; }
extern "C" int emd_raster_checkpoint_floats() { return CKPT_FLOATS; }
extern "C" int emd_raster_segout_floats() { return SEGOUT_FLOATS; }
// upper bound on the slots owned by multi-segment tiles: a tile of len > SEG has ceil(len/SEG) <= 2 len / SEG segments
extern "C" int64_t emd_raster_segment_slots(int64_t P) { return 2 * (P / SEG) + 2; }

// order / seg_prefix / ckpt_base: [C*tile_h*tile_w] int32 each (see tile_order_kernel).  The checkpoint buffer the
// forward fills and the backward reads holds emd_raster_segment_slots(P) * emd_raster_checkpoint_floats() floats, the
// forward's segment-output scratch emd_raster_segment_slots(P) * emd_raster_segout_floats(); forward and backward
// launch at most P / segment_size + n_cam_tiles CTAs.
extern "C" int emd_tile_order(const int32_t* tile_offsets, int64_t P, int64_t n_cam_tiles, int32_t* order,
                              int32_t* seg_prefix, int32_t* ckpt_base, cudaStream_t stream) {
    EMD_CHECK_ARG(n_cam_tiles >= 1 && n_cam_tiles < (1 << 30), "tile_order: bad tile count");
    EMD_CHECK_ARG((seg_prefix == nullptr) == (ckpt_base == nullptr), "tile_order: seg_prefix and ckpt_base go together");
    EMD_LAUNCH(EK_MISC, stream, tile_order_kernel<<<1, 1024, 0, stream>>>(tile_offsets, P, (int)n_cam_tiles, order, seg_prefix, ckpt_base));
    EMD_CHECK_LAUNCH("tile_order");
    return EMD_OK;
}

extern "C" int emd_rasterize_fwd(const float* recs, const int32_t* tile_offsets, const int32_t* flatten_ids,
                                 const int32_t* tile_order, const int32_t* seg_prefix, const int32_t* ckpt_base,
                                 int64_t P, int64_t C, int width, int height, int tile_w, int tile_h, int channels,
                                 int ed_mode, int flavour, const float* backgrounds, float* ckpt, float* seg_out,
                                 float* out_colors, float* out_alphas, int32_t* last_ids, cudaStream_t stream) {
    const RasterCfg cfg = flavour == 1 ? RasterCfg{0.0f, 0.99f, 1, 1} : RasterCfg{0.5f, 0.999f, 0, 0};
    EMD_CHECK_ARG(channels >= 1 && channels <= 4, "rasterize_fwd: channels must be 1..4");
    EMD_CHECK_ARG(C >= 1 && C * tile_w * tile_h < ((int64_t)1 << 30), "rasterize_fwd: grid too large");
    EMD_CHECK_ARG(tile_w == (width + EMD_TILE - 1) / EMD_TILE && tile_h == (height + EMD_TILE - 1) / EMD_TILE,
                  "rasterize_fwd: tile grid does not match image size (tile size is 16)");
    EMD_CHECK_ARG(tile_order && seg_prefix && ckpt_base && ckpt && seg_out,
                  "rasterize_fwd: needs tile_order, seg_prefix, ckpt_base and the ckpt / seg_out buffers");
    EMD_CHECK_ARG(P >= 0 && P < ((int64_t)1 << 31), "rasterize_fwd: too many intersections");
    if (!emd_aligned(recs, 16) || (channels == 4 && !emd_aligned(out_colors, 16))) {
        emd_set_error("rasterize_fwd: recs/out_colors must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int n_ct = (int)(C * tile_w * tile_h);
    // upper bound on the number of (tile, segment) CTAs; surplus CTAs exit at once
    dim3 grid((unsigned)(P / SEG + n_ct)), block(RB, 1, 1);
    const float4* r4 = reinterpret_cast<const float4*>(recs);
    if (P > SEG)  // only then can a tile have more than one segment
        EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_seg_kernel<false><<<grid, block, 0, stream>>>(
            r4, tile_offsets, flatten_ids, tile_order, seg_prefix, ckpt_base, P, (int)C, width, height, tile_w, tile_h,
            channels, ed_mode, cfg, backgrounds, ckpt, seg_out, out_colors, out_alphas, last_ids));
    EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_seg_kernel<true><<<grid, block, 0, stream>>>(
        r4, tile_offsets, flatten_ids, tile_order, seg_prefix, ckpt_base, P, (int)C, width, height, tile_w, tile_h,
        channels, ed_mode, cfg, backgrounds, ckpt, seg_out, out_colors, out_alphas, last_ids));
    if (P > SEG)
        EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_combine_kernel<<<n_ct, block, 0, stream>>>(
            tile_order, seg_prefix, ckpt_base, (int)C, width, height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds,
            ckpt, seg_out, out_colors, out_alphas, last_ids));
    EMD_CHECK_LAUNCH("rasterize_fwd");
    return EMD_OK;
}

extern "C" size_t emd_rasterize_bwd_workspace_bytes(int64_t P) {
    // [P][12] float partial slots + [P] touched flags
    const size_t part = ((size_t)P * NPART * sizeof(float) + 255) / 256 * 256;
    return part + (((size_t)P + 255) / 256) * 256 + 256;
}

extern "C" int emd_rasterize_bwd(const float* recs, const int32_t* tile_offsets, const int32_t* flatten_ids,
                                 const int32_t* tile_order, const int32_t* radii, const int64_t* cum_tiles, int64_t P, int64_t N, int64_t C,
                                 int width, int height, int tile_w, int tile_h, int channels, int ed_mode, int flavour,
                                 const float* backgrounds, const int32_t* seg_prefix, const int32_t* ckpt_base,
                                 const float* ckpt, const float* out_colors, const float* out_alphas,
                                 const int32_t* last_ids, const float* v_out_colors, const float* v_out_alphas,
                                 int d_color, int with_depth, float* v_means2d, float* v_means2d_abs, float* v_conics,
                                 float* v_colors, float* v_depths, float* v_opacities, void* workspace,
                                 size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(channels >= 1 && channels <= 4, "rasterize_bwd: channels must be 1..4");
    EMD_CHECK_ARG(d_color + (with_depth ? 1 : 0) == channels, "rasterize_bwd: channel bookkeeping mismatch");
    const RasterCfg cfg = flavour == 1 ? RasterCfg{0.0f, 0.99f, 1, 1} : RasterCfg{0.5f, 0.999f, 0, 0};
    EMD_CHECK_ARG(P < ((int64_t)1 << 32), "rasterize_bwd: too many intersections");
    EMD_CHECK_ARG(tile_order && seg_prefix && ckpt_base && ckpt, "rasterize_bwd: needs tile_order, seg_prefix, ckpt_base, ckpt");
    if (ws_bytes < emd_rasterize_bwd_workspace_bytes(P)) {
        emd_set_error("rasterize_bwd: workspace too small");
        return EMD_ERR_WORKSPACE;
    }
    if (!emd_aligned(workspace, 16) || !emd_aligned(recs, 16)) {
        emd_set_error("rasterize_bwd: workspace/recs must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t CN = C * N;
    if (CN == 0) return EMD_OK;
    float* partials = reinterpret_cast<float*>(workspace);
    const size_t part = ((size_t)P * NPART * sizeof(float) + 255) / 256 * 256;
    uint8_t* touched = reinterpret_cast<uint8_t*>(workspace) + part;
    if (P > 0) {
        cudaMemsetAsync(touched, 0, (size_t)P, stream);
        // upper bound on the number of (tile, segment) CTAs; surplus CTAs exit at once
        dim3 grid((unsigned)(P / SEG + C * tile_w * tile_h)), block(RB, 1, 1);
        if (g_raster_counters)
            EMD_LAUNCH(EK_RASTER_BWD, stream, raster_bwd_kernel<true><<<grid, block, 0, stream>>>(
                reinterpret_cast<const float4*>(recs), tile_offsets, flatten_ids, tile_order, radii, cum_tiles, P, (int)C, width,
                height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds, seg_prefix, ckpt_base, ckpt, out_colors, out_alphas,
                last_ids, v_out_colors, v_out_alphas, partials, touched, g_raster_counters));
        else
            EMD_LAUNCH(EK_RASTER_BWD, stream, raster_bwd_kernel<false><<<grid, block, 0, stream>>>(
                reinterpret_cast<const float4*>(recs), tile_offsets, flatten_ids, tile_order, radii, cum_tiles, P, (int)C, width,
                height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds, seg_prefix, ckpt_base, ckpt, out_colors, out_alphas,
                last_ids, v_out_colors, v_out_alphas, partials, touched, nullptr));
    }
    EMD_LAUNCH(EK_RASTER_GATHER, stream, raster_gather_kernel<<<(unsigned)emd_cdiv(CN, 256), 256, 0, stream>>>(partials, touched, cum_tiles, CN, d_color,
                                                                           with_depth, v_means2d, v_means2d_abs,
                                                                           v_conics, v_colors, v_depths, v_opacities));
    EMD_CHECK_LAUNCH("rasterize_bwd");
    return EMD_OK;
}
