// K6 / K7: front-to-back alpha compositing of RGB(+depth) and opacity per 16x16
// tile, forward and backward (gsplat rasterize_to_pixels semantics; the
// reference reaches it through OmniRe/models/trainers/base.py:393; the Inria /
// diff_gauss rule through S3Gaussian/gaussian_renderer/__init__.py:145).
//
// Data layout.  After the sort, `raster_sort_records` writes ONE depth-sorted, per-tile-contiguous stream of 48-byte
// records (one per (tile, Gaussian) intersection, in sorted order): everything a compositing CTA needs about a pair --
// mean, opacity, conic pre-scaled for exp2, the four channels, the 8-bit mask of the tile's 8x4 pixel blocks the
// Gaussian's alpha >= 1/255 box reaches, and the pair's gradient slot.  A CTA's batch of 256 staged Gaussians is then one
// contiguous 12 KB block: it is brought into shared memory by ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA
// path: SASS UBLKCP) signalled on an mbarrier, two stages deep, so the next batch travels while the current one is
// composited.  No indexed gathers, no per-thread staging arithmetic inside the compositing kernels.
//
// Forward: one CTA per (tile, 1024-Gaussian segment); warp = 8x4 pixel block; per-warp candidate lists from the
// block masks; alphas evaluated four at a time ahead of the sequential transmittance chain; tiles longer than a
// segment run segment-parallel (transmittance pass, compositing pass, combine).
//
// Backward: same CTA shape, back-to-front from the forward's checkpoints, in two phases per warp:
//   phase 1 (lane = pixel): the sequential part only -- alpha, T, and two scalars per (pixel, Gaussian): the blend
//            weight fac = alpha * T_before and w = exp(-sigma) * dL/dalpha -- stored in a per-warp [16 x 32] matrix;
//   phase 2 (lane = Gaussian): every lane owns one Gaussian of the group and walks the block's pixels, accumulating
//            its 12 gradient components in registers from the matrix -- NO cross-lane reduction per pair (the
//            13-shuffle butterfly of the first design is gone).
// A warp's per-Gaussian sums go to a compact per-warp slab; once per 256-Gaussian batch the CTA adds the warps'
// entries in a fixed order and writes ONE 48-byte slot per (tile, Gaussian) pair (slot = the pair's index in emission
// order, which makes every Gaussian's slots contiguous); a second kernel sums each Gaussian's run.  No float atomics
// anywhere; results are bit-reproducible.
#include <mutex>

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
constexpr int NWARP = RB / 32;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr int NPART = 12;  // floats per (tile, Gaussian) gradient slot
#ifndef EMD_SEG_BATCHES
#define EMD_SEG_BATCHES 4
#endif
constexpr int SEG_BATCHES = EMD_SEG_BATCHES;  // a segment = this many staged batches of 256 Gaussians of a tile's list
constexpr int SEG = SEG_BATCHES * RB;
constexpr int CKPT_FLOATS = 5 * RB;       // per segment boundary: T and acc[4] of the 256 pixels
constexpr int SEGOUT_FLOATS = 6 * RB;     // per segment of a multi-segment tile: T_end, local acc[4], last/stop code
constexpr int REC_F4 = 3;                 // float4 per record (48 bytes)
constexpr float L2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// A CTA's 8 warps own the 8 blocks of 8x4 pixels of a 16x16 tile (2 across, 4 down): compact footprints
// keep the per-warp "does any of my pixels see this Gaussian" rate low.
//   warp w -> block (bx = w & 1, by = w >> 1); lane l -> (lx = l & 7, ly = l >> 3)
// block_mask: bit w set iff the alpha box of a Gaussian reaches a pixel centre of block w.
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

// ---------------------------------------------------------------------------
// pack: gather the per-(camera,Gaussian) fields the compositor reads into three
// aligned float4 records.
//   rec0 = (mean_x, mean_y, opacity, conic_a)
//   rec1 = (conic_b, conic_c, ch0, ch1)
//   rec2 = (ch2, ch3, hx, hy)
// Channels: the D_color colour channels, then (optionally) the camera depth.
// (hx, hy): half extents of the axis-aligned box around the region where the
// Gaussian can pass the compositor's alpha test, opacity * exp(-sigma) >= 1/255
// <=> sigma <= ln(255 opacity): |dx| <= sqrt(2 tau c / det), |dy| <= sqrt(2 tau a / det).
// Conservative (tau, det and the result carry safety margins far above the
// rounding of ex2.approx / the conic), so skipping a (pixel block, Gaussian) pair
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

// ---------------------------------------------------------------------------
// sorted record stream: one 48-byte record per (tile, Gaussian) intersection, in sorted order
//   s0 = (mean_x, mean_y, opacity, A)       exp(-sigma) = exp2(A dx^2 + B dx dy + Cc dy^2),  A = -log2(e)/2 * conic_a ...
//   s1 = (B, Cc, ch0, ch1)
//   s2 = (ch2, ch3, bits(block mask), bits(slot))
// slot = the pair's index in EMISSION order (camera-major, Gaussian, row-major tiles of its rectangle).  The backward
// owns one gradient entry per set bit of the mask; cand[slot] = popc(mask) is written here (when asked for) and its
// exclusive scan over the slots gives every pair its first entry: the entries of one Gaussian are contiguous, which
// is what lets the gather kernel stream them without atomics.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raster_sort_records_kernel(
    const float4* __restrict__ recs, const int64_t* __restrict__ isect_ids, const int32_t* __restrict__ flatten_ids,
    const int32_t* __restrict__ radii, const int64_t* __restrict__ cum_tiles, int64_t P, int tile_w, int tile_h,
    int tile_n_bits, RasterCfg cfg, float4* __restrict__ srecs, uint8_t* __restrict__ cand) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t g = flatten_ids[i];
    const int tile = (int)((isect_ids[i] >> 32) & (((int64_t)1 << tile_n_bits) - 1));
    const int tile_y = tile / tile_w, tile_x = tile - tile_y * tile_w;
    const float4 r0 = __ldg(recs + g * 3 + 0);
    const float4 r1 = __ldg(recs + g * 3 + 1);
    const float4 r2 = __ldg(recs + g * 3 + 2);
    const float cx0 = (float)(tile_x * EMD_TILE) + cfg.px_off, cy0 = (float)(tile_y * EMD_TILE) + cfg.px_off;
    const uint32_t mask = block_mask(r0.x, r0.y, r2.z, r2.w, cx0, cy0);
    int x0, y0, x1, y1;
    if (cfg.dg_rect) tile_rect_dg(r0.x, r0.y, radii[g], tile_w, tile_h, x0, y0, x1, y1);
    else tile_rect_c(r0.x, r0.y, radii[g], tile_w, tile_h, x0, y0, x1, y1);
    const int64_t base = g == 0 ? 0 : cum_tiles[g - 1];
    const uint32_t slot = (uint32_t)(base + (int64_t)(tile_y - y0) * (x1 - x0) + (tile_x - x0));
    if (cand) cand[slot] = (uint8_t)__popc(mask);
    srecs[i * 3 + 0] = make_float4(r0.x, r0.y, r0.z, -0.5f * L2E * r0.w);
    srecs[i * 3 + 1] = make_float4(-L2E * r1.x, -0.5f * L2E * r1.y, r1.z, r1.w);
    srecs[i * 3 + 2] = make_float4(r2.x, r2.y, __uint_as_float(mask), __uint_as_float(slot));
}

// ---------------------------------------------------------------------------
// bulk-copy staging (cp.async.bulk global -> shared, completion on an mbarrier)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_init(uint64_t* bars /*[2]*/) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// one thread: bring `count` records starting at sorted index `first` into `dst`; completion flips `bar`
__device__ __forceinline__ void stage_issue(const float4* __restrict__ srecs, int64_t first, int count, float4* dst, uint64_t* bar) {
    const uint32_t bytes = (uint32_t)count * (REC_F4 * 16);
    const uint32_t b = smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(srecs + first * REC_F4), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void stage_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(b), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

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
__device__ __forceinline__ bool decode_segment(const int32_t* __restrict__ seg_prefix, const int2* __restrict__ cta_map,
                                               int n_cam_tiles, int tile_w, int tile_h, FwdTile& ft) {
    if ((int)blockIdx.x >= seg_prefix[n_cam_tiles - 1]) return false;
    const int2 cm = cta_map[blockIdx.x];     // written by tile_order_kernel: segments of all tiles, heavy tiles first
    ft.seg = cm.y & 0xffff;
    ft.nseg = cm.y >> 16;
    ft.tile_id = cm.x;
    ft.cam = ft.tile_id / (tile_w * tile_h);
    ft.tile_y = (ft.tile_id - ft.cam * tile_w * tile_h) / tile_w;
    ft.tile_x = ft.tile_id - (ft.cam * tile_h + ft.tile_y) * tile_w;
    return true;
}

template <bool COMPOSITE>
__global__ void __launch_bounds__(RB, 4) raster_fwd_seg_kernel(
    const float4* __restrict__ srecs, const int32_t* __restrict__ tile_offsets,
    const int2* __restrict__ cta_map, const int32_t* __restrict__ seg_prefix, const int32_t* __restrict__ ckpt_base,
    int64_t P, int C, int width, int height, int tile_w, int tile_h, int CH,
    int ed_mode, RasterCfg cfg, const float* __restrict__ backgrounds,
    float* __restrict__ ckpt, float* __restrict__ seg_out, float* __restrict__ out_colors, float* __restrict__ out_alphas,
    int32_t* __restrict__ last_ids) {
    __shared__ __align__(128) float4 s_rec[2][RB * REC_F4];         // two stages of 256 records
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ __align__(4) uint8_t s_list[NWARP][RB];              // per warp: staged slots whose alpha box reaches its block

    FwdTile ft;
    if (!decode_segment(seg_prefix, cta_map, C * tile_w * tile_h, tile_w, tile_h, ft)) return;
    if (!COMPOSITE && ft.seg == ft.nseg - 1) return;   // the product of a tile's last segment is never needed
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = ft.tile_y * EMD_TILE + (warp >> 1) * 4 + (lane >> 3);
    const int j = ft.tile_x * EMD_TILE + (warp & 1) * 8 + (lane & 7);
    const float px = (float)j + cfg.px_off, py = (float)i + cfg.px_off;
    const bool inside = i < height && j < width;
    bool done = !inside;

    // P < 2^31 (checked by the host wrapper): sorted indices fit 32 bits
    const int tile_start = tile_offsets[ft.tile_id];
    const int tile_end = (ft.tile_id == C * tile_h * tile_w - 1) ? (int)P : tile_offsets[ft.tile_id + 1];
    const int range_start = tile_start + ft.seg * SEG;
    const int range_end = min(tile_end, range_start + SEG);
    const int num_batches = (range_end - range_start + RB - 1) / RB;
    const int64_t slot0 = ft.nseg > 1 ? (int64_t)ckpt_base[ft.tile_id] : 0;

    stage_init(s_bar);
    __syncthreads();
    if (tr == 0 && num_batches > 0) stage_issue(srecs, range_start, min(RB, range_end - range_start), s_rec[0], &s_bar[0]);

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
        const int st = b & 1;
        // everyone is past batch b-1: its stage may be refilled.  Barrier-count early termination: when all 256
        // pixels have saturated, drain the copy already in flight (it targets this CTA's shared memory) and leave.
        const bool all_done = COMPOSITE ? (__syncthreads_count(done) >= RB) : (__syncthreads(), false);
        if (tr == 0 && !all_done && b + 1 < num_batches) {
            const int nxt = range_start + RB * (b + 1);
            stage_issue(srecs, nxt, min(RB, range_end - nxt), s_rec[st ^ 1], &s_bar[st ^ 1]);
        }
        stage_wait(&s_bar[st], (b >> 1) & 1);
        if (all_done) break;
        const int batch_start = range_start + RB * b;
        const int batch_size = min(RB, range_end - batch_start);
        const float4* rec = s_rec[st];
        if (__all_sync(0xffffffffu, done)) continue;  // every pixel of this warp's block has saturated
        // This warp's candidate list: the staged Gaussians whose alpha box reaches its 8x4 block, compacted in order.
        int cnt = 0;
#pragma unroll
        for (int chunk = 0; chunk < RB / 32; ++chunk) {
            const int t = chunk * 32 + lane;
            const bool c = t < batch_size && ((__float_as_uint(rec[t * REC_F4 + 2].z) >> warp) & 1u);
            const uint32_t word = __ballot_sync(0xffffffffu, c);
            if (c) s_list[warp][cnt + __popc(word & ((1u << lane) - 1u))] = (uint8_t)t;
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
                const int t = (q + u < cnt) ? tt[u] : tt[0];     // stale list bytes may exceed the staged range
                const float4 g0 = rec[t * REC_F4];
                const float2 g1 = *reinterpret_cast<const float2*>(&rec[t * REC_F4 + 1]);
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
                    const float2 c01 = *reinterpret_cast<const float2*>(&rec[tt[u] * REC_F4 + 1].z);
                    const float2 c23 = *reinterpret_cast<const float2*>(&rec[tt[u] * REC_F4 + 2]);
                    acc[0] += c01.x * vis; acc[1] += c01.y * vis; acc[2] += c23.x * vis; acc[3] += c23.y * vis;
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
constexpr int GRP = 16;        // Gaussians per transposed accumulation group (lane & 15 owns one, lane >> 4 picks the pixel half)
constexpr int MROW = 33;       // float2 per matrix row: 32 pixels + 1 pad (conflict-free column reads in phase 2)
constexpr int BB = 128;        // records per stage of the backward's ring
constexpr int NSTAGE = 4;      // ring depth: a warp may run up to 3 batches ahead of the slowest warp of its CTA
constexpr int ENTRY_F = 16;    // floats per (pair, warp) gradient entry: 12 used, padded to one 64-byte line

struct BwdSmem {
    float4 rec[NSTAGE][BB * REC_F4];     // ring of staged record batches                             24 576 B
    float2 M[NWARP][GRP][MROW];          // per warp: (fac, w) per (group member, pixel)             33 792 B
    float4 vc[NWARP][34];                // per warp: the 32 pixels' colour cotangents (+1 pad slot)   4 352 B
    float4 grp[NWARP][GRP][2];           // per warp: (mx, my, o, A), (B, Cc, slot, rank) of each member 4 096 B
    uint8_t list[NWARP][BB];             // per warp: candidate staged indices of the batch             1 024 B
    uint64_t full[NSTAGE];               // mbarriers: stage filled
    int done[NSTAGE];                    // warps finished with the stage's current batch
    int red[NWARP];
};

// The 8 warps of a CTA share the staged records and nothing else: no CTA-wide barrier inside the main loop.  A warp
// that is finished with a stage bumps the stage's counter; the LAST of the 8 refills it (bulk copy of the batch
// NSTAGE further on).  Per-warp results go straight to global memory: one 64-byte entry per (pair, warp) -- first
// entry of the pair (entry_base[slot], the scan of the per-pair candidate counts) + rank of the warp among the mask's
// set bits -- and the entry's `touched` byte; raster_gather adds a Gaussian's contiguous entries in a fixed order
// (tile of the rectangle ascending, warp ascending).
// COUNT: measurement build of the same kernel (bench.py / tools only): counters[0] += (warp, candidate) evaluations,
// [1] += evaluations in which at least one lane blended, [2] += blended (pixel, Gaussian) pairs, [3] += staged
// (tile, Gaussian) pairs, [4] += phase-2 groups run, [5] += group members.  The product path launches COUNT = false.
template <bool COUNT>
__global__ void __launch_bounds__(RB, 3) raster_bwd_kernel(
    const float4* __restrict__ srecs, const int32_t* __restrict__ tile_offsets, const int2* __restrict__ cta_map,
    int64_t P, int C, int width, int height, int tile_w, int tile_h, int CH, int ed_mode, RasterCfg cfg,
    const float* __restrict__ backgrounds, const int32_t* __restrict__ seg_prefix,
    const int32_t* __restrict__ ckpt_base, const float* __restrict__ ckpt,
    const float* __restrict__ out_colors, const float* __restrict__ out_alphas, const int32_t* __restrict__ last_ids,
    const float* __restrict__ v_out_colors, const float* __restrict__ v_out_alphas,
    const uint32_t* __restrict__ entry_base, float* __restrict__ partials,
    uint8_t* __restrict__ touched, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);
    unsigned cnt_eval = 0, cnt_hit = 0, cnt_pairs = 0, cnt_staged = 0, cnt_groups = 0, cnt_members = 0;

    // CTA -> (tile, segment): table written by tile_order_kernel (segments of all tiles, heavy tiles first)
    const int n_cam_tiles = C * tile_w * tile_h;
    if ((int)blockIdx.x >= seg_prefix[n_cam_tiles - 1]) return;
    const int2 cm = cta_map[blockIdx.x];
    const int tile_id = cm.x, seg = cm.y & 0xffff;
    const int cam = tile_id / (tile_w * tile_h);
    const int tile_y = (tile_id - cam * tile_w * tile_h) / tile_w;
    const int tile_x = tile_id - (cam * tile_h + tile_y) * tile_w;
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = tile_y * EMD_TILE + (warp >> 1) * 4 + (lane >> 3);   // same pixel <-> thread map as the forward
    const int j = tile_x * EMD_TILE + (warp & 1) * 8 + (lane & 7);
    const float px = (float)j + cfg.px_off, py = (float)i + cfg.px_off;
    const bool inside = i < height && j < width;
    const int64_t pix = ((int64_t)cam * height + min(i, height - 1)) * width + min(j, width - 1);

    // P < 2^31 (checked by the host wrapper): sorted indices fit 32 bits
    const int tile_start = tile_offsets[tile_id];
    const int tile_end = (tile_id == n_cam_tiles - 1) ? (int)P : tile_offsets[tile_id + 1];
    // this CTA's slice of the tile's sorted list
    const int range_start = tile_start + seg * SEG;
    const int range_end = min(tile_end, range_start + SEG);
    if (range_end <= range_start) return;

    // per-pixel state
    const float alpha_out = inside ? out_alphas[pix] : 0.f;
    const float T_final = 1.0f - alpha_out;
    float T = T_final;
    const int bin_final = inside ? last_ids[pix] : -1;
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
    // E = <colour accumulated BEHIND the current Gaussian, v_c> - T_final (v_alpha_out - <bg, v_c>): the one scalar the
    // per-pair dL/dalpha needs besides T:   dL/dalpha_i = T_i <c_i, v_c> - E / (1 - alpha_i)
    float E = -T_final * (v_a - bg_dot);
    if (inside && bin_final >= range_end) {
        // the pixel blended Gaussians beyond this segment: start from the forward checkpoint taken at the
        // segment's end (T before sorted index range_end; colour accumulated in front of it)
        const float* c = ckpt + ((int64_t)ckpt_base[tile_id] + seg) * CKPT_FLOATS;
        T = c[tr];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < CH) {
                float o = out_colors[pix * CH + k];
                if (ed_mode && k == CH - 1) o *= fmaxf(alpha_out, 1e-10f);   // undo the expected-depth normalisation
                if (backgrounds) o -= T_final * backgrounds[cam * CH + k];   // undo the background term
                E += (o - c[(k + 1) * RB + tr]) * v_c[k];                    // colour accumulated BEHIND the segment
            }
        }
    }
    S.vc[warp][lane + (lane >> 4)] = make_float4(v_c[0], v_c[1], v_c[2], v_c[3]);

    // last sorted index any pixel of this warp / this tile blended
    int wmax = bin_final;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) S.red[warp] = wmax;
    if (tr == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&S.full[s])) : "memory");
            S.done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    int tile_last = -1;
    for (int w = 0; w < NWARP; ++w) tile_last = max(tile_last, S.red[w]);
    if (tile_last < range_start) return;  // nothing was blended in this slice
    const int hi_end = min(range_end, tile_last + 1);  // exclusive

    // batch b covers sorted indices [max(range_start, hi_end - BB (b+1)), hi_end - BB b), walked from the back
    const int num_batches = (hi_end - range_start + BB - 1) / BB;
    if (tr == 0) {
        for (int b = 0; b < min(NSTAGE, num_batches); ++b) {
            const int bh = hi_end - BB * b, bl = max(range_start, bh - BB);
            stage_issue(srecs, bl, bh - bl, S.rec[b], &S.full[b]);
        }
        if (COUNT) cnt_staged += hi_end - range_start;
    }
    // block origin of this warp (pixel centres) for phase 2
    const float bx = (float)(tile_x * EMD_TILE + (warp & 1) * 8) + cfg.px_off;
    const float by = (float)(tile_y * EMD_TILE + (warp >> 1) * 4) + cfg.px_off;
    const int mem = lane & (GRP - 1), half = lane >> 4;
    int gcnt = 0;      // members of the open group (groups may span batches: members carry their own record copy)

    // phase 2 (lane & 15 = group member, lane >> 4 = pixel half): per-Gaussian sums over the warp's 8x4 block
    auto phase2 = [&]() {
        __syncwarp();
        if (COUNT) { cnt_groups += lane == 0; cnt_members += lane == 0 ? gcnt : 0; }
        const bool active = mem < gcnt;
        const float4 m0 = S.grp[warp][active ? mem : 0][0];
        const float4 m1 = S.grp[warp][active ? mem : 0][1];
        // first entry of the member's pair: a global load whose latency the pixel loop below covers
        const uint32_t ebase = (half == 0 && active) ? __ldg(entry_base + __float_as_uint(m1.z)) : 0u;
        const float ca = -2.0f * LN2 * m0.w, cb = -LN2 * m1.x, cc = -2.0f * LN2 * m1.y;   // the raw conic
        float a[NPART];
#pragma unroll
        for (int k = 0; k < NPART; ++k) a[k] = 0.f;
        const float2* Mrow = &S.M[warp][mem][half * 16];
        const float4* vcp = &S.vc[warp][half * 17];
        const float dx0 = m0.x - bx;
        const float dy0 = m0.y - (by + (float)(2 * half));
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const float2 fw = Mrow[jj];
            const float4 vc = vcp[jj];
            const float dx = dx0 - (float)(jj & 7);
            const float dy = dy0 - (float)(jj >> 3);
            a[0] = fmaf(fw.x, vc.x, a[0]); a[1] = fmaf(fw.x, vc.y, a[1]);
            a[2] = fmaf(fw.x, vc.z, a[2]); a[3] = fmaf(fw.x, vc.w, a[3]);
            const float s0 = fw.y * dx, s1 = fw.y * dy;
            a[4] = fmaf(s0, dx, a[4]); a[5] = fmaf(s0, dy, a[5]); a[6] = fmaf(s1, dy, a[6]);
            const float gx = fmaf(ca, s0, cb * s1), gy = fmaf(cb, s0, cc * s1);
            a[7] += gx; a[8] += gy; a[9] += fabsf(gx); a[10] += fabsf(gy);
            a[11] += fw.y;
        }
#pragma unroll
        for (int k = 0; k < NPART; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], 16);
        if (half == 0 && active) {
            const float o = m0.z;
            const int64_t entry = (int64_t)ebase + __float_as_uint(m1.w);    // the warp's own entry of this pair
            float4* dst = reinterpret_cast<float4*>(partials + entry * ENTRY_F);
            dst[0] = make_float4(a[0], a[1], a[2], a[3]);
            dst[1] = make_float4(-0.5f * o * a[4], -o * a[5], -0.5f * o * a[6], -o * a[7]);
            dst[2] = make_float4(-o * a[8], o * a[9], o * a[10], a[11]);
            touched[entry] = 1;
        }
        gcnt = 0;
        __syncwarp();
    };

    for (int b = 0; b < num_batches; ++b) {
        const int st = b % NSTAGE;
        const int batch_hi = hi_end - BB * b;
        const int batch_lo = max(range_start, batch_hi - BB);
        const int n = batch_hi - batch_lo;            // staged slot t <-> sorted index batch_lo + t
        stage_wait(&S.full[st], (b / NSTAGE) & 1);
        const float4* rec = S.rec[st];
        // this warp's candidates, back to front: alpha box reaches its 8x4 block, and the Gaussian is not behind
        // everything the block's pixels blended
        int cnt = 0;
        const int wrel = wmax - batch_lo;          // staged slots above this were blended by no pixel of the warp
        if (wrel >= 0) {
#pragma unroll
            for (int chunk = BB / 32 - 1; chunk >= 0; --chunk) {
                const int t = chunk * 32 + 31 - lane;
                const bool c = t < n && t <= wrel && ((__float_as_uint(rec[t * REC_F4 + 2].z) >> warp) & 1u);
                const uint32_t word = __ballot_sync(0xffffffffu, c);
                if (c) S.list[warp][cnt + __popc(word & ((1u << lane) - 1u))] = (uint8_t)t;
                cnt += __popc(word);
            }
            __syncwarp();
        }
        const int brel = bin_final - batch_lo;     // the pixel blended staged slots <= brel
        for (int q = 0; q < cnt; ++q) {
            // ---- phase 1 (lane = pixel): alpha, transmittance, the two scalars of this (pixel, Gaussian) pair
            const int t = S.list[warp][q];
            const float4 g0 = rec[t * REC_F4];
            const float4 g1 = rec[t * REC_F4 + 1];
            const float dx = g0.x - px, dy = g0.y - py;
            const float pw = fmaf(g1.y * dy, dy, fmaf(g1.x, dy, g0.w * dx) * dx);   // -sigma * log2(e)
            const float e = ex2_approx(pw);
            const float araw = g0.z * e;
            const float alpha = fminf(cfg.alpha_max, araw);
            const bool valid = t <= brel && !(pw > 0.f) && alpha >= ALPHA_MIN;
            if (COUNT) { cnt_eval += lane == 0; cnt_pairs += valid; }
            if (!__any_sync(0xffffffffu, valid)) continue;
            if (COUNT) cnt_hit += lane == 0;
            const float4 g2 = rec[t * REC_F4 + 2];
            const float D = fmaf(g1.z, v_c[0], fmaf(g1.w, v_c[1], fmaf(g2.x, v_c[2], g2.y * v_c[3])));
            const float ra = __fdividef(1.0f, 1.0f - alpha);   // 1 - alpha in [1e-3, 1]: the fast reciprocal is safe
            const float Ti = T * ra;                            // transmittance in front of this Gaussian
            const float fac = alpha * Ti;
            const float v_alpha = fmaf(Ti, D, -ra * E);
            const float w = araw <= cfg.alpha_max ? e * v_alpha : 0.f;   // capped alpha: no gradient through it
            if (valid) { T = Ti; E = fmaf(fac, D, E); }
            S.M[warp][gcnt][lane] = valid ? make_float2(fac, w) : make_float2(0.f, 0.f);
            if (lane == 0) {
                S.grp[warp][gcnt][0] = g0;
                const uint32_t rank = __popc(__float_as_uint(g2.z) & ((1u << warp) - 1u));   // of this warp among the pair's
                S.grp[warp][gcnt][1] = make_float4(g1.x, g1.y, g2.w, __uint_as_float(rank));

            }
            if (++gcnt == GRP) phase2();
        }
        // done with this stage: the last of the CTA's 8 warps to get here refills it
        __syncwarp();
        if (lane == 0) {
            const int old = atomicAdd(&S.done[st], 1);
            if (old == NWARP - 1) {
                S.done[st] = 0;
                const int nb = b + NSTAGE;
                if (nb < num_batches) {
                    const int bh = hi_end - BB * nb, bl = max(range_start, bh - BB);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    stage_issue(srecs, bl, bh - bl, S.rec[st], &S.full[st]);
                }
            }
        }
    }
    if (gcnt > 0) phase2();
    if (COUNT) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            cnt_eval += __shfl_xor_sync(0xffffffffu, cnt_eval, o);
            cnt_hit += __shfl_xor_sync(0xffffffffu, cnt_hit, o);
            cnt_pairs += __shfl_xor_sync(0xffffffffu, cnt_pairs, o);
            cnt_staged += __shfl_xor_sync(0xffffffffu, cnt_staged, o);
            cnt_groups += __shfl_xor_sync(0xffffffffu, cnt_groups, o);
            cnt_members += __shfl_xor_sync(0xffffffffu, cnt_members, o);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, (unsigned long long)cnt_eval);
            atomicAdd(counters + 1, (unsigned long long)cnt_hit);
            atomicAdd(counters + 2, (unsigned long long)cnt_pairs);
            atomicAdd(counters + 3, (unsigned long long)cnt_staged);
            atomicAdd(counters + 4, (unsigned long long)cnt_groups);
            atomicAdd(counters + 5, (unsigned long long)cnt_members);
        }
    }
}

// Heavy-first tile order + segment bookkeeping.  One CTA.
//   order[r]      : tile ids, bucketed by floor(log2(list length)), longest bucket first (coarse LPT; the order
//                   within a bucket is arbitrary -- it only affects scheduling, never results)
//   seg_prefix[r] : inclusive count of 1024-Gaussian segments of tiles order[0..r] (>= 1 per tile)
//   ckpt_base[t]  : first checkpoint / segment-output slot of TILE t (a tile with n > 1 segments owns n slots,
//                   a single-segment tile none)
//   cta_map[c]    : (tile id, segment | segments of the tile << 16) of compositing CTA c -- what the forward / backward
//                   CTAs read instead of searching seg_prefix
__global__ void __launch_bounds__(1024) tile_order_kernel(const int32_t* __restrict__ tile_offsets, int64_t P,
                                                          int n_cam_tiles, int32_t* __restrict__ order,
                                                          int32_t* __restrict__ seg_prefix,
                                                          int32_t* __restrict__ ckpt_base, int2* __restrict__ cta_map) {
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
            if (cta_map) {
                const int first = (int)(excl & 0xffffffffll);
                for (int sg = 0; sg < nseg; ++sg) cta_map[first + sg] = make_int2(t, sg | (nseg << 16));
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = incl;
        __syncthreads();
    }
}

// Sum each Gaussian's contiguous run of gradient entries (tiles of its rectangle in emission order x the warps whose
// block its alpha box reaches, warp order) -> dense per-(camera,Gaussian) grads.  FOUR lanes per Gaussian: lane q of the
// quad loads float4 #q of every 64-byte entry (one full line per quad and load) and sums components 4q .. 4q+3, so no
// cross-lane reduction is needed; runs longer than 64 entries (a Gaussian that covers many tiles) are spread over the
// warp's 8 quads and added in a fixed tree.  Deterministic either way.
__global__ void __launch_bounds__(256) raster_gather_kernel(
    const float* __restrict__ partials, const uint8_t* __restrict__ touched, const int64_t* __restrict__ cum_tiles,
    const uint32_t* __restrict__ entry_base, int64_t CN, int d_color, int with_depth, float* __restrict__ v_means2d,
    float* __restrict__ v_means2d_abs, float* __restrict__ v_conics, float* __restrict__ v_colors,
    float* __restrict__ v_depths, float* __restrict__ v_opacities) {
    const int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int lane = threadIdx.x & 31, q = lane & 3, quad = lane >> 2;
    const bool live = ci < CN;
    int64_t e0 = 0;
    int n = 0;
    if (live) {
        const int64_t lo = ci == 0 ? 0 : cum_tiles[ci - 1], hi = cum_tiles[ci];
        if (hi > lo) {
            e0 = entry_base[lo];
            n = (int)(entry_base[hi] - entry_base[lo]);   // entry_base holds P + 1 offsets
        }
    }
    const float4* ent = reinterpret_cast<const float4*>(partials) + q;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int LONG_RUN = 64;
    if (n <= LONG_RUN) {
        for (int e = 0; e < n; ++e) {
            if (touched[e0 + e]) {
                const float4 v = __ldg(ent + (e0 + e) * (ENTRY_F / 4));
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
        }
    }
    uint32_t longs = __ballot_sync(0xffffffffu, n > LONG_RUN && q == 0);
    while (longs) {
        const int src = __ffs(longs) - 1;    // lane 0 of the quad that owns the long run
        longs &= longs - 1;
        const int64_t be0 = __shfl_sync(0xffffffffu, e0, src);
        const int bn = __shfl_sync(0xffffffffu, n, src);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = quad; e < bn; e += 8) {
            if (touched[be0 + e]) {
                const float4 v = __ldg(ent + (be0 + e) * (ENTRY_F / 4));
                b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
            }
        }
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
            b.x += __shfl_xor_sync(0xffffffffu, b.x, o); b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
            b.z += __shfl_xor_sync(0xffffffffu, b.z, o); b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
        }
        if ((lane >> 2) == (src >> 2)) a = b;
    }
    if (!live) return;
    if (q == 0) {          // components 0..3: the channels (colours, then the depth channel)
        const float ch[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < d_color) v_colors[ci * d_color + k] = ch[k];
            if (with_depth && k == d_color) v_depths[ci] = ch[k];
        }
    } else if (q == 1) {   // 4..7: conic a, b, c, mean x
        v_conics[ci * 3 + 0] = a.x; v_conics[ci * 3 + 1] = a.y; v_conics[ci * 3 + 2] = a.z;
        v_means2d[ci * 2 + 0] = a.w;
    } else if (q == 2) {   // 8..11: mean y, |mean x|, |mean y|, opacity
        v_means2d[ci * 2 + 1] = a.x;
        if (v_means2d_abs) { v_means2d_abs[ci * 2 + 0] = a.y; v_means2d_abs[ci * 2 + 1] = a.z; }
        v_opacities[ci] = a.w;
    }
}

}  // namespace

static RasterCfg raster_cfg(int flavour) {
    return flavour == 1 ? RasterCfg{0.0f, 0.99f, 1, 1} : RasterCfg{0.5f, 0.999f, 0, 0};
}

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

// recs: emd_raster_pack's [C*N][12] floats; isect_ids / flatten_ids: the SORTED keys and values; srecs: [P][12] floats out.
// cand (may be NULL when no backward will follow): [P] uint8, per emission-order slot the number of gradient entries of
// the pair; the caller turns it into the (P+1)-long entry_base with emd_exclusive_scan_u8_u32(cand, entry_base, P,
// entry_base + P, ...).
// tile_n_bits: width of the tile field above bit 32 of a key (gsplat: floor(log2(tiles)) + 1; camera bits above it are
// ignored -- the camera comes from flatten_ids / N).
extern "C" int emd_raster_sort_records(const float* recs, const int64_t* isect_ids, const int32_t* flatten_ids,
                                       const int32_t* radii, const int64_t* cum_tiles, int64_t P, int tile_w, int tile_h,
                                       int tile_n_bits, int flavour, float* srecs, uint8_t* cand, cudaStream_t stream) {
    EMD_CHECK_ARG(P >= 0 && P < ((int64_t)1 << 31), "raster_sort_records: too many intersections");
    EMD_CHECK_ARG(tile_n_bits >= 1 && tile_n_bits <= 30, "raster_sort_records: bad tile_n_bits");
    if (!emd_aligned(recs, 16) || !emd_aligned(srecs, 16)) {
        emd_set_error("raster_sort_records: recs / srecs must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    if (P == 0) return EMD_OK;
    EMD_LAUNCH(EK_RASTER_PACK, stream, raster_sort_records_kernel<<<(unsigned)emd_cdiv(P, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(recs), isect_ids, flatten_ids, radii, cum_tiles, P, tile_w, tile_h,
        tile_n_bits, raster_cfg(flavour), reinterpret_cast<float4*>(srecs), cand));
    EMD_CHECK_LAUNCH("raster_sort_records");
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
// cta_map: [emd_raster_max_ctas(P, n_cam_tiles)] int2 (may be NULL when only the order is wanted).
extern "C" int64_t emd_raster_max_ctas(int64_t P, int64_t n_cam_tiles) { return P / SEG + n_cam_tiles; }

extern "C" int emd_tile_order(const int32_t* tile_offsets, int64_t P, int64_t n_cam_tiles, int32_t* order,
                              int32_t* seg_prefix, int32_t* ckpt_base, int32_t* cta_map, cudaStream_t stream) {
    EMD_CHECK_ARG(n_cam_tiles >= 1 && n_cam_tiles < (1 << 30), "tile_order: bad tile count");
    EMD_CHECK_ARG((seg_prefix == nullptr) == (ckpt_base == nullptr), "tile_order: seg_prefix and ckpt_base go together");
    EMD_CHECK_ARG(cta_map == nullptr || seg_prefix != nullptr, "tile_order: cta_map needs seg_prefix / ckpt_base");
    EMD_LAUNCH(EK_MISC, stream, tile_order_kernel<<<1, 1024, 0, stream>>>(tile_offsets, P, (int)n_cam_tiles, order, seg_prefix, ckpt_base,
                                                                            reinterpret_cast<int2*>(cta_map)));
    EMD_CHECK_LAUNCH("tile_order");
    return EMD_OK;
}

// srecs: the sorted record stream of emd_raster_sort_records.
extern "C" int emd_rasterize_fwd(const float* srecs, const int32_t* tile_offsets,
                                 const int32_t* tile_order, const int32_t* seg_prefix, const int32_t* ckpt_base,
                                 const int32_t* cta_map,
                                 int64_t P, int64_t C, int width, int height, int tile_w, int tile_h, int channels,
                                 int ed_mode, int flavour, const float* backgrounds, float* ckpt, float* seg_out,
                                 float* out_colors, float* out_alphas, int32_t* last_ids, cudaStream_t stream) {
    const RasterCfg cfg = raster_cfg(flavour);
    EMD_CHECK_ARG(channels >= 1 && channels <= 4, "rasterize_fwd: channels must be 1..4");
    EMD_CHECK_ARG(C >= 1 && C * tile_w * tile_h < ((int64_t)1 << 30), "rasterize_fwd: grid too large");
    EMD_CHECK_ARG(tile_w == (width + EMD_TILE - 1) / EMD_TILE && tile_h == (height + EMD_TILE - 1) / EMD_TILE,
                  "rasterize_fwd: tile grid does not match image size (tile size is 16)");
    EMD_CHECK_ARG(tile_order && seg_prefix && ckpt_base && cta_map && ckpt && seg_out,
                  "rasterize_fwd: needs tile_order, seg_prefix, ckpt_base, cta_map and the ckpt / seg_out buffers");
    EMD_CHECK_ARG(P >= 0 && P < ((int64_t)1 << 31), "rasterize_fwd: too many intersections");
    if (!emd_aligned(srecs, 16) || (channels == 4 && !emd_aligned(out_colors, 16))) {
        emd_set_error("rasterize_fwd: srecs/out_colors must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int n_ct = (int)(C * tile_w * tile_h);
    // upper bound on the number of (tile, segment) CTAs; surplus CTAs exit at once
    dim3 grid((unsigned)(P / SEG + n_ct)), block(RB, 1, 1);
    const float4* r4 = reinterpret_cast<const float4*>(srecs);
    if (P > SEG)  // only then can a tile have more than one segment
        EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_seg_kernel<false><<<grid, block, 0, stream>>>(
            r4, tile_offsets, reinterpret_cast<const int2*>(cta_map), seg_prefix, ckpt_base, P, (int)C, width, height, tile_w, tile_h,
            channels, ed_mode, cfg, backgrounds, ckpt, seg_out, out_colors, out_alphas, last_ids));
    EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_seg_kernel<true><<<grid, block, 0, stream>>>(
        r4, tile_offsets, reinterpret_cast<const int2*>(cta_map), seg_prefix, ckpt_base, P, (int)C, width, height, tile_w, tile_h,
        channels, ed_mode, cfg, backgrounds, ckpt, seg_out, out_colors, out_alphas, last_ids));
    if (P > SEG)
        EMD_LAUNCH(EK_RASTER_FWD, stream, raster_fwd_combine_kernel<<<n_ct, block, 0, stream>>>(
            tile_order, seg_prefix, ckpt_base, (int)C, width, height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds,
            ckpt, seg_out, out_colors, out_alphas, last_ids));
    EMD_CHECK_LAUNCH("rasterize_fwd");
    return EMD_OK;
}

// n_entries = entry_base[P] (the scan's total, read back by the caller; <= 8 P): [n_entries][16] float gradient entries
// followed by [n_entries] touched flags
static size_t bwd_entries_bytes(int64_t n_entries) { return ((size_t)n_entries * ENTRY_F * sizeof(float) + 255) / 256 * 256; }
extern "C" size_t emd_rasterize_bwd_workspace_bytes(int64_t n_entries) {
    return bwd_entries_bytes(n_entries) + ((size_t)n_entries + 255) / 256 * 256 + 256;
}

extern "C" int emd_rasterize_bwd(const float* srecs, const int32_t* tile_offsets, const int32_t* cta_map,
                                 const int64_t* cum_tiles, const uint32_t* entry_base, int64_t n_entries, int64_t P, int64_t N, int64_t C,
                                 int width, int height, int tile_w, int tile_h, int channels, int ed_mode, int flavour,
                                 const float* backgrounds, const int32_t* seg_prefix, const int32_t* ckpt_base,
                                 const float* ckpt, const float* out_colors, const float* out_alphas,
                                 const int32_t* last_ids, const float* v_out_colors, const float* v_out_alphas,
                                 int d_color, int with_depth, float* v_means2d, float* v_means2d_abs, float* v_conics,
                                 float* v_colors, float* v_depths, float* v_opacities, void* workspace,
                                 size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(channels >= 1 && channels <= 4, "rasterize_bwd: channels must be 1..4");
    EMD_CHECK_ARG(d_color + (with_depth ? 1 : 0) == channels, "rasterize_bwd: channel bookkeeping mismatch");
    const RasterCfg cfg = raster_cfg(flavour);
    EMD_CHECK_ARG(P >= 0 && P < ((int64_t)1 << 31), "rasterize_bwd: too many intersections");
    EMD_CHECK_ARG(cta_map && seg_prefix && ckpt_base && ckpt, "rasterize_bwd: needs cta_map, seg_prefix, ckpt_base, ckpt");
    EMD_CHECK_ARG(n_entries >= 0 && n_entries <= P * NWARP, "rasterize_bwd: n_entries must be entry_base[P]");
    if (ws_bytes < emd_rasterize_bwd_workspace_bytes(n_entries)) {
        emd_set_error("rasterize_bwd: workspace too small");
        return EMD_ERR_WORKSPACE;
    }
    if (!emd_aligned(workspace, 16) || !emd_aligned(srecs, 16)) {
        emd_set_error("rasterize_bwd: workspace/srecs must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t CN = C * N;
    if (CN == 0) return EMD_OK;
    float* partials = reinterpret_cast<float*>(workspace);
    uint8_t* touched = reinterpret_cast<uint8_t*>(workspace) + bwd_entries_bytes(n_entries);
    if (P > 0) {
        static std::once_flag once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [] {
            attr_err = cudaFuncSetAttribute(raster_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
            if (attr_err == cudaSuccess)
                attr_err = cudaFuncSetAttribute(raster_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
        });
        if (attr_err != cudaSuccess) {
            emd_set_error("rasterize_bwd: cannot reserve %zu bytes of shared memory: %s", sizeof(BwdSmem), cudaGetErrorString(attr_err));
            return EMD_ERR_CUDA;
        }
        cudaMemsetAsync(touched, 0, (size_t)n_entries, stream);
        // upper bound on the number of (tile, segment) CTAs; surplus CTAs exit at once
        dim3 grid((unsigned)(P / SEG + C * tile_w * tile_h)), block(RB, 1, 1);
        const float4* r4 = reinterpret_cast<const float4*>(srecs);
        if (g_raster_counters)
            EMD_LAUNCH(EK_RASTER_BWD, stream, raster_bwd_kernel<true><<<grid, block, sizeof(BwdSmem), stream>>>(
                r4, tile_offsets, reinterpret_cast<const int2*>(cta_map), P, (int)C, width, height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds,
                seg_prefix, ckpt_base, ckpt, out_colors, out_alphas, last_ids, v_out_colors, v_out_alphas, entry_base, partials,
                touched, g_raster_counters));
        else
            EMD_LAUNCH(EK_RASTER_BWD, stream, raster_bwd_kernel<false><<<grid, block, sizeof(BwdSmem), stream>>>(
                r4, tile_offsets, reinterpret_cast<const int2*>(cta_map), P, (int)C, width, height, tile_w, tile_h, channels, ed_mode, cfg, backgrounds,
                seg_prefix, ckpt_base, ckpt, out_colors, out_alphas, last_ids, v_out_colors, v_out_alphas, entry_base, partials,
                touched, nullptr));
    }
    EMD_LAUNCH(EK_RASTER_GATHER, stream, raster_gather_kernel<<<(unsigned)emd_cdiv(CN * 4, 256), 256, 0, stream>>>(partials, touched, cum_tiles, entry_base, CN, d_color,
                                                                           with_depth, v_means2d, v_means2d_abs,
                                                                           v_conics, v_colors, v_depths, v_opacities));
    EMD_CHECK_LAUNCH("rasterize_bwd");
    return EMD_OK;
}
