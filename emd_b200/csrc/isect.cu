// K3 / K5: tile-intersection bookkeeping.
//   * device-wide scans (tiles-per-Gaussian -> cumulative offsets, radix tables)
//   * emission of (64-bit key, 32-bit value) pairs, bit-exact with gsplat's
//     isect_tiles:  key = cam << (32+tile_bits) | tile << 32 | float_bits(depth)
//   * tile offsets from the sorted keys (isect_offset_encode)
#include "common.cuh"
#include "proj_math.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 elements per block

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += n;
    }
    return v;
}

// inclusive block scan of one value per thread; returns inclusive prefix, *total = block sum
template <typename T, int THREADS>
__device__ __forceinline__ T block_incl_scan(T v, T* s_warp /*[THREADS/32]*/, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < THREADS / 32 ? s_warp[lane] : T(0);
        T wi = warp_incl_scan(w);
        if (lane < THREADS / 32) s_warp[lane] = wi;
    }
    __syncthreads();
    const T base = warp > 0 ? s_warp[warp - 1] : T(0);
    *total = s_warp[THREADS / 32 - 1];
    __syncthreads();
    return inc + base;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const TIn* __restrict__ in, int64_t n,
                                                                   TOut* __restrict__ block_sums) {
    __shared__ TOut s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += (TOut)in[i];
    }
    TOut total;
    block_incl_scan<TOut, SCAN_THREADS>(acc, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place; writes grand total
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(TOut* __restrict__ block_sums, int64_t nb,
                                                                  TOut* __restrict__ total_out) {
    __shared__ TOut s_warp[SCAN_THREADS / 32];
    TOut carry = 0;
    for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const TOut v = i < nb ? block_sums[i] : TOut(0);
        TOut total;
        const TOut inc = block_incl_scan<TOut, SCAN_THREADS>(v, s_warp, &total);
        if (i < nb) block_sums[i] = carry + inc - v;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// per block: rescan with carry-in; thread owns SCAN_ITEMS consecutive elements
template <typename TIn, typename TOut, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const TIn* __restrict__ in, int64_t n,
                                                                 const TOut* __restrict__ block_sums,
                                                                 TOut* __restrict__ out) {
    __shared__ TOut s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? (TOut)in[i] : TOut(0);
        acc += v[k];
    }
    TOut total;
    const TOut inc = block_incl_scan<TOut, SCAN_THREADS>(acc, s_warp, &total);
    TOut run = block_sums[blockIdx.x] + inc - acc;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        if (INCLUSIVE) run += v[k];
        if (i < n) out[i] = run;
        if (!INCLUSIVE) run += v[k];
    }
}

template <typename TIn, typename TOut, bool INCLUSIVE>
int scan_impl(const TIn* in, TOut* out, int64_t n, TOut* total_out, void* workspace, size_t ws_bytes,
              cudaStream_t stream) {
    const int64_t nb = emd_cdiv(n, SCAN_TILE);
    if ((size_t)(nb > 0 ? nb : 1) * sizeof(TOut) > ws_bytes) {
        emd_set_error("scan: workspace too small (%zu < %zu)", ws_bytes, (size_t)nb * sizeof(TOut));
        return EMD_ERR_WORKSPACE;
    }
    TOut* sums = reinterpret_cast<TOut*>(workspace);
    if (n == 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(TOut), stream);
        return EMD_OK;
    }
    EMD_LAUNCH(EK_SCAN, stream, scan_reduce_kernel<TIn, TOut><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, sums));
    EMD_LAUNCH(EK_SCAN, stream, scan_spine_kernel<TOut><<<1, SCAN_THREADS, 0, stream>>>(sums, nb, total_out));
    EMD_LAUNCH(EK_SCAN, stream, scan_down_kernel<TIn, TOut, INCLUSIVE><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, sums, out));
    EMD_CHECK_LAUNCH("scan");
    return EMD_OK;
}

// ---------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS) isect_emit_kernel(
    const float* __restrict__ means2d, const int32_t* __restrict__ radii, const float* __restrict__ depths,
    const int64_t* __restrict__ cum_tiles, int64_t N, int64_t CN, int tile_w, int tile_h, int tile_n_bits,
    int64_t* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
    const int64_t ci = (int64_t)blockIdx.x * EMIT_THREADS + threadIdx.x;
    if (ci >= CN) return;
    const int r = radii[ci];
    if (r <= 0) return;
    const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + ci);
    int x0, y0, x1, y1;
    tile_rect_c(m.x, m.y, r, tile_w, tile_h, x0, y0, x1, y1);
    if (x1 <= x0 || y1 <= y0) return;
    const int64_t cam = ci / N;
    const int64_t hi = cam << (32 + tile_n_bits);
    const int64_t lo = (int64_t)(uint32_t)__float_as_int(depths[ci]);
    int64_t cur = ci == 0 ? 0 : cum_tiles[ci - 1];
    for (int ty = y0; ty < y1; ++ty) {
        for (int tx = x0; tx < x1; ++tx) {
            const int64_t tile = (int64_t)ty * tile_w + tx;
            isect_ids[cur] = hi | (tile << 32) | lo;
            flatten_ids[cur] = (int32_t)ci;
            ++cur;
        }
    }
}

__global__ void isect_offsets_kernel(const int64_t* __restrict__ sorted_ids, int64_t P, int64_t n_cam_tiles,
                                     int64_t n_tiles, int tile_n_bits, int32_t* __restrict__ offsets) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (P == 0) {
        if (idx < n_cam_tiles) offsets[idx] = 0;
        return;
    }
    if (idx >= P) return;
    const int64_t tile_mask = ((int64_t)1 << tile_n_bits) - 1;
    const int64_t hc = sorted_ids[idx] >> 32;
    const int64_t id_cur = (hc >> tile_n_bits) * n_tiles + (hc & tile_mask);
    if (idx == 0) {
        for (int64_t t = 0; t <= id_cur; ++t) offsets[t] = 0;
    } else {
        const int64_t hp = sorted_ids[idx - 1] >> 32;
        const int64_t id_prev = (hp >> tile_n_bits) * n_tiles + (hp & tile_mask);
        for (int64_t t = id_prev + 1; t <= id_cur; ++t) offsets[t] = (int32_t)idx;
    }
    if (idx == P - 1) {
        for (int64_t t = id_cur + 1; t < n_cam_tiles; ++t) offsets[t] = (int32_t)P;
    }
}

}  // namespace

extern "C" size_t emd_scan_workspace_bytes(int64_t n) {
    const int64_t nb = emd_cdiv(n > 0 ? n : 1, SCAN_TILE);
    return (size_t)nb * sizeof(int64_t) + 256;
}

// inclusive cumsum int32 -> int64 (gsplat: torch.cumsum(tiles_per_gauss)); total -> device scalar
extern "C" int emd_cumsum_i32_i64(const int32_t* in, int64_t* out, int64_t n, int64_t* total_out, void* workspace,
                                  size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(n >= 0, "cumsum: negative n");
    return scan_impl<int32_t, int64_t, true>(in, out, n, total_out, workspace, ws_bytes, stream);
}

extern "C" int emd_exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* workspace, size_t ws_bytes,
                                      cudaStream_t stream) {
    EMD_CHECK_ARG(n >= 0, "exclusive_scan: negative n");
    return scan_impl<uint32_t, uint32_t, false>(in, out, n, (uint32_t*)nullptr, workspace, ws_bytes, stream);
}

// exclusive scan uint8 -> uint32 with the grand total written to *total_out (a DEVICE pointer; out + n is a valid
// choice, which makes `out` an (n+1)-long offsets array)
extern "C" int emd_exclusive_scan_u8_u32(const uint8_t* in, uint32_t* out, int64_t n, uint32_t* total_out, void* workspace,
                                         size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(n >= 0, "exclusive_scan: negative n");
    return scan_impl<uint8_t, uint32_t, false>(in, out, n, total_out, workspace, ws_bytes, stream);
}

extern "C" int emd_isect_emit(const float* means2d, const int32_t* radii, const float* depths,
                              const int64_t* cum_tiles, int64_t N, int64_t C, int tile_w, int tile_h,
                              int tile_n_bits, int64_t* isect_ids, int32_t* flatten_ids, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && C >= 1, "isect_emit: bad sizes");
    EMD_CHECK_ARG(tile_n_bits >= 1 && tile_n_bits < 31, "isect_emit: bad tile_n_bits %d", tile_n_bits);
    const int64_t CN = C * N;
    if (CN == 0) return EMD_OK;
    EMD_LAUNCH(EK_ISECT_EMIT, stream, isect_emit_kernel<<<(unsigned)emd_cdiv(CN, EMIT_THREADS), EMIT_THREADS, 0, stream>>>(
        means2d, radii, depths, cum_tiles, N, CN, tile_w, tile_h, tile_n_bits, isect_ids, flatten_ids));
    EMD_CHECK_LAUNCH("isect_emit");
    return EMD_OK;
}

extern "C" int emd_isect_offsets(const int64_t* sorted_ids, int64_t P, int64_t C, int tile_w, int tile_h,
                                 int tile_n_bits, int32_t* offsets, cudaStream_t stream) {
    const int64_t n_tiles = (int64_t)tile_w * tile_h;
    const int64_t n_cam_tiles = C * n_tiles;
    EMD_CHECK_ARG(P >= 0 && n_cam_tiles > 0, "isect_offsets: bad sizes");
    const int64_t threads = P > 0 ? P : n_cam_tiles;
    EMD_LAUNCH(EK_ISECT_OFFSETS, stream, isect_offsets_kernel<<<(unsigned)emd_cdiv(threads, 256), 256, 0, stream>>>(sorted_ids, P, n_cam_tiles, n_tiles,
                                                                                tile_n_bits, offsets));
    EMD_CHECK_LAUNCH("isect_offsets");
    return EMD_OK;
}
