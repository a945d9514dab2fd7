// K4: stable LSD radix sort of (uint64 key, uint32 value) pairs over a bit range.
//
// Replaces cub::DeviceRadixSort::SortPairs as gsplat calls it from
// isect_tiles (reference call site OmniRe/models/trainers/base.py:393).
// 8-bit digits.  Per pass: (A) per-block digit histogram into a digit-major
// table, (B) device-wide exclusive scan of the table, (C) stable scatter: each
// warp ranks its 512 consecutive keys with match.any + per-warp counters, the
// block turns the counters into offsets, and every key goes to
// table[digit][block] + warp_offset + rank.  HBM-bound: 8 B (A) + 12 B read +
// 12 B written (C) per pair per pass.
#include "common.cuh"

extern "C" int emd_exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* workspace, size_t ws_bytes,
                                      cudaStream_t stream);
extern "C" size_t emd_scan_workspace_bytes(int64_t n);

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per block
constexpr int RS_WARP_KEYS = 32 * RS_ITEMS;     // 512 consecutive keys per warp
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                             uint32_t mask, uint32_t* __restrict__ table,
                                                             int64_t nblocks) {
    __shared__ uint32_t hist[RS_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(__ldg(keys + i) >> shift) & mask], 1u);
    }
    __syncthreads();
    table[(int64_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, int shift, uint32_t mask, const uint32_t* __restrict__ table,
    int64_t nblocks) {
    __shared__ uint32_t whist[RS_WARPS][RS_BINS];
    __shared__ uint32_t dig_base[RS_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int w = 0; w < RS_WARPS; ++w) whist[w][threadIdx.x] = 0;
    dig_base[threadIdx.x] = table[(int64_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WARP_KEYS;
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        key[r] = i < n ? __ldg(keys_in + i) : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = ok ? ((uint32_t)(key[r] >> shift) & mask) : 0xFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (ok && lane == leader) {
            prev = whist[warp][d];
            whist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = prev + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = whist[w][threadIdx.x];
            whist[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const int64_t pos = (int64_t)dig_base[d] + whist[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = __ldg(vals_in + i);
        }
    }
}

}  // namespace

extern "C" size_t emd_radix_sort_workspace_bytes(int64_t n) {
    const int64_t nb = emd_cdiv(n > 0 ? n : 1, RS_TILE);
    const size_t table = (size_t)nb * RS_BINS * sizeof(uint32_t);
    return ((table + 255) / 256) * 256 + emd_scan_workspace_bytes(nb * RS_BINS);
}

// Sorts bits [begin_bit, end_bit) ascending, stable.  Ping-pongs between buffer
// 0 (keys0/vals0, the input) and buffer 1; *result_buffer says where the
// result landed.
extern "C" int emd_radix_sort_pairs(uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1, int64_t n,
                                    int begin_bit, int end_bit, void* workspace, size_t ws_bytes, int* result_buffer,
                                    cudaStream_t stream) {
    EMD_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 32), "radix_sort: n out of range");
    EMD_CHECK_ARG(begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "radix_sort: bad bit range");
    if (result_buffer) *result_buffer = 0;
    if (n == 0 || begin_bit == end_bit) return EMD_OK;
    if (ws_bytes < emd_radix_sort_workspace_bytes(n)) {
        emd_set_error("radix_sort: workspace too small");
        return EMD_ERR_WORKSPACE;
    }
    const int64_t nb = emd_cdiv(n, RS_TILE);
    uint32_t* table = reinterpret_cast<uint32_t*>(workspace);
    const size_t table_bytes = (((size_t)nb * RS_BINS * sizeof(uint32_t) + 255) / 256) * 256;
    void* scan_ws = reinterpret_cast<char*>(workspace) + table_bytes;
    const size_t scan_ws_bytes = ws_bytes - table_bytes;
    uint64_t* kin = keys0;  uint32_t* vin = vals0;
    uint64_t* kout = keys1; uint32_t* vout = vals1;
    int cur = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        const uint32_t mask = (1u << bits) - 1u;
        EMD_LAUNCH(EK_SORT_HIST, stream, rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, stream>>>(kin, n, shift, mask, table, nb));
        int rc = emd_exclusive_scan_u32(table, table, nb * RS_BINS, scan_ws, scan_ws_bytes, stream);
        if (rc != EMD_OK) return rc;
        EMD_LAUNCH(EK_SORT_SCATTER, stream, rs_scatter_kernel<<<(unsigned)nb, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, mask, table, nb));
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
        cur ^= 1;
    }
    EMD_CHECK_LAUNCH("radix_sort");
    if (result_buffer) *result_buffer = cur;
    return EMD_OK;
}
