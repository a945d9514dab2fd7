// K4: stable LSD radix sort of (uint64 key, uint32 value) pairs over a bit range.
//
// Replaces cub::DeviceRadixSort::SortPairs as gsplat calls it from
// isect_tiles (reference call site OmniRe/models/trainers/base.py:393).
// 8-bit digits, one kernel per pass ("onesweep"): the digit histograms of ALL
// passes are taken in one read of the keys up front; a pass then ranks its
// 4096-key tile (match.any + per-warp counters), resolves the tile's global
// offsets by decoupled look-back over the tiles before it (tiles are numbered by
// an atomic ticket, so every predecessor is already running), reorders the tile
// by digit in shared memory and writes each digit run coalesced.
// HBM-bound: 8 B once + 12 B read + 12 B written per pair per pass.
// (n >= 2^30 falls back to the three-kernel histogram / scan / scatter pass.)
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

// Lanes holding the same digit, by ballots over the digit's bits (8 fast VOTEs).  The match.any instruction
// is far slower than this on sm_100: with it the ranking loop alone bounded a pass at ~55 us for 4.3 M keys.
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool ok) {
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
    if (!ok) peers = ~peers;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool set = (d >> bit) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, set);
        peers &= set ? m : ~m;
    }
    return peers;
}

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
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const uint32_t peers = match_digit(d, ok);
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


// ---- onesweep ------------------------------------------------------------------------------------
constexpr int OS_MAX_PASSES = 8;
constexpr int OS_ITEMS = 8;                      // keys per thread: 2048-key tiles, <= 64 registers -> 4 CTAs per SM
constexpr int OS_TILE = RS_THREADS * OS_ITEMS;
constexpr int OS_WARP_KEYS = 32 * OS_ITEMS;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VAL_MASK = (1u << 30) - 1u;
struct OsSmem {
    uint64_t keys[OS_TILE];
    uint32_t vals[OS_TILE];
    uint32_t whist[RS_WARPS][RS_BINS];
    uint32_t lstart[RS_BINS];   // first position of digit d inside the sorted tile
    uint32_t gbase[RS_BINS];    // global position of sorted-tile slot i with digit d = gbase[d] + i  (wraps mod 2^32)
    uint32_t wsum[RS_WARPS];
    uint32_t bid;
};

// histograms of every pass in one read of the keys: ghist[p][d]
__global__ void __launch_bounds__(RS_THREADS) rs_multi_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                                   int begin_bit, int end_bit, int npasses,
                                                                   uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[OS_MAX_PASSES][RS_BINS];
    for (int p = 0; p < npasses; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        const uint64_t k = __ldg(keys + i);
        for (int p = 0; p < npasses; ++p) {
            const int shift = begin_bit + 8 * p;
            const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
            atomicAdd(&h[p][(uint32_t)(k >> shift) & ((1u << bits) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < npasses; ++p)
        if (h[p][threadIdx.x]) atomicAdd(&ghist[p * RS_BINS + threadIdx.x], h[p][threadIdx.x]);
}

// exclusive prefix of each pass's 256 bins, in place (grid = npasses)
__global__ void __launch_bounds__(RS_BINS) rs_digit_prefix_kernel(uint32_t* __restrict__ ghist) {
    __shared__ uint32_t s_w[RS_BINS / 32];
    uint32_t* h = ghist + blockIdx.x * RS_BINS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_w[w];
    h[threadIdx.x] = base + inc - v;
}

__global__ void __launch_bounds__(RS_THREADS, 4) rs_onesweep_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, int shift, uint32_t mask, const uint32_t* __restrict__ digit_base,
    uint32_t* __restrict__ state /*[nblocks][256], zeroed*/, uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char os_raw[];
    OsSmem& sm = *reinterpret_cast<OsSmem*>(os_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (tid == 0) sm.bid = atomicAdd(ticket, 1u);
    for (int w = 0; w < RS_WARPS; ++w) sm.whist[w][tid] = 0;
    __syncthreads();
    const uint32_t b = sm.bid;
    const int64_t tile_base = (int64_t)b * OS_TILE;
    const int64_t wbase = tile_base + (int64_t)warp * OS_WARP_KEYS;
    uint64_t key[OS_ITEMS];
    uint32_t val[OS_ITEMS];
    uint32_t rank[OS_ITEMS];
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        key[r] = i < n ? __ldg(keys_in + i) : ~0ull;
        val[r] = i < n ? __ldg(vals_in + i) : 0u;
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const uint32_t peers = match_digit(d, ok);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (ok && lane == leader) {
            prev = sm.whist[warp][d];
            sm.whist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = prev + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // thread d owns digit d: per-warp counters -> exclusive offsets over warps, tile count, look-back
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = sm.whist[w][tid];
        sm.whist[w][tid] = cnt;
        cnt += c;
    }
    uint32_t* my_state = state + (int64_t)b * RS_BINS + tid;
    __stcg(my_state, (b == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | cnt);
    // exclusive scan of cnt over the 256 digits -> position of each digit run inside the sorted tile
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm.wsum[warp] = inc;
    __syncthreads();
    uint32_t lstart = inc - cnt;
    for (int w = 0; w < warp; ++w) lstart += sm.wsum[w];
    sm.lstart[tid] = lstart;
    __syncthreads();
    // reorder the tile by digit in shared memory first (needs tile-local offsets only): the keys leave the
    // registers, which the look-back below then uses for a wide window of independent loads
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const uint32_t pos = sm.lstart[d] + sm.whist[warp][d] + rank[r];
            sm.keys[pos] = key[r];
            sm.vals[pos] = val[r];
        }
    }
    // decoupled look-back: sum the counts of the tiles before this one for digit `tid`.  Windows of 8
    // predecessors: the loads are independent, so a long walk over tiles that have only published their
    // aggregate (the whole first wave of CTAs) costs one L2 round trip per 8 tiles instead of one per tile.
    uint32_t excl = 0;
    if (b > 0) {
        constexpr int WIN = 8;
        int64_t p = (int64_t)b - 1;
        bool found = false;
        while (!found) {
            uint32_t v[WIN];
#pragma unroll
            for (int k = 0; k < WIN; ++k)
                v[k] = p - k >= 0 ? __ldcg(state + (p - k) * RS_BINS + tid) : OS_FLAG_PREFIX;   // in front of tile 0: prefix 0
#pragma unroll
            for (int k = 0; k < WIN; ++k) {
                if (found) continue;
                while ((v[k] & ~OS_VAL_MASK) == 0) {
                    __nanosleep(20);
                    v[k] = __ldcg(state + (p - k) * RS_BINS + tid);
                }
                excl += v[k] & OS_VAL_MASK;
                if ((v[k] & ~OS_VAL_MASK) == OS_FLAG_PREFIX) found = true;
            }
            p -= WIN;
        }
        __stcg(my_state, OS_FLAG_PREFIX | (excl + cnt));
    }
    sm.gbase[tid] = digit_base[tid] + excl - lstart;
    __syncthreads();
    const int tile_n = (int)min((int64_t)OS_TILE, n - tile_base);
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = k * RS_THREADS + tid;
        if (i < tile_n) {
            const uint64_t kk = sm.keys[i];
            const uint32_t pos = sm.gbase[(uint32_t)(kk >> shift) & mask] + (uint32_t)i;
            keys_out[pos] = kk;
            vals_out[pos] = sm.vals[i];
        }
    }
}

}  // namespace

static size_t os_workspace_bytes(int64_t n) {
    const int64_t nb = emd_cdiv(n > 0 ? n : 1, OS_TILE);
    // [ghist 8*256][tickets 8][pad to 256 B][state passes*nb*256]
    return 256 * ((OS_MAX_PASSES * RS_BINS + OS_MAX_PASSES) * sizeof(uint32_t) / 256 + 1) +
           (size_t)OS_MAX_PASSES * nb * RS_BINS * sizeof(uint32_t);
}

extern "C" size_t emd_radix_sort_workspace_bytes(int64_t n) {
    const int64_t nb = emd_cdiv(n > 0 ? n : 1, RS_TILE);
    const size_t table = (size_t)nb * RS_BINS * sizeof(uint32_t);
    const size_t legacy = ((table + 255) / 256) * 256 + emd_scan_workspace_bytes(nb * RS_BINS);
    const size_t os = os_workspace_bytes(n);
    return legacy > os ? legacy : os;
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
    uint64_t* kin = keys0;  uint32_t* vin = vals0;
    uint64_t* kout = keys1; uint32_t* vout = vals1;
    int cur = 0;
    const int npasses = (end_bit - begin_bit + 7) / 8;
    if (n < ((int64_t)1 << 30) && npasses <= OS_MAX_PASSES) {
        const int64_t nb = emd_cdiv(n, OS_TILE);   // shadows the legacy tile count
        uint32_t* ghist = reinterpret_cast<uint32_t*>(workspace);
        uint32_t* tickets = ghist + OS_MAX_PASSES * RS_BINS;
        const size_t head = 256 * ((OS_MAX_PASSES * RS_BINS + OS_MAX_PASSES) * sizeof(uint32_t) / 256 + 1);
        uint32_t* state = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + head);
        cudaMemsetAsync(workspace, 0, head + (size_t)npasses * nb * RS_BINS * sizeof(uint32_t), stream);
        cudaFuncSetAttribute(rs_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem));
        const unsigned hb = (unsigned)(nb < 4 * EMD_NUM_SMS ? nb : 4 * EMD_NUM_SMS);
        EMD_LAUNCH(EK_SORT_HIST, stream, rs_multi_hist_kernel<<<hb, RS_THREADS, 0, stream>>>(kin, n, begin_bit, end_bit, npasses, ghist));
        EMD_LAUNCH(EK_SORT_HIST, stream, rs_digit_prefix_kernel<<<npasses, RS_BINS, 0, stream>>>(ghist));
        for (int p = 0; p < npasses; ++p) {
            const int shift = begin_bit + 8 * p;
            const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
            const uint32_t mask = (1u << bits) - 1u;
            EMD_LAUNCH(EK_SORT_SCATTER, stream, rs_onesweep_kernel<<<(unsigned)nb, RS_THREADS, sizeof(OsSmem), stream>>>(
                kin, vin, kout, vout, n, shift, mask, ghist + p * RS_BINS, state + (size_t)p * nb * RS_BINS, tickets + p));
            uint64_t* tk = kin; kin = kout; kout = tk;
            uint32_t* tv = vin; vin = vout; vout = tv;
            cur ^= 1;
        }
        EMD_CHECK_LAUNCH("radix_sort");
        if (result_buffer) *result_buffer = cur;
        return EMD_OK;
    }
    uint32_t* table = reinterpret_cast<uint32_t*>(workspace);
    const size_t table_bytes = (((size_t)nb * RS_BINS * sizeof(uint32_t) + 255) / 256) * 256;
    void* scan_ws = reinterpret_cast<char*>(workspace) + table_bytes;
    const size_t scan_ws_bytes = ws_bytes - table_bytes;
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
