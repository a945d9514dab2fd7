// K1g on the 5th-generation tensor cores -- EXPERIMENTAL, OFF BY DEFAULT (emd_dense_set_tc(1) or EMD_DENSE_TC=1).
//
// STATUS: written and compiled in round 1 after the round's GPU budget was spent; it has NOT run on hardware yet.
// The default path of emd_dense_fwd / emd_dense_bwd is the fp32 SIMT GEMM of deform_net.cu (parity-tested on B200);
// nothing selects this kernel unless the switch above is set.  tools/tc_dense_check.py is the first thing to run on a
// GPU: it compares this kernel with the SIMT kernel on the shapes of the DeformableNodes network.
//
// The 256-wide layers of ConditionalDeformNetwork (OmniRe/models/modules.py:411-457) as tcgen05.mma kernels:
//
//     Y[M, 0:N] = epilogue( A[M, 0:K] . B )          forward: A = X,  B(k, n) = W[n*ldw + k], epilogue = bias + ReLU
//                                                    dgrad  : A = dZ, B(k, n) = W[k*ldw + n], epilogue = (mask > 0)
//     dW[Nout, K] = dZ^T . X                         wgrad  : reduction over the rows, split across CTAs (wgrad_tc_kernel)
//
// Same numerics contract as csrc/mlp_tc.cu (3xTF32: every operand split into hi = TF32-exact part and lo = x - hi,
// products accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator -> fp32-class results) and the same
// hand-written descriptors / PTX (tc_common.cuh), but a different tile plan because N = 256 does not leave room to keep
// W resident in shared memory:
//   * CTA = 256 threads, persistent over 128-row tiles, ONE CTA per SM (198 KB of dynamic shared memory);
//   * accumulator: 128 lanes x 256 columns of TMEM (half of the SM's 512 columns);
//   * the reduction is consumed 32 columns at a time; a chunk's A (128 x 32) and B (N x 32) operands are split and
//     written to one of TWO shared-memory stages in the K-major no-swizzle canonical layout; one thread issues the
//     chunk's 4 k-steps x 3 MMAs and commits to that stage's mbarrier; staging of chunk c + 1 into the other stage runs
//     under the MMAs of chunk c, a stage is rewritten only after its barrier completed;
//   * epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31, warps 0-3 columns [0, 128), warps 4-7 columns [128, 256).
// B comes from L2 (the whole weight matrix is <= 365 KB); pre-splitting W once per step into the canonical layout so a
// chunk becomes one bulk copy, and a second TMEM accumulator so the epilogue overlaps the next tile, are the follow-ups.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stage_math.cuh"

namespace {

constexpr int DT_THREADS = DTS_THREADS;
constexpr int DT_KC = DTS_KC;                               // reduction columns per chunk (8 core-matrix columns)
constexpr int DT_NMAX = DTS_NMAX;                           // output columns per tile (UMMA N <= 256)
constexpr int DT_A_BYTES = (DT_KC / 4) * TC_A_LBO;          // one of {hi, lo} of an A chunk
constexpr int DT_B_LBO = DTS_B_LBO;                         // bytes between K-adjacent core-matrix columns of B
static_assert(TC_A_LBO == DTS_A_LBO && TC_SBO == DTS_SBO && TC_ROWS == DTS_ROWS, "tc_stage_math.cuh mirrors tc_common.cuh");
constexpr int DT_B_BYTES = (DT_KC / 4) * DT_B_LBO;          // one of {hi, lo} of a B chunk
constexpr int DT_STAGE_BYTES = 2 * DT_A_BYTES + 2 * DT_B_BYTES;
constexpr int DT_SMEM_BYTES = 2 * DT_STAGE_BYTES;           // 197 632 B
constexpr int DT_TMEM_COLS = 256;

struct DenseTcArgs {
    const float* A;        // [M, K] row stride lda (K = reduction length, multiple of 4; rows 16-byte aligned)
    int64_t lda;
    const float* W;        // TB ? W[k*ldw + n] : W[n*ldw + k]
    int64_t ldw;
    const float* bias;     // [N] or NULL
    const float* mask;     // or NULL: result *= (mask[m*ldmask + n] > 0)
    int64_t ldmask;
    float* Y;              // [M, N] row stride ldy
    int64_t ldy;
    int64_t M;
    int K, N;
    int relu;
    int vec_store;         // Y rows 16-byte aligned and ldy % 4 == 0
};

template <bool TB>
__global__ void __launch_bounds__(DT_THREADS, 1) dense_tc_kernel(const DenseTcArgs a) {
    extern __shared__ __align__(128) unsigned char dt_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npad = (a.N + 15) / 16 * 16;
    const int kq = a.K / 4;                         // core-matrix columns that hold data
    const int ksteps = (a.K + 7) / 8;               // MMA k-steps of the whole reduction
    const int nchunks = (a.K + DT_KC - 1) / DT_KC;

    for (int e = tid; e < DT_SMEM_BYTES / 16; e += DT_THREADS) reinterpret_cast<float4*>(dt_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(DT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, npad);
    uint32_t phase0 = 0, phase1 = 0;
    bool pend0 = false, pend1 = false;

    // thread -> (row, core-matrix column) of the 128 x 32 A chunk: 4 float4 per thread, rows cr + 32 i (as mlp_tc.cu)
    const int64_t n_tiles = (a.M + TC_ROWS - 1) / TC_ROWS;
    float4 pre[4];
    auto load_a = [&](int64_t row0, int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r, cj;
            dts_a_elem(tid, i, r, cj);
            const int jg = c * (DT_KC / 4) + cj;
            const int64_t row = row0 + r;
            pre[i] = (row < a.M && jg < kq) ? __ldg(reinterpret_cast<const float4*>(a.A + row * a.lda) + jg) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if ((int64_t)blockIdx.x < n_tiles) load_a((int64_t)blockIdx.x * TC_ROWS, 0);
    int it = 0;   // chunks issued so far: stage = it & 1
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        for (int c = 0; c < nchunks; ++c, ++it) {
            const int st = it & 1;
            unsigned char* sAhi = dt_smem + st * DT_STAGE_BYTES;
            unsigned char* sAlo = sAhi + DT_A_BYTES;
            unsigned char* sBhi = sAlo + DT_A_BYTES;
            unsigned char* sBlo = sBhi + DT_B_BYTES;
            // the MMAs that read this stage two chunks ago must have completed before it is overwritten
            if (st == 0) {
                if (pend0) { mbar_wait(bar0, phase0); phase0 ^= 1u; pend0 = false; }
            } else {
                if (pend1) { mbar_wait(bar1, phase1); phase1 ^= 1u; pend1 = false; }
            }
            // ---- A: registers (loaded one chunk ahead) -> hi/lo split -> canonical K-major layout ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 x = pre[i];
                float4 hi, lo;
                split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
                int r, cj;
                dts_a_elem(tid, i, r, cj);
                *reinterpret_cast<float4*>(sAhi + dts_a_store_offset(r, cj)) = hi;
                *reinterpret_cast<float4*>(sAlo + dts_a_store_offset(r, cj)) = lo;
            }
            // ---- B: this chunk of W (L2-resident) -> hi/lo split -> canonical layout: element (n, k) of the chunk at
            //      (k >> 2) * DT_B_LBO + n * 16 + (k & 3) * 4 ----
            if (!TB) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int n, j;
                    dts_b_elem_fwd(tid, i, n, j);
                    const int jg = c * (DT_KC / 4) + j;
                    if (n < npad) {
                        const float4 w = (n < a.N && jg < kq) ? __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)n * a.ldw) + jg)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 hi, lo;
                        split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y); split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
                        *reinterpret_cast<float4*>(sBhi + dts_b_store_offset_fwd(n, j)) = hi;
                        *reinterpret_cast<float4*>(sBlo + dts_b_store_offset_fwd(n, j)) = lo;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int k, n;                                       // W row k of the chunk, 4 consecutive output columns
                    dts_b_elem_dgrad(tid, i, k, n);
                    const int kg = c * DT_KC + k;
                    if (n < npad) {
                        const float4 w = (kg < a.K && n < a.N) ? __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)kg * a.ldw + n))
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);   // N % 4 == 0 on this path
                        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float hi, lo;
                            split_tf32(wv[q], hi, lo);
                            const int off = dts_b_store_offset_dgrad(k, n + q);
                            *reinterpret_cast<float*>(sBhi + off) = hi;
                            *reinterpret_cast<float*>(sBlo + off) = lo;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
            __syncthreads();
            // ---- next chunk's A loads fly under this chunk's MMAs ----
            if (c + 1 < nchunks) load_a(row0, c + 1);
            else if (tile + gridDim.x < n_tiles) load_a((tile + gridDim.x) * TC_ROWS, 0);
            // ---- MMA: one thread issues, completion arrives on this stage's mbarrier ----
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);
                const int s0 = c * (DT_KC / 8);
                const int s1 = min(s0 + DT_KC / 8, ksteps);
                for (int s = s0; s < s1; ++s) {
                    const int sl = s - s0;   // k-step inside the stage
                    const uint64_t dAh = umma_desc(aHi + dts_kstep_offset(sl, TC_A_LBO), TC_A_LBO, TC_SBO);
                    const uint64_t dAl = umma_desc(aLo + dts_kstep_offset(sl, TC_A_LBO), TC_A_LBO, TC_SBO);
                    const uint64_t dBh = umma_desc(bHi + dts_kstep_offset(sl, DT_B_LBO), DT_B_LBO, TC_SBO);
                    const uint64_t dBl = umma_desc(bLo + dts_kstep_offset(sl, DT_B_LBO), DT_B_LBO, TC_SBO);
                    umma_tf32(tmem, dAl, dBh, idesc, s > 0 ? 1u : 0u);   // small terms first
                    umma_tf32(tmem, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem, dAh, dBh, idesc, 1u);
                }
                umma_commit(st == 0 ? bar0 : bar1);
            }
            if (st == 0) pend0 = true; else pend1 = true;
        }
        // every MMA of the tile must have landed in TMEM (both stages' barriers are consumed so their phases stay in step)
        if (pend0) { mbar_wait(bar0, phase0); phase0 ^= 1u; pend0 = false; }
        if (pend1) { mbar_wait(bar1, phase1); phase1 ^= 1u; pend1 = false; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue ----
        const int64_t row = row0 + (warp & 3) * 32 + lane;
        const int cbeg = (warp >> 2) * 128;
        const int cend = min(npad, cbeg + 128);
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
            if (row < a.M) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float y = v[c];
                    if (c0 + c < a.N) {
                        if (a.bias) y += __ldg(a.bias + c0 + c);
                        if (a.relu) y = fmaxf(y, 0.f);
                        if (a.mask && !(__ldg(a.mask + row * a.ldmask + c0 + c) > 0.f)) y = 0.f;
                    }
                    v[c] = y;
                }
                float* yr = a.Y + row * a.ldy + c0;
                if (a.vec_store && c0 + 16 <= a.N) {
#pragma unroll
                    for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(yr + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c0 + c < a.N) yr[c] = v[c];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM is free again
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(DT_TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: partial[z][n][k] = sum over the rows r of split z of dZ[r][n] X[r][k]
//   UMMA M = 128 output features (block nb), UMMA N = up to 256 input features (block cb), reduction = rows, 32 per chunk.
//   Both operands are row-contiguous in memory across the UMMA M / N index (dZ[r][n .. n+3], X[r][k .. k+3]), so both are
//   staged with the transposing map (4 scalar stores per float4).  A CTA keeps its accumulator in TMEM over ALL chunks of
//   its row range and runs the epilogue once; the per-split partials are summed in a fixed order by the caller
//   (split_reduce_kernel of deform_net.cu): bit-reproducible, no atomics.
// ---------------------------------------------------------------------------------------------------------------
struct WgradTcArgs {
    const float* dZ;       // [M, Nout] row stride lddz
    int64_t lddz;
    const float* X;        // [M, K] row stride ldx
    int64_t ldx;
    float* partial;        // [splits][Nout][K]
    int64_t M;
    int K, Nout;
    int n_blocks;          // ceil(Nout / 128)
    int64_t rows_per_split;
};

__global__ void __launch_bounds__(DT_THREADS, 1) wgrad_tc_kernel(const WgradTcArgs a) {
    extern __shared__ __align__(128) unsigned char dt_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = blockIdx.x % a.n_blocks, cb = blockIdx.x / a.n_blocks;
    const int n0 = nb * TC_ROWS;                              // first output feature of this tile
    const int k0 = cb * DT_NMAX;                              // first input feature of this tile
    const int ncols = min(DT_NMAX, a.K - k0);
    const int npad = (ncols + 15) / 16 * 16;
    const int64_t rbeg = (int64_t)blockIdx.y * a.rows_per_split;
    const int64_t rend = min(a.M, rbeg + a.rows_per_split);
    const int nchunks = (int)((rend - rbeg + DT_KC - 1) / DT_KC);

    for (int e = tid; e < DT_SMEM_BYTES / 16; e += DT_THREADS) reinterpret_cast<float4*>(dt_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(DT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, npad);
    uint32_t phase0 = 0, phase1 = 0;
    bool pend0 = false, pend1 = false;

    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        unsigned char* sAhi = dt_smem + st * DT_STAGE_BYTES;
        unsigned char* sAlo = sAhi + DT_A_BYTES;
        unsigned char* sBhi = sAlo + DT_A_BYTES;
        unsigned char* sBlo = sBhi + DT_B_BYTES;
        if (st == 0) {
            if (pend0) { mbar_wait(bar0, phase0); phase0 ^= 1u; pend0 = false; }
        } else {
            if (pend1) { mbar_wait(bar1, phase1); phase1 ^= 1u; pend1 = false; }
        }
        const int64_t r0 = rbeg + (int64_t)c * DT_KC;
        // ---- A = dZ^T chunk: element (m = output feature, k = row) ----
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int k, m;
            dts_at_elem(tid, i, k, m);
            const int64_t row = r0 + k;
            const int n = n0 + m;
            float wv[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < rend && n < a.Nout) {
                if (n + 3 < a.Nout) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(a.dZ + row * a.lddz + n));   // Nout % 4 == 0, lddz % 4 == 0
                    wv[0] = w.x; wv[1] = w.y; wv[2] = w.z; wv[3] = w.w;
                } else {
                    for (int q = 0; q < 4; ++q)
                        if (n + q < a.Nout) wv[q] = __ldg(a.dZ + row * a.lddz + n + q);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float hi, lo;
                split_tf32(wv[q], hi, lo);
                const int off = dts_at_store_offset(k, m + q);
                *reinterpret_cast<float*>(sAhi + off) = hi;
                *reinterpret_cast<float*>(sAlo + off) = lo;
            }
        }
        // ---- B = X chunk: element (n = input feature, k = row) ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int k, n;
            dts_b_elem_dgrad(tid, i, k, n);
            if (n < npad) {
                const int64_t row = r0 + k;
                float wv[4] = {0.f, 0.f, 0.f, 0.f};
                if (row < rend && n < ncols) {
                    if (n + 3 < ncols) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(a.X + row * a.ldx + k0 + n));   // ldx % 4 == 0
                        wv[0] = w.x; wv[1] = w.y; wv[2] = w.z; wv[3] = w.w;
                    } else {
                        for (int q = 0; q < 4; ++q)
                            if (n + q < ncols) wv[q] = __ldg(a.X + row * a.ldx + k0 + n + q);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float hi, lo;
                    split_tf32(wv[q], hi, lo);
                    const int off = dts_b_store_offset_dgrad(k, n + q);
                    *reinterpret_cast<float*>(sBhi + off) = hi;
                    *reinterpret_cast<float*>(sBlo + off) = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);
#pragma unroll
            for (int sl = 0; sl < DT_KC / 8; ++sl) {     // rows beyond rend were staged as zeros: all 4 k-steps are safe
                const uint64_t dAh = umma_desc(aHi + dts_kstep_offset(sl, TC_A_LBO), TC_A_LBO, TC_SBO);
                const uint64_t dAl = umma_desc(aLo + dts_kstep_offset(sl, TC_A_LBO), TC_A_LBO, TC_SBO);
                const uint64_t dBh = umma_desc(bHi + dts_kstep_offset(sl, DT_B_LBO), DT_B_LBO, TC_SBO);
                const uint64_t dBl = umma_desc(bLo + dts_kstep_offset(sl, DT_B_LBO), DT_B_LBO, TC_SBO);
                umma_tf32(tmem, dAl, dBh, idesc, (c > 0 || sl > 0) ? 1u : 0u);
                umma_tf32(tmem, dAh, dBl, idesc, 1u);
                umma_tf32(tmem, dAh, dBh, idesc, 1u);
            }
            umma_commit(st == 0 ? bar0 : bar1);
        }
        if (st == 0) pend0 = true; else pend1 = true;
    }
    if (pend0) { mbar_wait(bar0, phase0); phase0 ^= 1u; pend0 = false; }
    if (pend1) { mbar_wait(bar1, phase1); phase1 ^= 1u; pend1 = false; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: lane = output feature, columns = input features -> partial[z][n0 + lane ..][k0 + c] ----
    const int n = n0 + (warp & 3) * 32 + lane;
    float* out = a.partial + ((int64_t)blockIdx.y * a.Nout + n) * a.K + k0;
    const int cbeg = (warp >> 2) * 128;
    const int cend = min(npad, cbeg + 128);
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
        float v[16];
        if (nchunks > 0) {
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = 0.f;      // empty row range: TMEM was never written
        }
        if (n < a.Nout) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c0 + c < ncols) out[c0 + c] = v[c];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(DT_TMEM_COLS) : "memory");
}

int g_dense_tc = -1;   // -1: read EMD_DENSE_TC on first use

}  // namespace

// 1 when the experimental tensor-core path of emd_dense_fwd / emd_dense_bwd(dgrad) is selected (default 0).
extern "C" int emd_dense_tc_enabled() {
    if (g_dense_tc < 0) {
        const char* e = getenv("EMD_DENSE_TC");
        g_dense_tc = (e && e[0] == '1') ? 1 : 0;
    }
    return g_dense_tc;
}

// Select (1) / deselect (0) the experimental tensor-core path.  Process-wide switch for measurement and bring-up.
extern "C" void emd_dense_set_tc(int on) { g_dense_tc = on ? 1 : 0; }

static bool dt_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Used by emd_dense_fwd / emd_dense_bwd (deform_net.cu).  Returns 1 when the shape qualifies and the kernel was launched
// (status in *rc), 0 when the caller must take the SIMT path.
int emd_dense_tc_try(int dgrad, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* mask,
                     int64_t ldmask, float* Y, int64_t ldy, int64_t M, int K, int N, int relu, cudaStream_t stream, int* rc) {
    if (!emd_dense_tc_enabled() || M <= 0) return 0;
    if (K < 4 || K % 4 != 0 || N < 1 || N > DT_NMAX || lda % 4 != 0 || ldw % 4 != 0 || !dt_al16(A) || !dt_al16(W)) return 0;
    if (dgrad && N % 4 != 0) return 0;
    DenseTcArgs a;
    a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.bias = bias; a.mask = mask; a.ldmask = ldmask; a.Y = Y; a.ldy = ldy;
    a.M = M; a.K = K; a.N = N; a.relu = relu; a.vec_store = dt_al16(Y) && ldy % 4 == 0;
    const int64_t tiles = emd_cdiv(M, TC_ROWS);
    const unsigned grid = (unsigned)(tiles < EMD_NUM_SMS ? tiles : EMD_NUM_SMS);
    cudaError_t e;
    if (dgrad) {
        e = cudaFuncSetAttribute(dense_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
        if (e == cudaSuccess)
            EMD_LAUNCH(EK_DENSE_BWD, stream, (dense_tc_kernel<true><<<grid, DT_THREADS, DT_SMEM_BYTES, stream>>>(a)));
    } else {
        e = cudaFuncSetAttribute(dense_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
        if (e == cudaSuccess)
            EMD_LAUNCH(EK_DENSE_FWD, stream, (dense_tc_kernel<false><<<grid, DT_THREADS, DT_SMEM_BYTES, stream>>>(a)));
    }
    if (e != cudaSuccess) {
        emd_set_error("emd_dense(tc): %s", cudaGetErrorString(e));
        *rc = EMD_ERR_CUDA;
        return 1;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        emd_set_error("emd_dense(tc): CUDA error: %s", cudaGetErrorString(e));
        *rc = EMD_ERR_CUDA;
        return 1;
    }
    *rc = EMD_OK;
    return 1;
}

// Floats of the per-split partial buffer the tensor-core weight gradient needs (0 when the path is off / the shape does
// not qualify); emd_dense_bwd_workspace_bytes takes the larger of this and the SIMT requirement.
size_t emd_dense_tc_wgrad_partial_floats(int64_t M, int K, int Nout) {
    if (!emd_dense_tc_enabled() || M <= 0 || Nout % 4 != 0 || K < 4) return 0;
    const DtsWgradSplit s = dts_wgrad_split(M, K, Nout, EMD_NUM_SMS);
    return (size_t)s.splits * (size_t)Nout * (size_t)K;
}

// Weight gradient on the tensor cores into `partial` ([splits][Nout][K]); *splits_out tells the caller how many partials
// to sum.  Returns 1 when launched (status in *rc), 0 when the caller must take the SIMT path.
int emd_dense_tc_try_wgrad(const float* X, int64_t ldx, const float* dZ, int64_t lddz, int64_t M, int K, int Nout, float* partial,
                           int* splits_out, cudaStream_t stream, int* rc) {
    if (!emd_dense_tc_enabled() || M <= 0) return 0;
    if (Nout % 4 != 0 || K < 4 || lddz % 4 != 0 || ldx % 4 != 0 || !dt_al16(dZ) || !dt_al16(X)) return 0;
    const DtsWgradSplit s = dts_wgrad_split(M, K, Nout, EMD_NUM_SMS);
    WgradTcArgs a;
    a.dZ = dZ; a.lddz = lddz; a.X = X; a.ldx = ldx; a.partial = partial; a.M = M; a.K = K; a.Nout = Nout;
    a.n_blocks = s.n_blocks; a.rows_per_split = s.rows_per_split;
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
    if (e == cudaSuccess) {
        const dim3 grid((unsigned)(s.n_blocks * s.col_blocks), (unsigned)s.splits);
        EMD_LAUNCH(EK_DENSE_BWD, stream, (wgrad_tc_kernel<<<grid, DT_THREADS, DT_SMEM_BYTES, stream>>>(a)));
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        emd_set_error("emd_dense(tc wgrad): %s", cudaGetErrorString(e));
        *rc = EMD_ERR_CUDA;
        return 1;
    }
    *splits_out = s.splits;
    *rc = EMD_OK;
    return 1;
}
