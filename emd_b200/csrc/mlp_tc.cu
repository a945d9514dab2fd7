// K1d on the 5th-generation tensor cores: the Linear layers of the S3Gaussian EMD deformation network
// (S3Gaussian/scene/deformation.py:100-185, 339-386) as tcgen05.mma kernels with the accumulator in TMEM.
//
//     Y[M,Nout] = act_out( act_in(X[M,K]) . W[Nout,K]^T + b )           act = ReLU or identity
//
// The reference computes these layers in fp32 (cuBLAS SGEMM, TF32 off), and the parity bar on what they feed is
// 1e-4 absolute on images / 1e-3 relative on gradients, so a single TF32 pass (10-bit mantissa) is not enough.
// Every operand is therefore split into hi = x with the low 13 mantissa bits cleared (exactly representable in
// TF32) and lo = x - hi (exact in fp32), and a product is accumulated as hi*hi + lo*hi + hi*lo in the fp32 TMEM
// accumulator ("3xTF32"): relative error ~2^-21 per product, i.e. fp32-class results at three tensor-core passes.
//
// CTA = 256 threads, one 128-row tile at a time (persistent over tiles), two CTAs per SM so one CTA's loads and
// epilogue run under the other's MMAs:
//   stage   : the tile's X block is consumed 32 columns at a time: global float4 loads one chunk ahead into
//             registers, act_in, hi/lo split, both copies written in the K-major no-swizzle UMMA canonical layout
//             (core matrix = 8 rows x 16 B contiguous; row groups 128 B apart; K-adjacent core matrices LBO apart)
//   mma     : one thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=Npad, K=8) per chunk and commits to
//             an mbarrier; the A buffer is rewritten only after that barrier completes
//   epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31 (tcgen05.ld 32x32b), adds the bias, applies act_out, stores Y
// W (hi and lo) is staged once per CTA.  No cuBLAS, no CUTLASS: descriptors and PTX are written out below.
#include "common.cuh"
#include "tc_common.cuh"

namespace {


// dynamic shared memory: [A_hi chunk | A_lo chunk | B_hi | B_lo], every region a multiple of 16 B.  A is staged
// TC_KC columns at a time (one buffer; the second CTA on the SM supplies the overlap), W stays resident.
constexpr int TC_THREADS = 256;
constexpr int TC_KC = 32;                    // A columns staged per chunk (8 core-matrix columns, 4 MMA k-steps)
constexpr int TC_A_CHUNK_BYTES = (TC_KC / 4) * TC_A_LBO;

struct TcLayout {
    int kchunks;        // Kpad / 4 core-matrix columns
    int npad;           // Nout padded to a multiple of 16
    int b_lbo;          // bytes between K-adjacent core-matrix columns of B
    int b_bytes;
    __host__ __device__ TcLayout(int K, int Nout) {
        const int kpad = (K + 7) / 8 * 8;
        kchunks = kpad / 4;
        npad = (Nout + 15) / 16 * 16;
        b_lbo = npad * 16 + 16;
        b_bytes = kchunks * b_lbo;
    }
    __host__ __device__ size_t total() const { return 2 * (size_t)TC_A_CHUNK_BYTES + 2 * (size_t)b_bytes; }
};

// Epilogue store of a 32-row x 16-column accumulator block held one row per lane (tcgen05.ld 32x32b).  Stored straight
// from those registers, every store instruction touches 32 different 128-byte lines with 16 bytes each; the launch
// list of the S3G step showed the epilogues, not HBM, bounding these kernels.  The block is therefore passed through a
// per-warp staging area (32 rows x 80 B: conflict-free float4 writes) and written back with FOUR lanes per row: one
// instruction covers 8 rows x 64 contiguous bytes -- whole 32-byte sectors, a quarter of the line visits.
//   dst / mask: address of (first row of the block, first column of the block); ld: row pitch in floats (% 4 == 0);
//   rows: valid rows of the block (<= 0: none); cols: valid columns (a multiple of 4, may exceed 16);
//   mask != nullptr: the value is zeroed where mask <= 0 (the ReLU derivative of the layer input).
constexpr int TC_STAGE_LD = 20;                                  // floats per staged row (16 + 4: bank spread)
constexpr int TC_STAGE_WARP = 32 * TC_STAGE_LD;                  // floats per warp; 8 warps = 20 480 B <= the A chunk buffers

__device__ __forceinline__ void warp_block_store16(float* __restrict__ stage, const float (&v)[16], int lane,
                                                   float* __restrict__ dst, const float* __restrict__ mask, int64_t ld,
                                                   int rows, int cols) {
    float4* w = reinterpret_cast<float4*>(stage + lane * TC_STAGE_LD);
    w[0] = make_float4(v[0], v[1], v[2], v[3]);
    w[1] = make_float4(v[4], v[5], v[6], v[7]);
    w[2] = make_float4(v[8], v[9], v[10], v[11]);
    w[3] = make_float4(v[12], v[13], v[14], v[15]);
    __syncwarp();
    const int q = lane & 3, rr = lane >> 2;
    if (4 * q + 4 <= cols) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = r * 8 + rr;
            if (row < rows) {
                float4 o = *reinterpret_cast<const float4*>(stage + row * TC_STAGE_LD + q * 4);
                if (mask) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(mask + row * ld + q * 4));
                    if (!(x.x > 0.f)) o.x = 0.f;
                    if (!(x.y > 0.f)) o.y = 0.f;
                    if (!(x.z > 0.f)) o.z = 0.f;
                    if (!(x.w > 0.f)) o.w = 0.f;
                }
                *reinterpret_cast<float4*>(dst + row * ld + q * 4) = o;
            }
        }
    }
    __syncwarp();   // the staging rows are rewritten by the next block
}

template <bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(TC_THREADS, 2) linear_fwd_tc_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                      const float* __restrict__ bias, int64_t M, int K,
                                                                      int Nout, float* __restrict__ Y) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[TC_NMAX];
    const TcLayout L(K, Nout);
    unsigned char* sAhi = tc_smem;
    unsigned char* sAlo = sAhi + TC_A_CHUNK_BYTES;
    unsigned char* sBhi = sAlo + TC_A_CHUNK_BYTES;
    unsigned char* sBlo = sBhi + L.b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kq = K / 4;   // float4 per row (K % 4 == 0)

    // ---- one-time setup: zero the padded operands, stage W (hi / lo), bias, barrier, TMEM ----
    for (int e = tid; e < (int)(L.total() / 16); e += TC_THREADS) reinterpret_cast<float4*>(tc_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < TC_NMAX) s_bias[tid] = tid < Nout ? __ldg(bias + tid) : 0.f;
    __syncthreads();
    for (int e = tid; e < Nout * kq; e += TC_THREADS) {
        const int n = e / kq, j = e - n * kq;
        const float4 w = __ldg(reinterpret_cast<const float4*>(W) + e);   // W[n][4j..4j+3]
        float4 hi, lo;
        split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y); split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sBhi + j * L.b_lbo + n * 16) = hi;
        *reinterpret_cast<float4*>(sBlo + j * L.b_lbo + n * 16) = lo;
    }
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, L.npad);
    const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);
    const int nchunks = (L.kchunks * 4 + TC_KC - 1) / TC_KC;   // A chunks per tile
    uint32_t phase = 0;

    // thread -> (row, core-matrix column) of the 128 x 32 chunk: 4 float4 per thread, rows r0 + 32 i
    const int cj = tid & 7, cr = tid >> 3;
    const int64_t n_tiles = (M + TC_ROWS - 1) / TC_ROWS;
    float4 pre[4];
    auto load_chunk = [&](int64_t row0, int c) {
        const int jg = c * (TC_KC / 4) + cj;   // core-matrix column in the full K
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + cr + 32 * i;
            pre[i] = (row < M && jg < kq) ? __ldg(reinterpret_cast<const float4*>(X + row * K) + jg) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if ((int64_t)blockIdx.x < n_tiles) load_chunk((int64_t)blockIdx.x * TC_ROWS, 0);
    bool mma_pending = false;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        for (int c = 0; c < nchunks; ++c) {
            if (mma_pending) {   // the MMAs that read the A buffer must have completed before it is overwritten
                mbar_wait(bar, phase);
                phase ^= 1u;
                mma_pending = false;
            }
            // ---- registers (loaded one step ahead) -> act_in, hi/lo split, canonical K-major layout ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 x = pre[i];
                if (RELU_IN) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                float4 hi, lo;
                split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
                *reinterpret_cast<float4*>(sAhi + cj * TC_A_LBO + (cr + 32 * i) * 16) = hi;
                *reinterpret_cast<float4*>(sAlo + cj * TC_A_LBO + (cr + 32 * i) * 16) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
            __syncthreads();
            // ---- next chunk's global loads fly under this chunk's MMAs ----
            if (c + 1 < nchunks) load_chunk(row0, c + 1);
            else if (tile + gridDim.x < n_tiles) load_chunk((tile + gridDim.x) * TC_ROWS, 0);
            // ---- MMA: one thread issues, completion arrives on the mbarrier ----
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int s0 = c * (TC_KC / 8);
                const int s1 = min(s0 + TC_KC / 8, L.kchunks / 2);
                for (int s = s0; s < s1; ++s) {
                    const int sl = s - s0;   // k-step inside the A chunk buffer
                    const uint64_t dAh = umma_desc(aHi + 2 * sl * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dAl = umma_desc(aLo + 2 * sl * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dBh = umma_desc(bHi + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                    const uint64_t dBl = umma_desc(bLo + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                    umma_tf32(tmem, dAl, dBh, idesc, s > 0 ? 1u : 0u);   // small terms first
                    umma_tf32(tmem, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem, dAh, dBh, idesc, 1u);
                }
                umma_commit(bar);
            }
            mma_pending = true;
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        mma_pending = false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31; warps 4-7 take the upper half of the columns.
        // The A chunk buffers are free here (every MMA of the tile has completed, the next chunk is still in registers):
        // they stage the coalesced store.
        const int64_t rbase = row0 + (warp & 3) * 32;
        const int64_t row = rbase + lane;
        const int rows_valid = (int)(M - rbase < 32 ? M - rbase : 32);
        float* stage = reinterpret_cast<float*>(tc_smem) + warp * TC_STAGE_WARP;
        const int cbeg = (warp >> 2) * 32;                       // warps 0-3: columns [0, 32), warps 4-7: [32, npad)
        const int cend = min(L.npad, cbeg + 32);
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float y = v[c] + s_bias[c0 + c];
                if (RELU_OUT) y = fmaxf(y, 0.f);
                v[c] = y;
            }
            if ((Nout & 3) == 0) {
                if (c0 < Nout) warp_block_store16(stage, v, lane, Y + rbase * Nout + c0, nullptr, Nout, rows_valid, Nout - c0);
            } else if (row < M) {
                float* yr = Y + row * Nout + c0;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c0 + c < Nout) yr[c] = v[c];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM is free again
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// backward, data gradient:  dG = dY (.) relu'(Y)  (written out for the weight-gradient kernel when the layer has an
//                           output ReLU; without one dG == dY and the weight-gradient kernel reads dY itself),
//                           dX = (dG . W) (.) relu'(X)
// The same GEMM skeleton as the forward with A = dG [rows x Nout] (reduction over Nout) and B[n][k] = W[k][n].
// ---------------------------------------------------------------------------------------------------------------
template <bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(TC_THREADS, 2) linear_dgrad_tc_kernel(
    const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ Y, const float* __restrict__ dY,
    int64_t M, int K, int Nout, int tmem_cols, float* __restrict__ dG, float* __restrict__ dX) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const TcLayout L(Nout, K);   // reduction = Nout, output columns = K
    unsigned char* sAhi = tc_smem;
    unsigned char* sAlo = sAhi + TC_A_CHUNK_BYTES;
    unsigned char* sBhi = sAlo + TC_A_CHUNK_BYTES;
    unsigned char* sBlo = sBhi + L.b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool vec = (Nout & 3) == 0;

    for (int e = tid; e < (int)(L.total() / 16); e += TC_THREADS) reinterpret_cast<float4*>(tc_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int e = tid; e < Nout * K; e += TC_THREADS) {
        const int k = e / K, n = e - k * K;   // W[k][n]: k = output feature (reduction index here), n = input feature
        float hi, lo;
        split_tf32(__ldg(W + e), hi, lo);
        *reinterpret_cast<float*>(sBhi + (k >> 2) * L.b_lbo + n * 16 + (k & 3) * 4) = hi;
        *reinterpret_cast<float*>(sBlo + (k >> 2) * L.b_lbo + n * 16 + (k & 3) * 4) = lo;
    }
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, L.npad);
    const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);
    const int nchunks = (L.kchunks * 4 + TC_KC - 1) / TC_KC;
    uint32_t phase = 0;
    const int cj = tid & 7, cr = tid >> 3;
    const int64_t n_tiles = (M + TC_ROWS - 1) / TC_ROWS;
    const int split = ((L.npad / 16 + 1) / 2) * 16;   // warps 0-3: columns [0, split), warps 4-7: [split, npad)

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        for (int c = 0; c < nchunks; ++c) {
            // dG chunk: rows cr + 32 i, reduction columns 4 jg .. 4 jg + 3
            const int jg = c * (TC_KC / 4) + cj;
            float4 g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t row = row0 + cr + 32 * i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < M && 4 * jg < Nout) {
                    if (vec) {
                        v = __ldg(reinterpret_cast<const float4*>(dY + row * Nout) + jg);
                        if (RELU_OUT) {
                            const float4 y = __ldg(reinterpret_cast<const float4*>(Y + row * Nout) + jg);
                            if (!(y.x > 0.f)) v.x = 0.f; if (!(y.y > 0.f)) v.y = 0.f; if (!(y.z > 0.f)) v.z = 0.f; if (!(y.w > 0.f)) v.w = 0.f;
                        }
                        if (RELU_OUT) reinterpret_cast<float4*>(dG + row * Nout)[jg] = v;
                    } else {
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int col = 4 * jg + q;
                            if (col < Nout) {
                                float gq = __ldg(dY + row * Nout + col);
                                if (RELU_OUT && !(__ldg(Y + row * Nout + col) > 0.f)) gq = 0.f;
                                if (RELU_OUT) dG[row * Nout + col] = gq;
                                t[q] = gq;
                            }
                        }
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
                g[i] = v;
            }
            if (dX == nullptr) continue;   // (kernel-uniform) only dG is wanted
            if (c > 0) {   // the MMAs reading the A buffer must have completed before it is overwritten
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 hi, lo;
                split_tf32(g[i].x, hi.x, lo.x); split_tf32(g[i].y, hi.y, lo.y); split_tf32(g[i].z, hi.z, lo.z); split_tf32(g[i].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(sAhi + cj * TC_A_LBO + (cr + 32 * i) * 16) = hi;
                *reinterpret_cast<float4*>(sAlo + cj * TC_A_LBO + (cr + 32 * i) * 16) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int s0 = c * (TC_KC / 8);
                const int s1 = min(s0 + TC_KC / 8, L.kchunks / 2);
                for (int s = s0; s < s1; ++s) {
                    const int sl = s - s0;
                    const uint64_t dAh = umma_desc(aHi + 2 * sl * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dAl = umma_desc(aLo + 2 * sl * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dBh = umma_desc(bHi + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                    const uint64_t dBl = umma_desc(bLo + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                    umma_tf32(tmem, dAl, dBh, idesc, s > 0 ? 1u : 0u);
                    umma_tf32(tmem, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem, dAh, dBh, idesc, 1u);
                }
                umma_commit(bar);
            }
        }
        if (dX == nullptr) continue;
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // coalesced store (and ReLU-derivative mask read) through the free A chunk buffers: see warp_block_store16
        const int64_t rbase = row0 + (warp & 3) * 32;
        const int rows_valid = (int)(M - rbase < 32 ? M - rbase : 32);
        float* stage = reinterpret_cast<float*>(tc_smem) + warp * TC_STAGE_WARP;
        const int cbeg = warp < 4 ? 0 : split, cend = warp < 4 ? split : L.npad;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
            if (c0 < K)   // K % 4 == 0: whole float4 groups only
                warp_block_store16(stage, v, lane, dX + rbase * K + c0, RELU_IN ? X + rbase * K + c0 : nullptr, K, rows_valid, K - c0);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// backward, weight gradient:  dW[Nout][K] = dG^T . act_in(X),  db = colsum(dG)  (a ones column appended to X)
// D[m = output feature (padded to 128)][n = input feature | ones] accumulates in TMEM over ALL row tiles of the CTA;
// the reduction runs over rows, 32 at a time.  Both operands are written TRANSPOSED into the K-major canonical
// layout (element (m, row) at (row / 4) * LBO + m * 16 + (row % 4) * 4): lanes run along the rows, which makes the
// 4-byte shared-memory stores conflict-free (32 lanes -> 32 banks).
// Per-CTA partials are reduced in a fixed order by linear_wgrad_tc_reduce_kernel: no float atomics.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_WR = 32;                         // rows (= reduction depth) staged per step: 8 core columns, 4 k-steps

template <bool RELU_IN>
__global__ void __launch_bounds__(TC_THREADS, 2) linear_wgrad_tc_kernel(const float* __restrict__ X, const float* __restrict__ dG,
                                                                        int64_t M, int K, int Nout, int tmem_cols,
                                                                        float* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int npad = (K + 1 + 15) / 16 * 16;          // columns of [X | 1], padded to N % 16 == 0
    const int x_lbo = npad * 16 + 16;
    const int g_bytes = (TC_WR / 4) * TC_A_LBO, x_bytes = (TC_WR / 4) * x_lbo;
    unsigned char* sGhi = tc_smem;
    unsigned char* sGlo = sGhi + g_bytes;
    unsigned char* sXhi = sGlo + g_bytes;
    unsigned char* sXlo = sXhi + x_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool gvec = (Nout & 3) == 0;
    const int ggroups = (Nout + 3) / 4, kq = K / 4;

    for (int e = tid; e < (2 * g_bytes + 2 * x_bytes) / 16; e += TC_THREADS) reinterpret_cast<float4*>(tc_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, npad);
    const uint32_t gHi = smem_u32(sGhi), gLo = smem_u32(sGlo), xHi = smem_u32(sXhi), xLo = smem_u32(sXlo);
    uint32_t phase = 0;
    bool pending = false, first = true;

    const int64_t n_tiles = (M + TC_ROWS - 1) / TC_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int rc = 0; rc < TC_ROWS / TC_WR; ++rc) {
            const int64_t row0 = tile * TC_ROWS + rc * TC_WR;
            if (row0 >= M) break;   // CTA-uniform
            const int64_t row = row0 + lane;                       // lane = row of the 32-row chunk
            const int roff = (lane >> 2) * TC_A_LBO + (lane & 3) * 4;
            const int xoff = (lane >> 2) * x_lbo + (lane & 3) * 4;
            // all global loads of the chunk first (<= 2 + 5 float4 per thread), so they fly together and under the
            // previous chunk's MMAs; warp w owns column groups w, w + 8, ...
            constexpr int NW = TC_THREADS / 32;
            float4 gv[2], xv[5];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = warp + NW * u;
                gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < M && j < ggroups) {
                    if (gvec) gv[u] = __ldg(reinterpret_cast<const float4*>(dG + row * Nout) + j);
                    else {
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c) if (4 * j + c < Nout) t[c] = __ldg(dG + row * Nout + 4 * j + c);
                        gv[u] = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int j = warp + NW * u;
                xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < M && j <= kq) {
                    if (j < kq) {
                        xv[u] = __ldg(reinterpret_cast<const float4*>(X + row * K) + j);
                        if (RELU_IN) { xv[u].x = fmaxf(xv[u].x, 0.f); xv[u].y = fmaxf(xv[u].y, 0.f); xv[u].z = fmaxf(xv[u].z, 0.f); xv[u].w = fmaxf(xv[u].w, 0.f); }
                    } else {
                        xv[u].x = 1.0f;   // the ones column: db = colsum(dG)
                    }
                }
            }
            if (pending) {   // the MMAs reading the buffers must have completed before they are overwritten
                mbar_wait(bar, phase);
                phase ^= 1u;
                pending = false;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = warp + NW * u;
                if (j < ggroups) {
                    const float t[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float hi, lo;
                        split_tf32(t[c], hi, lo);
                        *reinterpret_cast<float*>(sGhi + roff + (4 * j + c) * 16) = hi;
                        *reinterpret_cast<float*>(sGlo + roff + (4 * j + c) * 16) = lo;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int j = warp + NW * u;
                if (j <= kq) {
                    const float t[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float hi, lo;
                        split_tf32(t[c], hi, lo);
                        *reinterpret_cast<float*>(sXhi + xoff + (4 * j + c) * 16) = hi;
                        *reinterpret_cast<float*>(sXlo + xoff + (4 * j + c) * 16) = lo;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int ks = 0; ks < TC_WR / 8; ++ks) {   // one MMA k-step = 8 rows = 2 core-matrix columns
                    const uint64_t dAh = umma_desc(gHi + 2 * ks * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dAl = umma_desc(gLo + 2 * ks * TC_A_LBO, TC_A_LBO, TC_SBO);
                    const uint64_t dBh = umma_desc(xHi + 2 * ks * x_lbo, x_lbo, TC_SBO);
                    const uint64_t dBl = umma_desc(xLo + 2 * ks * x_lbo, x_lbo, TC_SBO);
                    umma_tf32(tmem, dAl, dBh, idesc, (first && ks == 0) ? 0u : 1u);
                    umma_tf32(tmem, dAh, dBl, idesc, 1u);
                    umma_tf32(tmem, dAh, dBh, idesc, 1u);
                }
                umma_commit(bar);
            }
            first = false;
            pending = true;
        }
    }
    float* out = partial + (size_t)blockIdx.x * ((size_t)Nout * K + Nout);
    if (first) {   // this CTA owned no rows
        for (int e = tid; e < Nout * K + Nout; e += TC_THREADS) out[e] = 0.f;
    } else {
        if (pending) mbar_wait(bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = (warp & 3) * 32 + lane;            // output feature = TMEM lane
        const int split = ((npad / 16 + 1) / 2) * 16;
        const int cbeg = warp < 4 ? 0 : split, cend = warp < 4 ? split : npad;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
            if (m < Nout) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int col = c0 + c;
                    if (col < K) out[(size_t)m * K + col] = v[c];
                    else if (col == K) out[(size_t)Nout * K + m] = v[c];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient, bulk-staged (K <= 64): the same MMA schedule as linear_wgrad_tc_kernel, but the operands reach
// shared memory differently.  The kernel above lets lane = row fetch one float4 of its own row, so every load
// instruction touches 32 different 128-byte lines: the launch list of the S3G step (profiles/r03f_ncu_s3g_launches.csv)
// shows it at 217-254 us per 64 x 64 layer against 85 us of HBM time, the L1 tag stage (one line per cycle) being the
// limiter.  Here a 32-row chunk of X and of dG -- each ONE contiguous block of global memory, rows being dense -- is
// brought in by two bulk asynchronous copies (cp.async.bulk, the 1-D TMA path) into a 3-deep ring, completion on an
// mbarrier per stage; the threads then read the raw chunk column-wise (lanes along the features: conflict-free),
// apply act_in, split hi / lo and store float4s of four consecutive rows straight into the K-major canonical layout
// (element (m, row) at (row / 4) * LBO + m * 16 + (row % 4) * 4).  A CTA owns a contiguous range of whole chunks;
// the < 32 rows behind the last whole chunk are handled by one CTA of the kernel above (one more partial slot).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_WST = 3;                         // raw-chunk stages in flight per CTA

template <bool RELU_IN>
__global__ void __launch_bounds__(TC_THREADS, 2) linear_wgrad_tc_bulk_kernel(const float* __restrict__ X, const float* __restrict__ dG,
                                                                             int64_t nch, int K, int Nout, int tmem_cols,
                                                                             float* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint64_t s_full[TC_WST];
    __shared__ uint32_t s_tmem;
    const int npad = (K + 1 + 15) / 16 * 16;          // columns of [X | 1], padded to N % 16 == 0
    const int x_lbo = npad * 16 + 16;
    const int g_bytes = (TC_WR / 4) * TC_A_LBO, x_bytes = (TC_WR / 4) * x_lbo;
    const int xr_bytes = TC_WR * K * 4, gr_bytes = TC_WR * Nout * 4;   // raw chunk: multiples of 16 (K % 4 == 0; 128 Nout)
    const int st_bytes = xr_bytes + gr_bytes;
    unsigned char* sGhi = tc_smem;
    unsigned char* sGlo = sGhi + g_bytes;
    unsigned char* sXhi = sGlo + g_bytes;
    unsigned char* sXlo = sXhi + x_bytes;
    unsigned char* raw = sXlo + x_bytes;              // 128-byte aligned: g_bytes and x_bytes are multiples of 128
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t beg = nch * (int64_t)blockIdx.x / gridDim.x, end = nch * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(end - beg);

    for (int e = tid; e < (2 * g_bytes + 2 * x_bytes) / 16; e += TC_THREADS) reinterpret_cast<float4*>(tc_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
#pragma unroll
        for (int s = 0; s < TC_WST; ++s) mbar_init(smem_u32(&s_full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // the ones column of [X | 1] (db = colsum(dG)): feature K, all 32 rows of the chunk; never overwritten afterwards
    if (tid < TC_WR / 4) *reinterpret_cast<float4*>(sXhi + tid * x_lbo + K * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(TC_ROWS, npad);
    const uint32_t gHi = smem_u32(sGhi), gLo = smem_u32(sGlo), xHi = smem_u32(sXhi), xLo = smem_u32(sXlo);

    // one thread: chunk i of this CTA -> stage i % TC_WST (two bulk copies, one transaction count)
    auto issue = [&](int i) {
        const int s = i % TC_WST;
        const uint32_t fb = smem_u32(&s_full[s]);
        const uint32_t dst = smem_u32(raw + s * st_bytes);
        const float* xs = X + (beg + i) * (int64_t)(TC_WR * K);
        const float* gs = dG + (beg + i) * (int64_t)(TC_WR * Nout);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(fb), "r"((uint32_t)st_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(xs), "r"((uint32_t)xr_bytes), "r"(fb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst + (uint32_t)xr_bytes), "l"(gs), "r"((uint32_t)gr_bytes), "r"(fb) : "memory");
    };
    if (tid == 0)
        for (int i = 0; i < min(TC_WST, n_my); ++i) issue(i);

    uint32_t phase = 0;
    bool pending = false;
    for (int i = 0; i < n_my; ++i) {
        const int s = i % TC_WST;
        mbar_wait(smem_u32(&s_full[s]), (uint32_t)((i / TC_WST) & 1));   // the raw chunk has landed
        if (pending) {   // the previous chunk's MMAs have read the operand buffers
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        const float* rx = reinterpret_cast<const float*>(raw + s * st_bytes);
        const float* rg = reinterpret_cast<const float*>(raw + s * st_bytes + xr_bytes);
        for (int e = tid; e < (TC_WR / 4) * K; e += TC_THREADS) {
            const int rq = e / K, m = e - rq * K;            // rows 4 rq .. 4 rq + 3 of feature m
            float4 x = make_float4(rx[(4 * rq + 0) * K + m], rx[(4 * rq + 1) * K + m], rx[(4 * rq + 2) * K + m], rx[(4 * rq + 3) * K + m]);
            if (RELU_IN) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            float4 hi, lo;
            split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(sXhi + rq * x_lbo + m * 16) = hi;
            *reinterpret_cast<float4*>(sXlo + rq * x_lbo + m * 16) = lo;
        }
        for (int e = tid; e < (TC_WR / 4) * Nout; e += TC_THREADS) {
            const int rq = e / Nout, m = e - rq * Nout;
            const float4 g = make_float4(rg[(4 * rq + 0) * Nout + m], rg[(4 * rq + 1) * Nout + m], rg[(4 * rq + 2) * Nout + m], rg[(4 * rq + 3) * Nout + m]);
            float4 hi, lo;
            split_tf32(g.x, hi.x, lo.x); split_tf32(g.y, hi.y, lo.y); split_tf32(g.z, hi.z, lo.z); split_tf32(g.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(sGhi + rq * TC_A_LBO + m * 16) = hi;
            *reinterpret_cast<float4*>(sGlo + rq * TC_A_LBO + m * 16) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();   // operands complete; every thread is done reading raw stage s
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ks = 0; ks < TC_WR / 8; ++ks) {   // one MMA k-step = 8 rows = 2 core-matrix columns
                const uint64_t dAh = umma_desc(gHi + 2 * ks * TC_A_LBO, TC_A_LBO, TC_SBO);
                const uint64_t dAl = umma_desc(gLo + 2 * ks * TC_A_LBO, TC_A_LBO, TC_SBO);
                const uint64_t dBh = umma_desc(xHi + 2 * ks * x_lbo, x_lbo, TC_SBO);
                const uint64_t dBl = umma_desc(xLo + 2 * ks * x_lbo, x_lbo, TC_SBO);
                umma_tf32(tmem, dAl, dBh, idesc, (i == 0 && ks == 0) ? 0u : 1u);
                umma_tf32(tmem, dAh, dBl, idesc, 1u);
                umma_tf32(tmem, dAh, dBh, idesc, 1u);
            }
            umma_commit(bar);
            if (i + TC_WST < n_my) issue(i + TC_WST);   // refill the stage just consumed
        }
        pending = true;
    }
    float* out = partial + (size_t)blockIdx.x * ((size_t)Nout * K + Nout);
    if (n_my == 0) {
        for (int e = tid; e < Nout * K + Nout; e += TC_THREADS) out[e] = 0.f;
    } else {
        mbar_wait(bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = (warp & 3) * 32 + lane;            // output feature = TMEM lane
        const int split = ((npad / 16 + 1) / 2) * 16;
        const int cbeg = warp < 4 ? 0 : split, cend = warp < 4 ? split : npad;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
            if (m < Nout) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int col = c0 + c;
                    if (col < K) out[(size_t)m * K + col] = v[c];
                    else if (col == K) out[(size_t)Nout * K + m] = v[c];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(tmem_cols) : "memory");
}

__global__ void linear_wgrad_tc_reduce_kernel(const float* __restrict__ partial, int nparts, int count,
                                              float* __restrict__ dW, float* __restrict__ db, int nW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * count + i];
    if (i < nW) dW[i] = s; else db[i - nW] = s;
}

int tc_grid(int64_t M) {
    const int64_t tiles = (M + TC_ROWS - 1) / TC_ROWS;
    return (int)(tiles < EMD_NUM_SMS * 2 ? (tiles > 0 ? tiles : 1) : EMD_NUM_SMS * 2);
}
int tmem_cols_for(int npad) { return npad <= 32 ? 32 : npad <= 64 ? 64 : npad <= 128 ? 128 : 256; }

}  // namespace

// Tensor-core forward of one Linear layer (see the header of this file).  K % 4 == 0, K <= 136, Nout <= 64.
extern "C" int emd_linear_fwd_tc(const float* X, const float* W, const float* b, int64_t M, int K, int Nout, int relu_in,
                                 int relu_out, float* Y, cudaStream_t stream) {
    EMD_CHECK_ARG(M >= 0 && K >= 4 && K <= TC_KMAX && (K % 4) == 0, "linear_fwd_tc: K must be a multiple of 4 in [4, %d]", TC_KMAX);
    EMD_CHECK_ARG(Nout >= 1 && Nout <= TC_NMAX, "linear_fwd_tc: Nout must be in [1, %d]", TC_NMAX);
    if (!emd_aligned(X, 16) || !emd_aligned(W, 16) || ((Nout & 3) == 0 && !emd_aligned(Y, 16))) {
        emd_set_error("linear_fwd_tc: X, W (and Y when Nout % 4 == 0) must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    if (M == 0) return EMD_OK;
    const TcLayout L(K, Nout);
    const size_t smem = L.total();
    const int64_t n_tiles = (M + TC_ROWS - 1) / TC_ROWS;
    const unsigned grid = (unsigned)(n_tiles < 2 * EMD_NUM_SMS ? n_tiles : 2 * EMD_NUM_SMS);
#define EMD_TC_LAUNCH(RI, RO)                                                                                          \
    do {                                                                                                               \
        cudaFuncSetAttribute(linear_fwd_tc_kernel<RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        EMD_LAUNCH(EK_MLP_FWD, stream, (linear_fwd_tc_kernel<RI, RO><<<grid, TC_THREADS, smem, stream>>>(X, W, b, M, K, Nout, Y))); \
    } while (0)
    if (relu_in && relu_out) EMD_TC_LAUNCH(true, true);
    else if (relu_in) EMD_TC_LAUNCH(true, false);
    else if (relu_out) EMD_TC_LAUNCH(false, true);
    else EMD_TC_LAUNCH(false, false);
#undef EMD_TC_LAUNCH
    EMD_CHECK_LAUNCH("linear_fwd_tc");
    return EMD_OK;
}

extern "C" size_t emd_linear_bwd_workspace_bytes(int64_t M, int K, int Nout);

// Tensor-core VJP of one Linear layer: same contract as emd_linear_bwd (dX may be NULL), same workspace size.
extern "C" int emd_linear_bwd_tc(const float* X, const float* W, const float* Y, const float* dY, int64_t M, int K,
                                 int Nout, int relu_in, int relu_out, float* dX, float* dW, float* db, void* workspace,
                                 size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(M >= 0 && K >= 4 && K <= TC_KMAX && (K % 4) == 0, "linear_bwd_tc: K must be a multiple of 4 in [4, %d]", TC_KMAX);
    EMD_CHECK_ARG(Nout >= 1 && Nout <= TC_NMAX, "linear_bwd_tc: Nout must be in [1, %d]", TC_NMAX);
    if (ws_bytes < emd_linear_bwd_workspace_bytes(M, K, Nout)) { emd_set_error("linear_bwd_tc: workspace too small"); return EMD_ERR_WORKSPACE; }
    if (!emd_aligned(X, 16) || !emd_aligned(dY, 16) || !emd_aligned(Y, 16) || !emd_aligned(workspace, 16) || (dX && !emd_aligned(dX, 16))) {
        emd_set_error("linear_bwd_tc: X, Y, dY, dX and the workspace must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int count = Nout * K + Nout;
    if (M == 0) {
        cudaMemsetAsync(dW, 0, (size_t)Nout * K * sizeof(float), stream);
        cudaMemsetAsync(db, 0, (size_t)Nout * sizeof(float), stream);
        return EMD_OK;
    }
    float* dG = reinterpret_cast<float*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((size_t)M * Nout * sizeof(float) + 255) / 256 * 256);
    const int grid = tc_grid(M);
    // without an output ReLU the masked gradient IS dY: nothing to write, the weight gradient reads dY
    const float* dGr = relu_out ? dG : dY;
    if (relu_out || dX) {
        const TcLayout L(Nout, K);
        const size_t smem = L.total();
        const int tcols = tmem_cols_for(L.npad);
#define EMD_TC_DG(RI, RO)                                                                                              \
    do {                                                                                                               \
        cudaFuncSetAttribute(linear_dgrad_tc_kernel<RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        EMD_LAUNCH(EK_MLP_BWD, stream, (linear_dgrad_tc_kernel<RI, RO><<<grid, TC_THREADS, smem, stream>>>(X, W, Y, dY, M, K, Nout, tcols, dG, dX))); \
    } while (0)
        if (relu_in && relu_out) EMD_TC_DG(true, true);
        else if (relu_in) EMD_TC_DG(true, false);
        else if (relu_out) EMD_TC_DG(false, true);
        else EMD_TC_DG(false, false);
#undef EMD_TC_DG
    }
    int nparts = grid;
    {
        const int npad = (K + 1 + 15) / 16 * 16;
        const size_t smem = 2 * (size_t)(TC_WR / 4) * (TC_A_LBO + npad * 16 + 16);
        const int tcols = tmem_cols_for(npad);
        const int64_t nch = M / TC_WR;                      // whole 32-row chunks
        const float* Xt = X;
        const float* Gt = dGr;
        int64_t Mt = M;
        int tgrid = grid;
        float* tpart = partial;
        if (K <= 64 && nch > 0) {
            // bulk-staged kernel over the whole chunks; the rows behind them (< 32) go through the row-wise kernel below
            const size_t bsmem = smem + (size_t)TC_WST * TC_WR * (K + Nout) * sizeof(float);
            const int bgrid = (int)(nch < grid ? nch : grid);
            if (relu_in) {
                cudaFuncSetAttribute(linear_wgrad_tc_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
                EMD_LAUNCH(EK_MLP_BWD, stream, (linear_wgrad_tc_bulk_kernel<true><<<bgrid, TC_THREADS, bsmem, stream>>>(X, dGr, nch, K, Nout, tcols, partial)));
            } else {
                cudaFuncSetAttribute(linear_wgrad_tc_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
                EMD_LAUNCH(EK_MLP_BWD, stream, (linear_wgrad_tc_bulk_kernel<false><<<bgrid, TC_THREADS, bsmem, stream>>>(X, dGr, nch, K, Nout, tcols, partial)));
            }
            Xt = X + nch * TC_WR * K;
            Gt = dGr + nch * TC_WR * Nout;
            Mt = M - nch * TC_WR;
            tgrid = 1;
            tpart = partial + (size_t)bgrid * count;
            nparts = bgrid + (Mt > 0 ? 1 : 0);
        }
        if (Mt > 0) {
            if (relu_in) {
                cudaFuncSetAttribute(linear_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                EMD_LAUNCH(EK_MLP_BWD, stream, (linear_wgrad_tc_kernel<true><<<tgrid, TC_THREADS, smem, stream>>>(Xt, Gt, Mt, K, Nout, tcols, tpart)));
            } else {
                cudaFuncSetAttribute(linear_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                EMD_LAUNCH(EK_MLP_BWD, stream, (linear_wgrad_tc_kernel<false><<<tgrid, TC_THREADS, smem, stream>>>(Xt, Gt, Mt, K, Nout, tcols, tpart)));
            }
        }
    }
    EMD_LAUNCH(EK_MLP_BWD, stream, (linear_wgrad_tc_reduce_kernel<<<(count + 127) / 128, 128, 0, stream>>>(partial, nparts, count, dW, db, Nout * K)));
    EMD_CHECK_LAUNCH("linear_bwd_tc");
    return EMD_OK;
}
