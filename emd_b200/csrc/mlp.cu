// K1d (building block): the Linear layers of the S3Gaussian EMD deformation network
// (S3Gaussian/scene/deformation.py:100-185: feature_out / pos_deform / opacity_deform / shs_deform / dino_head,
// coarse and fine copies), forward and backward, with the surrounding ReLUs fused in:
//
//     Y[M,Nout] = act_out( act_in(X[M,K]) . W[Nout,K]^T + b )           act = ReLU or identity
//
// M = number of Gaussians (1-2 M), K in {4, 64, 132}, Nout in {1, 3, 48, 64}: tall-skinny GEMMs.
// This file is the fp32 SIMT implementation (exact fp32 accumulation; the parity reference for the
// tensor-core path): CTA = 128 rows, one row per thread, the whole weight matrix resident in shared memory
// as W^T so every inner-loop operand read is a broadcast LDS.128, X staged in coalesced 32-column chunks.
// Weight gradients are reduced in a fixed order: per-CTA partials (persistent CTAs) -> one reduction kernel.
// No float atomics.
#include "common.cuh"

namespace {

constexpr int ML_ROWS = 128;      // rows per CTA tile == threads per CTA
constexpr int ML_KC = 32;         // X columns staged per chunk
constexpr int ML_NMAX = 64;       // output columns handled per pass
constexpr int ML_KMAX = 192;

// out[r][n] (+)= sum_k in[r][k] * B[k][n],  B in shared memory with leading dimension ldb (>= n_cols, mult of 4)
// in: global [M, KI] (row stride KI).  Each thread owns one row; acc[ML_NMAX] in registers.
// mask (optional, same shape as in): entries where mask <= 0 are treated as 0 (ReLU derivative).
template <bool RELU_IN>
__device__ __forceinline__ void rows_times_smem(const float* __restrict__ in, const float* __restrict__ mask,
                                                int64_t row0, int64_t M, int KI, const float* __restrict__ sB, int ldb,
                                                float* sX /*[128][33]*/, float (&acc)[ML_NMAX]) {
    const int tid = threadIdx.x;
    for (int k0 = 0; k0 < KI; k0 += ML_KC) {
        const int kc = min(ML_KC, KI - k0);
        __syncthreads();
        // coalesced stage of X[row0:row0+128, k0:k0+kc] -> sX[r][c]
        for (int e = tid; e < ML_ROWS * kc; e += ML_ROWS) {
            const int r = e / kc, c = e - r * kc;
            const int64_t row = row0 + r;
            float v = row < M ? __ldg(in + row * KI + k0 + c) : 0.f;
            if (RELU_IN) v = fmaxf(v, 0.f);
            if (mask != nullptr && row < M && !(__ldg(mask + row * KI + k0 + c) > 0.f)) v = 0.f;
            sX[r * (ML_KC + 1) + c] = v;
        }
        __syncthreads();
        for (int c = 0; c < kc; ++c) {
            const float x = sX[tid * (ML_KC + 1) + c];
            const float4* b4 = reinterpret_cast<const float4*>(sB + (size_t)(k0 + c) * ldb);
#pragma unroll
            for (int n4 = 0; n4 < ML_NMAX / 4; ++n4) {
                const float4 w = b4[n4];
                acc[n4 * 4 + 0] += x * w.x; acc[n4 * 4 + 1] += x * w.y;
                acc[n4 * 4 + 2] += x * w.z; acc[n4 * 4 + 3] += x * w.w;
            }
        }
    }
}

// forward: smem holds W^T padded to [K][64]
template <bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(ML_ROWS) linear_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                             const float* __restrict__ b, int64_t M, int K, int Nout,
                                                             float* __restrict__ Y) {
    extern __shared__ __align__(16) float smem[];
    float* sB = smem;                       // [K][64]
    float* sX = smem + (size_t)K * ML_NMAX;  // [128][33]
    for (int e = threadIdx.x; e < K * ML_NMAX; e += ML_ROWS) {
        const int k = e / ML_NMAX, n = e - k * ML_NMAX;
        sB[e] = n < Nout ? __ldg(W + (size_t)n * K + k) : 0.f;
    }
    const int64_t n_tiles = (M + ML_ROWS - 1) / ML_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * ML_ROWS, row = row0 + threadIdx.x;
        float acc[ML_NMAX];
#pragma unroll
        for (int n = 0; n < ML_NMAX; ++n) acc[n] = 0.f;
        rows_times_smem<RELU_IN>(X, nullptr, row0, M, K, sB, ML_NMAX, sX, acc);
        __syncthreads();
        // bias + activation, staged through sX (reused as [128][33] x 2 halves) for coalesced stores
        for (int half = 0; half < 2; ++half) {
            if (half * 32 >= Nout) break;
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const int col = half * 32 + n;
                float v = acc[col] + (col < Nout ? __ldg(b + col) : 0.f);
                if (RELU_OUT) v = fmaxf(v, 0.f);
                sX[threadIdx.x * (ML_KC + 1) + n] = v;
            }
            __syncthreads();
            const int nc = min(32, Nout - half * 32);
            for (int e = threadIdx.x; e < ML_ROWS * nc; e += ML_ROWS) {
                const int r = e / nc, c = e - r * nc;
                if (row0 + r < M) Y[(row0 + r) * Nout + half * 32 + c] = sX[r * (ML_KC + 1) + c];
            }
            __syncthreads();
        }
        (void)row;
    }
}

// dX[M,K] = (dY (.) relu'(Y)) . W   (W[Nout][K] row-major is already the [in][out] operand), masked by X > 0
// when the layer applied ReLU to its input.  Output columns processed 64 at a time.
template <bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(ML_ROWS) linear_dgrad_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                               const float* __restrict__ Y, const float* __restrict__ dY,
                                                               int64_t M, int K, int Nout, float* __restrict__ dG /*[M,Nout]*/,
                                                               float* __restrict__ dX) {
    extern __shared__ __align__(16) float smem[];
    const int Kp = (K + 63) / 64 * 64;
    float* sB = smem;                         // [Nout][Kp]
    float* sX = smem + (size_t)Nout * Kp;      // [128][33]
    for (int e = threadIdx.x; e < Nout * Kp; e += ML_ROWS) {
        const int n = e / Kp, k = e - n * Kp;
        sB[e] = k < K ? __ldg(W + (size_t)n * K + k) : 0.f;
    }
    const int64_t n_tiles = (M + ML_ROWS - 1) / ML_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * ML_ROWS;
        // masked upstream gradient dG = dY * (Y > 0) is materialised once for the weight-gradient kernel (the
        // GEMM below re-derives it from the read-only dY / Y while staging, never reading dG back)
        for (int e = threadIdx.x; e < ML_ROWS * Nout; e += ML_ROWS) {
            const int64_t idx = row0 * Nout + e;
            if (idx < M * Nout) {
                float g = __ldg(dY + idx);
                if (RELU_OUT && !(__ldg(Y + idx) > 0.f)) g = 0.f;
                dG[idx] = g;
            }
        }
        __syncthreads();
        if (dX == nullptr) continue;
        for (int kb = 0; kb < K; kb += ML_NMAX) {
            float acc[ML_NMAX];
#pragma unroll
            for (int n = 0; n < ML_NMAX; ++n) acc[n] = 0.f;
            rows_times_smem<false>(dY, RELU_OUT ? Y : nullptr, row0, M, Nout, sB + kb, Kp, sX, acc);
            __syncthreads();
            for (int half = 0; half < 2; ++half) {
                const int c0 = kb + half * 32;
                if (c0 >= K) break;
#pragma unroll
                for (int n = 0; n < 32; ++n) sX[threadIdx.x * (ML_KC + 1) + n] = acc[half * 32 + n];
                __syncthreads();
                const int nc = min(32, K - c0);
                for (int e = threadIdx.x; e < ML_ROWS * nc; e += ML_ROWS) {
                    const int r = e / nc, c = e - r * nc;
                    if (row0 + r < M) {
                        float v = sX[r * (ML_KC + 1) + c];
                        if (RELU_IN && !(__ldg(X + (row0 + r) * K + c0 + c) > 0.f)) v = 0.f;
                        dX[(row0 + r) * K + c0 + c] = v;
                    }
                }
                __syncthreads();
            }
        }
    }
}

// per-CTA partial of dW[Nout][K] = dG^T . act_in(X) and db = colsum(dG); persistent CTAs, fixed tile order.
// thread t owns output rows n = (t / 8) * 4 .. +3 (16 groups -> 64 rows) and, per 32-column chunk, columns
// k = (t % 8) * 4 .. +3.
template <bool RELU_IN>
__global__ void __launch_bounds__(ML_ROWS) linear_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dG,
                                                               int64_t M, int K, int Nout,
                                                               float* __restrict__ partial /*[grid][Nout*K + Nout]*/) {
    extern __shared__ __align__(16) float smem[];
    float (*sG)[ML_NMAX + 1] = reinterpret_cast<float (*)[ML_NMAX + 1]>(smem);
    float (*sXc)[ML_KC + 1] = reinterpret_cast<float (*)[ML_KC + 1]>(smem + ML_ROWS * (ML_NMAX + 1));
    const int tid = threadIdx.x;
    const int n0 = (tid >> 3) * 4, kq = (tid & 7) * 4;
    constexpr int MAXC = ML_KMAX / ML_KC;  // 6 chunks
    float acc[MAXC][16];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c][i] = 0.f;
    float bacc = 0.f;  // threads 0..63 own db[tid]
    const int nchunks = (K + ML_KC - 1) / ML_KC;
    const int64_t n_tiles = (M + ML_ROWS - 1) / ML_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * ML_ROWS;
        __syncthreads();
        for (int e = tid; e < ML_ROWS * ML_NMAX; e += ML_ROWS) {
            const int r = e / ML_NMAX, n = e - r * ML_NMAX;
            sG[r][n] = (n < Nout && row0 + r < M) ? __ldg(dG + (row0 + r) * Nout + n) : 0.f;
        }
        __syncthreads();
        if (tid < ML_NMAX) {
            float s = 0.f;
            for (int r = 0; r < ML_ROWS; ++r) s += sG[r][tid];
            bacc += s;
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c >= nchunks) break;
            const int k0 = c * ML_KC, kc = min(ML_KC, K - k0);
            __syncthreads();
            for (int e = tid; e < ML_ROWS * ML_KC; e += ML_ROWS) {
                const int r = e / ML_KC, cc = e - r * ML_KC;
                float v = (cc < kc && row0 + r < M) ? __ldg(X + (row0 + r) * K + k0 + cc) : 0.f;
                if (RELU_IN) v = fmaxf(v, 0.f);
                sXc[r][cc] = v;
            }
            __syncthreads();
            for (int r = 0; r < ML_ROWS; ++r) {
                const float g0 = sG[r][n0], g1 = sG[r][n0 + 1], g2 = sG[r][n0 + 2], g3 = sG[r][n0 + 3];
                const float x0 = sXc[r][kq], x1 = sXc[r][kq + 1], x2 = sXc[r][kq + 2], x3 = sXc[r][kq + 3];
                acc[c][0] += g0 * x0; acc[c][1] += g0 * x1; acc[c][2] += g0 * x2; acc[c][3] += g0 * x3;
                acc[c][4] += g1 * x0; acc[c][5] += g1 * x1; acc[c][6] += g1 * x2; acc[c][7] += g1 * x3;
                acc[c][8] += g2 * x0; acc[c][9] += g2 * x1; acc[c][10] += g2 * x2; acc[c][11] += g2 * x3;
                acc[c][12] += g3 * x0; acc[c][13] += g3 * x1; acc[c][14] += g3 * x2; acc[c][15] += g3 * x3;
            }
        }
    }
    float* out = partial + (size_t)blockIdx.x * ((size_t)Nout * K + Nout);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c >= nchunks) break;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + i, k = c * ML_KC + kq + j;
                if (n < Nout && k < K) out[(size_t)n * K + k] = acc[c][i * 4 + j];
            }
    }
    if (tid < Nout) out[(size_t)Nout * K + tid] = bacc;
}

__global__ void linear_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int count,
                                           float* __restrict__ dW, float* __restrict__ db, int nW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * count + i];
    if (i < nW) dW[i] = s; else db[i - nW] = s;
}

int wgrad_grid(int64_t M) {
    const int64_t tiles = (M + ML_ROWS - 1) / ML_ROWS;
    return (int)(tiles < EMD_NUM_SMS * 2 ? (tiles > 0 ? tiles : 1) : EMD_NUM_SMS * 2);
}

}  // namespace

extern "C" size_t emd_linear_bwd_workspace_bytes(int64_t M, int K, int Nout) {
    // masked upstream gradient [M,Nout] + per-CTA weight-gradient partials (+ one slot: the tensor-core path's tail CTA)
    return ((size_t)M * Nout * sizeof(float) + 255) / 256 * 256 +
           ((size_t)wgrad_grid(M) + 1) * ((size_t)Nout * K + Nout) * sizeof(float) + 256;
}

extern "C" int emd_linear_fwd(const float* X, const float* W, const float* b, int64_t M, int K, int Nout, int relu_in,
                              int relu_out, float* Y, cudaStream_t stream) {
    EMD_CHECK_ARG(K >= 1 && K <= ML_KMAX && Nout >= 1 && Nout <= ML_NMAX, "linear_fwd: need K <= %d, Nout <= %d", ML_KMAX, ML_NMAX);
    if (M == 0) return EMD_OK;
    const size_t smem = ((size_t)K * ML_NMAX + ML_ROWS * (ML_KC + 1)) * sizeof(float);
    const int64_t tiles = (M + ML_ROWS - 1) / ML_ROWS;
    const unsigned grid = (unsigned)(tiles < EMD_NUM_SMS * 8 ? tiles : EMD_NUM_SMS * 8);
#define EMD_FWD(RI, RO)                                                                                          \
    do {                                                                                                         \
        cudaFuncSetAttribute(linear_fwd_kernel<RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        EMD_LAUNCH(EK_MLP_FWD, stream, linear_fwd_kernel<RI, RO><<<grid, ML_ROWS, smem, stream>>>(X, W, b, M, K, Nout, Y)); \
    } while (0)
    if (relu_in && relu_out) EMD_FWD(true, true);
    else if (relu_in) EMD_FWD(true, false);
    else if (relu_out) EMD_FWD(false, true);
    else EMD_FWD(false, false);
#undef EMD_FWD
    EMD_CHECK_LAUNCH("linear_fwd");
    return EMD_OK;
}

// dX may be NULL (first layer of a branch whose input needs no gradient).
extern "C" int emd_linear_bwd(const float* X, const float* W, const float* Y, const float* dY, int64_t M, int K,
                              int Nout, int relu_in, int relu_out, float* dX, float* dW, float* db, void* workspace,
                              size_t ws_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(K >= 1 && K <= ML_KMAX && Nout >= 1 && Nout <= ML_NMAX, "linear_bwd: need K <= %d, Nout <= %d", ML_KMAX, ML_NMAX);
    if (ws_bytes < emd_linear_bwd_workspace_bytes(M, K, Nout)) { emd_set_error("linear_bwd: workspace too small"); return EMD_ERR_WORKSPACE; }
    const int count = Nout * K + Nout;
    if (M == 0) {
        cudaMemsetAsync(dW, 0, (size_t)Nout * K * sizeof(float), stream);
        cudaMemsetAsync(db, 0, (size_t)Nout * sizeof(float), stream);
        return EMD_OK;
    }
    float* dG = reinterpret_cast<float*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((size_t)M * Nout * sizeof(float) + 255) / 256 * 256);
    const int Kp = (K + 63) / 64 * 64;
    const size_t smem = ((size_t)Nout * Kp + ML_ROWS * (ML_KC + 1)) * sizeof(float);
    const int64_t tiles = (M + ML_ROWS - 1) / ML_ROWS;
    const unsigned grid = (unsigned)(tiles < EMD_NUM_SMS * 8 ? tiles : EMD_NUM_SMS * 8);
#define EMD_DG(RI, RO)                                                                                              \
    do {                                                                                                            \
        cudaFuncSetAttribute(linear_dgrad_kernel<RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        EMD_LAUNCH(EK_MLP_BWD, stream, linear_dgrad_kernel<RI, RO><<<grid, ML_ROWS, smem, stream>>>(X, W, Y, dY, M, K, Nout, dG, dX)); \
    } while (0)
    if (relu_in && relu_out) EMD_DG(true, true);
    else if (relu_in) EMD_DG(true, false);
    else if (relu_out) EMD_DG(false, true);
    else EMD_DG(false, false);
#undef EMD_DG
    const int wg = wgrad_grid(M);
    const size_t wsmem = (size_t)ML_ROWS * (ML_NMAX + 1 + ML_KC + 1) * sizeof(float);
    if (relu_in) {
        cudaFuncSetAttribute(linear_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        EMD_LAUNCH(EK_MLP_BWD, stream, linear_wgrad_kernel<true><<<wg, ML_ROWS, wsmem, stream>>>(X, dG, M, K, Nout, partial));
    } else {
        cudaFuncSetAttribute(linear_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        EMD_LAUNCH(EK_MLP_BWD, stream, linear_wgrad_kernel<false><<<wg, ML_ROWS, wsmem, stream>>>(X, dG, M, K, Nout, partial));
    }
    EMD_LAUNCH(EK_MLP_BWD, stream, linear_wgrad_reduce_kernel<<<(count + 127) / 128, 128, 0, stream>>>(partial, wg, count, dW, db, Nout * K));
    EMD_CHECK_LAUNCH("linear_bwd");
    return EMD_OK;
}
