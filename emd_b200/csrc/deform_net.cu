// K1g -- the non-rigid deformation network of OmniRe's DeformableNodes:
// ConditionalDeformNetwork (OmniRe/models/modules.py:411-457; D = 8 hidden layers of W = 256, positional
// encodings of the point and the time, a 16-float instance embedding, one skip connection) as queried by
// DeformableNodes.get_deformation / get_gaussians (OmniRe/models/nodes/deformable.py:35-77).
//
// Pieces (each a C-ABI entry point, orchestrated by emd_b200/deformable.py):
//   emd_deform_input_fwd    x = means / instance_height * 2, [x, sin/cos(2^f x)], [t, sin/cos(2^f t)], embedding[id]
//                           written straight into the layer-0 operand AND into the head of the skip-layer operand
//                           (the reference materialises three concatenations per call)
//   emd_dense_fwd / _bwd    Y = act(X W^T + b) on strided operands (so a layer can write into / read from a column
//                           window of a wider buffer: the skip concat costs nothing), dX restricted to a column
//                           window (the input gradient is only needed for the 16 embedding columns: the point is
//                           detached at deformable.py:43 and the time is a constant), ReLU mask of the producer
//                           fused into the dgrad epilogue, weight gradient as split-K partials reduced in a fixed
//                           order, bias gradient as fixed-order column sums.  No float atomics: bit-reproducible.
//   emd_deform_apply_fwd/_bwd   means + d_xyz, normalize(quats) + d_quat   (deformable.py:57-68, vanilla.py:142-146)
//   emd_deform_embed_grad   per-instance fixed-order sum of the embedding-column gradients (the gather's VJP)
//
// The GEMM here is fp32 SIMT (exact fp32 accumulation like the reference's cuBLAS SGEMM with TF32 off): 128x128x16
// CTA tiles, 8x8 register micro-tiles split 4+4 so every shared-memory read is a conflict-free LDS.128, global
// loads of the next k-tile in flight under the FMAs of the current one.  FP32-pipe bound; moving these 256-wide
// layers onto tcgen05 (3xTF32 like csrc/mlp_tc.cu, which is limited to K <= 136, Nout <= 64) is the open step.
#include <stdlib.h>

#include "common.cuh"
#include "dense_math.cuh"

#ifndef DG_DEFAULT_MINB
#define DG_DEFAULT_MINB 2
#endif

// experimental tensor-core path (deform_net_tc.cu): 1 when it took the call (status in *rc), 0 -> SIMT path below.
// Off unless emd_dense_set_tc(1) / EMD_DENSE_TC=1.
int emd_dense_tc_try(int dgrad, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* mask,
                     int64_t ldmask, float* Y, int64_t ldy, int64_t M, int K, int N, int relu, cudaStream_t stream, int* rc);

size_t emd_dense_tc_wgrad_partial_floats(int64_t M, int K, int Nout);
int emd_dense_tc_try_wgrad(const float* X, int64_t ldx, const float* dZ, int64_t lddz, int64_t M, int K, int Nout, float* partial,
                           int* splits_out, cudaStream_t stream, int* rc);

namespace {

// C[m,n] = sum_k A(m,k) B(k,n) -- tile logic in dense_math.cuh (shared with the host emulation the CPU tests run)
// MINB = resident CTAs per SM the register allocation targets: 2 (128 registers; two CTAs hide each other's barriers)
// or 1 (no register cap: no spills or rematerialised addresses in the k-loop, but 8 warps per SM).
template <bool TA, bool TB, int MINB>
__global__ void __launch_bounds__(DG_THREADS, MINB) sgemm_kernel(const GemmArgs g) {
    // two shared-memory stages: tile t lives in stage t & 1, so ONE barrier per k-tile suffices (a thread can only
    // overwrite stage s in iteration t + 1 after every thread passed the barrier of iteration t, i.e. finished reading
    // stage s in iteration t - 1)
    __shared__ __align__(16) float As[2][DG_BK][DG_PITCH];
    __shared__ __align__(16) float Bs[2][DG_BK][DG_PITCH];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * DG_BM;
    const int64_t n0 = (int64_t)blockIdx.y * DG_BN;
    const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
    const int64_t kend = min(g.K, kbeg + g.k_per_split);
    float* __restrict__ C = g.C + (int64_t)blockIdx.z * g.split_stride;

    float ra[8], rb[8];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    GemmLoadState S;
    gemm_prepare<TA, TB>(g, tid, m0, n0, kbeg, S);
    if (kbeg < kend) gemm_load_tile<TA, TB>(S, kbeg, kend, ra, rb);
    int stage = 0;
    for (int64_t k0 = kbeg; k0 < kend; k0 += DG_BK, stage ^= 1) {
        gemm_store<TA, TB>(tid, ra, rb, As[stage], Bs[stage]);
        __syncthreads();
        if (k0 + DG_BK < kend) {   // next tile's loads in flight under the FMAs of this one
            gemm_advance(S);
            gemm_load_tile<TA, TB>(S, k0 + DG_BK, kend, ra, rb);
        }
        gemm_compute(tid, As[stage], Bs[stage], acc);
    }
    gemm_epilogue(g, tid, m0, n0, C, acc);
}

// out[i] = sum over splits of partial[s][i], fixed order
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ partial, int64_t n, int splits,
                                                           int64_t stride, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(int64_t)s * stride + i];
    out[i] = acc;
}

// partial[s][c] = sum of Z[r, c] over the rows of chunk s (threads = columns: coalesced), fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ Z, int64_t ldz, int64_t M, int N,
                                                     int64_t rows_per_chunk, float* __restrict__ partial) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
    for (int c = threadIdx.x; c < N; c += 256) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc += __ldg(Z + r * ldz + c);
        partial[(int64_t)blockIdx.x * N + c] = acc;
    }
}

// tuning switch (measurement only): EMD_SGEMM_MINB=1|2 selects the register-allocation flavour; default below
int sgemm_minb() {
    static const int v = [] {
        const char* e = getenv("EMD_SGEMM_MINB");
        return (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : DG_DEFAULT_MINB;
    }();
    return v;
}

template <bool TA, bool TB>
void launch_sgemm(const GemmArgs& g, int splits, cudaStream_t stream) {
    dim3 grid((unsigned)emd_cdiv(g.M, DG_BM), (unsigned)emd_cdiv(g.N, DG_BN), (unsigned)splits);
    if (sgemm_minb() == 1) sgemm_kernel<TA, TB, 1><<<grid, DG_THREADS, 0, stream>>>(g);
    else sgemm_kernel<TA, TB, 2><<<grid, DG_THREADS, 0, stream>>>(g);
}

}  // namespace

// Y[M, 0:Nout] (row stride ldy) = act(X[M, 0:K] (row stride ldx) . W[Nout,K]^T + b);  b may be NULL.
// Replaces one nn.Linear (+ F.relu) of ConditionalDeformNetwork.forward (OmniRe/models/modules.py:438-455).
extern "C" int emd_dense_fwd(const float* X, int64_t ldx, const float* W, const float* b, int64_t M, int K, int Nout,
                             int relu_out, float* Y, int64_t ldy, cudaStream_t stream) {
    EMD_CHECK_ARG(M >= 0 && K >= 1 && Nout >= 1 && ldx >= K && ldy >= Nout, "emd_dense_fwd: M=%lld K=%d Nout=%d ldx=%lld ldy=%lld",
                  (long long)M, K, Nout, (long long)ldx, (long long)ldy);
    if (M == 0) return EMD_OK;
    EMD_CHECK_ARG(X && W && Y, "emd_dense_fwd: null argument");
    EMD_CHECK_ARG(emd_cdiv(Nout, DG_BN) <= 65535, "emd_dense_fwd: Nout=%d too wide", Nout);
    int rc_tc = EMD_OK;
    if (emd_dense_tc_try(0, X, ldx, W, K, b, nullptr, 0, Y, ldy, M, K, Nout, relu_out, stream, &rc_tc)) return rc_tc;
    const GemmArgs g = dense_fwd_args(X, ldx, W, b, M, K, Nout, relu_out, Y, ldy);
    EMD_LAUNCH(EK_DENSE_FWD, stream, (launch_sgemm<false, true>(g, 1, stream)));
    EMD_CHECK_LAUNCH("emd_dense_fwd");
    return EMD_OK;
}

static size_t dense_wgrad_partial_floats(int64_t M, int K, int Nout) {
    const DenseSplit s = dense_split(M, K, Nout);
    const size_t simt = (size_t)s.splits * (size_t)Nout * (size_t)K;
    const size_t tc = emd_dense_tc_wgrad_partial_floats(M, K, Nout);     // 0 unless the experimental path is selected
    return simt > tc ? simt : tc;
}

extern "C" size_t emd_dense_bwd_workspace_bytes(int64_t M, int K, int Nout) {
    const DenseSplit s = dense_split(M, K, Nout);
    return (dense_wgrad_partial_floats(M, K, Nout) + (size_t)s.col_chunks * (size_t)Nout) * sizeof(float);
}

// VJP of emd_dense_fwd given dZ[M, 0:Nout] (row stride lddz) = dL/d(pre-activation) (the caller's upstream kernel
// already applied this layer's ReLU mask -- see `mask` below):
//   dX[M, 0:ncols] (row stride lddx) = (dZ . W[:, col0:col0+ncols]) * (mask > 0)     skipped when dX is NULL;
//                                      mask[M, 0:ncols] (row stride ldmask) = the activations the PRODUCER of X wrote
//                                      (its post-ReLU output), or NULL when X is not a ReLU output
//   dW[Nout,K] = dZ^T . X ,  db[Nout] = column sums of dZ                             each skipped when NULL
extern "C" int emd_dense_bwd(const float* X, int64_t ldx, const float* W, const float* dZ, int64_t lddz, int64_t M, int K,
                             int Nout, float* dX, int64_t lddx, int col0, int ncols, const float* mask, int64_t ldmask,
                             float* dW, float* db, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    EMD_CHECK_ARG(M >= 0 && K >= 1 && Nout >= 1 && ldx >= K && lddz >= Nout, "emd_dense_bwd: M=%lld K=%d Nout=%d ldx=%lld lddz=%lld",
                  (long long)M, K, Nout, (long long)ldx, (long long)lddz);
    EMD_CHECK_ARG(M == 0 || (X && W && dZ), "emd_dense_bwd: null argument");
    if (dX) {
        EMD_CHECK_ARG(col0 >= 0 && ncols >= 1 && col0 + ncols <= K && lddx >= ncols && (!mask || ldmask >= ncols),
                      "emd_dense_bwd: column window [%d, %d) of K=%d, lddx=%lld ldmask=%lld", col0, col0 + ncols, K,
                      (long long)lddx, (long long)ldmask);
        if (M > 0) {
            int rc_tc = EMD_OK;
            if (emd_dense_tc_try(1, dZ, lddz, W + col0, K, nullptr, mask, ldmask, dX, lddx, M, Nout, ncols, 0, stream, &rc_tc)) {
                if (rc_tc != EMD_OK) return rc_tc;
            } else {
                const GemmArgs g = dense_dgrad_args(W, dZ, lddz, M, K, Nout, dX, lddx, col0, ncols, mask, ldmask);
                EMD_LAUNCH(EK_DENSE_BWD, stream, (launch_sgemm<false, false>(g, 1, stream)));
                EMD_CHECK_LAUNCH("emd_dense_bwd(dgrad)");
            }
        }
    }
    if (!dW && !db) return EMD_OK;
    const DenseSplit s = dense_split(M, K, Nout);
    if (!workspace || workspace_bytes < emd_dense_bwd_workspace_bytes(M, K, Nout)) {
        emd_set_error("emd_dense_bwd: workspace too small (%zu < %zu)", workspace_bytes, emd_dense_bwd_workspace_bytes(M, K, Nout));
        return EMD_ERR_WORKSPACE;
    }
    float* wpart = static_cast<float*>(workspace);
    float* bpart = wpart + dense_wgrad_partial_floats(M, K, Nout);
    if (dW) {
        if (M == 0) {
            cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)Nout * K, stream);
        } else {
            int splits = s.splits, rc_tc = EMD_OK;
            if (emd_dense_tc_try_wgrad(X, ldx, dZ, lddz, M, K, Nout, wpart, &splits, stream, &rc_tc)) {
                if (rc_tc != EMD_OK) return rc_tc;
            } else {
                const GemmArgs g = dense_wgrad_args(X, ldx, dZ, lddz, M, K, Nout, s, wpart);
                EMD_LAUNCH(EK_DENSE_BWD, stream, (launch_sgemm<true, false>(g, s.splits, stream)));
                EMD_CHECK_LAUNCH("emd_dense_bwd(wgrad)");
            }
            const int64_t n = (int64_t)Nout * K;
            EMD_LAUNCH(EK_DENSE_BWD, stream,
                       (split_reduce_kernel<<<(unsigned)emd_cdiv(n, 256), 256, 0, stream>>>(wpart, n, splits, n, dW)));
            EMD_CHECK_LAUNCH("emd_dense_bwd(wgrad reduce)");
        }
    }
    if (db) {
        if (M == 0) {
            cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Nout, stream);
        } else {
            EMD_LAUNCH(EK_DENSE_BWD, stream,
                       (colsum_kernel<<<(unsigned)s.col_chunks, 256, 0, stream>>>(dZ, lddz, M, Nout, s.rows_per_chunk, bpart)));
            EMD_CHECK_LAUNCH("emd_dense_bwd(bias partials)");
            EMD_LAUNCH(EK_DENSE_BWD, stream,
                       (split_reduce_kernel<<<(unsigned)emd_cdiv(Nout, 256), 256, 0, stream>>>(bpart, Nout, s.col_chunks, Nout, db)));
            EMD_CHECK_LAUNCH("emd_dense_bwd(bias reduce)");
        }
    }
    return EMD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// network input: positional encodings + gathered instance embedding
// ---------------------------------------------------------------------------------------------------------------
namespace {

// columns: [x(3) | for f < xm: sin(2^f x)(3), cos(2^f x)(3)] [t | for f < tm: sin(2^f t), cos(2^f t)] [embedding(E)]
// (Embedder.create_embedding_fn, OmniRe/models/modules.py:341-366: include_input, log-sampled bands 2^0 .. 2^(m-1))
__global__ void __launch_bounds__(256) deform_input_kernel(const float* __restrict__ means, const int64_t* __restrict__ ids,
                                                           const float* __restrict__ inst_size, const float* __restrict__ inst_emb,
                                                           float t, int xm, int tm, int E, int64_t N, float* __restrict__ out0,
                                                           int64_t ld0, float* __restrict__ out1, int64_t ld1) {
    const int xcols = 3 + 6 * xm, tcols = 1 + 2 * tm, cols = xcols + tcols + E;
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= N * cols) return;
    const int64_t n = e / cols;
    const int c = (int)(e - n * cols);
    const int64_t id = __ldg(ids + n);
    float v;
    if (c < xcols) {
        const int d = c < 3 ? c : (c - 3) % 3;
        // deformable.py:42-43: x = local_means.data / ins_height[:, None] * 2
        const float x = __fmul_rn(__fdiv_rn(__ldg(means + n * 3 + d), __ldg(inst_size + id * 3 + 2)), 2.0f);
        if (c < 3) {
            v = x;
        } else {
            const int j = c - 3, f = j / 6, fn = (j % 6) / 3;
            const float a = __fmul_rn(x, (float)(1u << f));
            v = fn ? cosf(a) : sinf(a);
        }
    } else if (c < xcols + tcols) {
        const int j = c - xcols;
        if (j == 0) {
            v = t;
        } else {
            const int f = (j - 1) / 2, fn = (j - 1) % 2;
            const float a = __fmul_rn(t, (float)(1u << f));
            v = fn ? cosf(a) : sinf(a);
        }
    } else {
        v = __ldg(inst_emb + id * E + (c - xcols - tcols));
    }
    out0[n * ld0 + c] = v;
    if (out1) out1[n * ld1 + c] = v;
}

// means_out = means + d[:, 0:3];  quats_out = quats / |quats| + d[:, 3:7]  (when dcols >= 7)
__global__ void __launch_bounds__(256) deform_apply_fwd_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                                                               const float* __restrict__ d, int dcols, int64_t N,
                                                               float* __restrict__ means_out, float* __restrict__ quats_out) {
    const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const float* dn = d + n * dcols;
#pragma unroll
    for (int a = 0; a < 3; ++a) means_out[n * 3 + a] = means[n * 3 + a] + dn[a];
    if (quats_out) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
        const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float4 o = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
        if (dcols >= 7) { o.x += dn[3]; o.y += dn[4]; o.z += dn[5]; o.w += dn[6]; }
        reinterpret_cast<float4*>(quats_out)[n] = o;
    }
}

// v_d[:, 0:3] = v_means_out (also v_means when requested);  v_d[:, 3:7] = v_quats_out;
// v_quats = (v_quats_out - qn <qn, v_quats_out>) / |quats|
__global__ void __launch_bounds__(256) deform_apply_bwd_kernel(const float* __restrict__ quats, const float* __restrict__ v_mo,
                                                               const float* __restrict__ v_qo, int dcols, int64_t N,
                                                               float* __restrict__ v_d, float* __restrict__ v_means,
                                                               float* __restrict__ v_quats) {
    const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    float* dn = v_d + n * dcols;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float g = v_mo[n * 3 + a];
        dn[a] = g;
        if (v_means) v_means[n * 3 + a] = g;
    }
    const float4 go = v_qo ? __ldg(reinterpret_cast<const float4*>(v_qo) + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (dcols >= 7) { dn[3] = go.x; dn[4] = go.y; dn[5] = go.z; dn[6] = go.w; }
    for (int c = 7; c < dcols; ++c) dn[c] = 0.f;
    if (v_quats) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
        const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float4 u = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
        const float dot = u.x * go.x + u.y * go.y + u.z * go.z + u.w * go.w;
        reinterpret_cast<float4*>(v_quats)[n] =
            make_float4((go.x - u.x * dot) / nrm, (go.y - u.y * dot) / nrm, (go.z - u.z * dot) / nrm, (go.w - u.w * dot) / nrm);
    }
}

// partial[i][c][e] = sum over the points of chunk c of instance i (emission order of the instance-sorted index) of
// g0[n, e] + g1[n, e].  grid (chunks, I): enough CTAs to hide the order[] -> g[] dependent-load latency.
constexpr int EMBG_CHUNK = 256;      // points per CTA
__global__ void __launch_bounds__(256) deform_embed_grad_kernel(const float* __restrict__ g0, const float* __restrict__ g1, int E,
                                                                const int64_t* __restrict__ order,
                                                                const int64_t* __restrict__ seg_start, float* __restrict__ partial) {
    __shared__ float s[256];
    const int i = blockIdx.y, chunk = blockIdx.x;
    const int rows = 256 / E;                       // E <= 256
    const int r = threadIdx.x / E, c = threadIdx.x - r * E;
    const int64_t lo = seg_start[i] + (int64_t)chunk * EMBG_CHUNK;
    const int64_t hi = min(seg_start[i + 1], lo + EMBG_CHUNK);
    float acc = 0.f;
    if (r < rows) {
        for (int64_t p = lo + r; p < hi; p += rows) {
            const int64_t n = order[p];
            acc += g0[n * E + c] + (g1 ? g1[n * E + c] : 0.f);
        }
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < E) {
        float tot = 0.f;
        for (int k = 0; k < rows; ++k) tot += s[k * E + threadIdx.x];
        partial[((int64_t)i * gridDim.x + chunk) * E + threadIdx.x] = tot;
    }
}

// v_emb[i][e] = sum over the chunks, fixed order
__global__ void __launch_bounds__(256) deform_embed_grad_finish_kernel(const float* __restrict__ partial, int chunks, int E, int I,
                                                                       float* __restrict__ v_emb) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= I * E) return;
    const int i = t / E, e = t - i * E;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += partial[((int64_t)i * chunks + c) * E + e];
    v_emb[t] = acc;
}

}  // namespace

// Network input of ConditionalDeformNetwork as DeformableNodes.get_deformation builds it (deformable.py:40-46 +
// modules.py:436-438): out0[N, cols] (row stride ld0) and, when out1 != NULL, the same columns into out1 (row stride
// ld1) -- the head of the skip layer's operand.  cols = 3 + 6 x_multires + 1 + 2 t_multires + E.
// means[N,3]; point_ids[N] int64; instances_size[I,3] (height = column 2); instances_embedding[I,E]; t = normalised time.
extern "C" int emd_deform_input_fwd(const float* means, const int64_t* point_ids, const float* instances_size,
                                    const float* instances_embedding, float t, int x_multires, int t_multires, int E, int64_t N,
                                    float* out0, int64_t ld0, float* out1, int64_t ld1, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && x_multires >= 0 && x_multires <= 24 && t_multires >= 0 && t_multires <= 24 && E >= 0,
                  "emd_deform_input_fwd: N=%lld x_multires=%d t_multires=%d E=%d", (long long)N, x_multires, t_multires, E);
    const int cols = 3 + 6 * x_multires + 1 + 2 * t_multires + E;
    EMD_CHECK_ARG(ld0 >= cols && (!out1 || ld1 >= cols), "emd_deform_input_fwd: row strides %lld / %lld < %d columns",
                  (long long)ld0, (long long)ld1, cols);
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(means && point_ids && instances_size && (E == 0 || instances_embedding) && out0, "emd_deform_input_fwd: null argument");
    const int64_t total = N * cols;
    EMD_LAUNCH(EK_DEFORM_IN, stream,
               (deform_input_kernel<<<(unsigned)emd_cdiv(total, 256), 256, 0, stream>>>(
                   means, point_ids, instances_size, instances_embedding, t, x_multires, t_multires, E, N, out0, ld0, out1, ld1)));
    EMD_CHECK_LAUNCH("emd_deform_input_fwd");
    return EMD_OK;
}

extern "C" size_t emd_deform_embed_grad_workspace_bytes(int I, int E, int64_t max_points_per_instance) {
    const int64_t chunks = emd_cdiv(max_points_per_instance > 0 ? max_points_per_instance : 1, EMBG_CHUNK);
    return (size_t)(I > 0 ? I : 1) * (size_t)chunks * (size_t)(E > 0 ? E : 1) * sizeof(float);
}

// Gradient of the instance embedding: v_emb[I,E] = per-instance sum of g0[N,E] (+ g1[N,E] when not NULL), the embedding
// columns of the layer-0 and skip-layer input gradients.  order / seg_start: points stably sorted by instance + the I+1
// boundaries (the index emd_rigid_deform_fwd uses); max_points_per_instance: any upper bound on the largest segment.
// Fixed-order two-level reduction (per 256-point chunk, then over the chunks).
extern "C" int emd_deform_embed_grad(const float* g0, const float* g1, int E, const int64_t* order, const int64_t* seg_start,
                                     int I, int64_t max_points_per_instance, void* workspace, size_t workspace_bytes,
                                     float* v_emb, cudaStream_t stream) {
    EMD_CHECK_ARG(E >= 1 && E <= 256 && I >= 0 && max_points_per_instance >= 0, "emd_deform_embed_grad: E=%d I=%d max=%lld", E, I,
                  (long long)max_points_per_instance);
    if (I == 0) return EMD_OK;
    EMD_CHECK_ARG(g0 && order && seg_start && v_emb, "emd_deform_embed_grad: null argument");
    const int64_t chunks = emd_cdiv(max_points_per_instance > 0 ? max_points_per_instance : 1, EMBG_CHUNK);
    EMD_CHECK_ARG(chunks <= 0x7fffffff && I <= 65535, "emd_deform_embed_grad: grid too large (chunks=%lld, I=%d)", (long long)chunks, I);
    if (!workspace || workspace_bytes < emd_deform_embed_grad_workspace_bytes(I, E, max_points_per_instance)) {
        emd_set_error("emd_deform_embed_grad: workspace too small (%zu < %zu)", workspace_bytes,
                      emd_deform_embed_grad_workspace_bytes(I, E, max_points_per_instance));
        return EMD_ERR_WORKSPACE;
    }
    float* partial = static_cast<float*>(workspace);
    EMD_LAUNCH(EK_DEFORM_IN, stream,
               (deform_embed_grad_kernel<<<dim3((unsigned)chunks, (unsigned)I), 256, 0, stream>>>(g0, g1, E, order, seg_start, partial)));
    EMD_CHECK_LAUNCH("emd_deform_embed_grad");
    EMD_LAUNCH(EK_DEFORM_IN, stream,
               (deform_embed_grad_finish_kernel<<<(unsigned)emd_cdiv((int64_t)I * E, 256), 256, 0, stream>>>(partial, (int)chunks, E, I, v_emb)));
    EMD_CHECK_LAUNCH("emd_deform_embed_grad(finish)");
    return EMD_OK;
}

// means_out[N,3] = means + d[:, 0:3];  quats_out[N,4] = normalize(quats) + d[:, 3:7]   (deformable.py:57-68;
// get_quats = quat_act(_quats), vanilla.py:142-146).  d[N,dcols], dcols = 3 (no quaternion head) or >= 7.
// quats / quats_out may be NULL together (dcols = 3).
extern "C" int emd_deform_apply_fwd(const float* means, const float* quats, const float* d, int dcols, int64_t N,
                                    float* means_out, float* quats_out, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && (dcols == 3 || dcols >= 7), "emd_deform_apply_fwd: N=%lld dcols=%d", (long long)N, dcols);
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(means && d && means_out && ((quats == nullptr) == (quats_out == nullptr)), "emd_deform_apply_fwd: null argument");
    if (quats && (!emd_aligned(quats, 16) || !emd_aligned(quats_out, 16))) {
        emd_set_error("emd_deform_apply_fwd: quats must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    EMD_LAUNCH(EK_DEFORM_IN, stream,
               (deform_apply_fwd_kernel<<<(unsigned)emd_cdiv(N, 256), 256, 0, stream>>>(means, quats, d, dcols, N, means_out, quats_out)));
    EMD_CHECK_LAUNCH("emd_deform_apply_fwd");
    return EMD_OK;
}

// VJP of emd_deform_apply_fwd: v_d[N,dcols] written (columns >= 7 zero); v_means[N,3] (NULL when the canonical means are
// detached: stop_optimizing_canonical_xyz) and v_quats[N,4] (NULL to skip).  v_quats_out may be NULL (treated as zero).
extern "C" int emd_deform_apply_bwd(const float* quats, const float* v_means_out, const float* v_quats_out, int dcols, int64_t N,
                                    float* v_d, float* v_means, float* v_quats, cudaStream_t stream) {
    EMD_CHECK_ARG(N >= 0 && (dcols == 3 || dcols >= 7), "emd_deform_apply_bwd: N=%lld dcols=%d", (long long)N, dcols);
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(v_means_out && v_d && (!v_quats || quats), "emd_deform_apply_bwd: null argument");
    if ((v_quats_out && !emd_aligned(v_quats_out, 16)) || (v_quats && (!emd_aligned(v_quats, 16) || !emd_aligned(quats, 16)))) {
        emd_set_error("emd_deform_apply_bwd: quaternion buffers must be 16-byte aligned");
        return EMD_ERR_ALIGN;
    }
    EMD_LAUNCH(EK_DEFORM_IN, stream,
               (deform_apply_bwd_kernel<<<(unsigned)emd_cdiv(N, 256), 256, 0, stream>>>(quats, v_means_out, v_quats_out, dcols, N, v_d,
                                                                                         v_means, v_quats)));
    EMD_CHECK_LAUNCH("emd_deform_apply_bwd");
    return EMD_OK;
}
