// Tile logic of the strided fp32 GEMM behind emd_dense_fwd / emd_dense_bwd (K1g), host/device.
//
// Everything a thread of sgemm_kernel (deform_net.cu) does -- which elements it loads, where it puts them in shared
// memory, its 8x8 micro-tile, the epilogue -- and how the three GEMMs of a layer (forward, data gradient, weight
// gradient) map onto the generic kernel, lives here so that hostmath.cpp can run the SAME code thread by thread on the
// CPU (`pytest -m "not gpu"` checks it against torch).  The kernel only adds the __syncthreads() between the phases.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef EMD_HD
#ifdef __CUDACC__
#define EMD_HD __host__ __device__ __forceinline__
#else
#define EMD_HD static inline
#endif
#endif

#ifdef __CUDA_ARCH__
#define DG_LDG(p) __ldg(p)
#else
#define DG_LDG(p) (*(p))
#endif

constexpr int DG_BM = 128, DG_BN = 128, DG_BK = 16, DG_THREADS = 256, DG_PITCH = DG_BM + 4;
constexpr int DG_NUM_SMS = 148;

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int64_t lda, ldb, ldc;
    int64_t M;            // rows of C
    int64_t N;            // columns of C
    int64_t K;            // reduction length
    int64_t k_per_split;  // reduction range per blockIdx.z (multiple of DG_BK)
    int64_t split_stride; // floats between the partial results of consecutive splits (0: single split)
    const float* bias;    // [N] or NULL
    int relu;             // ReLU after the bias
    const float* mask;    // or NULL: result *= (mask[m * ldmask + n] > 0)
    int64_t ldmask;
    int vec_store;        // C rows are 16-byte aligned and ldc % 4 == 0
};

// C[m,n] = sum_k A(m,k) B(k,n);   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k];   B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
// Each thread fetches 8 elements of the 128x16 A tile and 8 of the 16x128 B tile, consecutive threads along the
// operand's contiguous dimension.  Element i of a thread sits a FIXED stride after element 0 (16 rows when the
// operand is k-contiguous, 2 k-steps when it is row-contiguous), so a thread keeps one base offset per operand and the
// tile loop only adds k0 * (k stride): no per-element index arithmetic, no branches around the guarded loads.
struct GemmLoadState {
    const float* pa;        // this thread's element 0 of the CURRENT k-tile (advanced by gemm_advance)
    const float* pb;
    int64_t stepA, stepB;   // offset between the thread's consecutive elements
    int64_t advA, advB;     // offset between consecutive k-tiles
    int kA, kB;             // tile-local k of element 0
    unsigned maskA, maskB;  // bit i: the ROW (m resp. n) of element i is inside the matrix
};

template <bool TA, bool TB>
EMD_HD void gemm_prepare(const GemmArgs& g, int tid, int64_t m0, int64_t n0, int64_t kbeg, GemmLoadState& S) {
    const int rowA = TA ? (tid & (DG_BM - 1)) : (tid >> 4);
    const int rowB = TB ? (tid >> 4) : (tid & (DG_BN - 1));
    S.kA = TA ? (tid >> 7) : (tid & (DG_BK - 1));
    S.kB = TB ? (tid & (DG_BK - 1)) : (tid >> 7);
    S.pa = g.A + (TA ? (kbeg + S.kA) * g.lda + (m0 + rowA) : (m0 + rowA) * g.lda + (kbeg + S.kA));
    S.pb = g.B + (TB ? (n0 + rowB) * g.ldb + (kbeg + S.kB) : (kbeg + S.kB) * g.ldb + (n0 + rowB));
    S.stepA = TA ? 2 * g.lda : 16 * g.lda;      // TA: element i is k + 2 i of the same m; else: row m + 16 i of the same k
    S.stepB = TB ? 16 * g.ldb : 2 * g.ldb;
    S.advA = TA ? DG_BK * g.lda : DG_BK;
    S.advB = TB ? DG_BK : DG_BK * g.ldb;
    S.maskA = 0;
    S.maskB = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (m0 + rowA + (TA ? 0 : 16 * i) < g.M) S.maskA |= 1u << i;
        if (n0 + rowB + (TB ? 16 * i : 0) < g.N) S.maskB |= 1u << i;
    }
}

// loads the thread's 8 + 8 elements of the k-tile that starts at k0 (S.pa / S.pb point at it).  FULLK: the whole
// tile lies below kend, only the row masks guard the loads; otherwise every element's k is checked as well.
template <bool TA, bool TB, bool FULLK>
EMD_HD void gemm_load(const GemmLoadState& S, int64_t k0, int64_t kend, float (&ra)[8], float (&rb)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bool oka = (S.maskA >> i) & 1u, okb = (S.maskB >> i) & 1u;
        if (!FULLK) {
            oka = oka && (k0 + S.kA + (TA ? 2 * i : 0) < kend);
            okb = okb && (k0 + S.kB + (TB ? 0 : 2 * i) < kend);
        }
        const float* qa = S.pa + i * S.stepA;
        const float* qb = S.pb + i * S.stepB;
        float va = 0.f, vb = 0.f;
        if (oka) va = DG_LDG(qa);
        if (okb) vb = DG_LDG(qb);
        ra[i] = va;
        rb[i] = vb;
    }
}

EMD_HD void gemm_advance(GemmLoadState& S) {
    S.pa += S.advA;
    S.pb += S.advB;
}

// one k-tile's loads, choosing the unguarded-k flavour when the tile is full (uniform across the CTA)
template <bool TA, bool TB>
EMD_HD void gemm_load_tile(const GemmLoadState& S, int64_t k0, int64_t kend, float (&ra)[8], float (&rb)[8]) {
    if (k0 + DG_BK <= kend) gemm_load<TA, TB, true>(S, k0, kend, ra, rb);
    else gemm_load<TA, TB, false>(S, k0, kend, ra, rb);
}

// shared tiles are k-major: As[k][m], Bs[k][n]
template <bool TA, bool TB>
EMD_HD void gemm_store(int tid, const float (&ra)[8], const float (&rb)[8], float (*As)[DG_PITCH], float (*Bs)[DG_PITCH]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = tid + i * DG_THREADS;
        As[TA ? (e >> 7) : (e & (DG_BK - 1))][TA ? (e & (DG_BM - 1)) : (e >> 4)] = ra[i];
        Bs[TB ? (e & (DG_BK - 1)) : (e >> 7)][TB ? (e >> 4) : (e & (DG_BN - 1))] = rb[i];
    }
}

// thread (ty, tx) of the 16x16 grid owns rows {4 ty + i, 64 + 4 ty + i} x columns {4 tx + j, 64 + 4 tx + j}, i, j < 4:
// every shared-memory read is one aligned 16-byte vector, the 16 tx-lanes read consecutive vectors (no bank conflict)
EMD_HD void gemm_compute(int tid, const float (*As)[DG_PITCH], const float (*Bs)[DG_PITCH], float (&acc)[8][8]) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int k = 0; k < DG_BK; ++k) {
        float a[8], b[8];
#ifdef __CUDA_ARCH__
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#else
        for (int i = 0; i < 4; ++i) {
            a[i] = As[k][ty * 4 + i]; a[4 + i] = As[k][64 + ty * 4 + i];
            b[i] = Bs[k][tx * 4 + i]; b[4 + i] = Bs[k][64 + tx * 4 + i];
        }
#endif
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// bias, ReLU, the producer's ReLU mask; 16-byte stores when the destination allows
EMD_HD void gemm_epilogue(const GemmArgs& g, int tid, int64_t m0, int64_t n0, float* C, const float (&acc)[8][8]) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int64_t n = n0 + jh * 64 + tx * 4;
            if (n >= g.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][jh * 4 + j];
                if (n + j < g.N) {
                    if (g.bias) v[j] += DG_LDG(g.bias + n + j);
                    if (g.relu) v[j] = v[j] > 0.f ? v[j] : 0.f;
                    if (g.mask && !(DG_LDG(g.mask + m * g.ldmask + n + j) > 0.f)) v[j] = 0.f;
                }
            }
            float* dst = C + m * g.ldc + n;
#ifdef __CUDA_ARCH__
            if (g.vec_store && n + 3 < g.N) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                continue;
            }
#endif
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < g.N) dst[j] = v[j];
        }
    }
}

// ---- how a layer's three GEMMs map onto the kernel (host side) --------------------------------------------------
static inline int64_t dg_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool dg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct DenseSplit {
    int splits;              // weight-gradient partials (blockIdx.z)
    int64_t k_per_split;     // rows per partial, multiple of DG_BK
    int col_chunks;          // bias-gradient partials
    int64_t rows_per_chunk;
};

// how the weight / bias gradients of a layer with M rows are split (shared by the workspace query and the launch)
static inline DenseSplit dense_split(int64_t M, int K, int Nout) {
    DenseSplit s;
    const int64_t rows = M > 0 ? M : 1;
    const int64_t tiles = dg_cdiv(Nout, DG_BM) * dg_cdiv(K, DG_BN);
    int64_t want = dg_cdiv(2 * DG_NUM_SMS, tiles);
    const int64_t kt = dg_cdiv(rows, DG_BK);             // k-tiles available
    if (want > kt) want = kt;
    if (want < 1) want = 1;
    s.k_per_split = dg_cdiv(kt, want) * DG_BK;
    s.splits = (int)dg_cdiv(rows, s.k_per_split);
    int64_t chunks = dg_cdiv(rows, 128);
    if (chunks > 4 * DG_NUM_SMS) chunks = 4 * DG_NUM_SMS;
    s.rows_per_chunk = dg_cdiv(rows, chunks);
    s.col_chunks = (int)dg_cdiv(rows, s.rows_per_chunk);
    return s;
}

// forward: C = Y[M,Nout], A = X (row-major), B(k,n) = W[n*K + k]                       -> sgemm_kernel<false, true>
static inline GemmArgs dense_fwd_args(const float* X, int64_t ldx, const float* W, const float* b, int64_t M, int K, int Nout,
                                      int relu_out, float* Y, int64_t ldy) {
    GemmArgs g = {};
    g.A = X; g.lda = ldx; g.B = W; g.ldb = K; g.C = Y; g.ldc = ldy;
    g.M = M; g.N = Nout; g.K = K; g.k_per_split = dg_cdiv(K, DG_BK) * DG_BK; g.split_stride = 0;
    g.bias = b; g.relu = relu_out; g.mask = nullptr; g.ldmask = 0;
    g.vec_store = dg_aligned16(Y) && ldy % 4 == 0;
    return g;
}

// data gradient: C = dX[M,ncols], A = dZ (row-major), B(k,n) = W[k*K + col0 + n]       -> sgemm_kernel<false, false>
static inline GemmArgs dense_dgrad_args(const float* W, const float* dZ, int64_t lddz, int64_t M, int K, int Nout, float* dX,
                                        int64_t lddx, int col0, int ncols, const float* mask, int64_t ldmask) {
    GemmArgs g = {};
    g.A = dZ; g.lda = lddz; g.B = W + col0; g.ldb = K; g.C = dX; g.ldc = lddx;
    g.M = M; g.N = ncols; g.K = Nout; g.k_per_split = dg_cdiv(Nout, DG_BK) * DG_BK; g.split_stride = 0;
    g.bias = nullptr; g.relu = 0; g.mask = mask; g.ldmask = ldmask;
    g.vec_store = dg_aligned16(dX) && lddx % 4 == 0;
    return g;
}

// weight gradient: C = partial dW[Nout,K] per split, A(m,k) = dZ[k*lddz + m], B(k,n) = X[k*ldx + n], reduction over
// the M rows split across blockIdx.z                                                    -> sgemm_kernel<true, false>
static inline GemmArgs dense_wgrad_args(const float* X, int64_t ldx, const float* dZ, int64_t lddz, int64_t M, int K, int Nout,
                                        const DenseSplit& s, float* wpart) {
    GemmArgs g = {};
    g.A = dZ; g.lda = lddz; g.B = X; g.ldb = ldx; g.C = wpart; g.ldc = K;
    g.M = Nout; g.N = K; g.K = M; g.k_per_split = s.k_per_split; g.split_stride = (int64_t)Nout * K;
    g.bias = nullptr; g.relu = 0; g.mask = nullptr; g.ldmask = 0;
    g.vec_store = K % 4 == 0 && dg_aligned16(wpart);
    return g;
}
