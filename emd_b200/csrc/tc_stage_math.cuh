// Who-writes-what of the tensor-core staging in deform_net_tc.cu, host/device: the thread -> element maps and the
// shared-memory byte offsets of the K-major no-swizzle UMMA canonical layout (core matrix = 8 rows x 16 B, rows of a
// core-matrix column contiguous, K-adjacent columns LBO bytes apart -- the layout csrc/mlp_tc.cu runs on B200).
// hostmath.cpp replays every thread's writes on the CPU and reads the operand back the way the descriptor walks it
// (tests/test_cpu_hostmath.py), so an indexing slip shows up without a GPU.
#pragma once
#include <stdint.h>

#ifndef EMD_HD
#ifdef __CUDACC__
#define EMD_HD __host__ __device__ __forceinline__
#else
#define EMD_HD static inline
#endif
#endif

constexpr int DTS_ROWS = 128;                               // UMMA M
constexpr int DTS_THREADS = 256;
constexpr int DTS_KC = 32;                                  // reduction columns per chunk
constexpr int DTS_NMAX = 256;
constexpr int DTS_A_LBO = DTS_ROWS * 16 + 16;               // == TC_A_LBO of tc_common.cuh
constexpr int DTS_B_LBO = DTS_NMAX * 16 + 16;
constexpr int DTS_SBO = 128;                                // 8 rows x 16 B

// byte offset of element (row, k) of a staged operand whose K-adjacent core-matrix columns are `lbo` bytes apart --
// what a descriptor (start, LBO = lbo, SBO = 128) makes the tensor core read for k-step k / 8
EMD_HD int dts_canonical_offset(int row, int k, int lbo) { return (k >> 2) * lbo + (row >> 3) * DTS_SBO + (row & 7) * 16 + (k & 3) * 4; }

// A chunk: thread tid, i < 4 -> row cr + 32 i, core-matrix column cj (4 floats); float4 written at the returned offset
EMD_HD void dts_a_elem(int tid, int i, int& row, int& cj) {
    cj = tid & 7;
    row = (tid >> 3) + 32 * i;
}
EMD_HD int dts_a_store_offset(int row, int cj) { return cj * DTS_A_LBO + row * 16; }

// B chunk, forward (W[n][k], k contiguous): thread tid, i < 8 -> output column n, core-matrix column j; float4 store
EMD_HD void dts_b_elem_fwd(int tid, int i, int& n, int& j) {
    const int e = tid + DTS_THREADS * i;
    n = e >> 3;
    j = e & 7;
}
EMD_HD int dts_b_store_offset_fwd(int n, int j) { return j * DTS_B_LBO + n * 16; }

// B chunk, data gradient (W[k][n], n contiguous): thread tid, i < 8 -> reduction row k of the chunk, 4 consecutive output
// columns starting at n; four scalar stores
EMD_HD void dts_b_elem_dgrad(int tid, int i, int& k, int& n) {
    const int e = tid + DTS_THREADS * i;
    k = e >> 6;
    n = (e & 63) * 4;
}
EMD_HD int dts_b_store_offset_dgrad(int k, int n) { return (k >> 2) * DTS_B_LBO + n * 16 + (k & 3) * 4; }

// A chunk of the weight-gradient GEMM (A(m, k) = dZ[k][m], m contiguous in memory): thread tid, i < 4 -> reduction row k
// of the chunk, 4 consecutive operand rows starting at m; four scalar stores
EMD_HD void dts_at_elem(int tid, int i, int& k, int& m) {
    const int e = tid + DTS_THREADS * i;
    k = e >> 5;
    m = (e & 31) * 4;
}
EMD_HD int dts_at_store_offset(int k, int m) { return (k >> 2) * DTS_A_LBO + m * 16 + (k & 3) * 4; }

// how the weight gradient's reduction over the M rows is split across CTAs (shared by the workspace query and the launch)
struct DtsWgradSplit {
    int n_blocks, col_blocks, splits;
    int64_t rows_per_split;      // multiple of DTS_KC
};
static inline DtsWgradSplit dts_wgrad_split(int64_t M, int K, int Nout, int num_sms) {
    DtsWgradSplit s;
    s.n_blocks = (Nout + DTS_ROWS - 1) / DTS_ROWS;
    s.col_blocks = (K + DTS_NMAX - 1) / DTS_NMAX;
    const int64_t rows = M > 0 ? M : 1;
    const int64_t chunks = (rows + DTS_KC - 1) / DTS_KC;
    int64_t want = num_sms / (s.n_blocks * s.col_blocks);
    if (want < 1) want = 1;
    if (want > chunks) want = chunks;
    s.rows_per_split = ((chunks + want - 1) / want) * DTS_KC;
    s.splits = (int)((rows + s.rows_per_split - 1) / s.rows_per_split);
    return s;
}

// start offset (inside a stage's operand) of the two core-matrix columns MMA k-step `sl` of the chunk reads
EMD_HD int dts_kstep_offset(int sl, int lbo) { return 2 * sl * lbo; }
