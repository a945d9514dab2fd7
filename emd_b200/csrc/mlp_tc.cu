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
// CTA = 128 threads (4 warps) = one 128-row tile at a time (persistent over tiles):
//   stage   : all threads read the tile's contiguous 128 x K block coalesced (float4), apply act_in, split hi/lo
//             and write both copies in the K-major no-swizzle UMMA canonical layout
//             (core matrix = 8 rows x 16 B contiguous; row groups 128 B apart; K-adjacent core matrices LBO apart)
//   mma     : one elected thread issues K/8 x 3 tcgen05.mma.kind::tf32 (M=128, N=Npad, K=8) and commits to an mbarrier
//   epilogue: warp w reads TMEM lanes 32w..32w+31 (tcgen05.ld 32x32b), adds the bias, applies act_out, stores Y
// W (hi and lo) is staged once per CTA.  No cuBLAS, no CUTLASS: descriptors and PTX are written out below.
#include "common.cuh"

namespace {

constexpr int TC_ROWS = 128;                 // rows per tile == UMMA M == threads per CTA
constexpr int TC_KMAX = 136;                 // largest padded K (132 -> 136)
constexpr int TC_NMAX = 64;                  // largest padded Nout
constexpr int TC_A_LBO = TC_ROWS * 16 + 16;  // bytes between K-adjacent core-matrix columns of A (+16: bank spread)
constexpr int TC_SBO = 128;                  // bytes between consecutive 8-row groups (core matrices are contiguous)
constexpr int TC_TMEM_COLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: bits [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// kind::tf32, fp32 accumulate, A and B K-major: c_format [4,6) = 1, a_format [7,10) = 2, b_format [10,13) = 2,
// n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// dynamic shared memory: [A_hi | A_lo | B_hi | B_lo], every region a multiple of 16 B
struct TcLayout {
    int kchunks;        // Kpad / 4
    int npad;           // Nout padded to a multiple of 16
    int b_lbo;          // bytes between K-adjacent core-matrix columns of B
    int a_bytes, b_bytes;
    __host__ __device__ TcLayout(int K, int Nout) {
        const int kpad = (K + 7) / 8 * 8;
        kchunks = kpad / 4;
        npad = (Nout + 15) / 16 * 16;
        b_lbo = npad * 16 + 16;
        a_bytes = kchunks * TC_A_LBO;
        b_bytes = kchunks * b_lbo;
    }
    __host__ __device__ size_t total() const { return 2 * (size_t)a_bytes + 2 * (size_t)b_bytes; }
};

template <bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(TC_ROWS, 1) linear_fwd_tc_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, int64_t M, int K,
                                                                   int Nout, float* __restrict__ Y) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[TC_NMAX];
    const TcLayout L(K, Nout);
    unsigned char* sAhi = tc_smem;
    unsigned char* sAlo = sAhi + L.a_bytes;
    unsigned char* sBhi = sAlo + L.a_bytes;
    unsigned char* sBlo = sBhi + L.b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kq = K / 4;   // float4 per row (K % 4 == 0)

    // ---- one-time setup: zero the padded operands, stage W (hi / lo), bias, barrier, TMEM ----
    for (int e = tid; e < (int)(L.total() / 16); e += TC_ROWS) reinterpret_cast<float4*>(tc_smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < TC_NMAX) s_bias[tid] = tid < Nout ? __ldg(bias + tid) : 0.f;
    __syncthreads();
    for (int e = tid; e < Nout * kq; e += TC_ROWS) {
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
    const int ksteps = L.kchunks / 2;
    uint32_t phase = 0;

    const int64_t n_tiles = (M + TC_ROWS - 1) / TC_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TC_ROWS;
        const int rows = (int)min((int64_t)TC_ROWS, M - row0);
        // ---- stage A: the tile is one contiguous span of rows * K floats ----
        const float4* x4 = reinterpret_cast<const float4*>(X + row0 * K);
        for (int q = tid; q < TC_ROWS * kq; q += TC_ROWS) {
            const int r = q / kq, j = q - r * kq;
            float4 x = r < rows ? __ldg(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (RELU_IN) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            float4 hi, lo;
            split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(sAhi + j * TC_A_LBO + r * 16) = hi;
            *reinterpret_cast<float4*>(sAlo + j * TC_A_LBO + r * 16) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        // ---- MMA: one thread issues, completion arrives on the mbarrier ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int s = 0; s < ksteps; ++s) {
                const uint64_t dAh = umma_desc(aHi + 2 * s * TC_A_LBO, TC_A_LBO, TC_SBO);
                const uint64_t dAl = umma_desc(aLo + 2 * s * TC_A_LBO, TC_A_LBO, TC_SBO);
                const uint64_t dBh = umma_desc(bHi + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                const uint64_t dBl = umma_desc(bLo + 2 * s * L.b_lbo, L.b_lbo, TC_SBO);
                umma_tf32(tmem, dAl, dBh, idesc, s > 0 ? 1u : 0u);   // small terms first
                umma_tf32(tmem, dAh, dBl, idesc, 1u);
                umma_tf32(tmem, dAh, dBh, idesc, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: thread = row (TMEM lane), 16 columns at a time ----
        const int64_t row = row0 + tid;
        for (int c0 = 0; c0 < L.npad; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (row < M) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float y = v[c] + s_bias[c0 + c];
                    if (RELU_OUT) y = fmaxf(y, 0.f);
                    v[c] = y;
                }
                float* yr = Y + row * Nout + c0;
                if ((Nout & 3) == 0 && c0 + 16 <= Nout) {
#pragma unroll
                    for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(yr + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c0 + c < Nout) yr[c] = v[c];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM and the A tile are free again
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

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
    const unsigned grid = (unsigned)(n_tiles < EMD_NUM_SMS ? n_tiles : EMD_NUM_SMS);
#define EMD_TC_LAUNCH(RI, RO)                                                                                          \
    do {                                                                                                               \
        cudaFuncSetAttribute(linear_fwd_tc_kernel<RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        EMD_LAUNCH(EK_MLP_FWD, stream, (linear_fwd_tc_kernel<RI, RO><<<grid, TC_ROWS, smem, stream>>>(X, W, b, M, K, Nout, Y))); \
    } while (0)
    if (relu_in && relu_out) EMD_TC_LAUNCH(true, true);
    else if (relu_in) EMD_TC_LAUNCH(true, false);
    else if (relu_out) EMD_TC_LAUNCH(false, true);
    else EMD_TC_LAUNCH(false, false);
#undef EMD_TC_LAUNCH
    EMD_CHECK_LAUNCH("linear_fwd_tc");
    return EMD_OK;
}
