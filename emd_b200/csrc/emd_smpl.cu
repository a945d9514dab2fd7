// K1c: EMD motion-embedding deformation of SMPL nodes, forward and backward.
//
// Replaces SMPLNodes.transform_means_and_quats (OmniRe/models/nodes/smpl.py:438-532)
// + SMPLTemplate.forward's kinematic chain (human_body.py:158-172):
//   segmean  : per-instance mean motion embedding (instances own V contiguous points)
//   instance : one thread per instance -- 24 EMD joint-yaw offsets, pose quats -> rotations,
//              24-joint kinematic chain, A = chain * A0_inv                  [emd_math.cuh]
//   points   : one thread per Gaussian -- T = sum_j W[n,j] A[j] (linear-blend skinning),
//              x' = T.R x + T.t + trans, q' = normalize(mat2quat(T.R)) (x) normalize(q);
//              invisible instances get x' = trans, q' = identity (smpl.py:498-530)
// Backward reduces dL/dA (24x12 per instance) as a fixed-order [24x256]x[256x12]
// product per 256-point slab: no float atomics.
#include "common.cuh"
#include "emd_math.cuh"

namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_CHUNK = 1024;                 // points per (instance, chunk) block
constexpr int SM_AOUT = SMPL_J * 12;           // 288 floats of A per instance
constexpr int SM_RED = SM_AOUT + 3;            // + v_trans

struct SmplArgs {
    const float* table;   // [I][E][d]
    int I, E, d, g, V;
    float t;
    int cur_c, cur_f;
    SmplHeads H;
    const float* theta;    // [I][24][4]
    const float* trans;    // [I][3]
    const uint8_t* visible;  // [I]
    const float* J;        // [I][24][3]
    const float* A0inv;    // [I][24][16]
    const float* W;        // [I][V][24]
    const int* parents;    // [24]
    int max_chunks;
};

__global__ void __launch_bounds__(SM_THREADS) smpl_segmean_kernel(const float* __restrict__ emb, int g, int V,
                                                                  int max_chunks, float* __restrict__ partial) {
    __shared__ float s_scratch[SM_THREADS / 32 * EMD_GDIM_MAX];
    const int inst = blockIdx.y, chunk = blockIdx.x;
    const int64_t lo = (int64_t)inst * V + (int64_t)chunk * SM_CHUNK;
    const int64_t hi = min((int64_t)(inst + 1) * V, lo + SM_CHUNK);
    float acc[EMD_GDIM_MAX];
#pragma unroll
    for (int k = 0; k < EMD_GDIM_MAX; ++k) acc[k] = 0.f;
    for (int64_t n = lo + threadIdx.x; n < hi; n += SM_THREADS) {
#pragma unroll
        for (int k = 0; k < EMD_GDIM_MAX; ++k)
            if (k < g) acc[k] += emb[n * g + k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < EMD_GDIM_MAX; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    }
    if (lane == 0)
        for (int k = 0; k < EMD_GDIM_MAX; ++k) s_scratch[warp * EMD_GDIM_MAX + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < g) {
        float s = 0.f;
        for (int w = 0; w < SM_THREADS / 32; ++w) s += s_scratch[w * EMD_GDIM_MAX + threadIdx.x];
        partial[((int64_t)inst * max_chunks + chunk) * g + threadIdx.x] = s;
    }
}

// One CTA per instance: features, the 24 joint heads, joint rotations and the A = G * A0inv products run one
// thread per joint; only the parent-to-child chain itself is sequential (one thread, shared memory).
// Same arithmetic / summation order as smpl_instance_fwd (emd_math.cuh).
constexpr int SIF_THREADS = 64;
__global__ void __launch_bounds__(SIF_THREADS) smpl_instance_fwd_kernel(SmplArgs a, const float* __restrict__ seg_partial,
                                                                        float* __restrict__ mean_emb, float* __restrict__ A) {
    constexpr int IN_MAX = EMD_TDIM_MAX + EMD_GDIM_MAX;
    __shared__ float s_m[EMD_GDIM_MAX], s_hc[IN_MAX], s_hf[IN_MAX], s_ac[SMPL_J], s_af[SMPL_J];
    __shared__ float s_R[SMPL_J][9], s_G[SMPL_J][12];
    __shared__ int s_par[SMPL_J];
    __shared__ int s_skip;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int nch = (a.V + SM_CHUNK - 1) / SM_CHUNK;
    const int in = a.d + a.g;
    if (tid < a.g) {
        float s = 0.f;
        for (int c = 0; c < nch; ++c) s += seg_partial[((int64_t)i * a.max_chunks + c) * a.g + tid];
        s_m[tid] = s / (float)a.V;
        mean_emb[i * a.g + tid] = s_m[tid];
    }
    if (!a.visible[i]) return;  // block-uniform
    if (tid < SMPL_J) s_par[tid] = a.parents[tid];
    __syncthreads();
    const float* table = a.table + (int64_t)i * a.E * a.d;
    const float* theta = a.theta + (int64_t)i * SMPL_J * 4;
    const float* J = a.J + (int64_t)i * SMPL_J * 3;
    const float* A0inv = a.A0inv + (int64_t)i * SMPL_J * 16;
    TembTaps tc, tf;
    temb_taps(a.t, a.cur_c, a.E, tc);
    temb_taps(a.t, a.cur_f, a.E, tf);
    for (int k = tid; k < in; k += SIF_THREADS) {
        float c, f;
        if (k < a.d) {
            c = 0.f; f = 0.f;
            for (int q = 0; q < 4; ++q) { c += tc.w[q] * table[tc.row[q] * a.d + k]; f += tf.w[q] * table[tf.row[q] * a.d + k]; }
        } else {
            c = f = s_m[k - a.d];
        }
        s_hc[k] = c; s_hf[k] = f;
    }
    __syncthreads();
    if (tid < SMPL_J) {
        float sc = a.H.c_b[tid], sf = a.H.f_b[tid];
        for (int k = 0; k < in; ++k) { sc += a.H.c_w[tid * in + k] * s_hc[k]; sf += a.H.f_w[tid * in + k] * s_hf[k]; }
        s_ac[tid] = sc; s_af[tid] = sf;
    }
    __syncthreads();
    if (tid == 0) s_skip = (any_nan(s_ac, SMPL_J) || any_nan(s_af, SMPL_J)) ? 1 : 0;
    __syncthreads();
    if (tid < SMPL_J) {
        const int j = tid;
        float th[4];
        if (s_skip) { for (int k = 0; k < 4; ++k) th[k] = theta[j * 4 + k]; }
        else {
            const float qc[4] = {cosf(s_ac[j]), 0.f, 0.f, sinf(s_ac[j])}, qf[4] = {cosf(s_af[j]), 0.f, 0.f, sinf(s_af[j])};
            float qo[4];
            qmul(qc, qf, qo);
            qmul(theta + j * 4, qo, th);
        }
        float thn[4], R[9];
        qnormalize(th, thn);
        qrot(thn, R);
        for (int k = 0; k < 9; ++k) s_R[j][k] = R[k];
    }
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < SMPL_J; ++j) {
            const int p = s_par[j];
            if (p < 0) {
                for (int k = 0; k < 9; ++k) s_G[j][k] = s_R[j][k];
                for (int k = 0; k < 3; ++k) s_G[j][9 + k] = J[j * 3 + k];
            } else {
                const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
                mat3_mul(s_G[p], s_R[j], s_G[j]);
                float tr[3];
                mat3_vec(s_G[p], rel, tr);
                for (int k = 0; k < 3; ++k) s_G[j][9 + k] = tr[k] + s_G[p][9 + k];
            }
        }
    }
    __syncthreads();
    if (tid < SMPL_J) {
        // A' = [G.R | G.t - G.R J];  A = A' * A0inv
        const int j = tid;
        float G[12];
        for (int k = 0; k < 12; ++k) G[k] = s_G[j][k];
        float RJ[3];
        mat3_vec(G, J + j * 3, RJ);
        const float tp[3] = {G[9] - RJ[0], G[10] - RJ[1], G[11] - RJ[2]};
        const float* I4 = A0inv + j * 16;
        const float Ri[9] = {I4[0], I4[1], I4[2], I4[4], I4[5], I4[6], I4[8], I4[9], I4[10]};
        const float ti[3] = {I4[3], I4[7], I4[11]};
        float* Ao = A + (int64_t)i * SM_AOUT + j * 12;
        float AR[9], rt[3];
        mat3_mul(G, Ri, AR);
        mat3_vec(G, ti, rt);
        for (int k = 0; k < 9; ++k) Ao[k] = AR[k];
        for (int k = 0; k < 3; ++k) Ao[9 + k] = rt[k] + tp[k];
    }
}

__device__ __forceinline__ void blend_T(const float* __restrict__ Wn, const float* __restrict__ Ab, float* T) {
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = 0.f;
    for (int j = 0; j < SMPL_J; ++j) {
        const float w = Wn[j];
        if (w == 0.f) continue;  // LBS weights are sparse; 0 * finite adds nothing
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] += w * Ab[j * 12 + k];
    }
}

__global__ void __launch_bounds__(SM_THREADS) smpl_points_fwd_kernel(SmplArgs a, const float* __restrict__ A,
                                                                     const float* __restrict__ means,
                                                                     const float* __restrict__ quats, int64_t N,
                                                                     float* __restrict__ world_means,
                                                                     float* __restrict__ world_quats) {
    const int64_t n = (int64_t)blockIdx.x * SM_THREADS + threadIdx.x;
    if (n >= N) return;
    const int b = (int)(n / a.V);
    const float tx = a.trans[b * 3], ty = a.trans[b * 3 + 1], tz = a.trans[b * 3 + 2];
    if (!a.visible[b]) {
        world_means[n * 3] = tx; world_means[n * 3 + 1] = ty; world_means[n * 3 + 2] = tz;
        reinterpret_cast<float4*>(world_quats)[n] = make_float4(1.f, 0.f, 0.f, 0.f);
        return;
    }
    float T[12];
    blend_T(a.W + n * SMPL_J, A + (int64_t)b * SM_AOUT, T);
    const float x = means[n * 3], y = means[n * 3 + 1], z = means[n * 3 + 2];
    world_means[n * 3 + 0] = T[0] * x + T[1] * y + T[2] * z + T[9] + tx;
    world_means[n * 3 + 1] = T[3] * x + T[4] * y + T[5] * z + T[10] + ty;
    world_means[n * 3 + 2] = T[6] * x + T[7] * y + T[8] * z + T[11] + tz;
    float qR[4], qRn[4], qn[4], o[4];
    mat_to_quat(T, qR);
    qnormalize(qR, qRn);
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + n);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    qnormalize(q, qn);
    qmul(qRn, qn, o);
    reinterpret_cast<float4*>(world_quats)[n] = make_float4(o[0], o[1], o[2], o[3]);
}

// grid (max_chunks, I)
__global__ void __launch_bounds__(SM_THREADS) smpl_points_bwd_kernel(
    SmplArgs a, const float* __restrict__ A, const float* __restrict__ means, const float* __restrict__ quats,
    const float* __restrict__ v_world_means, const float* __restrict__ v_world_quats, float* __restrict__ v_means,
    float* __restrict__ v_quats, float* __restrict__ red_partial) {
    __shared__ float s_vT[SM_THREADS][13];
    __shared__ float s_W[SM_THREADS][SMPL_J + 1];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int64_t lo = (int64_t)b * a.V + (int64_t)chunk * SM_CHUNK;
    const int64_t hi = min((int64_t)(b + 1) * a.V, lo + SM_CHUNK);
    if (lo >= hi) return;  // block-uniform
    const bool vis = a.visible[b] != 0;
    const float* Ab = A + (int64_t)b * SM_AOUT;
    // each thread owns up to two of the 291 reduced outputs
    float acc0 = 0.f, acc1 = 0.f;
    const int o0 = threadIdx.x, o1 = threadIdx.x + SM_THREADS;
    for (int64_t base = lo; base < hi; base += SM_THREADS) {
        const int64_t n = base + threadIdx.x;
        float vT[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) vT[k] = 0.f;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (n < hi) {
            gx = v_world_means[n * 3]; gy = v_world_means[n * 3 + 1]; gz = v_world_means[n * 3 + 2];
            if (!vis) {
                v_means[n * 3] = 0.f; v_means[n * 3 + 1] = 0.f; v_means[n * 3 + 2] = 0.f;
                reinterpret_cast<float4*>(v_quats)[n] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float T[12];
                blend_T(a.W + n * SMPL_J, Ab, T);
                const float x = means[n * 3], y = means[n * 3 + 1], z = means[n * 3 + 2];
                v_means[n * 3 + 0] = T[0] * gx + T[3] * gy + T[6] * gz;
                v_means[n * 3 + 1] = T[1] * gx + T[4] * gy + T[7] * gz;
                v_means[n * 3 + 2] = T[2] * gx + T[5] * gy + T[8] * gz;
                vT[0] = gx * x; vT[1] = gx * y; vT[2] = gx * z;
                vT[3] = gy * x; vT[4] = gy * y; vT[5] = gy * z;
                vT[6] = gz * x; vT[7] = gz * y; vT[8] = gz * z;
                vT[9] = gx; vT[10] = gy; vT[11] = gz;
                // quaternion path
                float qR[4], qRn[4], qn[4];
                mat_to_quat(T, qR);
                const float invR = qnormalize(qR, qRn);
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + n);
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(v_world_quats) + n);
                const float q[4] = {q4.x, q4.y, q4.z, q4.w}, vg[4] = {g4.x, g4.y, g4.z, g4.w};
                const float invq = qnormalize(q, qn);
                float v_qRn[4], v_qn[4], v_qR[4], v_q[4], v_m[9];
                qmul_vjp(qRn, qn, vg, v_qRn, v_qn);
                qnormalize_vjp(qn, invq, v_qn, v_q);
                reinterpret_cast<float4*>(v_quats)[n] = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
                qnormalize_vjp(qRn, invR, v_qRn, v_qR);
                mat_to_quat_vjp(T, v_qR, v_m);
#pragma unroll
                for (int k = 0; k < 9; ++k) vT[k] += v_m[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 12; ++k) s_vT[threadIdx.x][k] = vT[k];
        s_vT[threadIdx.x][12] = 0.f;
        for (int j = 0; j < SMPL_J; ++j) s_W[threadIdx.x][j] = (n < hi && vis) ? a.W[n * SMPL_J + j] : 0.f;
        // v_trans contribution rides in slot 12.. via separate sums below
        __syncthreads();
        const int npts = (int)min((int64_t)SM_THREADS, hi - base);
        // outputs 0..287: (j, k) -> sum_pt W[pt][j] * vT[pt][k];  288..290: sum_pt g[pt]
        if (vis && o0 < SM_AOUT) {
            const int j = o0 / 12, k = o0 - j * 12;
            float s = 0.f;
            for (int pt = 0; pt < npts; ++pt) s += s_W[pt][j] * s_vT[pt][k];
            acc0 += s;
        }
        if (vis && o1 < SM_AOUT) {
            const int j = o1 / 12, k = o1 - j * 12;
            float s = 0.f;
            for (int pt = 0; pt < npts; ++pt) s += s_W[pt][j] * s_vT[pt][k];
            acc1 += s;
        }
        // v_trans: reuse s_vT slots 9..11 when visible; for invisible instances stash g there now
        __syncthreads();
        if (!vis) { s_vT[threadIdx.x][9] = gx; s_vT[threadIdx.x][10] = gy; s_vT[threadIdx.x][11] = gz; }
        __syncthreads();
        if (o1 >= SM_AOUT && o1 < SM_RED) {
            const int k = o1 - SM_AOUT;
            float s = 0.f;
            for (int pt = 0; pt < npts; ++pt) s += s_vT[pt][9 + k];
            acc1 += s;
        }
    }
    float* out = red_partial + ((int64_t)b * a.max_chunks + chunk) * SM_RED;
    if (o0 < SM_RED) out[o0] = acc0;
    if (o1 < SM_RED) out[o1] = acc1;
}

// One CTA per instance.  The per-joint work (EMD heads, joint rotations, v_A -> v_G, quaternion VJPs) runs one
// thread per joint, the bulk work (chunk sums of v_A, head-weight outer products, temporal-table VJP) across the
// whole CTA; only the 24-step kinematic chain and its reverse sweep stay on one thread, in shared memory.
// Same arithmetic and summation order as smpl_instance_bwd (emd_math.cuh).
constexpr int SIB_THREADS = 128;
__global__ void __launch_bounds__(SIB_THREADS) smpl_instance_bwd_kernel(
    SmplArgs a, const float* __restrict__ mean_emb, const float* __restrict__ red_partial,
    float* __restrict__ v_theta, float* __restrict__ v_trans, float* __restrict__ v_params_partial,
    float* __restrict__ v_table, float* __restrict__ v_mean_emb) {
    constexpr int IN_MAX = EMD_TDIM_MAX + EMD_GDIM_MAX;
    __shared__ float s_vA[SM_AOUT];
    __shared__ float s_hc[IN_MAX], s_hf[IN_MAX], s_vhc[IN_MAX], s_vhf[IN_MAX];
    __shared__ float s_ac[SMPL_J], s_af[SMPL_J], s_vac[SMPL_J];
    __shared__ float s_G[SMPL_J][12], s_vG[SMPL_J][12], s_R[SMPL_J][9], s_vR[SMPL_J][9];
    __shared__ float s_thn[SMPL_J][4], s_qo[SMPL_J][4], s_thinv[SMPL_J];
    __shared__ int s_par[SMPL_J];
    __shared__ int s_skip;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int nch = (a.V + SM_CHUNK - 1) / SM_CHUNK;
    const int in = a.d + a.g;
    const int pc = smpl_param_count(in);
    float* vp = v_params_partial + (int64_t)i * pc;
    const float* rp = red_partial + (int64_t)i * a.max_chunks * SM_RED;
    for (int k = tid; k < SM_RED; k += SIB_THREADS) {
        float s = 0.f;
        for (int c = 0; c < nch; ++c) s += rp[(int64_t)c * SM_RED + k];
        if (k < SM_AOUT) s_vA[k] = s; else v_trans[i * 3 + (k - SM_AOUT)] = s;
    }
    if (!a.visible[i]) {  // block-uniform
        for (int k = tid; k < SMPL_J * 4; k += SIB_THREADS) v_theta[(int64_t)i * SMPL_J * 4 + k] = 0.f;
        for (int k = tid; k < pc; k += SIB_THREADS) vp[k] = 0.f;
        for (int k = tid; k < a.g; k += SIB_THREADS) v_mean_emb[i * a.g + k] = 0.f;
        return;
    }
    const float* table = a.table + (int64_t)i * a.E * a.d;
    const float* theta = a.theta + (int64_t)i * SMPL_J * 4;
    const float* J = a.J + (int64_t)i * SMPL_J * 3;
    const float* A0inv = a.A0inv + (int64_t)i * SMPL_J * 16;
    TembTaps tc, tf;
    temb_taps(a.t, a.cur_c, a.E, tc);
    temb_taps(a.t, a.cur_f, a.E, tf);
    for (int k = tid; k < in; k += SIB_THREADS) {
        float c, f;
        if (k < a.d) {
            c = 0.f; f = 0.f;
            for (int q = 0; q < 4; ++q) { c += tc.w[q] * table[tc.row[q] * a.d + k]; f += tf.w[q] * table[tf.row[q] * a.d + k]; }
        } else {
            c = f = mean_emb[i * a.g + (k - a.d)];
        }
        s_hc[k] = c; s_hf[k] = f;
    }
    if (tid < SMPL_J) s_par[tid] = a.parents[tid];
    __syncthreads();
    if (tid < SMPL_J) {
        float sc = a.H.c_b[tid], sf = a.H.f_b[tid];
        for (int k = 0; k < in; ++k) { sc += a.H.c_w[tid * in + k] * s_hc[k]; sf += a.H.f_w[tid * in + k] * s_hf[k]; }
        s_ac[tid] = sc; s_af[tid] = sf;
    }
    __syncthreads();
    if (tid == 0) s_skip = (any_nan(s_ac, SMPL_J) || any_nan(s_af, SMPL_J)) ? 1 : 0;
    __syncthreads();
    const bool skip = s_skip != 0;
    if (tid < SMPL_J) {
        const int j = tid;
        float th[4];
        if (skip) { for (int k = 0; k < 4; ++k) th[k] = theta[j * 4 + k]; }
        else {
            const float qc[4] = {cosf(s_ac[j]), 0.f, 0.f, sinf(s_ac[j])}, qf[4] = {cosf(s_af[j]), 0.f, 0.f, sinf(s_af[j])};
            float qo[4];
            qmul(qc, qf, qo);
            qmul(theta + j * 4, qo, th);
            for (int k = 0; k < 4; ++k) s_qo[j][k] = qo[k];
        }
        float thn[4], R[9];
        s_thinv[j] = qnormalize(th, thn);
        qrot(thn, R);
        for (int k = 0; k < 4; ++k) s_thn[j][k] = thn[k];
        for (int k = 0; k < 9; ++k) s_R[j][k] = R[k];
        // v_G from v_A:  A.R = G.R Ri ; A.t = G.R ti + (G.t - G.R J)
        const float* I4 = A0inv + j * 16;
        const float Ri[9] = {I4[0], I4[1], I4[2], I4[4], I4[5], I4[6], I4[8], I4[9], I4[10]};
        const float ti[3] = {I4[3], I4[7], I4[11]};
        const float* vAR = s_vA + j * 12;
        const float* vAt = s_vA + j * 12 + 9;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                float s = vAR[r * 3] * Ri[c * 3] + vAR[r * 3 + 1] * Ri[c * 3 + 1] + vAR[r * 3 + 2] * Ri[c * 3 + 2];
                s += vAt[r] * (ti[c] - J[j * 3 + c]);
                s_vG[j][r * 3 + c] = s;
            }
        for (int k = 0; k < 3; ++k) s_vG[j][9 + k] = vAt[k];
    }
    __syncthreads();
    if (tid == 0) {
        // forward chain (parents precede children), then the reverse sweep
        for (int j = 0; j < SMPL_J; ++j) {
            const int p = s_par[j];
            if (p < 0) {
                for (int k = 0; k < 9; ++k) s_G[j][k] = s_R[j][k];
                for (int k = 0; k < 3; ++k) s_G[j][9 + k] = J[j * 3 + k];
            } else {
                const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
                mat3_mul(s_G[p], s_R[j], s_G[j]);
                float tr[3];
                mat3_vec(s_G[p], rel, tr);
                for (int k = 0; k < 3; ++k) s_G[j][9 + k] = tr[k] + s_G[p][9 + k];
            }
        }
        for (int j = SMPL_J - 1; j >= 0; --j) {
            const int p = s_par[j];
            if (p < 0) {
                for (int k = 0; k < 9; ++k) s_vR[j][k] = s_vG[j][k];
            } else {
                const float rel[3] = {J[j * 3] - J[p * 3], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        s_vG[p][r * 3 + c] += s_vG[j][r * 3] * s_R[j][c * 3] + s_vG[j][r * 3 + 1] * s_R[j][c * 3 + 1] +
                                              s_vG[j][r * 3 + 2] * s_R[j][c * 3 + 2] + s_vG[j][9 + r] * rel[c];
                        s_vR[j][r * 3 + c] = s_G[p][0 * 3 + r] * s_vG[j][0 * 3 + c] + s_G[p][1 * 3 + r] * s_vG[j][1 * 3 + c] +
                                             s_G[p][2 * 3 + r] * s_vG[j][2 * 3 + c];
                    }
                for (int k = 0; k < 3; ++k) s_vG[p][9 + k] += s_vG[j][9 + k];
            }
        }
    }
    __syncthreads();
    if (tid < SMPL_J) {
        const int j = tid;
        float vthn[4], vth[4];
        qrot_vjp(s_thn[j], s_vR[j], vthn);
        qnormalize_vjp(s_thn[j], s_thinv[j], vthn, vth);
        float* vt = v_theta + (int64_t)i * SMPL_J * 4 + j * 4;
        if (skip) {
            for (int k = 0; k < 4; ++k) vt[k] = vth[k];
            s_vac[j] = 0.0f;
        } else {
            float vqo[4];
            qmul_vjp(theta + j * 4, s_qo[j], vth, vt, vqo);
            s_vac[j] = -s_qo[j][3] * vqo[0] + s_qo[j][0] * vqo[3];
        }
    }
    __syncthreads();
    // head gradients: v_W[o][k] = v_ac[o] h[k], v_b[o] = v_ac[o]; layout c_w | c_b | f_w | f_b  (all zero when skipped)
    const int wsz = SMPL_J * in;
    for (int e = tid; e < pc; e += SIB_THREADS) {
        float v = 0.f;
        if (!skip) {
            if (e < wsz) v = s_vac[e / in] * s_hc[e % in];
            else if (e < wsz + SMPL_J) v = s_vac[e - wsz];
            else if (e < 2 * wsz + SMPL_J) { const int q = e - wsz - SMPL_J; v = s_vac[q / in] * s_hf[q % in]; }
            else v = s_vac[e - 2 * wsz - SMPL_J];
        }
        vp[e] = v;
    }
    for (int k = tid; k < in; k += SIB_THREADS) {
        float c = 0.f, f = 0.f;
        if (!skip)
            for (int o = 0; o < SMPL_J; ++o) { c += a.H.c_w[o * in + k] * s_vac[o]; f += a.H.f_w[o * in + k] * s_vac[o]; }
        s_vhc[k] = c; s_vhf[k] = f;
    }
    __syncthreads();
    // temporal-table VJP: thread k owns column k of this instance's [E][d] slab (taps applied in the reference order)
    float* vtab = v_table + (int64_t)i * a.E * a.d;
    for (int k = tid; k < a.d; k += SIB_THREADS) {
        for (int q = 0; q < 4; ++q) if (tc.w[q] != 0.0f) vtab[tc.row[q] * a.d + k] += tc.w[q] * s_vhc[k];
        for (int q = 0; q < 4; ++q) if (tf.w[q] != 0.0f) vtab[tf.row[q] * a.d + k] += tf.w[q] * s_vhf[k];
    }
    for (int k = tid; k < a.g; k += SIB_THREADS) v_mean_emb[i * a.g + k] = s_vhc[a.d + k] + s_vhf[a.d + k];
}

__global__ void smpl_params_reduce_kernel(const float* __restrict__ partial, int I, int pc, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pc) return;
    float s = 0.f;
    for (int i = 0; i < I; ++i) s += partial[(int64_t)i * pc + k];
    out[k] = s;
}

__global__ void smpl_embed_bwd_kernel(const float* __restrict__ v_mean_emb, int g, int V, int64_t N,
                                      float* __restrict__ v_emb) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * g) return;
    const int64_t n = e / g;
    const int k = (int)(e - n * g);
    v_emb[e] = v_mean_emb[(n / V) * g + k] / (float)V;
}

int fill(SmplArgs& a, const float* table, int I, int E, int d, int g, int V, float t, int cur_c, int cur_f,
         const float* const* heads, const float* theta, const float* trans, const uint8_t* visible, const float* J,
         const float* A0inv, const float* W, const int* parents) {
    EMD_CHECK_ARG(I >= 1 && E >= 2 && V >= 1, "smpl: need I >= 1, E >= 2, V >= 1");
    EMD_CHECK_ARG(d >= 1 && d <= EMD_TDIM_MAX && g >= 0 && g <= EMD_GDIM_MAX, "smpl: temporal dim <= %d, embedding dim <= %d",
                  EMD_TDIM_MAX, EMD_GDIM_MAX);
    a.table = table; a.I = I; a.E = E; a.d = d; a.g = g; a.V = V; a.t = t; a.cur_c = cur_c; a.cur_f = cur_f;
    a.H.c_w = heads[0]; a.H.c_b = heads[1]; a.H.f_w = heads[2]; a.H.f_b = heads[3];
    a.theta = theta; a.trans = trans; a.visible = visible; a.J = J; a.A0inv = A0inv; a.W = W; a.parents = parents;
    a.max_chunks = (V + SM_CHUNK - 1) / SM_CHUNK;
    EMD_CHECK_ARG(a.max_chunks <= 65535, "smpl: too many points per instance");
    return EMD_OK;
}

}  // namespace

extern "C" int emd_smpl_param_count(int d, int g) { return smpl_param_count(d + g); }
extern "C" int emd_smpl_max_chunks(int V) { return (V + SM_CHUNK - 1) / SM_CHUNK; }
extern "C" int emd_smpl_reduce_width() { return SM_RED; }

// heads: HOST array of 4 DEVICE pointers {track_smpl_c.weight[24,d+g], .bias[24], track_smpl_f.weight, .bias}.
// scratch: seg_partial [I*max_chunks*g].  Saved for backward: mean_emb [I,g], A [I,24,12].
extern "C" int emd_smpl_deform_fwd(const float* means, const float* quats, const float* embeddings, const float* table,
                                   const float* const* heads, const float* theta, const float* trans,
                                   const uint8_t* visible, const float* J, const float* A0inv, const float* W,
                                   const int* parents, int I, int V, int E, int d, int g, float t, int cur_c,
                                   int cur_f, float* seg_partial, float* mean_emb, float* A, float* world_means,
                                   float* world_quats, cudaStream_t stream) {
    SmplArgs a;
    int rc = fill(a, table, I, E, d, g, V, t, cur_c, cur_f, heads, theta, trans, visible, J, A0inv, W, parents);
    if (rc != EMD_OK) return rc;
    if (!emd_aligned(quats, 16) || !emd_aligned(world_quats, 16)) { emd_set_error("smpl_fwd: quats must be 16-B aligned"); return EMD_ERR_ALIGN; }
    const int64_t N = (int64_t)I * V;
    dim3 sg(a.max_chunks, I);
    EMD_LAUNCH(EK_SMPL_FWD, stream, smpl_segmean_kernel<<<sg, SM_THREADS, 0, stream>>>(embeddings, g, V, a.max_chunks, seg_partial));
    EMD_LAUNCH(EK_SMPL_FWD, stream, smpl_instance_fwd_kernel<<<I, SIF_THREADS, 0, stream>>>(a, seg_partial, mean_emb, A));
    EMD_LAUNCH(EK_SMPL_FWD, stream, smpl_points_fwd_kernel<<<(unsigned)emd_cdiv(N, SM_THREADS), SM_THREADS, 0, stream>>>(a, A, means, quats, N, world_means, world_quats));
    EMD_CHECK_LAUNCH("smpl_deform_fwd");
    return EMD_OK;
}

// scratch: red_partial [I*max_chunks*reduce_width], params_partial [I*param_count]; v_table zero-filled by the caller.
extern "C" int emd_smpl_deform_bwd(const float* means, const float* quats, const float* table,
                                   const float* const* heads, const float* theta, const float* trans,
                                   const uint8_t* visible, const float* J, const float* A0inv, const float* W,
                                   const int* parents, int I, int V, int E, int d, int g, float t, int cur_c,
                                   int cur_f, const float* mean_emb, const float* A, const float* v_world_means,
                                   const float* v_world_quats, float* red_partial, float* params_partial,
                                   float* v_means, float* v_quats, float* v_embeddings, float* v_table,
                                   float* v_params, float* v_theta, float* v_trans, float* v_mean_emb,
                                   cudaStream_t stream) {
    SmplArgs a;
    int rc = fill(a, table, I, E, d, g, V, t, cur_c, cur_f, heads, theta, trans, visible, J, A0inv, W, parents);
    if (rc != EMD_OK) return rc;
    if (!emd_aligned(quats, 16) || !emd_aligned(v_world_quats, 16) || !emd_aligned(v_quats, 16)) {
        emd_set_error("smpl_bwd: quats tensors must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    const int64_t N = (int64_t)I * V;
    dim3 sg(a.max_chunks, I);
    EMD_LAUNCH(EK_SMPL_BWD, stream, smpl_points_bwd_kernel<<<sg, SM_THREADS, 0, stream>>>(a, A, means, quats, v_world_means, v_world_quats, v_means,
                                                          v_quats, red_partial));
    EMD_LAUNCH(EK_SMPL_BWD, stream, smpl_instance_bwd_kernel<<<I, SIB_THREADS, 0, stream>>>(a, mean_emb, red_partial, v_theta, v_trans,
                                                               params_partial, v_table, v_mean_emb));
    const int pc = smpl_param_count(d + g);
    EMD_LAUNCH(EK_SMPL_BWD, stream, smpl_params_reduce_kernel<<<(pc + 127) / 128, 128, 0, stream>>>(params_partial, I, pc, v_params));
    if (g > 0) smpl_embed_bwd_kernel<<<(unsigned)emd_cdiv(N * g, 256), 256, 0, stream>>>(v_mean_emb, g, V, N, v_embeddings);
    EMD_CHECK_LAUNCH("smpl_deform_bwd");
    return EMD_OK;
}

// ---- gradient of the LBS weights (only when they come from the trainable voxel deformer) ------------------
__global__ void __launch_bounds__(SM_THREADS) smpl_weight_grad_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                                                                      const uint8_t* __restrict__ visible, const float* __restrict__ W,
                                                                      const float* __restrict__ A, int V, int64_t N,
                                                                      const float* __restrict__ v_world_means,
                                                                      const float* __restrict__ v_world_quats, float* __restrict__ v_W) {
    const int64_t n = (int64_t)blockIdx.x * SM_THREADS + threadIdx.x;
    if (n >= N) return;
    const int b = (int)(n / V);
    float out[SMPL_J];
    if (!visible[b]) {
#pragma unroll
        for (int j = 0; j < SMPL_J; ++j) out[j] = 0.f;
    } else {
        const float x[3] = {means[n * 3], means[n * 3 + 1], means[n * 3 + 2]};
        const float g[3] = {v_world_means[n * 3], v_world_means[n * 3 + 1], v_world_means[n * 3 + 2]};
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + n);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(v_world_quats) + n);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w}, vg[4] = {g4.x, g4.y, g4.z, g4.w};
        smpl_point_weight_grad(W + n * SMPL_J, A + (int64_t)b * SM_AOUT, x, q, g, vg, out);
    }
    float4* dst = reinterpret_cast<float4*>(v_W + n * SMPL_J);   // 96 B per point: 16-byte aligned rows
#pragma unroll
    for (int j4 = 0; j4 < SMPL_J / 4; ++j4) dst[j4] = make_float4(out[j4 * 4], out[j4 * 4 + 1], out[j4 * 4 + 2], out[j4 * 4 + 3]);
}

// v_W[I,V,24] = d loss / d W of emd_smpl_deform_fwd (W is the output of VoxelDeformer.forward when
// `use_voxel_deformer` is on: OmniRe/models/human_body.py:174-179).  A[I,24,12]: the skinning matrices the forward wrote.
extern "C" int emd_smpl_weight_grad(const float* means, const float* quats, const uint8_t* visible, const float* W,
                                    const float* A, int I, int V, const float* v_world_means, const float* v_world_quats,
                                    float* v_W, cudaStream_t stream) {
    EMD_CHECK_ARG(I >= 0 && V >= 0, "smpl_weight_grad: I=%d V=%d", I, V);
    const int64_t N = (int64_t)I * V;
    if (N == 0) return EMD_OK;
    EMD_CHECK_ARG(means && quats && visible && W && A && v_world_means && v_world_quats && v_W, "smpl_weight_grad: null argument");
    if (!emd_aligned(quats, 16) || !emd_aligned(v_world_quats, 16) || !emd_aligned(v_W, 16)) {
        emd_set_error("smpl_weight_grad: quats / v_world_quats / v_W must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    EMD_LAUNCH(EK_SMPL_BWD, stream,
               (smpl_weight_grad_kernel<<<(unsigned)emd_cdiv(N, SM_THREADS), SM_THREADS, 0, stream>>>(
                   means, quats, visible, W, A, V, N, v_world_means, v_world_quats, v_W)));
    EMD_CHECK_LAUNCH("smpl_weight_grad");
    return EMD_OK;
}
