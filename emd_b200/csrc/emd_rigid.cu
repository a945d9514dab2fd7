// K1a: EMD motion-embedding deformation of rigid nodes, forward and backward.
//
// Replaces the three per-instance Python loops of RigidNodes.transform_means /
// transform_quats (OmniRe/models/nodes/rigid.py:478-568; several hundred tiny
// launches and >= 2 host syncs per instance per step) with:
//   segmean  : per-instance mean of the Gaussian motion embeddings (fixed-order
//              chunked reduction over the instance-sorted point list)
//   instance : one thread per instance -- temporal-embedding taps, track_* heads,
//              yaw-offset quaternion, final pose (R, t, Q)        [emd_math.cuh]
//   points   : one thread per Gaussian -- x_w = R x + t, q_w = Q (x) normalize(q)
// and the mirrored backward (per-instance gradients are reduced in a fixed
// order: no float atomics, bit-reproducible).  HBM-bound on the per-point
// streams: 64 B/Gaussian forward (mean 12 + quat 16 + id 8 read; 12 + 16 written).
#include "common.cuh"
#include "emd_math.cuh"

namespace {

constexpr int RG_THREADS = 256;
constexpr int RG_CHUNK = 2048;  // points per (instance, chunk) block
constexpr int RG_INST = 16;     // floats of per-instance pose: R[9] t[3] Q[4]

// fixed-order block sum of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* s_scratch /*[RG_THREADS/32][NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s_scratch[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float s = 0.f;
            for (int w = 0; w < RG_THREADS / 32; ++w) s += s_scratch[w * NV + k];
            v[k] = s;
        }
    }
}

__global__ void __launch_bounds__(RG_THREADS) rigid_segmean_kernel(const float* __restrict__ emb, int g,
                                                                   const int64_t* __restrict__ order,
                                                                   const int64_t* __restrict__ seg_start,
                                                                   int max_chunks, float* __restrict__ partial) {
    __shared__ float s_scratch[RG_THREADS / 32 * EMD_GDIM_MAX];
    const int inst = blockIdx.y, chunk = blockIdx.x;
    const int64_t lo = seg_start[inst] + (int64_t)chunk * RG_CHUNK;
    const int64_t hi = min(seg_start[inst + 1], lo + RG_CHUNK);
    float acc[EMD_GDIM_MAX];
#pragma unroll
    for (int k = 0; k < EMD_GDIM_MAX; ++k) acc[k] = 0.f;
    for (int64_t p = lo + threadIdx.x; p < hi; p += RG_THREADS) {
        const int64_t n = order[p];
#pragma unroll
        for (int k = 0; k < EMD_GDIM_MAX; ++k)
            if (k < g) acc[k] += emb[n * g + k];
    }
    block_sum<EMD_GDIM_MAX>(acc, s_scratch);
    if (threadIdx.x == 0)
        for (int k = 0; k < g; ++k) partial[((int64_t)inst * max_chunks + chunk) * g + k] = acc[k];
}

struct RigidArgs {
    const float* table;  // weight [I][E][d]
    int I, E, d, g;
    float t;
    int cur_c, cur_f;
    RigidHeads H;
    const float* pose_q_means;  // [I,4]
    const float* pose_q_quats;  // [I,4]
    const float* pose_t;        // [I,3]
    const int64_t* seg_start;   // [I+1]
    int max_chunks;
};

// One warp-pair CTA per instance: features and the 8 head outputs in parallel, the quaternion algebra on one
// thread.  Same arithmetic / summation order as rigid_instance_fwd (emd_math.cuh).
constexpr int RI_THREADS = 64;
__device__ __forceinline__ void rigid_features(const RigidArgs& a, int i, const float* m /*smem [g]*/, float* s_hc,
                                               float* s_hf, TembTaps& tc, TembTaps& tf) {
    const float* table = a.table + (int64_t)i * a.E * a.d;
    temb_taps(a.t, a.cur_c, a.E, tc);
    temb_taps(a.t, a.cur_f, a.E, tf);
    for (int k = threadIdx.x; k < a.d + a.g; k += RI_THREADS) {
        float c, f;
        if (k < a.d) {
            c = 0.f; f = 0.f;
            for (int q = 0; q < 4; ++q) { c += tc.w[q] * table[tc.row[q] * a.d + k]; f += tf.w[q] * table[tf.row[q] * a.d + k]; }
        } else {
            c = f = m[k - a.d];
        }
        s_hc[k] = c; s_hf[k] = f;
    }
}
// head outputs y[0..2] = trans_c, y[3..5] = trans_f, y[6] = rot_c, y[7] = rot_f  (thread o < 8 computes y[o])
__device__ __forceinline__ float rigid_head(const RigidArgs& a, int o, const float* s_hc, const float* s_hf) {
    const int in = a.d + a.g;
    const float* W; const float* h; float s;
    if (o < 3) { W = a.H.trans_c_w + o * in; s = a.H.trans_c_b[o]; h = s_hc; }
    else if (o < 6) { W = a.H.trans_f_w + (o - 3) * in; s = a.H.trans_f_b[o - 3]; h = s_hf; }
    else if (o == 6) { W = a.H.rot_c_w; s = a.H.rot_c_b[0]; h = s_hc; }
    else { W = a.H.rot_f_w; s = a.H.rot_f_b[0]; h = s_hf; }
    for (int k = 0; k < in; ++k) s += W[k] * h[k];
    return s;
}

__global__ void __launch_bounds__(RI_THREADS) rigid_instance_fwd_kernel(RigidArgs a, const float* __restrict__ seg_partial,
                                                                        float* __restrict__ mean_emb,
                                                                        float* __restrict__ inst_out) {
    __shared__ float s_m[EMD_GDIM_MAX], s_hc[EMD_TDIM_MAX + EMD_GDIM_MAX], s_hf[EMD_TDIM_MAX + EMD_GDIM_MAX], s_y[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int64_t cnt = a.seg_start[i + 1] - a.seg_start[i];
    const int nch = (int)((cnt + RG_CHUNK - 1) / RG_CHUNK);
    if (tid < a.g) {
        float s = 0.f;
        for (int c = 0; c < nch; ++c) s += seg_partial[((int64_t)i * a.max_chunks + c) * a.g + tid];
        s_m[tid] = s / (float)cnt;  // 0/0 = NaN for an empty instance, exactly like torch.mean of an empty slice
        mean_emb[i * a.g + tid] = s_m[tid];
    }
    __syncthreads();
    TembTaps tc, tf;
    rigid_features(a, i, s_m, s_hc, s_hf, tc, tf);
    __syncthreads();
    if (tid < 8) s_y[tid] = rigid_head(a, tid, s_hc, s_hf);
    __syncthreads();
    if (tid != 0) return;
    const float* pose_q_means = a.pose_q_means + i * 4;
    const float* pose_q_quats = a.pose_q_quats + i * 4;
    const float* pose_t = a.pose_t + i * 3;
    const float dt[3] = {s_y[0] + s_y[3], s_y[1] + s_y[4], s_y[2] + s_y[5]};
    const float ac = s_y[6], af = s_y[7];
    const float qc[4] = {cosf(ac), 0.f, 0.f, sinf(ac)}, qf[4] = {cosf(af), 0.f, 0.f, sinf(af)};
    float qoff[4], qn[4], R[9], Qg[4], Q[4];
    qmul(qc, qf, qoff);
    qnormalize(pose_q_means, qn);
    qrot(qn, R);
    const bool skip_t = any_nan(dt, 3);
    const bool skip_q = any_nan(qoff, 4);
    if (skip_q) { for (int k = 0; k < 4; ++k) Qg[k] = pose_q_quats[k]; }
    else qmul(pose_q_quats, qoff, Qg);
    qnormalize(Qg, Q);
    float* out = inst_out + i * RG_INST;
    for (int k = 0; k < 9; ++k) out[k] = R[k];
    for (int k = 0; k < 3; ++k) out[9 + k] = pose_t[k] + (skip_t ? 0.0f : dt[k]);
    for (int k = 0; k < 4; ++k) out[12 + k] = Q[k];
}

__global__ void __launch_bounds__(RG_THREADS) rigid_points_fwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const int64_t* __restrict__ point_ids,
    const float* __restrict__ inst_out, int64_t N, float* __restrict__ world_means, float* __restrict__ world_quats) {
    const int64_t n = (int64_t)blockIdx.x * RG_THREADS + threadIdx.x;
    if (n >= N) return;
    const float* P = inst_out + point_ids[n] * RG_INST;
    const float x = means[n * 3 + 0], y = means[n * 3 + 1], z = means[n * 3 + 2];
    world_means[n * 3 + 0] = P[0] * x + P[1] * y + P[2] * z + P[9];
    world_means[n * 3 + 1] = P[3] * x + P[4] * y + P[5] * z + P[10];
    world_means[n * 3 + 2] = P[6] * x + P[7] * y + P[8] * z + P[11];
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + n);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float qn[4], o[4];
    qnormalize(q, qn);
    qmul(P + 12, qn, o);
    reinterpret_cast<float4*>(world_quats)[n] = make_float4(o[0], o[1], o[2], o[3]);
}

// grid (max_chunks, I): per-point input grads + fixed-order partial sums of the
// per-instance pose gradient (v_R 9, v_t 3, v_Q 4)
__global__ void __launch_bounds__(RG_THREADS) rigid_points_bwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const int64_t* __restrict__ order,
    const int64_t* __restrict__ seg_start, const float* __restrict__ inst_out, int max_chunks,
    const float* __restrict__ v_world_means, const float* __restrict__ v_world_quats, float* __restrict__ v_means,
    float* __restrict__ v_quats, float* __restrict__ pose_partial) {
    __shared__ float s_scratch[RG_THREADS / 32 * RG_INST];
    const int inst = blockIdx.y, chunk = blockIdx.x;
    const int64_t lo = seg_start[inst] + (int64_t)chunk * RG_CHUNK;
    const int64_t hi = min(seg_start[inst + 1], lo + RG_CHUNK);
    if (lo >= hi) return;  // block-uniform
    const float* P = inst_out + inst * RG_INST;
    float acc[RG_INST];
#pragma unroll
    for (int k = 0; k < RG_INST; ++k) acc[k] = 0.f;
    for (int64_t p = lo + threadIdx.x; p < hi; p += RG_THREADS) {
        const int64_t n = order[p];
        const float x = means[n * 3 + 0], y = means[n * 3 + 1], z = means[n * 3 + 2];
        const float gx = v_world_means[n * 3 + 0], gy = v_world_means[n * 3 + 1], gz = v_world_means[n * 3 + 2];
        v_means[n * 3 + 0] = P[0] * gx + P[3] * gy + P[6] * gz;
        v_means[n * 3 + 1] = P[1] * gx + P[4] * gy + P[7] * gz;
        v_means[n * 3 + 2] = P[2] * gx + P[5] * gy + P[8] * gz;
        acc[0] += gx * x; acc[1] += gx * y; acc[2] += gx * z;
        acc[3] += gy * x; acc[4] += gy * y; acc[5] += gy * z;
        acc[6] += gz * x; acc[7] += gz * y; acc[8] += gz * z;
        acc[9] += gx; acc[10] += gy; acc[11] += gz;
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(quats) + n);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(v_world_quats) + n);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w}, vg[4] = {g4.x, g4.y, g4.z, g4.w};
        float qn[4], vQ[4], vqn[4], vq[4];
        const float inv = qnormalize(q, qn);
        qmul_vjp(P + 12, qn, vg, vQ, vqn);
        qnormalize_vjp(qn, inv, vqn, vq);
        reinterpret_cast<float4*>(v_quats)[n] = make_float4(vq[0], vq[1], vq[2], vq[3]);
        acc[12] += vQ[0]; acc[13] += vQ[1]; acc[14] += vQ[2]; acc[15] += vQ[3];
    }
    block_sum<RG_INST>(acc, s_scratch);
    if (threadIdx.x == 0)
        for (int k = 0; k < RG_INST; ++k) pose_partial[((int64_t)inst * max_chunks + chunk) * RG_INST + k] = acc[k];
}

// One CTA per instance (see rigid_instance_fwd_kernel); same arithmetic as rigid_instance_bwd (emd_math.cuh).
__global__ void __launch_bounds__(RI_THREADS) rigid_instance_bwd_kernel(
    RigidArgs a, const float* __restrict__ mean_emb, const float* __restrict__ pose_partial,
    float* __restrict__ v_pose_q_means, float* __restrict__ v_pose_q_quats, float* __restrict__ v_pose_t,
    float* __restrict__ v_params_partial, float* __restrict__ v_table, float* __restrict__ v_mean_emb) {
    constexpr int IN_MAX = EMD_TDIM_MAX + EMD_GDIM_MAX;
    __shared__ float s_m[EMD_GDIM_MAX], s_hc[IN_MAX], s_hf[IN_MAX], s_vhc[IN_MAX], s_vhf[IN_MAX], s_y[8], s_vy[8], s_v[RG_INST];
    __shared__ int s_skip[2];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int in = a.d + a.g;
    const int64_t cnt = a.seg_start[i + 1] - a.seg_start[i];
    const int nch = (int)((cnt + RG_CHUNK - 1) / RG_CHUNK);
    if (tid < RG_INST) {
        float s = 0.f;
        for (int c = 0; c < nch; ++c) s += pose_partial[((int64_t)i * a.max_chunks + c) * RG_INST + tid];
        s_v[tid] = s;
    }
    if (tid < a.g) s_m[tid] = mean_emb[i * a.g + tid];
    __syncthreads();
    TembTaps tc, tf;
    rigid_features(a, i, s_m, s_hc, s_hf, tc, tf);
    __syncthreads();
    if (tid < 8) s_y[tid] = rigid_head(a, tid, s_hc, s_hf);
    __syncthreads();
    if (tid == 0) {
        const float* pose_q_means = a.pose_q_means + i * 4;
        const float* pose_q_quats = a.pose_q_quats + i * 4;
        const float* v_R = s_v; const float* v_t = s_v + 9; const float* v_Q = s_v + 12;
        const float dt[3] = {s_y[0] + s_y[3], s_y[1] + s_y[4], s_y[2] + s_y[5]};
        const float ac = s_y[6], af = s_y[7];
        const float qc[4] = {cosf(ac), 0.f, 0.f, sinf(ac)}, qf[4] = {cosf(af), 0.f, 0.f, sinf(af)};
        float qoff[4];
        qmul(qc, qf, qoff);
        const bool skip_t = any_nan(dt, 3), skip_q = any_nan(qoff, 4);
        float qn[4], vqn[4];
        const float inv = qnormalize(pose_q_means, qn);
        qrot_vjp(qn, v_R, vqn);
        qnormalize_vjp(qn, inv, vqn, v_pose_q_means + i * 4);
        for (int k = 0; k < 3; ++k) v_pose_t[i * 3 + k] = v_t[k];
        for (int k = 0; k < 3; ++k) { s_vy[k] = skip_t ? 0.f : v_t[k]; s_vy[3 + k] = s_vy[k]; }
        float Qg[4], Qn[4], v_Qg[4];
        if (skip_q) { for (int k = 0; k < 4; ++k) Qg[k] = pose_q_quats[k]; }
        else qmul(pose_q_quats, qoff, Qg);
        const float invQ = qnormalize(Qg, Qn);
        qnormalize_vjp(Qn, invQ, v_Q, v_Qg);
        float v_a = 0.f;
        if (skip_q) {
            for (int k = 0; k < 4; ++k) v_pose_q_quats[i * 4 + k] = v_Qg[k];
        } else {
            float v_qoff[4];
            qmul_vjp(pose_q_quats, qoff, v_Qg, v_pose_q_quats + i * 4, v_qoff);
            // qoff = qc (x) qf = (cos(ac+af), 0, 0, sin(ac+af)):  d/d ac = d/d af = (-qoff.z, 0, 0, qoff.w)
            v_a = -qoff[3] * v_qoff[0] + qoff[0] * v_qoff[3];
        }
        s_vy[6] = v_a; s_vy[7] = v_a;
        s_skip[0] = skip_t; s_skip[1] = skip_q;
    }
    __syncthreads();
    // parameter partial, layout rot_c_w[in] rot_c_b[1] rot_f_w[in] rot_f_b[1] trans_c_w[3in] trans_c_b[3] trans_f_w[3in] trans_f_b[3]
    const int pc = rigid_param_count(in);
    float* vp = v_params_partial + (int64_t)i * pc;
    for (int e = tid; e < pc; e += RI_THREADS) {
        float v;
        int q = e;
        bool rot = true;   // a skipped branch contributes exact zeros (its features may be NaN: empty instance)
        if (q < in) v = s_vy[6] * s_hc[q];
        else if ((q -= in) < 1) v = s_vy[6];
        else if ((q -= 1) < in) v = s_vy[7] * s_hf[q];
        else if ((q -= in) < 1) v = s_vy[7];
        else if ((q -= 1) < 3 * in) { v = s_vy[q / in] * s_hc[q % in]; rot = false; }
        else if ((q -= 3 * in) < 3) { v = s_vy[q]; rot = false; }
        else if ((q -= 3) < 3 * in) { v = s_vy[3 + q / in] * s_hf[q % in]; rot = false; }
        else { v = s_vy[3 + (q - 3 * in)]; rot = false; }
        if (s_skip[rot ? 1 : 0]) v = 0.f;
        vp[e] = v;
    }
    for (int k = tid; k < in; k += RI_THREADS) {
        float c = 0.f, f = 0.f;
        for (int o = 0; o < 3; ++o) { c += a.H.trans_c_w[o * in + k] * s_vy[o]; f += a.H.trans_f_w[o * in + k] * s_vy[3 + o]; }
        c += a.H.rot_c_w[k] * s_vy[6];
        f += a.H.rot_f_w[k] * s_vy[7];
        s_vhc[k] = c; s_vhf[k] = f;
    }
    __syncthreads();
    float* vtab = v_table + (int64_t)i * a.E * a.d;
    for (int k = tid; k < a.d; k += RI_THREADS) {
        for (int q = 0; q < 4; ++q) if (tc.w[q] != 0.0f) vtab[tc.row[q] * a.d + k] += tc.w[q] * s_vhc[k];
        for (int q = 0; q < 4; ++q) if (tf.w[q] != 0.0f) vtab[tf.row[q] * a.d + k] += tf.w[q] * s_vhf[k];
    }
    for (int k = tid; k < a.g; k += RI_THREADS) v_mean_emb[i * a.g + k] = s_vhc[a.d + k] + s_vhf[a.d + k];
}

// v_params[k] = sum_i partial[i][k], fixed order
__global__ void params_reduce_kernel(const float* __restrict__ partial, int I, int pc, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pc) return;
    float s = 0.f;
    for (int i = 0; i < I; ++i) s += partial[(int64_t)i * pc + k];
    out[k] = s;
}

__global__ void embed_bwd_kernel(const float* __restrict__ v_mean_emb, const int64_t* __restrict__ point_ids,
                                 const int64_t* __restrict__ seg_start, int g, int64_t N, float* __restrict__ v_emb) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * g) return;
    const int64_t n = e / g;
    const int k = (int)(e - n * g);
    const int64_t id = point_ids[n];
    const float cnt = (float)(seg_start[id + 1] - seg_start[id]);
    v_emb[e] = v_mean_emb[id * g + k] / cnt;
}

int fill_args(RigidArgs& a, const float* table, int I, int E, int d, int g, float t, int cur_c, int cur_f,
              const float* const* heads, const float* pose_q_means, const float* pose_q_quats, const float* pose_t,
              const int64_t* seg_start, int max_chunks) {
    EMD_CHECK_ARG(I >= 1 && E >= 2, "rigid: need I >= 1 instances and E >= 2 table rows");
    EMD_CHECK_ARG(d >= 1 && d <= EMD_TDIM_MAX && g >= 0 && g <= EMD_GDIM_MAX, "rigid: temporal dim <= %d, embedding dim <= %d",
                  EMD_TDIM_MAX, EMD_GDIM_MAX);
    EMD_CHECK_ARG(cur_c >= 1 && cur_f >= 1, "rigid: bad current embedding counts");
    EMD_CHECK_ARG(max_chunks >= 1 && max_chunks <= 65535, "rigid: bad chunk count");
    a.table = table; a.I = I; a.E = E; a.d = d; a.g = g; a.t = t; a.cur_c = cur_c; a.cur_f = cur_f;
    a.H.rot_c_w = heads[0]; a.H.rot_c_b = heads[1]; a.H.rot_f_w = heads[2]; a.H.rot_f_b = heads[3];
    a.H.trans_c_w = heads[4]; a.H.trans_c_b = heads[5]; a.H.trans_f_w = heads[6]; a.H.trans_f_b = heads[7];
    a.pose_q_means = pose_q_means; a.pose_q_quats = pose_q_quats; a.pose_t = pose_t;
    a.seg_start = seg_start; a.max_chunks = max_chunks;
    return EMD_OK;
}

}  // namespace

extern "C" int emd_rigid_chunk_size() { return RG_CHUNK; }
extern "C" int emd_rigid_param_count(int d, int g) { return rigid_param_count(d + g); }

// heads: HOST array of 8 DEVICE pointers {rot_c_w, rot_c_b, rot_f_w, rot_f_b, trans_c_w, trans_c_b, trans_f_w, trans_f_b}
// order/seg_start: points sorted by instance id (stable) and the I+1 segment boundaries.
// scratch: seg_partial [I*max_chunks*g] floats.  Saved for backward: mean_emb [I,g], inst_out [I,16].
extern "C" int emd_rigid_deform_fwd(const float* means, const float* quats, const float* embeddings,
                                    const int64_t* point_ids, const int64_t* order, const int64_t* seg_start,
                                    const float* table, const float* const* heads, const float* pose_q_means,
                                    const float* pose_q_quats, const float* pose_t, int64_t N, int I, int E, int d,
                                    int g, float t, int cur_c, int cur_f, int max_chunks, float* seg_partial,
                                    float* mean_emb, float* inst_out, float* world_means, float* world_quats,
                                    cudaStream_t stream) {
    RigidArgs a;
    int rc = fill_args(a, table, I, E, d, g, t, cur_c, cur_f, heads, pose_q_means, pose_q_quats, pose_t, seg_start, max_chunks);
    if (rc != EMD_OK) return rc;
    if (!emd_aligned(quats, 16) || !emd_aligned(world_quats, 16)) { emd_set_error("rigid_fwd: quats must be 16-B aligned"); return EMD_ERR_ALIGN; }
    dim3 sg(max_chunks, I);
    EMD_LAUNCH(EK_RIGID_FWD, stream, rigid_segmean_kernel<<<sg, RG_THREADS, 0, stream>>>(embeddings, g, order, seg_start, max_chunks, seg_partial));
    EMD_LAUNCH(EK_RIGID_FWD, stream, rigid_instance_fwd_kernel<<<I, RI_THREADS, 0, stream>>>(a, seg_partial, mean_emb, inst_out));
    if (N > 0)
        EMD_LAUNCH(EK_RIGID_FWD, stream, rigid_points_fwd_kernel<<<(unsigned)emd_cdiv(N, RG_THREADS), RG_THREADS, 0, stream>>>(
            means, quats, point_ids, inst_out, N, world_means, world_quats));
    EMD_CHECK_LAUNCH("rigid_deform_fwd");
    return EMD_OK;
}

// scratch: pose_partial [I*max_chunks*16], params_partial [I*param_count]; v_table must be zero-filled by the caller.
extern "C" int emd_rigid_deform_bwd(const float* means, const float* quats, const int64_t* point_ids,
                                    const int64_t* order, const int64_t* seg_start, const float* table,
                                    const float* const* heads, const float* pose_q_means, const float* pose_q_quats,
                                    const float* pose_t, int64_t N, int I, int E, int d, int g, float t, int cur_c,
                                    int cur_f, int max_chunks, const float* mean_emb, const float* inst_out,
                                    const float* v_world_means, const float* v_world_quats, float* pose_partial,
                                    float* params_partial, float* v_means, float* v_quats, float* v_embeddings,
                                    float* v_table, float* v_params, float* v_pose_q_means, float* v_pose_q_quats,
                                    float* v_pose_t, float* v_mean_emb, cudaStream_t stream) {
    RigidArgs a;
    int rc = fill_args(a, table, I, E, d, g, t, cur_c, cur_f, heads, pose_q_means, pose_q_quats, pose_t, seg_start, max_chunks);
    if (rc != EMD_OK) return rc;
    if (!emd_aligned(quats, 16) || !emd_aligned(v_world_quats, 16) || !emd_aligned(v_quats, 16)) {
        emd_set_error("rigid_bwd: quats tensors must be 16-B aligned");
        return EMD_ERR_ALIGN;
    }
    dim3 sg(max_chunks, I);
    EMD_LAUNCH(EK_RIGID_BWD, stream, rigid_points_bwd_kernel<<<sg, RG_THREADS, 0, stream>>>(means, quats, order, seg_start, inst_out, max_chunks,
                                                           v_world_means, v_world_quats, v_means, v_quats, pose_partial));
    EMD_LAUNCH(EK_RIGID_BWD, stream, rigid_instance_bwd_kernel<<<I, RI_THREADS, 0, stream>>>(a, mean_emb, pose_partial, v_pose_q_means,
                                                                v_pose_q_quats, v_pose_t, params_partial, v_table, v_mean_emb));
    const int pc = rigid_param_count(d + g);
    EMD_LAUNCH(EK_RIGID_BWD, stream, params_reduce_kernel<<<(pc + 127) / 128, 128, 0, stream>>>(params_partial, I, pc, v_params));
    if (N > 0 && g > 0)
        EMD_LAUNCH(EK_RIGID_BWD, stream, embed_bwd_kernel<<<(unsigned)emd_cdiv(N * g, 256), 256, 0, stream>>>(v_mean_emb, point_ids, seg_start, g, N, v_embeddings));
    EMD_CHECK_LAUNCH("rigid_deform_bwd");
    return EMD_OK;
}

// ---- stand-alone temporal embedding (S3Gaussian: one shared table, per-call scalar time on the device) ----
namespace {
__global__ void temb_fwd_kernel(const float* __restrict__ table, int E, int d, const float* __restrict__ t_dev, int cur,
                                float* __restrict__ emb) {
    TembTaps taps;
    temb_taps(*t_dev, cur, E, taps);
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += taps.w[k] * table[taps.row[k] * d + j];
        emb[j] = s;
    }
}
// d emb / d t: the taps are piecewise linear in the (reflected) row coordinate iy = t * (cur-1):
//   emb = (1-wy1) Wc[y0] + wy1 Wc[y0+1]  =>  d emb/d iy = Wc[y0+1] - Wc[y0]
__global__ void temb_bwd_kernel(const float* __restrict__ table, int E, int d, const float* __restrict__ t_dev, int cur,
                                const float* __restrict__ v_emb, float* __restrict__ v_table, float* __restrict__ v_t) {
    __shared__ float s_red[32];
    const float t = *t_dev;
    TembTaps taps;
    temb_taps(t, cur, E, taps);
    const float wy0 = taps.w[0] + taps.w[1], wy1 = taps.w[2] + taps.w[3];
    float local = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const float g = v_emb[j];
        for (int k = 0; k < 4; ++k)
            if (taps.w[k] != 0.f) v_table[taps.row[k] * d + j] += taps.w[k] * g;  // one thread per column: no race
        // resized rows: Wc0 = (w0 T[r0] + w1 T[r1]) / wy0 etc.; guard the degenerate weights
        const float a = wy0 > 0.f ? (taps.w[0] * table[taps.row[0] * d + j] + taps.w[1] * table[taps.row[1] * d + j]) / wy0 : 0.f;
        const float b = wy1 > 0.f ? (taps.w[2] * table[taps.row[2] * d + j] + taps.w[3] * table[taps.row[3] * d + j]) / wy1 : a;
        local += g * (b - a);
    }
    for (int o = 16; o >= 1; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0 && v_t) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += s_red[w];
        // iy = reflect(t * (cur-1)): slope +-(cur-1) depending on the reflection parity
        const float span = (float)(cur - 1);
        float sign = 1.f;
        if (span > 0.f) {
            const float iy = t * span;
            const int flips = (int)floorf(fabsf(iy) / span);
            sign = (flips % 2 == 0) ? 1.f : -1.f;
            if (iy < 0.f) sign = -sign;
        }
        *v_t = s * sign * span;
    }
}
}  // namespace

extern "C" int emd_temb_fwd(const float* table, int E, int d, const float* t_dev, int cur, float* emb,
                            cudaStream_t stream) {
    EMD_CHECK_ARG(E >= 2 && d >= 1 && cur >= 1, "temb_fwd: bad sizes");
    EMD_LAUNCH(EK_MISC, stream, temb_fwd_kernel<<<1, 64, 0, stream>>>(table, E, d, t_dev, cur, emb));
    EMD_CHECK_LAUNCH("temb_fwd");
    return EMD_OK;
}

extern "C" int emd_temb_bwd(const float* table, int E, int d, const float* t_dev, int cur, const float* v_emb,
                            float* v_table, float* v_t, cudaStream_t stream) {
    EMD_CHECK_ARG(E >= 2 && d >= 1 && cur >= 1, "temb_bwd: bad sizes");
    EMD_LAUNCH(EK_MISC, stream, temb_bwd_kernel<<<1, 64, 0, stream>>>(table, E, d, t_dev, cur, v_emb, v_table, v_t));
    EMD_CHECK_LAUNCH("temb_bwd");
    return EMD_OK;
}
