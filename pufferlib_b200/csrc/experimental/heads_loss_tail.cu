// heads_loss_tail.cu -- EXPERIMENTAL, NOT PART OF libpuffer_b200.so, NOT YET RUN ON HARDWARE (round-2 groundwork).
//
// One pass over the hidden layer for everything between the encoder GEMM and the dW_enc GEMM of the minibatch update
// (clean_pufferl._DefaultMLPUpdate), i.e. the present chain
//     out = hidden @ W_heads^T + b          (8-column head GEMM, reads hidden: 268 MB at M = 524288)
//     pb_ppo_loss(out, ...) -> dOut, statistics            (csrc/ppo_loss.cu, clean_pufferl.py:202-238)
//     pb_mlp_tail_backward(dOut, hidden) -> dPre, dW_heads, db_enc, db_heads   (csrc/mlp_tail.cu; reads hidden again)
// as ONE kernel that reads `hidden` once and writes dPre once: 536 MB instead of 805 MB + two launches' latency,
// measured today as 43 + 22 + 100 us per minibatch against a 82 us roofline for the fused traffic.  It does not depend
// on the number of input features, so it also serves the configs the planned tcgen05 update kernel (128 features) does
// not.
//
// Structure = csrc/mlp_tail.cu's TMA-ring kernel (column-owner lanes: lane l owns hidden columns 4l..4l+3, dPre / dW_heads
// / db_enc without any cross-lane traffic) preceded, per 64-row chunk, by a head phase built like csrc/policy_mlp.cu:
//   * rows land in shared memory as 512-byte bulk copies, pitch 136 floats (conflict-free 64-bit fragment loads);
//   * warps 0..3 each take 16 rows: out[16][8] = hidden . W_heads^T with 16 mma.sync.m16n8k8 TF32 (k slots (t, t+4) mapped
//     to adjacent columns (2t, 2t+1) in both operands), quad-shuffle gather, lanes t < 2 finish one row each: the
//     pb_ppo_loss row math (same expressions, same ATen tie rules) -> dOut[8] into shared memory, statistics in registers;
//   * after a CTA barrier all 8 warps run the tail loop on the chunk with dOut read from shared memory.
// Note: the head GEMM here rounds `hidden` to TF32 (cvt.rna) and W_heads is expected pre-rounded or is truncated by the
// tensor core -- the same precision class as the cuBLAS TF32 head GEMM it replaces.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../tma.cuh"

namespace {

char g_err2[512] = "";
void set_err2(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err2, sizeof(g_err2), fmt, ap);
    va_end(ap);
}

constexpr int H = 128, NO = 8;
constexpr int THREADS = 256, WARPS = 8;
constexpr int CHUNK = 64;                    // rows per pipeline stage
constexpr int STAGES = 2;
constexpr int PITCH = H + 8;                 // floats per shared row (544 B)
constexpr int ROWS_PER_BLOCK = 512;
constexpr int PSTRIDE = NO * H + H + NO;     // per-block partials: dW_heads | db_enc | db_heads

struct HltParams {
    const float* hidden;        // [M][128] post-ReLU
    const float* w_heads;       // [8][128]  (n_act logit rows | value row | zero rows)
    const float* b_heads;       // [8]
    const int64_t* actions;     // [M]
    const float* old_logprobs;  // [M]
    const float* adv;           // [M]
    const float* returns;       // [M]
    const float* old_values;    // [M]
    float* dpre;                // [M][128]
    float* partials;            // [blocks][PSTRIDE]
    double* stats;              // [8]: sums of pg, v (before the 0.5), entropy, -logratio, (ratio-1)-logratio, clipped
    int64_t m;
    int n_act;
    float clip, vclip, vf_coef, ent_coef;
    int clip_vloss;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// The row math of k_ppo_loss<PACKED> (csrc/ppo_loss.cu): z[0..n_act) logits, z[n_act] value -> dOut[8] (scaled by 1/M)
// and the six per-row statistics.
struct RowStats { float pg, v, ent, okl, kl, clipped; };
__device__ __forceinline__ RowStats ppo_row(const float (&zin)[8], const HltParams& p, int64_t i, float (&gro)[8]) {
    float z[8];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        z[k] = zin[k];
        if (k < p.n_act) mx = fmaxf(mx, z[k]);
    }
    float v_packed = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k == p.n_act) v_packed = z[k];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < p.n_act) sum += expf(z[k] - mx);
    const float lse = mx + logf(sum);
    int a = (int)p.actions[i];
    a = a < 0 ? 0 : (a >= p.n_act ? p.n_act - 1 : a);
    float ent = 0.f, nl_a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < p.n_act) {
            const float nl = z[k] - lse, pk = expf(nl);
            ent -= pk * nl;
            if (k == a) nl_a = nl;
            z[k] = nl;
        }
    const float logratio = nl_a - p.old_logprobs[i];
    const float ratio = expf(logratio);
    const float adv = p.adv[i];
    const float pg1 = -adv * ratio;
    const float rc = fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
    const float pg2 = -adv * rc;
    const float pg = fmaxf(pg1, pg2);
    const float in_range = (ratio >= 1.f - p.clip && ratio <= 1.f + p.clip) ? 1.f : 0.f;
    float g_ratio;
    if (pg1 > pg2) g_ratio = -adv;
    else if (pg1 < pg2) g_ratio = -adv * in_range;
    else g_ratio = 0.5f * (-adv) + 0.5f * (-adv * in_range);
    const float inv_m = 1.0f / (float)p.m;
    const float g_nlp = g_ratio * ratio * inv_m;
    const float ret = p.returns[i];
    const float dv = v_packed - ret;
    float vl, g_v;
    if (p.clip_vloss) {
        const float ov = p.old_values[i];
        const float d = v_packed - ov;
        const float dc = fminf(fmaxf(d, -p.vclip), p.vclip);
        const float vc = ov + dc;
        const float vu = dv * dv, vcl = (vc - ret) * (vc - ret);
        vl = fmaxf(vu, vcl);
        const float v_in = (d >= -p.vclip && d <= p.vclip) ? 1.f : 0.f;
        const float gu = 2.f * dv, gc = 2.f * (vc - ret) * v_in;
        g_v = vu > vcl ? gu : (vu < vcl ? gc : 0.5f * (gu + gc));
    } else {
        vl = dv * dv;
        g_v = 2.f * dv;
    }
    const float gv_out = 0.5f * p.vf_coef * g_v * inv_m;
    const float g_ent = p.ent_coef * inv_m;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        gro[k] = 0.f;
        if (k < p.n_act) {
            const float pk = expf(z[k]);
            gro[k] = g_nlp * ((k == a ? 1.f : 0.f) - pk) + g_ent * pk * (z[k] + ent);
        }
        if (k == p.n_act) gro[k] = gv_out;
    }
    RowStats s;
    s.pg = pg; s.v = vl; s.ent = ent; s.okl = -logratio; s.kl = (ratio - 1.f) - logratio;
    s.clipped = fabsf(ratio - 1.f) > p.clip ? 1.f : 0.f;
    return s;
}

__global__ void __launch_bounds__(THREADS, 2) k_heads_loss_tail(HltParams p) {
    extern __shared__ __align__(128) unsigned char dyn[];
    float* s_h = reinterpret_cast<float*>(dyn);                                        // [STAGES][CHUNK][PITCH]
    float* s_d = s_h + STAGES * CHUNK * PITCH;                                         // [CHUNK][NO]  dOut of the chunk
    float* s_wh = s_d + CHUNK * NO;                                                    // [NO][PITCH]
    float* s_red = s_wh + NO * PITCH;                                                  // [WARPS][PSTRIDE]
    __shared__ uint64_t bars[STAGES];
    __shared__ double s_stats[WARPS][6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    const int64_t row0 = (int64_t)blockIdx.x * ROWS_PER_BLOCK;
    const int64_t row_end = row0 + ROWS_PER_BLOCK < p.m ? row0 + ROWS_PER_BLOCK : p.m;
    const int n_chunks = (int)((row_end - row0 + CHUNK - 1) / CHUNK);

    // one 512-byte bulk copy per row into the padded stage; issued by warp 0 (2 rows per lane)
    auto issue = [&](int c) {
        const int st = c % STAGES;
        const int64_t r = row0 + (int64_t)c * CHUNK;
        const int rows = (int)((row_end - r) < CHUNK ? (row_end - r) : CHUNK);
        if (lane == 0) mbar_expect_tx(&bars[st], (uint32_t)rows * H * 4u);
        __syncwarp();
        for (int i = lane; i < rows; i += 32)
            tma_load_1d(s_h + ((size_t)st * CHUNK + i) * PITCH, p.hidden + (r + i) * H, H * 4u, &bars[st]);
    };
    if (tid == 0) {
        for (int st = 0; st < STAGES; ++st) mbar_init(&bars[st], 1);
        mbar_fence_init();
    }
    // head matrix for the mma B fragments, rounded to TF32 once
    for (int i = tid; i < NO * H; i += THREADS) s_wh[(i >> 7) * PITCH + (i & 127)] = __uint_as_float(to_tf32(p.w_heads[i]));
    __syncthreads();
    if (warp == 0)
        for (int c = 0; c < STAGES && c < n_chunks; ++c) issue(c);

    // ---- per-thread constants
    // tail loop: lane owns columns 4*lane .. 4*lane+3 (as in csrc/mlp_tail.cu)
    float4 w[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) w[k] = *reinterpret_cast<const float4*>(p.w_heads + (int64_t)k * H + 4 * lane);
    float4 acc_w[NO], acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc_o[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) { acc_o[k] = 0.f; acc_w[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    // head phase (warps 0..3): b_heads[2t], b_heads[2t+1] are the two output columns this lane accumulates
    float bias_lo = 0.f, bias_hi = 0.f;
    if (warp < 4) {
        bias_lo = p.b_heads[2 * t];
        bias_hi = p.b_heads[2 * t + 1];
    }
    double st_pg = 0, st_v = 0, st_ent = 0, st_okl = 0, st_kl = 0, st_clip = 0;

    for (int c = 0; c < n_chunks; ++c) {
        const int st = c % STAGES;
        mbar_wait(&bars[st], (uint32_t)((c / STAGES) & 1));
        const int64_t r0 = row0 + (int64_t)c * CHUNK;
        const int rows = (int)((row_end - r0) < CHUNK ? (row_end - r0) : CHUNK);
        const float* ch = s_h + (size_t)st * CHUNK * PITCH;

        // ---- head phase: warp w < 4 owns rows 16w .. 16w+15 of the chunk
        if (warp < 4) {
            float out[4] = {0.f, 0.f, 0.f, 0.f};
            const float* xa = ch + (16 * warp + g) * PITCH + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const float2 x0 = *reinterpret_cast<const float2*>(xa + 8 * ks);
                const float2 x1 = *reinterpret_cast<const float2*>(xa + 8 * PITCH + 8 * ks);
                const uint32_t a[4] = {to_tf32(x0.x), to_tf32(x1.x), to_tf32(x0.y), to_tf32(x1.y)};
                // B[k slot][n = g] with slots (t, t+4) <-> columns (8ks + 2t, 8ks + 2t + 1), like A
                const float2 bv = *reinterpret_cast<const float2*>(s_wh + g * PITCH + 8 * ks + 2 * t);
                mma_tf32(out, a, __float_as_uint(bv.x), __float_as_uint(bv.y));
            }
            // out: (row g, outputs 2t, 2t+1), (row g+8, outputs 2t, 2t+1); add the bias, gather a row's 8 outputs
            out[0] += bias_lo; out[1] += bias_hi; out[2] += bias_lo; out[3] += bias_hi;
            float z0[8], z1[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int src = (lane & ~3) | q;
                z0[2 * q] = __shfl_sync(0xffffffffu, out[0], src);
                z0[2 * q + 1] = __shfl_sync(0xffffffffu, out[1], src);
                z1[2 * q] = __shfl_sync(0xffffffffu, out[2], src);
                z1[2 * q + 1] = __shfl_sync(0xffffffffu, out[3], src);
            }
            if (t < 2) {   // lane t == 0 finishes row g, lane t == 1 row g + 8
                const int rl = 16 * warp + g + 8 * t;
                if (rl < rows) {
                    float z[8], gro[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[k] = t ? z1[k] : z0[k];
                    const RowStats s = ppo_row(z, p, r0 + rl, gro);
                    *reinterpret_cast<float4*>(s_d + rl * NO) = make_float4(gro[0], gro[1], gro[2], gro[3]);
                    *reinterpret_cast<float4*>(s_d + rl * NO + 4) = make_float4(gro[4], gro[5], gro[6], gro[7]);
                    st_pg += s.pg; st_v += s.v; st_ent += s.ent; st_okl += s.okl; st_kl += s.kl; st_clip += s.clipped;
                }
            }
        }
        __syncthreads();                                   // dOut of the chunk is in shared memory

        // ---- tail loop (all 8 warps): 8 rows per warp
#pragma unroll 2
        for (int i = 0; i < CHUNK / WARPS; ++i) {
            const int rl = warp + i * WARPS;
            if (rl < rows) {
                const float4 d0 = *reinterpret_cast<const float4*>(s_d + rl * NO);
                const float4 d1 = *reinterpret_cast<const float4*>(s_d + rl * NO + 4);
                const float d[NO] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const float4 h = *reinterpret_cast<const float4*>(ch + rl * PITCH + 4 * lane);
                float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < NO; ++k) {
                    acc_o[k] += d[k];
                    gr.x = fmaf(d[k], w[k].x, gr.x); gr.y = fmaf(d[k], w[k].y, gr.y);
                    gr.z = fmaf(d[k], w[k].z, gr.z); gr.w = fmaf(d[k], w[k].w, gr.w);
                    acc_w[k].x = fmaf(d[k], h.x, acc_w[k].x); acc_w[k].y = fmaf(d[k], h.y, acc_w[k].y);
                    acc_w[k].z = fmaf(d[k], h.z, acc_w[k].z); acc_w[k].w = fmaf(d[k], h.w, acc_w[k].w);
                }
                gr.x = h.x > 0.f ? gr.x : 0.f; gr.y = h.y > 0.f ? gr.y : 0.f;
                gr.z = h.z > 0.f ? gr.z : 0.f; gr.w = h.w > 0.f ? gr.w : 0.f;
                acc_b.x += gr.x; acc_b.y += gr.y; acc_b.z += gr.z; acc_b.w += gr.w;
                __stcs(reinterpret_cast<float4*>(p.dpre + (r0 + rl) * H + 4 * lane), gr);
            }
        }
        __syncthreads();                                   // stage st and s_d are free again
        if (warp == 0 && c + STAGES < n_chunks) issue(c + STAGES);
    }

    // ---- block reduction of the gradient partials (as csrc/mlp_tail.cu) and of the loss statistics
    float* mine = s_red + (size_t)warp * PSTRIDE;
#pragma unroll
    for (int k = 0; k < NO; ++k) *reinterpret_cast<float4*>(mine + k * H + 4 * lane) = acc_w[k];
    *reinterpret_cast<float4*>(mine + NO * H + 4 * lane) = acc_b;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NO; ++k) mine[NO * H + H + k] = acc_o[k];
    double sv[6] = {st_pg, st_v, st_ent, st_okl, st_kl, st_clip};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        double x = sv[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if (lane == 0) s_stats[warp][q] = x;
    }
    __syncthreads();
    float* outp = p.partials + (int64_t)blockIdx.x * PSTRIDE;
    for (int j = tid; j < PSTRIDE; j += THREADS) {
        float sum = 0.f;
#pragma unroll
        for (int wq = 0; wq < WARPS; ++wq) sum += s_red[(size_t)wq * PSTRIDE + j];
        outp[j] = sum;
    }
    if (tid < 6) {
        double tsum = 0;
        for (int wq = 0; wq < 4; ++wq) tsum += s_stats[wq][tid];      // only the head-phase warps hold statistics
        atomicAdd(p.stats + tid, tsum);
    }
}

// deterministic second stage, as in csrc/mlp_tail.cu: one warp per output element
__global__ void __launch_bounds__(256) k_reduce_partials2(const float* __restrict__ partials, int n_blocks, int pstride,
                                                         float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= pstride) return;
    float s = 0.f;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(int64_t)b * pstride + j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[j] = s;
}

constexpr size_t SMEM_BYTES = ((size_t)STAGES * CHUNK * PITCH + CHUNK * NO + NO * PITCH + (size_t)WARPS * PSTRIDE) * 4;

}  // namespace

extern "C" const char* pbx_hlt_last_error(void) { return g_err2; }
extern "C" size_t pbx_heads_loss_tail_workspace_bytes(int64_t m) {
    const int64_t blocks = m > 0 ? (m + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK : 1;
    return (size_t)blocks * PSTRIDE * sizeof(float);
}

// hidden [m][128] (post-ReLU) + the packed heads -> dpre [m][128], grads_out = [dW_heads 8x128 | db_enc 128 | db_heads 8]
// (the layout of pb_mlp_tail_backward), stats8 (the sums pb_ppo_loss produces; zeroed here).  Arguments as in pb_ppo_loss.
extern "C" int pbx_heads_loss_tail(const float* hidden, const float* w_heads, const float* b_heads, const int64_t* actions,
                                   const float* old_logprobs, const float* advantages, const float* returns,
                                   const float* old_values, int64_t m, int32_t n_act, float clip_coef, int32_t clip_vloss,
                                   float vf_clip_coef, float vf_coef, float ent_coef, float* dpre, float* grads_out,
                                   double* stats8, void* workspace, size_t workspace_bytes, void* stream) {
    if (!hidden || !w_heads || !b_heads || !actions || !old_logprobs || !advantages || !returns || !dpre || !grads_out ||
        !stats8 || !workspace || m < 1 || n_act < 1 || n_act > 7 || (clip_vloss && !old_values) ||
        ((uintptr_t)hidden & 15) || ((uintptr_t)dpre & 15) || ((uintptr_t)w_heads & 15) ||
        workspace_bytes < pbx_heads_loss_tail_workspace_bytes(m)) {
        set_err2("pbx_heads_loss_tail: bad arguments");
        return -1;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (int)((m + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
    HltParams p{hidden, w_heads, b_heads, actions, old_logprobs, advantages, returns, old_values, dpre, (float*)workspace,
                stats8, m, n_act, clip_coef, vf_clip_coef, vf_coef, ent_coef, clip_vloss};
    cudaError_t e = cudaMemsetAsync(stats8, 0, 8 * sizeof(double), s);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_heads_loss_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) {
        k_heads_loss_tail<<<blocks, THREADS, SMEM_BYTES, s>>>(p);
        k_reduce_partials2<<<(PSTRIDE * 32 + 255) / 256, 256, 0, s>>>((const float*)workspace, blocks, PSTRIDE, grads_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        set_err2("pbx_heads_loss_tail: %s", cudaGetErrorString(e));
        return -2;
    }
    return 0;
}
