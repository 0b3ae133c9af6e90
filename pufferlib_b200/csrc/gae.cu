// gae.cu -- GAE as a single-pass affine suffix scan with decoupled look-back (sm_100a).
//
// Replaces /root/reference/clean_pufferl.py:163-169 (sort_training_data + 3 numpy gathers) and
// /root/reference/c_gae.pyx:11-32 (compute_gae).  Recurrence over the sorted batch f = e*H + t, B = N*H:
//     A[B-1] = 0;   A[f] = a_f + b_f * A[f+1]
//     nnt = 1 - d[f+1];  a_f = r[f+1] + (gamma*v[f+1])*nnt - v[f];  b_f = (gamma*lambda)*nnt
// Affine maps compose associatively, (a,b)o(a',b') = (a + b a', b b'), so the chain is a suffix scan.
//
// Data layout: inputs are the arrival-order rollout tensors x[t*N + e] (time-major); the sorted order is never
// materialised.  A tile is a contiguous f-range: E whole envs (E a power of two, all H steps) or, for N == 1 /
// very long horizons, a flat chunk.  The loader walks the tile with e fastest, so every warp load is a run of E
// consecutive floats of one time row (coalesced, full 32 B sectors for E >= 8), and transposes through shared
// memory (row pitch H|1: conflict-free for both the e-fastest stores and the f-order reads).  Each warp scans
// its 32-element rounds with shuffles, (P,Q) partials stay in registers, tiles chain through a 16-byte status
// word each (aggregate / inclusive) with a warp-wide look-back window that stops as soon as the accumulated
// slope is exactly 0 (any done flag, or (gamma*lambda)^k underflow).  Outputs are written in sorted order,
// 128 B per warp store.  HBM traffic = 12 B read + 4 (or 8 with returns) B written per agent-step.
#include "pb_common.cuh"

namespace {

constexpr int GAE_THREADS = 256;
constexpr int GAE_WARPS = GAE_THREADS / 32;

struct __align__(16) GaeStatus {
    float P, Q, X;
    uint32_t flag;  // 0 = empty, 1 = aggregate (P,Q) valid, 2 = inclusive X valid
};

struct GaeHeader {
    uint32_t ticket, done, pad0, pad1;
};

struct GaeParams {
    const float* r;
    const float* v;
    const float* d;
    float* adv;
    float* ret;
    int64_t N, H, B;
    float gamma, gl;
    int E, logE;       // envs per tile (power of two) or 0 in flat mode
    int L;             // tile length in elements (E*H or flat chunk)
    int pitch;         // shared row pitch (E-mode) ; flat mode: unused
    uint32_t magicH;   // ceil(2^32 / H) for f_local / H (E-mode)
    int numTiles;
    GaeHeader* hdr;
    GaeStatus* status;
};

__device__ __forceinline__ void compose(float& a, float& b, float a2, float b2) {
    // (a,b) o (a2,b2): first apply the later map (a2,b2), then this one
    a = fmaf(b, a2, a);
    b = b * b2;
}

template <int RW>
__global__ void __launch_bounds__(GAE_THREADS) k_gae(GaeParams p) {
    extern __shared__ float smem[];
    __shared__ int s_tile;
    __shared__ float s_halo[3];
    __shared__ float s_wP[GAE_WARPS], s_wQ[GAE_WARPS];
    __shared__ float s_carry;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = p.numTiles - 1 - (int)atomicAdd(&p.hdr->ticket, 1u);  // suffix order: last tile first
    __syncthreads();
    const int tile = s_tile;

    // ---- tile geometry
    int64_t f0;       // first sorted index of the tile
    int Lt;           // elements in this tile
    int64_t e0 = 0;   // first env (E-mode)
    int Et = 0;       // envs in this tile (E-mode)
    const bool emode = p.E > 0;
    if (emode) {
        e0 = (int64_t)tile * p.E;
        Et = (int)min((int64_t)p.E, p.N - e0);
        f0 = e0 * p.H;
        Lt = Et * (int)p.H;
    } else {
        f0 = (int64_t)tile * p.L;
        Lt = (int)min((int64_t)p.L, p.B - f0);
    }
    const int stride_arr = emode ? p.E * p.pitch : p.L;
    float* sR = smem;
    float* sV = smem + stride_arr;
    float* sD = smem + 2 * stride_arr;

    // ---- load (transposing) : arrival order x[t*N+e] -> shared [e_local][t]
    if (emode) {
        const int total = p.E * (int)p.H;
        const int mask = p.E - 1;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += GAE_THREADS) {
            const int el = idx & mask, t = idx >> p.logE;
            if (el < Et) {
                const int64_t g = (int64_t)t * p.N + e0 + el;
                const int sp = el * p.pitch + t;
                sR[sp] = __ldcs(p.r + g);
                sV[sp] = __ldcs(p.v + g);
                sD[sp] = __ldcs(p.d + g);
            }
        }
    } else if (p.N == 1) {
#pragma unroll 4
        for (int idx = tid; idx < Lt; idx += GAE_THREADS) {
            sR[idx] = __ldcs(p.r + f0 + idx);
            sV[idx] = __ldcs(p.v + f0 + idx);
            sD[idx] = __ldcs(p.d + f0 + idx);
        }
    } else {  // long-horizon fallback: strided gathers
        for (int idx = tid; idx < Lt; idx += GAE_THREADS) {
            const int64_t f = f0 + idx, e = f / p.H, t = f - e * p.H;
            const int64_t g = t * p.N + e;
            sR[idx] = p.r[g];
            sV[idx] = p.v[g];
            sD[idx] = p.d[g];
        }
    }
    if (tid == 0) {  // halo: the element after the tile (the chain crosses env and tile boundaries)
        const int64_t fn = f0 + Lt;
        if (fn < p.B) {
            const int64_t e = fn / p.H, t = fn - e * p.H;
            const int64_t g = t * p.N + e;
            s_halo[0] = p.r[g];
            s_halo[1] = p.v[g];
            s_halo[2] = p.d[g];
        } else {
            s_halo[0] = s_halo[1] = 0.f;
            s_halo[2] = 1.f;
        }
    }
    __syncthreads();

    // ---- per-element maps into registers
    const int R = (Lt + 31) >> 5;                 // rounds of 32 in this tile
    const int Rw = (R + GAE_WARPS - 1) / GAE_WARPS;  // rounds per warp (<= RW)
    const int r_begin = warp * Rw;
    float a[RW], b[RW];
    const int dp = p.pitch - (int)p.H;            // shared index = i + (i / H) * dp   (E-mode)
#pragma unroll
    for (int k = 0; k < RW; ++k) {
        const int i = ((r_begin + k) << 5) + lane;
        a[k] = 0.f;
        b[k] = 1.f;  // identity for padding lanes
        if (k < Rw && i < Lt) {
            int sp0 = i, sp1 = i + 1;
            if (emode) {
                sp0 = i + (int)__umulhi((uint32_t)i, p.magicH) * dp;
                sp1 = i + 1 + (int)__umulhi((uint32_t)(i + 1), p.magicH) * dp;
            }
            float r1, v1, d1;
            if (i + 1 < Lt) {
                r1 = sR[sp1];
                v1 = sV[sp1];
                d1 = sD[sp1];
            } else {
                r1 = s_halo[0];
                v1 = s_halo[1];
                d1 = s_halo[2];
            }
            const float v0 = sV[sp0];
            const float nnt = __fsub_rn(1.0f, d1);
            // c_gae.pyx:28-29 association, no FMA contraction inside an element
            a[k] = __fsub_rn(__fadd_rn(r1, __fmul_rn(__fmul_rn(p.gamma, v1), nnt)), v0);
            b[k] = __fmul_rn(p.gl, nnt);
            if (f0 + i == p.B - 1) {  // A[B-1] = 0
                a[k] = 0.f;
                b[k] = 0.f;
            }
        }
    }

    // ---- warp-level suffix scan, rounds from last to first; (a,b) become tile-local partials (P,Q) w.r.t. the
    //      value entering this warp's range from the right
    float cP = 0.f, cQ = 1.f;
#pragma unroll
    for (int k = RW - 1; k >= 0; --k) {
        if (k < Rw) {
            float x = a[k], y = b[k];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float x2 = __shfl_down_sync(0xffffffffu, x, off);
                const float y2 = __shfl_down_sync(0xffffffffu, y, off);
                if (lane + off < 32) compose(x, y, x2, y2);
            }
            compose(x, y, cP, cQ);
            a[k] = x;
            b[k] = y;
            cP = __shfl_sync(0xffffffffu, x, 0);
            cQ = __shfl_sync(0xffffffffu, y, 0);
        }
    }
    if (lane == 0) {
        s_wP[warp] = cP;
        s_wQ[warp] = cQ;
    }
    __syncthreads();

    // ---- tile aggregate + decoupled look-back (warp 0)
    if (warp == 0) {
        float tP = 0.f, tQ = 1.f;  // composition of all warps, in order 0..7
#pragma unroll
        for (int w = GAE_WARPS - 1; w >= 0; --w) {
            float x = s_wP[w], y = s_wQ[w];
            compose(x, y, tP, tQ);
            tP = x;
            tQ = y;
        }
        GaeStatus* st = p.status;
        float carry = 0.f;
        if (tile == p.numTiles - 1) {
            if (lane == 0) {
                st[tile].X = tP;  // A beyond the batch is 0
                __threadfence();
                pb_st_release(&st[tile].flag, 2u);
            }
        } else {
            if (lane == 0) {
                st[tile].P = tP;
                st[tile].Q = tQ;
                __threadfence();
                pb_st_release(&st[tile].flag, 1u);
            }
            // window of 32 successor tiles per iteration
            float accP = 0.f, accQ = 1.f;  // composition of the tiles already walked
            int base = tile + 1;
            bool finished = false;
            while (!finished) {
                const int j = base + lane;
                uint32_t fl = 2u;
                float jP = 0.f, jQ = 0.f;  // beyond the last tile: inclusive value 0
                if (j < p.numTiles) {
                    uint32_t polls = 0;
                    do {
                        fl = pb_ld_acquire(&st[j].flag);
                        if (++polls == (1u << 27)) __trap();   // seconds of polling: abort rather than hang the GPU
                    } while (fl == 0u);
                    // status words share 128 B lines with their neighbours: read through L2 (.cg), never a stale L1 line
                    if (fl == 2u) { jP = __ldcg(&st[j].X); jQ = 0.f; }
                    else { jP = __ldcg(&st[j].P); jQ = __ldcg(&st[j].Q); }
                }
                // an inclusive tile or an exactly-zero slope ends the chain: A = P regardless of what follows
                const unsigned stop = __ballot_sync(0xffffffffu, fl == 2u || jQ == 0.f);
                const int last = stop ? (__ffs(stop) - 1) : 31;
                float x = (lane <= last) ? jP : 0.f, y = (lane <= last) ? jQ : 1.f;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float x2 = __shfl_down_sync(0xffffffffu, x, off);
                    const float y2 = __shfl_down_sync(0xffffffffu, y, off);
                    if (lane + off < 32) compose(x, y, x2, y2);
                }
                x = __shfl_sync(0xffffffffu, x, 0);
                y = __shfl_sync(0xffffffffu, y, 0);
                compose(accP, accQ, x, y);
                finished = stop != 0u;
                base += 32;
            }
            carry = accP;  // accQ == 0 here: value of A at the first element after this tile
            if (lane == 0) {
                st[tile].X = fmaf(tQ, carry, tP);
                __threadfence();
                pb_st_release(&st[tile].flag, 2u);
            }
        }
        if (lane == 0) s_carry = carry;
    }
    __syncthreads();

    // ---- carry entering this warp's range = later warps' aggregates applied to the tile carry
    float cin = s_carry;
    for (int w = GAE_WARPS - 1; w > warp; --w) cin = fmaf(s_wQ[w], cin, s_wP[w]);

    // ---- outputs in sorted order (coalesced 128 B per warp store)
#pragma unroll
    for (int k = 0; k < RW; ++k) {
        const int i = ((r_begin + k) << 5) + lane;
        if (k < Rw && i < Lt) {
            const float A = fmaf(b[k], cin, a[k]);
            __stcs(p.adv + f0 + i, A);
            if (p.ret) {
                const int sp0 = emode ? i + (int)__umulhi((uint32_t)i, p.magicH) * dp : i;
                __stcs(p.ret + f0 + i, A + sV[sp0]);
            }
        }
    }

    // ---- self-cleaning workspace: the last block to finish zeroes the header and every status word
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&p.hdr->done, 1u);
        s_tile = (prev == (uint32_t)p.numTiles - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (s_tile) {
        for (int j = tid; j < p.numTiles; j += GAE_THREADS) {
            p.status[j].P = 0.f; p.status[j].Q = 0.f; p.status[j].X = 0.f; p.status[j].flag = 0u;
        }
        if (tid == 0) { p.hdr->ticket = 0u; p.hdr->done = 0u; }
    }
}

struct GaePlan {
    int E, logE, L, pitch, numTiles, RW;
    uint32_t magicH;
    size_t smem;
};

GaePlan gae_plan(int64_t N, int64_t H) {
    GaePlan g{};
    const int64_t B = N * H;
    const int Ltarget = 4096, Lmax = 8192;
    if (N > 1 && H * 8 <= Lmax) {
        int E = 8;
        while ((int64_t)E * 2 * H <= Ltarget) E *= 2;   // largest power of two with E*H <= Ltarget, at least 8
        while (E > 8 && E / 2 >= N) E /= 2;             // do not over-allocate for tiny N
        g.E = E;
        g.logE = 0;
        while ((1 << g.logE) < E) ++g.logE;
        g.L = E * (int)H;
        g.pitch = (int)(H | 1);
        g.magicH = (uint32_t)(((1ull << 32) + (uint64_t)H - 1) / (uint64_t)H);
        g.numTiles = (int)pb_ceil_div(N, E);
        g.smem = (size_t)3 * E * g.pitch * sizeof(float);
    } else {
        g.E = 0;
        g.L = (int)(B < Ltarget ? (B > 0 ? B : 1) : Ltarget);
        g.pitch = 0;
        g.magicH = 0;
        g.numTiles = (int)pb_ceil_div(B, g.L);
        g.smem = (size_t)3 * g.L * sizeof(float);
    }
    const int R = (g.L + 31) / 32;
    const int Rw = (R + GAE_WARPS - 1) / GAE_WARPS;
    g.RW = Rw <= 16 ? 16 : 32;
    return g;
}

}  // namespace

extern "C" size_t pb_gae_workspace_bytes(int64_t num_envs, int64_t horizon) {
    if (num_envs <= 0 || horizon <= 0) return sizeof(GaeHeader);
    GaePlan g = gae_plan(num_envs, horizon);
    return sizeof(GaeHeader) + (size_t)g.numTiles * sizeof(GaeStatus);
}

extern "C" int pb_gae(const float* rewards, const float* values, const float* dones, float* advantages,
                      float* returns_sorted, int64_t num_envs, int64_t horizon, float gamma, float gae_lambda,
                      void* workspace, size_t workspace_bytes, void* stream) {
    PB_REQUIRE(num_envs >= 0 && horizon >= 0, PB_ERR_INVALID, "pb_gae: negative size");
    if (num_envs == 0 || horizon == 0) return PB_OK;
    PB_REQUIRE(rewards && values && dones && advantages, PB_ERR_INVALID, "pb_gae: null pointer");
    PB_REQUIRE(num_envs * horizon < (1ll << 40), PB_ERR_INVALID, "pb_gae: batch too large");
    GaePlan g = gae_plan(num_envs, horizon);
    const size_t need = sizeof(GaeHeader) + (size_t)g.numTiles * sizeof(GaeStatus);
    PB_REQUIRE(workspace && workspace_bytes >= need, PB_ERR_INVALID,
               "pb_gae: workspace too small (%zu < %zu)", workspace_bytes, need);
    GaeParams p{};
    p.r = rewards; p.v = values; p.d = dones; p.adv = advantages; p.ret = returns_sorted;
    p.N = num_envs; p.H = horizon; p.B = num_envs * horizon;
    p.gamma = gamma;
    p.gl = gamma * gae_lambda;  // float product, as `gamma * gae_lambda` in c_gae.pyx:29
    p.E = g.E; p.logE = g.logE; p.L = g.L; p.pitch = g.pitch; p.magicH = g.magicH; p.numTiles = g.numTiles;
    p.hdr = (GaeHeader*)workspace;
    p.status = (GaeStatus*)((char*)workspace + sizeof(GaeHeader));
    cudaStream_t s = (cudaStream_t)stream;
    if (g.RW == 16) {
        PB_CUDA(cudaFuncSetAttribute(k_gae<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k_gae<16><<<g.numTiles, GAE_THREADS, g.smem, s>>>(p);
    } else {
        PB_CUDA(cudaFuncSetAttribute(k_gae<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k_gae<32><<<g.numTiles, GAE_THREADS, g.smem, s>>>(p);
    }
    PB_LAUNCH_CHECK();
    return PB_OK;
}
