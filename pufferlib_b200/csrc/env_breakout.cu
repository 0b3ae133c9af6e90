// env_breakout.cu -- N Breakout instances (oracle/SPEC.md §Breakout) behind the reference vectoriser semantics.
//
// Dynamics are the builder's spec (the reference has no first-party breakout: SURVEY.md §0); conventions are the
// reference's: reset-on-next-send (vector.py:147-151), reset rows r=0/term=False/mask=True (emulation.py:187-192),
// EpisodeStats on the terminal row (postprocess.py:22-54).  Bit-exact against oracle/csrc/envs.c.
//
// Layout: per-env state is 28 B of SoA in HBM (two packed words, a uint4 brick bitmap, the draw counter),
// reloaded each step because the policy forward sits between steps.  One lane owns one env for the integer
// physics (EPW = 8 envs per warp, so N = 16384 still fills the chip with 2048 warps); the envs of a warp then emit
// their 512-byte observation rows cooperatively -- per row each lane builds one float4 (lane 0-1: the 8 header
// scalars, lanes 2-31: 4 bricks each from the shuffled bitmap) so every row is ONE fully coalesced 512 B warp store
// (4 x 128 B lines).  Reward / flag / done rows are [N]-contiguous.
#include "env_common.cuh"

namespace {

struct BreakoutState {
    uint32_t* s0;   // px(8) | lives(3)<<8 | in_play<<11 | wait(5)<<12 | (vx+3)(3)<<17 | (vy+2)(3)<<20
    uint32_t* s1;   // bx(8) | by(8)<<8 | tick(16)<<16
    uint4* bricks;  // 120 alive bits, brick i = row*20+col -> word i>>5, bit i&31
    uint32_t* ctr;  // RNG draw counter
    uint64_t seed;  // base seed + env_index_offset (env e uses seed + e)
    int max_ticks;
};

struct BkOut {
    float* obs;
    int64_t stride_f;
    float* rewards;
    uint8_t* terminals;
    uint8_t* truncations;
    uint8_t* masks;
    float* dones_f32;
    bool write_const;
};

__device__ __forceinline__ uint32_t bk_draw(uint64_t seed_e, uint32_t& ctr) {
    const uint32_t r = pb_mix32(seed_e * 0x9E3779B97F4A7C15ull + (uint64_t)ctr * 0xD1B54A32D192ED03ull);
    ctr += 1;
    return r;
}

// MODE 0: async_reset rows for every env;  MODE 1: vectoriser send (reset-or-step)
constexpr int EPW = 8;   // envs per warp: lanes 0..7 own one env each, all 32 lanes write the rows

template <int MODE>
__global__ void __launch_bounds__(128) k_breakout(BreakoutState st, int n, const int64_t* __restrict__ actions,
                                                 uint8_t* done, BkOut out, EpisodeAcc acc) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int e_base = warp * EPW;
    const int e = e_base + lane;
    const bool active = lane < EPW && e < n;
    int px = 68, lives = 5, in_play = 0, wait = 0, vx = 0, vy = 0, bx = 79, by = 188, tick = 0;
    uint4 bricks = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00ffffffu);
    uint32_t ctr = 0;
    int reward = 0;
    bool terminal = false, reset_row = true;
    float score = 0.f;
    if (active) {
        const uint64_t seed_e = st.seed + (uint64_t)e;
        bool do_reset = true;
        uint32_t a0 = 0, a1 = 0;
        uint4 bricks_in = bricks;
        int a = 0;
        if (MODE == 1) {   // all state loads are issued together (independent of `done`): one latency, not two
            ctr = st.ctr[e];
            a0 = st.s0[e]; a1 = st.s1[e];
            bricks_in = st.bricks[e];
            a = (int)actions[e];
            do_reset = done[e] != 0;
        }
        if (!do_reset) {
            reset_row = false;
            px = a0 & 0xff; lives = (a0 >> 8) & 7; in_play = (a0 >> 11) & 1; wait = (a0 >> 12) & 31;
            vx = (int)((a0 >> 17) & 7) - 3; vy = (int)((a0 >> 20) & 7) - 2;
            bx = a1 & 0xff; by = (a1 >> 8) & 0xff; tick = a1 >> 16;
            bricks = bricks_in;
            a = a < 0 ? 0 : (a > 3 ? 3 : a);
            if (a == 2) px = min(px + 4, 136);
            if (a == 3) px = max(px - 4, 0);
            if (!in_play) {
                wait += 1;
                bx = px + 11; by = 188;
                if (a == 1 || wait >= 16) {
                    in_play = 1; vy = -2;
                    const int k = (int)(bk_draw(seed_e, ctr) & 3u);
                    vx = k < 2 ? k - 2 : k - 1;   // {-2,-1,1,2}
                }
            } else {
                bx += vx; by += vy;
                if (bx < 0) { bx = -bx; vx = -vx; }
                if (bx > 158) { bx = 316 - bx; vx = -vx; }
                if (by < 0) { by = -by; vy = -vy; }
                const int cx = bx + 1, cy = by + 1;
                if (cy >= 30 && cy < 66) {
                    const int row = (cy - 30) / 6, col = cx >> 3;
                    const int i = row * 20 + col;
                    uint32_t* w = (i < 32) ? &bricks.x : (i < 64) ? &bricks.y : (i < 96) ? &bricks.z : &bricks.w;
                    const uint32_t bit = 1u << (i & 31);
                    if (*w & bit) {
                        *w &= ~bit;
                        reward += row < 2 ? 7 : (row < 4 ? 4 : 1);
                        vy = -vy;
                    }
                }
                if (vy > 0 && by >= 188 && by <= 192 && bx + 2 > px && bx < px + 24) {
                    int off = bx + 1 - px;
                    off = off < 0 ? 0 : (off > 23 ? 23 : off);
                    const int seg = off >> 2;
                    vy = -2; by = 188;
                    vx = seg < 3 ? seg - 3 : seg - 2;   // {-3,-2,-1,1,2,3}
                } else if (by >= 198) {
                    lives -= 1; in_play = 0; wait = 0;
                    bx = px + 11; by = 188; vx = 0; vy = 0;
                }
            }
            tick += 1;
            const int left = __popc(bricks.x) + __popc(bricks.y) + __popc(bricks.z) + __popc(bricks.w);
            terminal = lives == 0 || left == 0 || tick >= st.max_ticks;
            score = (float)(120 - left) / 120.0f;
        }
        st.s0[e] = (uint32_t)px | ((uint32_t)lives << 8) | ((uint32_t)in_play << 11) | ((uint32_t)wait << 12) |
                   ((uint32_t)(vx + 3) << 17) | ((uint32_t)(vy + 2) << 20);
        st.s1[e] = (uint32_t)bx | ((uint32_t)by << 8) | ((uint32_t)tick << 16);
        st.bricks[e] = bricks;
        st.ctr[e] = ctr;
        done[e] = terminal ? 1 : 0;
        out.rewards[e] = (float)reward;
        out.terminals[e] = terminal ? 1 : 0;
        if (out.write_const) out.truncations[e] = 0;
        if (out.write_const) out.masks[e] = 1;
        if (out.dones_f32) out.dones_f32[e] = terminal ? 1.f : 0.f;
    }
    episode_update(acc, e, active, reset_row, (double)reward, terminal, score);

    // ---- observation rows: 32 envs per warp, one coalesced 512 B store per env
    const int left_mine = __popc(bricks.x) + __popc(bricks.y) + __popc(bricks.z) + __popc(bricks.w);
    const int word_sel = (lane >= 2) ? ((4 * lane - 8) >> 5) : 0;   // which bitmap word this lane decodes
    const int bit0 = (4 * lane - 8) & 31;
#pragma unroll
    for (int j = 0; j < EPW; ++j) {
        if (e_base + j >= n) break;
        const int jpx = __shfl_sync(0xffffffffu, px, j), jbx = __shfl_sync(0xffffffffu, bx, j);
        const int jby = __shfl_sync(0xffffffffu, by, j), jvx = __shfl_sync(0xffffffffu, vx, j);
        const int jvy = __shfl_sync(0xffffffffu, vy, j), jlives = __shfl_sync(0xffffffffu, lives, j);
        const int jplay = __shfl_sync(0xffffffffu, in_play, j), jleft = __shfl_sync(0xffffffffu, left_mine, j);
        const uint32_t w0 = __shfl_sync(0xffffffffu, bricks.x, j), w1 = __shfl_sync(0xffffffffu, bricks.y, j);
        const uint32_t w2 = __shfl_sync(0xffffffffu, bricks.z, j), w3 = __shfl_sync(0xffffffffu, bricks.w, j);
        float4 v;
        if (lane == 0) {
            v = make_float4((float)jpx * (1.f / 256.f), (float)jbx * (1.f / 256.f), (float)jby * (1.f / 256.f),
                            (float)jvx * 0.25f);
        } else if (lane == 1) {
            v = make_float4((float)jvy * 0.25f, (float)jlives * 0.125f, (float)jplay, (float)jleft * (1.f / 128.f));
        } else {
            const uint32_t w = word_sel == 0 ? w0 : (word_sel == 1 ? w1 : (word_sel == 2 ? w2 : w3));
            const uint32_t b = w >> bit0;
            v = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
        }
        float4* row = reinterpret_cast<float4*>(out.obs + (int64_t)(e_base + j) * out.stride_f);
        row[lane] = v;
    }
}

int breakout_launch(pb_env* env, int mode, const int64_t* actions, const pb_env_out* out, cudaStream_t s) {
    BreakoutState* st = (BreakoutState*)env->kind;
    const int n = env->cfg.num_envs;
    PB_REQUIRE(out->obs_stride % 16 == 0 && ((uintptr_t)out->obs & 15) == 0, PB_ERR_INVALID,
               "breakout: obs pointer/stride must be 16-byte aligned");
    BkOut o{(float*)out->obs, out->obs_stride / 4, out->rewards, out->terminals, out->truncations, out->masks,
            out->dones_f32,
            env->write_const};
    const int blocks = (int)pb_ceil_div(n, (128 / 32) * EPW);
    if (mode == 0) k_breakout<0><<<blocks, 128, 0, s>>>(*st, n, actions, env->d_done, o, pb_episode_acc(env));
    else k_breakout<1><<<blocks, 128, 0, s>>>(*st, n, actions, env->d_done, o, pb_episode_acc(env));
    PB_LAUNCH_CHECK();
    return PB_OK;
}

int breakout_reset(pb_env* env, uint64_t seed, const pb_env_out* out, cudaStream_t s) {
    BreakoutState* st = (BreakoutState*)env->kind;
    st->seed = seed + (uint64_t)env->cfg.env_index_offset;
    return breakout_launch(env, 0, nullptr, out, s);
}

int breakout_step(pb_env* env, const int64_t* actions, const pb_env_out* out, cudaStream_t s) {
    return breakout_launch(env, 1, actions, out, s);
}

void breakout_destroy(pb_env* env) {
    BreakoutState* st = (BreakoutState*)env->kind;
    if (!st) return;
    cudaFree(st->s0); cudaFree(st->s1); cudaFree(st->bricks); cudaFree(st->ctr);
    delete st;
    env->kind = nullptr;
}

const pb_env_vtable BREAKOUT_VT = {breakout_reset, breakout_step, breakout_destroy};

}  // namespace

int pb_breakout_create(pb_env* env) {
    BreakoutState* st = new BreakoutState();
    env->kind = st;
    env->vt = &BREAKOUT_VT;
    st->max_ticks = env->cfg.iparam[0] > 0 ? env->cfg.iparam[0] : 4096;
    PB_REQUIRE(st->max_ticks <= 65535, PB_ERR_INVALID, "breakout: max_ticks must be <= 65535");
    const size_t n = (size_t)env->cfg.num_envs;
    PB_CUDA(cudaMalloc(&st->s0, n * 4));
    PB_CUDA(cudaMalloc(&st->s1, n * 4));
    PB_CUDA(cudaMalloc(&st->bricks, n * 16));
    PB_CUDA(cudaMalloc(&st->ctr, n * 4));
    PB_CUDA(cudaMemset(st->s0, 0, n * 4));
    PB_CUDA(cudaMemset(st->s1, 0, n * 4));
    PB_CUDA(cudaMemset(st->bricks, 0, n * 16));
    PB_CUDA(cudaMemset(st->ctr, 0, n * 4));
    env->info.obs_dtype = PB_DTYPE_F32;
    env->info.obs_ndim = 1;
    env->info.obs_shape[0] = 128;
    env->info.obs_bytes = 512;
    env->info.num_actions = 4;
    env->info.obs_low = -1.f;
    env->info.obs_high = 1.f;
    return PB_OK;
}
