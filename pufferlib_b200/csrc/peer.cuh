// peer.cuh -- device side of the NVLink peer-memory all-reduce (see peer.cu); included by the optimizer kernel.
#pragma once
#include <string.h>

#include "pb_common.cuh"

constexpr int PB_PEER_HEADER_BYTES = 1024;   // 8 flag lines of 128 B

#ifdef __CUDACC__
__device__ __forceinline__ void pb_st_release_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t pb_ld_acquire_sys_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float pb_ld_relaxed_sys_f32(const float* p) {   // peer data: never from a stale L1 line
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float4 pb_ld_relaxed_sys_f32x4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// One CTA (any size that is a multiple of 32): flat[0..n) <- sum over ranks, in rank order (bit-identical everywhere).
__device__ __forceinline__ void pb_peer_allreduce_sum(const pb_peer_comm& c, float* flat, int64_t n) {
    if (c.world <= 1) return;
    __shared__ uint64_t s_epoch;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_epoch = *c.epoch + 1;
    __syncthreads();
    const uint64_t e = s_epoch;
    const int64_t slot_off = PB_PEER_HEADER_BYTES / 4 + (int64_t)(e & 1) * c.capacity;
    float* mine = reinterpret_cast<float*>(c.base[c.rank]) + slot_off;
    // 128-bit accesses where the layout allows (capacity % 4 == 0 keeps both slots 16-byte aligned): at 68 KB per rank the
    // exchange is latency-bound, so every thread should have all its peer loads in flight at once
    const bool vec = (c.capacity & 3) == 0 && (reinterpret_cast<uintptr_t>(flat) & 15) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += nt) reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(flat)[i];
    for (int64_t i = 4 * n4 + tid; i < n; i += nt) mine[i] = flat[i];
    __threadfence_system();
    __syncthreads();
    if (tid < c.world) {   // raise my flag in every rank's buffer (mine included)
        uint64_t* flag = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(c.base[tid]) + 128 * c.rank);
        pb_st_release_sys_u64(flag, e);
        // ... and wait for rank `tid`'s flag in my buffer
        const uint64_t* theirs = reinterpret_cast<const uint64_t*>(reinterpret_cast<const char*>(c.base[c.rank]) + 128 * tid);
        const long long t0 = clock64();
        while (pb_ld_acquire_sys_u64(theirs) < e) {
            if (clock64() - t0 > 20000000000ll) __trap();   // ~10 s: a peer died; do not hang the box
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int64_t i = tid; i < n4; i += nt) {
        float4 v[PB_PEER_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < PB_PEER_MAX_RANKS; ++r)
            if (r < c.world) v[r] = pb_ld_relaxed_sys_f32x4(reinterpret_cast<const float*>(c.base[r]) + slot_off + 4 * i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < PB_PEER_MAX_RANKS; ++r)      // rank order: the same bits on every rank
            if (r < c.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        reinterpret_cast<float4*>(flat)[i] = s;
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nt) {
        float s = 0.f;
        for (int r = 0; r < c.world; ++r) s += pb_ld_relaxed_sys_f32(reinterpret_cast<const float*>(c.base[r]) + slot_off + i);
        flat[i] = s;
    }
    __syncthreads();
    if (tid == 0) *c.epoch = e;
}

constexpr int PB_PEER_SLICES = 16;           // flags of source rank r, slice b at byte 128 r + 8 b of the header

// Multi-CTA form: CTA b (of PB_PEER_SLICES) sums slice b over all ranks.  Slices are independent (own flag per rank and
// slice), so there is no grid-wide barrier; the epoch counter is NOT advanced here (every CTA reads it at its start): the
// kernel that follows in the stream does it (pb_clip_adam_parts).  sumsq[b] = sum of squares of the summed slice.
__device__ __forceinline__ void pb_peer_allreduce_slice(const pb_peer_comm& c, float* flat, int64_t n, double* sumsq) {
    __shared__ uint64_t s_epoch;
    __shared__ double s_sq[32];
    const int tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x;
    if (tid == 0) s_epoch = *c.epoch + 1;
    __syncthreads();
    const uint64_t e = s_epoch;
    const int64_t chunk = ((n + PB_PEER_SLICES - 1) / PB_PEER_SLICES + 3) & ~(int64_t)3;
    const int64_t lo = (int64_t)b * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const int64_t slot_off = PB_PEER_HEADER_BYTES / 4 + (int64_t)(e & 1) * c.capacity;
    float* mine = reinterpret_cast<float*>(c.base[c.rank]) + slot_off;
    // 128-bit accesses where the layout allows (slices start at multiples of 4 floats): one peer load per thread and rank,
    // all in flight at once -- the exchange is a handful of NVLink round trips, not bandwidth
    const bool vec = (c.capacity & 3) == 0 && (reinterpret_cast<uintptr_t>(flat) & 15) == 0;
    const int64_t hi4 = vec ? lo + ((hi - lo) & ~(int64_t)3) : lo;          // [lo, hi4) in float4 steps, [hi4, hi) scalar
    for (int64_t i = lo + 4 * tid; i < hi4; i += 4 * nt)
        *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(flat + i);
    for (int64_t i = hi4 + tid; i < hi; i += nt) mine[i] = flat[i];
    __syncthreads();
    if (tid < c.world) {
        __threadfence_system();      // (cumulative: the CTA barrier ordered every thread's slot writes before this thread)
        uint64_t* flag = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(c.base[tid]) + 128 * c.rank + 8 * b);
        pb_st_release_sys_u64(flag, e);
        const uint64_t* theirs = reinterpret_cast<const uint64_t*>(reinterpret_cast<const char*>(c.base[c.rank]) + 128 * tid + 8 * b);
        const long long t0 = clock64();
        while (pb_ld_acquire_sys_u64(theirs) < e) {
            if (clock64() - t0 > 20000000000ll) __trap();
        }
    }
    __syncthreads();
    double sq = 0.0;
    for (int64_t i = lo + 4 * tid; i < hi4; i += 4 * nt) {
        float4 v[PB_PEER_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < PB_PEER_MAX_RANKS; ++r)
            if (r < c.world) v[r] = pb_ld_relaxed_sys_f32x4(reinterpret_cast<const float*>(c.base[r]) + slot_off + i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < PB_PEER_MAX_RANKS; ++r)      // rank order: the same bits on every rank
            if (r < c.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        *reinterpret_cast<float4*>(flat + i) = s;
        sq += ((double)s.x * (double)s.x + (double)s.y * (double)s.y) + ((double)s.z * (double)s.z + (double)s.w * (double)s.w);
    }
    for (int64_t i = hi4 + tid; i < hi; i += nt) {
        float s = 0.f;
        for (int r = 0; r < c.world; ++r) s += pb_ld_relaxed_sys_f32(reinterpret_cast<const float*>(c.base[r]) + slot_off + i);
        flat[i] = s;
        sq += (double)s * (double)s;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((tid & 31) == 0) s_sq[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < (nt + 31) / 32; ++w) tot += s_sq[w];
        sumsq[b] = tot;
    }
}
#endif
