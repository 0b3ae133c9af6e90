// umma_probe.cu -- hardware probe for tcgen05.mma shared-memory / instruction descriptor encodings (test tool, not
// part of libpuffer_b200.so).  The host builds the exact shared-memory IMAGE of both operands (so every layout
// hypothesis -- K-major / MN-major, swizzle, LBO / SBO -- is expressed in Python), passes descriptor templates, and
// gets the raw TMEM accumulator back.  tests/experimental/check_umma_probe.py compares it with a matmul.
//
// One CTA, 128 threads.  a_img / b_img are copied verbatim to shared memory at offsets 0 and 96 KiB (1024-aligned)
// with ordinary stores (generic proxy) followed by fence.proxy.async -- the same hand-off the fused update kernel uses
// for operands produced by CUDA cores.  Thread 0 then issues n_mma MMAs
//     D[tmem] (+)= A(desc_a + i*a_step) * B(desc_b + i*b_step)
// commits to an mbarrier, and all four warps dump their TMEM lane quadrant (lane = row of D) to d_out[128][ncols].
#include <cuda_runtime.h>
#include <stdint.h>

#include "../tma.cuh"

namespace {

constexpr int PROBE_B_OFF = 96 * 1024;
constexpr int PROBE_SMEM = 192 * 1024 + 64;

__global__ void __launch_bounds__(128, 1)
k_umma_probe(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img, uint32_t a_bytes, uint32_t b_bytes,
             uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int n_mma, int inner, uint32_t a_step, uint32_t b_step,
             uint32_t a_step2, uint32_t b_step2, float* __restrict__ d_out, int ncols, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192 * 1024 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = a_img[i];
    for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem + PROBE_B_OFF)[i] = b_img[i];
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + PROBE_B_OFF);
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t ao = (i % inner) * a_step + (i / inner) * a_step2, bo = (i % inner) * b_step + (i / inner) * b_step2;
            const uint64_t da = a_desc | (uint64_t)(((a0 + ao) & 0x3FFFF) >> 4);
            const uint64_t db = b_desc | (uint64_t)(((b0 + bo) & 0x3FFFF) >> 4);
            const uint32_t acc = i ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                "}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                     : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < ncols; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) d_out[(32 * warp + lane) * ncols + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}


// Two-stage probe: stage 1 = n1 SS MMAs (A image at 0, B image at 64 KiB) into TMEM columns [0,128); stage 2 = n2 TS MMAs
// whose A operand is stage 1's accumulator (TMEM columns a2_col0 + i * a2_colstep, lanes = M) and whose B operand is a third
// image at 128 KiB, accumulating into columns [128, 256).  Both accumulators are dumped: d_out[128][256].
__global__ void __launch_bounds__(128, 1)
k_umma_probe2(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img, const uint4* __restrict__ c_img,
              uint32_t a_bytes, uint32_t b_bytes, uint32_t c_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t idesc1, int n1,
              int inner1, uint32_t a_step, uint32_t b_step, uint32_t a_step2, uint32_t b_step2, uint32_t a_off0,
              uint64_t c_desc, uint32_t idesc2, int n2, int inner2, uint32_t c_step, uint32_t c_step2, uint32_t a2_col0,
              uint32_t a2_colstep, float* __restrict__ d_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192 * 1024 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = a_img[i];
    for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem + 64 * 1024)[i] = b_img[i];
    for (uint32_t i = threadIdx.x; i < c_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem + 128 * 1024)[i] = c_img[i];
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem) + a_off0, b0 = smem_u32(smem + 64 * 1024), c0 = smem_u32(smem + 128 * 1024);
        for (int i = 0; i < n1; ++i) {
            const uint32_t ao = (i % inner1) * a_step + (i / inner1) * a_step2, bo = (i % inner1) * b_step + (i / inner1) * b_step2;
            const uint64_t da = a_desc | (uint64_t)(((a0 + ao) & 0x3FFFF) >> 4);
            const uint64_t db = b_desc | (uint64_t)(((b0 + bo) & 0x3FFFF) >> 4);
            const uint32_t acc = i ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                "}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc1), "r"(acc)
                : "memory");
        }
        // same issuing thread: the tensor pipe executes the MMAs in order, stage 2 reads what stage 1 wrote
        for (int i = 0; i < n2; ++i) {
            const uint32_t co = (i % inner2) * c_step + (i / inner2) * c_step2;
            const uint64_t dc = c_desc | (uint64_t)(((c0 + co) & 0x3FFFF) >> 4);
            const uint32_t acc = i ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "}\n" ::"r"(tmem + 128u),
                "r"(tmem + a2_col0 + (uint32_t)i * a2_colstep), "l"(dc), "r"(idesc2), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                     : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 256; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) d_out[(32 * warp + lane) * 256 + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

}  // namespace

// a_desc / b_desc: descriptor templates WITHOUT the start address (bits [0,14) zero).  Returns 0 or a cudaError code.
extern "C" int pbx_umma_probe(const void* a_img, const void* b_img, uint32_t a_bytes, uint32_t b_bytes, uint64_t a_desc,
                              uint64_t b_desc, uint32_t idesc, int n_mma, int inner, uint32_t a_step, uint32_t b_step,
                              uint32_t a_step2, uint32_t b_step2, float* d_out, int ncols, void* stream) {
    if (a_bytes > PROBE_B_OFF || b_bytes > PROBE_B_OFF || (a_bytes & 15) || (b_bytes & 15) || ncols < 8 || ncols > 256 ||
        (ncols & 7) || inner < 1)
        return -1;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < ncols) tmem_cols <<= 1;
    cudaError_t e = cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM);
    if (e != cudaSuccess) return (int)e;
    k_umma_probe<<<1, 128, PROBE_SMEM, (cudaStream_t)stream>>>((const uint4*)a_img, (const uint4*)b_img, a_bytes, b_bytes,
                                                              a_desc, b_desc, idesc, n_mma, inner, a_step, b_step, a_step2,
                                                              b_step2, d_out, ncols, tmem_cols);
    return (int)cudaGetLastError();
}

extern "C" int pbx_umma_probe2(const void* a_img, const void* b_img, const void* c_img, uint32_t a_bytes, uint32_t b_bytes,
                               uint32_t c_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t idesc1, int n1, int inner1,
                               uint32_t a_step, uint32_t b_step, uint32_t a_step2, uint32_t b_step2, uint32_t a_off0,
                               uint64_t c_desc, uint32_t idesc2, int n2, int inner2, uint32_t c_step, uint32_t c_step2,
                               uint32_t a2_col0, uint32_t a2_colstep, float* d_out, void* stream) {
    if (a_bytes > 65536 || b_bytes > 65536 || c_bytes > 65536 || ((a_bytes | b_bytes | c_bytes) & 15) || inner1 < 1 || inner2 < 1)
        return -1;
    cudaError_t e = cudaFuncSetAttribute(k_umma_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM);
    if (e != cudaSuccess) return (int)e;
    k_umma_probe2<<<1, 128, PROBE_SMEM, (cudaStream_t)stream>>>(
        (const uint4*)a_img, (const uint4*)b_img, (const uint4*)c_img, a_bytes, b_bytes, c_bytes, a_desc, b_desc, idesc1, n1,
        inner1, a_step, b_step, a_step2, b_step2, a_off0, c_desc, idesc2, n2, inner2, c_step, c_step2, a2_col0, a2_colstep, d_out);
    return (int)cudaGetLastError();
}
