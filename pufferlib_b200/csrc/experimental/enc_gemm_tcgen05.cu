// enc_gemm_tcgen05.cu -- EXPERIMENTAL, NOT PART OF libpuffer_b200.so, NOT YET RUN ON HARDWARE (round-2 groundwork).
//
// Stage A of the fused minibatch-update kernel planned in DESIGN.md ("Next"): the encoder forward of models.Default
//     hidden[M][128] = relu(x[M][128] @ W_enc[128][128]^T + b_enc)        (pufferlib/models.py:40-45)
// as a persistent, warp-specialised tcgen05 kernel -- the skeleton (TMA tensor-map loads with the 128-byte swizzle,
// UMMA shared-memory / instruction descriptors for kind::tf32, TMEM accumulator double buffering, tcgen05.ld epilogue)
// that the fused kernel extends with the heads / PPO-loss / dPre epilogue and the second UMMA (dW_enc += dPre^T x).
//
//   warp 0      TMA producer: W_enc once (4 boxes of 128 rows x 32 floats), then per 128-row tile 4 boxes of x
//   warp 1      TMEM allocation + MMA issue: per tile 16 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8)
//   warps 2..5  epilogue: tcgen05.ld 32x32b.x32 (thread = row, 32 columns at a time), bias + ReLU, row stores
//
// Shared memory (1024-byte aligned, SWIZZLE_128B K-major canonical layout: a K-block is [128 rows][32 floats = 128 B],
// 8-row groups of 1024 B, 16-byte chunk j of row r stored at chunk j ^ (r & 7)):
//   [0, 64K) W_enc: 4 K-blocks | [64K, 128K) x stage 0 | [128K, 192K) x stage 1 | barriers, TMEM base
// The same x tile read as an MN-major operand (N = input feature contiguous, K = row) is what the dW_enc UMMA of the
// fused kernel will consume: one staging of x serves both products.
//
// Descriptor encodings follow the PTX ISA tcgen05 tables (cross-checked against cute/arch/mma_sm100_desc.hpp):
//   smem descriptor: [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major; 1) | [32,46) SBO>>4 (1024 B between
//                    8-row groups -> 64) | [46,48) version = 1 | [61,64) layout: 2 = SWIZZLE_128B
//   instr descriptor: [4,6) D fmt F32 = 1 | [7,10) A fmt TF32 = 2 | [10,13) B fmt TF32 = 2 | bit 15/16 A/B major (0 = K)
//                     | [17,23) N>>3 | [24,29) M>>4
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../tma.cuh"

namespace {

char g_err[512] = "";
void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

constexpr int TILE_M = 128, HID = 128, KDIM = 128;
constexpr int KBLK = 32;                                  // floats per 128-byte swizzle row
constexpr int KBLK_BYTES = TILE_M * KBLK * 4;             // 16 KB: one [128][32] K-block
constexpr int TILE_BYTES = 4 * KBLK_BYTES;                // 64 KB
constexpr int SMEM_W = 0, SMEM_X0 = TILE_BYTES, SMEM_BAR = 3 * TILE_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 128;
constexpr int TMEM_COLS = 256;                            // 2 accumulator stages x 128 fp32 columns
constexpr int THREADS = 192;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((HID >> 3) << 17) | ((TILE_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all prior MMAs of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
k_enc_gemm_tcgen05(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const float* __restrict__ b_enc, float* __restrict__ hidden, int64_t m, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* full = bars;            // [2] x stage filled by TMA
    uint64_t* empty = bars + 2;       // [2] x stage consumed by the MMAs
    uint64_t* tfull = bars + 4;       // [2] accumulator stage complete
    uint64_t* tempty = bars + 6;      // [2] accumulator stage drained by the 4 epilogue warps
    uint64_t* wfull = bars + 8;       // W_enc resident
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 10);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        mbar_init(wfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) {   // TMEM allocation is warp-collective; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(wfull, TILE_BYTES);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + SMEM_W + kb * KBLK_BYTES, &map_w, kb * KBLK, 0, wfull);
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], TILE_BYTES);
                uint8_t* dst = smem + SMEM_X0 + s * TILE_BYTES;
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + kb * KBLK_BYTES, &map_x, kb * KBLK, tile * TILE_M, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            mbar_wait(wfull, 0);
            const uint32_t w_addr = smem_u32(smem + SMEM_W);
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                mbar_wait(&tempty[s], ph ^ 1);      // epilogue has drained accumulator stage s
                mbar_wait(&full[s], ph);            // x tile landed
                tc_fence_after();
                const uint32_t x_addr = smem_u32(smem + SMEM_X0 + s * TILE_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * HID);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 4 x K=8 (32 bytes) inside one 128-byte swizzle row
                        const uint64_t a = smem_desc_sw128(x_addr + kb * KBLK_BYTES + k * 32);
                        const uint64_t b = smem_desc_sw128(w_addr + kb * KBLK_BYTES + k * 32);
                        umma_tf32(d_tmem, a, b, IDESC_TF32_128x128, (kb | k) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);             // x stage s may be overwritten once these MMAs retire
                umma_commit(&tfull[s]);             // ... and the accumulator stage is complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp index % 4 (hardware rule for tcgen05.ld) =====
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            mbar_wait(&tfull[s], ph);
            tc_fence_after();
            const int64_t row = (int64_t)tile * TILE_M + 32 * q + lane;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(s * HID + 32 * c), v);
                if (row < m) {
                    float4* dst = reinterpret_cast<float4*>(hidden + row * HID + 32 * c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = *reinterpret_cast<const float4*>(b_enc + 32 * c + 4 * j);
                        float4 o;
                        o.x = fmaxf(__uint_as_float(v[4 * j + 0]) + b.x, 0.f);
                        o.y = fmaxf(__uint_as_float(v[4 * j + 1]) + b.y, 0.f);
                        o.z = fmaxf(__uint_as_float(v[4 * j + 2]) + b.z, 0.f);
                        o.w = fmaxf(__uint_as_float(v[4 * j + 3]) + b.w, 0.f);
                        dst[j] = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[s]);
        }
    }
    // ---- teardown: everyone is done with TMEM before the allocating warp frees it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn fn, CUtensorMap* map, const float* base, int64_t rows, int64_t row_stride_floats) {
    const cuuint64_t dims[2] = {(cuuint64_t)KDIM, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_stride_floats * 4};
    const cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_err("cuTensorMapEncodeTiled failed: %d", (int)r);
        return -1;
    }
    return 0;
}

}  // namespace

extern "C" const char* pbx_last_error(void) { return g_err; }

// hidden[m][128] = relu(x[m][128 (row stride ldx floats)] @ w_enc[128][128]^T + b_enc); TF32 products (operand bits are
// consumed as TF32: low 13 mantissa bits ignored), fp32 accumulate.  Returns 0 or a negative error (pbx_last_error()).
extern "C" int pbx_enc_gemm_tf32(const float* x, int64_t ldx, int64_t m, const float* w_enc, const float* b_enc,
                                 float* hidden, void* stream) {
    if (!x || !w_enc || !b_enc || !hidden || m < 1 || ldx < KDIM || ldx % 4 != 0 || ((uintptr_t)x & 15) ||
        ((uintptr_t)w_enc & 15) || ((uintptr_t)hidden & 15) || ((uintptr_t)b_enc & 15)) {
        set_err("pbx_enc_gemm_tf32: bad arguments");
        return -1;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        set_err("cuTensorMapEncodeTiled entry point not available");
        return -2;
    }
    alignas(64) CUtensorMap map_x, map_w;
    if (make_map((EncodeTiledFn)fn, &map_x, x, m, ldx) || make_map((EncodeTiledFn)fn, &map_w, w_enc, HID, KDIM)) return -3;
    const int n_tiles = (int)((m + TILE_M - 1) / TILE_M);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < sms ? n_tiles : sms;
    cudaError_t e = cudaFuncSetAttribute(k_enc_gemm_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e == cudaSuccess) {
        k_enc_gemm_tcgen05<<<grid, THREADS, SMEM_TOTAL, (cudaStream_t)stream>>>(map_x, map_w, b_enc, hidden, m, n_tiles);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        set_err("pbx_enc_gemm_tf32: %s", cudaGetErrorString(e));
        return -4;
    }
    return 0;
}
