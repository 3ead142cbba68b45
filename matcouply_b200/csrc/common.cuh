// Shared device/host helpers for the matcouply_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/matcouply_b200.h"

#define B2_MAX_RANK 32

// ---------------------------------------------------------------------------------------------------------
// host side: error plumbing (thread-local message, integer status codes; no exceptions cross the C ABI)
// ---------------------------------------------------------------------------------------------------------
void b2_set_error(const char* fmt, ...);

#define B2_CHECK_CUDA(expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            b2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
            return B2_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define B2_REQUIRE(cond, ...)                                                                           \
    do {                                                                                                \
        if (!(cond)) {                                                                                  \
            b2_set_error(__VA_ARGS__);                                                                  \
            return B2_ERR_INVALID;                                                                      \
        }                                                                                               \
    } while (0)

extern unsigned long long g_b2_launches;  // kernels launched through this library (bench.py's gpu_launches claim)
#define B2_LAUNCH_CHECK()                     \
    do {                                      \
        ++g_b2_launches;                      \
        B2_CHECK_CUDA(cudaGetLastError());    \
    } while (0)

int b2_num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device
int b2_option_value(int option);  // kernel-selection switches, see b2_set_option
// warp-per-slice polar steps selected by B2_OPT_POLAR_WARP inside b2_pf2_polar: 2 = Jacobi in registers
// (pf2_polar_reg.cu, ranks 2..B2_POLAR_REG_MAX_RANK, returns -1 otherwise), 1 = shared-memory warp kernel
// (pf2_polar_warp.cu), 0 = CTA-per-slice kernel (pf2_fused.cu)
#define B2_POLAR_REG_MAX_RANK 24
int b2_pf2_polar_reg(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                     void* num_part, void* Qstore, int warm, int dtype, cudaStream_t st);
int b2_pf2_polar_warp(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                      void* num_part, void* Qstore, int warm, int dtype, cudaStream_t st);

// dtype dispatch: calls `fn<float>(args...)` or `fn<double>(args...)`
#define B2_DISPATCH_DTYPE(dtype, ...)                                                                   \
    do {                                                                                                \
        if ((dtype) == B2_F64) {                                                                        \
            typedef double T;                                                                           \
            __VA_ARGS__;                                                                                \
        } else if ((dtype) == B2_F32) {                                                                 \
            typedef float T;                                                                            \
            __VA_ARGS__;                                                                                \
        } else {                                                                                        \
            b2_set_error("unknown dtype %d", (int)(dtype));                                             \
            return B2_ERR_INVALID;                                                                      \
        }                                                                                               \
    } while (0)

// rank dispatch: binds a compile-time bound RM >= R (per-row vectors stay in registers, loops fully unroll)
#define B2_DISPATCH_RANK(R, ...)                                                                        \
    do {                                                                                                \
        if ((R) <= 4) {                                                                                 \
            constexpr int RM = 4;                                                                       \
            __VA_ARGS__;                                                                                \
        } else if ((R) <= 8) {                                                                          \
            constexpr int RM = 8;                                                                       \
            __VA_ARGS__;                                                                                \
        } else if ((R) <= 12) {                                                                         \
            constexpr int RM = 12;                                                                      \
            __VA_ARGS__;                                                                                \
        } else if ((R) <= 16) {                                                                         \
            constexpr int RM = 16;                                                                      \
            __VA_ARGS__;                                                                                \
        } else if ((R) <= 20) {                                                                         \
            constexpr int RM = 20;                                                                      \
            __VA_ARGS__;                                                                                \
        } else if ((R) <= 24) {                                                                         \
            constexpr int RM = 24;                                                                      \
            __VA_ARGS__;                                                                                \
        } else {                                                                                        \
            constexpr int RM = 32;                                                                      \
            __VA_ARGS__;                                                                                \
        }                                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (shared::cta) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    // make barrier initialisation visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    // order generic-proxy shared-memory accesses before subsequent async-proxy (TMA) accesses
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug traps (-> CUDA error reported through the C ABI) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- TMA: 2D tiled tensor load, global -> shared, completion on an mbarrier ----
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- TMA: 1D bulk copy (16-byte aligned src/dst/size), global -> shared ----
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- explicit shared-space loads (LDS).  `volatile` keeps them ordered against the mbarrier wait / arrive asm ----
__device__ __forceinline__ double lds_f64(uint32_t saddr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ int2 lds_b64(uint32_t saddr) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ int4 lds_b128(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
template <typename T>
__device__ __forceinline__ T lds_elem(uint32_t saddr);
template <>
__device__ __forceinline__ double lds_elem<double>(uint32_t saddr) { return lds_f64(saddr); }
template <>
__device__ __forceinline__ float lds_elem<float>(uint32_t saddr) { return lds_f32(saddr); }

// ---- consumer-side release of a TMA ring stage ----------------------------------------------------------------
// The producer may overwrite the stage as soon as the `empty` barrier completes, so every lane's shared-memory reads of
// the stage must have RETURNED before the elected lane arrives.  __syncwarp() alone does not give that (ptxas hoists
// WARPSYNC above the loads and the arrive above the math that consumes them — observed as sporadically stale operand
// rows).  `dep_bits` must be computed from the results of the instructions that consumed every value loaded from the
// stage: the warp vote cannot execute before those results exist in all lanes, and its (always false) outcome feeds
// the barrier address, which pins the arrive behind it.  Combine the per-value bits with max(): the signed maximum of
// the high words equals the sentinel only if one of the values IS that NaN (an xor could hit it by chance).
__device__ __forceinline__ int dep_bits_of(double v) { return __double2hiint(v); }
__device__ __forceinline__ int dep_bits_of(float v) { return __float_as_int(v); }
__device__ __forceinline__ void stage_release(uint64_t* empty_bar, int lane, int dep_bits) {
    // 0x7ff7a5a5 / 0x7fb7a5a5: hi word of a NaN with a payload no arithmetic instruction produces
    const unsigned never = __any_sync(0xffffffffu, dep_bits == 0x7ff7a5a5) ? 1u : 0u;
    if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty_bar) + never) : "memory");
    }
}

// Byte offset of element (row, 16-byte chunk `c16`) inside a SWIZZLE_128B box whose rows are 128 bytes
// (box base 1024-byte aligned): chunk index is XORed with (row mod 8).
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t c16) { return row * 128u + ((c16 ^ (row & 7u)) << 4); }

// ---- fp64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col); SASS DMMA.8x8x4 ----
// lane = 4*g + t:  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- reductions ----
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all if `bcast`). `scratch` >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch, bool bcast = false) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        T u = lane < nw ? scratch[lane] : T(0);
        u = warp_sum(u);
        if (lane == 0) scratch[0] = u;
    }
    if (bcast) {
        __syncthreads();
        v = scratch[0];
    } else if (threadIdx.x == 0) {
        v = scratch[0];
    }
    return v;
}

template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
    typedef double2 type;
};
template <>
struct Vec2<float> {
    typedef float2 type;
};

#endif  // __CUDACC__
