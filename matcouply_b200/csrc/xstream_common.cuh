// Pieces shared by the X-stream translation units (xstream.cu, xfused.cu): staging of the small factor matrix for
// the DMMA path, the fixed-order reduction of per-CTA partials and the TMA tensor-map encoding of the packed data.
#pragma once
#include "common.cuh"

namespace {

// pad/copy C (K x R, ld R) into Cp (Kp x LDC), zero filled, so that a K-chunk is one contiguous 16B-aligned bulk copy.
// mma_order: rows are permuted inside every group of 16 into the order the DMMA path consumes them — staged row
// 4*k4 + t holds source row 8*(t>>1) + 2*k4 + (t&1) (see the A-fragment mapping in xstream_y_kernel).
template <typename T>
__global__ void pad_c_kernel(const T* __restrict__ C, T* __restrict__ Cp, int K, int R, int Kp, int LDC, int mma_order) {
    const int n = Kp * LDC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int k = i / LDC;
        const int c = i - k * LDC;
        if (mma_order) {
            const int j = k & 15, k4 = j >> 2, t = j & 3;
            k = (k & ~15) + 8 * (t >> 1) + 2 * k4 + (t & 1);
        }
        Cp[i] = (k < K && c < R) ? C[(size_t)k * R + c] : T(0);
    }
}

// fixed-order reduction of the per-row-group partials: Z[i] = sum_g part[g][i]
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ part, T* __restrict__ out, int n, int groups) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s = T(0);
    for (int g = 0; g < groups; ++g) s += part[(size_t)g * n + i];
    out[i] = s;
}

// ---------------------------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_x_map(CUtensorMap* map, const void* X, long long N, int K, int ldx, int dtype, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    B2_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const size_t es = dtype == B2_F64 ? 8 : 4;
    B2_REQUIRE(((size_t)ldx * es) % 16 == 0, "X row stride (%d elements) must be a multiple of 16 bytes", ldx);
    B2_REQUIRE(((uintptr_t)X) % 16 == 0, "X must be 16-byte aligned");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)ldx * es};
    cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == B2_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return B2_OK;
}

}  // namespace
