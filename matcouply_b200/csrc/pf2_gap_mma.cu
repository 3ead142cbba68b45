// Feasibility-gap terms of the PARAFAC2 penalty from the deferred state (reference decomposition.py:406-415 with
// penalties.py:1287-1304), tensor-core formulation: per row  pd = V (W_g Delta)  as one DMMA row-matrix product in
// the accumulator layout of mma_tiles.cuh (18 DMMAs per 8 rows at R = 20 instead of 100 DFMAs + 40 shuffles + 60
// shared loads per 8 rows in the shuffle kernel of pf2_fused.cu), then  sum (x - pd)^2, sum x^2, sum |x|.
// One CTA per slice, 8 warps x 8 rows per pass, rows loaded straight from HBM with 16-byte loads.
#include "admm_common.cuh"
#include "mma_tiles.cuh"

namespace {

template <class PL, typename T>
__device__ __forceinline__ void load_row_global(const T* __restrict__ grow, int t, int R, bool valid,
                                                double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        v[b][0] = v[b][1] = 0.0;
        if (valid) {
            if (c0 >= 0 && c1 >= 0) {  // adjacent columns, 2*sizeof(T)-aligned because R*sizeof(T) % 16 == 0
                const typename Vec2<T>::type pr = *(const typename Vec2<T>::type*)(grow + c0);
                v[b][0] = (double)pr.x;
                v[b][1] = (double)pr.y;
            } else {
                if (c0 >= 0) v[b][0] = (double)grow[c0];
                if (c1 >= 0) v[b][1] = (double)grow[c1];
            }
        }
    }
}

template <typename T, int NBF, int HALF>
__global__ void __launch_bounds__(256)
pf2_gap_mma_kernel(const int64_t* __restrict__ row_off, int R, const T* __restrict__ V, const T* __restrict__ x,
                   const T* __restrict__ Wmat, const T* __restrict__ Delta, double* __restrict__ part) {
    using PL = PosLayout<NBF, HALF>;
    constexpr int NB = PL::NB;
    extern __shared__ double gm_smem[];
    __shared__ double scratch[32];
    double* Ts = gm_smem;                       // T_g = W_g Delta in position order
    double* wsm = Ts + PL::NPOS * PL::LDM;
    double* dsm = wsm + R * R;
    const int gsl = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RR = R * R;
    const long long r_begin = row_off[gsl], r_end = row_off[gsl + 1];
    for (int e = tid; e < RR; e += blockDim.x) {
        wsm[e] = (double)Wmat[(size_t)gsl * RR + e];
        dsm[e] = (double)Delta[e];
    }
    __syncthreads();
    for (int e = tid; e < PL::NPOS * PL::LDM; e += blockDim.x) {
        const int pr = e / PL::LDM, pc = e - pr * PL::LDM;
        double v = 0.0;
        if (pc < PL::NPOS) {
            const int r = PL::col_of(pr, R), c = PL::col_of(pc, R);
            if (r >= 0 && c >= 0)
                for (int k = 0; k < R; ++k) v = fma(wsm[r * R + k], dsm[k * R + c], v);
        }
        Ts[e] = v;
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    double d2 = 0.0, x2 = 0.0, ab = 0.0;
    for (long long row0 = r_begin + warp * 8; row0 < r_end; row0 += 64) {
        const long long row = row0 + g;
        const bool valid = row < r_end;
        const size_t off = (size_t)(valid ? row : r_begin) * R;
        double v[NB][2], xv[NB][2], pd[NB][2];
        load_row_global<PL, T>(V + off, t, R, valid, v);
        load_row_global<PL, T>(x + off, t, R, valid, xv);
        mma_rowmat<PL>(v, Ts, g, t, pd);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // rows past the end and padding positions hold zeros in v and xv, hence in pd: they add nothing
                const double d = xv[b][e] - pd[b][e];
                d2 = fma(d, d, d2);
                x2 = fma(xv[b][e], xv[b][e], x2);
                ab += fabs(xv[b][e]);
            }
    }
    d2 = block_sum(d2, scratch);
    x2 = block_sum(x2, scratch);
    ab = block_sum(ab, scratch);
    if (tid == 0) {
        part[(size_t)gsl * 3 + 0] = d2;
        part[(size_t)gsl * 3 + 1] = x2;
        part[(size_t)gsl * 3 + 2] = ab;
    }
}

template <typename T, int NBF, int HALF>
int launch_gap_mma(const int64_t* row_off, int n_groups, int R, const void* V, const void* x, const void* Wmat,
                   const void* Delta, double* part, cudaStream_t st) {
    using PL = PosLayout<NBF, HALF>;
    const size_t smem = (size_t)(PL::NPOS * PL::LDM + 2 * R * R) * sizeof(double);
    auto kern = pf2_gap_mma_kernel<T, NBF, HALF>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_groups, 256, smem, st>>>(row_off, R, (const T*)V, (const T*)x, (const T*)Wmat, (const T*)Delta, part);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // namespace

// Returns -1 when this formulation does not apply (row size not a multiple of 16 bytes): the caller then uses the
// shuffle kernel.  part: 3 * n_groups doubles.
int b2_pf2_gap_mma_try(const void* V, const void* x, const int64_t* row_off, int n_groups, int R, const void* Wmat,
                       const void* Delta, double* part, int dtype, cudaStream_t st) {
    const size_t es = dtype == B2_F64 ? 8 : 4;
    if (((size_t)R * es) % 16 != 0) return -1;
    if (((uintptr_t)V) % 16 != 0 || ((uintptr_t)x) % 16 != 0) return -1;
    const int nbf = R / 8, rem = R % 8;
    const int NBF = rem >= 5 ? nbf + 1 : nbf, HALF = (rem >= 1 && rem <= 4) ? 1 : 0;
#define B2_GAP_CASE(F, H)                                                                                      \
    if (NBF == F && HALF == H) {                                                                               \
        B2_DISPATCH_DTYPE(dtype, return (launch_gap_mma<T, F, H>(row_off, n_groups, R, V, x, Wmat, Delta, part, st))); \
    }
    B2_GAP_CASE(0, 1)
    B2_GAP_CASE(1, 0)
    B2_GAP_CASE(1, 1)
    B2_GAP_CASE(2, 0)
    B2_GAP_CASE(2, 1)
    B2_GAP_CASE(3, 0)
    B2_GAP_CASE(3, 1)
    B2_GAP_CASE(4, 0)
#undef B2_GAP_CASE
    return -1;
}
