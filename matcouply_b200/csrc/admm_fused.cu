// Fully fused ADMM sub-solver for modes whose penalties are all row-local (NonNegativity, Box, L1 — or none):
// the whole inner loop of admm_update_B / _C / _A (decomposition.py:259-289, 325-342, 176-217) runs in registers,
// one pass over the state: read rhs, aux, dual once; iterate  x = (rho*sum(aux-dual) + rhs) Minv ; aux = prox(x+dual);
// dual = x + dual - aux  `n_inner` times; write x, aux, dual once (and W = x o a for the following X^T W pass).
//
// Thread mapping: 4 lanes per row, each lane owns CPL = ceil(R/4) consecutive columns, so a warp touches 8
// consecutive rows = one contiguous, fully coalesced 8*R-element segment per array.  The R x R solve needs the whole
// vector s: it is exchanged between the 4 lanes of a row with warp shuffles (no shared memory, no block barrier).
#include "admm_common.cuh"

namespace {

template <typename T, int CPL, int NP>
__global__ void __launch_bounds__(256)
admm_local_kernel(long long n, int R, const T* __restrict__ rhs, const T* __restrict__ rhs_scale, int group_mode,
                  const int32_t* __restrict__ gor, const T* __restrict__ rho, const T* __restrict__ Minv, PenArgs pa,
                  int n_inner, T* __restrict__ x, T* __restrict__ w_out, int ldw) {
    const int lane = threadIdx.x & 31, l4 = lane & 3;
    const long long rowid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool valid = rowid < n;
    const long long row = valid ? rowid : n - 1;  // clamp: all lanes stay in the shuffles, stores are masked
    const int g = group_mode == B2_GROUP_SINGLE ? 0 : (group_mode == B2_GROUP_INDEXED ? gor[row] : (int)row);
    const T rg = rho[g];
    const int c0 = l4 * CPL;
    const size_t base = (size_t)row * R + c0;
    T r_[CPL], xv[CPL], sc[CPL];
    T a_[NP > 0 ? NP : 1][CPL], d_[NP > 0 ? NP : 1][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const bool in = c0 + j < R;
        sc[j] = (in && rhs_scale) ? rhs_scale[(size_t)g * R + c0 + j] : T(1);
        r_[j] = in ? rhs[base + j] * sc[j] : T(0);
        xv[j] = T(0);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            a_[p][j] = in ? ((const T*)pa.aux[p])[base + j] : T(0);
            d_[p][j] = in ? ((const T*)pa.dual[p])[base + j] : T(0);
        }
    }
    const T* Mg = Minv + (size_t)g * R * R + c0;
    const int iters = NP == 0 ? 1 : n_inner;
    for (int it = 0; it < iters; ++it) {
        T s_[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            T sh = T(0);
#pragma unroll
            for (int p = 0; p < NP; ++p) sh += a_[p][j] - d_[p][j];
            s_[j] = NP > 0 ? rg * sh + r_[j] : r_[j];
            xv[j] = T(0);
        }
#pragma unroll
        for (int rr = 0; rr < 4 * CPL; ++rr) {
            const T sr = __shfl_sync(0xffffffffu, s_[rr % CPL], (lane & ~3) | (rr / CPL));
            if (rr < R) {
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    if (c0 + j < R) xv[j] = fma(sr, __ldg(Mg + (size_t)rr * R + j), xv[j]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int kind = pa.kind[p], nn = pa.nn[p];
            const T p0 = (T)pa.p0[p], p1 = (T)pa.p1[p];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const T v = xv[j] + d_[p][j];
                const T z = prox_elem<T>(v, kind, nn, p0, p1, rg);
                a_[p][j] = z;
                d_[p][j] = v - z;
            }
        }
    }
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        if (c0 + j < R) {
            x[base + j] = xv[j];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                ((T*)pa.aux[p])[base + j] = a_[p][j];
                ((T*)pa.dual[p])[base + j] = d_[p][j];
            }
            if (w_out) w_out[(size_t)row * ldw + c0 + j] = xv[j] * sc[j];
        }
    }
}

// CTA-per-group variant: Minv_g is staged once in shared memory in the padded RowLayout (128-bit loads at
// compile-time offsets, no runtime-R address arithmetic in the inner loops), the group's rows are processed in passes
// of 64, and B_g^T B_g can be accumulated on the fly with DMMA (BtB_out) for the following C-/A-updates.
// row_off == NULL: every block covers rows [64*blockIdx.x, +64) of a single group (C-mode).
template <typename T, int CPL, int NP>
__global__ void __launch_bounds__(256, 2)
admm_local_grouped_kernel(const int64_t* __restrict__ row_off, long long n, int R, const T* __restrict__ rhs,
                          const T* __restrict__ rhs_scale, const T* __restrict__ rho, const T* __restrict__ Minv,
                          PenArgs pa, int n_inner, T* __restrict__ x, T* __restrict__ w_out, int ldw,
                          T* __restrict__ BtB_out) {
    using L = RowLayout<T, CPL>;
    constexpr int NB = (CPL + 1) / 2;
    constexpr int LDT = 8 * NB + 2;  // 2*LDT == 4 (mod 16): conflict-free fragment reads
    extern __shared__ double al_smem[];
    T* Ms = (T*)al_smem;
    double* tile = al_smem + ((L::ELEMS * sizeof(T) + 7) / 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l4 = lane & 3;
    const int RR = R * R;
    const int g = row_off ? blockIdx.x : 0;
    long long r_begin, r_end;
    if (row_off) {
        r_begin = row_off[g];
        r_end = row_off[g + 1];
    } else {
        r_begin = (long long)blockIdx.x * kRowsPerPass;
        r_end = r_begin + kRowsPerPass < n ? r_begin + kRowsPerPass : n;
    }
    if (r_begin >= r_end) {
        if (BtB_out)
            for (int e = tid; e < RR; e += blockDim.x) BtB_out[(size_t)g * RR + e] = T(0);
        return;
    }
    for (int e = tid; e < L::ELEMS; e += blockDim.x) Ms[e] = T(0);
    if (BtB_out)
        for (int e = tid; e < kRowsPerPass * LDT; e += blockDim.x) tile[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < RR; e += blockDim.x) {
        const int i = e / R, c = e - i * R;
        Ms[i * L::LDM + (c / CPL) * L::CPLP + (c % CPL)] = Minv[(size_t)g * RR + e];
    }
    __syncthreads();
    const T rg = rho[g];
    const int c0 = l4 * CPL;
    const T* mseg = Ms + l4 * L::CPLP;
    T sc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) sc[j] = (c0 + j < R && rhs_scale) ? rhs_scale[(size_t)g * R + c0 + j] : T(1);
    double accB[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int gq = lane >> 2, tq = lane & 3;
    const int iters = NP == 0 ? 1 : n_inner;
    for (long long row0 = r_begin; row0 < r_end; row0 += kRowsPerPass) {
        const long long rowid = row0 + (tid >> 2);
        const bool valid = rowid < r_end;
        const long long row = valid ? rowid : r_end - 1;
        const size_t base = (size_t)row * R + c0;
        T r_[CPL], xv[CPL];
        T a_[NP > 0 ? NP : 1][CPL], d_[NP > 0 ? NP : 1][CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const bool in = c0 + j < R;
            r_[j] = in ? rhs[base + j] * sc[j] : T(0);
            xv[j] = T(0);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                a_[p][j] = in ? ((const T*)pa.aux[p])[base + j] : T(0);
                d_[p][j] = in ? ((const T*)pa.dual[p])[base + j] : T(0);
            }
        }
        for (int it = 0; it < iters; ++it) {
            T s_[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                T sh = T(0);
#pragma unroll
                for (int p = 0; p < NP; ++p) sh += a_[p][j] - d_[p][j];
                s_[j] = NP > 0 ? rg * sh + r_[j] : r_[j];
                xv[j] = T(0);
            }
            lane_matvec<T, CPL>(s_, mseg, lane, xv);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int kind = pa.kind[p], nn = pa.nn[p];
                const T p0 = (T)pa.p0[p], p1 = (T)pa.p1[p];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const T v = xv[j] + d_[p][j];
                    const T z = (c0 + j < R) ? prox_elem<T>(v, kind, nn, p0, p1, rg) : T(0);
                    a_[p][j] = z;
                    d_[p][j] = (c0 + j < R) ? v - z : T(0);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            if (c0 + j < R) {
                if (valid) {
                    x[base + j] = xv[j];
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        ((T*)pa.aux[p])[base + j] = a_[p][j];
                        ((T*)pa.dual[p])[base + j] = d_[p][j];
                    }
                    if (w_out) w_out[(size_t)row * ldw + c0 + j] = xv[j] * sc[j];
                }
                if (BtB_out) tile[(tid >> 2) * LDT + c0 + j] = valid ? (double)xv[j] : 0.0;
            }
        }
        if (BtB_out) {
            __syncthreads();
            tile_gram<NB>(tile, warp, gq, tq, accB);
            __syncthreads();
        }
    }
    if (BtB_out) store_gram<T, NB>(BtB_out + (size_t)g * RR, R, warp, gq, tq, accB);
}

template <typename T, int CPL>
int launch_local(int n_pen, long long n, int R, const void* rhs, const void* rhs_scale, int group_mode,
                 const int32_t* gor, const void* rho, const void* Minv, const PenArgs& pa, int n_inner, void* x,
                 void* w_out, int ldw, cudaStream_t st) {
    const long long threads = n * 4;
    const int grid = (int)((threads + 255) / 256);
#define B2_LAUNCH_LOCAL(NP)                                                                                       \
    admm_local_kernel<T, CPL, NP><<<grid, 256, 0, st>>>(n, R, (const T*)rhs, (const T*)rhs_scale, group_mode, gor, \
                                                        (const T*)rho, (const T*)Minv, pa, n_inner, (T*)x,        \
                                                        (T*)w_out, ldw)
    switch (n_pen) {
        case 0: B2_LAUNCH_LOCAL(0); break;
        case 1: B2_LAUNCH_LOCAL(1); break;
        default: B2_LAUNCH_LOCAL(2); break;
    }
#undef B2_LAUNCH_LOCAL
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T, int CPL>
int launch_local_grouped(int n_pen, const int64_t* row_off, int n_groups, long long n, int R, const void* rhs,
                         const void* rhs_scale, const void* rho, const void* Minv, const PenArgs& pa, int n_inner,
                         void* x, void* w_out, int ldw, void* BtB_out, cudaStream_t st) {
    using L = RowLayout<T, CPL>;
    constexpr int NB = (CPL + 1) / 2, LDT = 8 * NB + 2;
    const size_t smem = ((L::ELEMS * sizeof(T) + 7) / 8) * 8 + (size_t)kRowsPerPass * LDT * sizeof(double);
    const int grid = row_off ? n_groups : (int)((n + kRowsPerPass - 1) / kRowsPerPass);
#define B2_LAUNCH_GROUPED(NP)                                                                                      \
    admm_local_grouped_kernel<T, CPL, NP><<<grid, 256, smem, st>>>(row_off, n, R, (const T*)rhs, (const T*)rhs_scale, \
                                                                   (const T*)rho, (const T*)Minv, pa, n_inner, (T*)x, \
                                                                   (T*)w_out, ldw, (T*)BtB_out)
    switch (n_pen) {
        case 0: B2_LAUNCH_GROUPED(0); break;
        case 1: B2_LAUNCH_GROUPED(1); break;
        default: B2_LAUNCH_GROUPED(2); break;
    }
#undef B2_LAUNCH_GROUPED
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_admm_local(long long n, int R, const void* rhs, const void* rhs_scale, int group_mode,
                  const int32_t* group_of_row, const int64_t* row_off, int n_groups, const void* rho, const void* Minv,
                  const b2_penalty_desc* pens, int n_pen, int n_inner, void* x, void* w_out, int ldw, void* BtB_out,
                  int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(n_pen >= 0 && n_pen <= 2, "b2_admm_local fuses at most 2 penalties (got %d)", n_pen);
    B2_REQUIRE(group_mode != B2_GROUP_INDEXED || group_of_row != nullptr || row_off != nullptr,
               "group_of_row or row_off required");
    B2_REQUIRE(w_out == nullptr || rhs_scale != nullptr, "w_out needs rhs_scale (W = x o a)");
    B2_REQUIRE(BtB_out == nullptr || (group_mode == B2_GROUP_INDEXED && row_off != nullptr),
               "BtB_out needs the CTA-per-slice path (group_mode INDEXED with row_off)");
    if (n == 0) return B2_OK;
    PenArgs pa;
    {
        const int rc = b2_pack_penalties(pens, n_pen, &pa);
        if (rc != B2_OK) return rc;
    }
    for (int p = 0; p < n_pen; ++p)
        B2_REQUIRE(pa.kind[p] == B2_PEN_NONNEG || pa.kind[p] == B2_PEN_BOX || pa.kind[p] == B2_PEN_L1,
                   "b2_admm_local handles row-local penalties only (penalty %d has kind %d)", p, pa.kind[p]);
    if (group_mode == B2_GROUP_INDEXED && row_off != nullptr && n_inner > 0 && b2_option_value(B2_OPT_ADMM_LOCAL_MMA)) {
        // tensor-core formulation (admm_mma.cu) when it applies
        const int rc = b2_admm_local_mma_try(row_off, n_groups, R, rhs, rhs_scale, rho, Minv, pa, n_inner, x, w_out, ldw,
                                             BtB_out, dtype, st);
        if (rc >= 0) return rc;
    }
    const int CPL = (R + 3) / 4;
    // CTA-per-group path (operator staged in shared memory): slices via row_off, or a single group in 64-row blocks
    const bool grouped = (group_mode == B2_GROUP_INDEXED && row_off != nullptr) || group_mode == B2_GROUP_SINGLE;
    const int64_t* ro = group_mode == B2_GROUP_INDEXED ? row_off : nullptr;
#define B2_CASE_CPL(C)                                                                                            \
    case C:                                                                                                       \
        if (grouped)                                                                                              \
            B2_DISPATCH_DTYPE(dtype, return launch_local_grouped<T, C>(n_pen, ro, n_groups, n, R, rhs, rhs_scale, \
                                                                       rho, Minv, pa, n_inner, x, w_out, ldw,     \
                                                                       BtB_out, st));                             \
        B2_DISPATCH_DTYPE(dtype, return launch_local<T, C>(n_pen, n, R, rhs, rhs_scale, group_mode, group_of_row, \
                                                           rho, Minv, pa, n_inner, x, w_out, ldw, st));           \
        break
    switch (CPL) {
        B2_CASE_CPL(1);
        B2_CASE_CPL(2);
        B2_CASE_CPL(3);
        B2_CASE_CPL(4);
        B2_CASE_CPL(5);
        B2_CASE_CPL(6);
        B2_CASE_CPL(7);
        B2_CASE_CPL(8);
    }
#undef B2_CASE_CPL
    return B2_OK;
}

}  // extern "C"
