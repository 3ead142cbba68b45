// Tensor-core (DMMA) formulation of the fully fused row-local ADMM sub-solver for the B-mode (CTA per slice):
// the whole inner loop of admm_update_B (decomposition.py:259-289) for elementwise penalties (NonNegativity, Box, L1
// — or none) runs in registers, one pass over the state.  Same contract as admm_local_grouped_kernel (admm_fused.cu),
// selected by b2_admm_local when rows are 16-byte multiples.
//
// One CTA per slice: a producer warp streams 64-row tiles of rhs (= Y) and the aux/dual pairs into a shared-memory
// ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier); 8 consumer warps each own 8 rows of a tile in the MMA
// accumulator layout (mma_tiles.cuh) and iterate
//     x = (rho_g * sum_p(aux_p - dual_p) + rhs o a_g) Minv_g ;  aux_p = prox(x + dual_p) ;  dual_p = x + dual_p - aux_p
// n_inner times with the R x R product as DMMA.8x8x4 chains (the D registers of one product are the A operand of the
// next: no shuffles, no shared-memory round trip).  x, aux, dual and W = x o a_g go straight from registers to HBM;
// B_g^T B_g is accumulated with DMMA on the fly.
#include "admm_common.cuh"
#include "mma_tiles.cuh"

namespace {

constexpr int kConsWarps = 8;
constexpr int kTileRows = 8 * kConsWarps;
constexpr int kMaxIn = 5;  // rhs + 2 x (aux, dual)

struct LocalInputs {
    const void* ptr[kMaxIn];
    int n;
};

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kConsWarps * 32) : "memory"); }

template <class PL, typename T>
__device__ __forceinline__ void load_row(uint32_t srow, int t, int R, bool valid, double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = reg_col<PL>(b, e, t, R);
            v[b][e] = (valid && c >= 0) ? (double)lds_elem<T>(srow + (uint32_t)(c * sizeof(T))) : 0.0;
        }
    }
}

template <class PL, typename T>
__device__ __forceinline__ void store_row(T* __restrict__ grow, int t, int R, const double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        if (c0 >= 0 && c1 >= 0) {
            typename Vec2<T>::type pr;
            pr.x = (T)v[b][0];
            pr.y = (T)v[b][1];
            *(typename Vec2<T>::type*)(grow + c0) = pr;
        } else {
            if (c0 >= 0) grow[c0] = (T)v[b][0];
            if (c1 >= 0) grow[c1] = (T)v[b][1];
        }
    }
}

template <typename T, int NBF, int HALF, int NP>
__global__ void __launch_bounds__((kConsWarps + 1) * 32) __maxnreg__((NBF + HALF) <= 3 ? 96 : 168)
admm_local_mma_kernel(const int64_t* __restrict__ row_off, int R, LocalInputs in, const T* __restrict__ A,
                      const T* __restrict__ rho, const T* __restrict__ Minv, PenArgs pa, int n_inner,
                      T* __restrict__ x_out, T* __restrict__ w_out, int ldw, T* __restrict__ BtB_out, int stages) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    constexpr int NB = PL::NB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve: Ms | gram tiles (1 per consumer warp) | scale a_g | ring: stages x n_in x [64 x R] | barriers
    double* Ms = (double*)smem_raw;
    double* gtiles = Ms + PL::NPOS * PL::LDM;
    double* a_s = gtiles + kConsWarps * 8 * GA::LDT;
    unsigned char* ring = (unsigned char*)(((uintptr_t)(a_s + PL::NPOS) + 127) & ~(uintptr_t)127);
    const uint32_t arr_bytes = (uint32_t)(kTileRows * R * sizeof(T));
    const uint32_t stage_bytes = (uint32_t)in.n * arr_bytes;
    uint64_t* full = (uint64_t*)(ring + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const uint32_t ring_s = smem_u32(ring);

    const int g_slice = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long r_begin = row_off[g_slice], r_end = row_off[g_slice + 1];
    const int RR = R * R;
    if (r_begin >= r_end) {
        if (BtB_out)
            for (int e = tid; e < RR; e += blockDim.x) BtB_out[(size_t)g_slice * RR + e] = T(0);
        return;
    }
    const int n_tiles = (int)((r_end - r_begin + kTileRows - 1) / kTileRows);
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsWarps) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = 0; tile < n_tiles; ++tile) {
                const long long row0 = r_begin + (long long)tile * kTileRows;
                const int rows = (int)((r_end - row0) < kTileRows ? (r_end - row0) : kTileRows);
                const uint32_t bytes = (uint32_t)(rows * R * sizeof(T));
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = ring + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)in.n);
                for (int a = 0; a < in.n; ++a)
                    bulk_load_1d(st + (size_t)a * arr_bytes, (const T*)in.ptr[a] + (size_t)row0 * R, bytes, &full[s]);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }

    const int ctid = tid, cthreads = kConsWarps * 32;
    stage_operator<PL, T>(Minv + (size_t)g_slice * RR, R, Ms, ctid, cthreads);
    for (int e = ctid; e < PL::NPOS; e += cthreads) {
        const int c = PL::col_of(e, R);
        a_s[e] = (c >= 0 && A) ? (double)A[(size_t)g_slice * R + c] : (c >= 0 ? 1.0 : 0.0);
    }
    consumer_barrier();

    const int g = lane >> 2, t = lane & 3;
    const double rg = (double)rho[g_slice];
    double sc[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        sc[b][0] = a_s[8 * b + 2 * t];
        sc[b][1] = a_s[8 * b + 2 * t + 1];
    }
    GA accB;
    accB.clear();
    double* tileG = gtiles + (size_t)warp * 8 * GA::LDT;
    const int iters = NP == 0 ? 1 : n_inner;

    int s = 0;
    uint32_t ph = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        const long long row = r_begin + (long long)tile * kTileRows + warp * 8 + g;
        const bool valid = row < r_end;
        mbar_wait(&full[s], ph);
        const uint32_t st = ring_s + (uint32_t)s * stage_bytes + (uint32_t)((warp * 8 + g) * R * sizeof(T));
        double r_[NB][2], xv[NB][2], ax[NP > 0 ? NP : 1][NB][2], du[NP > 0 ? NP : 1][NB][2];
        load_row<PL, T>(st, t, R, valid, r_);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) r_[b][e] *= sc[b][e];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            load_row<PL, T>(st + (uint32_t)(1 + 2 * p) * arr_bytes, t, R, valid, ax[p]);
            load_row<PL, T>(st + (uint32_t)(2 + 2 * p) * arr_bytes, t, R, valid, du[p]);
        }
        for (int it = 0; it < iters; ++it) {
            double sv[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double sh = 0.0;
#pragma unroll
                    for (int p = 0; p < NP; ++p) sh += ax[p][b][e] - du[p][b][e];
                    sv[b][e] = NP > 0 ? fma(rg, sh, r_[b][e]) : r_[b][e];
                }
            mma_rowmat<PL>(sv, Ms, g, t, xv);
            if (it == 0) {  // x depends on every value loaded from the stage: once it exists the stage can be refilled
                int dep = 0;
#pragma unroll
                for (int b = 0; b < NB; ++b) dep = max(dep, max(dep_bits_of(xv[b][0]), dep_bits_of(xv[b][1])));
                stage_release(&empty[s], lane, dep);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int kind = pa.kind[p], nn = pa.nn[p];
                const T p0 = (T)pa.p0[p], p1 = (T)pa.p1[p];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        // prox in the storage precision T, like the row-wise kernels (admm.cu)
                        const T vv = (T)xv[b][e] + (T)du[p][b][e];
                        const T z = prox_elem<T>(vv, kind, nn, p0, p1, (T)rg);
                        ax[p][b][e] = (double)z;
                        du[p][b][e] = (double)(vv - z);
                    }
            }
        }
        const size_t goff = (size_t)row * R;
        if (valid) {
            store_row<PL, T>(x_out + goff, t, R, xv);
            if (w_out) {
                double wv[NB][2];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) wv[b][e] = xv[b][e] * sc[b][e];
                store_row<PL, T>(w_out + (size_t)row * ldw, t, R, wv);
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                store_row<PL, T>((T*)pa.aux[p] + goff, t, R, ax[p]);
                store_row<PL, T>((T*)pa.dual[p] + goff, t, R, du[p]);
            }
        }
        if (BtB_out) {
            double xz[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool pad = reg_col<PL>(b, e, t, R) < 0;
                    xz[b][e] = (valid && !pad) ? (double)(T)xv[b][e] : 0.0;  // Gram of x as stored
                }
            accB.add(xz, tileG, g, t);
        }
    }
    if (BtB_out) {
        consumer_barrier();  // the ring is free now: every warp has consumed every tile
        gram_reduce_store<PL, T>(accB, (double*)ring, warp, lane, kConsWarps, ctid, cthreads, R,
                                 BtB_out + (size_t)g_slice * RR, consumer_barrier);
    }
}

template <typename T, int NBF, int HALF, int NP>
int launch_k(const int64_t* row_off, int n_groups, int R, const LocalInputs& in, const void* A, const void* rho,
             const void* Minv, const PenArgs& pa, int n_inner, void* x, void* w_out, int ldw, void* BtB_out,
             cudaStream_t st) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    const size_t fixed = (size_t)(PL::NPOS * PL::LDM + kConsWarps * 8 * GA::LDT + PL::NPOS) * sizeof(double) + 128;
    const size_t stage_bytes = (size_t)in.n * kTileRows * R * sizeof(T);
    const size_t red_bytes = (size_t)kConsWarps * GA::NPAIR * 64 * sizeof(double);
    const size_t budget = 113 * 1024;  // two CTAs per SM when the tiles allow it
    int stages = (int)((budget - fixed - 64 - 128) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    size_t ring_bytes = (size_t)stages * stage_bytes;
    if (ring_bytes < red_bytes) ring_bytes = red_bytes;
    const size_t smem = fixed + ring_bytes + 2 * (size_t)stages * sizeof(uint64_t) + 64;
    if (smem > 227 * 1024) return -1;
    auto kern = admm_local_mma_kernel<T, NBF, HALF, NP>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ask for the largest shared-memory carve-out: otherwise the driver sizes it for ONE resident CTA
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<n_groups, (kConsWarps + 1) * 32, smem, st>>>(row_off, R, in, (const T*)A, (const T*)rho, (const T*)Minv, pa,
                                                        n_inner, (T*)x, (T*)w_out, ldw, (T*)BtB_out, stages);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T, int NBF, int HALF>
int launch_np(int n_pen, const int64_t* row_off, int n_groups, int R, const LocalInputs& in, const void* A,
              const void* rho, const void* Minv, const PenArgs& pa, int n_inner, void* x, void* w_out, int ldw,
              void* BtB_out, cudaStream_t st) {
    switch (n_pen) {
        case 0: return launch_k<T, NBF, HALF, 0>(row_off, n_groups, R, in, A, rho, Minv, pa, n_inner, x, w_out, ldw, BtB_out, st);
        case 1: return launch_k<T, NBF, HALF, 1>(row_off, n_groups, R, in, A, rho, Minv, pa, n_inner, x, w_out, ldw, BtB_out, st);
        case 2: return launch_k<T, NBF, HALF, 2>(row_off, n_groups, R, in, A, rho, Minv, pa, n_inner, x, w_out, ldw, BtB_out, st);
    }
    return -1;
}

template <typename T>
int dispatch(int n_pen, const int64_t* row_off, int n_groups, int R, const LocalInputs& in, const void* A,
             const void* rho, const void* Minv, const PenArgs& pa, int n_inner, void* x, void* w_out, int ldw,
             void* BtB_out, cudaStream_t st) {
    const int nbf = R / 8, rem = R % 8;
    const int NBF = rem >= 5 ? nbf + 1 : nbf, HALF = (rem >= 1 && rem <= 4) ? 1 : 0;
#define B2_MMA_CASE(F, H)                                                                                       \
    if (NBF == F && HALF == H)                                                                                  \
        return launch_np<T, F, H>(n_pen, row_off, n_groups, R, in, A, rho, Minv, pa, n_inner, x, w_out, ldw,      \
                                  BtB_out, st);
    B2_MMA_CASE(0, 1)
    B2_MMA_CASE(1, 0)
    B2_MMA_CASE(1, 1)
    B2_MMA_CASE(2, 0)
    B2_MMA_CASE(2, 1)
    B2_MMA_CASE(3, 0)
    B2_MMA_CASE(3, 1)
    B2_MMA_CASE(4, 0)
#undef B2_MMA_CASE
    return -1;
}

}  // namespace

// Returns B2_OK after launching, a positive error code on failure, or -1 when this formulation does not apply
// (the caller then uses admm_local_grouped_kernel).
int b2_admm_local_mma_try(const int64_t* row_off, int n_groups, int R, const void* rhs, const void* rhs_scale,
                          const void* rho, const void* Minv, const PenArgs& pa, int n_inner, void* x, void* w_out,
                          int ldw, void* BtB_out, int dtype, cudaStream_t st) {
    const size_t es = dtype == B2_F64 ? 8 : 4;
    if (((size_t)R * es) % 16 != 0 || pa.n_pen > 2 || !x) return -1;
    if (w_out && (ldw % 2 != 0 || ((uintptr_t)w_out) % (2 * es) != 0)) return -1;
    LocalInputs in;
    in.n = 0;
    in.ptr[in.n++] = rhs;
    for (int p = 0; p < pa.n_pen; ++p) {
        in.ptr[in.n++] = pa.aux[p];
        in.ptr[in.n++] = pa.dual[p];
    }
    for (int a = 0; a < in.n; ++a)
        if (((uintptr_t)in.ptr[a]) % 16 != 0) return -1;
    if (((uintptr_t)x) % 16 != 0) return -1;
    if (dtype == B2_F64)
        return dispatch<double>(pa.n_pen, row_off, n_groups, R, in, rhs_scale, rho, Minv, pa, n_inner, x, w_out, ldw,
                                BtB_out, st);
    return dispatch<float>(pa.n_pen, row_off, n_groups, R, in, rhs_scale, rho, Minv, pa, n_inner, x, w_out, ldw,
                           BtB_out, st);
}
