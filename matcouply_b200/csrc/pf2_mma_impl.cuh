// Tensor-core (DMMA) formulation of the fused PARAFAC2 B-mode row pass (reference decomposition.py:259-289 with
// penalties.py:1224-1281); same contract as pf2_rowpass_kernel in pf2_fused.cu, selected by b2_pf2_rowpass whenever
// the row size R*sizeof(T) is a multiple of 16 bytes (TMA bulk copies need 16-byte granularity).
//
// One CTA per slice: a producer warp streams 64-row tiles of every input array (Y, the PARAFAC2 pre-image V, the
// aux/dual pairs of the other penalties) into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier); 8 consumer warps each own 8 rows of a tile, pull them into registers in the MMA accumulator layout and
// run the whole row update as three small tensor-core products per row block (mma_tiles.cuh):
//     P Delta = V T_g            (deferred prox of the previous inner iteration, T_g = W_g Delta)
//     x       = (rho_g * sum_p(aux_p - dual_p) + Y o a_g) Minv_g
//     S_g    += V'^T V',  V' = x + dual_pf2        (and B_g^T B_g += x^T x on the last inner iteration)
// Results go straight from registers to HBM; no block-wide barrier inside the row loop.
#pragma once
#include "admm_common.cuh"
#include "mma_tiles.cuh"

namespace {

constexpr int kConsWarps = 8;
constexpr int kTileRows = 8 * kConsWarps;
constexpr int kMaxInputs = 3 + 2 * (kMaxPen - 1);
constexpr int kMaxExtra = 2;  // extra (non-PARAFAC2) penalties whose dual stays in registers

struct RowpassInputs {
    const void* ptr[kMaxInputs];  // [0] Y, [1] pf dual (V), [2] pf aux (P Delta, only when !deferred), then aux_p, dual_p
    int n;
};

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kConsWarps * 32) : "memory"); }

// row in D layout from a shared-memory tile row (shared-space byte address, element type T)
template <typename T>
__device__ __forceinline__ void lds_pair(uint32_t saddr, double& a, double& b);
template <>
__device__ __forceinline__ void lds_pair<double>(uint32_t saddr, double& a, double& b) {
    const int4 q = lds_b128(saddr);
    a = __hiloint2double(q.y, q.x);
    b = __hiloint2double(q.w, q.z);
}
template <>
__device__ __forceinline__ void lds_pair<float>(uint32_t saddr, double& a, double& b) {
    a = (double)lds_elem<float>(saddr);
    b = (double)lds_elem<float>(saddr + 4);
}

template <class PL, typename T>
__device__ __forceinline__ void load_row(uint32_t srow, int t, int R, bool valid, double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        v[b][0] = v[b][1] = 0.0;
        if (valid) {
            if (c0 >= 0 && c1 >= 0) {  // adjacent columns c0, c0 + 1 (c0 even; rows are 16-byte aligned): one load
                lds_pair<T>(srow + (uint32_t)(c0 * sizeof(T)), v[b][0], v[b][1]);
            } else {
                if (c0 >= 0) v[b][0] = (double)lds_elem<T>(srow + (uint32_t)(c0 * sizeof(T)));
                if (c1 >= 0) v[b][1] = (double)lds_elem<T>(srow + (uint32_t)(c1 * sizeof(T)));
            }
        }
    }
}

template <class PL, typename T>
__device__ __forceinline__ void store_row(T* __restrict__ grow, int t, int R, const double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        if (c0 >= 0 && c1 >= 0) {  // adjacent columns, 2*sizeof(T)-aligned because R*sizeof(T) % 16 == 0
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

// NEXTRA: number of penalties besides PARAFAC2 (compile time: their duals stay in registers).  LAST: the last inner
// iteration also emits x, W = x o a and B^T B.  Two CTAs per SM (9 warps each: 5 warps on one SM sub-partition x 96 registers fit its 16K-register file) matter: the
// row loop is latency bound, and a second resident CTA doubles the warps that hide it.
template <typename T, int NBF, int HALF, int NEXTRA, bool LAST, int K1>
__global__ void __launch_bounds__((kConsWarps + 1) * 32) __maxnreg__((NBF + HALF) <= 3 ? 96 : 168)
pf2_rowpass_mma_kernel(const int64_t* __restrict__ row_off, int R, RowpassInputs in, const T* __restrict__ A,
                       const T* __restrict__ rho, const T* __restrict__ Minv, PenArgs pa, int deferred,
                       const T* __restrict__ Wmat, const T* __restrict__ Delta, T* __restrict__ x_out,
                       T* __restrict__ w_out, int ldw, T* __restrict__ S_out, T* __restrict__ BtB_out, int stages) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    constexpr int NB = PL::NB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve: Ms | Ts | gram tiles (1 per consumer warp) | scale a_g | ring: stages x n_in x [64 x R] | barriers
    double* Ms = (double*)smem_raw;
    double* Ts = Ms + PL::NPOS * PL::LDM;
    double* gtiles = Ts + PL::NPOS * PL::LDM;
    double* a_s = gtiles + kConsWarps * 8 * GA::LDT;
    unsigned char* ring = (unsigned char*)(((uintptr_t)(a_s + PL::NPOS) + 127) & ~(uintptr_t)127);
    const uint32_t arr_bytes = (uint32_t)(kTileRows * R * sizeof(T));  // multiple of 16 (R*sizeof(T) % 16 == 0)
    const uint32_t stage_bytes = (uint32_t)in.n * arr_bytes;
    uint64_t* full = (uint64_t*)(ring + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const uint32_t ring_s = smem_u32(ring);

    const int g_slice = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long r_begin = row_off[g_slice], r_end = row_off[g_slice + 1];
    const int RR = R * R;
    if (r_begin >= r_end) {
        for (int e = tid; e < RR; e += blockDim.x) {
            S_out[(size_t)g_slice * RR + e] = T(0);
            if (LAST) BtB_out[(size_t)g_slice * RR + e] = T(0);
        }
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
        // ===== producer warp =====
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

    // ===== consumers: stage the slice operators while the first tiles are in flight =====
    const int ctid = tid, cthreads = kConsWarps * 32;
    stage_operator<PL, T>(Minv + (size_t)g_slice * RR, R, Ms, ctid, cthreads);
    if (deferred & 1) {
        // T_g = W_g Delta, formed in position order directly (R^3 FMAs per CTA).  W_g and Delta are first staged in
        // shared memory (the Gram tiles are free until the row loop): ONE global round trip instead of 2 R dependent
        // ones per thread in the k loop.
        double* wsm = gtiles;
        double* dsm = gtiles + RR;
        const T* Wg = Wmat + (size_t)g_slice * RR;
        for (int e = ctid; e < RR; e += cthreads) {
            wsm[e] = (double)Wg[e];
            dsm[e] = (double)Delta[e];
        }
        consumer_barrier();
        for (int e = ctid; e < PL::NPOS * PL::LDM; e += cthreads) {
            const int pr = e / PL::LDM, pc = e - pr * PL::LDM;
            double v = 0.0;
            if (pc < PL::NPOS) {
                const int r = PL::col_of(pr, R), c = PL::col_of(pc, R);
                if (r >= 0 && c >= 0)
                    for (int k = 0; k < R; ++k) v = fma(wsm[r * R + k], dsm[k * R + c], v);
            }
            Ts[e] = v;
        }
    }
    for (int e = ctid; e < PL::NPOS; e += cthreads) {
        const int c = PL::col_of(e, R);
        a_s[e] = c >= 0 ? (double)A[(size_t)g_slice * R + c] : 0.0;
    }
    consumer_barrier();

    const int g = lane >> 2, t = lane & 3;
    const double rg = (double)rho[g_slice];
    const bool tin = (deferred & 2) != 0, tout = (deferred & 4) != 0;  // T-only state of the elementwise extras
    deferred &= 1;
    double sc[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        sc[b][0] = a_s[8 * b + 2 * t];
        sc[b][1] = a_s[8 * b + 2 * t + 1];
    }
    GA accS, accB;
    accS.clear();
    if (LAST) accB.clear();
    double* tileG = gtiles + (size_t)warp * 8 * GA::LDT;  // shared by the S and B^T B accumulations
    T* pf_dual = (T*)pa.dual[0];

    int s = 0;
    uint32_t ph = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
        const long long row = r_begin + (long long)tile * kTileRows + warp * 8 + g;
        const bool valid = row < r_end;
        mbar_wait(&full[s], ph);
        const uint32_t st = ring_s + (uint32_t)s * stage_bytes + (uint32_t)((warp * 8 + g) * R * sizeof(T));
        double y[NB][2], v[NB][2], dpf[NB][2], sh[NB][2], du[NEXTRA > 0 ? NEXTRA : 1][NB][2];
        load_row<PL, T>(st, t, R, valid, y);
        load_row<PL, T>(st + arr_bytes, t, R, valid, v);
        int a_idx = 2;
        if (deferred) {
            double pd[NB][2];
            mma_rowmat<PL>(v, Ts, g, t, pd);
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    dpf[b][e] = v[b][e] - pd[b][e];    // dual = V - P Delta          (decomposition.py:282-285)
                    sh[b][e] = pd[b][e] - dpf[b][e];   // aux - dual = P Delta - dual (penalties.py:1280-1281)
                }
        } else {
            double pd[NB][2];
            load_row<PL, T>(st + 2 * arr_bytes, t, R, valid, pd);
            a_idx = 3;
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    dpf[b][e] = v[b][e];
                    sh[b][e] = pd[b][e] - v[b][e];
                }
        }
        {
            int ai = a_idx;
#pragma unroll
            for (int p = 0; p < NEXTRA; ++p) {
                const int kind = (K1 >= 0 && p == 0) ? K1 : pa.kind[p + 1];  // K1: compile-time kind of companion 0
                const bool elementwise = kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1;
                if (tin && elementwise) {
                    // T-only state: the dual slot holds the previous prox argument T = x + dual; aux = prox(T) and
                    // dual = T - aux are recomputed (bit-identical to what the explicit path would have stored)
                    double tt[NB][2];
                    load_row<PL, T>(st + (uint32_t)ai * arr_bytes, t, R, valid, tt);
                    ai += 1;
                    const int nn = pa.nn[p + 1];
                    const T p0 = (T)pa.p0[p + 1], p1 = (T)pa.p1[p + 1];
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const bool ok = valid && reg_col<PL>(b, e, t, R) >= 0;
                            const T tv = (T)tt[b][e];
                            const T z = ok ? prox_elem<T>(tv, kind, nn, p0, p1, (T)rg) : T(0);
                            du[p][b][e] = ok ? (double)(tv - z) : 0.0;
                            sh[b][e] += (double)z - du[p][b][e];
                        }
                } else {
                    double ax[NB][2];
                    load_row<PL, T>(st + (uint32_t)ai * arr_bytes, t, R, valid, ax);
                    load_row<PL, T>(st + (uint32_t)(ai + 1) * arr_bytes, t, R, valid, du[p]);
                    ai += 2;
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) sh[b][e] += ax[b][e] - du[p][b][e];
                }
            }
        }
        double sv[NB][2], xv[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) sv[b][e] = fma(rg, sh[b][e], y[b][e] * sc[b][e]);
        mma_rowmat<PL>(sv, Ms, g, t, xv);
        {   // x depends on every value loaded from the stage: once it exists the stage can be refilled
            int dep = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) dep = max(dep, max(dep_bits_of(xv[b][0]), dep_bits_of(xv[b][1])));
            stage_release(&empty[s], lane, dep);
            if (++s == stages) {
                s = 0;
                ph ^= 1;
            }
        }
        double vn[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) vn[b][e] = valid ? xv[b][e] + dpf[b][e] : 0.0;  // V' = x + dual_pf2
        const size_t goff = (size_t)row * R;
        if (valid) {
            store_row<PL, T>(pf_dual + goff, t, R, vn);
            if (LAST) store_row<PL, T>(x_out + goff, t, R, xv);
            if (LAST) {
                double wv[NB][2];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) wv[b][e] = xv[b][e] * sc[b][e];
                store_row<PL, T>(w_out + (size_t)row * ldw, t, R, wv);
            }
#pragma unroll
            for (int p = 0; p < NEXTRA; ++p) {
                {
                    const int kind = (K1 >= 0 && p == 0) ? K1 : pa.kind[p + 1], nn = pa.nn[p + 1];
                    const bool elementwise = kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1;
                    const T p0 = (T)pa.p0[p + 1], p1 = (T)pa.p1[p + 1];
                    double zo[NB][2], dn[NB][2];
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            // prox in the storage precision T, like the row-wise kernels (admm.cu)
                            const T vv = (T)xv[b][e] + (T)du[p][b][e];
                            const T z = elementwise ? prox_elem<T>(vv, kind, nn, p0, p1, (T)rg) : T(0);
                            zo[b][e] = (double)z;
                            dn[b][e] = elementwise ? (double)(vv - z) : (double)vv;
                        }
                    if (elementwise && tout) {
                        // T-only state: keep the prox argument, the next pass recomputes aux and dual from it
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int e = 0; e < 2; ++e) dn[b][e] = (double)((T)xv[b][e] + (T)du[p][b][e]);
                    } else if (elementwise) {
                        store_row<PL, T>((T*)pa.aux[p + 1] + goff, t, R, zo);
                    }
                    store_row<PL, T>((T*)pa.dual[p + 1] + goff, t, R, dn);  // column-coupled: V, finished later
                }
            }
        }
        accS.add(vn, tileG, g, t, LAST ? accB.dep() : 0);
        if (LAST) {
            double xz[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) xz[b][e] = valid ? xv[b][e] : 0.0;
            accB.add(xz, tileG, g, t, accS.dep());
        }
    }
    // cross-warp reduction of the Gram partials; the ring is free now (all tiles consumed by every warp after the barrier)
    consumer_barrier();
    double* red = (double*)ring;
    gram_reduce_store<PL, T>(accS, red, warp, lane, kConsWarps, ctid, cthreads, R, S_out + (size_t)g_slice * RR,
                             consumer_barrier);
    if (LAST) {
        consumer_barrier();
        gram_reduce_store<PL, T>(accB, red, warp, lane, kConsWarps, ctid, cthreads, R, BtB_out + (size_t)g_slice * RR,
                                 consumer_barrier);
    }
}

template <typename T, int NBF, int HALF, int NEXTRA, bool LAST, int K1>
int launch_mma_k(const int64_t* row_off, int n_groups, int R, const RowpassInputs& in, const void* A, const void* rho,
                 const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta, void* x,
                 void* w_out, int ldw, void* S_out, void* BtB_out, cudaStream_t st) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    const size_t fixed = (size_t)(2 * PL::NPOS * PL::LDM + kConsWarps * 8 * GA::LDT + PL::NPOS) * sizeof(double) + 128;
    const size_t stage_bytes = (size_t)in.n * kTileRows * R * sizeof(T);
    const size_t red_bytes = (size_t)kConsWarps * GA::NPAIR * 64 * sizeof(double);
    // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2 = 113 KB each
    const size_t budget = 113 * 1024;
    int stages = (int)((budget - fixed - 64 - 128) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    size_t ring_bytes = (size_t)stages * stage_bytes;
    if (ring_bytes < red_bytes) ring_bytes = red_bytes;
    const size_t smem = fixed + ring_bytes + 2 * (size_t)stages * sizeof(uint64_t) + 64;
    if (smem > 227 * 1024) return -1;  // too many / too wide input arrays for two stages: use the shuffle kernel
    auto kern = pf2_rowpass_mma_kernel<T, NBF, HALF, NEXTRA, LAST, K1>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ask for the largest shared-memory carve-out: otherwise the driver sizes it for ONE resident CTA
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<n_groups, (kConsWarps + 1) * 32, smem, st>>>(row_off, R, in, (const T*)A, (const T*)rho, (const T*)Minv, pa,
                                                        deferred, (const T*)Wmat, (const T*)Delta, (T*)x, (T*)w_out, ldw,
                                                        (T*)S_out, (T*)BtB_out, stages);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T, int NBF, int HALF>
int launch_mma(const int64_t* row_off, int n_groups, int R, const RowpassInputs& in, const void* A, const void* rho,
               const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta, void* x,
               void* w_out, int ldw, void* S_out, void* BtB_out, cudaStream_t st) {
    const int n_extra = pa.n_pen - 1;
    const bool last = x != nullptr;
    if (last && !(w_out && BtB_out)) return -1;   // the fused engine always asks for all three on the last iteration
    if (!last && (w_out || BtB_out)) return -1;
    // the common companion (non-negativity next to PARAFAC2) gets its kind at compile time: no per-element dispatch
    const bool nn1 = n_extra == 1 && pa.kind[1] == B2_PEN_NONNEG;
#define B2_MMA_K(NE, LA, KK)                                                                                          \
    return launch_mma_k<T, NBF, HALF, NE, LA, KK>(row_off, n_groups, R, in, A, rho, Minv, pa, deferred, Wmat, Delta, x, \
                                                  w_out, ldw, S_out, BtB_out, st);
    if (n_extra == 0 && !last) { B2_MMA_K(0, false, -1) }
    if (n_extra == 0 && last) { B2_MMA_K(0, true, -1) }
    if (nn1 && !last) { B2_MMA_K(1, false, B2_PEN_NONNEG) }
    if (nn1 && last) { B2_MMA_K(1, true, B2_PEN_NONNEG) }
    if (n_extra == 1 && !last) { B2_MMA_K(1, false, -1) }
    if (n_extra == 1 && last) { B2_MMA_K(1, true, -1) }
    if (n_extra == 2 && !last) { B2_MMA_K(2, false, -1) }
    if (n_extra == 2 && last) { B2_MMA_K(2, true, -1) }
#undef B2_MMA_K
    return -1;
}

// dtype-specific dispatch over the position layouts; defined in pf2_mma_f64.cu / pf2_mma_f32.cu
template <typename T>
int pf2_rowpass_mma_dispatch(const int64_t* row_off, int n_groups, int R, const RowpassInputs& in, const void* A,
                             const void* rho, const void* Minv, const PenArgs& pa, int deferred, const void* Wmat,
                             const void* Delta, void* x, void* w_out, int ldw, void* S_out, void* BtB_out,
                             cudaStream_t st) {
    const int nbf = R / 8, rem = R % 8;
    const int NBF = rem >= 5 ? nbf + 1 : nbf, HALF = (rem >= 1 && rem <= 4) ? 1 : 0;
#define B2_MMA_CASE(F, H)                                                                                           \
    if (NBF == F && HALF == H)                                                                                      \
        return launch_mma<T, F, H>(row_off, n_groups, R, in, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out, ldw, \
                                   S_out, BtB_out, st);
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
