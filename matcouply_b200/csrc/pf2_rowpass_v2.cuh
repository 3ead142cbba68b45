// Second-generation tensor-core PARAFAC2 row pass (fp64), the steady-state fast path of b2_pf2_rowpass.
//
// Same contract and the SAME operation order as pf2_rowpass_mma_kernel (pf2_mma_impl.cuh; reference
// decomposition.py:259-289 with penalties.py:1224-1281) — results are bit-identical — but specialised for what the
// steady state of a PARAFAC2 B-update actually runs, so that the per-row instruction stream is ~3x shorter:
//   * rank fixed at compile time, R = 8 NBF + 4 HALF (every position of the MMA layout is a real column: no
//     per-element column predicates, every register pair moves with one 128-bit access);
//   * the PARAFAC2 prox always deferred (V, W_g, Delta in; bit 0 of `deferred`), at most ONE elementwise companion
//     whose kind is a template parameter (non-negativity next to PARAFAC2 is the standard combination);
//   * full 64-row tiles take a path without row predicates; only the last, ragged tile of a slice pays for them;
//   * no producer warp: 8 consumer warps per CTA (two CTAs per SM at <= 128 registers instead of 96 with a ninth
//     warp).  The warp that releases a ring stage LAST (shared-memory counter) issues the TMA refill of that stage,
//     so nobody ever waits on an "empty" barrier.
// Everything else (one CTA per slice, 1-D TMA bulk copies into an mbarrier ring, rows in the MMA accumulator layout,
// x = s M chained without shuffles, Gram accumulation through a warp-private tile) is as in pf2_mma_impl.cuh.
#pragma once
#include <type_traits>

#include "admm_common.cuh"
#include "mma_tiles.cuh"

namespace rp2 {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 8 * kWarps;
constexpr int kMaxStages = 4;

struct Args {
    const int64_t* row_off;
    const double* in[4];  // ring inputs: [0] Y, [1] V (PARAFAC2 dual slot), then the companion's T | (aux, dual)
    const double *A, *rho, *Minv, *Wmat, *Delta;
    double* pf_dual;  // V' out (same buffer as in[1])
    double *c_aux, *c_dual;  // companion outputs: T' -> c_dual (TOUT) or prox -> c_aux, dual -> c_dual
    double *x_out, *w_out;
    int ldw;
    double *S_out, *BtB_out;
    double* stats_part;  // optional (LAST, one companion): per slice [sum (aux1 - x)^2, sum x^2, sum |x|]
    int stages;
};

__device__ __forceinline__ unsigned atom_inc_acqrel(unsigned* saddr) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(saddr)) : "memory");
    return old;
}

// one row (this lane's share) of a staged [64 x R] array -> D layout; invalid rows read as zeros
template <class PL, bool FULL>
__device__ __forceinline__ void ld_row(uint32_t srow, int t, bool valid, double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) v[b][0] = v[b][1] = 0.0;
    if (FULL || valid) {
#pragma unroll
        for (int b = 0; b < PL::NBF; ++b) {
            const int4 q = lds_b128(srow + (uint32_t)((8 * b + 2 * t) * 8));
            v[b][0] = __hiloint2double(q.y, q.x);
            v[b][1] = __hiloint2double(q.w, q.z);
        }
        if constexpr (PL::HALF) v[PL::NBF][0] = lds_f64(srow + (uint32_t)((8 * PL::NBF + t) * 8));
    }
}

template <class PL>
__device__ __forceinline__ void st_row(double* __restrict__ grow, int t, const double (&v)[PL::NB][2]) {
#pragma unroll
    for (int b = 0; b < PL::NBF; ++b) *(double2*)(grow + 8 * b + 2 * t) = make_double2(v[b][0], v[b][1]);
    if constexpr (PL::HALF) grow[8 * PL::NBF + t] = v[PL::NBF][0];
}

// K1: kind of the single elementwise companion (B2_PEN_NONNEG) or -1 for none.
// TIN / TOUT: the companion arrives / leaves as ONE array T = x + dual (bits 1 / 2 of `deferred`, see the header).
// LAST: also emit x, W = x o a and B^T B.
template <int NBF, int HALF, int K1, bool TIN, bool TOUT, bool LAST>
__global__ void __launch_bounds__(kThreads, (NBF + HALF) <= 3 ? 2 : 1)
pf2_rowpass_v2_kernel(const Args a) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    constexpr int NB = PL::NB, R = 8 * NBF + 4 * HALF, RR = R * R;
    constexpr int NIN = 2 + (K1 >= 0 ? (TIN ? 1 : 2) : 0);
    constexpr uint32_t ARR = (uint32_t)(kTile * R * sizeof(double));
    constexpr uint32_t STAGE = NIN * ARR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve: Ms | Ts | gram tiles (1 per warp) | scale a_g | release counters | ring: stages x NIN x [64 x R] | barriers
    double* Ms = (double*)smem_raw;
    double* Ts = Ms + PL::NPOS * PL::LDM;
    double* gtiles = Ts + PL::NPOS * PL::LDM;
    double* a_s = gtiles + kWarps * 8 * GA::LDT;
    unsigned* cnt = (unsigned*)(a_s + PL::NPOS);
    unsigned char* ring = (unsigned char*)(((uintptr_t)(cnt + kMaxStages) + 127) & ~(uintptr_t)127);
    const int stages = a.stages;
    uint64_t* full = (uint64_t*)(ring + (size_t)stages * STAGE);
    const uint32_t ring_s = smem_u32(ring);

    const int g_slice = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long r_begin = a.row_off[g_slice], r_end = a.row_off[g_slice + 1];
    if (r_begin >= r_end) {
        for (int e = tid; e < RR; e += kThreads) {
            a.S_out[(size_t)g_slice * RR + e] = 0.0;
            if (LAST) a.BtB_out[(size_t)g_slice * RR + e] = 0.0;
        }
        if (LAST && K1 >= 0 && a.stats_part != nullptr && tid < 3) a.stats_part[(size_t)g_slice * 3 + tid] = 0.0;
        return;
    }
    const int n_tiles = (int)((r_end - r_begin + kTile - 1) / kTile);
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            cnt[s] = 0u;
        }
        fence_mbar_init();
    }
    __syncthreads();

    // fill of ring slot `slot` with tile `tile` (one thread)
    auto issue = [&](int tile, int slot) {
        const long long row0 = r_begin + (long long)tile * kTile;
        const int rows = (int)((r_end - row0) < kTile ? (r_end - row0) : kTile);
        const uint32_t bytes = (uint32_t)(rows * R * sizeof(double));
        unsigned char* st = ring + (size_t)slot * STAGE;
        mbar_arrive_expect_tx(&full[slot], bytes * (uint32_t)NIN);
#pragma unroll
        for (int i = 0; i < NIN; ++i) bulk_load_1d(st + (size_t)i * ARR, a.in[i] + (size_t)row0 * R, bytes, &full[slot]);
    };
    if (tid == 0)
        for (int tile = 0; tile < stages && tile < n_tiles; ++tile) issue(tile, tile);

    // stage the slice operators while the first tiles are in flight: Ms = Minv_g, Ts = W_g Delta (position order)
    stage_operator<PL, double>(a.Minv + (size_t)g_slice * RR, R, Ms, tid, kThreads);
    {
        double* wsm = gtiles;  // free until the row loop
        double* dsm = gtiles + RR;
        const double* Wg = a.Wmat + (size_t)g_slice * RR;
        for (int e = tid; e < RR; e += kThreads) {
            wsm[e] = Wg[e];
            dsm[e] = a.Delta[e];
        }
        __syncthreads();
        for (int e = tid; e < PL::NPOS * PL::LDM; e += kThreads) {
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
    for (int e = tid; e < PL::NPOS; e += kThreads) {
        const int c = PL::col_of(e, R);
        a_s[e] = c >= 0 ? a.A[(size_t)g_slice * R + c] : 0.0;
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const double rg = a.rho[g_slice];
    double sc[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        sc[b][0] = a_s[8 * b + 2 * t];
        sc[b][1] = a_s[8 * b + 2 * t + 1];
    }
    GA accS, accB;
    accS.clear();
    if (LAST) accB.clear();
    double gap_d2 = 0.0, gap_x2 = 0.0, gap_ab = 0.0;
    double* tileG = gtiles + (size_t)warp * 8 * GA::LDT;
    const uint32_t lane_row_off = (uint32_t)((warp * 8 + g) * R * sizeof(double));

    auto body = [&](auto full_c, int tile, int s) {
        constexpr bool FULL = decltype(full_c)::value;
        const long long row = r_begin + (long long)tile * kTile + warp * 8 + g;
        const bool valid = FULL || row < r_end;
        const uint32_t st = ring_s + (uint32_t)s * STAGE + lane_row_off;
        double y[NB][2], v[NB][2], pd[NB][2], dpf[NB][2], sh[NB][2], du[NB][2];
        ld_row<PL, FULL>(st, t, valid, y);
        ld_row<PL, FULL>(st + ARR, t, valid, v);
        mma_rowmat<PL>(v, Ts, g, t, pd);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                dpf[b][e] = v[b][e] - pd[b][e];   // dual = V - P Delta          (decomposition.py:282-285)
                sh[b][e] = pd[b][e] - dpf[b][e];  // aux - dual = P Delta - dual (penalties.py:1280-1281)
            }
        if constexpr (K1 >= 0) {
            if constexpr (TIN) {
                // T-only state: aux = prox(T), dual = T - aux recomputed (bit-identical to the stored pair)
                double tt[NB][2];
                ld_row<PL, FULL>(st + 2 * ARR, t, valid, tt);
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double z = prox_elem<double>(tt[b][e], K1, 0, 0.0, 0.0, rg);
                        du[b][e] = tt[b][e] - z;
                        sh[b][e] += z - du[b][e];
                    }
            } else {
                double ax[NB][2];
                ld_row<PL, FULL>(st + 2 * ARR, t, valid, ax);
                ld_row<PL, FULL>(st + 3 * ARR, t, valid, du);
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) sh[b][e] += ax[b][e] - du[b][e];
            }
        }
        double sv[NB][2], xv[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) sv[b][e] = fma(rg, sh[b][e], y[b][e] * sc[b][e]);
        mma_rowmat<PL>(sv, Ms, g, t, xv);
        {
            // x depends on every value loaded from the stage.  Once it exists in all lanes (warp vote on its bits) the
            // warp has released the stage; the LAST warp to do so refills it with the tile `stages` ahead.
            int dep = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) dep = max(dep, max(dep_bits_of(xv[b][0]), dep_bits_of(xv[b][1])));
            const unsigned never = __any_sync(0xffffffffu, dep == 0x7ff7a5a5) ? 1u : 0u;
            if (lane == 0) {
                const unsigned old = atom_inc_acqrel(cnt + s + never);
                if (old == (unsigned)(kWarps - 1)) {
                    cnt[s] = 0u;
                    if (tile + stages < n_tiles) issue(tile + stages, s);
                }
            }
        }
        double vn[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) vn[b][e] = valid ? xv[b][e] + dpf[b][e] : 0.0;  // V' = x + dual_pf2
        const size_t goff = (size_t)row * R;
        if (valid) {
            st_row<PL>(a.pf_dual + goff, t, vn);
            if constexpr (LAST) {
                st_row<PL>(a.x_out + goff, t, xv);
                double wv[NB][2];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) wv[b][e] = xv[b][e] * sc[b][e];
                st_row<PL>(a.w_out + (size_t)row * a.ldw, t, wv);
            }
            if constexpr (K1 >= 0) {
                double zo[NB][2], dn[NB][2];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double vv = xv[b][e] + du[b][e];
                        const double z = prox_elem<double>(vv, K1, 0, 0.0, 0.0, rg);
                        zo[b][e] = z;
                        dn[b][e] = TOUT ? vv : vv - z;
                        if constexpr (LAST) {
                            // feasibility-gap terms of the companion (decomposition.py:406-415) while aux and x are in
                            // registers; padding positions hold x = aux = 0
                            const double df = z - xv[b][e];
                            gap_d2 = fma(df, df, gap_d2);
                            gap_x2 = fma(xv[b][e], xv[b][e], gap_x2);
                            gap_ab += fabs(xv[b][e]);
                        }
                    }
                if constexpr (!TOUT) st_row<PL>(a.c_aux + goff, t, zo);
                st_row<PL>(a.c_dual + goff, t, dn);
            }
        }
        accS.add(vn, tileG, g, t, LAST ? accB.dep() : 0);
        if constexpr (LAST) {
            double xz[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) xz[b][e] = valid ? xv[b][e] : 0.0;
            accB.add(xz, tileG, g, t, accS.dep());
        }
    };

    int s = 0;
    uint32_t ph = 0;
    const int n_full = (int)((r_end - r_begin) / kTile);
    for (int tile = 0; tile < n_tiles; ++tile) {
        mbar_wait(&full[s], ph);
        if (tile < n_full)
            body(std::true_type{}, tile, s);
        else
            body(std::false_type{}, tile, s);
        if (++s == stages) {
            s = 0;
            ph ^= 1;
        }
    }
    // cross-warp reduction of the Gram partials; the ring is free (every issued tile has been consumed by every warp)
    __syncthreads();
    double* red = (double*)ring;
    auto sync = [] { __syncthreads(); };
    gram_reduce_store<PL, double>(accS, red, warp, lane, kWarps, tid, kThreads, R, a.S_out + (size_t)g_slice * RR, sync);
    if (LAST) {
        __syncthreads();
        gram_reduce_store<PL, double>(accB, red, warp, lane, kWarps, tid, kThreads, R, a.BtB_out + (size_t)g_slice * RR,
                                      sync);
        if (K1 >= 0 && a.stats_part != nullptr) {  // fixed-order reduction: lanes (shuffle tree), then warps 0..7
            gap_d2 = warp_sum(gap_d2);
            gap_x2 = warp_sum(gap_x2);
            gap_ab = warp_sum(gap_ab);
            __syncthreads();
            if (lane == 0) {
                red[3 * warp] = gap_d2;
                red[3 * warp + 1] = gap_x2;
                red[3 * warp + 2] = gap_ab;
            }
            __syncthreads();
            if (tid < 3) {
                double sum = 0.0;
                for (int w = 0; w < kWarps; ++w) sum += red[3 * w + tid];
                a.stats_part[(size_t)g_slice * 3 + tid] = sum;
            }
        }
    }
}

template <int NBF, int HALF, int K1, bool TIN, bool TOUT, bool LAST>
int launch(const Args& a0, int n_groups, cudaStream_t st) {
    using PL = PosLayout<NBF, HALF>;
    using GA = GramAcc<PL>;
    constexpr int R = 8 * NBF + 4 * HALF;
    constexpr int NIN = 2 + (K1 >= 0 ? (TIN ? 1 : 2) : 0);
    const size_t fixed = (size_t)(2 * PL::NPOS * PL::LDM + kWarps * 8 * GA::LDT + PL::NPOS) * sizeof(double) +
                         kMaxStages * sizeof(unsigned) + 128;
    const size_t stage_bytes = (size_t)NIN * kTile * R * sizeof(double);
    const size_t red_bytes = (size_t)kWarps * GA::NPAIR * 64 * sizeof(double);
    // two CTAs per SM up to three column blocks: (228 KB - 2 x 1 KB reserved) / 2 = 113 KB each
    const size_t budget = (PL::NB <= 3 ? 113 : 226) * 1024;
    if (fixed + 2 * stage_bytes + 64 > budget) return -1;
    int stages = (int)((budget - fixed - 64) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    size_t ring_bytes = (size_t)stages * stage_bytes;
    if (ring_bytes < red_bytes) ring_bytes = red_bytes;
    const size_t smem = fixed + ring_bytes + (size_t)stages * sizeof(uint64_t) + 64;
    if (smem > 227 * 1024) return -1;
    Args a = a0;
    a.stages = stages;
    auto kern = pf2_rowpass_v2_kernel<NBF, HALF, K1, TIN, TOUT, LAST>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<n_groups, kThreads, smem, st>>>(a);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <int NBF, int HALF, int K1>
int launch_pass(const Args& a, int n_groups, bool tin, bool tout, bool last, cudaStream_t st) {
    // the three passes of a B-update: first (explicit aux/dual in, T out), middle (T in, T out), last (T in, explicit
    // out + x, W, B^T B); without a companion the T bits are meaningless and only `last` distinguishes the passes
    if (K1 < 0) {
        return last ? launch<NBF, HALF, K1, true, false, true>(a, n_groups, st)
                    : launch<NBF, HALF, K1, true, true, false>(a, n_groups, st);
    }
    if (!tin && tout && !last) return launch<NBF, HALF, K1, false, true, false>(a, n_groups, st);
    if (tin && tout && !last) return launch<NBF, HALF, K1, true, true, false>(a, n_groups, st);
    if (tin && !tout && last) return launch<NBF, HALF, K1, true, false, true>(a, n_groups, st);
    if (!tin && !tout && last) return launch<NBF, HALF, K1, false, false, true>(a, n_groups, st);  // inner_n_iter_max = 1
    // the companion stays ONE array across outer iterations too (T in the dual slot; the engine materialises
    // (aux, dual) on demand and takes the gap terms from stats_part)
    if (tin && tout && last) return launch<NBF, HALF, K1, true, true, true>(a, n_groups, st);
    if (!tin && tout && last) return launch<NBF, HALF, K1, false, true, true>(a, n_groups, st);
    return -1;
}

// Returns B2_OK after launching, a positive error code, or -1 when this specialisation does not apply.
inline int try_launch(const int64_t* row_off, int n_groups, int R, const void* Y, const void* A, const void* rho,
                      const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta, void* x,
                      void* w_out, int ldw, void* S_out, void* BtB_out, double* stats_part, cudaStream_t st) {
    if (!(deferred & 1) || R % 4 != 0 || R < 4 || R > 32) return -1;
    const int n_extra = pa.n_pen - 1;
    if (n_extra > 1) return -1;
    if (n_extra == 1 && pa.kind[1] != B2_PEN_NONNEG) return -1;
    const bool last = x != nullptr;
    if (last != (w_out != nullptr) || last != (BtB_out != nullptr)) return -1;
    if (last && (ldw % 2 != 0 || ((uintptr_t)w_out) % 16 != 0)) return -1;
    const bool tin = (deferred & 2) != 0, tout = (deferred & 4) != 0;
    Args a{};
    a.row_off = row_off;
    int n_in = 0;
    a.in[n_in++] = (const double*)Y;
    a.in[n_in++] = (const double*)pa.dual[0];
    if (n_extra == 1) {
        if (!tin) a.in[n_in++] = (const double*)pa.aux[1];
        a.in[n_in++] = (const double*)pa.dual[1];
        a.c_aux = (double*)pa.aux[1];
        a.c_dual = (double*)pa.dual[1];
    }
    for (int i = 0; i < n_in; ++i)
        if (((uintptr_t)a.in[i]) % 16 != 0) return -1;
    a.A = (const double*)A;
    a.rho = (const double*)rho;
    a.Minv = (const double*)Minv;
    a.Wmat = (const double*)Wmat;
    a.Delta = (const double*)Delta;
    a.pf_dual = (double*)pa.dual[0];
    a.x_out = (double*)x;
    a.w_out = (double*)w_out;
    a.ldw = ldw;
    a.S_out = (double*)S_out;
    a.BtB_out = (double*)BtB_out;
    a.stats_part = stats_part;
    const int NBF = R / 8, HALF = (R % 8) ? 1 : 0;
#define B2_RP2_CASE(F, H)                                                                              \
    if (NBF == F && HALF == H) {                                                                       \
        if (n_extra == 0) return launch_pass<F, H, -1>(a, n_groups, tin, tout, last, st);              \
        return launch_pass<F, H, B2_PEN_NONNEG>(a, n_groups, tin, tout, last, st);                     \
    }
    B2_RP2_CASE(0, 1)
    B2_RP2_CASE(1, 0)
    B2_RP2_CASE(1, 1)
    B2_RP2_CASE(2, 0)
    B2_RP2_CASE(2, 1)
    B2_RP2_CASE(3, 0)
    B2_RP2_CASE(3, 1)
    B2_RP2_CASE(4, 0)
#undef B2_RP2_CASE
    return -1;
}

}  // namespace rp2
