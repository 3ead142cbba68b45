// Single-read fused X-stream pass for slice-local coupled matrix factorization (SURVEY.md §7 hard part 1, §8b
// `b2_xstream_fused_local`): ONE kernel reads every slice X_i once from HBM per outer iteration and performs
//
//     Y_i = X_i C                                   (rhs of the B-update, decomposition.py:242:  X_i (C o a_i) = Y_i o a_i)
//     B_i <- inner ADMM loop of admm_update_B       (decomposition.py:259-289, row-local penalties: NonNegativity, Box, L1)
//     G_i = X_i^T B_i                               (K x R per slice; Z = sum_i G_i diag(a_i), the rhs of the C-update
//                                                    decomposition.py:312-315, is summed from the stored G_i right after)
//     B_i^T B_i                                     (lhs of the C-update :310 and cross products of the A-update :138-143)
//
// The A-update of the same outer iteration (which needs diag(B_i^T X_i C_new), decomposition.py:145-158) is served
// WITHOUT another pass over X from the stored G_i:  diag(B_i^T X_i C_new)[r] = sum_k G_i[k][r] C_new[k][r]
// (b2_slice_gdot).  The reference reads X three times per outer iteration, the two-pass schedule of xstream.cu twice.
//
// Work decomposition: persistent grid, one CTA per SM, every slice is owned by ONE CTA (host-side balanced schedule),
// so G_i and B_i^T B_i need no cross-CTA reduction.  A slice is processed in chunks of 128 rows; per chunk
//   Y-phase  the producer warp streams X[chunk, :] through the TMA ring (stage = 8 swizzled boxes [32 rows x 16
//            doubles]: 4 row quarters x 2 k-halves, 32 k's per stage; loads marked L2 evict_last); consumer warp (rq, kh)
//            contracts its box with DMMA.8x8x4 against the factor matrix C resident in shared memory; the 2 k-half
//            partials meet in shared memory
//   B-update every consumer warp owns 2 x 8 rows of the chunk in the MMA accumulator layout and runs the whole inner
//            loop in registers (same formulation as admm_mma.cu; the two row blocks are independent DMMA chains, the
//            fragments of Minv_g stay in registers), writes x / aux / dual to HBM and stages the new rows in shared
//            memory (the tile B_i^T B_i is accumulated from doubles as the operand of the Z-phase)
//   Z-phase  the producer re-streams the same 128 rows (stage = 8 boxes [32 rows x 16 doubles] = 128 k's, loads marked
//            evict_first; the chunk is 512 KB at K = 512, so this read is served by L2 — 148 CTAs keep 74 MB live, ncu:
//            5.0 GB of DRAM reads for 4.29 GB of X) and consumer warp w accumulates G_i[k-block + 16 w .. + 16][:] in
//            registers.
// The per-slice operands (Minv_g, a_g, rho_g) of the NEXT slice are fetched with cp.async while the current one runs.
// fp64 only (DMMA); the register-resident G_i accumulators bound K * ceil(R/8) <= 1024.  Measured and rejected: the
// B-update on two extra warps next to the contractions.
#include <stddef.h>

#include "admm_common.cuh"
#include "mma_tiles.cuh"
#include "xstream_common.cuh"

namespace {

constexpr int kCons = 8;  // consumer warps
constexpr int kThreadsF = (kCons + 1) * 32;
constexpr int kCR = 128;  // rows per chunk
constexpr int kBoxBytesF = 32 * 128;
constexpr int kStageBytesF = 8 * kBoxBytesF;

__device__ __forceinline__ void cons_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kCons * 32) : "memory"); }

struct FusedArgs {
    const double* Cp;        // [Kp x LDC] factor matrix staged by pad_c_kernel (MMA row order, zero padded)
    const int64_t* row_off;  // [n_slices + 1]
    const int32_t* sched;    // [gridDim.x x rounds] slice ids, -1 = none
    int rounds;
    const double* A;     // n_slices x R row scale a_i
    const double* rho;   // n_slices
    const double* Minv;  // n_slices x R x R
    PenArgs pa;
    int n_inner, R, K, Kp;
    double* B;      // N x R (out)
    double* G;      // n_slices x K x R (out)
    double* BtB;    // n_slices x R x R (out)
    int stages;
};

__device__ __forceinline__ double2 lds_f64x2(uint32_t saddr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}
// volatile: the B-state rows are requested where the call stands (ahead of the Y-phase), not where they are used
__device__ __forceinline__ double2 ldg_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <class PL>
__device__ __forceinline__ void load_row_g(const double* __restrict__ base, long long row, int R, int t, bool valid,
                                           double (&v)[PL::NB][2]) {
    const bool vec = (R & 1) == 0;
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        v[b][0] = v[b][1] = 0.0;
        if (valid) {
            if (vec && c0 >= 0 && c1 >= 0) {
                const double2 pr = ldg_f64x2(base + (size_t)row * R + c0);
                v[b][0] = pr.x;
                v[b][1] = pr.y;
            } else {
                if (c0 >= 0) v[b][0] = base[(size_t)row * R + c0];
                if (c1 >= 0) v[b][1] = base[(size_t)row * R + c1];
            }
        }
    }
}

template <class PL>
__device__ __forceinline__ void store_row_g(double* __restrict__ base, long long row, int R, int t,
                                            const double (&v)[PL::NB][2]) {
    const bool vec = (R & 1) == 0;
#pragma unroll
    for (int b = 0; b < PL::NB; ++b) {
        const int c0 = reg_col<PL>(b, 0, t, R), c1 = reg_col<PL>(b, 1, t, R);
        if (vec && c0 >= 0 && c1 >= 0) {
            *(double2*)(base + (size_t)row * R + c0) = make_double2(v[b][0], v[b][1]);
        } else {
            if (c0 >= 0) base[(size_t)row * R + c0] = v[b][0];
            if (c1 >= 0) base[(size_t)row * R + c1] = v[b][1];
        }
    }
}

// per-slice operands of the B-update (Minv_g in MMA position order, a_g, rho_g), double buffered: the next slice's set is
// fetched with cp.async while the current slice is processed
// TMA tile load with an L2 eviction-priority hint: the Y-phase read of a chunk is marked evict_last (the Z-phase reads
// the same rows a few microseconds later), the Z-phase read evict_first (the rows are dead afterwards)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

template <class PL>
struct SliceOps {
    double Ms[PL::NPOS * PL::LDM];
    double a[PL::NPOS];
    double rho[2];
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <class PL>
__device__ __forceinline__ void fetch_slice_ops(SliceOps<PL>* so, const FusedArgs& fa, int g, int tid) {
    const int R = fa.R;
    for (int e = tid; e < R * R; e += kCons * 32) {  // PosLayout<NBLK, 0>: position == column, pad rows/columns stay 0
        const int r = e / R, c = e - r * R;
        cp_async8(&so->Ms[r * PL::LDM + c], fa.Minv + (size_t)g * R * R + e);
    }
    if (tid < R) cp_async8(&so->a[tid], fa.A + (size_t)g * R + tid);
    if (tid == 32) cp_async8(&so->rho[0], fa.rho + g);
}

template <int NBLK, int KB, int NP>
__global__ void __launch_bounds__(kThreadsF, 1)
xfused_local_kernel(const __grid_constant__ CUtensorMap tmap_x, const FusedArgs fa) {
    using PL = PosLayout<NBLK, 0>;
    using GA = GramAcc<PL>;
    using SO = SliceOps<PL>;
    constexpr int NB = NBLK;
    constexpr int LDC = 8 * NBLK + 4, LDR = 8 * NBLK + 2;
    constexpr int NPm = NP > 0 ? NP : 1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: ring [stages x 32 KB] | Cp_s [Kp x LDC] | red [2 x 128 x LDR] | slice operands x 2 | full | empty
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stages = fa.stages, Kp = fa.Kp, R = fa.R, K = fa.K;
    double* Cp_s = (double*)(base + (size_t)stages * kStageBytesF);
    double* red = Cp_s + (size_t)Kp * LDC;
    SO* sops = (SO*)(red + 2 * kCR * LDR);
    uint64_t* full = (uint64_t*)(sops + 2);
    uint64_t* empty = full + stages;
    const uint32_t base_s = smem_u32(base);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kCons);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int nks = Kp / 32, nkb = Kp / 128;
    const int32_t* my_sched = fa.sched + (size_t)blockIdx.x * fa.rounds;

    if (warp == kCons) {
        // ===== producer warp: one elected lane issues all TMA traffic =====
        if (lane == 0) {
            prefetch_tmap(&tmap_x);
            const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < fa.rounds; ++i) {
                const int g = my_sched[i];
                if (g < 0) continue;
                const long long r_begin = fa.row_off[g], r_end = fa.row_off[g + 1];
                for (long long row0 = r_begin; row0 < r_end; row0 += kCR) {
                    const long long left = r_end - row0;
                    const int nrq = left >= kCR ? 4 : (int)((left + 31) >> 5);  // 32-row quarters in this chunk
                    for (int ks = 0; ks < nks; ++ks) {  // Y-phase: box (rq, kh) = rows 32 rq.., k's 32 ks + 16 kh..
                        mbar_wait(&empty[s], ph ^ 1);
                        unsigned char* st = base + (size_t)s * kStageBytesF;
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(nrq * 2 * kBoxBytesF));
                        for (int rq = 0; rq < nrq; ++rq)
#pragma unroll
                            for (int kh = 0; kh < 2; ++kh)
                                tma_load_2d_hint(st + (rq * 2 + kh) * kBoxBytesF, &tmap_x, ks * 32 + kh * 16,
                                                 (int)(row0 + rq * 32), &full[s], pol_keep);
                        if (++s == stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    for (int rt = 0; rt < nrq; ++rt) {  // Z-phase: box = 16 k's of a 128-k block, rows 32 rt..
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(&empty[s], ph ^ 1);
                            unsigned char* st = base + (size_t)s * kStageBytesF;
                            mbar_arrive_expect_tx(&full[s], (uint32_t)kStageBytesF);
#pragma unroll
                            for (int b = 0; b < 8; ++b)
                                tma_load_2d_hint(st + b * kBoxBytesF, &tmap_x, kb * 128 + b * 16,
                                                 (int)(row0 + rt * 32), &full[s], pol_drop);
                            if (++s == stages) {
                                s = 0;
                                ph ^= 1;
                            }
                        }
                    }
                }
            }
        }
        return;
    }

    // ===== consumers =====
    for (int e = tid; e < Kp * LDC; e += kCons * 32) Cp_s[e] = fa.Cp[e];
    for (int e = tid; e < (int)(2 * sizeof(SO) / sizeof(double)); e += kCons * 32) ((double*)sops)[e] = 0.0;
    const int g_ = lane >> 2, t = lane & 3;
    const int kh = warp & 1, rq = warp >> 1;
    const uint32_t a_chunk0 = 4u * (t >> 1), a_half = (t & 1) * 8u;
    const uint32_t cp_addr = smem_u32(Cp_s) + (uint32_t)(g_ * sizeof(double));
    const uint32_t red_addr = smem_u32(red);
    const uint32_t sops_addr = smem_u32(sops);
    const uint32_t wt_addr = red_addr + (uint32_t)(g_ * sizeof(double));  // the new rows of B_i live in partial 0 of `red`
    const int RR = R * R;
    const int iters = NP == 0 ? 1 : fa.n_inner;
    int s = 0;
    uint32_t ph = 0;
    int buf = 0;
    // first non-empty slice of this CTA: fetch its operands
    int i_next = 0;
    auto advance_next = [&]() {
        while (i_next < fa.rounds) {
            const int g = my_sched[i_next];
            if (g >= 0 && fa.row_off[g + 1] > fa.row_off[g]) break;
            ++i_next;
        }
    };
    advance_next();
    cons_barrier();  // the zero fill of the operand buffers is complete
    if (i_next < fa.rounds) fetch_slice_ops<PL>(&sops[0], fa, my_sched[i_next], tid);

    for (int i = 0; i < fa.rounds; ++i) {
        const int g = my_sched[i];
        if (g < 0) continue;
        const long long r_begin = fa.row_off[g], r_end = fa.row_off[g + 1];
        double* Gg = fa.G + (size_t)g * K * R;
        if (r_begin >= r_end) {  // empty slice: zero products
            for (int e = tid; e < RR; e += kCons * 32) fa.BtB[(size_t)g * RR + e] = 0.0;
            for (int e = tid; e < K * R; e += kCons * 32) Gg[e] = 0.0;
            continue;
        }
        cp_async_wait_all();  // this slice's operands (requested one slice ago) have landed
        cons_barrier();       // ... in every thread; and every warp is done with the previous slice's scratch
        const uint32_t so_addr = sops_addr + (uint32_t)(buf * sizeof(SO));
        i_next = i + 1;
        advance_next();
        if (i_next < fa.rounds) fetch_slice_ops<PL>(&sops[buf ^ 1], fa, my_sched[i_next], tid);
        buf ^= 1;
        const double rg = lds_f64(so_addr + (uint32_t)offsetof(SO, rho));
        double sc[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double2 a2 = lds_f64x2(so_addr + (uint32_t)(offsetof(SO, a) + (8 * b + 2 * t) * sizeof(double)));
            sc[b][0] = a2.x;
            sc[b][1] = a2.y;
        }
        GA accB;
        accB.clear();
        double Gacc[KB][2][NBLK][2];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NBLK; ++n) Gacc[kb][m][n][0] = Gacc[kb][m][n][1] = 0.0;

        for (long long row0 = r_begin; row0 < r_end; row0 += kCR) {
            const long long left = r_end - row0;
            const int nrq = left >= kCR ? 4 : (int)((left + 31) >> 5);
            // this lane's two rows in the B-update (m-blocks 2 warp, 2 warp + 1 of the chunk)
            const long long brow0 = row0 + warp * 16 + g_, brow1 = brow0 + 8;
            const bool valid0 = brow0 < r_end, valid1 = brow1 < r_end;
            // the B-state rows of the B-update are requested now and arrive behind the Y-phase
            double ax[2][NPm][NB][2], du[2][NPm][NB][2];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                load_row_g<PL>((const double*)fa.pa.aux[p], brow0, R, t, valid0, ax[0][p]);
                load_row_g<PL>((const double*)fa.pa.dual[p], brow0, R, t, valid0, du[0][p]);
                load_row_g<PL>((const double*)fa.pa.aux[p], brow1, R, t, valid1, ax[1][p]);
                load_row_g<PL>((const double*)fa.pa.dual[p], brow1, R, t, valid1, du[1][p]);
            }

            // ---- Y-phase: acc[m][n] = partial of Y[row0 + 32 rq + 8 m + g][8 n + 2 t (+1)] over k-half kh ----
            double acc[4][NBLK][2];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NBLK; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
            for (int ks = 0; ks < nks; ++ks) {
                mbar_wait(&full[s], ph);
                if (rq < nrq) {
                    const uint32_t box = base_s + (uint32_t)s * kStageBytesF + (uint32_t)(rq * 2 + kh) * kBoxBytesF;
                    const uint32_t crow0 = cp_addr + (uint32_t)((ks * 32 + kh * 16 + t) * LDC * sizeof(double));
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        double a[4], bf[NBLK];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            a[m] = lds_f64(box + swz128((uint32_t)(8 * m + g_), a_chunk0 + k4) + a_half);
                        const uint32_t crow = crow0 + (uint32_t)(k4 * 4 * LDC * sizeof(double));
#pragma unroll
                        for (int n = 0; n < NBLK; ++n) bf[n] = lds_f64(crow + n * 64);
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int n = 0; n < NBLK; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[m], bf[n]);
                    }
                }
                int dep = 0;
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < NBLK; ++n) dep = max(dep, dep_bits_of(acc[m][n][0]));
                stage_release(&empty[s], lane, dep);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
            cons_barrier();  // every warp has left the previous chunk's Z-phase, which reads partial 0 of `red`
            // k-half partials -> shared memory; quarters past the end of a short chunk are zeros
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < NBLK; ++n)
                    *(double2*)(red + (size_t)(kh * kCR + 32 * rq + 8 * m + g_) * LDR + 8 * n + 2 * t) =
                        make_double2(acc[m][n][0], acc[m][n][1]);
            cons_barrier();
            double r_[2][NB][2], xv[2][NB][2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const uint32_t row = red_addr + (uint32_t)(((16 * warp + 8 * mb + g_) * LDR + 2 * t) * sizeof(double));
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const double2 y0 = lds_f64x2(row + b * 64);
                    const double2 y1 = lds_f64x2(row + (uint32_t)(kCR * LDR * sizeof(double)) + b * 64);
                    r_[mb][b][0] = (y0.x + y1.x) * sc[b][0];
                    r_[mb][b][1] = (y0.y + y1.y) * sc[b][1];
                }
            }
            __syncwarp();  // the Gram tiles alias this warp's rows of partial 0: all lanes have read them

            // ---- B-update: the whole inner loop on this warp's 2 x 8 rows (decomposition.py:259-289) ----
            // B fragments of Minv_g for all k-steps: loaded once, reused by the n_inner iterations and both row blocks
            double mf[PL::KS][NB];
#pragma unroll
            for (int ks = 0; ks < PL::KS; ++ks)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                    mf[ks][nb] = lds_f64(so_addr + (uint32_t)(((8 * (ks >> 1) + 2 * t + (ks & 1)) * PL::LDM + g_ + 8 * nb) *
                                                             sizeof(double)));
            for (int it = 0; it < iters; ++it) {
                double sv[2][NB][2];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double sh = 0.0;
#pragma unroll
                            for (int p = 0; p < NP; ++p) sh += ax[mb][p][b][e] - du[mb][p][b][e];
                            sv[mb][b][e] = NP > 0 ? fma(rg, sh, r_[mb][b][e]) : r_[mb][b][e];
                        }
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) xv[mb][nb][0] = xv[mb][nb][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < PL::KS; ++ks)  // x = s Minv_g: the two row blocks are independent chains
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        dmma884(xv[0][nb][0], xv[0][nb][1], sv[0][ks >> 1][ks & 1], mf[ks][nb]);
                        dmma884(xv[1][nb][0], xv[1][nb][1], sv[1][ks >> 1][ks & 1], mf[ks][nb]);
                    }
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int kind = fa.pa.kind[p], nn = fa.pa.nn[p];
                    const double p0 = fa.pa.p0[p], p1 = fa.pa.p1[p];
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const double vv = xv[mb][b][e] + du[mb][p][b][e];
                                const double z = prox_elem<double>(vv, kind, nn, p0, p1, rg);
                                ax[mb][p][b][e] = z;
                                du[mb][p][b][e] = vv - z;
                            }
                }
            }
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const bool valid = mb ? valid1 : valid0;
                const long long brow = mb ? brow1 : brow0;
                double xz[NB][2];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool pad = reg_col<PL>(b, e, t, R) < 0;
                        xz[b][e] = (valid && !pad) ? xv[mb][b][e] : 0.0;
                    }
                if (valid) {
                    store_row_g<PL>(fa.B, brow, R, t, xv[mb]);
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        store_row_g<PL>((double*)fa.pa.aux[p], brow, R, t, ax[mb][p]);
                        store_row_g<PL>((double*)fa.pa.dual[p], brow, R, t, du[mb][p]);
                    }
                }
                // B_i^T B_i; the tile the rows are staged in (zero rows past the end of the slice, zero pad columns)
                // is at the same time the operand of the Z-phase
                accB.add(xz, red + (size_t)(16 * warp + 8 * mb) * LDR, g_, t);
            }
            cons_barrier();

            // ---- Z-phase: Gacc[kb][m][n] += X[rows, k]^T x[rows, :]  for k = 128 kb + 16 warp + 8 m + g ----
            // (contracts with the UNSCALED rows: G_i = X_i^T B_i also serves the A-update; a_i is applied when Z is
            // summed from the G_i — dividing X_i^T (B_i o a_i) by a_i afterwards would hit the zeros of a clipped a_i)
            for (int rt = 0; rt < nrq; ++rt) {
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    if (kb < nkb) {
                        mbar_wait(&full[s], ph);
                        const uint32_t box = base_s + (uint32_t)s * kStageBytesF + (uint32_t)warp * kBoxBytesF;
#pragma unroll
                        for (int rgi = 0; rgi < 4; ++rgi) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t row = rgi * 8 + h + 2 * t;
                                double a[2], bf[NBLK];
#pragma unroll
                                for (int m = 0; m < 2; ++m) {
                                    const uint32_t kk = 8 * m + g_;
                                    a[m] = lds_f64(box + swz128(row, kk >> 1) + (kk & 1) * 8);
                                }
                                const uint32_t wrow = wt_addr + (uint32_t)((rt * 32 + row) * LDR * sizeof(double));
#pragma unroll
                                for (int n = 0; n < NBLK; ++n) bf[n] = lds_f64(wrow + n * 64);
#pragma unroll
                                for (int m = 0; m < 2; ++m)
#pragma unroll
                                    for (int n = 0; n < NBLK; ++n)
                                        dmma884(Gacc[kb][m][n][0], Gacc[kb][m][n][1], a[m], bf[n]);
                            }
                        }
                        int dep = 0;
#pragma unroll
                        for (int m = 0; m < 2; ++m)
#pragma unroll
                            for (int n = 0; n < NBLK; ++n) dep = max(dep, dep_bits_of(Gacc[kb][m][n][0]));
                        stage_release(&empty[s], lane, dep);
                        if (++s == stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }

        // ---- end of slice: G_i -> HBM (fire-and-forget stores), B_i^T B_i -> HBM ----
        const bool vec = (R & 1) == 0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            if (kb < nkb) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int k = kb * 128 + warp * 16 + 8 * m + g_;
                    if (k < K) {
#pragma unroll
                        for (int n = 0; n < NBLK; ++n) {
                            const int c = 8 * n + 2 * t;
                            const size_t o = (size_t)k * R + c;
                            if (vec) {
                                if (c < R) *(double2*)(Gg + o) = make_double2(Gacc[kb][m][n][0], Gacc[kb][m][n][1]);
                            } else {
                                if (c < R) Gg[o] = Gacc[kb][m][n][0];
                                if (c + 1 < R) Gg[o + 1] = Gacc[kb][m][n][1];
                            }
                        }
                    }
                }
            }
        }
        cons_barrier();  // every warp has left the last Z-phase: `red` becomes the reduction scratch
        gram_reduce_store<PL, double>(accB, red, warp, lane, kCons, tid, kCons * 32, R, fa.BtB + (size_t)g * RR,
                                      cons_barrier);
    }
    cp_async_wait_all();
}

// rhs[g][c] = sum_k G[g][k][c] * C[k][c]   (= diag(B_g^T X_g C), decomposition.py:145-158, from G_g = X_g^T B_g)
__global__ void __launch_bounds__(256) slice_gdot_kernel(const double* __restrict__ G, const double* __restrict__ C,
                                                         int K, int R, double* __restrict__ rhs) {
    __shared__ double part[256];
    const int g = blockIdx.x;
    const int lanes = (256 / R) * R;  // threads in use: a multiple of R, so a thread's column is fixed
    const double* Gg = G + (size_t)g * K * R;
    double acc = 0.0;
    if ((int)threadIdx.x < lanes)
        for (int e = threadIdx.x; e < K * R; e += lanes) acc = fma(Gg[e], C[e], acc);
    part[threadIdx.x] = (int)threadIdx.x < lanes ? acc : 0.0;
    __syncthreads();
    if ((int)threadIdx.x < R) {
        double sum = 0.0;
        for (int q = threadIdx.x; q < lanes; q += R) sum += part[q];
        rhs[(size_t)g * R + threadIdx.x] = sum;
    }
}

// Zpart[s][e] = sum over the slices i of slab s of G[i][e] * A[i][e % R]   (Z = sum_i G_i diag(a_i), fixed order)
__global__ void __launch_bounds__(256) gsum_z_kernel(const double* __restrict__ G, const double* __restrict__ A,
                                                     int n_slices, int KR, int R, int per_slab,
                                                     double* __restrict__ Zpart) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= KR) return;
    const int c = e % R;
    const int i0 = blockIdx.y * per_slab, i1 = min(n_slices, i0 + per_slab);
    double acc = 0.0;
    int i = i0;
    for (; i + 4 <= i1; i += 4) {  // 4 independent loads in flight
        const double g0 = G[(size_t)i * KR + e], g1 = G[(size_t)(i + 1) * KR + e];
        const double g2 = G[(size_t)(i + 2) * KR + e], g3 = G[(size_t)(i + 3) * KR + e];
        acc = fma(g0, A[(size_t)i * R + c], acc);
        acc = fma(g1, A[(size_t)(i + 1) * R + c], acc);
        acc = fma(g2, A[(size_t)(i + 2) * R + c], acc);
        acc = fma(g3, A[(size_t)(i + 3) * R + c], acc);
    }
    for (; i < i1; ++i) acc = fma(G[(size_t)i * KR + e], A[(size_t)i * R + c], acc);
    Zpart[(size_t)blockIdx.y * KR + e] = acc;
}

struct FusedPlan {
    int NBLK, KB, Kp, stages;
    size_t smem;
};

// 0 = not applicable
int fused_plan(int K, int R, int dtype, int n_pen, FusedPlan* pl) {
    if (dtype != B2_F64 || R < 1 || R > B2_MAX_RANK || K < 1 || n_pen < 0 || n_pen > 2) return 0;
    const int NBLK = (R + 7) / 8;
    const int Kp = ((K + 127) / 128) * 128;
    const int nkb = Kp / 128;
    int KB = 1;
    while (KB < nkb) KB *= 2;
    if (NBLK * KB > 8) return 0;  // register-resident G_i and Z accumulators: 8 NBLK KB doubles per thread
    const int LDC = 8 * NBLK + 4, LDR = 8 * NBLK + 2, NPOS = 8 * NBLK, LDM = 8 * NBLK + 2;
    const size_t fixed = 1024 + ((size_t)Kp * LDC + 2 * kCR * LDR + 2 * (NPOS * LDM + NPOS + 2)) * sizeof(double) + 64;
    const size_t budget = 227 * 1024;
    if (fixed + 2 * (kStageBytesF + 16) > budget) return 0;
    int stages = (int)((budget - fixed) / (kStageBytesF + 16));
    if (stages > 6) stages = 6;
    pl->NBLK = NBLK;
    pl->KB = KB;
    pl->Kp = Kp;
    pl->stages = stages;
    pl->smem = fixed + (size_t)stages * (kStageBytesF + 16);
    return 1;
}

template <int NBLK, int KB, int NP>
int launch_fused(const CUtensorMap& map, const FusedArgs& fa, int grid, size_t smem, cudaStream_t st) {
    auto kern = xfused_local_kernel<NBLK, KB, NP>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreadsF, smem, st>>>(map, fa);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <int NBLK, int KB>
int launch_fused_np(const CUtensorMap& map, const FusedArgs& fa, int n_pen, int grid, size_t smem, cudaStream_t st) {
    switch (n_pen) {
        case 0: return launch_fused<NBLK, KB, 0>(map, fa, grid, smem, st);
        case 1: return launch_fused<NBLK, KB, 1>(map, fa, grid, smem, st);
        default: return launch_fused<NBLK, KB, 2>(map, fa, grid, smem, st);
    }
}

}  // namespace

extern "C" {

int b2_xstream_fused_local_supported(int K, int R, int dtype, int n_pen) {
    FusedPlan pl;
    return fused_plan(K, R, dtype, n_pen, &pl);
}

size_t b2_xstream_fused_workspace_bytes(int K, int R, int n_ctas) {
    const size_t Kp = (size_t)((K + 127) / 128) * 128;
    const size_t LDC = 8 * (size_t)((R + 7) / 8) + 4;
    return Kp * LDC * sizeof(double) + (size_t)n_ctas * K * R * sizeof(double) + 512;
}

int b2_xstream_fused_local(const void* X, long long n_rows, int K, int ldx, const int64_t* row_off, int n_slices,
                           const int32_t* sched, int n_ctas, int rounds, const void* C, const void* A,
                           const void* rho, const void* Minv, const b2_penalty_desc* pens_host, int n_pen,
                           int n_inner, int R, void* B, void* Z, void* G, void* BtB, int dtype, void* ws,
                           size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FusedPlan pl;
    B2_REQUIRE(fused_plan(K, R, dtype, n_pen, &pl), "b2_xstream_fused_local does not apply to K=%d R=%d dtype=%d n_pen=%d "
               "(fp64, K * ceil(R/8) <= 1024, at most 2 row-local penalties)", K, R, dtype, n_pen);
    B2_REQUIRE(n_ctas >= 1 && n_ctas <= b2_num_sms() && rounds >= 0, "bad schedule shape (%d CTAs x %d rounds)", n_ctas, rounds);
    B2_REQUIRE(n_inner >= 1 || n_pen == 0, "n_inner must be >= 1");
    PenArgs pa;
    int rc = b2_pack_penalties(pens_host, n_pen, &pa);
    if (rc != B2_OK) return rc;
    for (int p = 0; p < n_pen; ++p)
        B2_REQUIRE(pa.kind[p] == B2_PEN_NONNEG || pa.kind[p] == B2_PEN_BOX || pa.kind[p] == B2_PEN_L1,
                   "penalty %d (kind %d) is not row-local", p, pa.kind[p]);
    B2_REQUIRE(ws_bytes >= b2_xstream_fused_workspace_bytes(K, R, n_ctas), "b2_xstream_fused_local workspace too small");
    B2_REQUIRE(((uintptr_t)ws) % 16 == 0 && ((uintptr_t)B) % 16 == 0 && ((uintptr_t)G) % 16 == 0,
               "buffers must be 16-byte aligned");
    if (n_slices == 0 || n_rows == 0 || rounds == 0) {
        B2_CHECK_CUDA(cudaMemsetAsync(Z, 0, (size_t)K * R * sizeof(double), st));
        return B2_OK;
    }
    const int LDC = 8 * pl.NBLK + 4;
    double* Cp = (double*)ws;
    const size_t cp_bytes = ((size_t)pl.Kp * LDC * sizeof(double) + 255) & ~(size_t)255;
    double* Zpart = (double*)((unsigned char*)ws + cp_bytes);
    pad_c_kernel<double><<<(pl.Kp * LDC + 255) / 256, 256, 0, st>>>((const double*)C, Cp, K, R, pl.Kp, LDC, 1);
    B2_LAUNCH_CHECK();
    alignas(64) CUtensorMap map;
    rc = encode_x_map(&map, X, n_rows, K, ldx, B2_F64, 32);
    if (rc != B2_OK) return rc;
    FusedArgs fa;
    fa.Cp = Cp;
    fa.row_off = row_off;
    fa.sched = sched;
    fa.rounds = rounds;
    fa.A = (const double*)A;
    fa.rho = (const double*)rho;
    fa.Minv = (const double*)Minv;
    fa.pa = pa;
    fa.n_inner = n_inner;
    fa.R = R;
    fa.K = K;
    fa.Kp = pl.Kp;
    fa.B = (double*)B;
    fa.G = (double*)G;
    fa.BtB = (double*)BtB;
    fa.stages = pl.stages;
    rc = B2_ERR_INVALID;
#define B2_F_CASE(NB, KBV) \
    if (pl.NBLK == NB && pl.KB == KBV) rc = launch_fused_np<NB, KBV>(map, fa, n_pen, n_ctas, pl.smem, st);
    B2_F_CASE(1, 1) B2_F_CASE(1, 2) B2_F_CASE(1, 4) B2_F_CASE(1, 8)
    B2_F_CASE(2, 1) B2_F_CASE(2, 2) B2_F_CASE(2, 4)
    B2_F_CASE(3, 1) B2_F_CASE(3, 2)
    B2_F_CASE(4, 1) B2_F_CASE(4, 2)
#undef B2_F_CASE
    if (rc != B2_OK) return rc;
    // Z = sum_i G_i diag(a_i): slabs of slices in parallel, then the fixed-order sum of the slab partials
    const int n = K * R;
    const int slabs = n_slices < n_ctas ? n_slices : n_ctas;
    const int per_slab = (n_slices + slabs - 1) / slabs;
    gsum_z_kernel<<<dim3((n + 255) / 256, slabs), 256, 0, st>>>((const double*)G, (const double*)A, n_slices, n, R,
                                                                 per_slab, Zpart);
    B2_LAUNCH_CHECK();
    reduce_partials_kernel<double><<<(n + 255) / 256, 256, 0, st>>>(Zpart, (double*)Z, n, slabs);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_slice_gdot(const void* G, const void* C, int n_groups, int K, int R, void* rhs, int dtype, void* stream) {
    B2_REQUIRE(dtype == B2_F64, "b2_slice_gdot is fp64 only");
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    slice_gdot_kernel<<<n_groups, 256, 0, (cudaStream_t)stream>>>((const double*)G, (const double*)C, K, R, (double*)rhs);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // extern "C"
