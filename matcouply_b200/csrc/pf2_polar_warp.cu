// PARAFAC2 polar step, one WARP per slice (reference penalties.py:1233-1235: P_i = U Vh of svd((B_i + dual_i) Delta^T)).
//
// With S_i = V_i^T V_i (R x R, from the row pass) the polar factor is P_i = V_i W_i,
//     W_i = Delta^T (Delta S_i Delta^T)^(-1/2),
// so per slice only R x R work is left: G = Delta S Delta^T, the cyclic Jacobi eigen-decomposition G = Q Lam Q^T
// (round-robin ordering, all R/2 disjoint rotations of a round applied together), H = Q Lam^-1/2 Q^T, W = Delta^T H
// and the summand num = rho W^T S = rho H (Delta S) of the Delta update (:1240-1245).
//
// Why a warp and not a CTA per slice (pf2_polar_cta_kernel in pf2_fused.cu, kept for A/B checks): a 20 x 20 Jacobi round
// is ~400 thread-operations, i.e. one or two instructions per thread of a 128-thread CTA between two __syncthreads.
// ncu (profiles/r1_ncu_c2_top_kernels.txt) showed that kernel at 50 % issue utilisation with the barrier as the top
// stall and 126 k warp-instructions per slice, most of them index arithmetic (runtime-R div/mod).  Here a warp owns
// the whole slice: the three R x R work matrices live in its private shared-memory region (odd leading dimension:
// both row- and column-wise 64-bit accesses are bank-conflict free), lanes own a row (column pass) or a column (row
// pass) so there is no index division anywhere, and the only synchronisation is __syncwarp.
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;

template <typename T>
__device__ __forceinline__ T warp_max_d(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// C[i][j] = sum_k a(i, k) * b(k, j) for i, j < R; lane = j, four rows i per pass (b(k, j) is loaded once per four rows).
template <class FA, class FB, class FC>
__device__ __forceinline__ void warp_mm(int R, int lane, FA a, FB b, FC store) {
    if (lane < R) {
        for (int i0 = 0; i0 < R; i0 += 4) {
            const int i1 = min(i0 + 1, R - 1), i2 = min(i0 + 2, R - 1), i3 = min(i0 + 3, R - 1);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int k = 0; k < R; ++k) {
                const double bk = b(k, lane);
                s0 = fma(a(i0, k), bk, s0);
                s1 = fma(a(i1, k), bk, s1);
                s2 = fma(a(i2, k), bk, s2);
                s3 = fma(a(i3, k), bk, s3);
            }
            store(i0, lane, s0);
            if (i0 + 1 < R) store(i0 + 1, lane, s1);
            if (i0 + 2 < R) store(i0 + 2, lane, s2);
            if (i0 + 3 < R) store(i0 + 3, lane, s3);
        }
    }
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
pf2_polar_warp_kernel(const T* __restrict__ S, const T* __restrict__ Delta, const T* __restrict__ rho, int n_groups,
                      int R, T* __restrict__ Wmat, double* __restrict__ num_part, double* __restrict__ Qstore,
                      int warm) {
    extern __shared__ __align__(16) double pw_smem[];
    const int LD = R | 1, MS = (R * LD + 1) & ~1, RR = R * R;  // even matrix stride: the {c, s} pairs stay 16-byte aligned
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* Dm = pw_smem;
    double* G = pw_smem + MS + (size_t)w * (3 * MS + 48);
    double* Q = G + MS;
    double* Tm = Q + MS;
    double* cs = Tm + MS;          // {c, s}[16] (one 128-bit broadcast load per pair); later lam^-1/2 [32]
    double2* cs2 = (double2*)cs;
    int* pq = (int*)(cs + 32);     // p | q << 16 [16]
    for (int e = threadIdx.x; e < R * 32; e += blockDim.x) {
        const int i = e >> 5, j = e & 31;
        if (j < R) Dm[i * LD + j] = (double)Delta[i * R + j];
    }
    __syncthreads();
    const bool act = lane < R;
    const int Re = R + (R & 1), m = Re - 1, half = Re / 2;

    for (int g = blockIdx.x * kWarpsPerCta + w; g < n_groups; g += gridDim.x * kWarpsPerCta) {
        const T* Sg = S + (size_t)g * RR;
        if (act)
            for (int i = 0; i < R; ++i) G[i * LD + lane] = (double)Sg[i * R + lane];
        __syncwarp();
        // Tm = Delta S ; G = Tm Delta^T
        warp_mm(R, lane, [&](int i, int k) { return Dm[i * LD + k]; }, [&](int k, int j) { return G[k * LD + j]; },
                [&](int i, int j, double v) { Tm[i * LD + j] = v; });
        warp_mm(R, lane, [&](int i, int k) { return Tm[i * LD + k]; }, [&](int k, int j) { return Dm[j * LD + k]; },
                [&](int i, int j, double v) { G[i * LD + j] = v; });
        if (warm) {
            // rotate into the eigenbasis of the previous inner iteration: Q0^T G Q0 is nearly diagonal, so 1-3 sweeps
            // suffice; the rotations are accumulated onto Q0, so Q stays the eigenvector matrix of G itself
            const double* Q0 = Qstore + (size_t)g * RR;
            if (act)
                for (int i = 0; i < R; ++i) Q[i * LD + lane] = Q0[i * R + lane];
            __syncwarp();
            warp_mm(R, lane, [&](int i, int k) { return G[i * LD + k]; }, [&](int k, int j) { return Q[k * LD + j]; },
                    [&](int i, int j, double v) { Tm[i * LD + j] = v; });
            warp_mm(R, lane, [&](int i, int k) { return Q[k * LD + i]; }, [&](int k, int j) { return Tm[k * LD + j]; },
                    [&](int i, int j, double v) { G[i * LD + j] = v; });
        } else {
            if (act)
                for (int i = 0; i < R; ++i) Q[i * LD + lane] = (i == lane) ? 1.0 : 0.0;
            __syncwarp();
        }
        if (act)  // symmetrise round-off: lane j owns the pairs (i, j), i < j
            for (int i = 0; i < lane; ++i) {
                const double v = 0.5 * (G[i * LD + lane] + G[lane * LD + i]);
                G[i * LD + lane] = v;
                G[lane * LD + i] = v;
            }
        __syncwarp();

        for (int sweep = 0; sweep < 40 && R > 1; ++sweep) {
            double off = 0.0, dg = 0.0;
            if (act)
                for (int i = 0; i < R; ++i) {
                    const double v = G[i * LD + lane];
                    if (i == lane) dg += v * v; else off += v * v;
                }
            off = warp_sum_d(off);
            dg = warp_sum_d(dg);
            if (!(off > 1e-26 * dg && off > 0.0)) break;  // off/diag <= 1e-13: eigenvectors at round-off
            for (int t = 0; t < m; ++t) {
                int p = 0, q = 0;
                bool real_pair = false;
                if (lane < half) {
                    p = t + lane;
                    if (p >= m) p -= m;
                    q = t - lane + m;
                    if (q >= m) q -= m;
                    if (lane == 0) {
                        p = t;
                        q = m;
                    }
                    if (p > q) {
                        const int tmp = p;
                        p = q;
                        q = tmp;
                    }
                    double c = 1.0, s = 0.0;
                    if (q < R) {
                        real_pair = true;
                        const double apq = G[p * LD + q], app = G[p * LD + p], aqq = G[q * LD + q];
                        if (fabs(apq) > 1e-300 && fabs(apq) > 1e-20 * sqrt(fabs(app * aqq))) {
                            const double tau = (aqq - app) / (2.0 * apq);
                            const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = 1.0 / sqrt(1.0 + tt * tt);
                            s = tt * c;
                        }
                    } else {
                        q = p;  // dummy partner of an odd rank: identity rotation of column p with itself
                    }
                    pq[lane] = p | (q << 16);
                    cs2[lane] = make_double2(c, s);
                }
                __syncwarp();
                if (act) {
                    // column pass (lane = row i): G <- G J, Q <- Q J.  The pairs of a round are disjoint, so two of
                    // them are loaded before either is stored (the compiler cannot prove that on its own).
                    double* Gi = G + lane * LD;
                    double* Qi = Q + lane * LD;
                    int k = 0;
                    for (; k + 1 < half; k += 2) {
                        const int pqa = pq[k], pqb = pq[k + 1];
                        const double2 ra = cs2[k], rb = cs2[k + 1];
                        const int pa = pqa & 0xffff, qa = pqa >> 16, pb = pqb & 0xffff, qb = pqb >> 16;
                        const double ca = ra.x, sa = ra.y, cb = rb.x, sb = rb.y;
                        const double gpa = Gi[pa], gqa = Gi[qa], gpb = Gi[pb], gqb = Gi[qb];
                        const double qpa = Qi[pa], qqa = Qi[qa], qpb = Qi[pb], qqb = Qi[qb];
                        Gi[pa] = ca * gpa - sa * gqa;
                        Gi[qa] = sa * gpa + ca * gqa;
                        Gi[pb] = cb * gpb - sb * gqb;
                        Gi[qb] = sb * gpb + cb * gqb;
                        Qi[pa] = ca * qpa - sa * qqa;
                        Qi[qa] = sa * qpa + ca * qqa;
                        Qi[pb] = cb * qpb - sb * qqb;
                        Qi[qb] = sb * qpb + cb * qqb;
                    }
                    if (k < half) {
                        const int pqa = pq[k];
                        const double2 ra = cs2[k];
                        const int pa = pqa & 0xffff, qa = pqa >> 16;
                        const double ca = ra.x, sa = ra.y;
                        const double gpa = Gi[pa], gqa = Gi[qa], qpa = Qi[pa], qqa = Qi[qa];
                        Gi[pa] = ca * gpa - sa * gqa;
                        Gi[qa] = sa * gpa + ca * gqa;
                        Qi[pa] = ca * qpa - sa * qqa;
                        Qi[qa] = sa * qpa + ca * qqa;
                    }
                }
                __syncwarp();
                if (act) {
                    // row pass (lane = column j): G <- J^T G
                    double* Gj = G + lane;
                    int k = 0;
                    for (; k + 1 < half; k += 2) {
                        const int pqa = pq[k], pqb = pq[k + 1];
                        const double2 ra = cs2[k], rb = cs2[k + 1];
                        const int pa = (pqa & 0xffff) * LD, qa = (pqa >> 16) * LD;
                        const int pb = (pqb & 0xffff) * LD, qb = (pqb >> 16) * LD;
                        const double ca = ra.x, sa = ra.y, cb = rb.x, sb = rb.y;
                        const double gpa = Gj[pa], gqa = Gj[qa], gpb = Gj[pb], gqb = Gj[qb];
                        Gj[pa] = ca * gpa - sa * gqa;
                        Gj[qa] = sa * gpa + ca * gqa;
                        Gj[pb] = cb * gpb - sb * gqb;
                        Gj[qb] = sb * gpb + cb * gqb;
                    }
                    if (k < half) {
                        const int pqa = pq[k];
                        const double2 ra = cs2[k];
                        const int pa = (pqa & 0xffff) * LD, qa = (pqa >> 16) * LD;
                        const double ca = ra.x, sa = ra.y;
                        const double gpa = Gj[pa], gqa = Gj[qa];
                        Gj[pa] = ca * gpa - sa * gqa;
                        Gj[qa] = sa * gpa + ca * gqa;
                    }
                }
                __syncwarp();
                if (real_pair) {  // the annihilated pair: exact zeros
                    G[p * LD + q] = 0.0;
                    G[q * LD + p] = 0.0;
                }
                __syncwarp();
            }
        }

        // lam^-1/2 (directions with lam <= eps * lam_max dropped)
        const double lam = act ? G[lane * LD + lane] : 0.0;
        const double lmax = warp_max_d(lam);
        cs[lane] = (act && lam > 1e-28 * lmax && lam > 0.0) ? 1.0 / sqrt(lam) : 0.0;
        if (Qstore && act) {
            double* Qo = Qstore + (size_t)g * RR;
            for (int i = 0; i < R; ++i) Qo[i * R + lane] = Q[i * LD + lane];
        }
        __syncwarp();
        // H = Q lam^-1/2 Q^T -> G
        warp_mm(R, lane, [&](int i, int k) { return Q[i * LD + k]; },
                [&](int k, int j) { return Q[j * LD + k] * cs[k]; },
                [&](int i, int j, double v) { G[i * LD + j] = v; });
        // W = Delta^T H
        T* Wg = Wmat + (size_t)g * RR;
        warp_mm(R, lane, [&](int i, int k) { return Dm[k * LD + i]; }, [&](int k, int j) { return G[k * LD + j]; },
                [&](int i, int j, double v) { Wg[i * R + j] = (T)v; });
        // num = rho W^T S = rho H (Delta S)
        if (act)
            for (int i = 0; i < R; ++i) Q[i * LD + lane] = (double)Sg[i * R + lane];
        __syncwarp();
        warp_mm(R, lane, [&](int i, int k) { return Dm[i * LD + k]; }, [&](int k, int j) { return Q[k * LD + j]; },
                [&](int i, int j, double v) { Tm[i * LD + j] = v; });
        const double rg = (double)rho[g];
        double* Ng = num_part + (size_t)g * RR;
        warp_mm(R, lane, [&](int i, int k) { return G[i * LD + k]; }, [&](int k, int j) { return Tm[k * LD + j]; },
                [&](int i, int j, double v) { Ng[i * R + j] = rg * v; });
    }
}

}  // namespace

template <typename T>
static int launch_polar_warp(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                             void* num_part, void* Qstore, int warm, cudaStream_t st) {
    const int LD = R | 1, MS = (R * LD + 1) & ~1;
    const size_t smem = (size_t)(MS + kWarpsPerCta * (3 * MS + 48)) * sizeof(double);
    auto kern = pf2_polar_warp_kernel<T>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    const int ctas_needed = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm);
    const int grid = ctas_needed < b2_num_sms() * per_sm ? ctas_needed : b2_num_sms() * per_sm;
    kern<<<grid, 32 * kWarpsPerCta, smem, st>>>((const T*)S, (const T*)Delta, (const T*)rho, n_groups, R, (T*)Wmat,
                                                (double*)num_part, (double*)Qstore, warm);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_pf2_polar_warp(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                      void* num_part, void* Qstore, int warm, int dtype, cudaStream_t st) {
    B2_DISPATCH_DTYPE(dtype, return launch_polar_warp<T>(S, Delta, rho, n_groups, R, Wmat, num_part, Qstore, warm, st));
    return B2_OK;
}
