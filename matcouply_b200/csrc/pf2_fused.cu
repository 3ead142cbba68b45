// Fused PARAFAC2 B-mode ADMM path (reference decomposition.py:259-289 with penalties.py:1224-1281).
//
// One inner ADMM iteration = three launches:
//   b2_pf2_rowpass : CTA per slice, one pass over the slice's rows. Per row (4 lanes per row, warp shuffles for the
//                    R x R products):  [deferred prox of the previous iteration:  pd = V T_g, dual = V - pd]
//                    -> x = (rho*sum(aux-dual) + Y o a) Minv_g -> V' = x + dual_pf2 (stored in the dual slot)
//                    -> prox + dual update of the other row-local penalties -> S_g += V'^T V' with DMMA.8x8x4.
//   b2_pf2_polar   : CTA per slice, parallel-ordering Jacobi eigen-decomposition of Delta S_g Delta^T (R/2 rotations
//                    at once) -> W_g with P_i = V'_i W_g = polar(V'_i Delta^T), and rho_g W_g^T S_g.
//   b2_pf2_delta   : two-level fixed-order sum over slices -> Delta_new (sums exposed for the cross-rank all-reduce).
// The basis matrices P_i and the products P_i Delta are never materialised inside the inner loop: the next row pass
// applies T_g = W_g Delta_new on the fly ("deferred"), and b2_pf2_apply materialises them once per outer iteration.
#include "admm_common.cuh"

namespace {

template <typename T, int CPL>
__global__ void __launch_bounds__(256, 2)
pf2_rowpass_kernel(const int64_t* __restrict__ row_off, int R, const T* __restrict__ Y, const T* __restrict__ A,
                   const T* __restrict__ rho, const T* __restrict__ Minv, PenArgs pa, int deferred_flags,
                   const T* __restrict__ Wmat, const T* __restrict__ Delta, T* __restrict__ x, T* __restrict__ w_out,
                   int ldw, T* __restrict__ S_out, T* __restrict__ BtB_out) {
    using L = RowLayout<T, CPL>;
    // bit 0: PARAFAC2 prox deferred; bit 1 / 2: elementwise extras arrive / leave as T = x + dual in their dual slot
    const bool deferred = (deferred_flags & 1) != 0, tin = (deferred_flags & 2) != 0, tout = (deferred_flags & 4) != 0;
    extern __shared__ double rp_smem[];
    const int RR = R * R;
    constexpr int NB = (CPL + 1) / 2;  // == ceil(R / 8) for every R with ceil(R / 4) == CPL
    constexpr int LDT = 8 * NB + 2;  // 2*LDT == 4 (mod 16): conflict-free fragment reads
    T* Ms = (T*)rp_smem;                                              // Minv_g, padded layout
    T* Ts = Ms + L::ELEMS;                                            // T_g = W_g Delta (deferred mode), padded layout
    double* tile = rp_smem + ((2 * L::ELEMS * sizeof(T) + 7) / 8);    // [64 x LDT] staged V' for the Gram MMA
    double* tile2 = tile + kRowsPerPass * LDT;                        // [64 x LDT] staged x (only with BtB_out)
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l4 = lane & 3;
    const long long r_begin = row_off[g], r_end = row_off[g + 1];
    if (r_begin >= r_end) {
        for (int e = tid; e < RR; e += blockDim.x) {
            S_out[(size_t)g * RR + e] = T(0);
            if (BtB_out) BtB_out[(size_t)g * RR + e] = T(0);
        }
        return;
    }
    for (int e = tid; e < 2 * L::ELEMS; e += blockDim.x) Ms[e] = T(0);
    if (deferred) {  // stage W_g and Delta (tile memory doubles as scratch before the row loop)
        T* tw = (T*)tile;
        T* td = tw + RR;
        for (int e = tid; e < RR; e += blockDim.x) {
            tw[e] = Wmat[(size_t)g * RR + e];
            td[e] = Delta[e];
        }
    }
    __syncthreads();
    for (int e = tid; e < RR; e += blockDim.x) {
        const int i = e / R, c = e - i * R;
        const int dst = i * L::LDM + (c / CPL) * L::CPLP + (c % CPL);
        Ms[dst] = Minv[(size_t)g * RR + e];
        if (deferred) {
            const T* tw = (const T*)tile;
            const T* td = tw + RR;
            T s = T(0);
            for (int k = 0; k < R; ++k) s = fma(tw[i * R + k], td[k * R + c], s);
            Ts[dst] = s;
        }
    }
    __syncthreads();
    for (int e = tid; e < 2 * kRowsPerPass * LDT; e += blockDim.x) tile[e] = 0.0;  // pad columns stay zero
    __syncthreads();

    const T rg = rho[g];
    const int c0 = l4 * CPL;
    const T* mseg = Ms + l4 * L::CPLP;
    const T* tseg = Ts + l4 * L::CPLP;
    T sc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) sc[j] = (c0 + j < R) ? A[(size_t)g * R + c0 + j] : T(0);
    double accS[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, accB[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int gq = lane >> 2, tq = lane & 3;
    const int n_pen = pa.n_pen;
    const T* pf_aux = (const T*)pa.aux[0];
    T* pf_dual = (T*)pa.dual[0];

    for (long long row0 = r_begin; row0 < r_end; row0 += kRowsPerPass) {
        const long long rowid = row0 + (tid >> 2);
        const bool valid = rowid < r_end;
        const long long row = valid ? rowid : r_end - 1;
        const size_t base = (size_t)row * R + c0;
        T r_[CPL], sh[CPL], dpf[CPL], xv[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) r_[j] = (c0 + j < R) ? Y[base + j] * sc[j] : T(0);
        if (deferred) {
            T v_[CPL], pdv[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                v_[j] = (c0 + j < R) ? pf_dual[base + j] : T(0);
                pdv[j] = T(0);
            }
            lane_matvec<T, CPL>(v_, tseg, lane, pdv);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                dpf[j] = v_[j] - pdv[j];   // dual = V - P Delta          (decomposition.py:282-285)
                sh[j] = pdv[j] - dpf[j];   // aux - dual = P Delta - dual (penalties.py:1280-1281)
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const bool in = c0 + j < R;
                const T pdv = in ? pf_aux[base + j] : T(0);
                dpf[j] = in ? pf_dual[base + j] : T(0);
                sh[j] = pdv - dpf[j];
            }
        }
        for (int p = 1; p < n_pen; ++p) {
            const T* ax = (const T*)pa.aux[p];
            const T* du = (const T*)pa.dual[p];
            const int kind = pa.kind[p];
            const bool elementwise = kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1;
            if (tin && elementwise) {  // T-only state: aux = prox(T), dual = T - aux
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    if (c0 + j < R) {
                        const T tv = du[base + j];
                        const T z = prox_elem<T>(tv, kind, pa.nn[p], (T)pa.p0[p], (T)pa.p1[p], rg);
                        sh[j] += z - (tv - z);
                    }
            } else {
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    if (c0 + j < R) sh[j] += ax[base + j] - du[base + j];
            }
        }
        T s_[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            s_[j] = rg * sh[j] + r_[j];
            xv[j] = T(0);
        }
        lane_matvec<T, CPL>(s_, mseg, lane, xv);
        // PARAFAC2: V' = x + dual ; stage for the Gram(s)
        double* trow = tile + (tid >> 2) * LDT + c0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            if (c0 + j < R) {
                const T vnew = xv[j] + dpf[j];
                if (valid) pf_dual[base + j] = vnew;
                trow[j] = valid ? (double)vnew : 0.0;
                if (BtB_out) trow[kRowsPerPass * LDT + j] = valid ? (double)xv[j] : 0.0;
                if (valid && x) x[base + j] = xv[j];
                if (valid && w_out) w_out[(size_t)row * ldw + c0 + j] = xv[j] * sc[j];
            }
        }
        for (int p = 1; p < n_pen; ++p) {
            T* ax = (T*)pa.aux[p];
            T* du = (T*)pa.dual[p];
            const int kind = pa.kind[p], nn = pa.nn[p];
            const bool elementwise = kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1;
            const T p0 = (T)pa.p0[p], p1 = (T)pa.p1[p];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                if (valid && c0 + j < R) {
                    T dold = du[base + j];
                    if (tin && elementwise) dold = dold - prox_elem<T>(dold, kind, nn, p0, p1, rg);
                    const T v = xv[j] + dold;
                    if (elementwise && tout) {
                        du[base + j] = v;  // T-only state: the next pass recomputes aux and dual
                    } else if (elementwise) {
                        const T z = prox_elem<T>(v, kind, nn, p0, p1, rg);
                        ax[base + j] = z;
                        du[base + j] = v - z;
                    } else {
                        du[base + j] = v;  // finished by b2_prox_l2ball / b2_prox_unimodal
                    }
                }
            }
        }
        __syncthreads();
        // S += V'^T V' (and B^T B += x^T x).  MMA: M = column i (8), N = column j (8), K = rows (r0 + h + 2t)
        tile_gram<NB>(tile, warp, gq, tq, accS);
        if (BtB_out) tile_gram<NB>(tile2, warp, gq, tq, accB);
        __syncthreads();
    }
    store_gram<T, NB>(S_out + (size_t)g * RR, R, warp, gq, tq, accS);
    if (BtB_out) store_gram<T, NB>(BtB_out + (size_t)g * RR, R, warp, gq, tq, accB);
}

// Per-slice Gram B_g^T B_g with the same staged-tile DMMA scheme (used after the non-PARAFAC2 B-updates and at start-up)
template <typename T, int CPL>
__global__ void __launch_bounds__(256)
slice_gram_kernel(const T* __restrict__ B, const int64_t* __restrict__ row_off, int R, T* __restrict__ BtB) {
    extern __shared__ double rp_smem[];
    const int RR = R * R;
    constexpr int NB = (CPL + 1) / 2;
    constexpr int LDT = 8 * NB + 2;  // 2*LDT == 4 (mod 16): conflict-free fragment reads
    double* tile = rp_smem;
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l4 = lane & 3;
    const long long r_begin = row_off[g], r_end = row_off[g + 1];
    for (int e = tid; e < kRowsPerPass * LDT; e += blockDim.x) tile[e] = 0.0;
    __syncthreads();
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int gq = lane >> 2, tq = lane & 3, c0 = l4 * CPL;
    for (long long row0 = r_begin; row0 < r_end; row0 += kRowsPerPass) {
        const long long row = row0 + (tid >> 2);
        const bool valid = row < r_end;
        double* trow = tile + (tid >> 2) * LDT + c0;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (c0 + j < R) trow[j] = valid ? (double)B[(size_t)row * R + c0 + j] : 0.0;
        __syncthreads();
        tile_gram<NB>(tile, warp, gq, tq, acc);
        __syncthreads();
    }
    store_gram<T, NB>(BtB + (size_t)g * RR, R, warp, gq, tq, acc);
}

// rhs[g][c] = sum_j B[j][c] * Y[j][c] over the rows of slice g (= diag(B_g^T X_g C), decomposition.py:158); fixed order
template <typename T>
__global__ void __launch_bounds__(256)
slice_coldot_kernel(const T* __restrict__ B, const T* __restrict__ Y, const int64_t* __restrict__ row_off, int R,
                    T* __restrict__ rhs) {
    __shared__ double part[256];
    const int g = blockIdx.x, tid = threadIdx.x;
    const long long r_begin = row_off[g], r_end = row_off[g + 1];
    const long long cnt = (r_end - r_begin) * R;
    const int tpr = (int)blockDim.x / R * R;  // active threads: multiple of R so a thread always sees one column
    double acc = 0.0;
    if (tid < tpr) {
        const T* Bg = B + r_begin * R;
        const T* Yg = Y + r_begin * R;
        for (long long e = tid; e < cnt; e += tpr) acc += (double)Bg[e] * (double)Yg[e];
    }
    part[tid] = tid < tpr ? acc : 0.0;
    __syncthreads();
    if (tid < R) {
        double s = 0.0;
        for (int t = tid; t < tpr; t += R) s += part[t];
        rhs[(size_t)g * R + tid] = (T)s;
    }
}

// lhs (R x R) = sum_g (a_g a_g^T) o BtB_g  (= sum_i (B_i a_i)^T (B_i a_i), decomposition.py:312-314); one block per
// output element, fixed summation order over slices.
template <typename T>
__global__ void weighted_gram_sum_kernel(const T* __restrict__ BtB, const T* __restrict__ A, int n_groups, int R,
                                         T* __restrict__ out) {
    __shared__ double scratch[32];
    const int e = blockIdx.x, i = e / R, j = e - i * R;
    double acc = 0.0;
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x)
        acc += (double)BtB[(size_t)g * R * R + e] * (double)A[(size_t)g * R + i] * (double)A[(size_t)g * R + j];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[e] = (T)acc;
}

// cross[g] = BtB[g] o CtC  (decomposition.py:155)
template <typename T>
__global__ void hadamard_bcast_kernel(const T* __restrict__ BtB, const T* __restrict__ CtC, long long total, int RR,
                                      T* __restrict__ cross) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        cross[i] = BtB[i] * CtC[i % RR];
}

// ---------------------------------------------------------------------------------------------------------
// CTA-per-slice polar step: G = Delta S Delta^T = Q Lam Q^T by parallel-ordering (round-robin) Jacobi.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPolarThreads = 128;

__device__ __forceinline__ void cta_matmul(const double* A, const double* B, double* C, int R, bool transA, bool transB) {
    for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
        const int i = e / R, j = e - i * R;
        double s = 0.0;
        for (int k = 0; k < R; ++k) s += (transA ? A[k * R + i] : A[i * R + k]) * (transB ? B[j * R + k] : B[k * R + j]);
        C[e] = s;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kPolarThreads)
pf2_polar_cta_kernel(const T* __restrict__ S, const T* __restrict__ Delta, const T* __restrict__ rho, int R,
                     T* __restrict__ Wmat, double* __restrict__ num_part, double* __restrict__ Qstore, int warm) {
    extern __shared__ double pj_smem[];
    const int RR = R * R;
    double* D = pj_smem;
    double* Sg = D + RR;
    double* G = Sg + RR;
    double* Q = G + RR;
    double* Tm = Q + RR;
    double* cs = Tm + RR;            // c[16], s[16]
    int* pq = (int*)(cs + 32);       // p[16], q[16]
    __shared__ double red[2 * (kPolarThreads / 32)];
    __shared__ int s_continue;
    const int g = blockIdx.x, tid = threadIdx.x;
    for (int e = tid; e < RR; e += blockDim.x) {
        D[e] = (double)Delta[e];
        Sg[e] = (double)S[(size_t)g * RR + e];
        Q[e] = (e / R == e % R) ? 1.0 : 0.0;
    }
    __syncthreads();
    cta_matmul(D, Sg, Tm, R, false, false);
    cta_matmul(Tm, D, G, R, false, true);
    if (warm) {
        // Warm start: rotate into the eigenbasis of the previous inner iteration (G changes little between inner
        // iterations, so Q0^T G Q0 is nearly diagonal and Jacobi needs 1-3 sweeps instead of 6-8).  The accumulated
        // rotations start from Q0, so Q stays the eigenvector matrix of G itself.
        for (int e = tid; e < RR; e += blockDim.x) Q[e] = Qstore[(size_t)g * RR + e];
        __syncthreads();
        cta_matmul(G, Q, Tm, R, false, false);
        cta_matmul(Q, Tm, G, R, true, false);
    }
    for (int e = tid; e < RR; e += blockDim.x) {  // symmetrise round-off
        const int i = e / R, j = e - i * R;
        if (i < j) {
            const double v = 0.5 * (G[i * R + j] + G[j * R + i]);
            G[i * R + j] = v;
            G[j * R + i] = v;
        }
    }
    __syncthreads();
    const int Re = R + (R & 1), m = Re - 1, half = Re / 2;
    for (int sweep = 0; sweep < 40 && R > 1; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < RR; e += blockDim.x) {
            const double v = G[e];
            if (e / R == e % R) dg += v * v; else off += v * v;
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if ((tid & 31) == 0) {
            red[2 * (tid >> 5)] = off;
            red[2 * (tid >> 5) + 1] = dg;
        }
        __syncthreads();
        if (tid == 0) {
            double o = 0.0, d = 0.0;
            for (int w = 0; w < kPolarThreads / 32; ++w) {
                o += red[2 * w];
                d += red[2 * w + 1];
            }
            s_continue = (o > 1e-26 * d && o > 0.0) ? 1 : 0;  // off/diag <= 1e-13: eigenvectors at round-off
        }
        __syncthreads();
        if (!s_continue) break;
        for (int t = 0; t < m; ++t) {
            if (tid < half) {
                int p = (tid == 0) ? t : (t + tid) % m;
                int q = (tid == 0) ? m : (t - tid + m) % m;
                if (p > q) {
                    const int tmp = p;
                    p = q;
                    q = tmp;
                }
                double c = 1.0, s = 0.0;
                if (q < R) {
                    const double apq = G[p * R + q], app = G[p * R + p], aqq = G[q * R + q];
                    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-20 * sqrt(fabs(app * aqq))) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        s = tt * c;
                    }
                } else {
                    q = -1;  // dummy partner of an odd rank
                }
                pq[tid] = p;
                pq[16 + tid] = q;
                cs[tid] = c;
                cs[16 + tid] = s;
            }
            __syncthreads();
            // Two-sided update G <- J^T G J as independent 2x2 blocks: thread (k, k') owns rows {p,q} of pair k and
            // columns {p',q'} of pair k' (column rotation first, then row rotation: the same operation order as a
            // column pass followed by a row pass, without the barrier in between).
            for (int id = tid; id < half * half; id += blockDim.x) {
                const int k = id / half, k2 = id - k * half;
                const int p = pq[k], q = pq[16 + k], p2 = pq[k2], q2 = pq[16 + k2];
                const double c = cs[k], sn = cs[16 + k], c2 = cs[k2], sn2 = cs[16 + k2];
                const bool hq = q >= 0, hq2 = q2 >= 0;
                const double b00 = G[p * R + p2], b01 = hq2 ? G[p * R + q2] : 0.0;
                const double b10 = hq ? G[q * R + p2] : 0.0, b11 = (hq && hq2) ? G[q * R + q2] : 0.0;
                // columns (dummy partner: c2 = 1, sn2 = 0)
                const double d00 = c2 * b00 - sn2 * b01, d01 = sn2 * b00 + c2 * b01;
                const double d10 = c2 * b10 - sn2 * b11, d11 = sn2 * b10 + c2 * b11;
                // rows
                double e00 = c * d00 - sn * d10, e01 = c * d01 - sn * d11;
                double e10 = sn * d00 + c * d10, e11 = sn * d01 + c * d11;
                if (k == k2) e01 = e10 = 0.0;  // the annihilated pair
                G[p * R + p2] = e00;
                if (hq2) G[p * R + q2] = e01;
                if (hq) G[q * R + p2] = e10;
                if (hq && hq2) G[q * R + q2] = e11;
            }
            // eigenvectors: Q <- Q J (lane = row i, warps stride over the pairs)
            for (int k = tid >> 5; k < half; k += kPolarThreads / 32) {
                const int i = tid & 31;
                const int p = pq[k], q = pq[16 + k];
                if (i < R && q >= 0) {
                    const double c = cs[k], sn = cs[16 + k];
                    const double qp = Q[i * R + p], qq = Q[i * R + q];
                    Q[i * R + p] = c * qp - sn * qq;
                    Q[i * R + q] = sn * qp + c * qq;
                }
            }
            __syncthreads();
        }
    }
    // lam^-1/2 (directions with lam <= eps * lam_max dropped)
    if (tid == 0) {
        double lmax = 0.0;
        for (int k = 0; k < R; ++k) lmax = fmax(lmax, G[k * R + k]);
        cs[0] = lmax;
    }
    __syncthreads();
    const double lmax = cs[0];
    __syncthreads();
    if (Qstore)
        for (int e = tid; e < RR; e += blockDim.x) Qstore[(size_t)g * RR + e] = Q[e];
    for (int e = tid; e < RR; e += blockDim.x) {
        const int j = e % R;
        const double lam = G[j * R + j];
        const double isq = (lam > 1e-28 * lmax && lam > 0.0) ? 1.0 / sqrt(lam) : 0.0;
        Tm[e] = Q[e] * isq;
    }
    __syncthreads();
    cta_matmul(Tm, Q, G, R, false, true);   // G = Q lam^-1/2 Q^T
    cta_matmul(D, G, Tm, R, true, false);   // Tm = Delta^T G = W
    for (int e = tid; e < RR; e += blockDim.x) Wmat[(size_t)g * RR + e] = (T)Tm[e];
    cta_matmul(Tm, Sg, G, R, true, false);  // G = W^T S
    const double rg = (double)rho[g];
    for (int e = tid; e < RR; e += blockDim.x) num_part[(size_t)g * RR + e] = rg * G[e];
}

// sums[e] = sum_g num_part[g][e] (e < RR) ; sums[RR] = sum_g rho[g].  One block per element, fixed order.
template <typename T>
__global__ void pf2_sum_kernel(const double* __restrict__ num_part, const T* __restrict__ rho, int n_groups, int RR,
                               double* __restrict__ sums) {
    __shared__ double scratch[32];
    const int e = blockIdx.x;
    double acc = 0.0;
    if (e < RR)
        for (int g = threadIdx.x; g < n_groups; g += blockDim.x) acc += num_part[(size_t)g * RR + e];
    else
        for (int g = threadIdx.x; g < n_groups; g += blockDim.x) acc += (double)rho[g];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) sums[e] = acc;
}

template <typename T>
__global__ void pf2_normalise_kernel(const double* __restrict__ sums, int RR, T* __restrict__ Delta_new) {
    for (int e = threadIdx.x; e < RR; e += blockDim.x) Delta_new[e] = (T)(sums[e] / sums[RR]);
}

template <typename T, int CPL>
int launch_rowpass(const int64_t* row_off, int n_groups, int R, const void* Y, const void* A, const void* rho,
                   const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta, void* x,
                   void* w_out, int ldw, void* S_out, void* BtB_out, cudaStream_t st) {
    using L = RowLayout<T, CPL>;
    const int NB = (R + 7) / 8, LDT = 8 * NB + 2;
    const size_t smem = ((2 * (size_t)L::ELEMS * sizeof(T) + 7) / 8) * 8 + 2 * (size_t)kRowsPerPass * LDT * sizeof(double);
    auto kern = pf2_rowpass_kernel<T, CPL>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_groups, 256, smem, st>>>(row_off, R, (const T*)Y, (const T*)A, (const T*)rho, (const T*)Minv, pa, deferred,
                                      (const T*)Wmat, (const T*)Delta, (T*)x, (T*)w_out, ldw, (T*)S_out, (T*)BtB_out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}


// Feasibility-gap terms of the PARAFAC2 penalty straight from the deferred state (decomposition.py:406-415 with
// penalties.py:1287-1304): per slice  d2 = ||V_g T_g - x_g||^2 (T_g = W_g Delta, V_g T_g = P_g Delta),  x2 = ||x_g||^2,
// ab = sum |x_g|.  Nothing is materialised.  part[g*3 + {0,1,2}]; fixed summation order.
template <typename T, int CPL>
__global__ void __launch_bounds__(256)
pf2_gap_kernel(const int64_t* __restrict__ row_off, int R, const T* __restrict__ V, const T* __restrict__ x,
               const T* __restrict__ Wmat, const T* __restrict__ Delta, double* __restrict__ part) {
    using L = RowLayout<T, CPL>;
    extern __shared__ double rp_smem[];
    __shared__ double scratch[32];
    const int RR = R * R;
    T* Ts = (T*)rp_smem;
    T* tw = Ts + L::ELEMS;
    T* td = tw + RR;
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, l4 = lane & 3;
    const long long r_begin = row_off[g], r_end = row_off[g + 1];
    for (int e = tid; e < L::ELEMS; e += blockDim.x) Ts[e] = T(0);
    for (int e = tid; e < RR; e += blockDim.x) {
        tw[e] = Wmat[(size_t)g * RR + e];
        td[e] = Delta[e];
    }
    __syncthreads();
    for (int e = tid; e < RR; e += blockDim.x) {
        const int i = e / R, c = e - i * R;
        T s = T(0);
        for (int k = 0; k < R; ++k) s = fma(tw[i * R + k], td[k * R + c], s);
        Ts[i * L::LDM + (c / CPL) * L::CPLP + (c % CPL)] = s;
    }
    __syncthreads();
    const int c0 = l4 * CPL;
    const T* tseg = Ts + l4 * L::CPLP;
    double d2 = 0.0, x2 = 0.0, ab = 0.0;
    // software pipeline: the rows of pass p + 1 are in flight while pass p is contracted (the kernel is otherwise bound
    // by the global-load -> shuffle-matvec dependency with nothing to overlap it)
    T vn[CPL], xn[CPL];
    auto fetch = [&](long long row0) {
        const long long rowid = row0 + (tid >> 2);
        const bool ok = rowid < r_end;
        const size_t base = (size_t)(ok ? rowid : r_end - 1) * R + c0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const bool in = ok && c0 + j < R;
            vn[j] = in ? V[base + j] : T(0);
            xn[j] = in ? x[base + j] : T(0);
        }
    };
    fetch(r_begin);
    for (long long row0 = r_begin; row0 < r_end; row0 += kRowsPerPass) {
        T v_[CPL], xc[CPL], pdv[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            v_[j] = vn[j];
            xc[j] = xn[j];
            pdv[j] = T(0);
        }
        if (row0 + kRowsPerPass < r_end) fetch(row0 + kRowsPerPass);
        lane_matvec<T, CPL>(v_, tseg, lane, pdv);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            // rows past the end and pad columns were fetched as zeros and contribute nothing (pdv of a zero row is zero)
            const double xv = (double)xc[j];
            const double d = xv - (double)pdv[j];
            d2 += d * d;
            x2 += xv * xv;
            ab += fabs(xv);
        }
    }
    d2 = block_sum(d2, scratch);
    x2 = block_sum(x2, scratch);
    ab = block_sum(ab, scratch);
    if (tid == 0) {
        part[(size_t)g * 3 + 0] = d2;
        part[(size_t)g * 3 + 1] = x2;
        part[(size_t)g * 3 + 2] = ab;
    }
}

__global__ void pf2_gap_final_kernel(const double* __restrict__ part, int n_groups, double* __restrict__ out) {
    __shared__ double scratch[32];
    for (int k = 0; k < 3; ++k) {
        double acc = 0.0;
        for (int g = threadIdx.x; g < n_groups; g += blockDim.x) acc += part[(size_t)g * 3 + k];
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) out[k] = acc;
        __syncthreads();
    }
}

template <typename T, int CPL>
int launch_gap(const int64_t* row_off, int n_groups, int R, const void* V, const void* x, const void* Wmat,
               const void* Delta, double* part, cudaStream_t st) {
    using L = RowLayout<T, CPL>;
    const size_t smem = (((size_t)L::ELEMS + 2 * (size_t)R * R) * sizeof(T) + 7) / 8 * 8;
    auto kern = pf2_gap_kernel<T, CPL>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_groups, 256, smem, st>>>(row_off, R, (const T*)V, (const T*)x, (const T*)Wmat, (const T*)Delta, part);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T, int CPL>
int launch_slice_gram(const void* B, const int64_t* row_off, int n_groups, int R, void* BtB, cudaStream_t st) {
    const int NB = (R + 7) / 8, LDT = 8 * NB + 2;
    const size_t smem = (size_t)kRowsPerPass * LDT * sizeof(double);
    slice_gram_kernel<T, CPL><<<n_groups, 256, smem, st>>>((const T*)B, row_off, R, (T*)BtB);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_pf2_rowpass_fused_stats_supported(int R, int dtype, int n_pen, int companion_kind, int deferred) {
    return b2_pf2_rowpass_v2_applies(R, dtype, n_pen, companion_kind, deferred);
}

int b2_pf2_rowpass(const int64_t* row_off, int n_groups, int R, const void* Y, const void* A, const void* rho,
                   const void* Minv, const b2_penalty_desc* pens, int n_pen, int deferred, const void* Wmat,
                   const void* Delta, void* x, void* w_out, int ldw, void* S_out, void* BtB_out,
                   double* comp_stats_part, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(n_pen >= 1 && pens[0].kind == B2_PEN_PARAFAC2, "b2_pf2_rowpass: pens[0] must be the PARAFAC2 penalty");
    B2_REQUIRE((deferred & ~7) == 0, "b2_pf2_rowpass: unknown flag bits in `deferred`");
    if (n_groups == 0) return B2_OK;  // empty shard: zero-size buffers may arrive as NULL
    B2_REQUIRE(!(deferred & 1) || (Wmat && Delta), "deferred mode needs Wmat and Delta");
    PenArgs pa;
    {
        const int rc = b2_pack_penalties(pens, n_pen, &pa);
        if (rc != B2_OK) return rc;
    }
    if (comp_stats_part != nullptr) {
        B2_REQUIRE(x != nullptr && n_pen == 2, "b2_pf2_rowpass: comp_stats_part belongs to the last pass with one companion");
        B2_REQUIRE(b2_pf2_rowpass_v2_applies(R, dtype, n_pen, pa.kind[1], deferred),
                   "b2_pf2_rowpass: comp_stats_part needs the steady-state kernel "
                   "(check b2_pf2_rowpass_fused_stats_supported first)");
    }
    if (b2_option_value(B2_OPT_PF2_ROWPASS_MMA) >= 2) {  // steady-state specialisation (pf2_rowpass_v2.cu)
        const int rc = b2_pf2_rowpass_v2_try(row_off, n_groups, R, Y, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out,
                                             ldw, S_out, BtB_out, comp_stats_part, dtype, st);
        if (rc >= 0) return rc;
    }
    B2_REQUIRE(comp_stats_part == nullptr, "b2_pf2_rowpass: the steady-state kernel did not take this call "
                                           "(unaligned buffers?), comp_stats_part cannot be served");
    if (b2_option_value(B2_OPT_PF2_ROWPASS_MMA)) {  // tensor-core formulation (pf2_mma.cu) when it applies
        const int rc = b2_pf2_rowpass_mma_try(row_off, n_groups, R, Y, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out,
                                              ldw, S_out, BtB_out, dtype, st);
        if (rc >= 0) return rc;
    }
    const int CPL = (R + 3) / 4;
#define B2_CASE_CPL(C)                                                                                             \
    case C:                                                                                                        \
        B2_DISPATCH_DTYPE(dtype, return launch_rowpass<T, C>(row_off, n_groups, R, Y, A, rho, Minv, pa, deferred,  \
                                                             Wmat, Delta, x, w_out, ldw, S_out, BtB_out, st));     \
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

int b2_pf2_polar(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat, void* num_part,
                 void* Qstore, int warm, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;  // empty shard: zero-size buffers may arrive as NULL
    B2_REQUIRE(!warm || Qstore, "b2_pf2_polar: a warm start needs the eigenvector store");
    if (b2_option_value(B2_OPT_POLAR_WARP) >= 2) {
        const int rc = b2_pf2_polar_reg(S, Delta, rho, n_groups, R, Wmat, num_part, Qstore, warm, dtype, st);
        if (rc >= 0) return rc;
    }
    if (b2_option_value(B2_OPT_POLAR_WARP))
        return b2_pf2_polar_warp(S, Delta, rho, n_groups, R, Wmat, num_part, Qstore, warm, dtype, st);
    const size_t smem = (size_t)(5 * R * R + 32) * sizeof(double) + 32 * sizeof(int);
    B2_DISPATCH_DTYPE(dtype, {
        auto kern = pf2_polar_cta_kernel<T>;
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_groups, kPolarThreads, smem, st>>>((const T*)S, (const T*)Delta, (const T*)rho, R, (T*)Wmat,
                                                    (double*)num_part, (double*)Qstore, warm);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_pf2_delta(const void* num_part, const void* rho, int n_groups, int R, void* Delta_new, void* sums,
                 const void* sums_in, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int RR = R * R;
    B2_REQUIRE(sums_in != nullptr || sums != nullptr, "b2_pf2_delta needs the `sums` scratch (R*R+1 doubles)");
    B2_DISPATCH_DTYPE(dtype, {
        if (!sums_in) {
            pf2_sum_kernel<T><<<RR + 1, 256, 0, st>>>((const double*)num_part, (const T*)rho, n_groups, RR,
                                                      (double*)sums);
            B2_LAUNCH_CHECK();
        }
        pf2_normalise_kernel<T><<<1, 256, 0, st>>>((const double*)(sums_in ? sums_in : sums), RR, (T*)Delta_new);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_pf2_gap(const void* V, const void* x, const int64_t* row_off, int n_groups, int R, const void* Wmat,
               const void* Delta, double* out, int dtype, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(ws_bytes >= (size_t)3 * (n_groups > 0 ? n_groups : 1) * sizeof(double), "b2_pf2_gap workspace too small");
    if (n_groups == 0) {
        B2_CHECK_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(double), st));
        return B2_OK;
    }
    if (b2_option_value(B2_OPT_PF2_ROWPASS_MMA)) {  // tensor-core formulation (pf2_gap_mma.cu) when it applies
        const int rc_mma = b2_pf2_gap_mma_try(V, x, row_off, n_groups, R, Wmat, Delta, (double*)ws, dtype, st);
        if (rc_mma > 0) return rc_mma;
        if (rc_mma == B2_OK) {
            pf2_gap_final_kernel<<<1, 256, 0, st>>>((const double*)ws, n_groups, out);
            B2_LAUNCH_CHECK();
            return B2_OK;
        }
    }
    const int CPL = (R + 3) / 4;
    int rc = B2_OK;
#define B2_CASE_CPL(C)                                                                                           \
    case C:                                                                                                      \
        B2_DISPATCH_DTYPE(dtype, rc = launch_gap<T, C>(row_off, n_groups, R, V, x, Wmat, Delta, (double*)ws, st)); \
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
    if (rc != B2_OK) return rc;
    pf2_gap_final_kernel<<<1, 256, 0, st>>>((const double*)ws, n_groups, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_group_stats_sum(const double* part, int n_groups, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_groups == 0) {
        B2_CHECK_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(double), st));
        return B2_OK;
    }
    pf2_gap_final_kernel<<<1, 256, 0, st>>>(part, n_groups, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_slice_gram(const void* B, const int64_t* row_off, int n_groups, int R, void* BtB, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    const int CPL = (R + 3) / 4;
#define B2_CASE_CPL(C)                                                                                    \
    case C:                                                                                               \
        B2_DISPATCH_DTYPE(dtype, return launch_slice_gram<T, C>(B, row_off, n_groups, R, BtB, st));       \
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

int b2_slice_coldot(const void* B, const void* Y, const int64_t* row_off, int n_groups, int R, void* rhs, int dtype,
                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    B2_DISPATCH_DTYPE(dtype, {
        slice_coldot_kernel<T><<<n_groups, 256, 0, st>>>((const T*)B, (const T*)Y, row_off, R, (T*)rhs);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_weighted_gram_sum(const void* BtB, const void* A, int n_groups, int R, void* out, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_DISPATCH_DTYPE(dtype, {
        weighted_gram_sum_kernel<T><<<R * R, 256, 0, st>>>((const T*)BtB, (const T*)A, n_groups, R, (T*)out);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_hadamard_bcast(const void* BtB, const void* CtC, int n_groups, int R, void* cross, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_groups == 0) return B2_OK;
    const long long total = (long long)n_groups * R * R;
    long long blocks = (total + 255) / 256;
    if (blocks > b2_num_sms() * 8) blocks = b2_num_sms() * 8;
    B2_DISPATCH_DTYPE(dtype, {
        hadamard_bcast_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)BtB, (const T*)CtC, total, R * R, (T*)cross);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

}  // extern "C"
