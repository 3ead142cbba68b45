// Small dense and batched per-slice math of the AO-ADMM sub-solvers: Gram matrices, feasibility penalties rho,
// Cholesky-based R x R inverses (one warp per slice), per-slice cross products, and the PARAFAC2 Procrustes step
// materialisation pass (the Procrustes step itself lives in pf2_fused.cu).
#include <math.h>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// G = M^T M for n x R (n arbitrary): blocks own row ranges, partial Grams -> fixed-order final reduction
// ---------------------------------------------------------------------------------------------------------
constexpr int kGramRows = 32;  // rows staged per step

template <typename T>
__global__ void gram_partial_kernel(const T* __restrict__ M, long long n, int R, int ld, double* __restrict__ part) {
    __shared__ T tile[kGramRows * B2_MAX_RANK];
    const int RR = R * R;
    // each thread owns outputs e = tid, tid + blockDim, ... (at most 4 for R=32, 256 threads)
    double acc[4] = {0, 0, 0, 0};
    const long long rows_per_block = (n + gridDim.x - 1) / gridDim.x;
    const long long r_begin = (long long)blockIdx.x * rows_per_block;
    long long r_end = r_begin + rows_per_block;
    if (r_end > n) r_end = n;
    for (long long r0 = r_begin; r0 < r_end; r0 += kGramRows) {
        const int nr = (int)((r_end - r0) < kGramRows ? (r_end - r0) : kGramRows);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * R; i += blockDim.x) tile[i] = M[(r0 + i / R) * ld + i % R];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + u * blockDim.x;
            if (e < RR) {
                const int a = e / R, b = e - a * R;
                double s = 0.0;
                for (int j = 0; j < nr; ++j) s += (double)tile[j * R + a] * (double)tile[j * R + b];
                acc[u] += s;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + u * blockDim.x;
        if (e < RR) part[(size_t)blockIdx.x * RR + e] = acc[u];
    }
}

template <typename T>
__global__ void gram_final_kernel(const double* __restrict__ part, int blocks, int RR, T* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= RR) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += part[(size_t)b * RR + e];
    G[e] = (T)s;
}

// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void scale_gram_kernel(const T* __restrict__ G, const T* __restrict__ A, int n_groups, int R,
                                  T* __restrict__ lhs) {
    const int RR = R * R;
    const long long total = (long long)n_groups * RR;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / RR), e = (int)(i - (long long)g * RR);
        const int r = e / R, s = e - r * R;
        // (CtC[r][s] * a[s]) * a[r]   — same association as transpose(transpose(CtC * a) * a)
        lhs[i] = (G[e] * A[(size_t)g * R + s]) * A[(size_t)g * R + r];
    }
}

// rho[g] = 0.5*trace*scale ; rho_max via one block (n_groups up to ~1e5: grid-stride inside a single block pass 2)
template <typename T>
__global__ void rho_trace_kernel(const T* __restrict__ lhs, int n_groups, int R, double scale, T* __restrict__ rho) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const T* L = lhs + (size_t)g * R * R;
    T tr = T(0);
    for (int r = 0; r < R; ++r) tr += L[r * R + r];
    rho[g] = (T)(T(0.5) * tr * (T)scale);
}

template <typename T>
__global__ void max_kernel(const T* __restrict__ v, int n, T* __restrict__ out) {
    __shared__ T scratch[32];
    T m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = v[i] > m ? v[i] : m;
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        T u = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : (T)-INFINITY;
        u = warp_max(u);
        if (threadIdx.x == 0) out[0] = u;
    }
}

template <typename T>
__global__ void fill_kernel(T* __restrict__ out, int n, double value) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = (T)value;
}

// ---------------------------------------------------------------------------------------------------------
// Batched SPD inverse by Cholesky, one warp per matrix (lane = column). fp64 internally.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFactorWarps = 4;

// Fallback of factor_batch_kernel for a matrix that is not (numerically) positive definite: the reference applies
// lhs^-1 through an SVD, `x (U / s) Uh` (decomposition.py:172, 256, 321), which for a symmetric matrix equals
// Q diag(1 / lam) Q^T of its eigen-decomposition whatever the signs of lam — it does not need definiteness.  Cyclic
// Jacobi by one warp on the R x R matrix in shared memory (lane = row in the column pass, column in the row pass);
// only reached when a Cholesky pivot is <= 0 or NaN (no penalty and no ridge on a rank-deficient Gram matrix, or a
// negative l2_penalty), so it is written for clarity, not speed.  G: the matrix (destroyed), Q: R x R scratch.
template <typename T>
__device__ void eig_inverse_warp(double* __restrict__ G, double* __restrict__ Q, int R, int lane, T* __restrict__ dst) {
    for (int e = lane; e < R * R; e += 32) Q[e] = (e / R == e % R) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60 && R > 1; ++sweep) {
        double off = 0.0, dg = 0.0;
        if (lane < R)
            for (int i = 0; i < R; ++i) {
                const double v = G[i * R + lane];
                if (i == lane) dg += v * v; else off += v * v;
            }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (!(off > 1e-30 * dg && off > 0.0)) break;
        for (int p = 0; p < R - 1; ++p)
            for (int q = p + 1; q < R; ++q) {
                const double apq = G[p * R + q], app = G[p * R + p], aqq = G[q * R + q];
                __syncwarp();  // every lane has read the pivot entries before anyone rotates them
                if (!(fabs(apq) > 1e-300 && fabs(apq) > 1e-20 * sqrt(fabs(app * aqq)))) continue;  // warp-uniform
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
                if (lane < R) {  // G <- G J, Q <- Q J (lane = row)
                    const double gp = G[lane * R + p], gq = G[lane * R + q];
                    G[lane * R + p] = c * gp - sn * gq;
                    G[lane * R + q] = sn * gp + c * gq;
                    const double qp = Q[lane * R + p], qq = Q[lane * R + q];
                    Q[lane * R + p] = c * qp - sn * qq;
                    Q[lane * R + q] = sn * qp + c * qq;
                }
                __syncwarp();
                if (lane < R) {  // G <- J^T G (lane = column)
                    const double gp = G[p * R + lane], gq = G[q * R + lane];
                    G[p * R + lane] = c * gp - sn * gq;
                    G[q * R + lane] = sn * gp + c * gq;
                }
                __syncwarp();
            }
    }
    __syncwarp();
    if (lane < R)  // Minv = Q diag(1 / lam) Q^T ; lane = column s
        for (int r = 0; r < R; ++r) {
            double acc = 0.0;
            for (int k = 0; k < R; ++k) acc += Q[r * R + k] * Q[lane * R + k] / G[k * R + k];
            dst[r * R + lane] = (T)acc;
        }
}

// Shared-memory layout: row stride LD = R | 1 (odd), so that a warp access to one COLUMN (lane = row) and to one ROW
// (lane = column) are both bank-conflict free.  The right-looking trailing update runs with lane = column j:
// L[i][j] -= L[i][k] * L[j][k] for i = k+1..R-1 is one broadcast load + one FMA on a conflict-free row access per i
// (the first version walked lane = row with a lane-dependent inner loop and a 16-way conflicted stride R).  Every
// element sees the same operations in the same order as before (k ascending, one FMA each): bit-identical results.
template <typename T>
__global__ void factor_batch_kernel(const T* __restrict__ lhs, int n_groups, int R, T* __restrict__ rho,
                                    const T* __restrict__ rho_max, int n_reg, double l2, T* __restrict__ Minv) {
    extern __shared__ double fsm[];  // per warp: L (R*LD) + Z (R*LD)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kFactorWarps + warp;
    if (g >= n_groups) return;
    const int LD = R | 1;
    double* L = fsm + (size_t)warp * 2 * R * LD;
    double* Z = L + R * LD;
    double rho_g = rho_max ? (double)rho_max[0] : (double)rho[g];
    if (rho_max && lane == 0) rho[g] = (T)rho_g;
    const double shift = rho_g * n_reg + l2;
    const T* src = lhs + (size_t)g * R * R;
    for (int e = lane; e < R * R; e += 32) {
        const int r = e / R, c = e - r * R;
        L[r * LD + c] = (double)src[e] + (r == c ? shift : 0.0);
    }
    __syncwarp();
    // in-place Cholesky (lower), right-looking
    bool spd = true;
    for (int k = 0; k < R; ++k) {
        const double dd = L[k * LD + k];
        if (!(dd > 0.0)) {  // same value in every lane: warp-uniform exit
            spd = false;
            break;
        }
        const double d = sqrt(dd);
        __syncwarp();
        if (lane == k) L[k * LD + k] = d;
        if (lane > k && lane < R) L[lane * LD + k] /= d;  // lane = row: column k
        __syncwarp();
        const bool upd = lane > k && lane < R;
        const double ljk = upd ? L[lane * LD + k] : 0.0;  // lane = column j of the trailing block
        for (int i = k + 1; i < R; ++i) {
            const double lik = L[i * LD + k];  // broadcast
            if (upd && lane <= i) L[i * LD + lane] -= lik * ljk;
        }
        __syncwarp();
    }
    if (!spd) {  // not positive definite: inverse through the eigen-decomposition, like the reference's SVD route
        __syncwarp();
        for (int e = lane; e < R * R; e += 32) {  // eig_inverse_warp works on dense R x R scratch
            const int r = e / R, c = e - r * R;
            L[e] = 0.5 * ((double)src[e] + (double)src[c * R + r]) + (r == c ? shift : 0.0);
        }
        __syncwarp();
        eig_inverse_warp<T>(L, L + R * R, R, lane, Minv + (size_t)g * R * R);
        return;
    }
    // Z = L^-1 (lower): lane = column c, forward substitution in its right-looking form: once Z[i][c] is known, the
    // R - i - 1 updates Z[i'][c] -= L[i'][i] Z[i][c] (i' > i) are independent of each other (the left-looking form is
    // one dependent load-FMA chain per entry).  Every entry still receives its subtractions in ascending j, one FMA
    // each, so the values are those of the left-looking loop bit for bit.
    if (lane < R) {
        const int c = lane;
        for (int i = 0; i < R; ++i) Z[i * LD + c] = (i == c) ? 1.0 : 0.0;
        for (int i = c; i < R; ++i) {
            const double zi = Z[i * LD + c] / L[i * LD + i];
            Z[i * LD + c] = zi;
            for (int ip = i + 1; ip < R; ++ip) Z[ip * LD + c] -= L[ip * LD + i] * zi;
        }
    }
    __syncwarp();
    // Minv = Z^T Z ; lane = column s; four rows r at a time (independent accumulator chains).  The sums run over all
    // k: Z is lower triangular with exact zeros above the diagonal, and an FMA with a zero factor leaves the
    // accumulator unchanged, so this equals the sum from k = max(r, s).
    if (lane < R) {
        const int s = lane;
        T* dst = Minv + (size_t)g * R * R;
        for (int r0 = 0; r0 < R; r0 += 4) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = r0; k < R; ++k) {
                const double zks = Z[k * LD + s];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (r0 + u < R) acc[u] += Z[k * LD + r0 + u] * zks;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < R) dst[(r0 + u) * R + s] = (T)acc[u];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-slice cross products: CTA per slice.  cross[g] = (B_g^T B_g) [o CtC], rhs[g][r] = sum_j B[j][r] Y[j][r]
// ---------------------------------------------------------------------------------------------------------
constexpr int kCrossRows = 32;

template <typename T>
__global__ void slice_cross_kernel(const T* __restrict__ B, const T* __restrict__ Y, const int64_t* __restrict__ row_off,
                                   int R, const T* __restrict__ CtC, T* __restrict__ cross, T* __restrict__ rhs) {
    __shared__ T tb[kCrossRows * B2_MAX_RANK];
    __shared__ T ty[kCrossRows * B2_MAX_RANK];
    const int g = blockIdx.x;
    const long long r_begin = row_off[g], r_end = row_off[g + 1];
    const int RR = R * R;
    double acc[4] = {0, 0, 0, 0};
    double racc = 0.0;  // threads < R accumulate rhs column
    for (long long r0 = r_begin; r0 < r_end; r0 += kCrossRows) {
        const int nr = (int)((r_end - r0) < kCrossRows ? (r_end - r0) : kCrossRows);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * R; i += blockDim.x) {
            tb[i] = B[r0 * R + i];
            if (Y) ty[i] = Y[r0 * R + i];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + u * blockDim.x;
            if (e < RR) {
                const int a = e / R, b = e - a * R;
                double s = 0.0;
                for (int j = 0; j < nr; ++j) s += (double)tb[j * R + a] * (double)tb[j * R + b];
                acc[u] += s;
            }
        }
        if (Y && threadIdx.x < R) {
            double s = 0.0;
            for (int j = 0; j < nr; ++j) s += (double)tb[j * R + threadIdx.x] * (double)ty[j * R + threadIdx.x];
            racc += s;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + u * blockDim.x;
        if (e < RR) {
            double v = acc[u];
            if (CtC) v *= (double)CtC[e];
            cross[(size_t)g * RR + e] = (T)v;
        }
    }
    if (rhs && Y && threadIdx.x < R) rhs[(size_t)g * R + threadIdx.x] = (T)racc;
}

template <typename T>
__global__ void rowscale_kernel(const T* __restrict__ B, const T* __restrict__ A, const int32_t* __restrict__ gor,
                                long long n, int R, T* __restrict__ W, int ldw) {
    const long long total = n * R;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / R;
        const int c = (int)(i - row * R);
        W[row * ldw + c] = B[i] * A[(size_t)gor[row] * R + c];  // pad columns [R, ldw) stay zero
    }
}

// PARAFAC2: materialise P_i Delta / dual (and optionally P_i) from the pre-image V and W_g (see pf2_fused.cu).
template <typename T, int RM>
__global__ void pf2_apply_kernel(T* __restrict__ pd, T* __restrict__ dual, T* __restrict__ basis,
                                 const T* __restrict__ Wmat, const T* __restrict__ Delta_new,
                                 const int32_t* __restrict__ gor, long long n, int R) {
    // one thread per row: p = v W_g ; pdrow = p Delta_new   (RM = compile-time bound on R: arrays stay in registers)
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int g = gor[row];
    const T* Wg = Wmat + (size_t)g * R * R;
    T v[RM], p[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) v[r] = r < R ? dual[row * R + r] : T(0);
#pragma unroll
    for (int c = 0; c < RM; ++c) {
        T s = T(0);
        if (c < R) {
#pragma unroll
            for (int r = 0; r < RM; ++r)
                if (r < R) s = fma(v[r], __ldg(Wg + r * R + c), s);
        }
        p[c] = s;
    }
    if (basis) {
#pragma unroll
        for (int c = 0; c < RM; ++c)
            if (c < R) basis[row * R + c] = p[c];
    }
#pragma unroll
    for (int c = 0; c < RM; ++c) {
        if (c < R) {
            T s = T(0);
#pragma unroll
            for (int r = 0; r < RM; ++r)
                if (r < R) s = fma(p[r], __ldg(Delta_new + r * R + c), s);
            pd[row * R + c] = s;
            dual[row * R + c] = v[c] - s;
        }
    }
}

}  // namespace

extern "C" {

int b2_gram(const void* M, long long n, int R, int ld, void* G, int dtype, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    int blocks = (int)((n + 63) / 64);  // a factor matrix has few rows (K): spread them instead of walking them in one CTA
    const int cap = b2_num_sms() * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    B2_REQUIRE(ws_bytes >= (size_t)blocks * R * R * sizeof(double), "b2_gram workspace too small");
    B2_DISPATCH_DTYPE(dtype, {
        gram_partial_kernel<T><<<blocks, 256, 0, st>>>((const T*)M, n, R, ld, (double*)ws);
        B2_LAUNCH_CHECK();
        gram_final_kernel<T><<<(R * R + 255) / 256, 256, 0, st>>>((const double*)ws, blocks, R * R, (T*)G);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_scale_gram(const void* G, const void* A, int n_groups, int R, void* lhs, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_groups == 0) return B2_OK;
    const long long total = (long long)n_groups * R * R;
    int blocks = (int)((total + 255) / 256);
    if (blocks > b2_num_sms() * 8) blocks = b2_num_sms() * 8;
    B2_DISPATCH_DTYPE(dtype, {
        scale_gram_kernel<T><<<blocks, 256, 0, st>>>((const T*)G, (const T*)A, n_groups, R, (T*)lhs);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_rho_from_trace(const void* lhs, int n_groups, int R, double scale, void* rho, void* rho_max, int dtype,
                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_groups == 0) {
        // an empty shard contributes the identity of the MAX all-reduce that follows (a value left over from the
        // previous mode's update would otherwise take part in it)
        if (rho_max) {
            B2_DISPATCH_DTYPE(dtype, {
                fill_kernel<T><<<1, 32, 0, st>>>((T*)rho_max, 1, -INFINITY);
                B2_LAUNCH_CHECK();
            });
        }
        return B2_OK;
    }
    B2_DISPATCH_DTYPE(dtype, {
        rho_trace_kernel<T><<<(n_groups + 127) / 128, 128, 0, st>>>((const T*)lhs, n_groups, R, scale, (T*)rho);
        B2_LAUNCH_CHECK();
        if (rho_max) {
            max_kernel<T><<<1, 1024, 0, st>>>((const T*)rho, n_groups, (T*)rho_max);
            B2_LAUNCH_CHECK();
        }
    });
    return B2_OK;
}

int b2_factor_batch(const void* lhs, int n_groups, int R, void* rho, const void* rho_max, int n_reg, double l2,
                    void* Minv, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    const size_t smem = (size_t)kFactorWarps * 2 * R * (R | 1) * sizeof(double);
    B2_DISPATCH_DTYPE(dtype, {
        auto kern = factor_batch_kernel<T>;
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(n_groups + kFactorWarps - 1) / kFactorWarps, kFactorWarps * 32, smem, st>>>(
            (const T*)lhs, n_groups, R, (T*)rho, (const T*)rho_max, n_reg, l2, (T*)Minv);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_slice_cross(const void* B, const void* Y, const int64_t* row_off, int n_groups, int R, const void* CtC,
                   void* cross, void* rhs, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    B2_DISPATCH_DTYPE(dtype, {
        slice_cross_kernel<T><<<n_groups, 256, 0, st>>>((const T*)B, (const T*)Y, row_off, R, (const T*)CtC, (T*)cross,
                                                        (T*)rhs);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_rowscale(const void* B, const void* A, const int32_t* group_of_row, long long n, int R, void* W, int ldw,
                int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return B2_OK;
    long long blocks = (n * R + 255) / 256;
    if (blocks > b2_num_sms() * 16) blocks = b2_num_sms() * 16;
    B2_DISPATCH_DTYPE(dtype, {
        rowscale_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)B, (const T*)A, group_of_row, n, R, (T*)W, ldw);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_pf2_apply(void* pd, void* dual, void* basis, const void* Wmat, const void* Delta_new,
                 const int32_t* group_of_row, long long n, int R, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return B2_OK;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    const int grid = (int)((n + 127) / 128);
    B2_DISPATCH_DTYPE(dtype, B2_DISPATCH_RANK(R, {
        pf2_apply_kernel<T, RM><<<grid, 128, 0, st>>>((T*)pd, (T*)dual, (T*)basis, (const T*)Wmat,
                                                      (const T*)Delta_new, group_of_row, n, R);
        B2_LAUNCH_CHECK();
    }));
    return B2_OK;
}

}  // extern "C"
