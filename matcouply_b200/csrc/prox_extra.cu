// Column-coupled proximal operators of the "next" penalties:
//   UnitSimplex            reference penalties.py:928-980   (per column: bisection for the Lagrange multiplier)
//   TotalVariationPenalty  reference penalties.py:750-841   (per column: Condat's direct 1-D TV algorithm + L1 shrinkage)
//   GeneralizedL2Penalty   reference penalties.py:595-747   (per matrix: U diag(rho/2 / (s + rho/2)) U^T V)
// Convention of all b2_prox_* entry points: the pre-image V = x + dual arrives in `dual` (written by b2_admm_solve /
// b2_pf2_rowpass), the kernel writes aux = prox(V) and dual = V - aux (decomposition.py:275-285).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kSimplexThreads = 256;
constexpr int kSimplexWarps = kSimplexThreads / 32;

// ---------------------------------------------------------------------------------------------------------
// UnitSimplex.  One CTA per group; lane = column, warps stride over the rows (coalesced row reads).  All R columns
// bisect together; every iteration is one pass over the group's V (staged in shared memory when it fits).
// The bracket and the bisection loop restate penalties.py:953-969 + scipy.optimize.bisect (xtol 2e-12, rtol 4 eps,
// maxiter 100: dm *= .5; xm = xa + dm; if f(xm) * f(xa) >= 0: xa = xm; stop if f(xm) == 0 or |dm| < xtol + rtol |xm|).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kSimplexThreads)
prox_simplex_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int R,
                    int stage_rows) {
    extern __shared__ double sx_stage[];  // stage_rows x R doubles (0 rows: read V from global every pass)
    __shared__ double part[kSimplexWarps][32];
    __shared__ double part2[kSimplexWarps][32];
    __shared__ int s_active;
    const int g = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long r0 = row_off[g];
    const int rows = (int)(row_off[g + 1] - r0);
    if (rows == 0) return;
    T* V = dual + r0 * R;
    T* Z = aux + r0 * R;
    const bool col = lane < R;
    const bool staged = rows <= stage_rows;
    if (staged) {
        for (int e = threadIdx.x; e < rows * R; e += blockDim.x) sx_stage[e] = (double)V[e];
        __syncthreads();
    }
    auto val = [&](int row) -> double { return staged ? sx_stage[row * R + lane] : (double)V[(size_t)row * R + lane]; };

    // column min / max
    double mn = INFINITY, mx = -INFINITY;
    if (col)
        for (int row = w; row < rows; row += kSimplexWarps) {
            const double v = val(row);
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    part[w][lane] = mn;
    part2[w][lane] = mx;
    __syncthreads();
    for (int u = 0; u < kSimplexWarps; ++u) {
        mn = fmin(mn, part[u][lane]);
        mx = fmax(mx, part2[u][lane]);
    }
    __syncthreads();
    double xa = mn - 1.0;
    double xb = mx;
    xa -= 1e-5;
    xa = fmin(0.9 * xa, 1.1 * xa);
    xb += 1e-5;
    xb = fmax(0.9 * xb, 1.1 * xb);

    // f(mu) = sum_rows max(v - mu, 0) - 1, all columns at once, fixed summation order
    auto f_all = [&](double mu) -> double {
        double acc = 0.0;
        if (col)
            for (int row = w; row < rows; row += kSimplexWarps) acc += fmax(val(row) - mu, 0.0);
        part[w][lane] = acc;
        __syncthreads();
        double s = 0.0;
        for (int u = 0; u < kSimplexWarps; ++u) s += part[u][lane];
        __syncthreads();
        return s - 1.0;
    };
    const double fa = f_all(xa);
    const double fb = f_all(xb);
    double mu = 0.0;
    bool done = !col;
    if (col && fa == 0.0) {
        mu = xa;
        done = true;
    } else if (col && fb == 0.0) {
        mu = xb;
        done = true;
    }
    double dm = xb - xa;
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    for (int it = 0; it < 100; ++it) {
        if (threadIdx.x == 0) s_active = 0;
        __syncthreads();
        if (!done && w == 0) s_active = 1;  // benign race: every writer stores 1
        __syncthreads();
        if (!s_active) break;
        dm *= 0.5;
        const double xm = xa + dm;
        const double fm = f_all(xm);
        if (!done) {
            if (fm * fa >= 0.0) xa = xm;
            if (fm == 0.0 || fabs(dm) < xtol + rtol * fabs(xm)) {
                mu = xm;
                done = true;
            }
        }
    }
    if (!done) mu = xa + dm;  // scipy raises after maxiter; 100 halvings always meet xtol first
    if (col)
        for (int row = w; row < rows; row += kSimplexWarps) {
            const double v = val(row);
            const double z = fmax(v - mu, 0.0);
            Z[(size_t)row * R + lane] = (T)z;
            V[(size_t)row * R + lane] = (T)(v - (double)(T)z);
        }
}

// ---------------------------------------------------------------------------------------------------------
// Total variation.  Thread per (group, column): Condat's direct algorithm (L. Condat, "A direct algorithm for 1-D
// total variation denoising", IEEE SPL 20(11), 2013 — the algorithm behind the reference's un-vendored condat_tv
// dependency) for  argmin_x 0.5 ||x - v||^2 + lam TV(x),  lam = 2 reg_strength / rho  (penalties.py:821-823), then the
// optional L1 shrinkage by l1_strength / rho (:824-825).  The minimiser is unique, so every exact algorithm agrees to
// round-off; tests check the KKT certificate.  Sequential in the column length by construction.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void prox_tv_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off,
                               int n_groups, int R, const T* __restrict__ rho, int rho_stride, double reg, double l1) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = (int)(tid / 32), c = (int)(tid % 32);
    if (g >= n_groups || c >= R) return;
    const long long r0 = row_off[g];
    const int width = (int)(row_off[g + 1] - r0);
    if (width == 0) return;
    const double rg = (double)rho[(size_t)g * rho_stride];
    const double lambda = 2.0 * reg / rg;
    T* in = dual + r0 * R + c;   // element k at in[k * R]
    T* out = aux + r0 * R + c;
#define IN(k) ((double)in[(size_t)(k) * R])
#define OUT(k, v) out[(size_t)(k) * R] = (T)(v)
    {
        int k = 0, k0 = 0;                       // current sample, start of the current segment
        double umin = lambda, umax = -lambda;    // dual variable bounds
        double vmin = IN(0) - lambda, vmax = IN(0) + lambda;  // bounds of the segment's value
        int kplus = 0, kminus = 0;               // last positions where umax = -lambda, umin = lambda
        const double twolambda = 2.0 * lambda, minlambda = -lambda;
        for (;;) {
            bool finished = false;
            while (k == width - 1) {  // right boundary
                if (umin < 0.0) {     // vmin too high: negative jump
                    do { OUT(k0, vmin); ++k0; } while (k0 <= kminus);
                    k = kminus = k0;
                    vmin = IN(k0);
                    umin = lambda;
                    umax = vmin + umin - vmax;
                } else if (umax > 0.0) {  // vmax too low: positive jump
                    do { OUT(k0, vmax); ++k0; } while (k0 <= kplus);
                    k = kplus = k0;
                    vmax = IN(k0);
                    umax = minlambda;
                    umin = vmax + umax - vmin;
                } else {
                    vmin += umin / (double)(k - k0 + 1);
                    do { OUT(k0, vmin); ++k0; } while (k0 <= k);
                    finished = true;
                    break;
                }
            }
            if (finished) break;
            const double nxt = IN(k + 1);
            if ((umin += nxt - vmin) < minlambda) {  // negative jump
                do { OUT(k0, vmin); ++k0; } while (k0 <= kminus);
                k = kplus = kminus = k0;
                vmin = IN(k0);
                vmax = vmin + twolambda;
                umin = lambda;
                umax = minlambda;
            } else if ((umax += nxt - vmax) > lambda) {  // positive jump
                do { OUT(k0, vmax); ++k0; } while (k0 <= kplus);
                k = kplus = kminus = k0;
                vmax = IN(k0);
                vmin = vmax - twolambda;
                umin = lambda;
                umax = minlambda;
            } else {  // no jump
                ++k;
                if (umin >= lambda) {
                    kminus = k;
                    vmin += (umin - lambda) / (double)(kminus - k0 + 1);
                    umin = lambda;
                }
                if (umax <= minlambda) {
                    kplus = k;
                    vmax += (umax + lambda) / (double)(kplus - k0 + 1);
                    umax = minlambda;
                }
            }
        }
    }
    const double thr = l1 / rg;
    for (int k = 0; k < width; ++k) {
        const double v = IN(k);
        double z = (double)out[(size_t)k * R];
        if (l1 != 0.0) {
            const double a = fabs(z) - thr;
            const double sgn = z > 0.0 ? 1.0 : (z < 0.0 ? -1.0 : 0.0);
            z = sgn * (a > 0.0 ? a : 0.0);
            OUT(k, z);
        }
        in[(size_t)k * R] = (T)(v - (double)(T)z);
    }
#undef IN
#undef OUT
}

// part[g] = sum over columns and rows of |x[k+1] - x[k]| inside group g (TotalVariationPenalty._penalty, :832)
template <typename T>
__global__ void __launch_bounds__(256)
tv_norm_kernel(const T* __restrict__ x, const int64_t* __restrict__ row_off, int R, double* __restrict__ part) {
    __shared__ double scratch[32];
    const int g = blockIdx.x;
    const long long r0 = row_off[g];
    const long long cnt = (row_off[g + 1] - r0 - 1) * R;  // elements that have a successor row
    const T* X = x + r0 * R;
    double acc = 0.0;
    for (long long e = threadIdx.x; e < cnt; e += blockDim.x) acc += fabs((double)X[e + R] - (double)X[e]);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[g] = acc;
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// Generalized L2.  out_g[j][r] = w_g[j] * sum_l A(l, j) in_g[l][r]  for every group g of J rows, A = U (J x J):
//   TRANS = true :  A(l, j) = U[l][j]  (U^T in)   with w_g[j] = (rho_g / 2) / (s[j] + rho_g / 2)      (:726-728)
//   TRANS = false:  A(l, j) = U[j][l]  (U in)     with w = 1                                           (:729)
// and, with `sub_from` != NULL, sub_from_g <- sub_from_g - out_g (the dual update V - aux).
// 64 output rows x R columns per CTA, 32-deep l tiles staged in shared memory; fp64 accumulation.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool TRANS>
__global__ void __launch_bounds__(256)
gl2_gemm_kernel(const T* __restrict__ U, const T* __restrict__ s, const T* __restrict__ rho, int rho_stride,
                const T* __restrict__ in, T* __restrict__ out, T* __restrict__ sub_from, int J, int R) {
    __shared__ double As[32][65];
    __shared__ double Vs[32][33];
    const int g = blockIdx.y, j0 = blockIdx.x * 64;
    const int jr = threadIdx.x & 63, cg = threadIdx.x >> 6;  // output row, column group (columns cg, cg+4, ...)
    const T* In = in + (size_t)g * J * R;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
    for (int l0 = 0; l0 < J; l0 += 32) {
        for (int e = threadIdx.x; e < 32 * 64; e += 256) {
            int ll, jj;
            if (TRANS) {
                ll = e >> 6;
                jj = e & 63;
            } else {
                jj = e >> 5;
                ll = e & 31;
            }
            const int l = l0 + ll, j = j0 + jj;
            double v = 0.0;
            if (l < J && j < J) v = (double)(TRANS ? U[(size_t)l * J + j] : U[(size_t)j * J + l]);
            As[ll][jj] = v;
        }
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int ll = e >> 5, r = e & 31;
            Vs[ll][r] = (l0 + ll < J && r < R) ? (double)In[(size_t)(l0 + ll) * R + r] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int ll = 0; ll < 32; ++ll) {
            const double a = As[ll][jr];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fma(a, Vs[ll][cg + 4 * c], acc[c]);
        }
        __syncthreads();
    }
    const int j = j0 + jr;
    if (j >= J) return;
    double wj = 1.0;
    if (TRANS) {
        const double half_rho = 0.5 * (double)rho[(size_t)g * rho_stride];
        wj = half_rho / ((double)s[j] + half_rho);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int r = cg + 4 * c;
        if (r < R) {
            const size_t o = ((size_t)g * J + j) * R + r;
            const T z = (T)(wj * acc[c]);
            out[o] = z;
            if (sub_from) sub_from[o] = (T)((double)sub_from[o] - (double)z);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) dot_partial_kernel(const T* __restrict__ x, const T* __restrict__ y, long long n,
                                                           double* __restrict__ part) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i] * (double)y[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}


// ---------------------------------------------------------------------------------------------------------
// PARAFAC2 with frozen basis matrices (Parafac2(update_basis_matrices=False), penalties.py:1231-1248): the summand
// num[g] = rho_g P_g^T V_g of the coordinate-matrix update, and pd = P Delta, dual = V - pd.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
slice_atb_kernel(const T* __restrict__ P, const T* __restrict__ V, const int64_t* __restrict__ row_off, int R,
                 const T* __restrict__ rho, double* __restrict__ num) {
    __shared__ double Ps[32][33];
    __shared__ double Vs[32][33];
    const int g = blockIdx.x;
    const long long r0 = row_off[g], r1 = row_off[g + 1];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // outputs e = tid + 256 u -> (i, j) = (e / R, e % R)
    int oi[4], oj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + 256 * u;
        oi[u] = e < R * R ? e / R : 0;
        oj[u] = e < R * R ? e % R : 0;
    }
    for (long long row0 = r0; row0 < r1; row0 += 32) {
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int rr = e >> 5, c = e & 31;
            const bool ok = row0 + rr < r1 && c < R;
            Ps[rr][c] = ok ? (double)P[(size_t)(row0 + rr) * R + c] : 0.0;
            Vs[rr][c] = ok ? (double)V[(size_t)(row0 + rr) * R + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fma(Ps[rr][oi[u]], Vs[rr][oj[u]], acc[u]);
        __syncthreads();
    }
    const double rg = (double)rho[g];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + 256 * u;
        if (e < R * R) num[(size_t)g * R * R + e] = rg * acc[u];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
pf2_apply_fixed_kernel(T* __restrict__ pd, T* __restrict__ dual, const T* __restrict__ P, const T* __restrict__ Delta,
                       long long n, int R) {
    __shared__ double D[B2_MAX_RANK * B2_MAX_RANK];
    for (int e = threadIdx.x; e < R * R; e += blockDim.x) D[e] = (double)Delta[e];
    __syncthreads();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n * R;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / R;
        const int c = (int)(idx - row * R);
        double s = 0.0;
        for (int k = 0; k < R; ++k) s = fma((double)P[row * R + k], D[k * R + c], s);
        const T z = (T)s;
        pd[idx] = z;
        dual[idx] = (T)((double)dual[idx] - (double)z);
    }
}

}  // namespace

extern "C" {

int b2_prox_simplex(void* aux, void* dual, const int64_t* row_off, int n_groups, int max_rows, int R, int dtype,
                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0) return B2_OK;
    const size_t cap = 160 * 1024;
    int stage_rows = (size_t)max_rows * R * sizeof(double) <= cap ? max_rows : 0;
    const size_t smem = (size_t)stage_rows * R * sizeof(double);
    B2_DISPATCH_DTYPE(dtype, {
        auto kern = prox_simplex_kernel<T>;
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        kern<<<n_groups, kSimplexThreads, smem, st>>>((T*)aux, (T*)dual, row_off, R, stage_rows);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_prox_tv(void* aux, void* dual, const int64_t* row_off, int n_groups, int R, const void* rho, int rho_stride,
               double reg_strength, double l1_strength, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(reg_strength > 0.0 && l1_strength >= 0.0, "TV strength must be > 0 and the L1 strength >= 0");
    if (n_groups == 0) return B2_OK;  // empty shard: zero-size buffers may arrive as NULL
    B2_REQUIRE(rho != nullptr && (rho_stride == 0 || rho_stride == 1), "b2_prox_tv: rho (device) with stride 0 or 1");
    const long long threads = (long long)n_groups * 32;
    const int block = 128;
    B2_DISPATCH_DTYPE(dtype, {
        prox_tv_kernel<T><<<(unsigned)((threads + block - 1) / block), block, 0, st>>>(
            (T*)aux, (T*)dual, row_off, n_groups, R, (const T*)rho, rho_stride, reg_strength, l1_strength);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_tv_norm(const void* x, const int64_t* row_off, int n_groups, int R, double* out, double* part, int dtype,
               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(part != nullptr, "b2_tv_norm needs n_groups doubles of scratch");
    if (n_groups == 0) {
        B2_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
        return B2_OK;
    }
    B2_DISPATCH_DTYPE(dtype, {
        tv_norm_kernel<T><<<n_groups, 256, 0, st>>>((const T*)x, row_off, R, part);
        B2_LAUNCH_CHECK();
    });
    sum_partials_kernel<<<1, 256, 0, st>>>(part, n_groups, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_prox_gl2(void* aux, void* dual, int n_groups, int J, int R, const void* U, const void* s, const void* rho,
                int rho_stride, void* tmp, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0 || J == 0) return B2_OK;  // empty shard: zero-size buffers may arrive as NULL
    B2_REQUIRE(tmp != nullptr, "b2_prox_gl2 needs n_groups * J * R elements of scratch");
    B2_REQUIRE(rho != nullptr && (rho_stride == 0 || rho_stride == 1), "b2_prox_gl2: rho (device) with stride 0 or 1");
    const dim3 grid((J + 63) / 64, n_groups);
    B2_DISPATCH_DTYPE(dtype, {
        gl2_gemm_kernel<T, true><<<grid, 256, 0, st>>>((const T*)U, (const T*)s, (const T*)rho, rho_stride,
                                                       (const T*)dual, (T*)tmp, (T*)nullptr, J, R);
        B2_LAUNCH_CHECK();
        gl2_gemm_kernel<T, false><<<grid, 256, 0, st>>>((const T*)U, (const T*)s, (const T*)rho, rho_stride,
                                                        (const T*)tmp, (T*)aux, (T*)dual, J, R);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_quadform(const void* M, const void* x, int n_groups, int J, int R, double* out, void* tmp, double* part,
                int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0 || J == 0) {  // empty shard: zero-size buffers may arrive as NULL
        B2_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
        return B2_OK;
    }
    B2_REQUIRE(tmp != nullptr && part != nullptr, "b2_quadform needs scratch (n_groups*J*R elements, 256 doubles)");
    const dim3 grid((J + 63) / 64, n_groups);
    const long long n = (long long)n_groups * J * R;
    B2_DISPATCH_DTYPE(dtype, {
        gl2_gemm_kernel<T, false><<<grid, 256, 0, st>>>((const T*)M, (const T*)nullptr, (const T*)nullptr, 0,
                                                        (const T*)x, (T*)tmp, (T*)nullptr, J, R);
        B2_LAUNCH_CHECK();
        dot_partial_kernel<T><<<256, 256, 0, st>>>((const T*)x, (const T*)tmp, n, part);
        B2_LAUNCH_CHECK();
    });
    sum_partials_kernel<<<1, 256, 0, st>>>(part, 256, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_pf2_fixed_basis(void* pd, void* dual, const void* P, const void* Delta, const int64_t* row_off, int n_groups,
                       long long n, int R, const void* rho, double* num_part, int phase, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(phase == 1 || phase == 2, "b2_pf2_fixed_basis: phase 1 (numerator) or 2 (apply)");
    if (n_groups == 0 || n == 0) return B2_OK;
    B2_DISPATCH_DTYPE(dtype, {
        if (phase == 1) {
            slice_atb_kernel<T><<<n_groups, 256, 0, st>>>((const T*)P, (const T*)dual, row_off, R, (const T*)rho,
                                                          num_part);
        } else {
            const long long blocks = (n * R + 255) / 256;
            pf2_apply_fixed_kernel<T><<<(unsigned)(blocks < 65535 ? blocks : 65535), 256, 0, st>>>(
                (T*)pd, (T*)dual, (const T*)P, (const T*)Delta, n, R);
        }
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

}  // extern "C"
