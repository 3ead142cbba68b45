// Fused ADMM step (x-update + elementwise prox + dual update), the L2-ball column projection and the fused
// reductions behind the feasibility gaps / loss.  One thread owns one row of the packed (n x R) state, so the whole
// row stays in registers between the solve, the prox of every penalty and the dual update.
#include "admm_common.cuh"

namespace {

template <typename T, int RM>
__global__ void __launch_bounds__(128)
admm_solve_kernel(long long n, int R, const T* __restrict__ rhs, const T* __restrict__ rhs_scale, int group_mode,
                  const int32_t* __restrict__ gor, const T* __restrict__ rho, const T* __restrict__ Minv, PenArgs pa,
                  T* __restrict__ x) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int g = group_mode == B2_GROUP_SINGLE ? 0 : (group_mode == B2_GROUP_INDEXED ? gor[row] : (int)row);
    const T rg = rho[g];
    const size_t base = (size_t)row * R;
    T s[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        T v = T(0);
        if (r < R) {
            v = rhs[base + r];
            if (rhs_scale) v *= rhs_scale[(size_t)g * R + r];
        }
        s[r] = v;
    }
    // rho * sum_p (aux_p - dual_p), accumulated in penalty order like the reference's `sum_shifted_aux += ...`
    if (pa.n_pen > 0) {
        T sh[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) sh[r] = T(0);
        for (int p = 0; p < pa.n_pen; ++p) {
            const T* ax = (const T*)pa.aux[p];
            const T* du = (const T*)pa.dual[p];
#pragma unroll
            for (int r = 0; r < RM; ++r)
                if (r < R) sh[r] += ax[base + r] - du[base + r];
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) s[r] = rg * sh[r] + s[r];
    }
    const T* Mg = Minv + (size_t)g * R * R;
    T xr[RM];
#pragma unroll
    for (int c = 0; c < RM; ++c) xr[c] = T(0);
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        if (r < R) {
#pragma unroll
            for (int c = 0; c < RM; ++c)
                if (c < R) xr[c] = fma(s[r], __ldg(Mg + r * R + c), xr[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < RM; ++c)
        if (c < R) x[base + c] = xr[c];
    for (int p = 0; p < pa.n_pen; ++p) {
        T* ax = (T*)pa.aux[p];
        T* du = (T*)pa.dual[p];
        const int kind = pa.kind[p];
        const bool elementwise = kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1;
#pragma unroll
        for (int c = 0; c < RM; ++c) {
            if (c < R) {
                const T v = xr[c] + du[base + c];
                if (elementwise) {
                    const T z = prox_elem<T>(v, kind, pa.nn[p], (T)pa.p0[p], (T)pa.p1[p], rg);
                    ax[base + c] = z;
                    du[base + c] = v - z;  // x - (aux - dual) = (x + dual) - aux
                } else {
                    du[base + c] = v;
                }
            }
        }
    }
}

template <typename T>
__global__ void prox_elementwise_kernel(const T* __restrict__ v, T* __restrict__ out, long long n, int kind, int nn,
                                        T p0, T p1, T rho) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = prox_elem<T>(v[i], kind, nn, p0, p1, rho);
}

// L2 ball: CTA per group; pass 1 column sums of squares of clip(V), pass 2 scale + dual update.
template <typename T>
__global__ void prox_l2ball_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int R,
                                   T bound, int nn, double* __restrict__ colsq_io, int phase) {
    __shared__ double colsq[B2_MAX_RANK];
    __shared__ double part[8][B2_MAX_RANK];
    const int g = blockIdx.x;
    const long long r0 = row_off[g], r1 = row_off[g + 1];
    const long long cnt = (r1 - r0) * R;
    const T* V = dual + r0 * R;
    // thread t handles flat elements t, t + blockDim, ...; column = e % R.  blockDim is a multiple of R-lcm? no:
    // accumulate per-thread per-column requires fixed column per thread => stride must be a multiple of R.
    const int tpr = blockDim.x / R * R;  // active threads: multiple of R so a thread always sees the same column
    double acc = 0.0;
    if ((int)threadIdx.x < tpr && phase != 2) {
        for (long long e = threadIdx.x; e < cnt; e += tpr) {
            T v = V[e];
            if (nn && v < T(0)) v = T(0);
            acc += (double)v * (double)v;
        }
    }
    // reduce threads with equal column: column of thread t is t % R
    // stage through shared memory: sum over t' = c, c+R, c+2R, ... (fixed order)
    extern __shared__ double l2_scratch[];
    l2_scratch[threadIdx.x] = ((int)threadIdx.x < tpr) ? acc : 0.0;
    __syncthreads();
    if ((int)threadIdx.x < R) {
        double s = 0.0;
        if (phase == 2) {
            s = colsq_io[(size_t)g * R + threadIdx.x];  // sums of squares reduced over all ranks by the caller
        } else {
            for (int t = threadIdx.x; t < tpr; t += R) s += l2_scratch[t];
            if (phase == 1) colsq_io[(size_t)g * R + threadIdx.x] = s;
        }
        colsq[threadIdx.x] = s;
    }
    __syncthreads();
    (void)part;
    if (phase == 1) return;
    if ((int)threadIdx.x < tpr) {
        const int c = threadIdx.x % R;
        T nrm = (T)sqrt(colsq[c]);
        if (nrm < bound) nrm = bound;  // clip(norms, bound, inf)
        T* Ax = aux + r0 * R;
        T* Du = dual + r0 * R;
        for (long long e = threadIdx.x; e < cnt; e += tpr) {
            const T v = Du[e];
            T w = v;
            if (nn && w < T(0)) w = T(0);
            const T z = w * bound / nrm;  // (M * bound) / norms, same association as the reference
            Ax[e] = z;
            Du[e] = v - z;
        }
    }
}

// out[0] = sum (x-y)^2, out[1] = sum x^2, out[2] = sum |x| ; two stages, fixed order
template <typename T>
__global__ void reduce_stats_kernel(const T* __restrict__ x, const T* __restrict__ y, long long n,
                                    double* __restrict__ part) {
    __shared__ double scratch[32];
    double d2 = 0.0, x2 = 0.0, ab = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double xv = (double)x[i];
        if (y) {
            const double d = xv - (double)y[i];
            d2 += d * d;
        }
        x2 += xv * xv;
        ab += fabs(xv);
    }
    d2 = block_sum(d2, scratch);
    x2 = block_sum(x2, scratch);
    ab = block_sum(ab, scratch);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 3 + 0] = d2;
        part[blockIdx.x * 3 + 1] = x2;
        part[blockIdx.x * 3 + 2] = ab;
    }
}

__global__ void reduce_stats_final_kernel(const double* __restrict__ part, int blocks, int width,
                                          double* __restrict__ out) {
    __shared__ double scratch[32];
    for (int k = 0; k < width; ++k) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < blocks; b += blockDim.x) acc += part[b * width + k];
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) out[k] = acc;
    }
}

// fit terms: block-strided over groups
template <typename T>
__global__ void fit_terms_kernel(const T* __restrict__ rhs, const T* __restrict__ cross, const T* __restrict__ A,
                                 int n_groups, int R, double* __restrict__ part) {
    // one warp per slice: the R x R cross matrix is read with coalesced loads (lane = flat element), a_g is held one
    // entry per lane and broadcast with shuffles (a thread per slice walked 2 KB-strided rows: 49 us at 4 096 slices)
    __shared__ double scratch[32];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    double inner = 0.0, quad = 0.0;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const double al = lane < R ? (double)A[(size_t)g * R + lane] : 0.0;
        if (lane < R) inner += (double)rhs[(size_t)g * R + lane] * al;
        const T* cr = cross + (size_t)g * R * R;
        for (int e0 = 0; e0 < R * R; e0 += 32) {
            const int e = e0 + lane;
            const int r = e < R * R ? e / R : 0, c = e < R * R ? e - r * R : 0;
            const double ar = __shfl_sync(0xffffffffu, al, r), ac = __shfl_sync(0xffffffffu, al, c);
            if (e < R * R) quad += ar * ((double)cr[e] * ac);
        }
    }
    inner = block_sum(inner, scratch);
    quad = block_sum(quad, scratch);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 2 + 0] = inner;
        part[blockIdx.x * 2 + 1] = quad;
    }
}

}  // namespace

extern "C" {

int b2_admm_solve(long long n, int R, const void* rhs, const void* rhs_scale, int group_mode,
                  const int32_t* group_of_row, const void* rho, const void* Minv, const b2_penalty_desc* pens, int n_pen,
                  void* x, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n == 0) return B2_OK;  // empty shard: zero-size buffers may arrive as NULL
    B2_REQUIRE(group_mode != B2_GROUP_INDEXED || group_of_row != nullptr, "group_of_row required");
    PenArgs pa;
    {
        const int rc = b2_pack_penalties(pens, n_pen, &pa);
        if (rc != B2_OK) return rc;
    }
    const int grid = (int)((n + 127) / 128);
    B2_DISPATCH_DTYPE(dtype, B2_DISPATCH_RANK(R, {
        admm_solve_kernel<T, RM><<<grid, 128, 0, st>>>(n, R, (const T*)rhs, (const T*)rhs_scale, group_mode,
                                                       group_of_row, (const T*)rho, (const T*)Minv, pa, (T*)x);
        B2_LAUNCH_CHECK();
    }));
    return B2_OK;
}

int b2_prox_elementwise(const void* v, void* out, long long n, int kind, int non_negativity, double p0, double p1,
                        double rho, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(kind == B2_PEN_NONNEG || kind == B2_PEN_BOX || kind == B2_PEN_L1, "kind %d is not elementwise", kind);
    if (n == 0) return B2_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > b2_num_sms() * 16) blocks = b2_num_sms() * 16;
    B2_DISPATCH_DTYPE(dtype, {
        prox_elementwise_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)v, (T*)out, n, kind, non_negativity, (T)p0,
                                                                (T)p1, (T)rho);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_prox_l2ball(void* aux, void* dual, const int64_t* row_off, int n_groups, int R, double bound, int non_negativity,
                   double* colsq_io, int phase, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    B2_REQUIRE(phase >= 0 && phase <= 2 && (phase == 0 || colsq_io), "b2_prox_l2ball: phase 1/2 need colsq_io");
    if (n_groups == 0) return B2_OK;
    const int threads = 256;
    B2_DISPATCH_DTYPE(dtype, {
        prox_l2ball_kernel<T><<<n_groups, threads, threads * sizeof(double), st>>>((T*)aux, (T*)dual, row_off, R,
                                                                                   (T)bound, non_negativity, colsq_io,
                                                                                   phase);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

int b2_reduce_stats(const void* x, const void* y, long long n, double* out, int dtype, void* ws, size_t ws_bytes,
                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n + 1023) / 1024;
    if (blocks > b2_num_sms() * 4) blocks = b2_num_sms() * 4;
    if (blocks < 1) blocks = 1;
    B2_REQUIRE(ws_bytes >= (size_t)blocks * 3 * sizeof(double), "b2_reduce_stats workspace too small");
    B2_DISPATCH_DTYPE(dtype, {
        reduce_stats_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)x, (const T*)y, n, (double*)ws);
        B2_LAUNCH_CHECK();
    });
    reduce_stats_final_kernel<<<1, 256, 0, st>>>((const double*)ws, (int)blocks, 3, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int b2_fit_terms(const void* rhs, const void* cross, const void* A, int n_groups, int R, double* out, int dtype,
                 void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = (n_groups + 3) / 4;  // 4 warps per block, one warp per slice
    if (blocks > b2_num_sms() * 4) blocks = b2_num_sms() * 4;
    if (blocks < 1) blocks = 1;
    B2_REQUIRE(ws_bytes >= (size_t)blocks * 2 * sizeof(double), "b2_fit_terms workspace too small");
    B2_DISPATCH_DTYPE(dtype, {
        fit_terms_kernel<T><<<blocks, 128, 0, st>>>((const T*)rhs, (const T*)cross, (const T*)A, n_groups, R,
                                                    (double*)ws);
        B2_LAUNCH_CHECK();
    });
    reduce_stats_final_kernel<<<1, 256, 0, st>>>((const double*)ws, blocks, 2, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // extern "C"
