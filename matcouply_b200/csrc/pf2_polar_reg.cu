// PARAFAC2 polar step with the Jacobi iteration held in REGISTERS (one warp per slice, lane i owns row i of G and of
// the eigenvector matrix Q).  Same mathematics and interface as pf2_polar_warp.cu (reference penalties.py:1233-1245:
// P_i = polar(V_i Delta^T) = V_i W_i, W_i = Delta^T (Delta S_i Delta^T)^(-1/2), summand rho_i W_i^T S_i).
//
// Why: the shared-memory warp kernel is bound by shared-memory wavefronts — every Jacobi round re-reads and
// re-writes G twice and Q once (~290 wavefront-cycles per round at R = 20, measured 325 cycles per round per SM).
// With the rank padded to a compile-time RE and the column registers rotated one position per round (see the sweep
// loop), the column pass (G <- G J, Q <- Q J) indexes registers at compile-time positions and costs no memory
// traffic at all; the row pass (G <- J^T G) is one shuffle exchange of the partner's row; only the R/2 rotation
// pairs (c, s) of a round go through shared memory (one 16-byte store per pair, broadcast reads).  A first version
// that unrolled all RE - 1 rounds was SLOWER than the shared-memory kernel (170 KB of straight-line code per sweep:
// instruction-fetch bound); the rotating-register formulation keeps the round body at ~450 instructions.
// Measured at R = 20, 16 384 slices: 1.87 ms cold / 0.73 ms warm vs 2.89 / 1.26 ms (tools/bench_polar.py).
// The R x R products around the iteration keep one operand column in registers and read the other with 128-bit
// broadcasts.
#include "common.cuh"

namespace {

constexpr int kRegWarps = 4;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int RE>
struct Layout {
    static constexpr int LD = RE + 2;      // rows 16-byte aligned (128-bit broadcasts); RE % 4 == 0 => LD / 2 odd
    static constexpr int MS = RE * LD;
};

// dot of a shared-memory row (16-byte aligned, RE doubles) with a register column
template <int RE>
__device__ __forceinline__ double row_dot(const double* __restrict__ row, const double (&b)[RE]) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < RE; k += 2) {
        const double2 a = *(const double2*)(row + k);
        s0 = fma(a.x, b[k], s0);
        s1 = fma(a.y, b[k + 1], s1);
    }
    return s0 + s1;
}

// c[i] += sum_k A[k][i] * bk(k)   (A^T B, column `lane` of the product accumulated in registers)
template <int RE, class FB>
__device__ __forceinline__ void tn_acc(const double* __restrict__ A, FB bk, double (&c)[RE]) {
    constexpr int LD = Layout<RE>::LD;
#pragma unroll
    for (int i = 0; i < RE; ++i) c[i] = 0.0;
#pragma unroll 2
    for (int k = 0; k < RE; ++k) {
        const double b = bk(k);
#pragma unroll
        for (int i = 0; i < RE; i += 2) {
            const double2 a = *(const double2*)(A + k * LD + i);
            c[i] = fma(a.x, b, c[i]);
            c[i + 1] = fma(a.y, b, c[i + 1]);
        }
    }
}

template <typename T, int RE>
__global__ void __launch_bounds__(32 * kRegWarps, RE >= 24 ? 3 : (RE >= 8 ? 4 : 1))  // CTAs per SM that fit in shared memory
pf2_polar_reg_kernel(const T* __restrict__ S, const T* __restrict__ Delta, const T* __restrict__ rho, int n_groups,
                     int R, T* __restrict__ Wmat, double* __restrict__ num_part, double* __restrict__ Qstore,
                     int warm) {
    constexpr int LD = Layout<RE>::LD, MS = Layout<RE>::MS, M = RE - 1, HALF = RE / 2;
    extern __shared__ __align__(16) double pr_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* Dm = pr_smem;                                   // Delta, zero padded to RE x RE
    double* Gs = pr_smem + MS + (size_t)w * (3 * MS + 64);
    double* Qs = Gs + MS;
    double* Ts = Qs + MS;
    double2* cs2 = (double2*)(Ts + MS);                     // {c, s} of the HALF pairs of a round
    double* isq = Ts + MS + 32;                             // lam^-1/2 per eigenvalue
    const int RR = R * R;
    for (int e = threadIdx.x; e < RE * 32; e += blockDim.x) {
        const int i = e >> 5, j = e & 31;
        if (j < RE) Dm[i * LD + j] = (i < R && j < R) ? (double)Delta[i * R + j] : 0.0;
    }
    __syncthreads();
    const bool col = lane < RE;   // lanes that own a (possibly padded) row / column
    const bool act = lane < R;

    for (int g = blockIdx.x * kRegWarps + w; g < n_groups; g += gridDim.x * kRegWarps) {
        const T* Sg = S + (size_t)g * RR;
        if (col)
            for (int i = 0; i < RE; ++i) Gs[i * LD + lane] = (i < R && act) ? (double)Sg[i * R + lane] : 0.0;
        __syncwarp();
        {   // Ts = Delta S ; Gs = Ts Delta^T
            double b[RE];
            if (col) {
#pragma unroll
                for (int k = 0; k < RE; ++k) b[k] = Gs[k * LD + lane];
                for (int i = 0; i < RE; ++i) Ts[i * LD + lane] = row_dot<RE>(Dm + i * LD, b);
            }
            __syncwarp();
            if (col) {
#pragma unroll
                for (int k = 0; k < RE; ++k) b[k] = Dm[lane * LD + k];
                for (int i = 0; i < RE; ++i) Gs[i * LD + lane] = row_dot<RE>(Ts + i * LD, b);
            }
            __syncwarp();
        }
        if (warm) {
            // G <- Q0^T G Q0 with the eigenvectors of the previous call (nearly diagonal: 1-3 sweeps suffice)
            const double* Q0 = Qstore + (size_t)g * RR;
            if (col)
                for (int i = 0; i < RE; ++i)
                    Qs[i * LD + lane] = (i < R && act) ? Q0[i * R + lane] : ((i == lane) ? 1.0 : 0.0);
            __syncwarp();
            double b[RE];
            if (col) {
#pragma unroll
                for (int k = 0; k < RE; ++k) b[k] = Qs[k * LD + lane];
                for (int i = 0; i < RE; ++i) Ts[i * LD + lane] = row_dot<RE>(Gs + i * LD, b);   // Ts = G Q0
            }
            __syncwarp();
            if (col) {
                tn_acc<RE>(Qs, [&](int k) { return Ts[k * LD + lane]; }, b);                     // Q0^T (G Q0)
#pragma unroll
                for (int i = 0; i < RE; ++i) Gs[i * LD + lane] = b[i];
            }
            __syncwarp();
        }
        // rows into registers (symmetrised); padded rows / columns are zero and stay zero
        double gr[RE], qr[RE];
#pragma unroll
        for (int j = 0; j < RE; ++j) {
            gr[j] = col ? 0.5 * (Gs[lane * LD + j] + Gs[j * LD + lane]) : 0.0;
            qr[j] = col ? (warm ? Qs[lane * LD + j] : (j == lane ? 1.0 : 0.0)) : 0.0;
        }
        // Round-robin tournament on RE players (player M = RE - 1 fixed): in round t the logical indices (t + k) % M and
        // (t - k) % M meet (k = 1..HALF-1), and t meets M.  The COLUMN registers are rotated by one position per round
        // (logical column (j + t) % M sits in register j < M; free: the passes simply write their results one slot
        // down), so the pairs of every round sit at the compile-time positions {k, M - k} / {0, M} and the round body
        // is ONE piece of code in a runtime loop.  Rows stay in their lanes; after the M rounds of a sweep the
        // registers are back in logical order.
        double diag = 0.0;
        for (int sweep = 0; sweep < 40 && R > 1; ++sweep) {
            double off = 0.0;
            diag = 0.0;
#pragma unroll
            for (int j = 0; j < RE; ++j) {
                const bool own = j == lane;
                off = own ? off : fma(gr[j], gr[j], off);
                diag = own ? gr[j] : diag;
            }
            const double dg = wsum(diag * diag);
            off = wsum(off);
            if (!(off > 1e-26 * dg && off > 0.0)) break;  // off/diag <= 1e-13: eigenvectors at round-off
#pragma unroll 1
            for (int t = 0; t < M; ++t) {
                int pt = 2 * t - lane;                       // logical partner of this lane's row
                pt = pt < 0 ? pt + M : (pt >= M ? pt - M : pt);
                pt = (lane == M) ? t : ((lane == t) ? M : pt);
                if (!col) pt = lane;
                int pd = lane - t;                           // register positions of G[l][l] and G[l][partner]
                pd = pd < 0 ? pd + M : pd;
                pd = (lane == M) ? M : pd;
                int po = pt - t;
                po = po < 0 ? po + M : po;
                po = (pt == M) ? M : po;
                double d = 0.0, o = 0.0;
#pragma unroll
                for (int j = 0; j < RE; ++j) {
                    d = (j == pd) ? gr[j] : d;
                    o = (j == po) ? gr[j] : o;
                }
                const bool low = pd < po;                    // first element of the pair {k, M - k} (or 0 of {0, M})
                const double dp = shfl_d(d, pt), op = shfl_d(o, pt);
                const double app = low ? d : dp, aqq = low ? dp : d, apq = low ? o : op;
                double c = 1.0, s = 0.0;
                if (fabs(apq) > 1e-300 && apq * apq > 1e-40 * fabs(app * aqq)) {
                    // tan(theta) = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 apq), written with ONE
                    // division: numerator and denominator times 2 |apq| (the dependent div/sqrt chain is what bounds a round)
                    const double dl = aqq - app, a2 = 2.0 * apq;
                    const double num = (dl >= 0.0) ? a2 : -a2;   // sgn(tau) * 2 |apq| = sgn(dl) * 2 apq
                    const double tt = num / (fabs(dl) + sqrt(fma(dl, dl, a2 * a2)));
                    c = rsqrt(fma(tt, tt, 1.0));
                    s = tt * c;
                }
                if (col && low) cs2[pd] = make_double2(c, s);   // slot = the pair's first register position
                __syncwarp();
                // column pass, registers only: G <- G J, Q <- Q J (Q written one slot down = next round's order)
                double g2[RE], q2[RE];
#pragma unroll
                for (int k = 0; k < HALF; ++k) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int p = k, q = (k == 0) ? M : M - k;
                    const double2 r = cs2[k];
                    g2[p] = r.x * gr[p] - r.y * gr[q];
                    g2[q] = r.y * gr[p] + r.x * gr[q];
                    const int pn = (p == 0) ? M - 1 : p - 1, qn = (q == M) ? M : q - 1;
                    q2[pn] = r.x * qr[p] - r.y * qr[q];
                    q2[qn] = r.y * qr[p] + r.x * qr[q];
                }
                // row pass: G <- J^T G, one exchange of the partner's row, written one slot down
                const double sg = low ? -s : s;
#pragma unroll
                for (int j = 0; j < RE; ++j) {
                    const double other = shfl_d(g2[j], pt);
                    const int jn = (j == M) ? M : ((j == 0) ? M - 1 : j - 1);
                    gr[jn] = fma(sg, other, c * g2[j]);
                }
#pragma unroll
                for (int j = 0; j < RE; ++j) qr[j] = q2[j];
                __syncwarp();  // cs2 is rewritten by the next round
            }
        }
#pragma unroll
        for (int j = 0; j < RE; ++j) diag = (j == lane) ? gr[j] : diag;

        // lam^-1/2 (directions with lam <= eps * lam_max dropped); eigenvectors back to shared memory
        const double lam = act ? diag : 0.0;
        const double lmax = wmax(lam);
        isq[lane] = (act && lam > 1e-28 * lmax && lam > 0.0) ? 1.0 / sqrt(lam) : 0.0;
        if (col) {
#pragma unroll
            for (int j = 0; j < RE; ++j) Qs[lane * LD + j] = qr[j];
        }
        __syncwarp();
        if (Qstore && act) {
            double* Qo = Qstore + (size_t)g * RR;
            for (int i = 0; i < R; ++i) Qo[i * R + lane] = Qs[i * LD + lane];
        }
        double b[RE];
        if (col) {   // H = Q lam^-1/2 Q^T -> Gs
#pragma unroll
            for (int k = 0; k < RE; ++k) b[k] = Qs[lane * LD + k] * isq[k];
            for (int i = 0; i < RE; ++i) Gs[i * LD + lane] = row_dot<RE>(Qs + i * LD, b);
        }
        __syncwarp();
        if (col) {   // W = Delta^T H
            tn_acc<RE>(Dm, [&](int k) { return Gs[k * LD + lane]; }, b);
            if (act) {
                T* Wg = Wmat + (size_t)g * RR;
#pragma unroll
                for (int i = 0; i < RE; ++i)
                    if (i < R) Wg[i * R + lane] = (T)b[i];
            }
        }
        // num = rho W^T S = rho H (Delta S)
        if (col)
            for (int i = 0; i < RE; ++i) Qs[i * LD + lane] = (i < R && act) ? (double)Sg[i * R + lane] : 0.0;
        __syncwarp();
        if (col) {
#pragma unroll
            for (int k = 0; k < RE; ++k) b[k] = Qs[k * LD + lane];
            for (int i = 0; i < RE; ++i) Ts[i * LD + lane] = row_dot<RE>(Dm + i * LD, b);
        }
        __syncwarp();
        if (col) {
#pragma unroll
            for (int k = 0; k < RE; ++k) b[k] = Ts[k * LD + lane];
            if (act) {
                const double rg = (double)rho[g];
                double* Ng = num_part + (size_t)g * RR;
                for (int i = 0; i < R; ++i) Ng[i * R + lane] = rg * row_dot<RE>(Gs + i * LD, b);
            }
        }
        __syncwarp();
    }
}

template <typename T, int RE>
int launch_polar_reg(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                     void* num_part, void* Qstore, int warm, cudaStream_t st) {
    const size_t smem = (size_t)(Layout<RE>::MS + kRegWarps * (3 * Layout<RE>::MS + 64)) * sizeof(double);
    auto kern = pf2_polar_reg_kernel<T, RE>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    const int ctas_needed = (n_groups + kRegWarps - 1) / kRegWarps;
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const int grid = ctas_needed < b2_num_sms() * per_sm ? ctas_needed : b2_num_sms() * per_sm;
    kern<<<grid, 32 * kRegWarps, smem, st>>>((const T*)S, (const T*)Delta, (const T*)rho, n_groups, R, (T*)Wmat,
                                             (double*)num_part, (double*)Qstore, warm);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // namespace

// Returns -1 when the register formulation does not cover this rank (the caller then uses pf2_polar_warp.cu).
int b2_pf2_polar_reg(const void* S, const void* Delta, const void* rho, int n_groups, int R, void* Wmat,
                     void* num_part, void* Qstore, int warm, int dtype, cudaStream_t st) {
    if (R < 2 || R > B2_POLAR_REG_MAX_RANK) return -1;
    const int RE = (R + 3) / 4 * 4;
#define B2_POLAR_CASE(N)                                                                                          \
    if (RE == N) {                                                                                                \
        B2_DISPATCH_DTYPE(dtype, return (launch_polar_reg<T, N>(S, Delta, rho, n_groups, R, Wmat, num_part,       \
                                                                Qstore, warm, st)));                              \
    }
    B2_POLAR_CASE(4)
    B2_POLAR_CASE(8)
    B2_POLAR_CASE(12)
    B2_POLAR_CASE(16)
    B2_POLAR_CASE(20)
    B2_POLAR_CASE(24)
#undef B2_POLAR_CASE
    return -1;
}
