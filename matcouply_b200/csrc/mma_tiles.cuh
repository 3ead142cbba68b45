// Register-resident row tiles for the fp64 tensor-core (DMMA.8x8x4) formulation of the per-slice ADMM row math.
//
// A warp owns an "m-block" of 8 rows of an (n x R) state matrix.  A row vector lives in the MMA accumulator ("D")
// layout: lane = 4g + t holds, for every 8-column block b, the two POSITIONS 8b + 2t and 8b + 2t + 1 of row g.
// Chaining  x = s M  without any shuffle: the D registers of block b are fed back as the A operand of k-steps 2b
// (positions 8b + 2t) and 2b + 1 (positions 8b + 2t + 1) — the k index of an MMA is only a summation label, so any
// consistent assignment works as long as the B operand (rows of M) uses the same one.
//
// Positions vs. columns.  With R = 8 NBF + rem, 1 <= rem <= 4, the last block is a HALF block: its 4 columns sit on
// the even positions (column 8 NBF + t on position 8 NBF + 2t), so that they are covered by ONE k-step (instead of
// two half-empty ones) and the operators need ceil(R/4) k-steps x ceil(R/8) n-blocks.  rem >= 5 pads to a full block.
// Operators (R x R, row-major) are staged in shared memory in position order, zero padded, with row stride
// LDM = 8 NB + 2 words: the B-fragment read of lane (g, t) is word (8b + 2t + e) LDM + 8 nb + g, and 2 LDM = 4
// (mod 16) makes the 16 lanes of a half-warp hit 16 distinct words (64-bit shared loads are served per half-warp).
#pragma once
#include "common.cuh"

template <int NBF_, int HALF_>
struct PosLayout {
    static constexpr int NBF = NBF_, HALF = HALF_;
    static constexpr int NB = NBF + HALF;       // n-blocks
    static constexpr int KS = 2 * NBF + HALF;   // k-steps of 4 positions
    static constexpr int NPOS = 8 * NB;
    static constexpr int LDM = 8 * NB + 2;
    // column held on position p, or -1 (padding)
    __host__ __device__ static __forceinline__ int col_of(int p, int R) {
        const int b = p >> 3, w = p & 7;
        if (b < NBF) return p < R ? p : -1;
        if (w & 1) return -1;
        const int c = 8 * NBF + (w >> 1);
        return c < R ? c : -1;
    }
};

// column of register (block b, element e) for lane quad index t, or -1
template <class PL>
__device__ __forceinline__ int reg_col(int b, int e, int t, int R) {
    if (b < PL::NBF) {
        const int c = 8 * b + 2 * t + e;
        return c < R ? c : -1;
    }
    if (e) return -1;
    const int c = 8 * PL::NBF + t;
    return c < R ? c : -1;
}

// Stage an R x R row-major operator M (global or shared, element type TS) into shared memory in position order.
template <class PL, typename TS>
__device__ __forceinline__ void stage_operator(const TS* __restrict__ M, int R, double* __restrict__ Ms, int tid, int nthreads) {
    for (int e = tid; e < PL::NPOS * PL::LDM; e += nthreads) {
        const int pr = e / PL::LDM, pc = e - pr * PL::LDM;
        double v = 0.0;
        if (pc < PL::NPOS) {
            const int r = PL::col_of(pr, R), c = PL::col_of(pc, R);
            if (r >= 0 && c >= 0) v = (double)M[r * R + c];
        }
        Ms[e] = v;
    }
}

// out = s * M  for the 8 rows of a warp (s, out in D layout; M staged by stage_operator)
template <class PL>
__device__ __forceinline__ void mma_rowmat(const double (&s)[PL::NB][2], const double* __restrict__ Ms, int g, int t,
                                           double (&out)[PL::NB][2]) {
#pragma unroll
    for (int nb = 0; nb < PL::NB; ++nb) out[nb][0] = out[nb][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < PL::KS; ++ks) {
        const int b = ks >> 1, e = ks & 1;
        const double* mrow = Ms + (8 * b + 2 * t + e) * PL::LDM + g;
#pragma unroll
        for (int nb = 0; nb < PL::NB; ++nb) dmma884(out[nb][0], out[nb][1], s[b][e], mrow[8 * nb]);
    }
}

// ---- Gram accumulation  S += V^T V  over the 8 rows of a warp -------------------------------------------------
// The rows are staged in a warp-private shared tile [8][LDT] (position order); the MMA contracts over the rows with
// k index t <-> row h + 2t (h = 0, 1), which keeps the fragment reads conflict free for LDT = 8 NB + 2.
// Only the block pairs bi <= bj are accumulated: pair index q = bi * NB - bi (bi - 1) / 2 + (bj - bi).
template <class PL>
struct GramAcc {
    static constexpr int NPAIR = PL::NB * (PL::NB + 1) / 2;
    static constexpr int LDT = 8 * PL::NB + 2;
    double acc[NPAIR][2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) acc[q][0] = acc[q][1] = 0.0;
    }
    // high words of the accumulators: a value that cannot exist before every fragment load feeding them has returned
    __device__ __forceinline__ int dep() const {
        int d = 0;
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) d = max(d, __double2hiint(acc[q][0]));
        return d;
    }
    // v: the warp's rows in D layout (padding positions MUST be zero); tile: 8 * LDT doubles owned by this warp.
    // extra_dep: dep() of another GramAcc that shares `tile` (0 if none).
    __device__ __forceinline__ void add(const double (&v)[PL::NB][2], double* __restrict__ tile, int g, int t,
                                        int extra_dep = 0) {
        // WAR guard: the fragment loads of the previous call(s) on this tile must have returned in every lane before
        // the tile is overwritten.  ptxas may sink loads below a WARPSYNC, so order the stores behind a warp vote on
        // the accumulators those loads fed (same device as stage_release in common.cuh).
        const int d = max(dep(), extra_dep);
        const unsigned never = __any_sync(0xffffffffu, d == 0x7ff7a5a5) ? 1u : 0u;
        double* tl = tile + never;
        __syncwarp();  // the memory-model barrier between the previous call's fragment reads and these stores (the vote
                       // above only pins the INSTRUCTION order; compute-sanitizer racecheck needs the real barrier)
#pragma unroll
        for (int b = 0; b < PL::NB; ++b) *(double2*)(tl + g * LDT + 8 * b + 2 * t) = make_double2(v[b][0], v[b][1]);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double f[PL::NB];
#pragma unroll
            for (int b = 0; b < PL::NB; ++b) f[b] = tile[(h + 2 * t) * LDT + 8 * b + g];
            int q = 0;
#pragma unroll
            for (int bi = 0; bi < PL::NB; ++bi)
#pragma unroll
                for (int bj = bi; bj < PL::NB; ++bj, ++q) dmma884(acc[q][0], acc[q][1], f[bi], f[bj]);
        }
    }
};

// Sum the per-warp Gram partials (fixed warp order) and write the R x R result (both triangles).
// red: shared scratch of n_warps * NPAIR * 64 doubles.  Must be called by all `nthreads` threads; `sync` is a
// barrier over exactly those threads.
template <class PL, typename T, class Sync>
__device__ __forceinline__ void gram_reduce_store(const GramAcc<PL>& ga, double* __restrict__ red, int warp, int lane,
                                                  int n_warps, int tid, int nthreads, int R, T* __restrict__ out,
                                                  Sync sync) {
    constexpr int NPAIR = GramAcc<PL>::NPAIR;
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
        red[((size_t)warp * NPAIR + q) * 64 + 2 * lane] = ga.acc[q][0];
        red[((size_t)warp * NPAIR + q) * 64 + 2 * lane + 1] = ga.acc[q][1];
    }
    sync();
    for (int e = tid; e < NPAIR * 64; e += nthreads) {
        const int q = e >> 6, w = e & 63, ln = w >> 1, el = w & 1;
        double s = 0.0;
        for (int wi = 0; wi < n_warps; ++wi) s += red[((size_t)wi * NPAIR + q) * 64 + w];
        int bi = 0, rem = q;
        while (rem >= PL::NB - bi) {
            rem -= PL::NB - bi;
            ++bi;
        }
        const int bj = bi + rem;
        const int gg = ln >> 2, tt = ln & 3;
        const int i = PL::col_of(8 * bi + gg, R), j = PL::col_of(8 * bj + 2 * tt + el, R);
        if (i >= 0 && j >= 0) {
            out[i * R + j] = (T)s;
            if (bi != bj) out[j * R + i] = (T)s;
        }
    }
}
