// X-stream contractions over the packed varlen data matrix X (N = sum_i J_i rows, K columns, row-major, ldx >= K):
//
//   b2_xstream_y : Y = X * C           (N x R)   — serves rhs of the B-update  X_i (C*a_i) = Y_i*a_i  (decomposition.py:242)
//                                                  and of the A-update  diag(B_i^T X_i C) = colsum(B_i o Y_i) (:145-158)
//   b2_xstream_z : Z = X^T * W         (K x R)   — W = B o a (row-scaled), rhs of the C-update  sum_i X_i^T (B_i*a_i) (:312-315)
//   b2_sumsq     : sum(X**2)                      — ||X||^2 for the fit term (:906)
//
// Both contractions are HBM-streaming tall-skinny products (R << K, J_i): X is read exactly once per launch through
// TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes) into an mbarrier-synchronised shared-memory ring filled by a
// dedicated producer warp; consumer warps contract from shared memory with FP64/FP32 FMAs or, for fp64, with
// DMMA.8x8x4 tensor-core instructions.  Grids are persistent (<= one CTA per SM, static tile striding).
#include "xstream_common.cuh"

namespace {

constexpr int kConsumerThreads = 256;
constexpr int kThreads = kConsumerThreads + 32;  // + producer warp

// ---------------------------------------------------------------------------------------------------------
// Y = X * C
// ---------------------------------------------------------------------------------------------------------
// Stage layout (1024-byte aligned): NBOX swizzled boxes [TM rows x 128 B] of X, then the C chunk [KC x LDC].
// FMA path: 128-row tiles, 2 boxes per stage.  DMMA path: 256-row tiles (4 m-blocks = 12..16 independent accumulator
// chains per warp: the DMMA pipe needs that much ILP with only 8 consumer warps per SM), 1 box per stage.
template <typename T, bool DMMA, int CW = 8>
struct YCfg {
    static constexpr int TM = DMMA ? 32 * CW : 128;   // rows per tile (DMMA: 32 rows = 4 m-blocks per consumer warp)
    static constexpr int RBOX = TM > 256 ? 128 : TM;  // rows per TMA box (a box dimension is limited to 256)
    static constexpr int NRB = TM / RBOX;             // row boxes per k-box
    static constexpr int MB = DMMA ? 4 : TM / 64;     // 8-row m-blocks per consumer warp (DMMA path)
    static constexpr int EPB = 128 / (int)sizeof(T);  // elements per 128-byte box row
    static constexpr int NBOX = DMMA ? 1 : 2;
    static constexpr int KC = NBOX * EPB;  // K-chunk per stage
    static constexpr int BOX_BYTES = TM * 128;
    static constexpr int X_BYTES = NBOX * BOX_BYTES;
};

// CT = columns per thread in the FMA path (thread tile 2 rows x CT cols, 4 column groups) ; LDC = 4*CT.
// DMMA path (fp64): warp w owns rows [16w,16w+16) = 2 m-blocks, NBLK n-blocks of 8 columns; LDC = 8*NBLK + 4.
// Shared-memory operand reads are 64-bit, i.e. served per half-warp (g = 0..3, t = 0..3) in 128-byte wavefronts:
//   A: MMA k-index t of k-step k4 is mapped to box column 8*(t>>1) + 2*k4 + (t&1), so that under SWIZZLE_128B the 16
//      lanes hit 16 distinct 8-byte bank pairs (the natural 4*k4 + t mapping is 2-way conflicted: rows g and g^1 land
//      in the same 32-byte window);
//   B: the staged C rows are stored in that same MMA order (pad_c_kernel) with LDC = 4 or 12 (mod 16), so lane (g,t)
//      reads word (4*k4 + t)*LDC + g: 16 distinct words mod 16.
// EX (DMMA path): columns 8*NBLK .. 8*NBLK+EX-1 (R = 8*NBLK + 1..4) are NOT padded to a third tensor-core block:
// they are contracted with DFMAs on the A fragments the lane already holds (lane (g,t) has X[row g][k(t)] and needs
// C[k(t)][8*NBLK + e], which sits in the 4 pad words of the staged C row) and summed over the 4 t-lanes at the end
// of the tile.  DMMA and DFMA share the fp64 units (b2_microbench_flops kind 3), so what this buys is the saved
// padding: R = 20 costs 20 columns of pipe time instead of 24.
// CW = consumer warps (DMMA path only: 8, or 12 = three per SM sub-partition for more issue slack around the DMMAs)
template <typename T, int CT, bool DMMA, int NBLK, int EX, int CW = 8>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
xstream_y_kernel(const __grid_constant__ CUtensorMap tmap_x, const T* __restrict__ Cp, T* __restrict__ Y, long long N,
                 int R, int Kp, int LDC, int num_tiles, int stages) {
    using Cfg = YCfg<T, DMMA, CW>;
    static_assert(DMMA || CW == 8, "the FMA path is laid out for 8 consumer warps");
    constexpr int TM = Cfg::TM, EPB = Cfg::EPB, NBOX = Cfg::NBOX, KC = Cfg::KC, MB = Cfg::MB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stages x (X boxes | C chunk)] | full[stages] | empty[stages]
    const uint32_t c_bytes = (uint32_t)(KC * LDC * sizeof(T));
    const uint32_t stage_bytes = (Cfg::X_BYTES + c_bytes + 1023u) & ~1023u;
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(base + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const uint32_t base_s = smem_u32(base);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int kchunks = Kp / KC;

    if (warp == CW) {
        // ===== producer warp: one elected lane issues all TMA traffic =====
        if (lane == 0) {
            prefetch_tmap(&tmap_x);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int row0 = tile * TM;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char* st = base + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], Cfg::X_BYTES + c_bytes);
#pragma unroll
                    for (int b = 0; b < NBOX; ++b)
#pragma unroll
                        for (int rb = 0; rb < Cfg::NRB; ++rb)  // a k-box of TM rows = NRB TMA boxes of RBOX rows
                            tma_load_2d(st + b * Cfg::BOX_BYTES + rb * Cfg::RBOX * 128, &tmap_x, kc * KC + b * EPB,
                                        row0 + rb * Cfg::RBOX, &full[s]);
                    bulk_load_1d(st + Cfg::X_BYTES, Cp + (size_t)kc * KC * LDC, c_bytes, &full[s]);
                    if (++s == stages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        return;
    }

    // ===== consumers =====
    int s = 0;
    uint32_t ph = 0;
    if constexpr (!DMMA) {
        const int tc = tid & 3, tr = tid >> 2;  // 4 column groups x 64 row pairs (rows tr, tr+64)
        constexpr int EPC = 16 / (int)sizeof(T);  // elements per 16-byte chunk
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            T acc0[CT], acc1[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) acc0[c] = acc1[c] = T(0);
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&full[s], ph);
                const uint32_t st = base_s + (uint32_t)s * stage_bytes;
                const uint32_t Cs = st + Cfg::X_BYTES + (uint32_t)(tc * CT * sizeof(T));
#pragma unroll
                for (int b = 0; b < NBOX; ++b) {
                    const uint32_t box = st + b * Cfg::BOX_BYTES;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        T x0[EPC], x1[EPC];
                        *(int4*)x0 = lds_b128(box + swz128(tr, q));
                        *(int4*)x1 = lds_b128(box + swz128(tr + 64, q));
#pragma unroll
                        for (int e = 0; e < EPC; ++e) {
                            const uint32_t crow = Cs + (uint32_t)((b * EPB + q * EPC + e) * LDC * sizeof(T));
                            T cv[CT];
                            if constexpr ((CT * sizeof(T)) % 16 == 0) {
#pragma unroll
                                for (int v = 0; v < (int)(CT * sizeof(T) / 16); ++v) *((int4*)cv + v) = lds_b128(crow + 16 * v);
                            } else if constexpr ((CT * sizeof(T)) % 8 == 0) {
#pragma unroll
                                for (int v = 0; v < (int)(CT * sizeof(T) / 8); ++v) *((int2*)cv + v) = lds_b64(crow + 8 * v);
                            } else {
#pragma unroll
                                for (int c = 0; c < CT; ++c) cv[c] = lds_elem<T>(crow + c * (uint32_t)sizeof(T));
                            }
#pragma unroll
                            for (int c = 0; c < CT; ++c) {
                                acc0[c] = fma(x0[e], cv[c], acc0[c]);
                                acc1[c] = fma(x1[e], cv[c], acc1[c]);
                            }
                        }
                    }
                }
                int dep = 0;
#pragma unroll
                for (int c = 0; c < CT; ++c) dep = max(dep, max(dep_bits_of(acc0[c]), dep_bits_of(acc1[c])));
                stage_release(&empty[s], lane, dep);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
            const long long r0 = (long long)tile * TM + tr, r1 = r0 + 64;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int col = tc * CT + c;
                if (col < R) {
                    if (r0 < N) Y[r0 * R + col] = acc0[c];
                    if (r1 < N) Y[r1 * R + col] = acc1[c];
                }
            }
        }
    } else {
        // fp64 tensor-core path. lane = 4g + t. A frag: X[16w + 8m + g][col(k4, t)]; B frag: C[col(k4, t)][8n + g].
        const int g = lane >> 2, t = lane & 3;
        const uint32_t a_chunk0 = 4u * (t >> 1), a_half = (t & 1) * 8u;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            double acc[MB][NBLK][2];
            double accx[MB][EX > 0 ? EX : 1];
#pragma unroll
            for (int m = 0; m < MB; ++m) {
#pragma unroll
                for (int n = 0; n < NBLK; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
#pragma unroll
                for (int e = 0; e < (EX > 0 ? EX : 1); ++e) accx[m][e] = 0.0;
            }
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&full[s], ph);
                const uint32_t st = base_s + (uint32_t)s * stage_bytes;
                const uint32_t Cs = st + Cfg::X_BYTES + (uint32_t)(g * sizeof(double));
#pragma unroll
                for (int b = 0; b < NBOX; ++b) {
                    const uint32_t box = st + b * Cfg::BOX_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < EPB / 4; ++k4) {
                        double a[MB], bf[NBLK];
#pragma unroll
                        for (int m = 0; m < MB; ++m) {
                            const uint32_t row = warp * (8 * MB) + m * 8 + g;
                            a[m] = lds_f64(box + swz128(row, a_chunk0 + k4) + a_half);
                        }
                        const uint32_t crow = Cs + (uint32_t)((b * EPB + k4 * 4 + t) * LDC * sizeof(double));
#pragma unroll
                        for (int n = 0; n < NBLK; ++n) bf[n] = lds_f64(crow + n * 64);
                        if constexpr (EX > 0) {
                            // extra columns: C[k(t)][8*NBLK + e] (same address for all g: broadcast)
                            double ce[EX];
                            const uint32_t cex = crow - (uint32_t)(g * sizeof(double)) + NBLK * 64;
#pragma unroll
                            for (int v = 0; v < EX / 2; ++v) *((int4*)ce + v) = lds_b128(cex + 16 * v);
#pragma unroll
                            for (int m = 0; m < MB; ++m)
#pragma unroll
                                for (int e = 0; e < EX; ++e) accx[m][e] = fma(a[m], ce[e], accx[m][e]);
                        }
#pragma unroll
                        for (int m = 0; m < MB; ++m)
#pragma unroll
                            for (int n = 0; n < NBLK; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[m], bf[n]);
                    }
                }
                int dep = 0;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
#pragma unroll
                    for (int n = 0; n < NBLK; ++n) dep = max(dep, dep_bits_of(acc[m][n][0]));
                    if constexpr (EX > 0) {
#pragma unroll
                        for (int e = 0; e < EX; ++e) dep = max(dep, dep_bits_of(accx[m][e]));
                    }
                }
                stage_release(&empty[s], lane, dep);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if constexpr (EX > 0) {  // sum the DFMA partials over the 4 lanes (t) that share a row
#pragma unroll
                for (int m = 0; m < MB; ++m)
#pragma unroll
                    for (int e = 0; e < EX; ++e) {
                        accx[m][e] += __shfl_xor_sync(0xffffffffu, accx[m][e], 1);
                        accx[m][e] += __shfl_xor_sync(0xffffffffu, accx[m][e], 2);
                    }
            }
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                const long long row = (long long)tile * TM + warp * (8 * MB) + m * 8 + g;
                if (row < N) {
#pragma unroll
                    for (int n = 0; n < NBLK; ++n) {
                        const int col = n * 8 + 2 * t;
                        if (col < R) Y[row * R + col] = (T)acc[m][n][0];
                        if (col + 1 < R) Y[row * R + col + 1] = (T)acc[m][n][1];
                    }
                    if constexpr (EX > 0) {  // lane t writes extra column t
#pragma unroll
                        for (int e = 0; e < EX; ++e)
                            if (e == t && 8 * NBLK + e < R) Y[row * R + 8 * NBLK + e] = (T)accx[m][e];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Z = X^T * W   (FMA path: thread owns 2 consecutive k's x all R columns; grid = row-groups x k-blocks)
// ---------------------------------------------------------------------------------------------------------
// Stage layout: KZB swizzled boxes [TMZ rows x 128 B] of X (KZB = KZ / EPB), then the W tile [TMZ x R] (contiguous rows).
template <typename T>
struct ZCfg {
    static constexpr int TMZ = 16;
    static constexpr int EPB = 128 / (int)sizeof(T);
    static constexpr int EPC = 16 / (int)sizeof(T);  // k's per thread (one 16-byte chunk)
    static constexpr int BOX_BYTES = TMZ * 128;
};

// VECW: the W tile rows are 16-byte aligned and exactly RP wide, so a row is read with 128-bit broadcast loads.
template <typename T, int RP, bool VECW>
__global__ void __launch_bounds__(kThreads, 1)
xstream_z_kernel(const __grid_constant__ CUtensorMap tmap_x, const T* __restrict__ W, int ldw, T* __restrict__ part, int K,
                 int R, int KZ, int num_tiles, int stages) {
    using Cfg = ZCfg<T>;
    constexpr int TMZ = Cfg::TMZ, EPB = Cfg::EPB, EPC = Cfg::EPC;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int kzb = KZ / EPB;
    const uint32_t x_bytes = (uint32_t)kzb * Cfg::BOX_BYTES;
    const uint32_t w_bytes = (uint32_t)(TMZ * ldw * sizeof(T));
    const uint32_t stage_bytes = (x_bytes + w_bytes + 1023u) & ~1023u;
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(base + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const uint32_t base_s = smem_u32(base);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_cons_warps = (blockDim.x >> 5) - 1;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], n_cons_warps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int k_base = blockIdx.y * KZ;

    if (warp == n_cons_warps) {
        if (lane == 0) {
            prefetch_tmap(&tmap_x);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int row0 = tile * TMZ;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = base + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], x_bytes + w_bytes);
                for (int b = 0; b < kzb; ++b)
                    tma_load_2d(st + b * Cfg::BOX_BYTES, &tmap_x, k_base + b * EPB, row0, &full[s]);
                bulk_load_1d(st + x_bytes, W + (size_t)row0 * ldw, w_bytes, &full[s]);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }

    // consumer thread `tid` owns k's [k_base + EPC*tid, +EPC) for all RP (>= R) columns
    T acc[EPC][RP];
#pragma unroll
    for (int e = 0; e < EPC; ++e)
#pragma unroll
        for (int c = 0; c < RP; ++c) acc[e][c] = T(0);
    const int box_id = tid / 8, chunk = tid & 7;  // 8 chunks of 16 B per 128-byte box row
    const bool active = tid * EPC < KZ;           // consumer count is rounded up to whole warps
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&full[s], ph);
        const uint32_t st = base_s + (uint32_t)s * stage_bytes;
        const uint32_t box = st + (uint32_t)box_id * Cfg::BOX_BYTES;
        const uint32_t Ws = st + x_bytes;
#pragma unroll 4
        for (int r = 0; r < (active ? TMZ : 0); ++r) {
            T x[EPC];
            *(int4*)x = lds_b128(box + swz128(r, chunk));
            const uint32_t wr = Ws + (uint32_t)(r * ldw * sizeof(T));
            if constexpr (VECW) {
                T wv[RP];
#pragma unroll
                for (int v = 0; v < (int)(RP * sizeof(T) / 16); ++v) *((int4*)wv + v) = lds_b128(wr + 16 * v);
#pragma unroll
                for (int c = 0; c < RP; ++c)
#pragma unroll
                    for (int e = 0; e < EPC; ++e) acc[e][c] = fma(x[e], wv[c], acc[e][c]);
            } else {
#pragma unroll
                for (int c = 0; c < RP; ++c) {
                    const T w = (c < R) ? lds_elem<T>(wr + c * (uint32_t)sizeof(T)) : T(0);
#pragma unroll
                    for (int e = 0; e < EPC; ++e) acc[e][c] = fma(x[e], w, acc[e][c]);
                }
            }
        }
        int dep = 0;
#pragma unroll
        for (int e = 0; e < EPC; ++e)
#pragma unroll
            for (int c = 0; c < RP; ++c) dep = max(dep, dep_bits_of(acc[e][c]));
        stage_release(&empty[s], lane, dep);
        if (++s == stages) {
            s = 0;
            ph ^= 1;
        }
    }
    // partials: part[blockIdx.x][k][c]
    T* out = part + (size_t)blockIdx.x * K * R;
#pragma unroll
    for (int e = 0; e < EPC; ++e) {
        const int k = k_base + tid * EPC + e;
        if (active && k < K) {  // inactive (warp-rounding) threads would alias the next k-block's rows
#pragma unroll
            for (int c = 0; c < RP; ++c)
                if (c < R) out[(size_t)k * R + c] = acc[e][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Z = X^T * W, fp64 tensor-core path.  CTA = (row group, k-block of 128 columns); stage = 8 swizzled boxes
// [32 rows x 16 doubles] of X plus the W tile [32 x ldw].  Warp w owns the 16 k's of box w (2 m-blocks).
// MMA roles: M = k index (8), N = factor column (8), K = data rows (4, taken as r0+h+2t so that the swizzled
// A-fragment reads are bank-conflict free; the B-fragment reads word (r0+h+2t)*ldw + g, conflict free per half-warp
// for ldw = 8*NBLK + 2, i.e. 2*ldw = 4 (mod 16)).
// ---------------------------------------------------------------------------------------------------------
constexpr int kZdTM = 32;
constexpr int kZdKZ = 128;
constexpr int kZdBoxBytes = kZdTM * 128;
constexpr int kZdXBytes = (kZdKZ / 16) * kZdBoxBytes;

// EX: as in xstream_y_kernel — the remainder columns 8*NBLK .. 8*NBLK+EX-1 of W are contracted with DFMAs on the A
// fragments (lane (g,t) holds X[row(t)][k(g)] and needs W[row(t)][8*NBLK + e]: a broadcast over g) and summed over
// the 4 t-lanes once, after the last tile.
// CW = consumer warps: 8 (warp w owns the 16 k's of box w over all 32 rows of a stage, two accumulator sets), or 16 for
// the fp64-pipe-bound ranks: two warps per box, each on 16 of the 32 rows with ONE accumulator set — the same number of
// DMMA chains per SM on twice as many warps (four per SM sub-partition), and every warp pair writes its own partial.
template <int NBLK, int EX, int CW = 8>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
xstream_z_dmma_kernel(const __grid_constant__ CUtensorMap tmap_x, const double* __restrict__ W, int ldw,
                      double* __restrict__ part, int K, int R, int num_tiles, int stages) {
    static_assert(CW == 8 || (CW == 16 && EX == 0), "16 consumer warps: padded column blocks only");
    constexpr int SETS = CW == 8 ? 2 : 1;   // accumulator sets per warp
    constexpr int RGS = CW == 8 ? 4 : 2;    // 8-row groups of a stage per warp
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t w_bytes = (uint32_t)(kZdTM * ldw * sizeof(double));
    const uint32_t stage_bytes = (kZdXBytes + w_bytes + 1023u) & ~1023u;
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(base + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const uint32_t base_s = smem_u32(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int k_base = blockIdx.y * kZdKZ;
    if (warp == CW) {
        if (lane == 0) {
            prefetch_tmap(&tmap_x);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int row0 = tile * kZdTM;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = base + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], kZdXBytes + w_bytes);
#pragma unroll
                for (int b = 0; b < kZdKZ / 16; ++b)
                    tma_load_2d(st + b * kZdBoxBytes, &tmap_x, k_base + b * 16, row0, &full[s]);
                bulk_load_1d(st + kZdXBytes, W + (size_t)row0 * ldw, w_bytes, &full[s]);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }
    const int g = lane >> 2, t = lane & 3;
    const int kbox = warp & 7, half = warp >> 3;  // CW == 8: half == 0
    // CW == 8: two independent accumulator sets (rows r0+h+2t, h = 0 / 1, summed at the end): 4*NBLK dependent DMMA
    // chains per warp instead of 2*NBLK — with 8 consumer warps per SM the DMMA pipe needs that much ILP.
    double acc[SETS][2][NBLK][2];
    double accx[2][EX > 0 ? EX : 1];  // [m][extra column], summed over this lane's rows
#pragma unroll
    for (int u = 0; u < SETS; ++u)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NBLK; ++n) acc[u][m][n][0] = acc[u][m][n][1] = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int e = 0; e < (EX > 0 ? EX : 1); ++e) accx[m][e] = 0.0;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&full[s], ph);
        const uint32_t st = base_s + (uint32_t)s * stage_bytes;
        const uint32_t box = st + (uint32_t)kbox * kZdBoxBytes;
        const uint32_t Ws = st + kZdXBytes + (uint32_t)(g * sizeof(double));
#pragma unroll
        for (int rgi = 0; rgi < RGS; ++rgi) {
            const int rg = half * RGS + rgi;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t row = rg * 8 + h + 2 * t;
                double a[2], bf[NBLK];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t kk = 8 * m + g;
                    a[m] = lds_f64(box + swz128(row, kk >> 1) + (kk & 1) * 8);
                }
                const uint32_t wrow = Ws + (uint32_t)(row * ldw * sizeof(double));
#pragma unroll
                for (int n = 0; n < NBLK; ++n) bf[n] = lds_f64(wrow + n * 64);
                if constexpr (EX > 0) {
                    double we[EX];
                    const uint32_t wex = wrow - (uint32_t)(g * sizeof(double)) + NBLK * 64;
#pragma unroll
                    for (int v = 0; v < EX / 2; ++v) *((int4*)we + v) = lds_b128(wex + 16 * v);
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int e = 0; e < EX; ++e) accx[m][e] = fma(a[m], we[e], accx[m][e]);
                }
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < NBLK; ++n)
                        dmma884(acc[h % SETS][m][n][0], acc[h % SETS][m][n][1], a[m], bf[n]);
            }
        }
        int dep = 0;
#pragma unroll
        for (int u = 0; u < SETS; ++u)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NBLK; ++n) dep = max(dep, dep_bits_of(acc[u][m][n][0]));
        if constexpr (EX > 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int e = 0; e < EX; ++e) dep = max(dep, dep_bits_of(accx[m][e]));
        }
        stage_release(&empty[s], lane, dep);
        if (++s == stages) {
            s = 0;
            ph ^= 1;
        }
    }
    // one partial per (row group, row half): part[(blockIdx.x * (CW / 8) + half)][k][c]
    double* out = part + ((size_t)blockIdx.x * (CW / 8) + half) * K * R;
    if constexpr (EX > 0) {  // sum the DFMA partials over the 4 lanes (t) that share a k
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int e = 0; e < EX; ++e) {
                accx[m][e] += __shfl_xor_sync(0xffffffffu, accx[m][e], 1);
                accx[m][e] += __shfl_xor_sync(0xffffffffu, accx[m][e], 2);
            }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int k = k_base + kbox * 16 + m * 8 + g;
        if (k < K) {
#pragma unroll
            for (int n = 0; n < NBLK; ++n) {
                const int col = n * 8 + 2 * t;
                const double v0 = SETS == 2 ? acc[0][m][n][0] + acc[SETS - 1][m][n][0] : acc[0][m][n][0];
                const double v1 = SETS == 2 ? acc[0][m][n][1] + acc[SETS - 1][m][n][1] : acc[0][m][n][1];
                if (col < R) out[(size_t)k * R + col] = v0;
                if (col + 1 < R) out[(size_t)k * R + col + 1] = v1;
            }
            if constexpr (EX > 0) {  // lane t writes extra column t
#pragma unroll
                for (int e = 0; e < EX; ++e)
                    if (e == t && 8 * NBLK + e < R) out[(size_t)k * R + 8 * NBLK + e] = accx[m][e];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// sum of squares of X (only the K valid columns of each ldx-strided row)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sumsq_kernel(const T* __restrict__ X, long long N, int K, int ldx, double* __restrict__ part) {
    __shared__ double scratch[32];
    double acc = 0.0;
    const long long total = N * (long long)ldx;
    constexpr int V = 16 / (int)sizeof(T);
    // ldx is a multiple of V and padding columns are zero, so the padded buffer can be summed as a flat vector
    const long long nvec = total / V;
    const int4* Xv = (const int4*)X;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        int4 raw = __ldg(Xv + i);
        const T* v = (const T*)&raw;
#pragma unroll
        for (int e = 0; e < V; ++e) acc += (double)v[e] * (double)v[e];
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void reduce_double_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc;
}

// R = 8*NBLK + (1..4) with NBLK >= 1: the remainder goes to the DFMA path (EX = 2 or 4 columns incl. padding);
// otherwise R is padded up to whole tensor-core blocks.
void y_dmma_split(int R, int* NBLK, int* EX) {
    const int nb = R / 8, rem = R % 8;
    if (nb >= 1 && rem >= 1 && rem <= 4 && b2_option_value(B2_OPT_XSTREAM_HYBRID) == 1) {
        *NBLK = nb;
        *EX = rem <= 2 ? 2 : 4;
    } else {
        *NBLK = (R + 7) / 8;
        *EX = 0;
    }
}

template <typename T, int CT>
int launch_y_fma(const CUtensorMap& map, const T* Cp, T* Y, long long N, int R, int Kp, int num_tiles, int grid,
                 int stages, size_t smem, cudaStream_t st) {
    auto kern = xstream_y_kernel<T, CT, false, 1, 0>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(map, Cp, Y, N, R, Kp, 4 * CT, num_tiles, stages);
    B2_LAUNCH_CHECK();
    return B2_OK;
}
template <int NBLK, int EX, int CW = 8>
int launch_y_dmma(const CUtensorMap& map, const double* Cp, double* Y, long long N, int R, int Kp, int LDC, int num_tiles,
                  int grid, int stages, size_t smem, cudaStream_t st) {
    auto kern = xstream_y_kernel<double, 1, true, NBLK, EX, CW>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, (CW + 1) * 32, smem, st>>>(map, Cp, Y, N, R, Kp, LDC, num_tiles, stages);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T, bool DMMA, int CW = 8>
int xstream_y_impl(const void* X, long long N, int K, int ldx, const void* C, int R, void* Y, void* ws, size_t ws_bytes,
                   int max_ctas, cudaStream_t st) {
    using Cfg = YCfg<T, DMMA, CW>;
    const int dtype = sizeof(T) == 8 ? B2_F64 : B2_F32;
    constexpr bool dmma = DMMA;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (N == 0) return B2_OK;
    const int Kp = ((K + Cfg::KC - 1) / Cfg::KC) * Cfg::KC;
    int CT = (R + 3) / 4, LDC = 4 * CT, NBLK = (R + 7) / 8, EX = 0;
    if (dmma) {
        y_dmma_split(R, &NBLK, &EX);
        LDC = 8 * NBLK + 4;  // == 4 or 12 (mod 16): conflict-free B-fragment reads; the 4 pad words hold the EX columns
    }
    const size_t cp_bytes = (size_t)Kp * LDC * sizeof(T);
    B2_REQUIRE(ws_bytes >= cp_bytes, "xstream_y workspace too small: need %zu bytes, got %zu", cp_bytes, ws_bytes);
    B2_REQUIRE(((uintptr_t)ws) % 16 == 0, "workspace must be 16-byte aligned");
    pad_c_kernel<T><<<(Kp * LDC + 255) / 256, 256, 0, st>>>((const T*)C, (T*)ws, K, R, Kp, LDC, dmma ? 1 : 0);
    B2_LAUNCH_CHECK();

    alignas(64) CUtensorMap map;
    int rc = encode_x_map(&map, X, N, K, ldx, dtype, Cfg::RBOX);
    if (rc != B2_OK) return rc;
    const int num_tiles = (int)((N + Cfg::TM - 1) / Cfg::TM);
    const uint32_t c_bytes = (uint32_t)(Cfg::KC * LDC * sizeof(T));
    const uint32_t stage_bytes = (Cfg::X_BYTES + c_bytes + 1023u) & ~1023u;
    int stages = (int)((220 * 1024 - 1024 - 256) / stage_bytes);
    if (stages > 6) stages = 6;
    B2_REQUIRE(stages >= 2, "xstream_y: stage does not fit shared memory");
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 2 * stages * sizeof(uint64_t);
    int grid = b2_num_sms();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
    if (grid > num_tiles) grid = num_tiles;

    if constexpr (dmma) {
        const double* Cp = (const double*)ws;
        double* Yd = (double*)Y;
#define B2_Y_CASE(NB, E) \
    if (NBLK == NB && EX == E) return launch_y_dmma<NB, E, CW>(map, Cp, Yd, N, R, Kp, LDC, num_tiles, grid, stages, smem, st);
        B2_Y_CASE(1, 0) B2_Y_CASE(2, 0) B2_Y_CASE(3, 0) B2_Y_CASE(4, 0)
        if constexpr (CW == 8) {
            B2_Y_CASE(1, 2) B2_Y_CASE(1, 4) B2_Y_CASE(2, 2) B2_Y_CASE(2, 4) B2_Y_CASE(3, 2) B2_Y_CASE(3, 4)
        }
#undef B2_Y_CASE
        B2_REQUIRE(false, "xstream_y: no kernel for NBLK=%d EX=%d", NBLK, EX);
    }
    const T* Cp = (const T*)ws;
    T* Yt = (T*)Y;
    switch (CT) {
        case 1: return launch_y_fma<T, 1>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 2: return launch_y_fma<T, 2>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 3: return launch_y_fma<T, 3>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 4: return launch_y_fma<T, 4>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 5: return launch_y_fma<T, 5>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 6: return launch_y_fma<T, 6>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        case 7: return launch_y_fma<T, 7>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
        default: return launch_y_fma<T, 8>(map, Cp, Yt, N, R, Kp, num_tiles, grid, stages, smem, st);
    }
}

template <typename T, int RP>
int launch_z(const CUtensorMap& map, const T* W, int ldw, T* part, int K, int R, int KZ, int num_tiles, dim3 grid,
             int threads, int stages, size_t smem, cudaStream_t st) {
    const bool vecw = RP == R && ((size_t)ldw * sizeof(T)) % 16 == 0 && (RP * sizeof(T)) % 16 == 0;
    if (vecw) {
        auto kern = xstream_z_kernel<T, RP, true>;
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           (int)cudaSharedmemCarveoutMaxShared));
        kern<<<grid, threads, smem, st>>>(map, W, ldw, part, K, R, KZ, num_tiles, stages);
    } else {
        auto kern = xstream_z_kernel<T, RP, false>;
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           (int)cudaSharedmemCarveoutMaxShared));
        kern<<<grid, threads, smem, st>>>(map, W, ldw, part, K, R, KZ, num_tiles, stages);
    }
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <int NBLK, int EX, int CW = 8>
int launch_z_dmma(const CUtensorMap& map, const double* W, int ldw, double* part, int K, int R, int num_tiles, dim3 grid,
                  int stages, size_t smem, cudaStream_t st) {
    auto kern = xstream_z_dmma_kernel<NBLK, EX, CW>;
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, (CW + 1) * 32, smem, st>>>(map, W, ldw, part, K, R, num_tiles, stages);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int z_ldw_for(int R, int dtype, int variant) {
    if (variant == B2_VARIANT_DMMA && dtype == B2_F64) {
        const int nblk = (R + 7) / 8;
        return 8 * nblk + 2;  // 2*ldw == 4 (mod 16): B-fragment rows r0+h+2t, t = 0..3, start 4 words apart
    }
    return R;
}

int xstream_z_dmma_impl(const void* X, long long N, int K, int ldx, const void* W, int ldw, int R, void* Z, void* ws,
                        size_t ws_bytes, int max_ctas, cudaStream_t st) {
    int NBLK, EX;
    y_dmma_split(R, &NBLK, &EX);  // ldw keeps the padded-block formula, so the EX columns are simply W[:, 8*NBLK ...]
    B2_REQUIRE(ldw == z_ldw_for(R, B2_F64, B2_VARIANT_DMMA), "xstream_z (DMMA): W row stride must be %d, got %d",
               z_ldw_for(R, B2_F64, B2_VARIANT_DMMA), ldw);
    const int kblocks = (K + kZdKZ - 1) / kZdKZ;
    const int num_tiles = (int)((N + kZdTM - 1) / kZdTM);
    int groups = b2_num_sms() / kblocks;
    if (max_ctas > 0 && max_ctas / kblocks < groups) groups = max_ctas / kblocks;
    if (groups < 1) groups = 1;
    if (groups > num_tiles) groups = num_tiles;
    // 16 consumer warps (two per 16-k box, one partial per row half) where the kernel is bound by the fp64 tensor pipe
    // (three column blocks and more), like the Y kernel's 12; 8 warps for the HBM-bound ranks and the DFMA-remainder variant
    const int opt = b2_option_value(B2_OPT_XSTREAM_HYBRID);
    const bool wide = EX == 0 && (opt == 2 || (opt == 0 && R > 16));
    const int halves = wide ? 2 : 1;
    const size_t part_bytes = (size_t)groups * halves * K * R * sizeof(double);
    B2_REQUIRE(ws_bytes >= part_bytes, "xstream_z workspace too small: need %zu bytes, got %zu", part_bytes, ws_bytes);
    alignas(64) CUtensorMap map;
    int rc = encode_x_map(&map, X, N, K, ldx, B2_F64, kZdTM);
    if (rc != B2_OK) return rc;
    const uint32_t w_bytes = (uint32_t)(kZdTM * ldw * sizeof(double));
    const uint32_t stage_bytes = (kZdXBytes + w_bytes + 1023u) & ~1023u;
    int stages = (int)((224 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 2 * stages * sizeof(uint64_t);
    dim3 grid(groups, kblocks);
    const double* Wd = (const double*)W;
    double* part = (double*)ws;
    rc = B2_ERR_INVALID;
#define B2_Z_CASE(NB, E) \
    if (NBLK == NB && EX == E) rc = launch_z_dmma<NB, E>(map, Wd, ldw, part, K, R, num_tiles, grid, stages, smem, st);
#define B2_Z_WIDE(NB) \
    if (NBLK == NB) rc = launch_z_dmma<NB, 0, 16>(map, Wd, ldw, part, K, R, num_tiles, grid, stages, smem, st);
    if (wide) {
        B2_Z_WIDE(1) B2_Z_WIDE(2) B2_Z_WIDE(3) B2_Z_WIDE(4)
    } else {
        B2_Z_CASE(1, 0) B2_Z_CASE(2, 0) B2_Z_CASE(3, 0) B2_Z_CASE(4, 0)
        B2_Z_CASE(1, 2) B2_Z_CASE(1, 4) B2_Z_CASE(2, 2) B2_Z_CASE(2, 4) B2_Z_CASE(3, 2) B2_Z_CASE(3, 4)
    }
#undef B2_Z_CASE
#undef B2_Z_WIDE
    if (rc != B2_OK) return rc;
    const int n = K * R;
    reduce_partials_kernel<double><<<(n + 255) / 256, 256, 0, st>>>(part, (double*)Z, n, groups * halves);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

template <typename T>
int xstream_z_impl(const void* X, long long N, int K, int ldx, const void* W, int ldw, int R, void* Z, void* ws,
                   size_t ws_bytes, int variant, int max_ctas, cudaStream_t st) {
    using Cfg = ZCfg<T>;
    const int dtype = sizeof(T) == 8 ? B2_F64 : B2_F32;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (N == 0) {
        B2_CHECK_CUDA(cudaMemsetAsync(Z, 0, (size_t)K * R * sizeof(T), st));
        return B2_OK;
    }
    if (variant == B2_VARIANT_DMMA) {
        B2_REQUIRE(dtype == B2_F64, "the DMMA variant exists for fp64 only");
        return xstream_z_dmma_impl(X, N, K, ldx, W, ldw, R, Z, ws, ws_bytes, max_ctas, st);
    }
    B2_REQUIRE(ldw >= R && ((size_t)Cfg::TMZ * ldw * sizeof(T)) % 16 == 0, "xstream_z: bad W row stride %d", ldw);
    // k-block per CTA: up to 256 consumer threads x EPC k's; rounded to whole 128-byte boxes
    const int kmax = kConsumerThreads * Cfg::EPC;
    const int Kbox = ((K + Cfg::EPB - 1) / Cfg::EPB) * Cfg::EPB;
    const int kblocks = (Kbox + kmax - 1) / kmax;
    int KZ = (Kbox + kblocks - 1) / kblocks;
    KZ = ((KZ + Cfg::EPB - 1) / Cfg::EPB) * Cfg::EPB;
    // consumer threads: one per 16-byte chunk, rounded up to whole warps
    const int cons = ((KZ / Cfg::EPC + 31) / 32) * 32;
    const int threads = cons + 32;
    const int num_tiles = (int)((N + Cfg::TMZ - 1) / Cfg::TMZ);
    // narrow data (few consumer warps per k-block, e.g. fp32 with K = 256: 64 threads): several CTAs share an SM, each
    // with its share of the shared memory, instead of one CTA with two busy warps
    int ctas_per_sm = kConsumerThreads / cons;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1 || kblocks > 1) ctas_per_sm = 1;
    int groups = b2_num_sms() * ctas_per_sm / kblocks;
    if (max_ctas > 0 && max_ctas / kblocks < groups) groups = max_ctas / kblocks;
    if (groups < 1) groups = 1;
    if (groups > num_tiles) groups = num_tiles;
    size_t part_bytes = (size_t)groups * K * R * sizeof(T);
    while (part_bytes > ws_bytes && ctas_per_sm > 1) {  // a caller-sized workspace for one CTA per SM still works
        --ctas_per_sm;
        groups = b2_num_sms() * ctas_per_sm / kblocks;
        if (groups > num_tiles) groups = num_tiles;
        part_bytes = (size_t)groups * K * R * sizeof(T);
    }
    B2_REQUIRE(ws_bytes >= part_bytes, "xstream_z workspace too small: need %zu bytes, got %zu", part_bytes, ws_bytes);

    alignas(64) CUtensorMap map;
    int rc = encode_x_map(&map, X, N, K, ldx, dtype, Cfg::TMZ);
    if (rc != B2_OK) return rc;
    const uint32_t x_bytes = (uint32_t)(KZ / Cfg::EPB) * Cfg::BOX_BYTES;
    const uint32_t w_bytes = (uint32_t)(Cfg::TMZ * ldw * sizeof(T));
    const uint32_t stage_bytes = (x_bytes + w_bytes + 1023u) & ~1023u;
    int stages = (int)(((224 * 1024) / ctas_per_sm - 2048) / stage_bytes);
    if (stages > 8) stages = 8;
    B2_REQUIRE(stages >= 2, "xstream_z: stage does not fit shared memory");
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 2 * stages * sizeof(uint64_t);
    dim3 grid(groups, kblocks);
    const T* Wt = (const T*)W;
    T* part = (T*)ws;
    const int RP = ((R + 3) / 4) * 4;
    switch (RP) {
        case 4: rc = launch_z<T, 4>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 8: rc = launch_z<T, 8>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 12: rc = launch_z<T, 12>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 16: rc = launch_z<T, 16>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 20: rc = launch_z<T, 20>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 24: rc = launch_z<T, 24>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        case 28: rc = launch_z<T, 28>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
        default: rc = launch_z<T, 32>(map, Wt, ldw, part, K, R, KZ, num_tiles, grid, threads, stages, smem, st); break;
    }
    if (rc != B2_OK) return rc;
    const int n = K * R;
    reduce_partials_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(part, (T*)Z, n, groups);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_xstream_y(const void* X, long long n_rows, int K, int ldx, const void* C, int R, void* Y, int dtype, void* ws,
                 size_t ws_bytes, int variant, int max_ctas, void* stream) {
    if (variant == B2_VARIANT_DMMA) {
        B2_REQUIRE(dtype == B2_F64, "the DMMA variant exists for fp64 only");
        // 12 consumer warps (384-row tiles, three warps per SM sub-partition) where the kernel is bound by the fp64
        // tensor pipe, i.e. from three column blocks on (R >= 17): the third warp fills issue gaps around the DMMAs
        // (R = 20: 7.92 -> 7.4-7.6 ms per 39 GB under sustained load, R = 32: 4.03 -> 3.84 ms per 16 GB; bit-identical).
        // HBM-bound ranks keep 8 warps and 256-row tiles (finer tile granularity over 148 CTAs).
        const int opt = b2_option_value(B2_OPT_XSTREAM_HYBRID);
        if (opt == 2 || (opt == 0 && R > 16))
            return xstream_y_impl<double, true, 12>(X, n_rows, K, ldx, C, R, Y, ws, ws_bytes, max_ctas,
                                                    (cudaStream_t)stream);
        return xstream_y_impl<double, true>(X, n_rows, K, ldx, C, R, Y, ws, ws_bytes, max_ctas, (cudaStream_t)stream);
    }
    B2_DISPATCH_DTYPE(dtype, return xstream_y_impl<T, false>(X, n_rows, K, ldx, C, R, Y, ws, ws_bytes, max_ctas,
                                                             (cudaStream_t)stream));
}

int b2_xstream_z(const void* X, long long n_rows, int K, int ldx, const void* W, int ldw, int R, void* Z, int dtype,
                 void* ws, size_t ws_bytes, int variant, int max_ctas, void* stream) {
    B2_DISPATCH_DTYPE(dtype, return xstream_z_impl<T>(X, n_rows, K, ldx, W, ldw, R, Z, ws, ws_bytes, variant, max_ctas,
                                                      (cudaStream_t)stream));
}

int b2_xstream_z_ldw(int R, int dtype, int variant) { return z_ldw_for(R, dtype, variant); }

size_t b2_xstream_workspace_bytes(int K, int R, int dtype) {
    const size_t es = dtype == B2_F64 ? 8 : 4;
    const size_t kp = (size_t)((K + 63) / 64) * 64;
    const size_t y = kp * 40 * es;
    // Z partials: one K x R block per CTA; narrow data runs up to 4 CTAs per SM (K <= 1024 elements per SM either way)
    const size_t z = (size_t)b2_num_sms() * (K < 1024 ? 1024 : K) * R * es;
    return (y > z ? y : z) + 256;
}

int b2_sumsq(const void* X, long long n_rows, int K, int ldx, int dtype, double* out, void* ws, size_t ws_bytes,
             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = b2_num_sms() * 4;
    B2_REQUIRE(ws_bytes >= blocks * sizeof(double), "b2_sumsq workspace too small");
    (void)K;
    if (dtype == B2_F64)
        sumsq_kernel<double><<<blocks, 256, 0, st>>>((const double*)X, n_rows, K, ldx, (double*)ws);
    else if (dtype == B2_F32)
        sumsq_kernel<float><<<blocks, 256, 0, st>>>((const float*)X, n_rows, K, ldx, (double*)ws);
    else
        B2_REQUIRE(false, "unknown dtype %d", dtype);
    B2_LAUNCH_CHECK();
    reduce_double_kernel<<<1, 256, 0, st>>>((const double*)ws, blocks, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // extern "C"
