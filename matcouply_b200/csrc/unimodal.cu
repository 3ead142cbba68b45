// Unimodal regression prox (penalties.py:1014-1015 -> _unimodal_regression.py:24-141).
//
// Two threads per (slice, column): lane l < 16 of a warp runs the prefix-isotonic regression (PAVA) of the column, lane
// l + 16 the one of the reversed column (_unimodal_regression.py:96-97), so a warp works on 16 columns.  The
// arithmetic is IEEE round-to-nearest fp64 with explicit non-contracted intrinsics in the reference's operation order,
// so block levels, prefix errors and therefore the peak index t* are bit-identical to the NumPy/numba reference on
// identical input.
//
// Why not one warp per column (BASELINE.json's sketch): PAVA is a sequential stack algorithm and the bit-exact peak
// forbids re-associating its sums, so the parallelism is across columns and across the two directions.  What makes a
// thread-per-sequence PAVA slow on a GPU is divergence: inside a warp every element step costs the LONGEST merge
// cascade of its 32 lanes.  Here the loop is a state machine that does ONE unit of work per trip — either one merge
// (cur absorbs the block below) or one finalisation (store the results of element i, fetch element i + 1) — and both
// need exactly one fp64 division, which is shared.  Lanes drift apart in the element index instead of waiting for
// each other: the warp runs max_lanes(n + merges) trips, not sum_steps max_lanes(merges per step).
//
// Per element and direction the thread records what the reference keeps in (sumwy, sumwy2, index_range, error)
// (_unimodal_regression.py:31-37): the merged sums at each block end double as the block stack (a pop re-reads them),
// the prefix errors feed the peak search (:84-92), and (start, sum) reconstruct the fit of ANY prefix (:72-81) — no
// PAVA re-run once t* is known.  The two top blocks stay in registers; a pop refills the second one early.
#include "common.cuh"

namespace {

struct Block {
    int start;
    double sy, sy2, level, err_after;
};

// Per-thread scratch in global memory: one 32-byte record (= one DRAM/L2 sector) per element, contiguous per thread.
// rec[i] holds what the reference keeps at index i after processing element i: error[i+1], sumwy[i], sumwy2[i],
// index_range[i] (_unimodal_regression.py:31-37).
struct __align__(16) Rec {
    double err_after, sy, sy2;
    int start, pad;
};

constexpr int kThreads = 128;
// YRING: elements of the column in flight ahead of the PAVA front (cp.async into shared memory);
// STACK: most recent stack blocks (below the register-resident top) cached in shared memory — a pop below the cached
// window re-reads the recorded block end from global memory, and since a warp trip costs the slowest of its 32 lanes,
// a few per cent of such pops per lane put a global-memory round trip into a quarter of all trips.

// shared memory, all arrays [depth][thread] so that a warp access is conflict free
template <int YRING, int STACK>
struct SharedState {
    double y[YRING][kThreads];
    double2 sums[STACK][kThreads];   // sy, sy2
    double2 lvl_err[STACK][kThreads];  // level, err_after
    int start[STACK][kThreads];
};

template <typename T>
__device__ __forceinline__ void cp_async_elem(void* smem_dst, const T* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void store_rec(Rec* r, double err_after, double sy, double sy2, int start) {
    double2* q = (double2*)r;
    q[0] = make_double2(err_after, sy);
    q[1] = make_double2(sy2, __hiloint2double(0, start));
}
__device__ __forceinline__ int rec_start(const Rec* r) { return r->start; }

// Prefix-isotonic PAVA over seq(j) = col[(rev ? n-1-j : j) * ld], j in [0, n); fills rec[0..n).
template <typename T, int kYRing, int kStack>
__device__ __forceinline__ void pava_prefix(const T* __restrict__ col, long long ld, int n, bool rev, bool nn,
                                            Rec* __restrict__ rec, SharedState<kYRing, kStack>& sh, int tid) {
    const long long step = rev ? -ld : ld;
    const T* p = col + (rev ? (long long)(n - 1) * ld : 0LL);
    // element j lands in ring slot j % kYRing, stored as T in the first sizeof(T) bytes of the slot
#pragma unroll
    for (int u = 1; u <= kYRing; ++u) {
        if (u < n) cp_async_elem<T>(&sh.y[u % kYRing][tid], p + (long long)u * step);
        cp_async_commit();
    }
    Block cur, top;
    bool has_top = false;
    int depth = 0, cached = 0;  // blocks below `top`; how many of the most recent ones sit in the shared ring
    top.start = 0;
    top.sy = top.sy2 = top.level = top.err_after = 0.0;
    int i = 0;
    {
        const double y0 = (double)p[0];
        cur.start = 0;
        cur.sy = y0;
        cur.sy2 = __dmul_rn(y0, y0);
        cur.level = y0;
    }
    double cum = cur.sy2;  // running sum of y^2 (cumsumwy2, :42)
    while (i < n) {
        // `<=` pooling (:53); index_range[i] != 0 <=> there is a block below
        const bool merge = has_top && cur.level <= top.level;
        double num;
        if (merge) {
            cur.sy = __dadd_rn(cur.sy, top.sy);     // _merge_intervals_inplace (:15-17)
            cur.sy2 = __dadd_rn(cur.sy2, top.sy2);
            cur.start = top.start;
            has_top = cur.start > 0;  // blocks tile [0, i]: something lies below iff the merged block does not start at 0
            if (has_top) {
                --depth;
                if (cached > 0) {
                    --cached;
                    const int d = depth % kStack;
                    const double2 a = sh.sums[d][tid], b = sh.lvl_err[d][tid];
                    top.start = sh.start[d][tid];
                    top.sy = a.x;
                    top.sy2 = a.y;
                    top.level = b.x;
                    top.err_after = b.y;
                } else {  // cascade deeper than the cached window: re-read the recorded block end (rare)
                    const Rec* r = rec + (cur.start - 1);
                    top.start = r->start;
                    top.sy = r->sy;
                    top.sy2 = r->sy2;
                    top.err_after = r->err_after;
                    top.level = __ddiv_rn(top.sy, (double)(cur.start - top.start));
                }
            }
            num = cur.sy;
        } else {
            num = __dmul_rn(cur.sy, cur.sy);
        }
        const double q = __ddiv_rn(num, (double)(i - cur.start + 1));
        if (merge) {
            cur.level = q;  // (:21)
        } else {
            const double levelerror = __dsub_rn(cur.sy2, q);  // (:57)
            const double before = has_top ? top.err_after : 0.0;
            cur.err_after = (nn && cur.level < 0.0) ? cum : __dadd_rn(levelerror, before);  // (:58-62)
            store_rec(rec + i, cur.err_after, cur.sy, cur.sy2, cur.start);
            if (has_top) {  // push the old top
                const int d = depth % kStack;
                sh.sums[d][tid] = make_double2(top.sy, top.sy2);
                sh.lvl_err[d][tid] = make_double2(top.level, top.err_after);
                sh.start[d][tid] = top.start;
                ++depth;
                cached = cached < kStack ? cached + 1 : kStack;
            }
            top = cur;
            has_top = true;
            ++i;
            cp_async_wait<kYRing - 1>();  // element i has landed
            const double yi = (double)*(const T*)&sh.y[i % kYRing][tid];
            if (i + kYRing < n) cp_async_elem<T>(&sh.y[i % kYRing][tid], p + (long long)(i + kYRing) * step);
            cp_async_commit();
            cur.start = i;
            cur.sy = yi;
            cur.sy2 = __dmul_rn(yi, yi);
            cur.level = yi;
            cum = __dadd_rn(cum, cur.sy2);
        }
    }
    cp_async_wait<0>();
}

// Phase 1 of the fit reconstruction (_compute_isotonic_from_index, :72-81): walk the block ends of the length-`len`
// prefix from the back and write the block list (start, thresholded level :64-67) top-down into rec[len-1-k] — always
// inside the part of the prefix the walk has already passed.  Returns the number of blocks K; the list, read from
// rec[len-K] upwards, is sorted by ascending block start.
__device__ __forceinline__ int list_blocks(int len, bool nn, Rec* __restrict__ rec) {
    int idx = len - 1, k = 0;
    while (idx >= 0) {
        const int s = rec[idx].start;
        const double level = __ddiv_rn(rec[idx].sy, (double)(idx - s + 1));
        Rec* out = rec + (len - 1 - k);
        out->err_after = (nn && level < 0.0) ? 0.0 : level;
        out->start = s;
        ++k;
        idx = s - 1;
    }
    return k;
}

// Phase 2: uniform pass over the elements: aux = level of the block the element belongs to, dual = V - aux (V arrives
// in dual).  Sequence position j -> row (rev ? n-1-j : j).
template <typename T>
__device__ __forceinline__ void fill_prefix(T* __restrict__ aux, T* __restrict__ dual, long long ld, int n, int len,
                                            bool rev, const Rec* __restrict__ rec, int K) {
    const long long step = rev ? -ld : ld;
    const long long o0 = rev ? (long long)(n - 1) * ld : 0LL;
    const Rec* e = rec + (len - K);  // first block (start 0)
    const Rec* e_end = rec + len;
    T v = (T)e->err_after;
    ++e;
    int next_s = e < e_end ? e->start : len;
    T next_v = e < e_end ? (T)e->err_after : T(0);
    constexpr int U = 8;  // independent loads in flight per trip
    for (int j0 = 0; j0 < len; j0 += U) {
        T vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vv[u] = (j0 + u < len) ? dual[o0 + (long long)(j0 + u) * step] : T(0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            if (j < len) {
                if (j == next_s) {
                    v = next_v;
                    ++e;
                    next_s = e < e_end ? e->start : len;
                    next_v = e < e_end ? (T)e->err_after : T(0);
                }
                const long long o = o0 + (long long)j * step;
                aux[o] = v;
                dual[o] = vv[u] - v;
            }
        }
    }
}

// Residency: measured at config 3 (8 192 slices x 8 columns x 1 024 points) in round 1 with YRING = 8, STACK = 4: 4 CTAs
// per SM (<= 128 registers, ~104 KB of shared memory, the rest of the 256 KB stays L1 for the per-thread records)
// 3.94 ms; 5 CTAs (89 registers, the compiler's own choice) 4.27 ms; 6 CTAs 4.51 ms; 8 CTAs with the maximum
// shared-memory carve-out 5.50 ms.  The variant (ring depth, stack-cache depth, CTAs per SM) is a template parameter
// pack selected by B2_OPT_UNIMODAL_VARIANT so that alternatives can be A/B-ed inside one process.
template <typename T, int YRING, int STACK, int MINCTAS>
__global__ void __launch_bounds__(kThreads, MINCTAS)
unimodal_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int n_groups, int R,
                int max_rows, int nn_flag, int32_t* __restrict__ peaks, unsigned char* __restrict__ ws,
                long long ncolslots, size_t thread_bytes) {
    extern __shared__ __align__(16) unsigned char uni_smem[];
    SharedState<YRING, STACK>& sh = *(SharedState<YRING, STACK>*)uni_smem;
    const int tid = threadIdx.x, lane = tid & 31;
    const int rev = lane >> 4;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long slot = warp * 16 + (lane & 15);  // column slot shared by the lane pair (l, l + 16)
    const bool nn = nn_flag != 0;
    const long long sl = slot < ncolslots ? slot : 0;
    Rec* mine = (Rec*)(ws + (size_t)(2 * sl + rev) * thread_bytes);
    const Rec* other = (const Rec*)(ws + (size_t)(2 * sl + (rev ^ 1)) * thread_bytes);
    const long long total = (long long)n_groups * R;
    const long long rounds = (total + ncolslots - 1) / ncolslots;
    for (long long rd = 0; rd < rounds; ++rd) {
        const long long colid = rd * ncolslots + slot;
        const bool active = slot < ncolslots && colid < total;
        int n = 0;
        long long base = 0;
        if (active) {
            const int g = (int)(colid / R), c = (int)(colid - (long long)g * R);
            const long long r0 = row_off[g];
            n = (int)(row_off[g + 1] - r0);
            base = r0 * R + c;
        }
        if (active && n > 0) pava_prefix<T, YRING, STACK>(dual + base, R, n, rev != 0, nn, mine, sh, tid);
        __syncwarp();  // the partner lane's prefix errors are read below
        // peak: first strict minimum of errL[i] + errR[n - i], i = 0..n (:84-92); the pair splits the range.
        // error[0] = 0 is implicit, error[k] = rec[k-1].err_after.
        double best = 0.0;
        int bidx = 0;
        if (active && n > 0) {
            const Rec* rL = rev ? other : mine;
            const Rec* rR = rev ? mine : other;
            const int mid = (n + 1) / 2;
            const int lo = rev ? mid : 0, hi = rev ? n + 1 : mid;
            auto cand_at = [&](int i) {
                const double eL = i > 0 ? rL[i - 1].err_after : 0.0;
                const double eR = i < n ? rR[n - i - 1].err_after : 0.0;
                return __dadd_rn(eL, eR);
            };
            best = cand_at(lo);
            bidx = lo;
            for (int i0 = lo + 1; i0 < hi; i0 += 8) {
                double cand[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) cand[u] = (i0 + u < hi) ? cand_at(i0 + u) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (i0 + u < hi && cand[u] < best) {
                        best = cand[u];
                        bidx = i0 + u;
                    }
                }
            }
        }
        {
            const double ob = __shfl_xor_sync(0xffffffffu, best, 16);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, 16);
            // the lower half wins ties (first minimum); for the rev lane the partner holds the lower half
            if (rev ? !(best < ob) : (ob < best)) {
                best = ob;
                bidx = oi;
            }
        }
        __syncwarp();  // both lanes are done reading each other's records before they are recycled as block lists
        const int len = (active && n > 0) ? (rev ? n - bidx : bidx) : 0;
        const int K = list_blocks(len, nn, mine);
        if (len > 0) fill_prefix<T>(aux + base, dual + base, R, n, len, rev != 0, mine, K);
        if (active && n > 0 && !rev && peaks) peaks[colid] = bidx;
        __syncwarp();  // scratch is reused by the next round
    }
}

size_t per_thread_bytes(int max_rows) {
    return (size_t)max_rows * 32;  // one Rec per element
}

}  // namespace

extern "C" {

size_t b2_unimodal_workspace_bytes(int n_groups, int R, int max_rows) {
    long long total = (long long)n_groups * R;
    const long long cap = (long long)b2_num_sms() * 1024;  // one resident wave of thread pairs is enough
    if (total > cap) total = cap;
    total = (total + 15) / 16 * 16;
    return (size_t)total * 2 * per_thread_bytes(max_rows) + 256;
}

int b2_prox_unimodal(void* aux, void* dual, const int64_t* row_off, int n_groups, int R, int max_rows,
                     int non_negativity, int32_t* peaks, int dtype, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0 || max_rows == 0) return B2_OK;
    const long long total = (long long)n_groups * R;
    const size_t tb = per_thread_bytes(max_rows);
    long long ncolslots = (long long)((ws_bytes > 256 ? ws_bytes - 256 : 0) / (2 * tb));
    ncolslots = ncolslots / 16 * 16;
    const long long want = (total + 15) / 16 * 16;
    if (ncolslots > want) ncolslots = want;
    B2_REQUIRE(ncolslots >= 16, "b2_prox_unimodal workspace too small (%zu bytes for max_rows=%d)", ws_bytes, max_rows);
    B2_REQUIRE(((uintptr_t)ws) % 16 == 0, "workspace must be 16-byte aligned");
    const long long threads = ncolslots * 2;  // 16 column slots per warp
    const int grid = (int)((threads + kThreads - 1) / kThreads);
    const int variant = b2_option_value(B2_OPT_UNIMODAL_VARIANT);
#define B2_UNI_LAUNCH(YR, SK, MC)                                                                                     \
    B2_DISPATCH_DTYPE(dtype, {                                                                                        \
        auto kern = unimodal_kernel<T, YR, SK, MC>;                                                                   \
        const int smem = (int)sizeof(SharedState<YR, SK>);                                                            \
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                 \
        kern<<<grid, kThreads, smem, st>>>((T*)aux, (T*)dual, row_off, n_groups, R, max_rows, non_negativity, peaks,  \
                                           (unsigned char*)ws, ncolslots, tb);                                        \
        B2_LAUNCH_CHECK();                                                                                            \
    })
    switch (variant) {
        case 1: B2_UNI_LAUNCH(4, 8, 4); break;    // deeper stack cache, shorter ring: 40 KB per CTA
        case 2: B2_UNI_LAUNCH(8, 8, 4); break;    // deeper stack cache: 44 KB per CTA
        case 3: B2_UNI_LAUNCH(4, 16, 3); break;   // stack cache covers noise-like data completely: 76 KB per CTA
        case 4: B2_UNI_LAUNCH(4, 8, 5); break;
        case 5: B2_UNI_LAUNCH(4, 12, 4); break;
        default: B2_UNI_LAUNCH(8, 4, 4); break;   // round-1 configuration
    }
#undef B2_UNI_LAUNCH
    return B2_OK;
}

}  // extern "C"
