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
template <int YRING, int STACK, int NT = kThreads>
struct SharedState {
    double y[YRING][NT];
    double2 sums[STACK][NT];   // sy, sy2
    double2 lvl_err[STACK][NT];  // level, err_after
    int start[STACK][NT];
};

template <typename T>
__device__ __forceinline__ void cp_async_elem(void* smem_dst, const T* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// FAST variants: one 256-bit store / load per record (STG.256 / LDG.256 on sm_100a; records are 32-byte aligned)
template <bool FAST>
__device__ __forceinline__ void store_rec(Rec* r, double err_after, double sy, double sy2, int start) {
    if constexpr (FAST) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(r), "d"(err_after), "d"(sy), "d"(sy2),
                     "d"(__hiloint2double(0, start))
                     : "memory");
    } else {
        double2* q = (double2*)r;
        q[0] = make_double2(err_after, sy);
        q[1] = make_double2(sy2, __hiloint2double(0, start));
    }
}
template <bool FAST>
__device__ __forceinline__ void load_rec(const Rec* r, double& err_after, double& sy, double& sy2, int& start) {
    if constexpr (FAST) {
        double w;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(err_after), "=d"(sy), "=d"(sy2), "=d"(w) : "l"(r) : "memory");
        start = __double2loint(w);
    } else {
        start = r->start;
        sy = r->sy;
        sy2 = r->sy2;
        err_after = r->err_after;
    }
}

// a / cnt, correctly rounded (== __ddiv_rn), for an integer-valued divisor.  y = RN(1 / cnt) and two Markstein
// correction steps: q1 = q0 + y (a - cnt q0) is a faithful quotient, and for a faithful q1 and a correctly rounded
// reciprocal RN(q1 + y (a - cnt q1)) is the correctly rounded quotient (Markstein 1990; checked against exact rational
// arithmetic and on the device against __ddiv_rn, tests/test_gpu_kernels.py).  5 fp64 operations after the reciprocal
// instead of the ~35-instruction IEEE division sequence with its slow-path branch.  Tiny numerators (the residuals
// could go subnormal) and non-finite ones take the IEEE division.
template <bool FAST>
__device__ __forceinline__ double div_count(double a, double cnt) {
    if constexpr (FAST) {
        const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
        if (e - 128u < 1919u) {  // 2^-895 <= |a| < 2^1024 (finite): residuals stay normal
            const double y = __drcp_rn(cnt);
            const double q0 = __dmul_rn(a, y);
            const double r0 = __fma_rn(-cnt, q0, a);
            const double q1 = __fma_rn(r0, y, q0);
            const double r1 = __fma_rn(-cnt, q1, a);
            return __fma_rn(r1, y, q1);
        }
    }
    return __ddiv_rn(a, cnt);
}

// div_count with the correctly rounded reciprocal y = RN(1 / cnt) supplied by the caller (one reciprocal serves the two
// divisions of a trip, see pava_prefix_ec)
__device__ __forceinline__ double div_count_y(double a, double cnt, double y) {
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    if (e - 128u < 1919u) {
        const double q0 = __dmul_rn(a, y);
        const double r0 = __fma_rn(-cnt, q0, a);
        const double q1 = __fma_rn(r0, y, q0);
        const double r1 = __fma_rn(-cnt, q1, a);
        return __fma_rn(r1, y, q1);
    }
    return __ddiv_rn(a, cnt);
}

// shared state of the variants that also keep a COMPACT copy of the prefix errors (pava_prefix_ec): four staged
// values per thread, written out as one 32-byte sector per four elements
template <int YRING, int STACK, int NT = kThreads>
struct SharedStateE : SharedState<YRING, STACK, NT> {
    double e[4][NT];
};

template <int YRING, int STACK, int NT, bool WITH_E>
struct UniShared {
    using type = SharedState<YRING, STACK, NT>;
};
template <int YRING, int STACK, int NT>
struct UniShared<YRING, STACK, NT, true> {
    using type = SharedStateE<YRING, STACK, NT>;
};

__device__ __forceinline__ void store4(double* dst, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void load4(const double* src, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(src) : "memory");
}

// Prefix-isotonic PAVA over seq(j) = col[(rev ? n-1-j : j) * ld], j in [0, n); fills rec[0..n).
template <typename T, int kYRing, int kStack, bool FAST, class SH>
__device__ __forceinline__ void pava_prefix(const T* __restrict__ col, long long ld, int n, bool rev, bool nn,
                                            Rec* __restrict__ rec, SH& sh, int tid) {
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
                    load_rec<FAST>(rec + (cur.start - 1), top.err_after, top.sy, top.sy2, top.start);
                    top.level = div_count<FAST>(top.sy, (double)(cur.start - top.start));
                }
            }
            num = cur.sy;
        } else {
            num = __dmul_rn(cur.sy, cur.sy);
        }
        const double q = div_count<FAST>(num, (double)(i - cur.start + 1));
        if (merge) {
            cur.level = q;  // (:21)
        } else {
            const double levelerror = __dsub_rn(cur.sy2, q);  // (:57)
            const double before = has_top ? top.err_after : 0.0;
            cur.err_after = (nn && cur.level < 0.0) ? cum : __dadd_rn(levelerror, before);  // (:58-62)
            store_rec<FAST>(rec + i, cur.err_after, cur.sy, cur.sy2, cur.start);
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

// Same PAVA, two units of work per trip: a lane whose current block must absorb the block below does that merge, and
// — if the merged block now sits above its new neighbour (or on the floor) — finalises the element in the SAME trip
// instead of the next one.  The instruction stream of a trip is [merge part | finalise part], which is what a divergent
// trip of the one-unit loop executes anyway (merging and finalising lanes are serialised), plus the second division;
// a column that merges once per element (noise-like data) then takes ~n trips instead of ~2 n.  Same operations in the
// same order per column, so levels, errors and the peak index stay bit-identical.
template <typename T, int kYRing, int kStack, class SH>
__device__ __forceinline__ void pava_prefix_dual(const T* __restrict__ col, long long ld, int n, bool rev, bool nn,
                                                 Rec* __restrict__ rec, SH& sh, int tid) {
    const long long step = rev ? -ld : ld;
    const T* p = col + (rev ? (long long)(n - 1) * ld : 0LL);
#pragma unroll
    for (int u = 1; u <= kYRing; ++u) {
        if (u < n) cp_async_elem<T>(&sh.y[u % kYRing][tid], p + (long long)u * step);
        cp_async_commit();
    }
    Block cur, top;
    bool has_top = false;
    int depth = 0, cached = 0;
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
    double cum = cur.sy2;
    while (i < n) {
        bool merge = has_top && cur.level <= top.level;
        if (merge) {
            cur.sy = __dadd_rn(cur.sy, top.sy);
            cur.sy2 = __dadd_rn(cur.sy2, top.sy2);
            cur.start = top.start;
            has_top = cur.start > 0;
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
                } else {
                    load_rec<true>(rec + (cur.start - 1), top.err_after, top.sy, top.sy2, top.start);
                    top.level = div_count<true>(top.sy, (double)(cur.start - top.start));
                }
            }
            cur.level = div_count<true>(cur.sy, (double)(i - cur.start + 1));
            merge = has_top && cur.level <= top.level;  // another merge is due: next trip
        }
        if (!merge) {
            const double q = div_count<true>(__dmul_rn(cur.sy, cur.sy), (double)(i - cur.start + 1));
            const double levelerror = __dsub_rn(cur.sy2, q);
            const double before = has_top ? top.err_after : 0.0;
            cur.err_after = (nn && cur.level < 0.0) ? cum : __dadd_rn(levelerror, before);
            store_rec<true>(rec + i, cur.err_after, cur.sy, cur.sy2, cur.start);
            if (has_top) {
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
            cp_async_wait<kYRing - 1>();
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

// pava_prefix_dual plus a COMPACT copy of the prefix errors for the peak search.  The peak search needs only
// error[i] (8 of the 32 bytes of a record), so reading it from the records costs a 32-byte sector per candidate and
// direction — at config 3 the whole 4.3 GB record array a second time, a DRAM-bound phase that took a quarter of the
// kernel (ncu source view, profiles/r2_ncu_unimodal_v9_src_s34.txt).  Here every thread also writes E[p] = the error both
// directions contribute to peak candidate p (forward: error[p]; reversed: error[n - p], i.e. stored back to front), four
// values staged on chip and written as ONE aligned 32-byte sector per four elements, so the peak search reads
// 2 x 8 bytes per candidate with 256-bit loads (4 candidates per load and direction).
// EC = 1: the four values are staged in shared memory; EC = 2: in a register shift chain.
// Also: one reciprocal per trip.  A finalisation either follows a merge of the same trip (same block count as the level
// division just made: same reciprocal) or belongs to a fresh one-element block (count 1: RN(1/1) = 1 and the Markstein
// steps return the numerator unchanged), so the second __drcp_rn of a trip is never needed.
template <typename T, int kYRing, int kStack, int EC, class SH>
__device__ __forceinline__ void pava_prefix_ec(const T* __restrict__ col, long long ld, int n, bool rev, bool nn,
                                               Rec* __restrict__ rec, double* __restrict__ E, SH& sh, int tid) {
    const long long step = rev ? -ld : ld;
    const T* p = col + (rev ? (long long)(n - 1) * ld : 0LL);
#pragma unroll
    for (int u = 1; u <= kYRing; ++u) {
        if (u < n) cp_async_elem<T>(&sh.y[u % kYRing][tid], p + (long long)u * step);
        cp_async_commit();
    }
    Block cur, top;
    bool has_top = false;
    int depth = 0, cached = 0;
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
    double cum = cur.sy2;
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;  // EC == 2: E[pos], E[pos -+ 1], ... of the group being filled
    const int trig = rev ? 0 : 3;                   // position (mod 4) that completes a group of four
    auto stage = [&](int pos, double v) {
        if constexpr (EC == 1) {
            sh.e[pos & 3][tid] = v;
            if ((pos & 3) == trig) store4(E + (pos & ~3), sh.e[0][tid], sh.e[1][tid], sh.e[2][tid], sh.e[3][tid]);
        } else {
            e3 = e2;
            e2 = e1;
            e1 = e0;
            e0 = v;
            if ((pos & 3) == trig) {
                if (rev) store4(E + pos, e0, e1, e2, e3);
                else store4(E + pos - 3, e3, e2, e1, e0);
            }
        }
    };
    stage(rev ? n : 0, 0.0);  // error[0] = 0
    while (i < n) {
        bool merge = has_top && cur.level <= top.level;
        double cnt = 1.0, y = 1.0;
        if (merge) {
            cur.sy = __dadd_rn(cur.sy, top.sy);
            cur.sy2 = __dadd_rn(cur.sy2, top.sy2);
            cur.start = top.start;
            cnt = (double)(i - cur.start + 1);
            y = __drcp_rn(cnt);
            has_top = cur.start > 0;
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
                } else {
                    load_rec<true>(rec + (cur.start - 1), top.err_after, top.sy, top.sy2, top.start);
                    top.level = div_count<true>(top.sy, (double)(cur.start - top.start));
                }
            }
            cur.level = div_count_y(cur.sy, cnt, y);
            merge = has_top && cur.level <= top.level;  // another merge is due: next trip
        }
        if (!merge) {
            const double q = div_count_y(__dmul_rn(cur.sy, cur.sy), cnt, y);
            const double levelerror = __dsub_rn(cur.sy2, q);
            const double before = has_top ? top.err_after : 0.0;
            cur.err_after = (nn && cur.level < 0.0) ? cum : __dadd_rn(levelerror, before);
            store_rec<true>(rec + i, cur.err_after, cur.sy, cur.sy2, cur.start);
            stage(rev ? n - 1 - i : i + 1, cur.err_after);
            if (has_top) {
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
            cp_async_wait<kYRing - 1>();
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
    if (!rev && (n & 3) != 3) {  // the forward direction ends inside a group: positions (n & ~3) .. n
        double* dst = E + (n & ~3);
        if constexpr (EC == 1) {
            store4(dst, sh.e[0][tid], sh.e[1][tid], sh.e[2][tid], sh.e[3][tid]);
        } else {
            const int c = n & 3;  // c + 1 valid values, the newest (position n) in e0
            if (c == 0) store4(dst, e0, 0.0, 0.0, 0.0);
            else if (c == 1) store4(dst, e1, e0, 0.0, 0.0);
            else store4(dst, e2, e1, e0, 0.0);
        }
    }
}

// Phase 1 of the fit reconstruction (_compute_isotonic_from_index, :72-81): walk the block ends of the length-`len`
// prefix from the back and write the block list (start, thresholded level :64-67) top-down into rec[len-1-k] — always
// inside the part of the prefix the walk has already passed.  Returns the number of blocks K; the list, read from
// rec[len-K] upwards, is sorted by ascending block start.
template <bool FAST>
__device__ __forceinline__ int list_blocks(int len, bool nn, Rec* __restrict__ rec) {
    int idx = len - 1, k = 0;
    while (idx >= 0) {
        int s;
        double sy;
        if constexpr (FAST) {
            double e_, sy2_;
            load_rec<true>(rec + idx, e_, sy, sy2_, s);
        } else {
            s = rec[idx].start;
            sy = rec[idx].sy;
        }
        const double level = div_count<FAST>(sy, (double)(idx - s + 1));
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
template <typename T, int U = 8>
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
    // U independent loads in flight per trip
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
template <typename T, int YRING, int STACK, int MINCTAS, bool FAST, bool DUAL, int NT = kThreads, int EC = 0, int FU = 8>
__global__ void __launch_bounds__(NT, MINCTAS)
unimodal_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int n_groups, int R,
                int max_rows, int nn_flag, int32_t* __restrict__ peaks, unsigned char* __restrict__ ws,
                long long ncolslots, size_t thread_bytes) {
    extern __shared__ __align__(16) unsigned char uni_smem[];
    using SH = typename UniShared<YRING, STACK, NT, EC == 1>::type;
    SH& sh = *(SH*)uni_smem;
    const int tid = threadIdx.x, lane = tid & 31;
    const int rev = lane >> 4;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long slot = warp * 16 + (lane & 15);  // column slot shared by the lane pair (l, l + 16)
    const bool nn = nn_flag != 0;
    const long long sl = slot < ncolslots ? slot : 0;
    Rec* mine = (Rec*)(ws + (size_t)(2 * sl + rev) * thread_bytes);
    const Rec* other = (const Rec*)(ws + (size_t)(2 * sl + (rev ^ 1)) * thread_bytes);
    // compact prefix errors (EC variants): behind the max_rows records of the thread
    double* mineE = (double*)(mine + max_rows);
    const double* otherE = (const double*)(other + max_rows);
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
        if (active && n > 0) {
            if constexpr (EC != 0)
                pava_prefix_ec<T, YRING, STACK, EC, SH>(dual + base, R, n, rev != 0, nn, mine, mineE, sh, tid);
            else if constexpr (DUAL)
                pava_prefix_dual<T, YRING, STACK, SH>(dual + base, R, n, rev != 0, nn, mine, sh, tid);
            else
                pava_prefix<T, YRING, STACK, FAST, SH>(dual + base, R, n, rev != 0, nn, mine, sh, tid);
        }
        __syncwarp();  // the partner lane's prefix errors are read below
        // peak: first strict minimum of errL[i] + errR[n - i], i = 0..n (:84-92); the pair splits the range.
        // error[0] = 0 is implicit, error[k] = rec[k-1].err_after.
        double best = 0.0;
        int bidx = 0;
        if constexpr (EC != 0) {
            // candidates in aligned groups of four from the compact copies: EL[i] + ER[i], i = 0..n
            if (active && n > 0) {
                const double* EL = rev ? otherE : mineE;
                const double* ER = rev ? mineE : otherE;
                const int G = n / 4 + 1, gm = (G + 1) / 2;
                const int glo = rev ? gm : 0, ghi = rev ? G : gm;
                best = __longlong_as_double(0x7ff0000000000000LL);  // +inf: the first finite candidate replaces it
                bidx = 4 * glo;
                constexpr int PG = 2;  // groups (2 x 256-bit loads each) in flight per trip
                for (int g0 = glo; g0 < ghi; g0 += PG) {
                    double l[PG][4], r[PG][4];
#pragma unroll
                    for (int u = 0; u < PG; ++u) {
                        if (g0 + u < ghi) {
                            load4(EL + 4 * (g0 + u), l[u][0], l[u][1], l[u][2], l[u][3]);
                            load4(ER + 4 * (g0 + u), r[u][0], r[u][1], r[u][2], r[u][3]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PG; ++u) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int c = 4 * (g0 + u) + k;
                            if (g0 + u < ghi && c <= n) {
                                const double cand = __dadd_rn(l[u][k], r[u][k]);
                                if (cand < best) {
                                    best = cand;
                                    bidx = c;
                                }
                            }
                        }
                    }
                }
            }
        } else if (active && n > 0) {
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
            constexpr int PU = FAST ? 16 : 8;  // candidates (2 loads each) in flight per trip
            for (int i0 = lo + 1; i0 < hi; i0 += PU) {
                double cand[PU];
#pragma unroll
                for (int u = 0; u < PU; ++u) cand[u] = (i0 + u < hi) ? cand_at(i0 + u) : 0.0;
#pragma unroll
                for (int u = 0; u < PU; ++u) {
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
        const int K = list_blocks<FAST>(len, nn, mine);
        if constexpr (FU > 0) {
            if (len > 0) fill_prefix<T, FU>(aux + base, dual + base, R, n, len, rev != 0, mine, K);
        } else {
            // deferred fill (one round only: every column owns its scratch until unimodal_fill_kernel has run):
            // leave (prefix length, number of blocks) behind the compact errors
            if (slot < ncolslots) *(int2*)((unsigned char*)mine + thread_bytes - 32) = make_int2(len, K);
        }
        if (active && n > 0 && !rev && peaks) peaks[colid] = bidx;
        __syncwarp();  // scratch is reused by the next round
    }
}

// Second kernel of the deferred-fill variants: aux = fitted value, dual = V - aux from the block lists the PAVA kernel
// left in the scratch area (list_blocks: (start, level) of the blocks of the winning prefix per column and direction).
// Inside the PAVA kernel this pass costs a quarter of the time (ncu, profiles/r2_ncu_unimodal_v11_full_s36.txt): every lane
// walks its own prefix (the warp runs max_lanes(len) steps with most lanes idle near the end, the two directions of a
// column write the halves of a 32-byte sector at different times -> partial-sector writes).  As a kernel of its own it is a
// plain streaming pass at full occupancy: a CTA takes a slice, thread (chunk k, column c) walks rows [k CH, (k+1) CH)
// of column c, so a warp step touches whole rows (R consecutive values) of 32 / R chunks.
// Rows below the peak t* = len_L come from the forward list at position r, rows from t* on from the reversed list at
// position n - 1 - r (walked downwards).
template <typename T>
__global__ void __launch_bounds__(256)
unimodal_fill_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int n_groups, int R,
                     const unsigned char* __restrict__ ws, size_t thread_bytes) {
    const int cpb = blockDim.x / R;  // row chunks of a slice in flight
    const int c = threadIdx.x % R, k = threadIdx.x / R;
    if (k >= cpb) return;
    constexpr int U = 8;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long r0 = row_off[g];
        const int n = (int)(row_off[g + 1] - r0);
        if (n == 0) continue;
        const long long colid = (long long)g * R + c;
        const unsigned char* bL = ws + (size_t)(2 * colid) * thread_bytes;
        const unsigned char* bR = bL + thread_bytes;
        const int2 hL = *(const int2*)(bL + thread_bytes - 32), hR = *(const int2*)(bR + thread_bytes - 32);
        const int tstar = hL.x;  // rows [0, t*) forward fit, [t*, n) reversed fit; hR.x == n - t*
        const Rec* lL = (const Rec*)bL + (hL.x - hL.y);  // (level in err_after, start), ascending block starts
        const Rec* lR = (const Rec*)bR + (hR.x - hR.y);
        const int nch = cpb * (int)gridDim.y;  // gridDim.y CTAs share a slice when there are few slices
        const int CH = (n + nch - 1) / nch;
        const int ra = min(n, ((int)blockIdx.y * cpb + k) * CH), rb = min(n, ra + CH);
        T* ap = aux + r0 * R + c;
        T* dp = dual + r0 * R + c;
        // forward part: rows [ra, min(rb, t*))
        int r = ra;
        const int re = min(rb, tstar);
        if (r < re) {
            int lo = 0, hi = hL.y - 1;  // last block with start <= r
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (lL[mid].start <= r) lo = mid;
                else hi = mid - 1;
            }
            int b = lo;
            T v = (T)lL[b].err_after;
            int next_s = b + 1 < hL.y ? lL[b + 1].start : tstar;
            for (int j0 = r; j0 < re; j0 += U) {
                T vv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) vv[u] = (j0 + u < re) ? dp[(long long)(j0 + u) * R] : T(0);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u;
                    if (j < re) {
                        if (j == next_s) {
                            ++b;
                            v = (T)lL[b].err_after;
                            next_s = b + 1 < hL.y ? lL[b + 1].start : tstar;
                        }
                        ap[(long long)j * R] = v;
                        dp[(long long)j * R] = vv[u] - v;
                    }
                }
            }
        }
        // reversed part: rows [max(ra, t*), rb), sequence position q = n - 1 - row, descending
        r = max(ra, tstar);
        if (r < rb) {
            const int q0 = n - 1 - r;
            int lo = 0, hi = hR.y - 1;  // last block with start <= q0
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (lR[mid].start <= q0) lo = mid;
                else hi = mid - 1;
            }
            int b = lo;
            T v = (T)lR[b].err_after;
            int cur_s = lR[b].start;
            for (int j0 = r; j0 < rb; j0 += U) {
                T vv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) vv[u] = (j0 + u < rb) ? dp[(long long)(j0 + u) * R] : T(0);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u;
                    if (j < rb) {
                        const int q = n - 1 - j;
                        if (q < cur_s) {
                            --b;
                            v = (T)lR[b].err_after;
                            cur_s = lR[b].start;
                        }
                        ap[(long long)j * R] = v;
                        dp[(long long)j * R] = vv[u] - v;
                    }
                }
            }
        }
    }
}

// self-test of div_count<true> against the IEEE division: pseudo-random numerators over the whole exponent range (and
// values next to the guard thresholds), every divisor 1..max_cnt
__global__ void div_selftest_kernel(long long n, unsigned long long seed, int max_cnt, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);  // splitmix64
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        unsigned long long y = x * 0xD1342543DE82EF95ull + 1;
        const int cnt = 1 + (int)(y % (unsigned long long)max_cnt);
        double a;
        const int kind = (int)((y >> 40) & 7);
        if (kind < 5) {  // moderate magnitudes (sums of data values and their squares)
            const unsigned long long e = 1023 - 40 + ((y >> 20) % 80);
            a = __longlong_as_double((long long)((x & 0x800FFFFFFFFFFFFFull) | (e << 52)));
        } else if (kind < 7) {  // any finite bit pattern, zeros and subnormals included
            a = __longlong_as_double((long long)x);
            if (!isfinite(a)) a = 0.0;
        } else {  // exact multiples and near-multiples of the divisor
            a = (double)cnt * (double)(long long)((x >> 11) & 0xFFFFFFFFFFull) * (1.0 + (double)((y >> 8) & 3) * 0x1p-52);
        }
        const double f = div_count<true>(a, (double)cnt), r = __ddiv_rn(a, (double)cnt);
        if (__double_as_longlong(f) != __double_as_longlong(r) && !(f == 0.0 && r == 0.0)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// Measured and rejected (profiles/r2_ab_unimodal_*.log, config-3 size, bit-identical results):
//  * the reciprocals 1/count from an L1-resident table instead of __drcp_rn's 6 dependent operations: 3.89 ms against
//    3.37 ms — one more memory operation per division costs more than the arithmetic it replaces;
//  * more resident threads through 64- or 96-thread CTAs (576 / 640 threads per SM instead of 512): 3.8 - 7.5 ms against
//    3.24 ms — like 5, 6 and 8 CTAs of 128 threads in round 1, every thread beyond ~512 per SM costs more L1 (per-thread
//    records) than it hides latency;
//  * the block below the top prefetched into registers (a merge promotes it at once, the next one is requested early):
//    4.18 ms against 3.34 ms — the extra state and instructions cost more than the hidden round trip;
//  * a two-pass "replay" formulation (pass 1 writes only the 8-byte prefix errors, the block stack lives in the
//    shared-memory window and spills when it outgrows it; pass 2 re-runs the PAVA on the winning prefix): less than
//    half the DRAM bytes, but 4.01 ms against 3.30 ms on noise-like input and the same 4.63 ms on peak-like input — the
//    kernel is bound by the serial dependency chain per thread, not by DRAM, so a second pass over part of the column
//    costs more than the traffic it saves.
//  * (last session, profiles/r2_ab_unimodal_compact_s40.log, r2_ncu_unimodal_v16_full_s39.txt) 20-byte records
//    (error, level, block start) with the sums of deep pops in a spill stack indexed by stack depth, written / re-read
//    four entries at a time: DRAM bytes 10.6 -> 7.6 GB, 25 % fewer instructions, 87 registers, bit-identical - and
//    6.3 ms against 2.69 ms: two scattered store instructions per trip (one sector per lane each) instead of 1.25 back up
//    the L1 data pipe, and the shared-memory pops behind them wait;
//  * (profiles/r2_ab_unimodal_inplace_s42.log) the top block updated in place (a pooling element is added to the
//    register block, only an element that starts a block pushes): half the ring traffic on paper, but three paths per
//    trip instead of two, which a divergent warp all executes: 3.38 ms against 2.69 ms.
//  * (profiles/r2_ab_unimodal_pg4_fill64_s48.log) four instead of two groups of peak candidates in flight (spills at
//    the 128-register cap: 2.93 against 2.69 ms) and >= 64 instead of 32 rows per thread in the fill kernel (fewer
//    block-list searches, but half the threads per slice: 2.85 ms).
size_t per_thread_bytes(int max_rows) {
    // one Rec per element + the compact prefix errors E[0..max_rows] in whole 32-byte sectors + a 32-byte header
    // ((prefix length, number of blocks) for the deferred fill)
    return (size_t)max_rows * 32 + (size_t)(max_rows / 4 + 1) * 32 + 32;
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

int b2_selftest_div_count(long long n, unsigned long long seed, int max_cnt, unsigned long long* mismatches_dev,
                          void* stream) {
    B2_REQUIRE(n >= 0 && max_cnt >= 1 && mismatches_dev, "bad arguments");
    B2_CHECK_CUDA(cudaMemsetAsync(mismatches_dev, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    div_selftest_kernel<<<b2_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(n, seed, max_cnt, mismatches_dev);
    B2_LAUNCH_CHECK();
    return B2_OK;
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
    B2_REQUIRE(((uintptr_t)ws) % 32 == 0, "workspace must be 32-byte aligned");
    const long long threads = ncolslots * 2;  // 16 column slots per warp
    const int variant = b2_option_value(B2_OPT_UNIMODAL_VARIANT);
#define B2_UNI_LAUNCH_NT(YR, SK, MC, FAST, DUAL, NTV, ECV, FUV)                                                       \
    B2_DISPATCH_DTYPE(dtype, {                                                                                        \
        auto kern = unimodal_kernel<T, YR, SK, MC, FAST, DUAL, NTV, ECV, FUV>;                                        \
        const int smem = (int)sizeof(typename UniShared<YR, SK, NTV, ECV == 1>::type);                                \
        const int grid_nt = (int)((threads + NTV - 1) / NTV);                                                         \
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                 \
        kern<<<grid_nt, NTV, smem, st>>>((T*)aux, (T*)dual, row_off, n_groups, R, max_rows, non_negativity, peaks,    \
                                         (unsigned char*)ws, ncolslots, tb);                                          \
        B2_LAUNCH_CHECK();                                                                                            \
    })
#define B2_UNI_LAUNCH(YR, SK, MC, FAST, DUAL) B2_UNI_LAUNCH_NT(YR, SK, MC, FAST, DUAL, kThreads, 0, 8)
#define B2_UNI_LAUNCH_EC(YR, SK, MC, ECV, FUV) B2_UNI_LAUNCH_NT(YR, SK, MC, true, true, kThreads, ECV, FUV)
    switch (variant) {
        case 1: B2_UNI_LAUNCH(4, 8, 4, false, false); break;    // deeper stack cache, shorter ring: 40 KB per CTA
        case 2: B2_UNI_LAUNCH(8, 8, 4, false, false); break;    // deeper stack cache: 44 KB per CTA
        case 3: B2_UNI_LAUNCH(4, 16, 3, false, false); break;   // stack cache covers noise-like data completely: 76 KB per CTA
        case 4: B2_UNI_LAUNCH(4, 8, 5, false, false); break;
        case 5: B2_UNI_LAUNCH(4, 12, 4, false, false); break;
        case 6: B2_UNI_LAUNCH(4, 8, 4, true, false); break;     // variant 1 + Markstein division + 256-bit record I/O
        case 9: B2_UNI_LAUNCH(4, 8, 4, true, true); break;   // variant 6 + merge and finalise in one trip
        case 10: B2_UNI_LAUNCH_EC(4, 8, 4, 1, 8); break;     // variant 9 + compact prefix errors (staged in shared memory), one reciprocal per trip
        case 11: B2_UNI_LAUNCH_EC(4, 8, 4, 2, 8); break;     // the same, staged in registers
        case 12: B2_UNI_LAUNCH_EC(4, 8, 4, 1, 16); break;    // variant 10 with 16 loads in flight in the fill pass
        case 13: B2_UNI_LAUNCH_EC(4, 8, 4, 2, 16); break;
        case 14:  // variant 11 with the fill as a second, streaming kernel (needs one round: a scratch slot per column)
        case 15:  // variant 10 likewise
            if (ncolslots >= want) {
                if (variant == 14) B2_UNI_LAUNCH_EC(4, 8, 4, 2, -1);
                else B2_UNI_LAUNCH_EC(4, 8, 4, 1, -1);
                const int nthr = R <= 256 ? 256 / R * R : R;
                const int grid_f = n_groups < b2_num_sms() * 8 ? n_groups : b2_num_sms() * 8;
                int parts = (b2_num_sms() * 4 + grid_f - 1) / grid_f;             // few slices: split the rows of a slice
                const int max_parts = (max_rows + (nthr / R) * 8 - 1) / ((nthr / R) * 8);  // >= 8 rows per thread
                parts = parts > max_parts ? max_parts : parts;
                parts = parts < 1 ? 1 : (parts > 65535 ? 65535 : parts);
                B2_DISPATCH_DTYPE(dtype, {
                    unimodal_fill_kernel<T><<<dim3(grid_f, parts), nthr, 0, st>>>(
                        (T*)aux, (T*)dual, row_off, n_groups, R, (const unsigned char*)ws, tb);
                    B2_LAUNCH_CHECK();
                });
            } else {
                B2_UNI_LAUNCH_EC(4, 8, 4, 2, 8);
            }
            break;
        default: B2_UNI_LAUNCH(8, 4, 4, false, false); break;   // round-1 configuration
    }
#undef B2_UNI_LAUNCH
#undef B2_UNI_LAUNCH_EC
#undef B2_UNI_LAUNCH_NT
    return B2_OK;
}

}  // extern "C"
