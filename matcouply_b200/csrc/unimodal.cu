// Unimodal regression prox (penalties.py:1014-1015 -> _unimodal_regression.py:24-141), one thread per
// (slice, column).  The arithmetic is IEEE round-to-nearest fp64 with explicit non-contracted intrinsics in the
// reference's operation order, so block levels, prefix errors and therefore the peak index t* are bit-identical to
// the NumPy/numba reference on identical input.
//
// Memory-light formulation: instead of materialising both prefix structures (level_set, index_range, error for the
// forward and the reversed vector) we
//   1. run PAVA on the reversed column and store only its prefix errors errR[0..n],
//   2. run PAVA forward, forming errL[i] on the fly and tracking the first strict minimum of errL[i] + errR[n-i],
//   3. re-run PAVA on the two winning prefixes (y[:t*] forward, y[t*:] reversed) and expand their block stacks.
// Re-running PAVA on a prefix reproduces exactly the blocks the reference reads back through index_range, because
// the block stack of a prefix is never modified by later elements.
// PAVA keeps the two top blocks in registers (lazily spilled); deeper blocks live in a per-thread stack in global memory that is
// interleaved across threads (element j of slot t at [j * nslots + t]) so that warps access it coalesced.
#include "common.cuh"

namespace {

struct Block {
    int start;
    double sy, sy2, level, err_after;
};

struct Stack {
    double* sy;
    double* sy2;
    double* level;
    double* err;
    int* start;
    size_t stride;  // nslots
    __device__ __forceinline__ void put(int d, const Block& b) const {
        const size_t o = (size_t)d * stride;
        sy[o] = b.sy;
        sy2[o] = b.sy2;
        level[o] = b.level;
        err[o] = b.err_after;
        start[o] = b.start;
    }
    __device__ __forceinline__ Block get(int d) const {
        const size_t o = (size_t)d * stride;
        Block b;
        b.sy = sy[o];
        b.sy2 = sy2[o];
        b.level = level[o];
        b.err_after = err[o];
        b.start = start[o];
        return b;
    }
};

// Runs prefix-isotonic PAVA over seq(j) = col[(REV ? n-1-j : j) * ld], j in [0, len).
// MODE 0: store prefix errors to errR[(j+1)*stride]           (reversed pass)
// MODE 1: track first strict minimum of errL[i] + errR[n-i]    (forward pass); returns best index through *best_idx
// MODE 2: expand the final block stack into out (same indexing as the input sequence)
template <typename T, bool REV, int MODE>
__device__ __forceinline__ void pava_pass(const T* __restrict__ col, long long ld, int n, int len, bool nn,
                                          const Stack& stk, double* __restrict__ errR, size_t stride,
                                          T* __restrict__ out, int* best_idx) {
    if (MODE == 0) errR[0] = 0.0;
    double best = 0.0;
    int bidx = 0;
    if (MODE == 1) best = errR[(size_t)n * stride];  // error_right[-1]  (:85)
    if (len <= 0) {
        if (MODE == 1) *best_idx = 0;
        return;
    }
    int depth = 0;
    bool has_sec = false;
    Block top, sec;
    sec.start = 0;
    sec.sy = sec.sy2 = sec.level = sec.err_after = 0.0;
    double cum = 0.0;
    {
        const double y0 = (double)col[(REV ? (long long)(n - 1) : 0LL) * ld];
        top.start = 0;
        top.sy = y0;
        top.sy2 = __dmul_rn(y0, y0);
        top.level = y0;
        cum = top.sy2;
        top.err_after = (nn && y0 < 0.0) ? cum : 0.0;  // (:46-48); error[1] stays 0 otherwise
        if (MODE == 0) errR[stride] = top.err_after;
        if (MODE == 1) {
            const double cand = __dadd_rn(top.err_after, errR[(size_t)(n - 1) * stride]);
            // i = 0 candidate equals `best` exactly (0 + errR[n]) and is never strictly smaller; i = 1:
            if (cand < best) {
                best = cand;
                bidx = 1;
            }
        }
    }
    // The loads of y (and of errR in the forward pass) do not depend on the PAVA state: keep PF of them in flight in
    // a register ring so that only stack pops stay on the sequential critical path.
    constexpr int PF = 8;
    double ybuf[PF], ebuf[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        const int i = 1 + u;
        ybuf[u] = (i < len) ? (double)col[(REV ? (long long)(n - 1 - i) : (long long)i) * ld] : 0.0;
        ebuf[u] = (MODE == 1 && i < len) ? errR[(size_t)(n - i - 1) * stride] : 0.0;
    }
    for (int i0 = 1; i0 < len; i0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 + u;
            if (i >= len) break;
            const double yi = ybuf[u];
            const double eR = ebuf[u];
            {
                const int ip = i + PF;
                ybuf[u] = (ip < len) ? (double)col[(REV ? (long long)(n - 1 - ip) : (long long)ip) * ld] : 0.0;
                ebuf[u] = (MODE == 1 && ip < len) ? errR[(size_t)(n - ip - 1) * stride] : 0.0;
            }
            // new single-element block; it absorbs the blocks below while its level is <= theirs (`<=` pooling, :53).
            // Lazy stack: `top`/`sec` live in registers, memory is touched only when a third live block must be
            // spilled (rising runs) or when a cascade drains both registers (falling runs).
            Block cur;
            cur.start = i;
            cur.sy = yi;
            cur.sy2 = __dmul_rn(yi, yi);
            cur.level = yi;
            cum = __dadd_rn(cum, cur.sy2);
            bool has_top = true;
            while (has_top && cur.level <= top.level) {
                cur.sy = __dadd_rn(cur.sy, top.sy);
                cur.sy2 = __dadd_rn(cur.sy2, top.sy2);
                cur.start = top.start;
                cur.level = __ddiv_rn(cur.sy, (double)(i - cur.start + 1));
                if (has_sec) {
                    top = sec;
                    has_sec = false;
                } else if (depth > 0) {
                    top = stk.get(--depth);
                } else {
                    has_top = false;
                }
            }
            const double cnt = (double)(i - cur.start + 1);
            const double levelerror = __dsub_rn(cur.sy2, __ddiv_rn(__dmul_rn(cur.sy, cur.sy), cnt));  // (:57)
            const double before = has_top ? top.err_after : 0.0;
            cur.err_after = (nn && cur.level < 0.0) ? cum : __dadd_rn(levelerror, before);  // (:58-62)
            if (has_top) {
                if (has_sec) stk.put(depth++, sec);
                sec = top;
                has_sec = true;
            }
            top = cur;
            if (MODE == 0) errR[(size_t)(i + 1) * stride] = top.err_after;
            if (MODE == 1) {
                const double cand = __dadd_rn(top.err_after, eR);
                if (cand < best) {  // strict: first minimum wins (:88-91)
                    best = cand;
                    bidx = i + 1;
                }
            }
        }
    }
    if (MODE == 1) *best_idx = bidx;
    if (MODE == 2) {
        // expand blocks from the top of the stack down (:72-81); negative levels were zeroed under non-negativity (:64-67)
        int end = len - 1;
        Block b = top;
        while (true) {
            const T v = (T)((nn && b.level < 0.0) ? 0.0 : b.level);
            for (int j = b.start; j <= end; ++j) out[(REV ? (long long)(n - 1 - j) : (long long)j) * ld] = v;
            end = b.start - 1;
            if (end < 0) break;
            if (has_sec) {
                b = sec;
                has_sec = false;
            } else {
                b = stk.get(--depth);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
unimodal_kernel(T* __restrict__ aux, T* __restrict__ dual, const int64_t* __restrict__ row_off, int n_groups, int R,
                int max_rows, int nn, int32_t* __restrict__ peaks, double* __restrict__ wsd, int* __restrict__ wsi,
                long long nslots) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const size_t stride = (size_t)nslots;
    // workspace carve (doubles): errR[(max_rows+1)] | sy | sy2 | level | err  (each max_rows) ; ints: start[max_rows]
    double* errR = wsd + slot;
    Stack stk;
    stk.stride = stride;
    stk.sy = wsd + (size_t)(max_rows + 1) * stride + slot;
    stk.sy2 = stk.sy + (size_t)max_rows * stride;
    stk.level = stk.sy2 + (size_t)max_rows * stride;
    stk.err = stk.level + (size_t)max_rows * stride;
    stk.start = wsi + slot;
    const long long total = (long long)n_groups * R;
    for (long long colid = slot; colid < total; colid += nslots) {
        const int g = (int)(colid / R), c = (int)(colid - (long long)g * R);
        const long long r0 = row_off[g];
        const int n = (int)(row_off[g + 1] - r0);
        if (n <= 0) continue;
        const T* V = dual + r0 * R + c;
        T* out = aux + r0 * R + c;
        int t = 0;
        pava_pass<T, true, 0>(V, R, n, n, nn != 0, stk, errR, stride, nullptr, nullptr);
        pava_pass<T, false, 1>(V, R, n, n, nn != 0, stk, errR, stride, nullptr, &t);
        pava_pass<T, false, 2>(V, R, n, t, nn != 0, stk, errR, stride, out, nullptr);
        pava_pass<T, true, 2>(V, R, n, n - t, nn != 0, stk, errR, stride, out, nullptr);
        if (peaks) peaks[colid] = t;
        T* D = dual + r0 * R + c;
        for (int j = 0; j < n; ++j) {
            const long long o = (long long)j * R;
            D[o] = D[o] - out[o];  // dual = (x + dual) - aux
        }
    }
}

size_t per_slot_bytes(int max_rows) { return (size_t)(5 * (size_t)max_rows + 1) * 8 + (size_t)max_rows * 4; }

}  // namespace

extern "C" {

size_t b2_unimodal_workspace_bytes(int n_groups, int R, int max_rows) {
    long long total = (long long)n_groups * R;
    const long long cap = (long long)b2_num_sms() * 2048;  // one resident wave of threads is enough
    if (total > cap) total = cap;
    total = (total + 31) / 32 * 32;
    return (size_t)total * per_slot_bytes(max_rows) + 256;
}

int b2_prox_unimodal(void* aux, void* dual, const int64_t* row_off, int n_groups, int R, int max_rows,
                     int non_negativity, int32_t* peaks, int dtype, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(R >= 1 && R <= B2_MAX_RANK, "rank %d outside [1, %d]", R, B2_MAX_RANK);
    if (n_groups == 0 || max_rows == 0) return B2_OK;
    const long long total = (long long)n_groups * R;
    long long nslots = (long long)((ws_bytes > 256 ? ws_bytes - 256 : 0) / per_slot_bytes(max_rows));
    nslots = nslots / 32 * 32;
    const long long want = (total + 31) / 32 * 32;
    if (nslots > want) nslots = want;
    B2_REQUIRE(nslots >= 32, "b2_prox_unimodal workspace too small (%zu bytes for max_rows=%d)", ws_bytes, max_rows);
    B2_REQUIRE(((uintptr_t)ws) % 8 == 0, "workspace must be 8-byte aligned");
    double* wsd = (double*)ws;
    int* wsi = (int*)(wsd + (size_t)(5 * (size_t)max_rows + 1) * nslots);
    const int grid = (int)((nslots + 127) / 128);
    B2_DISPATCH_DTYPE(dtype, {
        unimodal_kernel<T><<<grid, 128, 0, st>>>((T*)aux, (T*)dual, row_off, n_groups, R, max_rows, non_negativity,
                                                 peaks, wsd, wsi, nslots);
        B2_LAUNCH_CHECK();
    });
    return B2_OK;
}

}  // extern "C"
