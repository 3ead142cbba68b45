// fp64 instantiations of the tensor-core PARAFAC2 row pass (pf2_mma_impl.cuh) + the entry used by b2_pf2_rowpass.
#include "pf2_mma_impl.cuh"

int b2_pf2_rowpass_mma_f32(const int64_t* row_off, int n_groups, int R, const void* in_ptrs, int n_in, const void* A,
                           const void* rho, const void* Minv, const PenArgs& pa, int deferred, const void* Wmat,
                           const void* Delta, void* x, void* w_out, int ldw, void* S_out, void* BtB_out, cudaStream_t st);

// Returns B2_OK after launching, a positive error code on failure, or -1 when this formulation does not apply
// (the caller then uses the shuffle-based kernel of pf2_fused.cu).
int b2_pf2_rowpass_mma_try(const int64_t* row_off, int n_groups, int R, const void* Y, const void* A, const void* rho,
                           const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta,
                           void* x, void* w_out, int ldw, void* S_out, void* BtB_out, int dtype, cudaStream_t st) {
    const size_t es = dtype == B2_F64 ? 8 : 4;
    if (((size_t)R * es) % 16 != 0) return -1;
    if (pa.n_pen - 1 > kMaxExtra) return -1;
    if (w_out && (((size_t)ldw * es) % (2 * es) != 0)) return -1;
    RowpassInputs in;
    in.n = 0;
    in.ptr[in.n++] = Y;
    in.ptr[in.n++] = pa.dual[0];
    if (!(deferred & 1)) in.ptr[in.n++] = pa.aux[0];
    for (int p = 1; p < pa.n_pen; ++p) {
        const bool elementwise = pa.kind[p] == B2_PEN_NONNEG || pa.kind[p] == B2_PEN_BOX || pa.kind[p] == B2_PEN_L1;
        if (!((deferred & 2) && elementwise)) in.ptr[in.n++] = pa.aux[p];  // T-only state: the aux slot is not read
        in.ptr[in.n++] = pa.dual[p];
    }
    for (int a = 0; a < in.n; ++a)
        if (((uintptr_t)in.ptr[a]) % 16 != 0) return -1;
    if (dtype == B2_F64)
        return pf2_rowpass_mma_dispatch<double>(row_off, n_groups, R, in, A, rho, Minv, pa, deferred, Wmat, Delta, x,
                                                w_out, ldw, S_out, BtB_out, st);
    return b2_pf2_rowpass_mma_f32(row_off, n_groups, R, in.ptr, in.n, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out,
                                  ldw, S_out, BtB_out, st);
}
