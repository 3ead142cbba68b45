// fp32 instantiations of the tensor-core PARAFAC2 row pass (pf2_mma_impl.cuh).
#include "pf2_mma_impl.cuh"

int b2_pf2_rowpass_mma_f32(const int64_t* row_off, int n_groups, int R, const void* in_ptrs, int n_in, const void* A,
                           const void* rho, const void* Minv, const PenArgs& pa, int deferred, const void* Wmat,
                           const void* Delta, void* x, void* w_out, int ldw, void* S_out, void* BtB_out, cudaStream_t st) {
    RowpassInputs in;
    in.n = n_in;
    for (int a = 0; a < n_in; ++a) in.ptr[a] = ((const void* const*)in_ptrs)[a];
    return pf2_rowpass_mma_dispatch<float>(row_off, n_groups, R, in, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out,
                                           ldw, S_out, BtB_out, st);
}
