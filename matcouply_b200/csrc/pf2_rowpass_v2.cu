// fp64 instantiations of the second-generation PARAFAC2 row pass (pf2_rowpass_v2.cuh) + the entry b2_pf2_rowpass tries
// first (B2_OPT_PF2_ROWPASS_MMA >= 2).
#include "pf2_rowpass_v2.cuh"

int b2_pf2_rowpass_v2_try(const int64_t* row_off, int n_groups, int R, const void* Y, const void* A, const void* rho,
                          const void* Minv, const PenArgs& pa, int deferred, const void* Wmat, const void* Delta,
                          void* x, void* w_out, int ldw, void* S_out, void* BtB_out, double* stats_part, int dtype,
                          cudaStream_t st) {
    if (dtype != B2_F64) return -1;
    return rp2::try_launch(row_off, n_groups, R, Y, A, rho, Minv, pa, deferred, Wmat, Delta, x, w_out, ldw, S_out,
                           BtB_out, stats_part, st);
}

// Whether a call with these properties is served by the steady-state kernel (the only one that can emit the
// companion's gap terms and keep the companion as one array on the last pass).
int b2_pf2_rowpass_v2_applies(int R, int dtype, int n_pen, int companion_kind, int deferred) {
    if (dtype != B2_F64 || !(deferred & 1) || R % 4 != 0 || R < 4 || R > 32) return 0;
    if (n_pen < 1 || n_pen > 2) return 0;
    if (n_pen == 2 && companion_kind != B2_PEN_NONNEG) return 0;
    return b2_option_value(B2_OPT_PF2_ROWPASS_MMA) >= 2 ? 1 : 0;
}
