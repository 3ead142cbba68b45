// Shared pieces of the ADMM kernels: penalty argument pack and the elementwise proximal operators.
#pragma once
#include <math.h>

#include "common.cuh"

constexpr int kMaxPen = 4;

struct PenArgs {
    int n_pen;
    int kind[kMaxPen];
    int nn[kMaxPen];
    double p0[kMaxPen], p1[kMaxPen];
    void* aux[kMaxPen];
    void* dual[kMaxPen];
};

template <typename T>
__device__ __forceinline__ T prox_elem(T v, int kind, int nn, T p0, T p1, T rho) {
    // penalties.py:503-508 (NonNegativity), :537-542 (Box), :573-586 (L1)
    if (kind == B2_PEN_NONNEG) return v > T(0) ? v : T(0);
    if (kind == B2_PEN_BOX) return v < p0 ? p0 : (v > p1 ? p1 : v);
    // L1
    const T thr = p0 / rho;
    if (nn) {
        const T u = v - thr;
        return u > T(0) ? u : T(0);
    }
    const T a = fabs(v) - thr;
    const T m = a > T(0) ? a : T(0);
    const T sgn = v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0));
    return sgn * m;
}


// host: validate and pack the caller's descriptors
inline int b2_pack_penalties(const b2_penalty_desc* pens, int n_pen, PenArgs* pa) {
    B2_REQUIRE(n_pen >= 0 && n_pen <= kMaxPen, "at most %d penalties per mode are supported (got %d)", kMaxPen, n_pen);
    pa->n_pen = n_pen;
    for (int p = 0; p < n_pen; ++p) {
        B2_REQUIRE(pens[p].kind >= B2_PEN_NONNEG && pens[p].kind <= B2_PEN_PARAFAC2, "unknown penalty kind %d",
                   pens[p].kind);
        B2_REQUIRE(pens[p].aux && pens[p].dual, "penalty %d: aux/dual pointers must be set", p);
        pa->kind[p] = pens[p].kind;
        pa->nn[p] = pens[p].non_negativity;
        pa->p0[p] = pens[p].p0;
        pa->p1[p] = pens[p].p1;
        pa->aux[p] = pens[p].aux;
        pa->dual[p] = pens[p].dual;
    }
    return B2_OK;
}
