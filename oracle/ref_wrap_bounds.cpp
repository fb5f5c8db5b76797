// TEST INFRASTRUCTURE ONLY. extern "C" gateway onto the unmodified reference,
// box-constraint family (separate TU: see ref_wrap_l1.cpp).
#include <stdlib.h>
#include "PFDR_graph_quadratic_d1_bounds.hpp"
#include "CP_PFDR_graph_quadratic_d1_bounds.hpp"

#define WRAP_B(SUF, real)                                                               \
extern "C" void ref_PFDR_graph_quadratic_d1_bounds_##SUF(int V, int E, int N, real *X,  \
    const real *Y, const real *A, const int *Eu, const int *Ev, const real *La_d1,      \
    real lo, real hi, int Ltype, const real *L, real rho, real condMin, real difRcd,    \
    real difTol, int itMax, int *it, real *Obj, real *Dif, int verbose)                 \
{                                                                                       \
    PFDR_graph_quadratic_d1_bounds<real>(V, E, N, X, Y, A, Eu, Ev, La_d1, lo, hi,       \
        (Lipschtype) Ltype, L, rho, condMin, difRcd, difTol, itMax, it, Obj, Dif,       \
        verbose);                                                                       \
}                                                                                       \
extern "C" void ref_CP_PFDR_graph_quadratic_d1_bounds_##SUF(int V, int E, int N,        \
    int *rV, int *Cv, real **rX, const real *Y, const real *A, const int *Eu,           \
    const int *Ev, const real *La_d1, real lo, real hi, real CP_difTol, int CP_itMax,   \
    int *CP_it, real rho, real condMin, real difRcd, real difTol, int itMax,            \
    double *Time, real *Obj, real *Dif, int verbose)                                    \
{                                                                                       \
    CP_PFDR_graph_quadratic_d1_bounds<real>(V, E, N, rV, Cv, rX, Y, A, Eu, Ev, La_d1,   \
        lo, hi, CP_difTol, CP_itMax, CP_it, rho, condMin, difRcd, difTol, itMax, Time,  \
        Obj, Dif, verbose, NULL);                                                       \
}

WRAP_B(f32, float)
WRAP_B(f64, double)
