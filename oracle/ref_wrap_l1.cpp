// TEST INFRASTRUCTURE ONLY. extern "C" gateway onto the *unmodified* reference
// (compiled where it lies under /root/reference by oracle/build_ref.sh; outputs
// go to oracle/_ref/ only). This TU includes the l1 header; the bounds header
// re-defines `Lipschtype` and therefore lives in ref_wrap_bounds.cpp.
#include <stdlib.h>
#include "PFDR_graph_quadratic_d1_l1.hpp"      // /root/reference/include (via -I)
#include "CP_PFDR_graph_quadratic_d1_l1.hpp"

#define WRAP_L1(SUF, real)                                                              \
extern "C" void ref_PFDR_graph_quadratic_d1_l1_##SUF(int V, int E, int N, real *X,      \
    const real *Y, const real *A, const int *Eu, const int *Ev, const real *La_d1,      \
    const real *La_l1, int positivity, int Ltype, const real *L, real rho,              \
    real condMin, real difRcd, real difTol, int itMax, int *it, real *Obj, real *Dif,   \
    int verbose)                                                                        \
{                                                                                       \
    PFDR_graph_quadratic_d1_l1<real>(V, E, N, X, Y, A, Eu, Ev, La_d1, La_l1,            \
        positivity, (Lipschtype) Ltype, L, rho, condMin, difRcd, difTol, itMax, it,     \
        Obj, Dif, verbose);                                                             \
}                                                                                       \
extern "C" void ref_CP_PFDR_graph_quadratic_d1_l1_##SUF(int V, int E, int N, int *rV,   \
    int *Cv, real **rX, const real *Y, const real *A, const int *Eu, const int *Ev,     \
    const real *La_d1, const real *La_l1, int positivity, real CP_difTol, int CP_itMax, \
    int *CP_it, real rho, real condMin, real difRcd, real difTol, int itMax,            \
    double *Time, real *Obj, real *Dif, int verbose, int duplex)                        \
{                                                                                       \
    if (duplex)                                                                         \
        CP_PFDR_graph_quadratic_d1_l1_duplex<real>(V, E, N, rV, Cv, rX, Y, A, Eu, Ev,   \
            La_d1, La_l1, positivity, CP_difTol, CP_itMax, CP_it, rho, condMin, difRcd, \
            difTol, itMax, Time, Obj, Dif, verbose, NULL);                              \
    else                                                                                \
        CP_PFDR_graph_quadratic_d1_l1<real>(V, E, N, rV, Cv, rX, Y, A, Eu, Ev,          \
            La_d1, La_l1, positivity, CP_difTol, CP_itMax, CP_it, rho, condMin, difRcd, \
            difTol, itMax, Time, Obj, Dif, verbose, NULL);                              \
}

WRAP_L1(f32, float)
WRAP_L1(f64, double)
