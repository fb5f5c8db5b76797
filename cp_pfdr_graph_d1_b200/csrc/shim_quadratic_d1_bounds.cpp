// Drop-in replacement object for the reference's src/PFDR_graph_quadratic_d1_bounds.cpp (instantiations :532-543,
// symbols _Z30PFDR_graph_quadratic_d1_boundsI{f,d}EviiiPT_PKS0_S3_PKiS5_S3_S0_S0_10LipschtypeS3_S0_S0_S0_S0_iPiS1_S1_i).
// Separate translation unit from the l1 shim for the same reason as in the reference: both headers define
// `Lipschtype`.
#include "shim_common.hpp"
#ifdef PFDR_B200_REFERENCE_INCLUDE
#include "PFDR_graph_quadratic_d1_bounds.hpp"
#else
typedef enum {SCAL, DIAG} Lipschtype;      // include/PFDR_graph_quadratic_d1_bounds.hpp:32
template <typename real>
void PFDR_graph_quadratic_d1_bounds(const int V, const int E, const int N, real *X, const real *Y, const real *A,
    const int *Eu, const int *Ev, const real *La_d1, const real min, const real max, const Lipschtype Ltype,
    const real *L, const real rho, const real condMin, real difRcd, const real difTol, const int itMax, int *it,
    real *Obj, real *Dif, const int verbose);
#endif

template <>
void PFDR_graph_quadratic_d1_bounds<float>(const int V, const int E, const int N, float *X, const float *Y,
    const float *A, const int *Eu, const int *Ev, const float *La_d1, const float min, const float max,
    const Lipschtype Ltype, const float *L, const float rho, const float condMin, float difRcd, const float difTol,
    const int itMax, int *it, float *Obj, float *Dif, const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_quadratic_d1_bounds_f32(V, E, N, X, Y, A, Eu, Ev, La_d1, min, max, (int) Ltype, L, rho,
        condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_quadratic_d1_bounds<float>");
}

template <>
void PFDR_graph_quadratic_d1_bounds<double>(const int V, const int E, const int N, double *X, const double *Y,
    const double *A, const int *Eu, const int *Ev, const double *La_d1, const double min, const double max,
    const Lipschtype Ltype, const double *L, const double rho, const double condMin, double difRcd,
    const double difTol, const int itMax, int *it, double *Obj, double *Dif, const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_quadratic_d1_bounds_f64(V, E, N, X, Y, A, Eu, Ev, La_d1, min, max, (int) Ltype, L, rho,
        condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_quadratic_d1_bounds<double>");
}
