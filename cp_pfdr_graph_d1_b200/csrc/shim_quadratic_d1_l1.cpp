// Drop-in replacement object for the reference's src/PFDR_graph_quadratic_d1_l1.cpp: defines the same
// explicit instantiations (reference :555-566, symbols
// _Z26PFDR_graph_quadratic_d1_l1I{f,d}EviiiPT_PKS0_S3_PKiS5_S3_S3_i10LipschtypeS3_S0_S0_S0_S0_iPiS1_S1_i)
// and forwards them to the C ABI of libpfdr_b200.so.  Compile it INSTEAD of that file; every caller
// (MEX gateways, Boost-Python module, CP_PFDR_graph_quadratic_d1_l1*.cpp) links unchanged.
// -DPFDR_B200_REFERENCE_INCLUDE=<dir> makes it include the reference's own header instead of re-declaring.
#include "shim_common.hpp"
#ifdef PFDR_B200_REFERENCE_INCLUDE
#include "PFDR_graph_quadratic_d1_l1.hpp"
#else
typedef enum {SCAL, DIAG} Lipschtype;      // include/PFDR_graph_quadratic_d1_l1.hpp:34 (part of the mangled name)
template <typename real>
void PFDR_graph_quadratic_d1_l1(const int V, const int E, const int N, real *X, const real *Y, const real *A,
    const int *Eu, const int *Ev, const real *La_d1, const real *La_l1, const int positivity,
    const Lipschtype Ltype, const real *L, const real rho, const real condMin, real difRcd, const real difTol,
    const int itMax, int *it, real *Obj, real *Dif, const int verbose);
#endif

template <>
void PFDR_graph_quadratic_d1_l1<float>(const int V, const int E, const int N, float *X, const float *Y,
    const float *A, const int *Eu, const int *Ev, const float *La_d1, const float *La_l1, const int positivity,
    const Lipschtype Ltype, const float *L, const float rho, const float condMin, float difRcd, const float difTol,
    const int itMax, int *it, float *Obj, float *Dif, const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_quadratic_d1_l1_f32(V, E, N, X, Y, A, Eu, Ev, La_d1, La_l1, positivity, (int) Ltype,
        L, rho, condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_quadratic_d1_l1<float>");
}

template <>
void PFDR_graph_quadratic_d1_l1<double>(const int V, const int E, const int N, double *X, const double *Y,
    const double *A, const int *Eu, const int *Ev, const double *La_d1, const double *La_l1, const int positivity,
    const Lipschtype Ltype, const double *L, const double rho, const double condMin, double difRcd,
    const double difTol, const int itMax, int *it, double *Obj, double *Dif, const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_quadratic_d1_l1_f64(V, E, N, X, Y, A, Eu, Ev, La_d1, La_l1, positivity, (int) Ltype,
        L, rho, condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_quadratic_d1_l1<double>");
}
