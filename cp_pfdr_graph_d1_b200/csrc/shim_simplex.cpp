// Drop-in replacement object for the reference's src/PFDR_graph_loss_d1_simplex.cpp (:717-726) and
// src/proj_simplex_metric.cpp (:85-89): symbols
// _Z26PFDR_graph_loss_d1_simplexI{f,d}EviiiT_PKS0_PS0_S2_PKiS5_S2_S0_S0_S0_S0_iPiS3_S3_i and
// _Z19proj_simplex_metricI{f,d}EvPT_PKS0_iiiS3_i.
#include "shim_common.hpp"
template <typename real>
void PFDR_graph_loss_d1_simplex(const int K, const int V, const int E, const real al, const real *La_f, real *P,
    const real *Q, const int *Eu, const int *Ev, const real *La_d1, const real rho, const real condMin, real difRcd,
    const real difTol, const int itMax, int *it, real *Obj, real *Dif, const int verbose);
template <typename real>
void proj_simplex_metric(real *X, const real *M, const int D, const int N, const int nm, const real *A, const int na);

template <>
void PFDR_graph_loss_d1_simplex<float>(const int K, const int V, const int E, const float al, const float *La_f,
    float *P, const float *Q, const int *Eu, const int *Ev, const float *La_d1, const float rho, const float condMin,
    float difRcd, const float difTol, const int itMax, int *it, float *Obj, float *Dif, const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_loss_d1_simplex_f32(K, V, E, al, La_f, P, Q, Eu, Ev, La_d1, rho, condMin, difRcd,
        difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_loss_d1_simplex<float>");
}
template <>
void PFDR_graph_loss_d1_simplex<double>(const int K, const int V, const int E, const double al, const double *La_f,
    double *P, const double *Q, const int *Eu, const int *Ev, const double *La_d1, const double rho,
    const double condMin, double difRcd, const double difTol, const int itMax, int *it, double *Obj, double *Dif,
    const int verbose)
{
    pfdr_b200_shim_check(pfdr_b200_loss_d1_simplex_f64(K, V, E, al, La_f, P, Q, Eu, Ev, La_d1, rho, condMin, difRcd,
        difTol, itMax, it, Obj, Dif, verbose), "PFDR_graph_loss_d1_simplex<double>");
}
template <>
void proj_simplex_metric<float>(float *X, const float *M, const int D, const int N, const int nm, const float *A,
                                const int na)
{
    pfdr_b200_shim_check(pfdr_b200_proj_simplex_metric_f32(X, M, D, N, nm, A, na), "proj_simplex_metric<float>");
}
template <>
void proj_simplex_metric<double>(double *X, const double *M, const int D, const int N, const int nm, const double *A,
                                 const int na)
{
    pfdr_b200_shim_check(pfdr_b200_proj_simplex_metric_f64(X, M, D, N, nm, A, na), "proj_simplex_metric<double>");
}
