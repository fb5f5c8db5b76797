// Drop-in replacement object for the reference's src/operator_norm_matrix.cpp (:214-219):
// symbols _Z20operator_norm_matrixI{f,d}ET_iiPKS0_S0_iii.
#include "shim_common.hpp"
template <typename real>
real operator_norm_matrix(int M, int N, const real *A, const real nTol, const int itMax, int nbInit, const int verbose);

template <>
float operator_norm_matrix<float>(int M, int N, const float *A, const float nTol, const int itMax, int nbInit,
                                  const int verbose)
{
    float r = 0;
    pfdr_b200_shim_check(pfdr_b200_operator_norm_matrix_f32(M, N, A, nTol, itMax, nbInit, verbose, &r),
                         "operator_norm_matrix<float>");
    return r;
}
template <>
double operator_norm_matrix<double>(int M, int N, const double *A, const double nTol, const int itMax, int nbInit,
                                    const int verbose)
{
    double r = 0;
    pfdr_b200_shim_check(pfdr_b200_operator_norm_matrix_f64(M, N, A, nTol, itMax, nbInit, verbose, &r),
                         "operator_norm_matrix<double>");
    return r;
}
