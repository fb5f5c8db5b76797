// cp_gateway.cpp - extern "C" gateway onto the reference's cut-pursuit, for the Python module libCP.py (the Boost.Python
// module of the reference, python/CP_quadratic_l1_py.cpp:107-160, needs Boost; this one needs ctypes only).
//
// NOT part of libpfdr_b200.so: a maintainer compiles this file together with the reference's UNMODIFIED cut-pursuit sources
// (src/CP_PFDR_graph_quadratic_d1_l1.cpp, src/graph.cpp, src/maxflow.cpp) and links the result against
// libpfdr_b200_shim.so, which serves PFDR_graph_quadratic_d1_l1<real> and operator_norm_matrix<real> from the GPU
// (recipe: INTEGRATION.md, "Python module").  The gateway only fixes the template type and releases what the reference
// allocates; nothing is computed here.
#include <stdlib.h>
#include "CP_PFDR_graph_quadratic_d1_l1.hpp"      // the reference's include/ (via -I)

#define CP_GATEWAY(SUF, real)                                                                                     \
extern "C" int cp_b200_quadratic_d1_l1_##SUF(int V, int E, int N, const real *Y, const real *A, const int *Eu,    \
    const int *Ev, const real *La_d1, const real *La_l1, int positivity, real CP_difTol, int CP_itMax,            \
    real PFDR_rho, real PFDR_condMin, real PFDR_difRcd, real PFDR_difTol, int PFDR_itMax, int verbose,            \
    int *Cv /* V */, real *rX /* room for V values */, int *CP_it)                                                \
{                                                                                                                 \
    int rV = 0;                                                                                                   \
    real *values = NULL;                                                                                          \
    CP_PFDR_graph_quadratic_d1_l1<real>(V, E, N, &rV, Cv, &values, Y, A, Eu, Ev, La_d1, La_l1, positivity,        \
        CP_difTol, CP_itMax, CP_it, PFDR_rho, PFDR_condMin, PFDR_difRcd, PFDR_difTol, PFDR_itMax,                 \
        (double *) NULL, (real *) NULL, (real *) NULL, verbose, (struct CPql1_Restart<real> *) NULL);             \
    for (int c = 0; c < rV; c++) rX[c] = values[c];                                                               \
    free(values);                                                                                                 \
    return rV;                                                                                                    \
}

CP_GATEWAY(f32, float)
CP_GATEWAY(f64, double)
