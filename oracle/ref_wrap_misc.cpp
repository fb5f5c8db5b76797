// TEST INFRASTRUCTURE ONLY. extern "C" gateway onto the unmodified reference:
// simplex family, proj_simplex_metric, operator_norm_matrix, plus a pinned time()
// so the power method's rand_r seeds (operator_norm_matrix.cpp:182) are reproducible.
#include <stdlib.h>
#include <time.h>
#include <omp.h>
#include "PFDR_graph_loss_d1_simplex.hpp"
#include "CP_PFDR_graph_loss_d1_simplex.hpp"
#include "proj_simplex.hpp"
#include "operator_norm_matrix.hpp"

// bound inside libref.so through -Wl,-Bsymbolic-functions
extern "C" time_t time(time_t *t) { if (t) *t = 42; return 42; }
extern "C" void ref_free(void *p) { free(p); }
extern "C" int ref_num_procs(void) { return omp_get_num_procs(); }

#define WRAP_M(SUF, real)                                                               \
extern "C" void ref_PFDR_graph_loss_d1_simplex_##SUF(int K, int V, int E, real al,      \
    const real *La_f, real *P, const real *Q, const int *Eu, const int *Ev,             \
    const real *La_d1, real rho, real condMin, real difRcd, real difTol, int itMax,     \
    int *it, real *Obj, real *Dif, int verbose)                                         \
{                                                                                       \
    PFDR_graph_loss_d1_simplex<real>(K, V, E, al, La_f, P, Q, Eu, Ev, La_d1, rho,       \
        condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose);                         \
}                                                                                       \
extern "C" void ref_CP_PFDR_graph_loss_d1_simplex_##SUF(int K, int V, int E, real al,   \
    int *rV, int *Cv, real **rP, const real *Q, const int *Eu, const int *Ev,           \
    const real *La_d1, real CP_difTol, int CP_itMax, int *CP_it, real rho,              \
    real condMin, real difRcd, real difTol, int itMax, double *Time, real *Obj,         \
    real *Dif, int verbose)                                                             \
{                                                                                       \
    CP_PFDR_graph_loss_d1_simplex<real>(K, V, E, al, rV, Cv, rP, Q, Eu, Ev, La_d1,      \
        CP_difTol, CP_itMax, CP_it, rho, condMin, difRcd, difTol, itMax, Time, Obj,     \
        Dif, verbose, NULL);                                                            \
}                                                                                       \
extern "C" void ref_proj_simplex_metric_##SUF(real *X, const real *M, int D, int N,     \
    int nm, const real *A, int na)                                                      \
{                                                                                       \
    proj_simplex_metric<real>(X, M, D, N, nm, A, na);                                   \
}                                                                                       \
extern "C" real ref_operator_norm_matrix_##SUF(int M, int N, const real *A, real nTol,  \
    int itMax, int nbInit, int verbose)                                                 \
{                                                                                       \
    return operator_norm_matrix<real>(M, N, A, nTol, itMax, nbInit, verbose);           \
}

WRAP_M(f32, float)
WRAP_M(f64, double)
