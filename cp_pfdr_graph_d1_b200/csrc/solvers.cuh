// solvers.cuh - host-side entry points behind the C ABI (capi.cu).
#pragma once
#include "common.cuh"
#include "comm.cuh"

namespace pfdr {

// arguments of PFDR_graph_quadratic_d1_{l1,bounds}<real> (reference include/PFDR_graph_quadratic_d1_l1.hpp:36-42,
// include/PFDR_graph_quadratic_d1_bounds.hpp:34-40); prox_kind 0 = l1 (+positivity), 1 = box [lo, hi]
template <typename real>
struct QuadProblem {
    int V, E, N;
    real *X; const real *Y, *A; const int *Eu, *Ev; const real *La_d1, *La_l1;
    int positivity, prox_kind; real lo, hi;
    int Ltype; const real *L; real rho, condMin, difRcd, difTol;
    int itMax; int *it; real *Obj, *Dif; int verbose;
    const ShardInfo *shard = nullptr;   // non-NULL: this rank's part of a vertex-sharded problem (comm.cuh)
    const pfdr_b200_slab *slab = nullptr;   // non-NULL: this rank's slab of a stencil graph (slab.cuh)
};

// A solver resident on the device: layout + preconditioner built once, then reset()/run() any number
// of times (the handle API of include/pfdr_b200.h; the one-shot entry points are create+run+fetch).
struct SolverBase {
    virtual ~SolverBase() {}
    virtual void reset() = 0;                                   // back to the initial iterate (device only)
    virtual void run() = 0;                                     // iterate until the stopping rule fires
    virtual void fetch(void *X, int *it, void *Obj, void *Dif) = 0;
    // slab-sharded solves (slab.cuh): publish this rank's arena / attach the other ranks' (collective)
    virtual void slab_export(void *blob) { throw Error(PFDR_B200_ERR_ARG, "not a slab-sharded solver"); }
    virtual void slab_connect(const void *blobs, bool same_process) { throw Error(PFDR_B200_ERR_ARG, "not a slab-sharded solver"); }
};
template <typename real> SolverBase *make_quadratic(const QuadProblem<real> &q);
template <typename real> void solve_quadratic(const QuadProblem<real> &q);

// arguments of PFDR_graph_loss_d1_simplex<real> (reference include/PFDR_graph_loss_d1_simplex.hpp:24-30)
template <typename real>
struct SimplexProblem {
    int K, V, E; real al; const real *La_f; real *P; const real *Q; const int *Eu, *Ev; const real *La_d1;
    real rho, condMin, difRcd, difTol; int itMax; int *it; real *Obj, *Dif; int verbose;
    const ShardInfo *shard = nullptr;
    const pfdr_b200_slab *slab = nullptr;
};
template <typename real> SolverBase *make_simplex(const SimplexProblem<real> &q);
template <typename real> void solve_simplex(const SimplexProblem<real> &q);

// PFDR_B200_GPUS=N: the one-shot solves of large grid problems over N GPUs (multi.cu); false = not taken, solve on one GPU
template <typename real> bool solve_simplex_multi(const SimplexProblem<real> &q);
template <typename real> bool solve_quadratic_multi(const QuadProblem<real> &q);

// dense passes of the cut-pursuit's outer iteration with the operator resident on the device (cpreduce.cu)
struct CpOperatorBase {
    virtual ~CpOperatorBase() {}
    virtual void set_components(int rV, const int *Cv) = 0;
    virtual void reduce(const void *Y, int preAt, void *rA, void *rAA, void *rY) = 0;
    virtual void gradient(const void *rX, const void *YorR, void *DfS) = 0;
    virtual void residual(const void *rX, const void *Y, void *R) = 0;
};
template <typename real> CpOperatorBase *make_cp_operator(int V, int N, const real *A);

// graph passes of the cut-pursuit's outer iteration (cpgraph.cu)
void cp_components(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, int *Cv, int *rV);
template <typename real>
void cp_reduced_edges(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, const real *La_d1, const int *Cv, int rV,
                      real eps, int cap, int *rE, int *rEu, int *rEv, real *rLa_d1);
template <typename real>
void cp_merge(int V, int E, const int *Eu, const int *Ev, const int *Cv, int rV, const real *rX, real CP_difTol, real eps,
              unsigned char *active, int *count);

template <typename real> void proj_simplex_metric(real *X, const real *M, int D, int N, int nm, const real *A, int na);
template <typename real> real operator_norm_matrix(int M, int N, const real *A, real nTol, int itMax, int nbInit, int verbose);

}  // namespace pfdr
