/* pfdr_b200.h - C ABI of the B200-native PFDR hot path (libpfdr_b200.so).
 *
 * Drop-in boundary for ai3DVision/CP_PFDR_graph_d1: one extern "C" entry point per reference
 * routine x {f32,f64}.  Argument lists, meanings and ownership are exactly those of the reference
 * templates they replace (reference file:line cited at each prototype); `Lipschtype` travels as
 * an int (SCAL = 0, DIAG = 1; include/PFDR_graph_quadratic_d1_l1.hpp:34) and every function
 * returns an int status (0 = ok) because the reference's `void` cannot report a CUDA failure.
 * csrc/shim_*.cpp re-creates the reference's C++ template symbols on top of this ABI so that
 * the MEX / Boost-Python gateways and the cut-pursuit sources link unchanged.
 *
 * Pointers are plain host pointers, as in the reference.  As an extension every array argument
 * may instead be a CUDA device pointer (detected with cudaPointerGetAttributes): it is then used
 * in place / written in place without a host round trip.
 *
 * There is no CPU fallback: every entry point needs a CUDA device and fails with a non-zero
 * status (message in pfdr_b200_last_error()) when none is usable.
 */
#ifndef PFDR_B200_H
#define PFDR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PFDR_B200_SCAL 0
#define PFDR_B200_DIAG 1

/* status codes */
#define PFDR_B200_OK            0
#define PFDR_B200_ERR_CUDA      1   /* CUDA runtime failure (no device, OOM, launch error)      */
#define PFDR_B200_ERR_ARG       2   /* unsupported / inconsistent argument                       */
#define PFDR_B200_ERR_NCCL      3   /* NCCL failure (sharded entry points only)                  */

/* ---- quadratic + d1 + l1 (+ positivity) ----------------------------------------------------
 * replaces PFDR_graph_quadratic_d1_l1<real>  (include/PFDR_graph_quadratic_d1_l1.hpp:36-42,
 * src/PFDR_graph_quadratic_d1_l1.cpp:270-553).  N>0: A is N-by-V column major, Y length N;
 * N<0: A is (A^tA) V-by-V, Y is A^tY; N==0: A is diag(A^tA) or NULL (identity).
 * X[V] in/out (must be initialised); *it always written; Obj needs itMax+1 entries, Dif itMax. */
int pfdr_b200_quadratic_d1_l1_f32(int V, int E, int N, float *X, const float *Y, const float *A,
    const int *Eu, const int *Ev, const float *La_d1, const float *La_l1, int positivity,
    int Ltype, const float *L, float rho, float condMin, float difRcd, float difTol, int itMax,
    int *it, float *Obj, float *Dif, int verbose);
int pfdr_b200_quadratic_d1_l1_f64(int V, int E, int N, double *X, const double *Y, const double *A,
    const int *Eu, const int *Ev, const double *La_d1, const double *La_l1, int positivity,
    int Ltype, const double *L, double rho, double condMin, double difRcd, double difTol, int itMax,
    int *it, double *Obj, double *Dif, int verbose);

/* ---- quadratic + d1 + box ------------------------------------------------------------------
 * replaces PFDR_graph_quadratic_d1_bounds<real> (include/PFDR_graph_quadratic_d1_bounds.hpp:34-40,
 * src/PFDR_graph_quadratic_d1_bounds.cpp:244-530).  min/max = -/+HUGE_VAL(F) for "no bound". */
int pfdr_b200_quadratic_d1_bounds_f32(int V, int E, int N, float *X, const float *Y, const float *A,
    const int *Eu, const int *Ev, const float *La_d1, float min, float max,
    int Ltype, const float *L, float rho, float condMin, float difRcd, float difTol, int itMax,
    int *it, float *Obj, float *Dif, int verbose);
int pfdr_b200_quadratic_d1_bounds_f64(int V, int E, int N, double *X, const double *Y, const double *A,
    const int *Eu, const int *Ev, const double *La_d1, double min, double max,
    int Ltype, const double *L, double rho, double condMin, double difRcd, double difTol, int itMax,
    int *it, double *Obj, double *Dif, int verbose);

/* ---- loss + d1 + simplex -------------------------------------------------------------------
 * replaces PFDR_graph_loss_d1_simplex<real> (include/PFDR_graph_loss_d1_simplex.hpp:24-30,
 * src/PFDR_graph_loss_d1_simplex.cpp:372-715).  P, Q are K-by-V column major (P[v*K+k]);
 * al = 0 linear, 0<al<1 smoothed Kullback-Leibler, al = 1 quadratic (al > 1 is undefined in the
 * reference, SURVEY 5.1-10, and rejected here with PFDR_B200_ERR_ARG). */
int pfdr_b200_loss_d1_simplex_f32(int K, int V, int E, float al, const float *La_f, float *P,
    const float *Q, const int *Eu, const int *Ev, const float *La_d1, float rho, float condMin,
    float difRcd, float difTol, int itMax, int *it, float *Obj, float *Dif, int verbose);
int pfdr_b200_loss_d1_simplex_f64(int K, int V, int E, double al, const double *La_f, double *P,
    const double *Q, const int *Eu, const int *Ev, const double *La_d1, double rho, double condMin,
    double difRcd, double difTol, int itMax, int *it, double *Obj, double *Dif, int verbose);

/* ---- proj_simplex_metric -------------------------------------------------------------------
 * replaces proj_simplex_metric<real> (include/proj_simplex.hpp:33-35, src/proj_simplex_metric.cpp:18-83):
 * in-place projection of the N columns (length D) of X onto {x>=0, sum x = A[min(n,na-1)]} in the
 * metric diag(1/M[:,min(n,nm-1)]). */
int pfdr_b200_proj_simplex_metric_f32(float *X, const float *M, int D, int N, int nm, const float *A, int na);
int pfdr_b200_proj_simplex_metric_f64(double *X, const double *M, int D, int N, int nm, const double *A, int na);

/* ---- operator_norm_matrix ------------------------------------------------------------------
 * replaces operator_norm_matrix<real> (include/operator_norm_matrix.hpp:12-14,
 * src/operator_norm_matrix.cpp:84-212): squared operator norm (largest eigenvalue of A^tA) by the
 * power method, max over nbInit random starts; M or N == 0 means A already is A^tA / AA^t.
 * The value is returned through *result. */
int pfdr_b200_operator_norm_matrix_f32(int M, int N, const float *A, float nTol, int itMax, int nbInit,
                                       int verbose, float *result);
int pfdr_b200_operator_norm_matrix_f64(int M, int N, const double *A, double nTol, int itMax, int nbInit,
                                       int verbose, double *result);

/* ---- resident solver handles (no counterpart in the reference) --------------------------------
 * The one-shot entry points above are create + run + fetch + destroy.  A handle keeps the inputs,
 * the half-edge layout and all work arrays in HBM so that a caller (a benchmark, or a cut-pursuit
 * loop re-solving on one graph) can reset()/run() repeatedly without touching the host.
 * create: same arguments as the one-shot call; prox_kind 0 = l1 (+positivity), 1 = box [lo, hi];
 * want_obj / want_dif select whether the per-iteration records are kept.  fetch: any pointer may be
 * NULL; Obj needs it+1 entries, Dif it. */
typedef struct pfdr_b200_solver pfdr_b200_solver;
int pfdr_b200_solver_create_quadratic_f32(pfdr_b200_solver **out, int prox_kind, int V, int E, int N,
    const float *X0, const float *Y, const float *A, const int *Eu, const int *Ev, const float *La_d1,
    const float *La_l1, int positivity, float lo, float hi, int Ltype, const float *L, float rho,
    float condMin, float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_quadratic_f64(pfdr_b200_solver **out, int prox_kind, int V, int E, int N,
    const double *X0, const double *Y, const double *A, const int *Eu, const int *Ev, const double *La_d1,
    const double *La_l1, int positivity, double lo, double hi, int Ltype, const double *L, double rho,
    double condMin, double difRcd, double difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_simplex_f32(pfdr_b200_solver **out, int K, int V, int E, float al,
    const float *La_f, const float *P0, const float *Q, const int *Eu, const int *Ev, const float *La_d1,
    float rho, float condMin, float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_simplex_f64(pfdr_b200_solver **out, int K, int V, int E, double al,
    const double *La_f, const double *P0, const double *Q, const int *Eu, const int *Ev, const double *La_d1,
    double rho, double condMin, double difRcd, double difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_reset(pfdr_b200_solver *s);   /* back to the initial iterate + first preconditioning */
int pfdr_b200_solver_run(pfdr_b200_solver *s);     /* iterate until the stopping rule fires (<= itMax)    */
int pfdr_b200_solver_fetch(pfdr_b200_solver *s, void *X, int *it, void *Obj, void *Dif);
int pfdr_b200_solver_destroy(pfdr_b200_solver *s);

/* ---- cut-pursuit outer iteration: the dense passes around the PFDR call --------
 * The reference's cut-pursuit (src/CP_PFDR_graph_quadratic_d1_l1.cpp, ..._bounds.cpp) runs, once per outer
 * iteration and on the host, three passes over the full operator: the gradient of the quadratic term that feeds
 * the steepest cut (:340-374), the reduced operator of the new components (:664-772) and, for N > 0, the residual
 * of the reduced solution (:889-903).  A pfdr_b200_cp_operator keeps A (N > 0: N x V column major; N = -V: the
 * V x V matrix A^tA; N = 0: the diagonal of A^tA, NULL = identity) in HBM across outer iterations; the calls below
 * replace those blocks one for one.  `real` is the type of the create call; host or device pointers.
 *   set_components  rV, Cv[V] as the reference holds them after its connected-components pass (:574-600)
 *   reduce          N > 0: rA (N x rV), and when preAt: rAA = rA^t rA (rV x rV, exactly symmetric), rY = rA^t Y (Y: N values)
 *                   N < 0: rAA (rV x rV), rY (Y: the V values of A^tY);  N = 0: rAA (rV: summed diagonal or component
 *                   sizes), rY.   rA may be NULL (kept on the device for residual()); rAA / rY may be NULL when N > 0 && !preAt
 *   gradient        DfS[V] before the d1 contribution; YorR: N > 0: the residual R (NULL = the one kept by the last
 *                   residual()), else the observation A^tY (V values)
 *   residual        N > 0 only: R = Y - rA rX for the components and rA of the last reduce()
 * Component sums are taken in ascending vertex order, in double: results agree with the reference's to rounding
 * (tests: 1e-10 relative in float64, 1e-3 in float32). */
typedef struct pfdr_b200_cp_operator pfdr_b200_cp_operator;
int pfdr_b200_cp_operator_create_f32(pfdr_b200_cp_operator **out, int V, int N, const float *A);
int pfdr_b200_cp_operator_create_f64(pfdr_b200_cp_operator **out, int V, int N, const double *A);
int pfdr_b200_cp_set_components(pfdr_b200_cp_operator *op, int rV, const int *Cv);
int pfdr_b200_cp_reduce(pfdr_b200_cp_operator *op, const void *Y, int preAt, void *rA, void *rAA, void *rY);
int pfdr_b200_cp_gradient(pfdr_b200_cp_operator *op, const void *rX, const void *YorR, void *DfS);
int pfdr_b200_cp_residual(pfdr_b200_cp_operator *op, const void *rX, const void *Y, void *R);
int pfdr_b200_cp_operator_destroy(pfdr_b200_cp_operator *op);

/* ---- cut-pursuit outer iteration: the graph passes ---------------------------------
 * `active[e]` = 1 when edge e is cut (the reference keeps the flag in its graph structure, arcs[2e].is_active).
 *   components     Cv[V], *rV: connected components of the NON-active edges, numbered by smallest vertex - the numbering
 *                  of the reference's sweep (src/CP_PFDR_graph_quadratic_d1_l1.cpp:570-600), bit-identical
 *   reduced_edges  reduced connectivity of :603-661: one edge (rEu <= rEv) per pair of components joined by active edges
 *                  of nonzero weight, weights added (an active edge inside one component gives the loop (c, c) and counts
 *                  twice, as in the reference), the loop (c, c) of weight eps on components no such edge touches.  Sorted by
 *                  (rEu, rEv) - the reference emits first-encounter order.  `capacity` = room in the three output arrays
 *                  (E + rV always suffices); *rE is set even when the call fails for lack of room
 *   merge          :863-886: clears active[e] where |a - b| / max(|a|, |b|, eps) <= CP_difTol, a, b the values rX of the two
 *                  components; *deactivated (may be NULL) = how many
 * Host or device pointers. */
int pfdr_b200_cp_components(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, int *Cv, int *rV);
int pfdr_b200_cp_reduced_edges_f32(int V, int E, const int *Eu, const int *Ev, const unsigned char *active,
    const float *La_d1, const int *Cv, int rV, float eps, int capacity, int *rE, int *rEu, int *rEv, float *rLa_d1);
int pfdr_b200_cp_reduced_edges_f64(int V, int E, const int *Eu, const int *Ev, const unsigned char *active,
    const double *La_d1, const int *Cv, int rV, double eps, int capacity, int *rE, int *rEu, int *rEv, double *rLa_d1);
int pfdr_b200_cp_merge_f32(int V, int E, const int *Eu, const int *Ev, const int *Cv, int rV, const float *rX,
    float CP_difTol, float eps, unsigned char *active, int *deactivated);
int pfdr_b200_cp_merge_f64(int V, int E, const int *Eu, const int *Ev, const int *Cv, int rV, const double *rX,
    double CP_difTol, double eps, unsigned char *active, int *deactivated);

/* ---- vertex-sharded solves over NCCL (no counterpart in the reference) ---------------------
 * One process per GPU.  Rank r owns a contiguous vertex range plus every edge touching it; vertices at the
 * far end of a cut edge are ghosts numbered after the owned ones (cp_pfdr_graph_d1_b200/shard.py builds the
 * plan).  Per iteration a rank exchanges, with each neighbour, the forward point of its boundary vertices and
 * its half-edge copies on the cut edges (ncclSend/ncclRecv), and all ranks all-reduce the one or two scalars
 * of the stopping rule.  Index arrays are HOST arrays in LOCAL numbering; *_off have nnb+1 entries.
 * The problem arrays passed to the *_sharded constructors are the LOCAL ones (Vown + ghosts vertices, local
 * edges); fetch() returns the Vown owned entries.  This build shards N == 0 problems without Obj/reconditioning. */
typedef struct {
    int Vown;                 /* vertices owned by this rank (local ids 0..Vown-1; ghosts follow)            */
    int Vglobal;              /* vertices of the whole problem                                                */
    int nnb;                  /* neighbouring ranks                                                           */
    const int *nb_rank;       /* [nnb]                                                                        */
    const int *sv_off, *sv;   /* vertices whose forward point neighbour i needs (owned ids)                   */
    const int *sh_off, *sh;   /* our half-edges 2*e_local+side on the edges cut with neighbour i              */
    const int *rv_off, *rv;   /* ghost ids filled by neighbour i's message                                    */
    const int *rh_off, *rh;   /* their half-edges on the same cut edges, same order                           */
} pfdr_b200_shard;
int pfdr_b200_comm_unique_id(void *out128);                           /* rank 0: ncclGetUniqueId (128 bytes)  */
int pfdr_b200_comm_init(const void *id128, int world, int rank);      /* every rank, after pfdr_b200_set_device */
int pfdr_b200_comm_destroy(void);
int pfdr_b200_solver_create_quadratic_sharded_f32(pfdr_b200_solver **out, const pfdr_b200_shard *shard,
    int prox_kind, int V, int E, int N, const float *X0, const float *Y, const float *A, const int *Eu,
    const int *Ev, const float *La_d1, const float *La_l1, int positivity, float lo, float hi, int Ltype,
    const float *L, float rho, float condMin, float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_quadratic_sharded_f64(pfdr_b200_solver **out, const pfdr_b200_shard *shard,
    int prox_kind, int V, int E, int N, const double *X0, const double *Y, const double *A, const int *Eu,
    const int *Ev, const double *La_d1, const double *La_l1, int positivity, double lo, double hi, int Ltype,
    const double *L, double rho, double condMin, double difRcd, double difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_simplex_sharded_f32(pfdr_b200_solver **out, const pfdr_b200_shard *shard, int K, int V,
    int E, float al, const float *La_f, const float *P0, const float *Q, const int *Eu, const int *Ev,
    const float *La_d1, float rho, float condMin, float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_simplex_sharded_f64(pfdr_b200_solver **out, const pfdr_b200_shard *shard, int K, int V,
    int E, double al, const double *La_f, const double *P0, const double *Q, const int *Eu, const int *Ev,
    const double *La_d1, double rho, double condMin, double difRcd, double difTol, int itMax, int want_obj, int want_dif);

/* ---- slab-sharded solves of grid (constant-offset) graphs: peer-memory halo exchange, no collective library ----
 * Rank r of `world` owns the contiguous global vertex range [v0, v1) and is given every vertex array over
 * [lo, hi) = [v0 - band, v1 + band) clipped to [0, Vglobal) (band = largest edge offset, n^2 on an n^3 grid; the values
 * of the ghost rows are not read), the edges with at least one owned end in GLOBAL order with vertex numbers relative
 * to lo, and their weights.  Create the solver on every rank, exchange the PFDR_B200_SLAB_BLOB-byte blobs of
 * pfdr_b200_slab_export (any transport: torch.distributed, MPI, a file ...), hand the table of all ranks' blobs to
 * pfdr_b200_slab_connect, then reset / run / fetch collectively.  Ranks live in separate processes (CUDA IPC) or in
 * one process (same_process = 1, peer access); each on its OWN GPU - their kernels wait on one another.
 * fetch returns the owned rows.  Replaces nothing in the reference (which has no distributed mode). */
#define PFDR_B200_SLAB_BLOB 256
typedef struct { int world, rank, Vglobal, v0, v1, lo, hi; } pfdr_b200_slab;
int pfdr_b200_solver_create_simplex_slab_f32(pfdr_b200_solver **out, const pfdr_b200_slab *slab, int K, int Vloc, int Eloc,
    float al, const float *La_f, const float *P0, const float *Q, const int *Eu, const int *Ev, const float *La_d1,
    float rho, float condMin, float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_simplex_slab_f64(pfdr_b200_solver **out, const pfdr_b200_slab *slab, int K, int Vloc, int Eloc,
    double al, const double *La_f, const double *P0, const double *Q, const int *Eu, const int *Ev, const double *La_d1,
    double rho, double condMin, double difRcd, double difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_quadratic_slab_f32(pfdr_b200_solver **out, const pfdr_b200_slab *slab, int prox_kind, int Vloc,
    int Eloc, int N, const float *X0, const float *Y, const float *A, const int *Eu, const int *Ev, const float *La_d1,
    const float *La_l1, int positivity, float lo, float hi, int Ltype, const float *L, float rho, float condMin,
    float difRcd, float difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_solver_create_quadratic_slab_f64(pfdr_b200_solver **out, const pfdr_b200_slab *slab, int prox_kind, int Vloc,
    int Eloc, int N, const double *X0, const double *Y, const double *A, const int *Eu, const int *Ev, const double *La_d1,
    const double *La_l1, int positivity, double lo, double hi, int Ltype, const double *L, double rho, double condMin,
    double difRcd, double difTol, int itMax, int want_obj, int want_dif);
int pfdr_b200_slab_export(pfdr_b200_solver *s, void *blob);                              /* PFDR_B200_SLAB_BLOB bytes out */
int pfdr_b200_slab_connect(pfdr_b200_solver *s, const void *blobs, int same_process);    /* world x PFDR_B200_SLAB_BLOB in */

/* ---- run-time controls and introspection (no counterpart in the reference) ----------------- */
const char *pfdr_b200_last_error(void);      /* message of the last non-zero status (thread local) */
int pfdr_b200_version(void);
int pfdr_b200_set_device(int device);        /* CUDA device used by subsequent calls (default: current) */
int pfdr_b200_set_stream(void *cuda_stream); /* cudaStream_t all kernels are launched on (default 0)    */

typedef struct {
    double setup_ms;       /* upload + half-edge layout + preconditioning (CUDA events)           */
    double loop_ms;        /* the iteration loop only (CUDA events on the launch stream)          */
    double total_ms;       /* whole call, host wall clock                                         */
    long long launches;    /* kernels of this library launched by the call                        */
    long long loop_launches;    /* ... of which inside the iteration loop                         */
    int iterations;        /* == *it                                                              */
    long long h2d_bytes, d2h_bytes;  /* host<->device bytes moved by the call                     */
    long long slots;       /* padded half-edge slots of the SELL layout (>= 2E); 2E on the stencil layout */
    int layout;            /* 0: half-edge table (graph.cuh); 1: constant-offset / stencil layout (dia.cuh) */
    int gpus;              /* devices the solve ran on                                            */
} pfdr_b200_stats;
int pfdr_b200_get_last_stats(pfdr_b200_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* PFDR_B200_H */
