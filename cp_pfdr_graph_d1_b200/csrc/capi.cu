// capi.cu - the extern "C" surface declared in include/pfdr_b200.h.
#include "solvers.cuh"
#include <chrono>

namespace pfdr {

Context &ctx()
{
    static thread_local Context c;
    return c;
}

static thread_local std::string g_err;

// library-owned launch stream: graph capture is not allowed on the legacy default stream
cudaStream_t own_stream()
{
    static thread_local cudaStream_t s = nullptr;
    if (!s) PFDR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}

static thread_local pfdr_b200_stats g_last_stats{};

template <typename F>
static int guarded(F &&f)
{
    try {
        Context &c = ctx();
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Error(PFDR_B200_ERR_CUDA, std::string("no usable CUDA device (libpfdr_b200 has no CPU fallback): ") +
                                                cudaGetErrorString(e));
        if (c.device >= 0) PFDR_CUDA(cudaSetDevice(c.device));
        if (!c.user_stream) c.stream = own_stream();
        c.stats = pfdr_b200_stats{};
        c.stats.gpus = 1;
        c.in_loop = false;
        c.ordered = false;
        const auto t0 = std::chrono::steady_clock::now();
        f();
        c.stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        g_last_stats = c.stats;
        return PFDR_B200_OK;
    } catch (const Error &e) {
        g_err = e.what();
        cudaGetLastError();
        return e.status;
    } catch (const std::exception &e) {
        g_err = e.what();
        return PFDR_B200_ERR_CUDA;
    }
}

}  // namespace pfdr

using namespace pfdr;

extern "C" {

const char *pfdr_b200_last_error(void) { return g_err.c_str(); }
int pfdr_b200_version(void) { return 100; }
int pfdr_b200_set_device(int device) { ctx().device = device; return PFDR_B200_OK; }
int pfdr_b200_set_stream(void *s)
{
    ctx().stream = (cudaStream_t) s;
    ctx().user_stream = (s != nullptr);
    return PFDR_B200_OK;
}
int pfdr_b200_get_last_stats(pfdr_b200_stats *out) { if (!out) return PFDR_B200_ERR_ARG; *out = g_last_stats; return PFDR_B200_OK; }

#define QUAD_L1(SUF, real)                                                                              \
int pfdr_b200_quadratic_d1_l1_##SUF(int V, int E, int N, real *X, const real *Y, const real *A,         \
    const int *Eu, const int *Ev, const real *La_d1, const real *La_l1, int positivity, int Ltype,      \
    const real *L, real rho, real condMin, real difRcd, real difTol, int itMax, int *it, real *Obj,     \
    real *Dif, int verbose)                                                                             \
{                                                                                                       \
    return guarded([&] {                                                                                \
        QuadProblem<real> q{V, E, N, X, Y, A, Eu, Ev, La_d1, La_l1, positivity, 0, 0, 0, Ltype, L, rho, \
                            condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose};                     \
        if (!solve_quadratic_multi<real>(q)) solve_quadratic<real>(q);                                  \
    });                                                                                                 \
}                                                                                                       \
int pfdr_b200_quadratic_d1_bounds_##SUF(int V, int E, int N, real *X, const real *Y, const real *A,     \
    const int *Eu, const int *Ev, const real *La_d1, real lo, real hi, int Ltype, const real *L,        \
    real rho, real condMin, real difRcd, real difTol, int itMax, int *it, real *Obj, real *Dif,         \
    int verbose)                                                                                        \
{                                                                                                       \
    return guarded([&] {                                                                                \
        QuadProblem<real> q{V, E, N, X, Y, A, Eu, Ev, La_d1, nullptr, 0, 1, lo, hi, Ltype, L, rho,      \
                            condMin, difRcd, difTol, itMax, it, Obj, Dif, verbose};                     \
        if (!solve_quadratic_multi<real>(q)) solve_quadratic<real>(q);                                  \
    });                                                                                                 \
}

QUAD_L1(f32, float)
QUAD_L1(f64, double)

#define SIMPLEX(SUF, real)                                                                              \
int pfdr_b200_loss_d1_simplex_##SUF(int K, int V, int E, real al, const real *La_f, real *P,            \
    const real *Q, const int *Eu, const int *Ev, const real *La_d1, real rho, real condMin,             \
    real difRcd, real difTol, int itMax, int *it, real *Obj, real *Dif, int verbose)                    \
{                                                                                                       \
    return guarded([&] {                                                                                \
        SimplexProblem<real> q{K, V, E, al, La_f, P, Q, Eu, Ev, La_d1, rho, condMin, difRcd, difTol,    \
                               itMax, it, Obj, Dif, verbose};                                           \
        if (!solve_simplex_multi<real>(q)) solve_simplex<real>(q);                                      \
    });                                                                                                 \
}                                                                                                       \
int pfdr_b200_proj_simplex_metric_##SUF(real *X, const real *M, int D, int N, int nm, const real *A,    \
                                        int na)                                                         \
{                                                                                                       \
    return guarded([&] { proj_simplex_metric<real>(X, M, D, N, nm, A, na); });                          \
}                                                                                                       \
int pfdr_b200_operator_norm_matrix_##SUF(int M, int N, const real *A, real nTol, int itMax,             \
                                         int nbInit, int verbose, real *result)                         \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(result != nullptr, "operator_norm_matrix: result must not be NULL");               \
        *result = operator_norm_matrix<real>(M, N, A, nTol, itMax, nbInit, verbose);                    \
    });                                                                                                 \
}

SIMPLEX(f32, float)
SIMPLEX(f64, double)

/* ---- resident solver handles ---- */
#define HANDLE(SUF, real)                                                                               \
int pfdr_b200_solver_create_quadratic_##SUF(pfdr_b200_solver **out, int prox_kind, int V, int E, int N, \
    const real *X0, const real *Y, const real *A, const int *Eu, const int *Ev, const real *La_d1,      \
    const real *La_l1, int positivity, real lo, real hi, int Ltype, const real *L, real rho,            \
    real condMin, real difRcd, real difTol, int itMax, int want_obj, int want_dif)                      \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr, "solver_create: out must not be NULL");                            \
        real *flag = reinterpret_cast<real *>(sizeof(real));   /* non-NULL marker, never dereferenced */\
        QuadProblem<real> q{V, E, N, const_cast<real *>(X0), Y, A, Eu, Ev, La_d1,                       \
                            prox_kind == 0 ? La_l1 : nullptr, prox_kind == 0 ? positivity : 0,          \
                            prox_kind, lo, hi, Ltype, L, rho, condMin, difRcd, difTol, itMax, nullptr,  \
                            want_obj ? flag : nullptr, want_dif ? flag : nullptr, 0};                   \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_quadratic<real>(q));                           \
    });                                                                                                 \
}                                                                                                       \
int pfdr_b200_solver_create_simplex_##SUF(pfdr_b200_solver **out, int K, int V, int E, real al,         \
    const real *La_f, const real *P0, const real *Q, const int *Eu, const int *Ev, const real *La_d1,   \
    real rho, real condMin, real difRcd, real difTol, int itMax, int want_obj, int want_dif)            \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr, "solver_create: out must not be NULL");                            \
        real *flag = reinterpret_cast<real *>(sizeof(real));                                            \
        SimplexProblem<real> q{K, V, E, al, La_f, const_cast<real *>(P0), Q, Eu, Ev, La_d1, rho,        \
                               condMin, difRcd, difTol, itMax, nullptr, want_obj ? flag : nullptr,      \
                               want_dif ? flag : nullptr, 0};                                           \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_simplex<real>(q));                             \
    });                                                                                                 \
}

HANDLE(f32, float)
HANDLE(f64, double)

/* ---- sharded handles: the shard description is only read during create ---- */
#define HANDLE_SHARDED(SUF, real)                                                                       \
int pfdr_b200_solver_create_quadratic_sharded_##SUF(pfdr_b200_solver **out, const pfdr_b200_shard *sh,  \
    int prox_kind, int V, int E, int N, const real *X0, const real *Y, const real *A, const int *Eu,    \
    const int *Ev, const real *La_d1, const real *La_l1, int positivity, real lo, real hi, int Ltype,   \
    const real *L, real rho, real condMin, real difRcd, real difTol, int itMax, int want_obj,           \
    int want_dif)                                                                                       \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr && sh != nullptr, "solver_create_sharded: NULL argument");          \
        real *flag = reinterpret_cast<real *>(sizeof(real));                                            \
        QuadProblem<real> q{V, E, N, const_cast<real *>(X0), Y, A, Eu, Ev, La_d1,                       \
                            prox_kind == 0 ? La_l1 : nullptr, prox_kind == 0 ? positivity : 0,          \
                            prox_kind, lo, hi, Ltype, L, rho, condMin, difRcd, difTol, itMax, nullptr,  \
                            want_obj ? flag : nullptr, want_dif ? flag : nullptr, 0, sh};               \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_quadratic<real>(q));                           \
    });                                                                                                 \
}                                                                                                       \
int pfdr_b200_solver_create_simplex_sharded_##SUF(pfdr_b200_solver **out, const pfdr_b200_shard *sh,    \
    int K, int V, int E, real al, const real *La_f, const real *P0, const real *Q, const int *Eu,       \
    const int *Ev, const real *La_d1, real rho, real condMin, real difRcd, real difTol, int itMax,      \
    int want_obj, int want_dif)                                                                         \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr && sh != nullptr, "solver_create_sharded: NULL argument");          \
        real *flag = reinterpret_cast<real *>(sizeof(real));                                            \
        SimplexProblem<real> q{K, V, E, al, La_f, const_cast<real *>(P0), Q, Eu, Ev, La_d1, rho,        \
                               condMin, difRcd, difTol, itMax, nullptr, want_obj ? flag : nullptr,      \
                               want_dif ? flag : nullptr, 0, sh};                                       \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_simplex<real>(q));                             \
    });                                                                                                 \
}

HANDLE_SHARDED(f32, float)
HANDLE_SHARDED(f64, double)

#define HANDLE_SLAB(SUF, real)                                                                          \
int pfdr_b200_solver_create_simplex_slab_##SUF(pfdr_b200_solver **out, const pfdr_b200_slab *sl, int K, \
    int Vloc, int Eloc, real al, const real *La_f, const real *P0, const real *Q, const int *Eu,        \
    const int *Ev, const real *La_d1, real rho, real condMin, real difRcd, real difTol, int itMax,      \
    int want_obj, int want_dif)                                                                         \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr && sl != nullptr, "solver_create_slab: NULL argument");             \
        real *flag = reinterpret_cast<real *>(sizeof(real));                                            \
        SimplexProblem<real> q{K, Vloc, Eloc, al, La_f, const_cast<real *>(P0), Q, Eu, Ev, La_d1, rho,  \
                               condMin, difRcd, difTol, itMax, nullptr, want_obj ? flag : nullptr,      \
                               want_dif ? flag : nullptr, 0, nullptr, sl};                              \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_simplex<real>(q));                             \
    });                                                                                                 \
}                                                                                                       \
int pfdr_b200_solver_create_quadratic_slab_##SUF(pfdr_b200_solver **out, const pfdr_b200_slab *sl,      \
    int prox_kind, int Vloc, int Eloc, int N, const real *X0, const real *Y, const real *A,             \
    const int *Eu, const int *Ev, const real *La_d1, const real *La_l1, int positivity, real lo,        \
    real hi, int Ltype, const real *L, real rho, real condMin, real difRcd, real difTol, int itMax,     \
    int want_obj, int want_dif)                                                                         \
{                                                                                                       \
    return guarded([&] {                                                                                \
        PFDR_REQUIRE(out != nullptr && sl != nullptr, "solver_create_slab: NULL argument");             \
        real *flag = reinterpret_cast<real *>(sizeof(real));                                            \
        QuadProblem<real> q{Vloc, Eloc, N, const_cast<real *>(X0), Y, A, Eu, Ev, La_d1, La_l1,          \
                            positivity, prox_kind, lo, hi, Ltype, L, rho, condMin, difRcd, difTol,      \
                            itMax, nullptr, want_obj ? flag : nullptr, want_dif ? flag : nullptr, 0,    \
                            nullptr, sl};                                                               \
        *out = reinterpret_cast<pfdr_b200_solver *>(make_quadratic<real>(q));                           \
    });                                                                                                 \
}
HANDLE_SLAB(f32, float)
HANDLE_SLAB(f64, double)

int pfdr_b200_slab_export(pfdr_b200_solver *s, void *blob)
{
    return guarded([&] { PFDR_REQUIRE(s && blob, "NULL argument"); reinterpret_cast<SolverBase *>(s)->slab_export(blob); });
}
int pfdr_b200_slab_connect(pfdr_b200_solver *s, const void *blobs, int same_process)
{
    return guarded([&] { PFDR_REQUIRE(s && blobs, "NULL argument"); reinterpret_cast<SolverBase *>(s)->slab_connect(blobs, same_process != 0); });
}

/* ---- cut-pursuit outer iteration: resident operator, reduced-problem construction ---- */
int pfdr_b200_cp_operator_create_f32(pfdr_b200_cp_operator **out, int V, int N, const float *A)
{
    return guarded([&] { PFDR_REQUIRE(out != nullptr, "cp_operator_create: out must not be NULL");
                         *out = reinterpret_cast<pfdr_b200_cp_operator *>(make_cp_operator<float>(V, N, A)); });
}
int pfdr_b200_cp_operator_create_f64(pfdr_b200_cp_operator **out, int V, int N, const double *A)
{
    return guarded([&] { PFDR_REQUIRE(out != nullptr, "cp_operator_create: out must not be NULL");
                         *out = reinterpret_cast<pfdr_b200_cp_operator *>(make_cp_operator<double>(V, N, A)); });
}
int pfdr_b200_cp_set_components(pfdr_b200_cp_operator *op, int rV, const int *Cv)
{
    return guarded([&] { PFDR_REQUIRE(op, "NULL operator"); reinterpret_cast<CpOperatorBase *>(op)->set_components(rV, Cv); });
}
int pfdr_b200_cp_reduce(pfdr_b200_cp_operator *op, const void *Y, int preAt, void *rA, void *rAA, void *rY)
{
    return guarded([&] { PFDR_REQUIRE(op, "NULL operator"); reinterpret_cast<CpOperatorBase *>(op)->reduce(Y, preAt, rA, rAA, rY); });
}
int pfdr_b200_cp_gradient(pfdr_b200_cp_operator *op, const void *rX, const void *YorR, void *DfS)
{
    return guarded([&] { PFDR_REQUIRE(op, "NULL operator"); reinterpret_cast<CpOperatorBase *>(op)->gradient(rX, YorR, DfS); });
}
int pfdr_b200_cp_residual(pfdr_b200_cp_operator *op, const void *rX, const void *Y, void *R)
{
    return guarded([&] { PFDR_REQUIRE(op, "NULL operator"); reinterpret_cast<CpOperatorBase *>(op)->residual(rX, Y, R); });
}
int pfdr_b200_cp_operator_destroy(pfdr_b200_cp_operator *op)
{
    return guarded([&] { delete reinterpret_cast<CpOperatorBase *>(op); });
}

int pfdr_b200_cp_components(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, int *Cv, int *rV)
{
    return guarded([&] { cp_components(V, E, Eu, Ev, active, Cv, rV); });
}
#define CP_GRAPH(SUF, real)                                                                                       \
int pfdr_b200_cp_reduced_edges_##SUF(int V, int E, const int *Eu, const int *Ev, const unsigned char *active,     \
    const real *La_d1, const int *Cv, int rV, real eps, int capacity, int *rE, int *rEu, int *rEv, real *rLa_d1)  \
{                                                                                                                 \
    return guarded([&] { cp_reduced_edges<real>(V, E, Eu, Ev, active, La_d1, Cv, rV, eps, capacity, rE, rEu, rEv, \
                                                rLa_d1); });                                                      \
}                                                                                                                 \
int pfdr_b200_cp_merge_##SUF(int V, int E, const int *Eu, const int *Ev, const int *Cv, int rV, const real *rX,   \
    real CP_difTol, real eps, unsigned char *active, int *deactivated)                                            \
{                                                                                                                 \
    return guarded([&] { cp_merge<real>(V, E, Eu, Ev, Cv, rV, rX, CP_difTol, eps, active, deactivated); });       \
}
CP_GRAPH(f32, float)
CP_GRAPH(f64, double)

int pfdr_b200_comm_unique_id(void *out128)
{
    return guarded([&] { PFDR_REQUIRE(out128, "NULL id buffer"); comm_unique_id(out128); });
}
int pfdr_b200_comm_init(const void *id128, int world, int rank)
{
    return guarded([&] { PFDR_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
                         comm_init(id128, world, rank); });
}
int pfdr_b200_comm_destroy(void) { return guarded([&] { comm_destroy(); }); }

int pfdr_b200_solver_reset(pfdr_b200_solver *s)
{
    return guarded([&] { PFDR_REQUIRE(s, "NULL solver"); reinterpret_cast<SolverBase *>(s)->reset(); });
}
int pfdr_b200_solver_run(pfdr_b200_solver *s)
{
    return guarded([&] { PFDR_REQUIRE(s, "NULL solver"); reinterpret_cast<SolverBase *>(s)->run(); });
}
int pfdr_b200_solver_fetch(pfdr_b200_solver *s, void *X, int *it, void *Obj, void *Dif)
{
    return guarded([&] { PFDR_REQUIRE(s, "NULL solver"); reinterpret_cast<SolverBase *>(s)->fetch(X, it, Obj, Dif); });
}
int pfdr_b200_solver_destroy(pfdr_b200_solver *s)
{
    return guarded([&] { delete reinterpret_cast<SolverBase *>(s); });
}

}  // extern "C"
