// multi.cu - several GPUs behind the reference's one-shot entry points (PFDR_graph_loss_d1_simplex,
// PFDR_graph_quadratic_d1_l1 / _bounds): with PFDR_B200_GPUS=N a large problem on a constant-offset (grid) graph is
// cut into N vertex slabs inside the call - one host thread per GPU, each building its slab solver from SUB-RANGES of
// the caller's host arrays (a slab of a vertex-major problem is contiguous in every argument), the slabs linked
// through peer memory (slab.cuh), the result written back into the caller's X / P.  Callers - the MEX gateways, the
// cut-pursuit loops - are unchanged: configuration 5 of BASELINE.json is a plain call of PFDR_graph_quadratic_d1_l1.
// Anything the slab path does not cover (irregular graphs, Obj record, reconditioning, dense operators, device
// pointers) is solved on one GPU as before; the decision is taken before any work is done.
#include "solvers.cuh"
#include "slab.cuh"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace pfdr {

cudaStream_t own_stream();          // capi.cu: this thread's library-owned stream

namespace {

struct Gate {                        // reusable barrier that can be broken by a failing rank
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
    bool broken = false;
    explicit Gate(int n_) : n(n_) {}
    bool wait()                      // false: another rank failed
    {
        std::unique_lock<std::mutex> l(m);
        if (broken) return false;
        const int ph = phase;
        if (++waiting == n) { waiting = 0; phase++; cv.notify_all(); return true; }
        cv.wait(l, [&] { return phase != ph || broken; });
        return !broken;
    }
    void abort() { std::lock_guard<std::mutex> l(m); broken = true; cv.notify_all(); }
};

int gpus_wanted()
{
    const char *e = getenv("PFDR_B200_GPUS");
    if (!e || !*e) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 1; }
    const int want = (!strcmp(e, "all") || !strcmp(e, "ALL")) ? ndev : atoi(e);
    return std::max(1, std::min(std::min(want, ndev), SLAB_MAXW));
}

// largest edge offset seen on a sample of the edge list (the slab solvers validate the whole list)
long long sample_band(const int *Eu, const int *Ev, int E)
{
    long long band = 0;
    const int step = std::max(1, E / 65536);
    for (long long e = 0; e < E; e += step) band = std::max<long long>(band, (long long) Ev[e] - Eu[e]);
    for (long long e = std::max(0, E - 64); e < E; e++) band = std::max<long long>(band, (long long) Ev[e] - Eu[e]);
    return band;
}

std::vector<long long> aligned_bounds(long long V, long long band, int world)
{
    std::vector<long long> b(world + 1);
    const long long planes = (band > 0 && V % band == 0) ? V / band : 0;
    for (int r = 0; r <= world; r++) b[r] = planes >= 2LL * world ? band * ((planes * r) / world) : (V * r) / world;
    return b;
}

struct RankEdges { std::vector<int> Eu, Ev; long long e0 = 0; };

// edges with an owned end are a contiguous run of a vertex-major list: Eu in [lo, v1)
RankEdges local_edges(const int *Eu, const int *Ev, int E, long long lo, long long v1)
{
    RankEdges r;
    const int *a = std::lower_bound(Eu, Eu + E, (int) lo), *b = std::lower_bound(Eu, Eu + E, (int) v1);
    r.e0 = a - Eu;
    const size_t n = (size_t) (b - a);
    r.Eu.resize(n); r.Ev.resize(n);
    for (size_t i = 0; i < n; i++) { r.Eu[i] = a[i] - (int) lo; r.Ev[i] = Ev[r.e0 + i] - (int) lo; }
    return r;
}

// runs make(rank, slab info, edges) on `world` threads, one per device, links the solvers and solves
template <typename MakeFn, typename FetchFn>
void run_ranks(int world, int base_device, const std::vector<long long> &bounds, long long V, long long band,
               const int *Eu, const int *Ev, int E, MakeFn make, FetchFn fetch)
{
    Gate gate(world);
    std::vector<unsigned char> table((size_t) world * PFDR_B200_SLAB_BLOB);
    std::vector<std::string> errors(world);
    std::vector<pfdr_b200_stats> stats(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; r++) {
        threads.emplace_back([&, r] {
            SolverBase *s = nullptr;
            try {
                Context &c = ctx();
                c.device = base_device + r;
                PFDR_CUDA(cudaSetDevice(c.device));
                c.stream = own_stream();
                c.stats = pfdr_b200_stats{};
                c.in_loop = false;
                pfdr_b200_slab sl{world, r, (int) V, (int) bounds[r], (int) bounds[r + 1],
                                  (int) std::max<long long>(bounds[r] - band, 0), (int) std::min<long long>(bounds[r + 1] + band, V)};
                RankEdges ed = local_edges(Eu, Ev, E, sl.lo, sl.v1);
                s = make(r, sl, ed);
                s->slab_export(table.data() + (size_t) r * PFDR_B200_SLAB_BLOB);
                if (!gate.wait()) throw Error(PFDR_B200_ERR_CUDA, "another rank failed");
                s->slab_connect(table.data(), true);
                if (!gate.wait()) throw Error(PFDR_B200_ERR_CUDA, "another rank failed");
                s->run();
                fetch(r, sl, s);
                stats[r] = c.stats;
                gate.wait();                                 // nobody frees memory a neighbour may still write
            } catch (const std::exception &e) {
                errors[r] = e.what();
                cudaGetLastError();
                gate.abort();
            }
            delete s;
            slab_release_cache();                            // this thread ends here: its cached arena goes back now
        });
    }
    for (auto &t : threads) t.join();
    for (int r = 0; r < world; r++)
        if (!errors[r].empty()) throw Error(PFDR_B200_ERR_CUDA, "multi-GPU solve, rank " + std::to_string(r) + ": " + errors[r]);
    pfdr_b200_stats &o = ctx().stats;
    o = stats[0];
    for (int r = 1; r < world; r++) {
        o.loop_ms = std::max(o.loop_ms, stats[r].loop_ms); o.setup_ms = std::max(o.setup_ms, stats[r].setup_ms);
        o.launches += stats[r].launches; o.loop_launches += stats[r].loop_launches;
        o.h2d_bytes += stats[r].h2d_bytes; o.d2h_bytes += stats[r].d2h_bytes; o.slots += stats[r].slots;
    }
    o.gpus = world; o.layout = 1;
}

int current_device()
{
    int d = 0;
    if (ctx().device >= 0) return ctx().device;
    cudaGetDevice(&d);
    return d;
}

}  // namespace

template <typename real>
bool solve_simplex_multi(const SimplexProblem<real> &q)
{
    const int world = gpus_wanted();
    if (world < 2 || q.K < 3 || q.E < 1 || q.itMax < 1) return false;
    if (is_device_ptr(q.P) || is_device_ptr(q.Q) || is_device_ptr(q.Eu) || is_device_ptr(q.Ev) || is_device_ptr(q.La_d1)) return false;
    if ((long long) q.V * q.K < (1LL << 22)) return false;                       // small problems: one GPU is faster
    const long long band = sample_band(q.Eu, q.Ev, q.E);
    if (band < 1 || !std::is_sorted(q.Eu, q.Eu + std::min(q.E, 4096))) return false;
    const std::vector<long long> bounds = aligned_bounds(q.V, band, world);
    for (int r = 0; r < world; r++) if (bounds[r + 1] - bounds[r] < band) return false;
    const size_t K = (size_t) q.K;
    int it = 0;
    try {
    run_ranks(world, current_device(), bounds, q.V, band, q.Eu, q.Ev, q.E,
        [&](int r, const pfdr_b200_slab &sl, const RankEdges &ed) {
            SimplexProblem<real> l = q;
            l.V = sl.hi - sl.lo; l.E = (int) ed.Eu.size();
            l.P = q.P + (size_t) sl.lo * K; l.Q = q.Q + (size_t) sl.lo * K;
            l.La_f = q.La_f ? q.La_f + sl.lo : nullptr;
            l.Eu = ed.Eu.data(); l.Ev = ed.Ev.data(); l.La_d1 = q.La_d1 + ed.e0;
            l.it = nullptr; l.Obj = q.Obj ? reinterpret_cast<real *>(sizeof(real)) : nullptr;
            l.Dif = q.Dif ? reinterpret_cast<real *>(sizeof(real)) : nullptr;
            l.verbose = 0; l.shard = nullptr; l.slab = &sl;
            return make_simplex<real>(l);
        },
        [&](int r, const pfdr_b200_slab &sl, SolverBase *s) {
            int it_r = 0;
            s->fetch(q.P + (size_t) sl.v0 * K, &it_r, r == 0 ? q.Obj : nullptr, r == 0 ? q.Dif : nullptr);
            if (r == 0) it = it_r;
        });
    } catch (const Error &e) {       // not a grid graph after all, GPUs without peer access, ...: one GPU takes it
        if (getenv("PFDR_B200_TRACE")) fprintf(stderr, "[pfdr_b200] multi-GPU path not taken: %s\n", e.what());
        return false;
    }
    if (q.it) *q.it = it;
    if (q.verbose) { printf("PFDR simplex (B200 x%d, slab-sharded): %d iterations (max. %d)\n", world, it, q.itMax); fflush(stdout); }
    return true;
}
template bool solve_simplex_multi<float>(const SimplexProblem<float> &);
template bool solve_simplex_multi<double>(const SimplexProblem<double> &);

template <typename real>
bool solve_quadratic_multi(const QuadProblem<real> &q)
{
    const int world = gpus_wanted();
    if (world < 2 || q.N != 0 || q.E < 1 || q.itMax < 1) return false;
    if (is_device_ptr(q.X) || is_device_ptr(q.Y) || is_device_ptr(q.Eu) || is_device_ptr(q.Ev) || is_device_ptr(q.La_d1)) return false;
    if (q.V < (1 << 21)) return false;                                           // small problems: one GPU is faster
    const long long band = sample_band(q.Eu, q.Ev, q.E);
    if (band < 1 || !std::is_sorted(q.Eu, q.Eu + std::min(q.E, 4096))) return false;
    const std::vector<long long> bounds = aligned_bounds(q.V, band, world);
    for (int r = 0; r < world; r++) if (bounds[r + 1] - bounds[r] < band) return false;
    int it = 0;
    try {
    run_ranks(world, current_device(), bounds, q.V, band, q.Eu, q.Ev, q.E,
        [&](int r, const pfdr_b200_slab &sl, const RankEdges &ed) {
            QuadProblem<real> l = q;
            l.V = sl.hi - sl.lo; l.E = (int) ed.Eu.size();
            l.X = q.X + sl.lo; l.Y = q.Y + sl.lo; l.A = q.A ? q.A + sl.lo : nullptr;
            l.La_l1 = q.La_l1 ? q.La_l1 + sl.lo : nullptr;
            l.L = (q.L && q.Ltype == PFDR_B200_DIAG) ? q.L + sl.lo : q.L;
            l.Eu = ed.Eu.data(); l.Ev = ed.Ev.data(); l.La_d1 = q.La_d1 + ed.e0;
            l.it = nullptr; l.Obj = q.Obj ? reinterpret_cast<real *>(sizeof(real)) : nullptr;
            l.Dif = q.Dif ? reinterpret_cast<real *>(sizeof(real)) : nullptr;
            l.verbose = 0; l.shard = nullptr; l.slab = &sl;
            return make_quadratic<real>(l);
        },
        [&](int r, const pfdr_b200_slab &sl, SolverBase *s) {
            int it_r = 0;
            s->fetch(q.X + sl.v0, &it_r, r == 0 ? q.Obj : nullptr, r == 0 ? q.Dif : nullptr);
            if (r == 0) it = it_r;
        });
    } catch (const Error &e) {
        if (getenv("PFDR_B200_TRACE")) fprintf(stderr, "[pfdr_b200] multi-GPU path not taken: %s\n", e.what());
        return false;
    }
    if (q.it) *q.it = it;
    if (q.verbose) { printf("PFDR (B200 x%d, slab-sharded): %d iterations (max. %d)\n", world, it, q.itMax); fflush(stdout); }
    return true;
}
template bool solve_quadratic_multi<float>(const QuadProblem<float> &);
template bool solve_quadratic_multi<double>(const QuadProblem<double> &);

}  // namespace pfdr
