// cpgraph.cu - the graph passes of the cut-pursuit's outer iteration on the device:
//   components     connected components of the non-active edges, numbered by smallest vertex   reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:570-600
//   reduced_edges  one edge per pair of neighbouring components, weights added up, loops        reference :603-661
//   merge          deactivate active edges between components of (relatively) equal value       reference :863-886
// The reference runs them as serial O(V + E) host sweeps.  Here: lock-free union-find over the edge list (roots only ever
// point to smaller vertices, so a component's root is its smallest vertex and the numbering is the reference's), a
// stable radix sort of the cut edges by component pair with per-pair sums in edge order (deterministic), and an
// element-wise pass.  Integer outputs (Cv, rV, the reduced pairs, the flags) are bit-identical to the reference's;
// the reduced pairs come out sorted instead of in first-encounter order.
#include "solvers.cuh"
#include <cub/cub.cuh>

namespace pfdr {
namespace {

__device__ __forceinline__ int uf_find(int *parent, int v)
{
    volatile int *vp = parent;
    int p = vp[v];
    while (p != v) {
        const int gp = vp[p];
        if (gp != p) vp[v] = gp;          // path halving: any ancestor is a valid parent
        v = p; p = gp;
    }
    return v;
}

__global__ void uf_init_kernel(int *parent, int V)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) parent[v] = v;
}

__global__ void uf_hook_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, const unsigned char *__restrict__ active, int E, int V,
                               int *parent, int *bad)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int u = Eu[e], v = Ev[e];
    if (u < 0 || u >= V || v < 0 || v >= V) { atomicOr(bad, 1); return; }
    if (active[e]) return;
    int a = uf_find(parent, u), b = uf_find(parent, v);
    while (a != b) {
        if (a < b) { const int t = a; a = b; b = t; }          // a > b: the larger root goes under the smaller one
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) break;
        a = uf_find(parent, old);
        b = uf_find(parent, b);
    }
}

__global__ void uf_flatten_kernel(int *parent, int V, int *isroot)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    volatile int *vp = parent;
    int r = vp[v];
    while (vp[r] != r) r = vp[r];
    isroot[v] = (r == v);
    __syncwarp();
    if (r != v) parent[v] = r;            // roots are never written here; non-roots only ever move closer to their root
}

__global__ void uf_number_kernel(const int *__restrict__ parent, const int *__restrict__ rank, int V, int *__restrict__ Cv)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int r = parent[v];
    while (parent[r] != r) r = parent[r];   // a flattened pointer may be one step behind its root's final position
    Cv[v] = rank[r];
}

// ---- reduced edges -------------------------------------------------------------------------
template <typename real>
__global__ void cut_touch_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, const unsigned char *__restrict__ active,
                                 const real *__restrict__ La_d1, const int *__restrict__ Cv, int E, unsigned char *touched)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || !active[e] || La_d1[e] == (real) 0) return;
    touched[Cv[Eu[e]]] = 1;
    touched[Cv[Ev[e]]] = 1;
}

// entries 0..E-1: the edges; E..E+rV-1: one per component (valid when nothing touches it: the loop of weight eps)
template <typename real>
__global__ void cut_entry_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, const unsigned char *__restrict__ active,
                                 const real *__restrict__ La_d1, const int *__restrict__ Cv, const unsigned char *__restrict__ touched,
                                 int E, int rV, real eps, unsigned long long *__restrict__ key, double *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E + rV) return;
    const unsigned long long none = (unsigned long long) rV * (unsigned long long) rV;
    if (i < E) {
        if (!active[i] || La_d1[i] == (real) 0) { key[i] = none; val[i] = 0.0; return; }
        const int a = Cv[Eu[i]], b = Cv[Ev[i]], lo = min(a, b), hi = max(a, b);
        key[i] = (unsigned long long) lo * (unsigned long long) rV + (unsigned long long) hi;
        val[i] = (a == b ? 2.0 : 1.0) * (double) La_d1[i];       // a cut edge inside a component is met from both of its ends
    } else {
        const int c = i - E;
        if (touched[c]) { key[i] = none; val[i] = 0.0; }
        else { key[i] = (unsigned long long) c * (unsigned long long) rV + (unsigned long long) c; val[i] = (double) eps; }
    }
}

__global__ void seg_flag_kernel(const unsigned long long *__restrict__ key, int n, unsigned long long none, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (key[i] != none && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
}

// start[s] = position of the first entry of reduced edge s;  start[rE] = number of valid entries
__global__ void seg_start_kernel(const unsigned long long *__restrict__ key, const int *__restrict__ flag, const int *__restrict__ seg, int n,
                                 unsigned long long none, int *__restrict__ start, int *__restrict__ total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) start[seg[i]] = i;
    const bool valid = key[i] != none;
    if (valid && (i == n - 1 || key[i + 1] == none)) { start[seg[i] + flag[i]] = i + 1; total[0] = seg[i] + flag[i]; }
}

// one warp per reduced edge: the entries of a pair are in edge order (stable sort); lane-strided partial sums, fixed tree
template <typename real>
__global__ void __launch_bounds__(256) seg_sum_kernel(const unsigned long long *__restrict__ key, const double *__restrict__ val,
                                                     const int *__restrict__ start, const int *__restrict__ total, int rV, int cap,
                                                     int *__restrict__ rEu, int *__restrict__ rEv, real *__restrict__ rLa)
{
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= total[0] || s >= cap) return;
    const int b = start[s], e = start[s + 1];
    double acc = 0.0;
    for (int i = b + lane; i < e; i += 32) acc += val[i];
    acc = warp_sum(acc);
    if (lane == 0) {
        const unsigned long long k = key[b];
        rEu[s] = (int) (k / (unsigned long long) rV);
        rEv[s] = (int) (k % (unsigned long long) rV);
        rLa[s] = (real) acc;
    }
}

template <typename real>
__global__ void merge_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, const int *__restrict__ Cv, const real *__restrict__ rX,
                             int E, real tol, real eps, unsigned char *active, int *count)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || !active[e]) return;
    real a = rX[Cv[Eu[e]]], b = rX[Cv[Ev[e]]], d = a - b;
    a = a < (real) 0 ? -a : a;
    b = b < (real) 0 ? -b : b;
    d = d < (real) 0 ? -d : d;
    if (a < b) a = b;
    d = (a > eps) ? d / a : d / eps;
    if (d <= tol) { active[e] = 0; atomicAdd(count, 1); }
}

}  // namespace

// Cv[V] and *rV from the edge list and its active flags
void cp_components(int V, int E, const int *Eu_, const int *Ev_, const unsigned char *active_, int *Cv_out, int *rV_out)
{
    PFDR_REQUIRE(V >= 1 && E >= 0, "cp_components: need V >= 1, E >= 0");
    PFDR_REQUIRE(Cv_out && rV_out, "cp_components: Cv and rV must not be NULL");
    PFDR_REQUIRE(E == 0 || (Eu_ && Ev_ && active_), "cp_components: edge arrays must not be NULL");
    cudaStream_t st = ctx().stream;
    DevIn<int> Eu, Ev; DevIn<unsigned char> active;
    Eu.bind(Eu_, E); Ev.bind(Ev_, E); active.bind(active_, E);
    DevBuf<int> parent(V), isroot(V), rank(V + 1), Cv(V), bad(1);
    bad.zero();
    PFDR_LAUNCH(uf_init_kernel, cdiv(V, 256), 256, 0, parent.p, V);
    if (E) PFDR_LAUNCH(uf_hook_kernel, cdiv(E, 256), 256, 0, Eu.p, Ev.p, active.p, E, V, parent.p, bad.p);
    PFDR_LAUNCH(uf_flatten_kernel, cdiv(V, 256), 256, 0, parent.p, V, isroot.p);
    size_t tb = 0;
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, isroot.p, rank.p, V, st));
    DevBuf<unsigned char> tmp(tb + 16);
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, isroot.p, rank.p, V, st));
    ctx().stats.launches += 1;
    PFDR_LAUNCH(uf_number_kernel, cdiv(V, 256), 256, 0, parent.p, rank.p, V, Cv.p);
    int tail[2] = {0, 0}, hbad = 0;
    PFDR_CUDA(cudaMemcpyAsync(&tail[0], rank.p + (V - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaMemcpyAsync(&tail[1], isroot.p + (V - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    copy_out(Cv_out, Cv.p, (size_t) V);
    PFDR_CUDA(cudaStreamSynchronize(st));
    PFDR_REQUIRE(!hbad, "cp_components: an edge names a vertex outside [0, V)");
    *rV_out = tail[0] + tail[1];
}

template <typename real>
void cp_reduced_edges(int V, int E, const int *Eu_, const int *Ev_, const unsigned char *active_, const real *La_d1_, const int *Cv_,
                      int rV, real eps, int cap, int *rE_out, int *rEu_out, int *rEv_out, real *rLa_out)
{
    PFDR_REQUIRE(V >= 1 && E >= 0 && rV >= 1 && rV <= V, "cp_reduced_edges: need V >= 1, E >= 0, 1 <= rV <= V");
    PFDR_REQUIRE(Cv_ && rE_out && rEu_out && rEv_out && rLa_out && cap >= 1, "cp_reduced_edges: NULL argument");
    PFDR_REQUIRE(E == 0 || (Eu_ && Ev_ && active_ && La_d1_), "cp_reduced_edges: edge arrays must not be NULL");
    cudaStream_t st = ctx().stream;
    DevIn<int> Eu, Ev, Cv; DevIn<unsigned char> active; DevIn<real> La;
    Eu.bind(Eu_, E); Ev.bind(Ev_, E); active.bind(active_, E); La.bind(La_d1_, E); Cv.bind(Cv_, V);
    const int n = E + rV;
    DevBuf<unsigned char> touched(rV);
    touched.zero();
    if (E) PFDR_LAUNCH(cut_touch_kernel<real>, cdiv(E, 256), 256, 0, Eu.p, Ev.p, active.p, La.p, Cv.p, E, touched.p);
    DevBuf<unsigned long long> key(n), skey(n);
    DevBuf<double> val(n), sval(n);
    PFDR_LAUNCH(cut_entry_kernel<real>, cdiv(n, 256), 256, 0, Eu.p, Ev.p, active.p, La.p, Cv.p, touched.p, E, rV, eps, key.p, val.p);
    const unsigned long long none = (unsigned long long) rV * (unsigned long long) rV;
    int bits = 1; while (bits < 64 && (1ull << bits) <= none) bits++;
    DevBuf<int> flag(n), seg(n), start(n + 2), total(1);
    total.zero();
    size_t t1 = 0, t2 = 0;
    PFDR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, key.p, skey.p, val.p, sval.p, n, 0, bits, st));
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, flag.p, seg.p, n, st));
    DevBuf<unsigned char> tmp(std::max(t1, t2) + 16);
    size_t tb = tmp.n;
    PFDR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, skey.p, val.p, sval.p, n, 0, bits, st));   // stable: edge order inside a pair
    PFDR_LAUNCH(seg_flag_kernel, cdiv(n, 256), 256, 0, skey.p, n, none, flag.p);
    tb = tmp.n;
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, seg.p, n, st));
    ctx().stats.launches += 2;
    PFDR_LAUNCH(seg_start_kernel, cdiv(n, 256), 256, 0, skey.p, flag.p, seg.p, n, none, start.p, total.p);
    int rE = 0;
    PFDR_CUDA(cudaMemcpyAsync(&rE, total.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaStreamSynchronize(st));
    *rE_out = rE;
    if (rE > cap) throw Error(PFDR_B200_ERR_ARG, "cp_reduced_edges: " + std::to_string(rE) + " reduced edges, room for " + std::to_string(cap));
    if (!rE) return;
    DevBuf<int> rEu(rE), rEv(rE);
    DevBuf<real> rLa(rE);
    PFDR_LAUNCH(seg_sum_kernel<real>, cdiv(rE, 8), 256, 0, skey.p, sval.p, start.p, total.p, rV, rE, rEu.p, rEv.p, rLa.p);
    copy_out(rEu_out, rEu.p, (size_t) rE);
    copy_out(rEv_out, rEv.p, (size_t) rE);
    copy_out(rLa_out, rLa.p, (size_t) rE);
    PFDR_CUDA(cudaStreamSynchronize(st));
}

template <typename real>
void cp_merge(int V, int E, const int *Eu_, const int *Ev_, const int *Cv_, int rV, const real *rX_, real tol, real eps, unsigned char *active_io,
              int *count_out)
{
    PFDR_REQUIRE(E >= 0 && rV >= 1, "cp_merge: need E >= 0, rV >= 1");
    if (count_out) *count_out = 0;
    if (!E) return;
    PFDR_REQUIRE(Eu_ && Ev_ && Cv_ && rX_ && active_io, "cp_merge: NULL argument");
    cudaStream_t st = ctx().stream;
    DevIn<int> Eu, Ev, Cv; DevIn<real> rX;
    Eu.bind(Eu_, E); Ev.bind(Ev_, E); Cv.bind(Cv_, V); rX.bind(rX_, rV);
    const bool on_device = is_device_ptr(active_io);
    DevBuf<unsigned char> own;
    unsigned char *act = active_io;
    if (!on_device) {
        own.alloc(E);
        PFDR_CUDA(cudaMemcpyAsync(own.p, active_io, E, cudaMemcpyHostToDevice, st));
        ctx().stats.h2d_bytes += E;
        act = own.p;
    } else order_after_caller();
    DevBuf<int> count(1);
    count.zero();
    PFDR_LAUNCH(merge_kernel<real>, cdiv(E, 256), 256, 0, Eu.p, Ev.p, Cv.p, rX.p, E, tol, eps, act, count.p);
    int h = 0;
    PFDR_CUDA(cudaMemcpyAsync(&h, count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (!on_device) { PFDR_CUDA(cudaMemcpyAsync(active_io, own.p, E, cudaMemcpyDeviceToHost, st)); ctx().stats.d2h_bytes += E; }
    PFDR_CUDA(cudaStreamSynchronize(st));
    if (count_out) *count_out = h;
}

template void cp_reduced_edges<float>(int, int, const int *, const int *, const unsigned char *, const float *, const int *, int, float, int, int *, int *, int *, float *);
template void cp_reduced_edges<double>(int, int, const int *, const int *, const unsigned char *, const double *, const int *, int, double, int, int *, int *, int *, double *);
template void cp_merge<float>(int, int, const int *, const int *, const int *, int, const float *, float, float, unsigned char *, int *);
template void cp_merge<double>(int, int, const int *, const int *, const int *, int, const double *, double, double, unsigned char *, int *);

}  // namespace pfdr
