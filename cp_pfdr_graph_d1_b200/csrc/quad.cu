// quad.cu - PFDR for  1/2||y-Ax||^2 + sum_e La_d1|x_u-x_v| + { sum_v La_l1|x_v| (+ x>=0) | box },
// the B200-native replacement of src/PFDR_graph_quadratic_d1_l1.cpp and
// src/PFDR_graph_quadratic_d1_bounds.cpp of the reference (same iteration, different schedule).
//
// Schedule of one iteration ("trip") on the device:
//   [N>0]  ax_partial -> ax_finish          R = Y - A X            (ref S1  :356-367)
//          atr                              P = 2X - Ga.(-A^t R)   (ref S4+S6 :432-440,:463-464)
//   [N<0]  ata_gemv                         P = 2X - Ga.(AtA X - AtY)          (ref :368-376,:441-445)
//   [Obj]  objective                                              (ref S2  :388-422)
//   iter   ONE fused kernel, one thread per vertex walking that vertex's half-edge row:
//          d1 prox of every incident edge + relaxation (S7 :467-489), new edge copy written for
//          the thread's own half only (the other half is recomputed by the neighbour's thread, the
//          old copies come from the ping buffer), weighted average in the reference's summation
//          order (S8 :492-497), vertex prox (S9 :500-512 / bounds :473-490), next forward point
//          when A is diagonal (S1+S4+S6 of the next trip) and the evolution sums (S10 :515-529).
// Per-edge constants of the reference (Wu, Wv, W_d1u, W_d1v, Th_d1: five E-arrays) are not stored:
// with S[v] = sum of raw weights at v and T[v] = Ga[v].S[v] they are  Wu = w/S[u],
// W_d1u = T[v]/(T[u]+T[v]), W_d1v = T[u]/(T[u]+T[v]), Th_d1 = (La_d1/w).(T[u]+T[v])   (ref :253-261)
// so only the raw weight w is streamed per half-edge.
#include "graph.cuh"
#include "solvers.cuh"
#include "quad.cuh"
#include "dia.cuh"
#include "quad_dia.cuh"
#include <cooperative_groups.h>
#include <chrono>
#include <vector>

namespace cg = cooperative_groups;
namespace pfdr {


// ---------------------------------------------------------------------------------------------
// preconditioning (ref src/PFDR_graph_quadratic_d1_l1.cpp:57-268)
// ---------------------------------------------------------------------------------------------
// Ga0 = diag(A^tA) (:102-122) and the amplitude constant c (:125-154).  recond = 0: over the
// coordinate-wise pseudo-inverse of Y; recond = 1: over the current iterate.
template <typename real>
__global__ void __launch_bounds__(256) amp_kernel(QP<real> p, int recond)
{
    __shared__ double smem[64];
    const int v = p.vbeg + blockIdx.x * blockDim.x + threadIdx.x;       // owned vertices
    double r[2] = {0.0, 0.0};
    if (v < p.vend) {
        real a;
        if (recond) {
            a = ((p.ctl->it & 1) ? p.X1 : p.X0)[v];
        } else {
            real g0;
            if (p.N > 0) { g0 = p.Ga0[v]; a = p.AtY[v] / g0; }           // colstats filled Ga0, AtY
            else {
                g0 = (p.N < 0) ? p.A[((size_t) p.V + 1) * (size_t) v] : (p.A ? p.A[v] : (real) 1);
                p.Ga0[v] = g0;
                a = (g0 > (real) 0) ? p.Y[v] / g0 : (real) 0;
            }
        }
        if (a > (real) 0) { r[0] = (double) a; r[1] = 1.0; }          // ref :149-153
        else if (a < (real) 0) { r[0] = -(double) a; r[1] = 1.0; }
    }
    if (grid_sum<2>(p.ctl, p.partial, r, smem) && threadIdx.x == 0) {
        if (p.sharded) { p.amp_sums[0] = r[0]; p.amp_sums[1] = r[1]; } // finished by amp_finish_kernel after the all-reduce
        else p.ctl->amp = recond ? (double) ((real) r[0] / (real) r[1]) : (double) ((real) r[1] / (real) r[0]);
    }
}

template <typename real>
__global__ void amp_finish_kernel(QP<real> p, int recond)
{
    const double a = p.amp_sums[0], n = p.amp_sums[1];
    p.ctl->amp = recond ? (double) ((real) a / (real) n) : (double) ((real) n / (real) a);
}


// first preconditioning + initial state: w = c La_d1 (:160-166), S, Ga, T, Th_l1, z_e = x (:321-324)
template <typename real>
__global__ void __launch_bounds__(256) init_vertex_kernel(QP<real> p, const real *__restrict__ X0)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.V) return;
    const real c = (real) p.ctl->amp;
    const real x = X0[v];
    const int lane = v & 31;
    const long long base = p.slice_off[v >> 5];
    const int width = (int) ((p.slice_off[(v >> 5) + 1] - base) >> 5);
    real S = 0;
    for (int j = 0; j < width; j++) {
        const long long s = base + (long long) j * 32 + lane;
        const int h = p.hkey[s];
        if (h >= 0) {
            const real wv = mul_rn(c, p.La_d1[h >> 1]);
            p.w[s] = wv; S = add_rn(S, wv); p.Z0[s] = x;
        }
    }
    real ga = add_rn(p.Ga0[v], S);
    if (p.La_l1) ga = fma_rn(c, p.La_l1[v], ga);
    ga = clamp_metric(p, v, (real) 1 / ga);
    p.Ga[v] = ga; p.Sinv[v] = (real) 1 / S; p.T[v] = mul_rn(ga, S);
    if (p.La_l1) p.Thl1[v] = mul_rn(ga, p.La_l1[v]);
    p.X0[v] = x;
    p.P0[v] = (p.N == 0) ? forward_point_of(x, ga, grad_at(p, v, x)) : (real) 0;
}

// reconditioning, everything that is local to a vertex row in one kernel (ref :89-99, :167-250):
// subgradients from the old metric, new raw weights, new metric, new edge copies, new forward point.
template <typename real>
__device__ __forceinline__ void recond_vertex(const QP<real> &p, int v, const real *X, real *Z, real *Pout, real c)
{
    const real x = X[v], g = grad_at(p, v, x);
    const real ga_old = p.Ga[v], sinv_old = p.Sinv[v];
    const int lane = v & 31;
    const long long base = p.slice_off[v >> 5];
    const int width = (int) ((p.slice_off[(v >> 5) + 1] - base) >> 5);
    real S = 0;
    for (int j = 0; j < width; j++) {
        const long long s = base + (long long) j * 32 + lane;
        const int n = p.nbr[s];
        if (n >= 0) {
            const real Wold = mul_rn(p.w[s], sinv_old);
            Z[s] = mul_rn(Wold / ga_old, add_rn(fma_rn(-ga_old, g, x), -Z[s]));   // ref :96-97
            const real xn = X[n];
            real a = rabs(x), b = rabs(xn), d = rabs(x - xn);                   // ref :172-186
            if (a < b) a = b;
            if (a < c) a = c;
            a *= p.condMin;
            if (d < a) d = a;
            const real wv = p.La_d1[p.hkey[s] >> 1] / d;
            p.w[s] = wv; p.thc[s] = d; S = add_rn(S, wv);
        }
    }
    real ga = add_rn(p.Ga0[v], S);
    if (p.La_l1) {                                                             // ref :209-218
        c *= p.condMin;
        real d = rabs(x);
        if (d < c) d = c;
        ga += p.La_l1[v] / d;
    }
    ga = clamp_metric(p, v, (real) 1 / ga);
    const real sinv = (real) 1 / S;
    p.Ga[v] = ga; p.Sinv[v] = sinv; p.T[v] = mul_rn(ga, S);
    if (p.La_l1) p.Thl1[v] = mul_rn(ga, p.La_l1[v]);
    for (int j = 0; j < width; j++) {
        const long long s = base + (long long) j * 32 + lane;
        if (p.nbr[s] >= 0) Z[s] = fma_rn(-ga, add_rn(g, Z[s] / mul_rn(p.w[s], sinv)), x);   // ref :247-248
    }
    Pout[v] = forward_point_of(x, ga, g);
}

template <typename real>
__global__ void __launch_bounds__(256) recond_vertex_kernel(QP<real> p)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int cur = p.ctl->it & 1;
    if (v < p.V) recond_vertex<real>(p, v, cur ? p.X1 : p.X0, cur ? p.Z1 : p.Z0, cur ? p.P1 : p.P0, (real) p.ctl->amp);
}

template <typename real>
__global__ void resume_kernel(QP<real> p, double rcd_factor)
{
    p.ctl->rcd *= rcd_factor; p.ctl->state = ST_RUN; p.ctl->hold = 0;
    p.ctl->scratch[1] = (double) (p.ctl->it + 1);            // slab-sharded: this trip's reconditioning test has been made
}

template <typename real>
__global__ void quad_obj_final_kernel(QP<real> p) { p.dObj[p.ctl->it] = (real) *p.obj_part; }

// ---------------------------------------------------------------------------------------------
// the fused iteration kernel
// ---------------------------------------------------------------------------------------------

// sharded path: the sums were all-reduced over the ranks between the iteration kernel and this one
template <typename real>
__global__ void quad_finalize_kernel(QP<real> p)
{
    if (p.ctl->hold != 0 || p.ctl->state != ST_RUN) return;
    double d = 0.0, c = 0.0;
    for (int r = 0; r < p.world; r++) { d += p.sums[2 * r]; c += p.sums[2 * r + 1]; }     // rank order: same bits everywhere
    quad_finalize(p, p.ctl->it, d, c);
}

// The iteration-invariant per-slot data side by side, 16 bytes per slot: float32 {nbr, twin, w, La_d1/w}, float64
// {nbr, twin, w}.  One 128-bit load per row replaces three or four 32/64-bit ones (fewer instructions in a kernel bound
// by its dependent instruction stream, fewer concurrent DRAM streams): C2 f32 51.2 -> 46.4 us per iteration with the
// index pair alone (profiles/r01_notes.md).  Rebuilt after every (re)conditioning; the other kernels keep the plain arrays.
template <typename real>
__global__ void pack_slots_kernel(QP<real> p, int4 *__restrict__ pk, long long nslots)
{
    const long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const real w = p.w[s];
    int4 r;
    r.x = p.nbr[s]; r.y = p.twin[s];
    if (pk == nullptr) { const_cast<int2 *>(p.pk2)[s] = make_int2(r.x, r.y); return; }
    if (sizeof(real) == 4) { r.z = __float_as_int((float) w); r.w = __float_as_int((float) (p.thc ? p.thc[s] : p.thc0)); }
    else { r.z = __double2loint((double) w); r.w = __double2hiint((double) w); }
    pk[s] = r;
}
template <typename real>
__device__ __forceinline__ void unpack_slot(const int4 r, int &n, int &tw, real &w, real &th)
{
    n = r.x; tw = r.y;
    if (sizeof(real) == 4) { w = (real) __int_as_float(r.z); th = (real) __int_as_float(r.w); }
    else w = (real) __hiloint2double(r.w, r.z);
}

// Pipelined variant: one row at a time, the next row's indices / weight / own copy are requested while the
// current row's gathers are in flight (as simplex_iter_kernel).  Few registers -> up to 64 warps per SM.
template <typename real, bool PER_SLOT_TH, int MINB, bool HUB, int NT = 0>
__global__ void __launch_bounds__(256, MINB) quad_iter_pipe_kernel(QP<real> p)
{
    __shared__ double smem[64];
    Ctl *ctl = p.ctl;
    const CtlView cv = load_ctl(ctl);
    if (cv.hold != 0 || cv.state != ST_RUN) {
        if (cv.hold == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctl->hold = 1;
        return;
    }
    const int it = cv.it, cur = it & 1;
    const real *__restrict__ Zin = cur ? p.Z1 : p.Z0;
    real *__restrict__ Zout = cur ? p.Z0 : p.Z1;
    const real *__restrict__ Pc = cur ? p.P1 : p.P0;
    const real *__restrict__ Tg = p.T;
    const int *__restrict__ nbr = p.nbr;
    const int *__restrict__ twin = p.twin;
    const real *__restrict__ wg = p.w;
    const real rho = p.rho;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    real red[2] = {(real) 0, (real) 0};
    const bool own = v < p.Vown;
    long long base = 0;
    int width = 0;                                       // of the warp's slice (warp-uniform)
    if ((v & ~31) < p.Vown) {
        base = p.slice_off[v >> 5];
        width = (int) ((p.slice_off[(v >> 5) + 1] - base) >> 5);
    }
    real xo = 0, po_v = 0, Tv = 1, Sv = 0, acc = 0;
    if (own) {
        const int wl = HUB ? min(width, HUB_ROWS) : width;   // HUB: rows past HUB_ROWS are shared by the warp, below
        int s = (int) base + (v & 31);
        xo = (cur ? p.X1 : p.X0)[v]; po_v = Pc[v];
        Tv = Tg[v]; Sv = p.Sinv[v];
        int n_c = -1, tw_c = 0;
        real w_c = 0, zo_c = 0, th_c = p.thc0;
        if (wl > 0) {
            if (NT == 2) {
                real th_pk = th_c;
                unpack_slot<real>(__ldg(p.pk + s), n_c, tw_c, w_c, th_pk);
                if (PER_SLOT_TH) th_c = sizeof(real) == 4 ? th_pk : p.thc[s];
            } else if (NT == 1) { const int2 t = __ldg(p.pk2 + s); n_c = t.x; tw_c = t.y; w_c = wg[s]; if (PER_SLOT_TH) th_c = p.thc[s]; }
            else { n_c = nbr[s]; tw_c = twin[s]; w_c = wg[s]; if (PER_SLOT_TH) th_c = p.thc[s]; }
            zo_c = Zin[s];
        }
        for (int j = 0; j < wl; j++, s += 32) {
            const bool live = n_c >= 0;
            real pn = 0, Tn = 1, zt = 0;
            if (live) { pn = Pc[n_c]; Tn = Tg[n_c]; zt = Zin[tw_c]; }
            int n_n = -1, tw_n = 0;
            real w_n = 0, zo_n = 0, th_n = p.thc0;
            if (j + 1 < wl) {
                if (NT == 2) {
                    real th_pk = th_n;
                    unpack_slot<real>(__ldg(p.pk + s + 32), n_n, tw_n, w_n, th_pk);
                    if (PER_SLOT_TH) th_n = sizeof(real) == 4 ? th_pk : p.thc[s + 32];
                } else if (NT == 1) {
                    const int2 t = __ldg(p.pk2 + s + 32); n_n = t.x; tw_n = t.y; w_n = wg[s + 32];
                    if (PER_SLOT_TH) th_n = p.thc[s + 32];
                } else { n_n = nbr[s + 32]; tw_n = twin[s + 32]; w_n = wg[s + 32]; if (PER_SLOT_TH) th_n = p.thc[s + 32]; }
                zo_n = Zin[s + 32];
            }
            if (live) {
                const real zn = d1_prox_half<real>(rho, xo, po_v, Tv, zo_c, zt, pn, Tn, th_c);
                Zout[s] = zn;
                acc = fma_rn(mul_rn(w_c, Sv), zn, acc);
            }
            n_c = n_n; tw_c = tw_n; w_c = w_n; zo_c = zo_n; th_c = th_n;
        }
    }
    if (HUB && width > HUB_ROWS) {
        // A hub (the background component of a cut-pursuit reduced graph touches thousands of others): one thread
        // walking its row serially would hold the whole launch.  The warp takes the rows past HUB_ROWS of every hub
        // of its slice 32 at a time, every lane doing the full prox of one row; partial sums folded by shuffles.
        // (HUB variant: launched only when the graph has such a slice, with a larger register budget.)
        const int lane = threadIdx.x & 31;
        unsigned m = __ballot_sync(0xffffffffu, own && nbr[(int) base + HUB_ROWS * 32 + lane] >= 0);
        while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
            const real xo_c = __shfl_sync(0xffffffffu, xo, c), po_c = __shfl_sync(0xffffffffu, po_v, c);
            const real Tv_c = __shfl_sync(0xffffffffu, Tv, c), Sv_c = __shfl_sync(0xffffffffu, Sv, c);
            real part = 0;
            constexpr int HU = 4;                                             // rows in flight per lane
            for (int j0 = HUB_ROWS + lane; j0 < width; j0 += 32 * HU) {
                int sk[HU], n[HU], tw[HU];
                real zo[HU], ww[HU], thc[HU], zt[HU], Tn[HU], pn[HU];
#pragma unroll
                for (int u = 0; u < HU; u++) {
                    const int j = j0 + 32 * u;
                    sk[u] = (int) base + j * 32 + c;
                    n[u] = j < width ? nbr[sk[u]] : -1;
                }
#pragma unroll
                for (int u = 0; u < HU; u++) {
                    const bool live = n[u] >= 0;
                    tw[u] = live ? twin[sk[u]] : 0; zo[u] = live ? Zin[sk[u]] : (real) 0; ww[u] = live ? wg[sk[u]] : (real) 0;
                    thc[u] = (PER_SLOT_TH && live) ? p.thc[sk[u]] : p.thc0;
                    Tn[u] = live ? Tg[n[u]] : (real) 1; pn[u] = live ? Pc[n[u]] : (real) 0;
                }
#pragma unroll
                for (int u = 0; u < HU; u++) zt[u] = n[u] >= 0 ? Zin[tw[u]] : (real) 0;
#pragma unroll
                for (int u = 0; u < HU; u++) {
                    if (n[u] >= 0) {
                        const real zn = d1_prox_half<real>(rho, xo_c, po_c, Tv_c, zo[u], zt[u], pn[u], Tn[u], thc[u]);
                        Zout[sk[u]] = zn;
                        part = fma_rn(mul_rn(ww[u], Sv_c), zn, part);
                    }
                }
            }
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == c) acc += part;
        }
    }
    if (own) {
        real x = acc;
        if (p.prox_kind == 0) {                                                // ref l1.cpp :500-512
            if (p.La_l1) {
                const real th = p.Thl1[v];
                if (x > th) x -= th;
                else if (!p.positivity && x < -th) x += th;
                else x = 0;
            } else if (p.positivity) { if (x < (real) 0) x = 0; }
        } else {                                                               // ref bounds.cpp :473-490
            if (p.has_lo && x < p.lo) x = p.lo;
            else if (p.has_hi && x > p.hi) x = p.hi;
        }
        (cur ? p.X0 : p.X1)[v] = x;
        if (p.N == 0) (cur ? p.P0 : p.P1)[v] = forward_point_diag(p, v, x);
        const real dx = xo - x;
        red[0] = dx * dx; red[1] = x * x;
    }
    double tot[2];
    if (grid_sum_fast<2, real>(ctl, p.partial, red, tot, smem) && threadIdx.x == 0) {
        if (p.sharded) { p.sums[2 * p.rank] = tot[0]; p.sums[2 * p.rank + 1] = tot[1]; }
        else quad_finalize(p, it, tot[0], tot[1]);
    }
}

// ---------------------------------------------------------------------------------------------
// Small-problem mode: the whole solve in ONE launch of ONE thread-block cluster (up to 16 CTAs x 512 threads).
// The reduced problems that the cut-pursuit outer loop hands over (a few hundred to a few thousand vertices, up to
// 1e4 iterations, reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:848-859) are latency bound on the large path: one
// launch (or four) per iteration and, inside it, a chain of dependent L2 round trips per half-edge row - with a
// hub vertex (the background component of a reduced graph touches most others) the chain is its whole degree.
//   * an iteration is two to four hardware cluster barriers (~0.2 us each) instead of kernel boundaries;
//   * each CTA owns a range of SELL slices (vertices) and a range of edges;
//   * the vertex arrays every edge gathers (X, forward point P, T) are replicated in each CTA's SHARED memory and
//     refreshed from global memory (L2) with coalesced loads after the barrier that publishes them;
//   * the edge prox is edge-parallel: one thread computes BOTH copies of an edge (same arithmetic as the two
//     half-edge threads of the large path, bit for bit), reads the old copies from an edge-ordered array (coalesced)
//     and scatters the new ones to their SELL slots, where the owner of the vertex reads its rows coalesced;
//   * dense operators are applied by column slabs, the partial A x of the CTAs is exchanged through L2;
//   * stopping rule, reconditioning and objective are evaluated on the device, identically in every CTA
//     (partial sums exchanged through L2 and added in rank order).
// Everything crossing CTAs goes through L2 (ld.global.cg) between cluster barriers (release / acquire).
// ---------------------------------------------------------------------------------------------
#ifdef PFDR_SMALL_PROF
#define SPROF_DECL long long spt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, spc = clock64();
#define SPROF(k) { const long long t_ = clock64(); spt[k] += t_ - spc; spc = t_; }
#define SPROF_OUT if (rank == 0 && tid == 0) { for (int k = 0; k < 8; k++) printf("phase %d: %.3f us/it\n", k, (double) spt[k] / 1.9e3 / (it > 0 ? it : 1)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GSTAMP(k) if (tid == 0 && it == itMax - 2) gst[k] = gtime();
#define GSTAMP_DECL unsigned long long gst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define GSTAMP_OUT if (tid == 0) printf("rank %d: edge0 %llu zpush %llu zwait %llu vertex %llu pub %llu xwait %llu\n", rank, gst[0] % 1000000, gst[1] % 1000000, gst[2] % 1000000, gst[3] % 1000000, gst[4] % 1000000, gst[5] % 1000000);
#else
#define SPROF_DECL
#define SPROF(k)
#define SPROF_OUT
#define GSTAMP(k)
#define GSTAMP_DECL
#define GSTAMP_OUT
#endif

constexpr int SMALL_MAX_SLICES = 1024;   // slices a CTA may own (32768 vertices)
constexpr int SMALL_MAX_WIDE = 128;      // wide slices a CTA handles cooperatively; further ones take the per-thread path
constexpr int SMALL_WIDE = 24;           // rows from which a slice counts as wide

template <typename real> __device__ __forceinline__ void ld2(const real *q, real &a, real &b);
template <> __device__ __forceinline__ void ld2<float>(const float *q, float &a, float &b) { const float2 t = *reinterpret_cast<const float2 *>(q); a = t.x; b = t.y; }
template <> __device__ __forceinline__ void ld2<double>(const double *q, double &a, double &b) { const double2 t = *reinterpret_cast<const double2 *>(q); a = t.x; b = t.y; }
template <typename real> __device__ __forceinline__ void st2(real *q, real a, real b);
template <> __device__ __forceinline__ void st2<float>(float *q, float a, float b) { *reinterpret_cast<float2 *>(q) = make_float2(a, b); }
template <> __device__ __forceinline__ void st2<double>(double *q, double a, double b) { *reinterpret_cast<double2 *>(q) = make_double2(a, b); }

// Cluster barrier of the iteration loop: release on arrive (the CTA's global writes are performed before any CTA
// passes), NO acquire on wait - an acquiring wait invalidates the whole L1 (CCTL.IVALL), and then every constant of
// the iteration (indices, weights, metric) is fetched from L2 again, iteration after iteration.  Data written by
// another CTA is only ever read with ld.global.cg (L2, never L1), so nothing stale can be observed.
__device__ __forceinline__ void cluster_bar()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.aligned;\n" ::: "memory");
}

template <typename real>
struct SmallArgs {
    int nslots, nslices;
    int *pos;          // [2E] SELL slot of every half-edge key
    real *Ze;          // [2E] edge copies in edge order
    double *xs;        // [2][16][2] partial sums of the CTAs, double-buffered
    real *parts;       // [16][N] partial A x of the CTAs
};

template <typename real>
__device__ __forceinline__ void small_block_sum2(double &a, double &b, double *sm /* 64 doubles */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum(a); b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { sm[wid] = a; sm[32 + wid] = b; }
    __syncthreads();
    if (wid == 0) {
        double x = lane < nw ? sm[lane] : 0.0, y = lane < nw ? sm[32 + lane] : 0.0;
        x = warp_sum(x); y = warp_sum(y);
        if (lane == 0) { sm[0] = x; sm[32] = y; }
    }
    __syncthreads();
    a = sm[0]; b = sm[32];
    __syncthreads();
}

// sum of (a, b) over the CTAs of the cluster, the same bits in every CTA; contains one cluster barrier, which also
// publishes every global write made before the call
template <typename real>
__device__ __forceinline__ void cluster_sum2(cg::cluster_group &cl, double *xs, int &ph, int rank, int C, double &a, double &b,
                                             double *sm)
{
    small_block_sum2<real>(a, b, sm);
    if (threadIdx.x == 0) { xs[(ph * 16 + rank) * 2] = a; xs[(ph * 16 + rank) * 2 + 1] = b; }
    cl.sync();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double x = lane < C ? __ldcg(xs + (ph * 16 + lane) * 2) : 0.0, y = lane < C ? __ldcg(xs + (ph * 16 + lane) * 2 + 1) : 0.0;
        x = warp_sum(x); y = warp_sum(y);
        if (lane == 0) { sm[0] = x; sm[32] = y; }
    }
    __syncthreads();
    a = sm[0]; b = sm[32];
    __syncthreads();
    ph ^= 1;
}

template <typename real>
__global__ void __launch_bounds__(512, 1) quad_small_kernel(QP<real> p, SmallArgs<real> a)
{
    extern __shared__ double dsm_small[];
    cg::cluster_group cl = cg::this_cluster();
    Ctl *ctl = p.ctl;
    const int C = (int) gridDim.x, rank = (int) blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int V = p.V, N = p.N, E = p.E, itMax = ctl->itMax, track = ctl->track;
    const int Np = N > 0 ? N : 0, Vp = (V + 3) & ~3;
    double *red = dsm_small;                                     // [64] reductions
    real *Xs = reinterpret_cast<real *>(dsm_small + 64);         // [V] iterate         } replicas of the arrays the edges gather
    real *Ps = Xs + Vp;                                          // [V] forward point   }
    real *Ts = Ps + Vp;                                          // [V] Ga.S            }
    real *Rs = Ts + Vp;                                          // [N] residual
    real *part = Rs + ((Np + 3) & ~3);                           // [nw * N] per-warp partial A x
    __shared__ int sbase[SMALL_MAX_SLICES + 1];                  // slot offset of the own slices
    __shared__ unsigned char isw[SMALL_MAX_SLICES];              // own slice is wide (handled by the whole CTA)
    __shared__ int widelist[SMALL_MAX_WIDE], nwide;
    __shared__ real wpart[32 * 32];
    constexpr int EU = 2, VU = 8;
    // ownership: SELL slices (vertices, slots) and edges in contiguous ranges
    const int sl0 = (int) ((long long) a.nslices * rank / C), sl1 = (int) ((long long) a.nslices * (rank + 1) / C);
    const int v0 = sl0 * 32, v1 = min(V, sl1 * 32);
    const int s0 = (int) p.slice_off[sl0], s1 = (int) p.slice_off[sl1];
    const int e0 = (int) ((long long) E * rank / C), e1 = (int) ((long long) E * (rank + 1) / C);
    real *Xg = p.X0, *Pg = p.P0, *Zg = p.Z0;                     // one buffer each in this mode
    int it = ctl->it;
    if (it & 1) {                                                // resumed at an odd iteration: state sits in the other buffers
        for (int v = v0 + tid; v < v1; v += nt) { Xg[v] = p.X1[v]; Pg[v] = p.P1[v]; }
        for (int s = s0 + tid; s < s1; s += nt) Zg[s] = p.Z1[s];
    }
    for (int s = s0 + tid; s < s1; s += nt) { const int h = p.hkey[s]; if (h >= 0) a.pos[h] = s; }
    for (int q = tid; q <= sl1 - sl0; q += nt) sbase[q] = (int) p.slice_off[sl0 + q];
    __syncthreads();
    if (tid == 0) {                                               // in slice order: deterministic
        int n = 0;
        for (int q = 0; q < sl1 - sl0; q++) {
            const bool wide = ((sbase[q + 1] - sbase[q]) >> 5) > SMALL_WIDE && n < SMALL_MAX_WIDE;
            isw[q] = wide;
            if (wide) widelist[n++] = sl0 + q;
        }
        nwide = n;
    }
    cl.sync();
    for (int h = 2 * e0 + tid; h < 2 * e1; h += nt) a.Ze[h] = __ldcg(Zg + __ldcg(a.pos + h));
    for (int v = tid; v < V; v += nt) { Xs[v] = __ldcg(Xg + v); Ps[v] = __ldcg(Pg + v); Ts[v] = __ldcg(p.T + v); }
    const real tol = (real) ctl->tol, rho = p.rho;
    const double eps = ctl->eps;
    double dif = ctl->dif, rcd = ctl->rcd;
    bool per_slot = false;                           // thresholds per half-edge once reconditioned
    int ph = 0;
    const int *__restrict__ nbr = p.nbr;
    __syncthreads();
    SPROF_DECL
    for (;;) {
        const bool stop = (it == itMax || (real) dif < tol);                  // ref :429 (uniform over the cluster)
        if (stop && !p.dObj) break;
        SPROF(0)
        // ---- forward point for dense operators (ref S1, S4, S6), own columns ----
        if (N > 0) {
            // partial A X: warp w walks the columns v0 + w, v0 + w + nw, ...; lanes along n
            for (int n0 = 0; n0 < N; n0 += 32) {
                const int n = n0 + lane;
                if (n < N) {
                    real acc = 0;
#pragma unroll 4
                    for (int v = v0 + wid; v < v1; v += nw) acc += __ldg(p.A + (size_t) N * v + n) * Xs[v];
                    part[(size_t) wid * N + n] = acc;
                }
            }
            __syncthreads();
            for (int n = tid; n < N; n += nt) {
                real acc = 0;
                for (int w = 0; w < nw; w++) acc += part[(size_t) w * N + n];
                a.parts[(size_t) rank * N + n] = acc;
            }
            cluster_bar();
            for (int n = tid; n < N; n += nt) {
                real acc = 0;
                for (int q = 0; q < C; q++) acc += __ldcg(a.parts + (size_t) q * N + n);      // rank order: same bits everywhere
                Rs[n] = p.Y[n] - acc;
            }
            __syncthreads();
            // G = -A^t R: one warp per column, lanes along n
            for (int v = v0 + wid; v < v1; v += nw) {
                const real *col = p.A + (size_t) N * v;
                real acc = 0;
                for (int n = lane; n < N; n += 32) acc += __ldg(col + n) * Rs[n];
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) { p.G[v] = -acc; Pg[v] = (real) 2 * Xs[v] - p.Ga[v] * (-acc); }
            }
        } else if (N < 0) {
            // G = (A^tA) X: one warp per column;  P = 2X - Ga.(G - A^tY)
            for (int v = v0 + wid; v < v1; v += nw) {
                const real *col = p.A + (size_t) V * v;
                real acc = 0;
#pragma unroll 4
                for (int u = lane; u < V; u += 32) acc += __ldg(col + u) * Xs[u];
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) { p.G[v] = acc; Pg[v] = (real) 2 * Xs[v] - p.Ga[v] * (acc - p.Y[v]); }
            }
        }
        SPROF(1)
        if (N != 0) {
            cluster_bar();
            for (int v = tid; v < V; v += nt) Ps[v] = __ldcg(Pg + v);
            __syncthreads();
        }
        SPROF(2)
        // ---- objective of the current iterate (ref :388-422) ----
        if (p.dObj) {
            double o = 0.0, zero = 0.0;
            for (int e = e0 + tid; e < e1; e += nt) o += (double) p.La_d1[e] * (double) rabs(Xs[p.Eu[e]] - Xs[p.Ev[e]]);
            for (int v = v0 + tid; v < v1; v += nt) {
                const real x = Xs[v];
                if (N == 0) o += (double) x * ((double) 0.5 * (double) (p.A ? p.A[v] * x : x) - (double) p.Y[v]);
                else if (N < 0) o += (double) x * ((double) 0.5 * (double) p.G[v] - (double) p.Y[v]);
                if (p.La_l1) o += (double) p.La_l1[v] * (double) rabs(x);
            }
            if (N > 0 && rank == 0) for (int n = tid; n < N; n += nt) o += 0.5 * (double) Rs[n] * (double) Rs[n];
            cluster_sum2<real>(cl, a.xs, ph, rank, C, o, zero, red);
            if (rank == 0 && tid == 0) p.dObj[it] = (real) o;
            if (stop) break;
        }
        // ---- reconditioning (ref :448-459): new amplitude constant, new metric, edge copies rebuilt ----
        if ((real) dif < (real) rcd) {
            double sa = 0.0, sn = 0.0;
            for (int v = v0 + tid; v < v1; v += nt) {
                const real x = Xs[v];
                if (x != (real) 0) { sa += (double) rabs(x); sn += 1.0; }
            }
            cluster_sum2<real>(cl, a.xs, ph, rank, C, sa, sn, red);
            const real c = (real) sa / (real) sn;
            for (int v = v0 + tid; v < v1; v += nt) recond_vertex<real>(p, v, Xs, Zg, Pg, c);
            per_slot = true;
            rcd *= 0.01;
            cl.sync();
            for (int h = 2 * e0 + tid; h < 2 * e1; h += nt) a.Ze[h] = __ldcg(Zg + a.pos[h]);
            for (int v = tid; v < V; v += nt) { Ps[v] = __ldcg(Pg + v); Ts[v] = __ldcg(p.T + v); }
            __syncthreads();
        }
        // ---- edge prox + relaxation (ref S7, S8): both copies of an edge by one thread, EU edges in flight ----
        for (int eb = e0 + tid; eb < e1; eb += EU * nt) {
            int su[EU], sv[EU], u[EU], v[EU];
            real zu[EU], zv[EU], thc[EU];
#pragma unroll
            for (int k = 0; k < EU; k++) {
                const int e = eb + k * nt;
                const bool in = e < e1;
                const int2 ss = in ? *reinterpret_cast<const int2 *>(a.pos + 2 * e) : make_int2(0, 0);
                su[k] = ss.x; sv[k] = ss.y;
                u[k] = in ? __ldg(p.Eu + e) : 0; v[k] = in ? __ldg(p.Ev + e) : 0;
                if (in) ld2(a.Ze + 2 * e, zu[k], zv[k]); else { zu[k] = 0; zv[k] = 0; }
                thc[k] = (per_slot && in) ? p.thc[su[k]] : p.thc0;
            }
#pragma unroll
            for (int k = 0; k < EU; k++) {
                const int e = eb + k * nt;
                if (e < e1) {
                    const real Tu = Ts[u[k]], Tv = Ts[v[k]];
                    const real tsum = Tu + Tv, inv = fast_rcp(tsum), th = thc[k] * tsum;
                    const real wu = Tv * inv, wv = Tu * inv;     // W_d1 on the u side / on the v side
                    const real pu = Ps[u[k]] - zu[k], pv = Ps[v[k]] - zv[k], d0 = pu - pv;
                    const real b = d0 - min(max(d0, -th), th);   // soft threshold, exact 0 inside
                    const real nu = zu[k] + rho * (fma(wu, d0, pv) + wv * b - Xs[u[k]]);
                    const real nv = zv[k] + rho * (fma(wv, -d0, pu) + wu * (-b) - Xs[v[k]]);
                    st2(a.Ze + 2 * e, nu, nv);
                    Zg[su[k]] = nu; Zg[sv[k]] = nv;
                }
            }
        }
        SPROF(3)
        cluster_bar();
        SPROF(4)
        // ---- average + vertex prox + evolution (ref S9, S10), own vertices ----
        double sd = 0.0, sc = 0.0;
        auto finish_vertex = [&](int v, real x, real xo) {
            if (p.prox_kind == 0) {
                if (p.La_l1) {
                    const real th = p.Thl1[v];
                    if (x > th) x -= th;
                    else if (!p.positivity && x < -th) x += th;
                    else x = 0;
                } else if (p.positivity) { if (x < (real) 0) x = 0; }
            } else {
                if (p.has_lo && x < p.lo) x = p.lo;
                else if (p.has_hi && x > p.hi) x = p.hi;
            }
            Xg[v] = x;
            if (N == 0) Pg[v] = (real) 2 * x - p.Ga[v] * grad_at(p, v, x);
            const real dx = xo - x;
            sd += (double) (dx * dx); sc += (double) (x * x);
        };
        // (a) one thread per vertex, VU rows of its SELL column in flight, summed in the reference's order
        for (int v = v0 + tid; v < v1; v += nt) {
            const int sl = v >> 5;
            if (isw[sl - sl0]) continue;                 // wide slice (a hub): all warps of the CTA, below
            const int base = sbase[sl - sl0], width = (sbase[sl - sl0 + 1] - base) >> 5;
            const int s = base + (v & 31);
            const real Sv = p.Sinv[v];
            real acc = 0;
            for (int j0 = 0; j0 < width; j0 += VU) {
                int nb[VU];
                real ww[VU], zz[VU];
#pragma unroll
                for (int k = 0; k < VU; k++) {
                    const bool in = j0 + k < width;
                    const int sk = s + (j0 + k) * 32;
                    nb[k] = in ? nbr[sk] : -1; ww[k] = in ? p.w[sk] : (real) 0; zz[k] = in ? __ldcg(Zg + sk) : (real) 0;
                }
#pragma unroll
                for (int k = 0; k < VU; k++) if (nb[k] >= 0) acc += (ww[k] * Sv) * zz[k];
            }
            finish_vertex(v, acc, Xs[v]);
        }
        // (b) wide slices: warp w takes rows w, w + nw, ... of the slice (lanes = its 32 vertices), partial sums
        //     folded through shared memory in warp order
        for (int q = 0; q < nwide; q++) {
            const int sl = widelist[q];
            const int base = sbase[sl - sl0], width = (sbase[sl - sl0 + 1] - base) >> 5;
            const int v = sl * 32 + lane;
            const real Sv = v < V ? p.Sinv[v] : (real) 0;
            real acc = 0;
            for (int j0 = wid; j0 < width; j0 += VU * nw) {
                int nb[VU];
                real ww[VU], zz[VU];
#pragma unroll
                for (int k = 0; k < VU; k++) {
                    const int j = j0 + k * nw;
                    const bool in = j < width;
                    const int sk = base + j * 32 + lane;
                    nb[k] = in ? nbr[sk] : -1; ww[k] = in ? p.w[sk] : (real) 0; zz[k] = in ? __ldcg(Zg + sk) : (real) 0;
                }
#pragma unroll
                for (int k = 0; k < VU; k++) if (nb[k] >= 0) acc += (ww[k] * Sv) * zz[k];
            }
            wpart[wid * 32 + lane] = acc;
            __syncthreads();
            if (wid == 0 && v < V) {
                real t = 0;
                for (int w = 0; w < nw; w++) t += wpart[w * 32 + lane];
                finish_vertex(v, t, Xs[v]);
            }
            __syncthreads();
        }
        SPROF(5)
        // ---- evolution sums over the cluster; the barrier publishes Xg / Pg; one L2 round trip refreshes the replicas ----
        small_block_sum2<real>(sd, sc, red);
        if (tid == 0) { a.xs[(ph * 16 + rank) * 2] = sd; a.xs[(ph * 16 + rank) * 2 + 1] = sc; }
        cluster_bar();
        SPROF(6)
        {
            double x = 0.0, y = 0.0;
            if (tid < 32 && lane < C) { x = __ldcg(a.xs + (ph * 16 + lane) * 2); y = __ldcg(a.xs + (ph * 16 + lane) * 2 + 1); }
            for (int v = tid; v < V; v += nt) { Xs[v] = __ldcg(Xg + v); if (N == 0) Ps[v] = __ldcg(Pg + v); }
            if (tid < 32) {
                x = warp_sum(x); y = warp_sum(y);
                if (lane == 0) { red[0] = x; red[32] = y; }
            }
            __syncthreads();
            sd = red[0]; sc = red[32];
            ph ^= 1;
        }
        if (track) {
            dif = (sc > eps) ? sd / sc : sd / eps;
            if (rank == 0 && tid == 0 && p.dDif) p.dDif[it] = (real) dif;
        }
        it++;
        SPROF(7)
    }
    SPROF_OUT
    // the large path's readers (extract_x_kernel, a later run) look at the buffers of parity it & 1
    for (int v = v0 + tid; v < v1; v += nt) { p.X1[v] = Xs[v]; p.P1[v] = Ps[v]; }
    for (int s = s0 + tid; s < s1; s += nt) p.Z1[s] = __ldcg(Zg + s);
    if (rank == 0 && tid == 0) { ctl->it = it; ctl->dif = dif; ctl->rcd = rcd; ctl->state = ST_FINAL; ctl->hold = 1; }
}

// ---------------------------------------------------------------------------------------------
// Small-problem mode, push variant (the default when neither the objective record nor reconditioning is asked
// for - the way cut-pursuit calls PFDR): same ownership and arithmetic as quad_small_kernel, but NOTHING crosses
// CTAs through global memory inside the loop.  A cluster barrier with release/acquire costs a gpu-scope MEMBAR
// plus an invalidation of the whole L1 (CCTL.IVALL), i.e. every constant of the iteration comes from L2 again.
// Here every exchange is a push into the consumer's SHARED memory that completes a transaction count on the
// consumer's mbarrier:
//   * new edge copies:  st.async (4/8 bytes) to the SELL slot in the shared memory of the CTA that owns the vertex;
//   * new iterate / forward point of the own vertices, partial A x: one bulk copy (cp.async.bulk shared::cta ->
//     shared::cluster) per destination CTA;
//   * evolution sums: one 16-byte st.async per destination CTA.
// Consumers wait on their own mbarrier (try_wait.parity) - no fence, no L1 invalidation, constants stay in L1.
// Replicas of X and P and the SELL-ordered edge copies are double-buffered by iteration parity (a CTA may be one
// phase ahead of another); the old copy of an edge lives in an edge-ordered array private to the edge's owner.
// ---------------------------------------------------------------------------------------------
template <typename real>
struct PushArgs {
    int nslices, nso;  // nso: shared-memory slots per CTA for the own SELL rows (max over CTAs, multiple of 4)
    int slb[17];       // slice range of CTA c: [slb[c], slb[c+1])
    int *dst;          // [2E] (owner CTA << 24 | slot index inside the owner's range) of every half-edge key
    real *Ze;          // [2E] edge copies in edge order
    real *wS;          // [nslots] w / S of the slot's vertex, 0 on padding
};

__device__ __forceinline__ uint32_t smem_u32(const void *q) { return (uint32_t) __cvta_generic_to_shared(q); }
__device__ __forceinline__ uint32_t map_rank(uint32_t addr, int rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, int &phase)
{
    uint32_t ok = 0;
    for (long long spin = 0; !ok; spin++) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(phase) : "memory");
        if (spin > (1ll << 24)) __trap();            // seconds: a lost message must not hang the device
    }
    phase ^= 1;
}
__device__ __forceinline__ void push_val(uint32_t raddr, float v, uint32_t rmbar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void push_val(uint32_t raddr, double v, uint32_t rmbar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void push_pair(uint32_t raddr, double a, double b, uint32_t rmbar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "l"(__double_as_longlong(a)),
                 "l"(__double_as_longlong(b)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void push_bulk(uint32_t raddr, uint32_t laddr, uint32_t bytes, uint32_t rmbar)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(raddr), "r"(laddr), "r"(bytes),
                 "r"(rmbar) : "memory");
}
// 16 bytes of the own shared memory (src, 16-byte aligned) to the same offset in every other CTA of the cluster
__device__ __forceinline__ void push_chunks(const void *src, uint32_t bytes, int rank, int C, uint32_t mbar)
{
    const uint32_t base = smem_u32(src), nch = bytes >> 4, total = nch * (uint32_t) (C - 1);
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint32_t d = i / nch, ch = i - d * nch;
        const int dest = (int) d < rank ? (int) d : (int) d + 1;
        const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(src) + 16 * ch);
        asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(map_rank(base + 16 * ch, dest)),
                     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(map_rank(mbar, dest)) : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename real>
__global__ void __launch_bounds__(512, 1) quad_push_kernel(QP<real> p, PushArgs<real> a)
{
    extern __shared__ __align__(16) unsigned char push_smem[];
    cg::cluster_group cl = cg::this_cluster();
    Ctl *ctl = p.ctl;
    const int C = (int) gridDim.x, rank = (int) blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int V = p.V, N = p.N, E = p.E, itMax = ctl->itMax, track = ctl->track;
    const int Vq = a.nslices * 32, Nq = N > 0 ? ((N + 3) & ~3) : 0, nso = a.nso;
    // mbarriers: 0 Z, 1 / 4 X (+P, sums, partial A x) of even / odd iterations, 2 partial A x (first iteration), 3 P (dense).
    // A CTA with no edge into a slow CTA's vertices can run one iteration ahead of it: what it pushes for iteration
    // it + 1 must neither be counted in the slow CTA's still open phase of iteration it (two X barriers, by parity)
    // nor overwrite sums it has not read yet (two sets of sums).
    unsigned long long *mbars = reinterpret_cast<unsigned long long *>(push_smem);
    double *sums = reinterpret_cast<double *>(push_smem + 64);   // [2][16][2] evolution sums of the CTAs, by iteration parity
    double *red = sums + 64;                                     // [64]
    real *Xs = reinterpret_cast<real *>(red + 64);               // [2][Vq] iterate         } replicas, by iteration parity
    real *Ps = Xs + 2 * Vq;                                      // [2][Vq] forward point   }
    real *Ts = Ps + 2 * Vq;                                      // [Vq]    Ga.S
    real *Zs = Ts + Vq;                                          // [2][nso] new edge copies of the own SELL rows
    real *partsS = Zs + 2 * nso;                                 // [16][Nq] partial A x of the CTAs
    real *Rs = partsS + 16 * Nq;                                 // [Nq] residual
    __shared__ int sbase[SMALL_MAX_SLICES + 1];
    __shared__ unsigned char isw[SMALL_MAX_SLICES];
    __shared__ int widelist[SMALL_MAX_WIDE], nwide, s0r[17];
    __shared__ real wpart[32 * 32];
    constexpr int EU = 2;
    const int sl0 = a.slb[rank], sl1 = a.slb[rank + 1];
    const int v0 = sl0 * 32, v1 = min(V, sl1 * 32), v1q = sl1 * 32;
    const int e0 = (int) ((long long) E * rank / C), e1 = (int) ((long long) E * (rank + 1) / C);
    const uint32_t mb0 = smem_u32(mbars), mb2 = mb0 + 16, mb3 = mb0 + 24;
    int it = ctl->it;
    if (tid <= C) s0r[tid] = (int) p.slice_off[a.slb[tid]];
    for (int q = tid; q <= sl1 - sl0; q += nt) sbase[q] = (int) p.slice_off[sl0 + q];
    if (tid == 0) {
        for (int k = 0; k < 5; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb0 + 8 * k), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int s0 = s0r[rank], s1 = s0r[rank + 1];
    if (tid == 0) {                                               // in slice order: deterministic
        int n = 0;
        for (int q = 0; q < sl1 - sl0; q++) {
            const bool wide = ((sbase[q + 1] - sbase[q]) >> 5) > SMALL_WIDE && n < SMALL_MAX_WIDE;
            isw[q] = wide;
            if (wide) widelist[n++] = sl0 + q;
        }
        nwide = n;
    }
    {
        const int cur = it & 1;                                  // this mode keeps its global state in the buffers of parity 0
        if (cur) {
            for (int v = v0 + tid; v < v1; v += nt) { p.X0[v] = p.X1[v]; p.P0[v] = p.P1[v]; }
            for (int s = s0 + tid; s < s1; s += nt) p.Z0[s] = p.Z1[s];
        }
    }
    double cnt = 0.0, zero = 0.0;
    for (int s = s0 + tid; s < s1; s += nt) {
        const int h = p.hkey[s];
        if (h >= 0) { a.dst[h] = (rank << 24) | (s - s0); cnt += 1.0; }
    }
    for (int v = v0 + tid; v < v1; v += nt) {                    // w/S per own slot: constant without reconditioning
        const int base = sbase[(v >> 5) - sl0], width = (sbase[(v >> 5) - sl0 + 1] - base) >> 5;
        const real Sv = p.Sinv[v];
        for (int j = 0; j < width; j++) {
            const int sk = base + j * 32 + (v & 31);
            a.wS[sk] = p.nbr[sk] >= 0 ? p.w[sk] * Sv : (real) 0;
        }
    }
    for (int v = v1 + tid; v < v1q; v += nt) {                   // lanes past V in the last slice
        const int base = sbase[(v >> 5) - sl0], width = (sbase[(v >> 5) - sl0 + 1] - base) >> 5;
        for (int j = 0; j < width; j++) a.wS[base + j * 32 + (v & 31)] = 0;
    }
    small_block_sum2<real>(cnt, zero, red);
    const uint32_t nz_own = (uint32_t) cnt;                      // half-edges whose copy lands in this CTA
    cl.sync();                                                   // dst complete, mbarriers initialised everywhere
    for (int h = 2 * e0 + tid; h < 2 * e1; h += nt) {
        const int d = __ldcg(a.dst + h);
        a.Ze[h] = __ldcg(p.Z0 + s0r[d >> 24] + (d & 0xffffff));
    }
    for (int v = tid; v < Vq; v += nt) {
        const bool in = v < V;
        Xs[v] = in ? __ldcg(p.X0 + v) : (real) 0; Ps[v] = in ? __ldcg(p.P0 + v) : (real) 0; Ts[v] = in ? p.T[v] : (real) 1;
        Xs[Vq + v] = 0; Ps[Vq + v] = 0;
    }
    const real tol = (real) ctl->tol, rho = p.rho, thc0 = p.thc0;
    const double eps = ctl->eps;
    double dif = ctl->dif;
    int phZ = 0, phA = 0, phP = 0, cb = 0;
    const int it_start = it;
    const int *__restrict__ nbr = p.nbr;
    const uint32_t own_bytes = (uint32_t) ((v1q - v0) * sizeof(real)), all_bytes = (uint32_t) (Vq * sizeof(real));
    __syncthreads();
    // partial A x over the own columns (N > 0): G threads per observation n, folded by shuffles
    auto own_partial = [&](const real *Xv) {
        int G = 1;
        while (G < 32 && 2 * G * N <= nt) G *= 2;
        const int g = tid & (G - 1);
        for (int n0 = 0; n0 < N; n0 += nt / G) {                              // uniform trip count: shuffles below
            const int n = n0 + tid / G;
            real acc = 0;
            if (n < N) {
#pragma unroll 4
                for (int v = v0 + g; v < v1; v += G) acc += __ldg(p.A + (size_t) N * v + n) * Xv[v];
            }
            for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (n < N && g == 0) partsS[rank * Nq + n] = acc;
        }
    };
    SPROF_DECL
    GSTAMP_DECL
    for (;;) {
        if (it == itMax || (real) dif < tol) break;                           // ref :429 (uniform over the cluster)
        const int nb = cb ^ 1, zb = it & 1;
        const int rel_it = it - it_start;                                     // X barrier and sums of this iteration's parity
        const uint32_t mb1 = mb0 + ((rel_it & 1) ? 32 : 8);
        double *sums_it = sums + (rel_it & 1) * 32;
        real *Xc = Xs + cb * Vq, *Pc = Ps + cb * Vq, *Xn = Xs + nb * Vq, *Pn = Ps + nb * Vq, *Zc = Zs + zb * nso;
        SPROF(0)
        // ---- forward point for dense operators (ref S1, S4, S6), own columns ----
        if (N != 0) {
            constexpr int GL = 8;                                             // lanes per column in the transposed products
            if (N > 0) {
                // partial A X over the own columns: after the first iteration it arrived with the iterate (see publish)
                if (it == it_start) {
                    if (tid == 0) mbar_expect(mb2, (uint32_t) ((C - 1) * Nq * sizeof(real)));
                    own_partial(Xc);
                    __syncthreads();
                    push_chunks(partsS + rank * Nq, (uint32_t) (Nq * sizeof(real)), rank, C, mb2);
                    mbar_wait(mb2, phA);
                }
                for (int n = tid; n < N; n += nt) {
                    real acc = 0;
                    for (int q = 0; q < C; q++) acc += partsS[q * Nq + n];                    // rank order: same bits everywhere
                    Rs[n] = p.Y[n] - acc;
                }
                __syncthreads();
            }
            if (tid == 0) mbar_expect(mb3, all_bytes - own_bytes);
            // G = -A^t R  |  G = (A^tA) X - A^tY : GL lanes per own column
            const int M = N > 0 ? N : V;
            const real *rhs = N > 0 ? Rs : Xc;
            for (int c0 = v0; c0 < v1; c0 += nt / GL) {                       // uniform trip count
                const int v = c0 + tid / GL, gl = tid & (GL - 1);
                real acc = 0;
                if (v < v1) {
                    const real *col = p.A + (size_t) M * v;
#pragma unroll 4
                    for (int m = gl; m < M; m += GL) acc += __ldg(col + m) * rhs[m];
                }
                for (int o = GL >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (v < v1 && gl == 0) Pc[v] = (real) 2 * Xc[v] - p.Ga[v] * (N > 0 ? -acc : acc - p.Y[v]);
            }
            __syncthreads();
            push_chunks(Pc + v0, own_bytes, rank, C, mb3);
            mbar_wait(mb3, phP);
        }
        SPROF(1)
        GSTAMP(0)
        // ---- edge prox + relaxation (ref S7, S8): both copies of an edge by one thread, pushed to the vertex owners ----
        if (tid == 0) mbar_expect(mb0, nz_own * (uint32_t) sizeof(real));
        const uint32_t zc0 = smem_u32(Zc);
        for (int eb = e0 + tid; eb < e1; eb += EU * nt) {
            int du[EU], dv[EU], u[EU], v[EU];
            real zu[EU], zv[EU];
#pragma unroll
            for (int k = 0; k < EU; k++) {
                const int e = eb + k * nt;
                const bool in = e < e1;
                const int2 dd = in ? *reinterpret_cast<const int2 *>(a.dst + 2 * e) : make_int2(0, 0);
                du[k] = dd.x; dv[k] = dd.y;
                u[k] = in ? __ldg(p.Eu + e) : 0; v[k] = in ? __ldg(p.Ev + e) : 0;
                if (in) ld2(a.Ze + 2 * e, zu[k], zv[k]); else { zu[k] = 0; zv[k] = 0; }
            }
#pragma unroll
            for (int k = 0; k < EU; k++) {
                const int e = eb + k * nt;
                if (e < e1) {
                    const real Tu = Ts[u[k]], Tv = Ts[v[k]];
                    const real tsum = Tu + Tv, inv = fast_rcp(tsum), th = thc0 * tsum;
                    const real wu = Tv * inv, wv = Tu * inv;     // W_d1 on the u side / on the v side
                    const real pu = Pc[u[k]] - zu[k], pv = Pc[v[k]] - zv[k], d0 = pu - pv;
                    const real b = d0 - min(max(d0, -th), th);   // soft threshold, exact 0 inside
                    const real nu = zu[k] + rho * (fma(wu, d0, pv) + wv * b - Xc[u[k]]);
                    const real nv = zv[k] + rho * (fma(wv, -d0, pu) + wu * (-b) - Xc[v[k]]);
                    st2(a.Ze + 2 * e, nu, nv);
                    push_val(map_rank(zc0 + (uint32_t) ((du[k] & 0xffffff) * sizeof(real)), du[k] >> 24), nu, map_rank(mb0, du[k] >> 24));
                    push_val(map_rank(zc0 + (uint32_t) ((dv[k] & 0xffffff) * sizeof(real)), dv[k] >> 24), nv, map_rank(mb0, dv[k] >> 24));
                }
            }
        }
        SPROF(2)
        GSTAMP(1)
        mbar_wait(mb0, phZ);
        SPROF(3)
        GSTAMP(2)
        // ---- average + vertex prox + evolution (ref S9, S10), own vertices ----
        if (tid == 0) mbar_expect(mb1, (all_bytes - own_bytes) * (N == 0 ? 2u : 1u) + (uint32_t) C * 16u + (N > 0 ? (uint32_t) ((C - 1) * Nq * sizeof(real)) : 0u));
        double sd = 0.0, sc = 0.0;
        auto finish_vertex = [&](int v, real x, real xo) {
            if (p.prox_kind == 0) {
                if (p.La_l1) {
                    const real th = p.Thl1[v];
                    if (x > th) x -= th;
                    else if (!p.positivity && x < -th) x += th;
                    else x = 0;
                } else if (p.positivity) { if (x < (real) 0) x = 0; }
            } else {
                if (p.has_lo && x < p.lo) x = p.lo;
                else if (p.has_hi && x > p.hi) x = p.hi;
            }
            Xn[v] = x;
            if (N == 0) Pn[v] = (real) 2 * x - p.Ga[v] * grad_at(p, v, x);
            const real dx = xo - x;
            sd += (double) (dx * dx); sc += (double) (x * x);
        };
        // stage 1: every own slot times its constant weight w/S (0 on padding), all threads, coalesced
        for (int i = tid; i < s1 - s0; i += nt) {
            const real c = a.wS[s0 + i];
            Zc[i] = c != (real) 0 ? c * Zc[i] : (real) 0;
        }
        __syncthreads();
        // stage 2: column sums in the reference's order (adding the zeros of the padding changes nothing)
        for (int v = v0 + tid; v < v1; v += nt) {
            const int sl = v >> 5;
            if (isw[sl - sl0]) continue;                 // wide slice (a hub): all warps of the CTA, below
            const int base = sbase[sl - sl0], width = (sbase[sl - sl0 + 1] - base) >> 5;
            const real *col = Zc + (base - s0) + (v & 31);
            real acc = 0;
#pragma unroll 8
            for (int j = 0; j < width; j++) acc += col[j * 32];
            finish_vertex(v, acc, Xc[v]);
        }
        for (int q = 0; q < nwide; q++) {
            const int sl = widelist[q];
            const int base = sbase[sl - sl0], width = (sbase[sl - sl0 + 1] - base) >> 5;
            const int v = sl * 32 + lane;
            const real *col = Zc + (base - s0) + lane;
            real acc = 0;
#pragma unroll 4
            for (int j = wid; j < width; j += nw) acc += col[j * 32];
            wpart[wid * 32 + lane] = acc;
            __syncthreads();
            if (wid == 0 && v < V) {
                real t = 0;
                for (int w = 0; w < nw; w++) t += wpart[w * 32 + lane];
                finish_vertex(v, t, Xc[v]);
            }
            __syncthreads();
        }
        SPROF(4)
        GSTAMP(3)
        // ---- publish: own chunk of the new iterate (and forward point) + evolution sums into every CTA ----
        {
            double *rb = red + (it & 1) * 32;                                 // by iteration parity: one barrier is enough
            sd = warp_sum(sd); sc = warp_sum(sc);
            if (lane == 0) { rb[2 * wid] = sd; rb[2 * wid + 1] = sc; }
            __syncthreads();                                                  // also orders the chunk writes above
            if (tid < C) {
                double d = 0.0, c = 0.0;
                for (int w = 0; w < nw; w++) { d += rb[2 * w]; c += rb[2 * w + 1]; }
                push_pair(map_rank(smem_u32(sums_it + 2 * rank), tid), d, c, map_rank(mb1, tid));
            }
            push_chunks(Xn + v0, own_bytes, rank, C, mb1);
            if (N == 0) push_chunks(Pn + v0, own_bytes, rank, C, mb1);
            if (N > 0) {                                                      // the next iteration's partial A x rides along
                own_partial(Xn);
                __syncthreads();
                push_chunks(partsS + rank * Nq, (uint32_t) (Nq * sizeof(real)), rank, C, mb1);
            }
        }
        SPROF(5)
        GSTAMP(4)
        { int phX = (rel_it >> 1) & 1; mbar_wait(mb1, phX); }
        SPROF(6)
        GSTAMP(5)
        if (track) {
            double d = 0.0, c = 0.0;
            for (int q = 0; q < C; q++) { d += sums_it[2 * q]; c += sums_it[2 * q + 1]; }     // rank order: same bits everywhere
            dif = (c > eps) ? d / c : d / eps;
            if (rank == 0 && tid == 0 && p.dDif) p.dDif[it] = (real) dif;
        }
        cb = nb;
        it++;
        SPROF(7)
    }
    SPROF_OUT
    GSTAMP_OUT
    cl.sync();          // every CTA has received everything it was sent: outgoing copies are complete, peers may exit
    // the large path's readers (extract_x_kernel, a later run) look at the buffers of parity it & 1
    {
        const real *Xc = Xs + cb * Vq, *Pc = Ps + cb * Vq;
        for (int v = v0 + tid; v < v1; v += nt) {
            p.X0[v] = Xc[v]; p.X1[v] = Xc[v];
            if (N == 0) { p.P0[v] = Pc[v]; p.P1[v] = Pc[v]; }
        }
        if (it > it_start)
            for (int h = 2 * e0 + tid; h < 2 * e1; h += nt) {
                const int d = a.dst[h];
                const real z = a.Ze[h];
                p.Z0[s0r[d >> 24] + (d & 0xffffff)] = z; p.Z1[s0r[d >> 24] + (d & 0xffffff)] = z;
            }
    }
    if (rank == 0 && tid == 0) { ctl->it = it; ctl->dif = dif; ctl->state = ST_FINAL; ctl->hold = 1; }
}

template <typename real>
__global__ void extract_x_kernel(QP<real> p, real *__restrict__ out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.Vown) return;
    out[v] = ((p.ctl->it & 1) ? p.X1 : p.X0)[p.vbeg + v];
}

// ---------------------------------------------------------------------------------------------
// objective (ref :388-422).  Runs once per trip when Obj is requested, before the stop test.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) objective_kernel(QP<real> p)
{
    __shared__ double smem[64];
    Ctl *ctl = p.ctl;
    const CtlView cv = load_ctl(ctl);
    if (cv.hold) return;
    const int it = cv.it;
    const real *__restrict__ X = (it & 1) ? p.X1 : p.X0;
    double r[1] = {0.0};
    const size_t stride = (size_t) gridDim.x * blockDim.x, t0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const bool slab = p.slab.on != 0;                        // slab-sharded: owned vertices, edges whose first end is owned
    if (slab) {                                              // ... whose other end may be a ghost: the neighbours' stores have landed
        if (threadIdx.x == 0) slab_wait_neighbours(p.slab, it);
        __syncthreads();
    }
    for (size_t e = t0; e < (size_t) p.E; e += stride) {
        const int u = p.Eu[e];
        if (slab && (u < p.vbeg || u >= p.vend)) continue;
        r[0] += (double) p.La_d1[e] * (double) rabs(X[u] - X[p.Ev[e]]);
    }
    for (size_t v = (slab ? (size_t) p.vbeg : 0) + t0; v < (slab ? (size_t) p.vend : (size_t) p.V); v += stride) {
        const real x = X[v];
        if (p.N == 0) r[0] += (double) x * ((double) 0.5 * (double) (p.A ? p.A[v] * x : x) - (double) p.Y[v]);
        else if (p.N < 0) r[0] += (double) x * ((double) 0.5 * (double) p.G[v] - (double) p.Y[v]);
        if (p.La_l1) r[0] += (double) p.La_l1[v] * (double) rabs(x);
    }
    if (p.N > 0)
        for (size_t n = t0; n < (size_t) p.N; n += stride) r[0] += 0.5 * (double) p.R[n] * (double) p.R[n];
    if (grid_sum<1>(ctl, p.partial, r, smem) && threadIdx.x == 0) {
        if (slab) *p.obj_part = r[0];                            // travels with the trip's stopping sums (quad_slab_decide)
        else p.dObj[it] = (real) r[0];
    }
}

// ---------------------------------------------------------------------------------------------
// dense operators
// ---------------------------------------------------------------------------------------------
template <typename real, int VEC> struct VecT;
template <> struct VecT<float, 4> { typedef float4 type; };
template <> struct VecT<float, 1> { typedef float type; };
template <> struct VecT<double, 2> { typedef double2 type; };
template <> struct VecT<double, 1> { typedef double type; };

template <typename real, int VEC>
__device__ __forceinline__ void vload(const real *q, real (&o)[VEC])
{
    typedef typename VecT<real, VEC>::type vt;
    const vt t = __ldg(reinterpret_cast<const vt *>(q));
    const real *tp = reinterpret_cast<const real *>(&t);
#pragma unroll
    for (int k = 0; k < VEC; k++) o[k] = tp[k];
}

// N<0:  G = (A^tA) X column by column (ref :368-376), then P = 2X - Ga.(G - AtY) (:441-445,:463-464).
// CG columns per CTA pass so every X vector load feeds CG independent A streams.
template <typename real, int VEC, int CG>
__global__ void __launch_bounds__(256) ata_gemv_kernel(QP<real> p)
{
    __shared__ double smem[CG * 32];
    const CtlView cv = load_ctl(p.ctl);
    if (cv.hold) return;
    real *P = (cv.it & 1) ? p.P1 : p.P0;
    const size_t V = (size_t) p.V;
    const real *__restrict__ X = (cv.it & 1) ? p.X1 : p.X0;
    for (size_t c0 = (size_t) blockIdx.x * CG; c0 < V; c0 += (size_t) gridDim.x * CG) {
        double acc[CG];
        real part[CG];
        const real *col[CG];
#pragma unroll
        for (int c = 0; c < CG; c++) { part[c] = 0; col[c] = p.A + V * (c0 + c < V ? c0 + c : V - 1); }
        for (size_t u = (size_t) threadIdx.x * VEC; u < V; u += (size_t) blockDim.x * VEC) {
            real xv[VEC];
            vload<real, VEC>(X + u, xv);
            real av[CG][VEC];
#pragma unroll
            for (int c = 0; c < CG; c++) vload<real, VEC>(col[c] + u, av[c]);
#pragma unroll
            for (int c = 0; c < CG; c++)
#pragma unroll
                for (int k = 0; k < VEC; k++) part[c] += av[c][k] * xv[k];
        }
#pragma unroll
        for (int c = 0; c < CG; c++) acc[c] = (double) part[c];
        block_sum<CG>(acc, smem);
        if (threadIdx.x < CG && c0 + threadIdx.x < V) {
            const size_t v = c0 + threadIdx.x;
            real g = 0;
#pragma unroll
            for (int c = 0; c < CG; c++) if (c == (int) threadIdx.x) g = (real) acc[c];
            p.G[v] = g;
            const real x = X[v];
            P[v] = (real) 2 * x - p.Ga[v] * (g - p.Y[v]);
        }
    }
}

// N>0, pass 1:  partial sums of A X over the CTA's block of columns (ref :356-367).
// Lanes run along n (contiguous in the column-major A), TY columns are walked side by side.
template <typename real, int VEC>
__global__ void __launch_bounds__(256) ax_partial_kernel(QP<real> p, int TX, double *__restrict__ parts)
{
    extern __shared__ double dsm[];
    const CtlView cv = load_ctl(p.ctl);
    if (cv.hold) return;
    const int N = p.N, TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const size_t V = (size_t) p.Vown;
    const size_t v0 = V * blockIdx.x / gridDim.x, v1 = V * (blockIdx.x + 1) / gridDim.x;
    const real *__restrict__ X = (cv.it & 1) ? p.X1 : p.X0;
    const int nvec = (N + VEC - 1) / VEC;
    for (int nb0 = 0; nb0 < nvec; nb0 += TX) {     // uniform trip count: every thread reaches the barriers
        const int nb = nb0 + tx;
        const bool on = nb < nvec;
        double acc[VEC];
        real part[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) part[k] = 0;
        if (on) {
            const real *__restrict__ Ab = p.A + (size_t) nb * VEC;
            size_t v = v0 + ty;
            // eight columns in flight per thread (the same two groups of four, added in the same order as the loop below)
            for (; v + 7 * (size_t) TY < v1; v += 8 * (size_t) TY) {
                real a0[VEC], a1[VEC], a2[VEC], a3[VEC], a4[VEC], a5[VEC], a6[VEC], a7[VEC];
                vload<real, VEC>(Ab + (size_t) N * v, a0);
                vload<real, VEC>(Ab + (size_t) N * (v + TY), a1);
                vload<real, VEC>(Ab + (size_t) N * (v + 2 * (size_t) TY), a2);
                vload<real, VEC>(Ab + (size_t) N * (v + 3 * (size_t) TY), a3);
                vload<real, VEC>(Ab + (size_t) N * (v + 4 * (size_t) TY), a4);
                vload<real, VEC>(Ab + (size_t) N * (v + 5 * (size_t) TY), a5);
                vload<real, VEC>(Ab + (size_t) N * (v + 6 * (size_t) TY), a6);
                vload<real, VEC>(Ab + (size_t) N * (v + 7 * (size_t) TY), a7);
                const real x0 = X[v], x1 = X[v + TY], x2 = X[v + 2 * (size_t) TY], x3 = X[v + 3 * (size_t) TY];
                const real x4 = X[v + 4 * (size_t) TY], x5 = X[v + 5 * (size_t) TY], x6 = X[v + 6 * (size_t) TY], x7 = X[v + 7 * (size_t) TY];
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    part[k] += a0[k] * x0 + a1[k] * x1 + a2[k] * x2 + a3[k] * x3;
                    part[k] += a4[k] * x4 + a5[k] * x5 + a6[k] * x6 + a7[k] * x7;
                }
            }
            for (; v + 3 * (size_t) TY < v1; v += 4 * (size_t) TY) {
                real a0[VEC], a1[VEC], a2[VEC], a3[VEC];
                vload<real, VEC>(Ab + (size_t) N * v, a0);
                vload<real, VEC>(Ab + (size_t) N * (v + TY), a1);
                vload<real, VEC>(Ab + (size_t) N * (v + 2 * (size_t) TY), a2);
                vload<real, VEC>(Ab + (size_t) N * (v + 3 * (size_t) TY), a3);
                const real x0 = X[v], x1 = X[v + TY], x2 = X[v + 2 * (size_t) TY], x3 = X[v + 3 * (size_t) TY];
#pragma unroll
                for (int k = 0; k < VEC; k++) part[k] += a0[k] * x0 + a1[k] * x1 + a2[k] * x2 + a3[k] * x3;
            }
            for (; v < v1; v += TY) {
                real a0[VEC];
                vload<real, VEC>(Ab + (size_t) N * v, a0);
                const real x0 = X[v];
#pragma unroll
                for (int k = 0; k < VEC; k++) part[k] += a0[k] * x0;
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] = (double) part[k];
        if (TY > 1) {       // fold the TY column lanes in a fixed order through shared memory
            __syncthreads();
#pragma unroll
            for (int k = 0; k < VEC; k++) dsm[((size_t) ty * TX + tx) * VEC + k] = acc[k];
            __syncthreads();
            if (ty == 0) {
                for (int y = 1; y < TY; y++)
#pragma unroll
                    for (int k = 0; k < VEC; k++) acc[k] += dsm[((size_t) y * TX + tx) * VEC + k];
            }
        }
        if (on && ty == 0) {
#pragma unroll
            for (int k = 0; k < VEC; k++)
                if (nb * VEC + k < N) parts[(size_t) blockIdx.x * N + (size_t) nb * VEC + k] = acc[k];
        }
    }
}

template <typename real>
__global__ void ax_finish_kernel(QP<real> p, const double *__restrict__ parts)
{
    if (load_ctl(p.ctl).hold) return;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.N) return;
    double a = 0;
    for (int b = 0; b < p.nparts; b++) a += parts[(size_t) b * p.N + n];
    if (p.sharded) p.rsum[n] = a;                 // all-reduced over the ranks, then ax_apply_kernel
    else p.R[n] = p.Y[n] - (real) a;
}

template <typename real>
__global__ void ax_apply_kernel(QP<real> p)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < p.N) p.R[n] = p.Y[n] - (real) p.rsum[n];
}

// N>0, pass 2 (mode 0):  G = -A^t R, P = 2X - Ga.G  (ref :432-440, :463-464); GL lanes per column.
// mode 1 (setup): Ga0 = ||A_v||^2 and AtY = <A_v, Y>   (ref :102-110, :126-134).
constexpr int ATR_COLS = 4;            // columns a lane group works on at a time: 4 x (N / GL) vector loads in flight per lane
template <typename real, int VEC>
__global__ void __launch_bounds__(256) atr_kernel(QP<real> p, int GL, int mode)
{
    const CtlView cv = load_ctl(p.ctl);
    if (mode == 0 && cv.hold) return;
    const int N = p.N;
    const size_t grp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) / GL;
    const int gl = threadIdx.x % GL;
    const size_t vb = grp * ATR_COLS, Vown = (size_t) p.Vown;
    const real *__restrict__ rhs = mode ? p.Y : p.R;
    const real *col[ATR_COLS];
#pragma unroll
    for (int c = 0; c < ATR_COLS; c++) col[c] = p.A + (size_t) N * (vb + c < Vown ? vb + c : Vown - 1);
    const int nvec = (N + VEC - 1) / VEC;
    real s0[ATR_COLS], s1[ATR_COLS];
#pragma unroll
    for (int c = 0; c < ATR_COLS; c++) { s0[c] = 0; s1[c] = 0; }
    for (int nb = gl; nb < nvec; nb += GL) {
        real a[ATR_COLS][VEC], r[VEC];
#pragma unroll
        for (int c = 0; c < ATR_COLS; c++) vload<real, VEC>(col[c] + (size_t) nb * VEC, a[c]);
        vload<real, VEC>(rhs + (size_t) nb * VEC, r);
#pragma unroll
        for (int c = 0; c < ATR_COLS; c++)
#pragma unroll
            for (int k = 0; k < VEC; k++) { s0[c] += a[c][k] * r[k]; if (mode) s1[c] += a[c][k] * a[c][k]; }
    }
    for (int o = GL >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < ATR_COLS; c++) {
            s0[c] += __shfl_xor_sync(0xffffffffu, s0[c], o);
            if (mode) s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
        }
    }
    if (gl == 0) {
#pragma unroll
        for (int c = 0; c < ATR_COLS; c++) {
            const size_t v = vb + c;
            if (v >= Vown) continue;
            if (mode) { p.Ga0[v] = s1[c]; p.AtY[v] = s0[c]; }
            else {
                const real g = -s0[c];
                p.G[v] = g;
                ((cv.it & 1) ? p.P1 : p.P0)[v] = (real) 2 * ((cv.it & 1) ? p.X1 : p.X0)[v] - p.Ga[v] * g;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename real>
struct QuadSolver : SolverBase {
    const QuadProblem<real> q;
    bool keep_x0 = false;
    SellGraph g;
    DiaGraph dg;
    bool dia = false;               // constant-offset layout (dia.cuh) instead of the half-edge table
    long long nhalf = 0;            // half-edge copies held (slots of the table / 2.ND.V)
    QP<real> p{};
    DevBuf<Ctl> ctl;
    DevBuf<real> Wrow;
    DevIn<real> Y, A, La_d1, La_l1, L, X0;
    DevIn<int> Eu, Ev;
    DevBuf<real> w, thc, Z0, Z1, X0b, X1b, P0b, P1b, T, Sinv, Ga, Ga0, G, Thl1, R, AtY, dObj, dDif, Xout;
    DevBuf<double> partial, parts, sums, rsum, amp_sums, obj_part;
    DevBuf<int4> pkb;               // per-slot packs for the iteration kernel
    DevBuf<int2> pk2b;
    DevBuf<int> pos, dstb;          // small mode: slot / (owner CTA, slot) of every half-edge key
    DevBuf<real> ze, sparts, ws;        //             edge-ordered copies, partial A x of the CTAs
    DevBuf<double> xsum;            //             partial sums of the CTAs
    Halo<real> halo;
    SlabLink link;                  // slab-sharded solve (slab.cuh)
    bool slab = false;
    int recond_parity = 0;          // parity of the trip a reconditioning interrupts
    bool sharded = false;
    int ogrid = 0;                  // CTAs covering the owned vertices
    Ctl *hctl = nullptr;            // pinned mirror of the control block, one per batch in flight
    cudaEvent_t ev[2]{}, ev_a{}, ev_b{}, ev_c{};
    cudaGraphExec_t gexec = nullptr;
    ProgressPrinter progress;
    int graph_trips = 0;
    int vec = 1, TX = 1, GL = 1, vgrid = 0;
    size_t ax_smem = 0;

    QuadSolver(const QuadProblem<real> &q_, bool keep) : q(q_), keep_x0(keep) {}
    ~QuadSolver() override
    {
        if (gexec) cudaGraphExecDestroy(gexec);
        pinned_put(hctl);
        for (auto e : {ev[0], ev[1], ev_a, ev_b, ev_c}) event_put(e);
    }

    void setup()
    {
        const int V = q.V, E = q.E, N = q.N;
        PFDR_REQUIRE(V > 0 && E >= 0, "V must be positive and E non-negative");
        PFDR_REQUIRE(q.itMax >= 0, "itMax must be non-negative");
        PFDR_REQUIRE(N == 0 || q.A != nullptr, "A must not be NULL when N != 0");
        sharded = q.shard != nullptr;
        slab = q.slab != nullptr;
        if (slab) {
            PFDR_REQUIRE(!sharded, "a solve is sharded either by slabs or by a shard plan");
            PFDR_REQUIRE(N == 0, "slab-sharded solves support diagonal / identity A (N = 0) in this build");
            PFDR_REQUIRE(q.slab->world >= 1 && q.slab->lo <= q.slab->v0 && q.slab->v0 < q.slab->v1 && q.slab->v1 <= q.slab->hi &&
                         q.slab->hi - q.slab->lo == V, "bad slab: need lo <= v0 < v1 <= hi and Vloc = hi - lo");
        }
        if (sharded) {
            PFDR_REQUIRE(N >= 0, "sharded solves support N >= 0 (A diagonal, identity or dense N-by-Vown) in this build");
            PFDR_REQUIRE(q.Obj == nullptr && q.difRcd == (real) 0, "sharded solves do not record Obj nor recondition in this build");
            PFDR_REQUIRE(q.shard->Vown > 0 && q.shard->Vown <= V, "bad shard: Vown");
        }
        cudaStream_t st = ctx().stream;
        for (auto *e : {&ev[0], &ev[1], &ev_a, &ev_b, &ev_c}) *e = event_get();
        PFDR_CUDA(cudaEventRecord(ev_a, st));
        static_assert(2 * sizeof(Ctl) <= PINNED_BLOCK, "control block mirror");
        hctl = (Ctl *) pinned_get();

        const size_t ylen = (N > 0) ? (size_t) N : (size_t) V;
        // sharded dense A holds the OWNED columns only (ghost columns are never applied on this rank)
        const size_t acols = (sharded && N > 0) ? (size_t) q.shard->Vown : (size_t) V;
        const size_t alen = (N > 0) ? (size_t) N * acols : (N < 0 ? (size_t) V * V : (size_t) V);
        Eu.bind(q.Eu, E); Ev.bind(q.Ev, E);
        Y.bind(q.Y, ylen); A.bind(q.A, q.A ? alen : 0);
        La_d1.bind(q.La_d1, E); La_l1.bind(q.prox_kind == 0 ? q.La_l1 : nullptr, V);
        L.bind(q.L, q.L ? (q.Ltype == PFDR_B200_SCAL ? 1 : (size_t) V) : 0);
        X0.bind(q.X, V);
        {   // Constant-offset graph (grids) that will not run in the single-launch cluster kernels: index-free layout.
            // PFDR_B200_DIA=0 keeps the half-edge table, =1 tries the stencil layout at any size (tests compare the two).
            const char *de = getenv("PFDR_B200_DIA"), *se = getenv("PFDR_B200_SMALL");
            const bool small_off = se && se[0] == '0';
            const bool forced = de && de[0] == '1';
            if (!sharded && !(de && de[0] == '0') && (forced || small_off || slab || !small_candidate()))
                dia = build_dia(dg, Eu, Ev, V, E);
            PFDR_REQUIRE(dia || !slab, "slab sharding needs a constant-offset (grid) graph with its edges in vertex-major order");
        }
        if (dia) {
            nhalf = 2LL * dg.ND * V;
            Wrow.alloc((size_t) nhalf);
            ctx().stats.slots = 2LL * E; ctx().stats.layout = 1;
        } else {
            build_sell(g, Eu, Ev, V, E, 32);
            rowsW = g.max_width;
            nhalf = g.nslots;
            w.alloc((size_t) std::max<long long>(nhalf, 1));
        }

        const size_t ns = (size_t) std::max<long long>(nhalf, 1);
        if (slab) {          // what the neighbouring ranks write / copy into lives in the peer-visible arena
            link.reserve(SLAB_Z0, ns * sizeof(real)); link.reserve(SLAB_Z1, ns * sizeof(real));
            link.reserve(SLAB_F0, (size_t) V * sizeof(real)); link.reserve(SLAB_F1, (size_t) V * sizeof(real)); link.reserve(SLAB_T, (size_t) V * sizeof(real));
            link.reserve(SLAB_X0, (size_t) V * sizeof(real)); link.reserve(SLAB_X1, (size_t) V * sizeof(real));
            link.commit(q.slab->world, q.slab->rank, q.slab->lo, q.slab->hi, q.slab->v0, q.slab->v1, 1);
            Z0.borrow(link.ptr<real>(SLAB_Z0), ns); Z1.borrow(link.ptr<real>(SLAB_Z1), ns);
            P0b.borrow(link.ptr<real>(SLAB_F0), V); P1b.borrow(link.ptr<real>(SLAB_F1), V); T.borrow(link.ptr<real>(SLAB_T), V);
            X0b.borrow(link.ptr<real>(SLAB_X0), V); X1b.borrow(link.ptr<real>(SLAB_X1), V);
            obj_part.alloc(1); obj_part.zero();
        } else {
            Z0.alloc(ns); Z1.alloc(ns); P0b.alloc(V); P1b.alloc(V); T.alloc(V); X0b.alloc(V); X1b.alloc(V);
        }
        Sinv.alloc(V); Ga.alloc(V); Ga0.alloc(V);
        if (La_l1.p) Thl1.alloc(V);
        if (N != 0) G.alloc(V);
        if (N > 0) { R.alloc(N); AtY.alloc(V); rsum.alloc(N); p.rsum = rsum.p; }
        if (q.Obj) dObj.alloc((size_t) q.itMax + 1);
        if (q.Dif || q.verbose > 0) dDif.alloc((size_t) std::max(q.itMax, 1));
        vgrid = (int) cdiv(V, 256);
        ogrid = sharded ? (int) cdiv(q.shard->Vown, 256) : (slab ? (int) cdiv(q.slab->v1 - q.slab->v0, 256) : vgrid);
        ctl.alloc(1); sums.alloc(2 * (size_t) std::max(1, comm().world)); sums.zero(); amp_sums.alloc(2);

        // padded slots of Z must never hold garbage that could turn into NaN traps: zero them
        Z0.zero(); Z1.zero(); w.zero();

        p.ctl = ctl.p; p.V = V; p.N = N; p.E = E;
        p.Vown = sharded ? q.shard->Vown : (slab ? q.slab->v1 - q.slab->v0 : V); p.sharded = sharded ? 1 : 0; p.sums = sums.p; p.amp_sums = amp_sums.p;
        p.world = sharded ? comm().world : 1; p.rank = sharded ? comm().rank : 0;
        p.slice_off = g.slice_off.p; p.nbr = g.nbr.p; p.twin = g.twin.p; p.hkey = g.hkey.p;
        p.dia_nd = dia ? dg.ND : 0;
        for (int i = 0; i < 3; i++) p.dia_off[i] = dia ? dg.off[i] : 0;
        p.vbeg = slab ? q.slab->v0 - q.slab->lo : 0; p.vend = p.vbeg + p.Vown; p.dmask = dg.mask.p; p.eoff = dg.eoff.p; p.Wrow = Wrow.p;
        p.slab = SlabDev{}; p.plo = SlabPeer<real>{}; p.phi = SlabPeer<real>{};
        if (slab) { p.slab = link.dev(); slab_geometry(); }
        p.w = w.p; p.thc = nullptr; p.thc0 = 0; p.Z0 = Z0.p; p.Z1 = Z1.p;
        p.pk = nullptr; p.pk2 = nullptr;           // set by pack_slots() after each (re)conditioning 
        p.X0 = X0b.p; p.X1 = X1b.p; p.P0 = P0b.p; p.P1 = P1b.p; p.T = T.p; p.Sinv = Sinv.p; p.Ga = Ga.p; p.Ga0 = Ga0.p; p.G = G.p; p.Thl1 = Thl1.p;
        p.R = R.p; p.AtY = AtY.p; p.Y = Y.p; p.A = A.p; p.La_d1 = La_d1.p; p.La_l1 = La_l1.p; p.L = L.p;
        p.Eu = Eu.p; p.Ev = Ev.p; p.Ltype = q.Ltype; p.positivity = q.positivity; p.prox_kind = q.prox_kind;
        p.lo = q.lo; p.hi = q.hi;
        p.has_lo = (q.prox_kind == 1 && -HUGE_VAL < (double) q.lo);
        p.has_hi = (q.prox_kind == 1 && (double) q.hi < HUGE_VAL);
        p.rho = q.rho; p.condMin = q.condMin; p.dObj = dObj.p; p.dDif = dDif.p; p.obj_part = obj_part.p;

        // dense operator geometry
        const int vmax = (sizeof(real) == 4) ? 4 : 2;
        if (N > 0) {
            const bool al = (N % vmax == 0) && (((uintptr_t) A.p) % 16 == 0) && (((uintptr_t) Y.p) % 16 == 0);
            vec = al ? vmax : 1;
            const int nvec = (N + vec - 1) / vec;
            TX = 1; while (TX < nvec && TX < 256) TX <<= 1;
            GL = 1; while (GL < nvec && GL < 32) GL <<= 1;
            p.nparts = std::max(1, std::min(num_sms() * 4, ((sharded ? q.shard->Vown : V) + 63) / 64));
            parts.alloc((size_t) p.nparts * N);
            ax_smem = (256 / TX > 1) ? (size_t) 256 * vec * sizeof(double) : 0;
        } else if (N < 0) {
            vec = ((V % vmax == 0) && (((uintptr_t) A.p) % 16 == 0)) ? vmax : 1;
        }
        const int maxgrid = std::max(std::max(std::max(vgrid, dia ? 0 : g.nslices), p.nparts), num_sms() * 8);
        partial.alloc((size_t) maxgrid * 4);
        p.partial = partial.p;

        if (sharded) {
            halo.setup(*q.shard, g.hkey.p, g.nslots, 2 * (size_t) E, 1);
            halo.build(HALO_VA0 | HALO_VA1); halo.build(HALO_VA0 | HALO_Z); halo.build(HALO_VA0); halo.build(HALO_Z);
        }
        if (!(slab && link.world > 1)) reset();            // a slab-sharded solve is initialised by slab_connect (collective)
        if (!keep_x0) X0.own.release();
        PFDR_CUDA(cudaEventRecord(ev_b, st));
    }

    // boundary vertices (within the band of a neighbouring rank) are worked on first: slab.cuh
    void slab_geometry()
    {
        const int nown = p.vend - p.vbeg;
        p.slab.band = dg.band();
        p.slab.lowB = link.nb_lo >= 0 ? std::min(dg.band(), nown) : 0;
        p.slab.highB = link.nb_hi >= 0 ? std::min(dg.band(), nown - p.slab.lowB) : 0;
        p.slab.nbchunks = (p.slab.lowB + p.slab.highB + 31) / 32;
    }
    // ghost zones of the neighbours <- this rank's boundary: metric T, forward point and cut-edge copies of the parity-0 buffers
    void slab_refresh(int par = 0)
    {
        const size_t Vp = (size_t) q.V;
        const int ND = dg.ND;
        for (int side = 0; side < 2; side++) {
            const int nb = side ? link.nb_hi : link.nb_lo;
            if (nb < 0) continue;
            const SlabBlob &b = link.blobs[nb];
            const long long depth = side ? b.v0 - b.lo : b.hi - b.v1;          // its ghost zone on our side
            const long long first = side ? p.vend - depth : p.vbeg;
            const long long dshift = link.lo - b.lo;
            const size_t Vn = (size_t) (b.hi - b.lo);
            link.push<real>(SLAB_T, nb, (size_t) first, (size_t) depth, dshift);
            link.push<real>(par ? SLAB_F1 : SLAB_F0, nb, (size_t) first, (size_t) depth, dshift);
            link.push<real>(par ? SLAB_X1 : SLAB_X0, nb, (size_t) first, (size_t) depth, dshift);
            for (int i = 0; i < ND; i++) {                                       // copies on the cut edges of offset d_i
                const long long d = std::min<long long>(dg.off[i], depth), f = side ? p.vend - d : p.vbeg;
                const int plane = side ? i : ND + i;                             // our out_i copies go up, our in_i copies down
                link.push<real>(par ? SLAB_Z1 : SLAB_Z0, nb, (size_t) plane * Vp + (size_t) f, (size_t) d,
                                (long long) plane * ((long long) Vn - (long long) Vp) + dshift);
            }
        }
        link.barrier();
    }
    void slab_check()
    {
        int err = 0;
        PFDR_CUDA(cudaMemcpy(&err, &link.ptr<SlabMail>(SLAB_MAILBOX)->error, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) throw Error(PFDR_B200_ERR_CUDA, "slab-sharded solve: a rank stopped answering (timed out waiting for its flag)");
    }
    void slab_export(void *blob) override
    {
        PFDR_REQUIRE(slab, "not a slab-sharded solver");
        memset(blob, 0, PFDR_B200_SLAB_BLOB);
        memcpy(blob, &link.mine, sizeof(SlabBlob));
    }
    void slab_connect(const void *blobs, bool same_process) override
    {
        PFDR_REQUIRE(slab, "not a slab-sharded solver");
        std::vector<SlabBlob> all(link.world);
        for (int r = 0; r < link.world; r++) memcpy(&all[r], (const char *) blobs + (size_t) r * PFDR_B200_SLAB_BLOB, sizeof(SlabBlob));
        link.connect(all.data(), same_process);
        p.slab = link.dev(); slab_geometry();
        for (int side = 0; side < 2; side++) {
            const int nb = side ? link.nb_hi : link.nb_lo;
            if (nb < 0) continue;
            SlabPeer<real> &pe = side ? p.phi : p.plo;
            pe.Z[0] = link.ptr<real>(SLAB_Z0, nb); pe.Z[1] = link.ptr<real>(SLAB_Z1, nb);
            pe.F[0] = link.ptr<real>(SLAB_F0, nb); pe.F[1] = link.ptr<real>(SLAB_F1, nb);
            pe.X[0] = link.ptr<real>(SLAB_X0, nb); pe.X[1] = link.ptr<real>(SLAB_X1, nb);
            pe.plane = (size_t) (link.blobs[nb].hi - link.blobs[nb].lo);
            pe.shift = link.lo - link.blobs[nb].lo;
        }
        if (link.world > 1) reset();                          // collective: first preconditioning, ghost zones, barrier
    }

    // initial iterate, edge copies, first preconditioning and control block: device work only
    void reset() override
    {
        cudaStream_t st = ctx().stream;
        PFDR_REQUIRE(X0.p != nullptr, "reset: the initial iterate was not kept on the device");
        if (gexec && p.thc) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        p.thc = nullptr;
        // control block (ref :286-292, :339-343)
        Ctl h{};
        const real tol2 = q.difTol * q.difTol, rcd2 = q.difRcd * q.difRcd;
        h.it = 0; h.state = (q.itMax == 0) ? ST_FINAL : ST_RUN; h.hold = 0; h.itMax = q.itMax;
        h.track = (q.difTol > (real) 0 || q.difRcd > (real) 0 || q.Dif != nullptr);
        h.tol = (double) tol2; h.rcd = (double) rcd2;
        h.dif = (double) (tol2 > rcd2 ? tol2 : rcd2);
        h.eps = (double) (((real) 0 < q.difTol && q.difTol < real_eps<real>()) ? q.difTol : real_eps<real>());
        hctl[0] = h;
        PFDR_CUDA(cudaMemcpyAsync(ctl.p, &hctl[0], sizeof(Ctl), cudaMemcpyHostToDevice, st));
        const bool linked = slab && link.world > 1;
        if (linked) {                                        // nobody still iterates on the buffers about to be rewritten
            PFDR_REQUIRE(link.connected, "slab-sharded solver: reset before slab_connect");
            link.barrier();
            link.run_id++;
            p.slab.run_id = link.run_id;
            if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        }
        p.sharded = (sharded || linked) ? 1 : 0;             // amp_kernel leaves its sums for the all-reduce
        // first preconditioning
        if (q.N > 0) launch_atr(1);
        PFDR_LAUNCH(amp_kernel<real>, ogrid, 256, 0, p, 0);
        if (sharded) { comm_allreduce_sum(amp_sums.p, 2); PFDR_LAUNCH(amp_finish_kernel<real>, 1, 1, 0, p, 0); }
        if (linked) { link.allreduce_sum(amp_sums.p, 2); PFDR_LAUNCH(amp_finish_kernel<real>, 1, 1, 0, p, 0); p.sharded = 0; }
        if (dia) quad_dia_launch_init<real>(p, X0.p, vgrid);
        else PFDR_LAUNCH(init_vertex_kernel<real>, vgrid, 256, 0, p, X0.p);
        // ghosts: the metric constants and forward point come from their owners
        if (sharded) halo.exchange(HALO_VA0 | HALO_VA1, PingPong<real>(T.p), PingPong<real>(P0b.p), PingPong<real>(), nullptr, PARITY_ZERO);
        if (linked) slab_refresh();
        PFDR_CUDA(cudaMemcpyAsync(&hctl[1], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        if (linked) slab_check();
        const real thc0 = (real) 1 / (real) hctl[1].amp;       // La_d1 / (c La_d1)
        if (gexec && thc0 != p.thc0) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        p.thc0 = thc0;
        pack_slots();
    }

    // float32 before a reconditioning: {nbr, twin} in 8 bytes (the weight stays a separate stream: 16 bytes would carry
    // an unused word: -4 % instead of -9.5 % on C2).  float64, and float32 with per-slot thresholds: 16 bytes.
    void pack_slots()
    {
        p.pk = nullptr; p.pk2 = nullptr;
        if (dia) return;
        const size_t ns = (size_t) std::max<long long>(g.nslots, 1);
        // (buffers are kept across reset(): a captured graph holds their addresses)
        if (sizeof(real) == 4 && !p.thc) { if (pk2b.n != ns) pk2b.alloc(ns); p.pk2 = pk2b.p; }
        else { if (pkb.n != ns) pkb.alloc(ns); p.pk = pkb.p; }
        PFDR_LAUNCH(pack_slots_kernel<real>, cdiv(ns, 256), 256, 0, p, (int4 *) p.pk, g.nslots);
    }

    void launch_atr(int mode)
    {
        const size_t threads = (((size_t) p.Vown + ATR_COLS - 1) / ATR_COLS) * GL;
        if (vec > 1) PFDR_LAUNCH((atr_kernel<real, (sizeof(real) == 4 ? 4 : 2)>), cdiv(threads, 256), 256, 0, p, GL, mode);
        else PFDR_LAUNCH((atr_kernel<real, 1>), cdiv(threads, 256), 256, 0, p, GL, mode);
    }

    void enqueue_trip()
    {
        constexpr int VM = (sizeof(real) == 4) ? 4 : 2;
        if (q.N > 0) {
            if (vec > 1) PFDR_LAUNCH((ax_partial_kernel<real, VM>), p.nparts, 256, ax_smem, p, TX, parts.p);
            else PFDR_LAUNCH((ax_partial_kernel<real, 1>), p.nparts, 256, ax_smem, p, TX, parts.p);
            PFDR_LAUNCH(ax_finish_kernel<real>, cdiv(q.N, 128), 128, 0, p, parts.p);
            if (sharded) {
                comm_allreduce_sum(rsum.p, q.N);
                PFDR_LAUNCH(ax_apply_kernel<real>, cdiv(q.N, 128), 128, 0, p);
            }
            launch_atr(0);
            // the forward point of this iterate exists only now: ghosts get it before the edge prox
            if (sharded) halo.exchange(HALO_VA0, PingPong<real>(P0b.p, P1b.p), PingPong<real>(), PingPong<real>(), ctl.p, PARITY_CUR);
        } else if (q.N < 0) {
            const int grid = std::min((q.V + 3) / 4, num_sms() * 4);
            if (vec > 1) PFDR_LAUNCH((ata_gemv_kernel<real, VM, 4>), grid, 256, 0, p);
            else PFDR_LAUNCH((ata_gemv_kernel<real, 1, 4>), grid, 256, 0, p);
        }
        if (q.Obj) {
            const size_t work = std::max((size_t) q.E, std::max((size_t) q.V, (size_t) std::max(q.N, 0)));
            PFDR_LAUNCH(objective_kernel<real>, std::min<unsigned>(cdiv(work, 256), num_sms() * 8), 256, 0, p);
        }
        launch_iter();
        if (sharded) {     // new boundary forward points and cut-edge copies to the neighbours (taken from the buffers the
                           // iteration kernel just wrote), evolution sums over all ranks, then the iteration is closed
            halo.exchange(q.N == 0 ? (HALO_VA0 | HALO_Z) : HALO_Z, PingPong<real>(P0b.p, P1b.p), PingPong<real>(),
                          PingPong<real>(Z0.p, Z1.p), ctl.p, PARITY_NEXT_IF_RAN, sums.p, 2);
            PFDR_LAUNCH(quad_finalize_kernel<real>, 1, 1, 0, p);
        }
    }

    // one thread per vertex, rows software-pipelined (quad_iter_pipe_kernel), 8 (float32) / 6 (float64) CTAs per SM
    int rowsW = 0;
    void launch_iter()
    {
        if (dia) { quad_dia_launch_iter<real>(p, ogrid); return; }
        constexpr int MB = sizeof(real) == 4 ? 8 : 6;
        // a slice wider than HUB_ROWS (a hub vertex): variant whose warps share the hub's rows
        if (g.max_width > HUB_ROWS) {
            if (p.thc) PFDR_LAUNCH((quad_iter_pipe_kernel<real, true, 2, true>), ogrid, 256, 0, p);
            else PFDR_LAUNCH((quad_iter_pipe_kernel<real, false, 2, true>), ogrid, 256, 0, p);
        } else if (p.pk) {
            if (p.thc) PFDR_LAUNCH((quad_iter_pipe_kernel<real, true, MB, false, 2>), ogrid, 256, 0, p);
            else PFDR_LAUNCH((quad_iter_pipe_kernel<real, false, MB, false, 2>), ogrid, 256, 0, p);
        } else if (p.pk2 && !p.thc) PFDR_LAUNCH((quad_iter_pipe_kernel<real, false, MB, false, 1>), ogrid, 256, 0, p);
        else if (p.thc) PFDR_LAUNCH((quad_iter_pipe_kernel<real, true, MB, false>), ogrid, 256, 0, p);
        else PFDR_LAUNCH((quad_iter_pipe_kernel<real, false, MB, false>), ogrid, 256, 0, p);
    }

    int launches_per_trip() const { return (q.N > 0 ? 3 : (q.N < 0 ? 1 : 0)) + (q.Obj ? 1 : 0) + 1 + (sharded ? (q.N > 0 ? 9 : 4) : 0); }

    // enqueue `trips` iterations; replays a captured CUDA graph when a whole batch is requested
    void enqueue_batch(int trips, int batch)
    {
        cudaStream_t st = ctx().stream;
        if (trips == batch && use_graph()) {
            if (!gexec || graph_trips != batch) {
                if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
                cudaGraph_t graph;
                const long long l0 = ctx().stats.launches, l1 = ctx().stats.loop_launches;
                PFDR_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                for (int i = 0; i < batch; i++) enqueue_trip();
                PFDR_CUDA(cudaStreamEndCapture(st, &graph));
                ctx().stats.launches = l0; ctx().stats.loop_launches = l1;
                PFDR_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
                cudaGraphDestroy(graph);
                graph_trips = batch;
            }
            PFDR_CUDA(cudaGraphLaunch(gexec, st));
            ctx().stats.launches += (long long) batch * launches_per_trip();
            ctx().stats.loop_launches += (long long) batch * launches_per_trip();
        } else {
            for (int i = 0; i < trips; i++) enqueue_trip();
        }
    }

    static bool use_graph() { return true; }            // batches of iterations replay as CUDA graphs

    void recondition()
    {
        if (!thc.p) thc.alloc(std::max<size_t>(nhalf, 1));
        if (!p.thc) { thc.zero(); p.thc = thc.p; }                       // first reconditioning since setup() / reset()
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }   // kernel parameters changed
        const bool linked = slab && link.world > 1;
        if (linked) {               // mean amplitude over all ranks' vertices
            p.sharded = 1;
            PFDR_LAUNCH(amp_kernel<real>, ogrid, 256, 0, p, 1);
            link.allreduce_sum(amp_sums.p, 2);
            PFDR_LAUNCH(amp_finish_kernel<real>, 1, 1, 0, p, 1);
            p.sharded = 0;
        } else PFDR_LAUNCH(amp_kernel<real>, vgrid, 256, 0, p, 1);
        if (dia) quad_dia_launch_recond<real>(p, vgrid);
        else PFDR_LAUNCH(recond_vertex_kernel<real>, vgrid, 256, 0, p);
        pack_slots();
        if (linked) slab_refresh(recond_parity);     // new metric, forward point and copies of the boundary -> neighbours' ghost zones
        PFDR_LAUNCH(resume_kernel<real>, 1, 1, 0, p, 0.01);                   // ref :458
    }

    // small-problem mode (quad_small_kernel): the whole solve in one launch of one thread-block cluster
    size_t small_smem() const
    {
        const size_t Np = std::max(q.N, 0), Vp = ((size_t) q.V + 3) & ~(size_t) 3;
        return 64 * sizeof(double) + (3 * Vp + ((Np + 3) & ~(size_t) 3) + 32 * Np) * sizeof(real);
    }
    // could the single-launch cluster kernels take this problem?  (decided before any layout exists: sizes only)
    bool small_candidate() const
    {
        const long long dense = q.N > 0 ? (long long) q.N * q.V : (q.N < 0 ? (long long) q.V * q.V : 0);
        if (q.N == 0 && 2LL * q.E > 49152) return false;
        return small_smem() <= 200 * 1024 && q.V <= SMALL_MAX_SLICES * 32 && q.E <= (1 << 19) && dense <= (1 << 22);
    }
    bool small_mode() const
    {
        const char *e = getenv("PFDR_B200_SMALL");        // read per call: the tests switch it
        if ((e && e[0] == '0') || sharded || dia) return false;
        const long long dense = q.N > 0 ? (long long) q.N * q.V : (q.N < 0 ? (long long) q.V * q.V : 0);
        // Regular graphs without dense operator: the large path costs ~7.5 us per iteration whatever the size, the
        // cluster kernels pass that near 50k slots (8000 vertices of a 3-D grid; profiles/r01_notes.md).  With a hub
        // (a slice wider than SMALL_WIDE rows) the large path walks the hub's row serially: cluster kernels whenever they fit.
        if (q.N == 0 && g.max_width <= SMALL_WIDE && g.nslots > 49152) return false;
        return small_smem() <= 200 * 1024 && g.nslices <= SMALL_MAX_SLICES && g.nslots <= (1 << 20) && q.E <= (1 << 19) && dense <= (1 << 22);
    }

    void run_small()
    {
        cudaStream_t st = ctx().stream;
        static thread_local int attr_dev = -1, cmax = 8;                      // function attributes are per device
        int dev = 0;
        PFDR_CUDA(cudaGetDevice(&dev));
        if (attr_dev != dev) {
            PFDR_CUDA(cudaFuncSetAttribute(quad_small_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            cmax = cudaFuncSetAttribute(quad_small_kernel<real>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 16 : 8;
            cudaGetLastError();
            attr_dev = dev;
        }
        int C = cmax;
        if (const char *e = getenv("PFDR_B200_SMALL_C")) C = std::max(1, std::min(cmax, atoi(e)));
        SmallArgs<real> a{};
        a.nslots = (int) g.nslots; a.nslices = (int) g.nslices;
        pos.alloc(std::max<size_t>(2 * (size_t) q.E, 1));
        ze.alloc(std::max<size_t>(2 * (size_t) q.E, 1));
        xsum.alloc(64);
        sparts.alloc(std::max<size_t>(16 * (size_t) std::max(q.N, 0), 1));
        a.pos = pos.p; a.Ze = ze.p; a.xs = xsum.p; a.parts = sparts.p;
        if (q.difRcd > (real) 0) {                               // (reset() clears p.thc, the buffer is kept)
            if (!thc.p) thc.alloc(std::max<size_t>(g.nslots, 1));
            thc.zero(); p.thc = thc.p;
        }
        const size_t sm = small_smem();
        ctx().in_loop = true;
        PFDR_CUDA(cudaEventRecord(ev_b, st));
        for (;; C /= 2) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(C); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = sm; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const cudaError_t rc = cudaLaunchKernelEx(&cfg, quad_small_kernel<real>, p, a);
            if (rc == cudaSuccess) break;
            cudaGetLastError();
            if (C == 1) PFDR_CUDA(rc);                       // a cluster this large may not be schedulable: halve it
        }
        ctx().count_launch();
        PFDR_CUDA(cudaEventRecord(ev_c, st));
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        ctx().in_loop = false;
        float ms = 0;
        PFDR_CUDA(cudaEventElapsedTime(&ms, ev_b, ev_c));
        ctx().stats.loop_ms = ms;
        ctx().stats.iterations = hctl[0].it;
    }

    // push variant (quad_push_kernel): needs the slice offsets on the host to balance the CTAs and size their shared memory
    bool run_push()
    {
        const char *e = getenv("PFDR_B200_PUSH");
        if ((e && e[0] == '0') || q.Obj != nullptr || q.difRcd > (real) 0) return false;
        cudaStream_t st = ctx().stream;
        static thread_local int attr_dev = -1, cmax = 8;                      // function attributes are per device
        int dev = 0;
        PFDR_CUDA(cudaGetDevice(&dev));
        if (attr_dev != dev) {
            PFDR_CUDA(cudaFuncSetAttribute(quad_push_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            cmax = cudaFuncSetAttribute(quad_push_kernel<real>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 16 : 8;
            cudaGetLastError();
            attr_dev = dev;
        }
        const int S = g.nslices;
        std::vector<long long> off((size_t) S + 1);
        PFDR_CUDA(cudaMemcpyAsync(off.data(), g.slice_off.p, ((size_t) S + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        // the largest cluster is the fastest at every size measured (profiles/r01_notes.md): phases are latency chains
        int C = cmax;
        if (const char *ec = getenv("PFDR_B200_SMALL_C")) C = std::max(1, std::min(cmax, atoi(ec)));
        const size_t Vq = (size_t) S * 32, Nq = q.N > 0 ? (((size_t) q.N + 3) & ~(size_t) 3) : 0;
        for (;; C /= 2) {
            // slices to CTAs, balanced by slots (a hub slice is as heavy as many ordinary ones)
            PushArgs<real> a{};
            a.nslices = S;
            long long nso = 0;
            a.slb[0] = 0;
            for (int c = 1, sl = 0; c <= C; c++) {
                const long long target = off[S] * c / C;
                while (sl < S && off[sl + 1] <= target) sl++;
                if (c == C) sl = S;
                a.slb[c] = std::max(sl, a.slb[c - 1]);
                nso = std::max(nso, off[a.slb[c]] - off[a.slb[c - 1]]);
            }
            for (int c = C + 1; c <= 16; c++) a.slb[c] = S;
            a.nso = (int) ((nso + 3) & ~3LL);
            const size_t sm = 64 + 512 + 512 + (5 * Vq + 2 * (size_t) a.nso + 17 * Nq) * sizeof(real);
            bool fits = sm <= 200 * 1024;
            for (int c = 0; c < C; c++) fits = fits && a.slb[c + 1] - a.slb[c] <= SMALL_MAX_SLICES;
            if (fits) {
                dstb.alloc(std::max<size_t>(2 * (size_t) q.E, 2));
                ze.alloc(std::max<size_t>(2 * (size_t) q.E, 2));
                ws.alloc(std::max<size_t>(g.nslots, 1));
                a.dst = dstb.p; a.Ze = ze.p; a.wS = ws.p;
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(C); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = sm; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                ctx().in_loop = true;
                PFDR_CUDA(cudaEventRecord(ev_b, st));
                const cudaError_t rc = cudaLaunchKernelEx(&cfg, quad_push_kernel<real>, p, a);
                if (rc == cudaSuccess) break;
                cudaGetLastError();
                ctx().in_loop = false;
                if (C == 1) return false;                    // a cluster this large may not be schedulable: halve it
            } else if (C == 1) return false;                 // does not fit the shared memory of one CTA either
        }
        ctx().count_launch();
        PFDR_CUDA(cudaEventRecord(ev_c, st));
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        ctx().in_loop = false;
        float ms = 0;
        PFDR_CUDA(cudaEventElapsedTime(&ms, ev_b, ev_c));
        ctx().stats.loop_ms = ms;
        ctx().stats.iterations = hctl[0].it;
        return true;
    }

    void run() override
    {
        cudaStream_t st = ctx().stream;
        if (small_mode()) { if (!run_push()) run_small(); return; }
        // batch: enough trips per host round trip to hide it, small enough to bound the no-op tail
        ctx().stats.layout = dia ? 1 : 0; ctx().stats.gpus = slab ? link.world : 1;
        const size_t work = (size_t) nhalf + (size_t) q.V + (size_t) std::abs((double) q.N) * q.V;
        int batch = work > (1u << 24) ? 4 : (work > (1u << 20) ? 16 : 64);
        batch = std::max(1, std::min(batch, q.itMax + 1));
        ctx().in_loop = true;
        PFDR_CUDA(cudaEventRecord(ev_b, st));
        // Pipelined: batch k is enqueued before the verdict of batch k-1 is read, so the GPU never
        // idles on the host.  Launches enqueued after the device froze are no-ops.
        progress.start(q.verbose);
        long long enqueued = 0, budget = (long long) q.itMax + 1 + 2LL * batch;
        int fresh_from = 0;             // batches older than this predate the last reconditioning
        for (int k = 0;; k++) {
            enqueue_batch(batch, batch);
            enqueued += batch;
            PFDR_CUDA(cudaMemcpyAsync(&hctl[k & 1], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
            PFDR_CUDA(cudaEventRecord(ev[k & 1], st));
            int verdict = 0;
            if (k >= 1 && k - 1 >= fresh_from) verdict = inspect((k - 1) & 1);
            if (verdict == 0 && enqueued >= budget) verdict = (k >= fresh_from) ? inspect(k & 1) : 0;
            if (progress.every && k >= 1 && k - 1 >= fresh_from)
                progress.upto(hctl[(k - 1) & 1].it, q.itMax, dDif.p, hctl[(k - 1) & 1].track != 0, hctl[(k - 1) & 1].tol, hctl[(k - 1) & 1].rcd);
            if (verdict == 1) break;
            if (verdict == 2) {         // frozen for a reconditioning; batches <= k were no-ops past the freeze
                PFDR_CUDA(cudaStreamSynchronize(st));
                recond_parity = hctl[k & 1].it & 1;            // (the device froze at the trip that asked for it)
                recondition();
                fresh_from = k + 1;
                budget = enqueued + (long long) q.itMax + 1 + 2LL * batch;
            } else if (enqueued >= budget)
                throw Error(PFDR_B200_ERR_CUDA, "PFDR loop did not terminate on the device");
        }
        PFDR_CUDA(cudaEventRecord(ev_c, st));
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        ctx().in_loop = false;
        float ms = 0;
        PFDR_CUDA(cudaEventElapsedTime(&ms, ev_b, ev_c));
        ctx().stats.loop_ms = ms;
        ctx().stats.iterations = hctl[0].it;
        if (slab && link.world > 1) {
            if (q.Obj) {             // objective of the final iterate: the last trip's parts were not carried by an iteration
                link.allreduce_sum(obj_part.p, 1);
                PFDR_LAUNCH(quad_obj_final_kernel<real>, 1, 1, 0, p);
                PFDR_CUDA(cudaStreamSynchronize(st));
            }
            slab_check();
        }
    }

    // waits for batch slot b: 0 = still running, 1 = finished, 2 = frozen for a reconditioning
    int inspect(int b)
    {
        PFDR_CUDA(cudaEventSynchronize(ev[b]));
        const Ctl &h = hctl[b];
        if (!h.hold) return 0;
        return h.state == ST_FINAL ? 1 : (h.state == ST_RECOND ? 2 : 0);
    }

    void fetch(void *Xo, int *ito, void *Objo, void *Difo) override
    {
        cudaStream_t st = ctx().stream;
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        const Ctl h = hctl[0];
        if (Xo) {
            if (is_device_ptr(Xo)) PFDR_LAUNCH(extract_x_kernel<real>, ogrid, 256, 0, p, (real *) Xo);
            else {
                Xout.alloc(p.Vown);
                PFDR_LAUNCH(extract_x_kernel<real>, ogrid, 256, 0, p, Xout.p);
                copy_out((real *) Xo, Xout.p, p.Vown);
            }
        }
        if (Objo && dObj.p) copy_out((real *) Objo, dObj.p, (size_t) h.it + 1);
        if (Difo && dDif.p && h.it > 0) copy_out((real *) Difo, dDif.p, (size_t) h.it);
        PFDR_CUDA(cudaStreamSynchronize(st));
        if (ito) *ito = h.it;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev_a, ev_b) == cudaSuccess) ctx().stats.setup_ms = ms;
        if (cudaEventElapsedTime(&ms, ev_b, ev_c) == cudaSuccess) ctx().stats.loop_ms = ms;
        cudaGetLastError();
        ctx().stats.iterations = h.it;
        if (q.verbose) {                                                       // final information (ref :535-539)
            printf("iteration %d (max. %d)\n", h.it, q.itMax);
            if (h.track) printf("iterate evolution %g (recond. %g; tol. %g)\n", h.dif, h.rcd, h.tol);
            printf("PFDR on B200: loop %.3f ms, %lld kernel launches\n", ctx().stats.loop_ms, ctx().stats.loop_launches);
            fflush(stdout);
        }
    }
};

template <typename real>
SolverBase *make_quadratic(const QuadProblem<real> &q)
{
    QuadSolver<real> *s = new QuadSolver<real>(q, true);
    try { s->setup(); } catch (...) { delete s; throw; }
    return s;
}
template SolverBase *make_quadratic<float>(const QuadProblem<float> &);
template SolverBase *make_quadratic<double>(const QuadProblem<double> &);

template <typename real>
void solve_quadratic(const QuadProblem<real> &q)
{
    const bool trace = getenv("PFDR_B200_TRACE") != nullptr;      // one line per call on stderr (who spends what inside cut-pursuit)
    const auto t0 = std::chrono::steady_clock::now();
    QuadSolver<real> s(q, false);
    s.setup();
    const auto t1 = std::chrono::steady_clock::now();
    s.run();
    const auto t2 = std::chrono::steady_clock::now();
    s.fetch(q.X, q.it, q.Obj, q.Dif);
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        fprintf(stderr, "[pfdr_b200] quadratic V=%d E=%d N=%d f%d: %d it, setup %.2f ms, run %.2f ms (device loop %.2f ms, %s), fetch %.2f ms\n",
                q.V, q.E, q.N, (int) sizeof(real) * 8, (int) ctx().stats.iterations, ms(t0, t1), ms(t1, t2), ctx().stats.loop_ms,
                s.small_mode() ? "cluster kernel" : "large path", ms(t2, t3));
    }
}

template void solve_quadratic<float>(const QuadProblem<float> &);
template void solve_quadratic<double>(const QuadProblem<double> &);

}  // namespace pfdr
