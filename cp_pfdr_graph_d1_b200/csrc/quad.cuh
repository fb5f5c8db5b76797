// quad.cuh - kernel parameter block and device helpers shared by the layouts of the quadratic family
// (quad.cu: half-edge table and the single-launch cluster kernels; quad_dia.cu: constant-offset graphs).
#pragma once
#include "common.cuh"
#include "slab.cuh"

namespace pfdr {

template <typename real> using pair_t = typename Pair<real>::type;

template <typename real>
struct QP {                       // kernel parameter block
    Ctl *ctl;
    int V, N, E;
    int Vown, sharded;            // vertices this rank owns (== V unless sharded: ghosts follow)
    int world, rank;              // sharded: sums holds one pair per rank, sums[2*rank..] is ours
    double *sums;                 // [2*world] evolution sums of every rank (sharded path), gathered with the halo exchange
    double *amp_sums;             // [2] amplitude sums (all-reduced)
    double *rsum;                 // [N] sum of A x over this rank's columns (sharded dense A)
    const long long *slice_off; const int *nbr, *twin, *hkey;
    const int4 *pk;               // per slot {nbr, twin, w, [threshold]} in 16 bytes: one vector load per row in the iteration kernel
    const int2 *pk2;              // per slot {nbr, twin}: float32 before a reconditioning (16 bytes would carry an unused word)
    real *w;                      // [slots] raw d1 weight of the half-edge's edge
    real *thc;                    // [slots] La_d1/w after a reconditioning (NULL: thc0 for all)
    real thc0;
    real *Z0, *Z1;                // [slots] edge copies, ping/pong by iteration parity
    real *X0, *X1;                // [V] iterate, ping/pong by iteration parity
    real *P0, *P1;                // [V] forward point 2x - Ga.grad(x), ping/pong (the only vertex value a
                                  //     neighbour gathers, hence stored apart from x)
    real *T, *Sinv, *Ga, *Ga0, *G, *Thl1, *R, *AtY;
    const real *Y, *A, *La_d1, *La_l1, *L;
    const int *Eu, *Ev;
    int Ltype, positivity, prox_kind, has_lo, has_hi;
    real lo, hi, rho, condMin;
    double *partial;
    real *dObj, *dDif;
    int nparts;                   // CTAs of ax_partial
    // constant-offset layout (dia.cuh; dia_nd > 0): Z0/Z1 are 2.ND planes of V edge copies; Wrow (row weight w/S) and,
    // after a reconditioning, thc (La_d1/w) hold 2.ND planes of V in the order the half-edges are summed
    int dia_nd, dia_off[3];
    int vbeg, vend;               // vertices processed by the iteration kernel: [vbeg, vend)
    const unsigned char *dmask;   // [V]
    const int *eoff;              // [V]
    real *Wrow;                   // [2.ND * V]
    // slab-sharded solve (slab.cuh): V counts the vertices held (owned + ghost zones), [vbeg, vend) are the owned ones
    SlabDev slab;
    SlabPeer<real> plo, phi;      // the ghost zones this rank writes: in the rank below / above (F = forward point P)
    double *obj_part;             // [1] this rank's part of the objective of the current iterate (Obj requested)
};

// explicitly rounded operations: the d1 prox below is shared by the kernels of both layouts and must give the same bits
// in each of them, whatever contraction (a*b+c -> fma) the compiler would choose in the surrounding code
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// d1 prox + relaxation of one half-edge (ref src/PFDR_graph_quadratic_d1_l1.cpp:467-489): new copy of the vertex that
// owns it.  xo, po_v, Tv: iterate, forward point and T of that vertex; zo / zt: its copy / the twin copy; pn, Tn: forward
// point and T of the neighbour; thc: La_d1 / w.  W_d1u.pu + W_d1v.pv = pt + wo.(po - pt) since the two weights sum to one.
template <typename real>
__device__ __forceinline__ real d1_prox_half(real rho, real xo, real po_v, real Tv, real zo, real zt, real pn, real Tn, real thc)
{
    const real tsum = add_rn(Tv, Tn), inv = fast_rcp(tsum);
    const real th = mul_rn(thc, tsum);                                    // Th_d1
    const real wo = mul_rn(Tn, inv), wt = mul_rn(Tv, inv);                // W_d1 own / twin side
    const real po = add_rn(po_v, -zo), pt = add_rn(pn, -zt), d0 = add_rn(po, -pt);
    const real b = add_rn(d0, -min(max(d0, -th), th));                    // soft threshold (ref :477-488), exact 0 inside
    return fma_rn(rho, add_rn(fma_rn(wt, b, fma_rn(wo, d0, pt)), -xo), zo);
}

constexpr int HUB_ROWS = 32;      // rows of a SELL slice walked by one thread per vertex; past that the warp shares a hub's row

template <typename real> __device__ __forceinline__ real rabs(real a) { return a < (real) 0 ? -a : a; }

// gradient of the smooth part at vertex v (x = X[v]); dense cases read what the A-kernels left in G
template <typename real>
__device__ __forceinline__ real grad_at(const QP<real> &p, int v, real x)
{
    if (p.N == 0) return p.A ? fma_rn(p.A[v], x, -p.Y[v]) : add_rn(x, -p.Y[v]);
    if (p.N < 0) return p.G[v] - p.Y[v];
    return p.G[v];
}

// forward point 2x - Ga.grad f(x) for diagonal / identity A (S1 + S4 + S6 of the next trip, ref :377-385, :441-445, :463-464)
template <typename real>
__device__ __forceinline__ real forward_point_of(real x, real ga, real g) { return fma_rn(-ga, g, mul_rn((real) 2, x)); }
template <typename real>
__device__ __forceinline__ real forward_point_diag(const QP<real> &p, int v, real x)
{
    return forward_point_of(x, p.Ga[v], grad_at(p, v, x));
}

// Sum NV doubles over the whole grid, deterministically: per-CTA partials, the last CTA to finish
// adds them in a fixed order.  Returns true (in every thread) only in that last CTA, v valid in thread 0.
template <int NV>
__device__ bool grid_sum(Ctl *ctl, double *partial, double (&v)[NV], double *smem)
{
    __shared__ int s_last;
    block_sum<NV>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) partial[(size_t) blockIdx.x * NV + i] = v[i];
        __threadfence();
        const unsigned t = atomicAdd(&ctl->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    const volatile double *vp = partial;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] += vp[(size_t) b * NV + i];
    }
    block_sum<NV>(acc, smem);
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = acc[i];
    if (threadIdx.x == 0) ctl->ticket = 0;
    return true;
}


template <typename real>
__device__ __forceinline__ real clamp_metric(const QP<real> &p, int v, real ga)
{   // ref :222-239
    const real bound = (real) 1.9 * ((real) 2 - p.rho);
    if (p.Ltype == PFDR_B200_SCAL || p.L == nullptr) {
        const real b = p.L ? bound / p.L[0] : bound;
        if (ga > b) ga = b;
    } else if (p.L[v] > (real) 0) {
        const real b = bound / p.L[v];
        if (ga > b) ga = b;
    }
    return ga;
}

// end of an iteration: evolution measure, iteration counter, stop / recondition decision (ref :429, :448, :515-529)
template <typename real>
__device__ void quad_finalize(const QP<real> &p, int it, double d, double c)
{
    Ctl *ctl = p.ctl;
    double dif = ctl->dif;
    if (ctl->track) {
        dif = (c > ctl->eps) ? d / c : d / ctl->eps;
        ctl->dif = dif;
        if (p.dDif) p.dDif[it] = (real) dif;
    }
    const int itn = it + 1;
    ctl->it = itn;
    if (itn == ctl->itMax || (real) dif < (real) ctl->tol) ctl->state = ST_FINAL;
    else if ((real) dif < (real) ctl->rcd) ctl->state = ST_RECOND;
}

// slab-sharded solve: what the kernel of trip `it` does before touching any data (slab.cuh).  Called by one thread of every
// CTA; all of them (on every rank) return the same state.
template <typename real>
__device__ int quad_slab_decide(const QP<real> &p, int it)
{
    Ctl *ctl = p.ctl;
    if (it == 0) return ST_RUN;
    SlabMail *me = p.slab.mail[p.slab.rank];
    const unsigned long long target = (p.slab.run_id << 32) | (unsigned long long) it;
    bool ok = true;
    if (p.slab.nb_lo >= 0) ok = slab_wait(&me->halo_flag[p.slab.nb_lo], target, me) && ok;
    if (p.slab.nb_hi >= 0) ok = slab_wait(&me->halo_flag[p.slab.nb_hi], target, me) && ok;
    const int track = ctl->track;
    const bool all = track || p.dObj;                               // someone needs every rank's sums of the previous trip
    double dif = ctl->dif, d = 0.0, c = 0.0, obj = 0.0;
    if (ok && all) {
        for (int q = 0; q < p.slab.world; q++) {                    // rank order: same bits everywhere
            ok = slab_wait(&me->sum_flag[q], target, me) && ok;
            d += me->sums[(it - 1) & 1][q][0]; c += me->sums[(it - 1) & 1][q][1]; obj += me->sums[(it - 1) & 1][q][2];
        }
        if (track) dif = (c > ctl->eps) ? d / c : d / ctl->eps;     // ref :515-529
    }
    const bool recond_done = ctl->scratch[1] == (double) (it + 1);  // this trip already reconditioned (the test is made once per trip)
    int state = ST_RUN;
    if (!ok || it == ctl->itMax || (real) dif < (real) ctl->tol) state = ST_FINAL;
    else if (!recond_done && (real) dif < (real) ctl->rcd) state = ST_RECOND;
    if (blockIdx.x == 0) {
        if (track && ok) { ctl->dif = dif; if (p.dDif) p.dDif[it - 1] = (real) dif; }
        if (p.dObj && ok) p.dObj[it - 1] = (real) obj;              // objective of the iterate the previous trip started from
        if (state != ST_RUN) { ctl->state = state; ctl->hold = 1; }
    }
    return state;
}

}  // namespace pfdr
