// quad_dia.cu - the quadratic-family kernels on constant-offset graphs (dia.cuh).  Same arithmetic in the same order as
// the half-edge-table kernels of quad.cu (bit-identical results); one thread per vertex, every address computed from
// the vertex number: the 2.ND half-edges of a vertex need 5 independent unit-stride loads each (own copy, twin copy,
// neighbour's forward point and T, row weight), all issued before the first use - no index stream (24 B per edge
// less HBM traffic than the table) and no dependent index -> gather chain.
// ref src/PFDR_graph_quadratic_d1_l1.cpp:463-529 (iteration), :57-268 (preconditioning).
#include "dia.cuh"
#include "quad_dia.cuh"

namespace pfdr {

namespace {

template <int ND>
__device__ __forceinline__ int dia_edge_id(const int *__restrict__ eoff, const unsigned char *__restrict__ dmask,
                                           int v, unsigned m, int h, const int (&off)[3])
{
    typedef DiaHalf<ND> H;
    const int i = H::idx(h);
    if (H::out(h)) return eoff[v] + __popc(m & ((1u << i) - 1u));
    const int nb = v - off[i];
    return eoff[nb] + __popc((unsigned) dmask[nb] & ((1u << i) - 1u));
}

// first preconditioning + initial state (as init_vertex_kernel of quad.cu): w = c La_d1 (:160-166), S, Ga, T, Th_l1, z_e = x
template <typename real, int ND>
__global__ void __launch_bounds__(256) quad_dia_init_kernel(QP<real> p, const real *__restrict__ X0)
{
    typedef DiaHalf<ND> H;
    const int v = p.vbeg + blockIdx.x * blockDim.x + threadIdx.x;   // owned vertices (slab-sharded: ghosts come from their owners)
    if (v >= p.vend) return;
    const size_t Vp = (size_t) p.V;
    const real c = (real) p.ctl->amp;
    const real x = X0[v];
    const unsigned m = p.dmask[v];
    const int off[3] = {p.dia_off[0], p.dia_off[1], p.dia_off[2]};
    real S = 0, wv[2 * ND];
#pragma unroll
    for (int h = 0; h < 2 * ND; h++) {
        wv[h] = 0;
        if ((m >> H::mask_bit(h)) & 1u) {
            wv[h] = mul_rn(c, p.La_d1[dia_edge_id<ND>(p.eoff, p.dmask, v, m, h, off)]);
            S = add_rn(S, wv[h]);
            p.Z0[(size_t) H::own_plane(h) * Vp + v] = x;
        }
    }
    real ga = add_rn(p.Ga0[v], S);
    if (p.La_l1) ga = fma_rn(c, p.La_l1[v], ga);
    ga = clamp_metric(p, v, (real) 1 / ga);
    const real sinv = (real) 1 / S;
    p.Ga[v] = ga; p.Sinv[v] = sinv; p.T[v] = mul_rn(ga, S);
#pragma unroll
    for (int h = 0; h < 2 * ND; h++) p.Wrow[(size_t) h * Vp + v] = mul_rn(wv[h], sinv);
    if (p.La_l1) p.Thl1[v] = mul_rn(ga, p.La_l1[v]);
    p.X0[v] = x;
    p.P0[v] = (p.N == 0) ? forward_point_of(x, ga, grad_at(p, v, x)) : (real) 0;
}

// reconditioning (as recond_vertex of quad.cu, ref :89-99, :167-250)
template <typename real, int ND>
__global__ void __launch_bounds__(256) quad_dia_recond_kernel(QP<real> p)
{
    typedef DiaHalf<ND> H;
    const int v = p.vbeg + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.vend) return;
    const size_t Vp = (size_t) p.V;
    const int cur = p.ctl->it & 1;
    const real *X = cur ? p.X1 : p.X0;
    real *Z = cur ? p.Z1 : p.Z0, *Pout = cur ? p.P1 : p.P0;
    real c = (real) p.ctl->amp;
    const real x = X[v], g = grad_at(p, v, x);
    const real ga_old = p.Ga[v];
    const unsigned m = p.dmask[v];
    const int off[3] = {p.dia_off[0], p.dia_off[1], p.dia_off[2]};
    real S = 0, wv[2 * ND];
#pragma unroll
    for (int h = 0; h < 2 * ND; h++) {
        wv[h] = 0;
        if (!((m >> H::mask_bit(h)) & 1u)) continue;
        const int n = H::out(h) ? v + off[H::idx(h)] : v - off[H::idx(h)];
        const size_t zk = (size_t) H::own_plane(h) * Vp + v, hk = (size_t) h * Vp + v;
        const real Wold = p.Wrow[hk];
        Z[zk] = mul_rn(Wold / ga_old, add_rn(fma_rn(-ga_old, g, x), -Z[zk]));  // ref :96-97
        const real xn = X[n];
        real a = rabs(x), b = rabs(xn), d = rabs(x - xn);                       // ref :172-186
        if (a < b) a = b;
        if (a < c) a = c;
        a *= p.condMin;
        if (d < a) d = a;
        wv[h] = p.La_d1[dia_edge_id<ND>(p.eoff, p.dmask, v, m, h, off)] / d;
        p.thc[hk] = d; S = add_rn(S, wv[h]);
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
#pragma unroll
    for (int h = 0; h < 2 * ND; h++) {
        const real W = mul_rn(wv[h], sinv);
        p.Wrow[(size_t) h * Vp + v] = W;
        if ((m >> H::mask_bit(h)) & 1u) {
            const size_t zk = (size_t) H::own_plane(h) * Vp + v;
            Z[zk] = fma_rn(-ga, add_rn(g, Z[zk] / W), x);                      // ref :247-248
        }
    }
    Pout[v] = forward_point_of(x, ga, g);
}

// fused iteration: S6-S10 of SURVEY 3.3 + the next forward point when A is diagonal.
// Persistent warps (one wave of resident CTAs): a warp walks chunks of 32 consecutive vertices, no block-level
// synchronisation inside the loop; loads never depend on the edge mask (neighbour index clamped into [0, V), a
// half-edge that does not exist is dropped at compute time); G half-edges (5 loads each) are requested at a time.
template <typename real, bool PER_SLOT_TH, int ND, int G, int MINB, bool SLAB = false>
__global__ void __launch_bounds__(256, MINB) quad_dia_iter_kernel(QP<real> p)
{
    typedef DiaHalf<ND> H;
    __shared__ double smem[64];
    Ctl *ctl = p.ctl;
    const CtlView cv = load_ctl(ctl);
    if (cv.hold != 0 || cv.state != ST_RUN) {
        if (cv.hold == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctl->hold = 1;
        return;
    }
    const int it = cv.it, cur = it & 1;
    if (SLAB) {             // neighbours' pushes of the previous trip have landed; stop / recondition decision over all ranks
        __shared__ int s_state;
        if (threadIdx.x == 0) s_state = quad_slab_decide(p, it);
        __syncthreads();
        if (s_state != ST_RUN) return;
    }
    const real *__restrict__ Zin = cur ? p.Z1 : p.Z0;
    real *__restrict__ Zout = cur ? p.Z0 : p.Z1;
    const real *__restrict__ Xc = cur ? p.X1 : p.X0;
    real *__restrict__ Xn = cur ? p.X0 : p.X1;
    const real *__restrict__ Pc = cur ? p.P1 : p.P0;
    real *__restrict__ Pn = cur ? p.P0 : p.P1;
    const real *__restrict__ Tg = p.T;
    const real *__restrict__ Wrow = p.Wrow;
    const real rho = p.rho, thc0 = p.thc0;
    const size_t Vp = (size_t) p.V;
    const int Vm1 = p.V - 1;
    const int off[3] = {p.dia_off[0], p.dia_off[1], p.dia_off[2]};
    const int stride = (int) (gridDim.x * blockDim.x);              // a warp takes 32 consecutive vertices per trip
    const int nown = p.vend - p.vbeg, lane = threadIdx.x & 31;
    // slab-sharded: the vertices within `band` of a neighbouring rank come first (work index -> vertex below), so that the
    // neighbours learn early that their ghost zones are written; the interior follows and touches no ghost data
    const int lowB = SLAB ? p.slab.lowB : 0, nB = SLAB ? p.slab.lowB + p.slab.highB : 0;
    int bcount = 0;
    double red0 = 0.0, red1 = 0.0;
    for (int i0 = (int) (blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < nown; i0 += stride) {      // warp-uniform
        const bool bchunk = SLAB && i0 < nB;
        if (SLAB && !bchunk && bcount) { slab_boundary_done(p.slab, &ctl->bdone, bcount, it); bcount = 0; }
        if (bchunk) bcount++;
        const int i = i0 + lane;
        if (i >= nown) continue;
        const int v = !SLAB ? p.vbeg + i : (i < lowB ? p.vbeg + i : (i < nB ? p.vend - nB + i : p.vbeg + lowB + (i - nB)));
        const unsigned m = __ldg(p.dmask + v);
        const real xo = Xc[v], po_v = Pc[v], Tv = Tg[v];
        real acc = 0;
#pragma unroll
        for (int h0 = 0; h0 < 2 * ND; h0 += G) {
            real zo[G], zt[G], pn[G], Tn[G], wr[G], thc[G];
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int h = h0 + u;
                if (h < 2 * ND) {
                    const int n = H::out(h) ? min(v + off[H::idx(h)], Vm1) : max(v - off[H::idx(h)], 0);
                    zo[u] = Zin[(size_t) H::own_plane(h) * Vp + v];
                    zt[u] = Zin[(size_t) H::twin_plane(h) * Vp + n];
                    pn[u] = Pc[n]; Tn[u] = Tg[n];
                    wr[u] = __ldg(Wrow + ((size_t) h * Vp + v));
                    thc[u] = PER_SLOT_TH ? __ldg(p.thc + ((size_t) h * Vp + v)) : thc0;
                }
            }
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int h = h0 + u;
                if (h < 2 * ND && ((m >> H::mask_bit(h)) & 1u)) {
                    const real zn = d1_prox_half<real>(rho, xo, po_v, Tv, zo[u], zt[u], pn[u], Tn[u], thc[u]);
                    Zout[(size_t) H::own_plane(h) * Vp + v] = zn;
                    acc = fma_rn(wr[u], zn, acc);
                    if (SLAB) {     // the other end belongs to the rank above / below: its twin copy lives there too
                        if (H::out(h)) { if (v + off[H::idx(h)] >= p.vend) p.phi.Z[cur ^ 1][(size_t) H::own_plane(h) * p.phi.plane + (size_t) ((long long) v + p.phi.shift)] = zn; }
                        else if (v - off[H::idx(h)] < p.vbeg) p.plo.Z[cur ^ 1][(size_t) H::own_plane(h) * p.plo.plane + (size_t) ((long long) v + p.plo.shift)] = zn;
                    }
                }
            }
        }
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
        Xn[v] = x;
        if (p.N == 0) {
            const real fp = forward_point_diag(p, v, x);
            Pn[v] = fp;
            if (SLAB) {     // boundary vertices: the neighbouring rank gathers this forward point from its ghost zone
                if (p.plo.F[0] && v < p.vbeg + p.slab.band) { p.plo.F[cur ^ 1][(long long) v + p.plo.shift] = fp; p.plo.X[cur ^ 1][(long long) v + p.plo.shift] = x; }
                if (p.phi.F[0] && v >= p.vend - p.slab.band) { p.phi.F[cur ^ 1][(long long) v + p.phi.shift] = fp; p.phi.X[cur ^ 1][(long long) v + p.phi.shift] = x; }
            }
        }
        const real dx = xo - x;
        red0 += (double) (dx * dx); red1 += (double) (x * x);
    }
    if (SLAB && bcount) slab_boundary_done(p.slab, &ctl->bdone, bcount, it);
    double tot[2] = {red0, red1};
    if (grid_sum<2>(ctl, p.partial, tot, smem) && threadIdx.x == 0) {
        if (SLAB) {
            ctl->it = it + 1;
            const double v4[SLAB_NSUM] = {tot[0], tot[1], p.obj_part ? *p.obj_part : 0.0, 0.0};
            if (ctl->track || p.dObj) slab_post(p.slab, v4, (unsigned long long) (it + 1));   // only when someone waits for them
        } else if (p.sharded) { p.sums[2 * p.rank] = tot[0]; p.sums[2 * p.rank + 1] = tot[1]; }
        else quad_finalize(p, it, tot[0], tot[1]);
    }
}

}  // namespace

#define PFDR_ND_SWITCH(CALL)                       \
    do {                                           \
        if (p.dia_nd == 1) { CALL(1); }            \
        else if (p.dia_nd == 2) { CALL(2); }       \
        else { CALL(3); }                          \
    } while (0)

template <typename real>
void quad_dia_launch_init(const QP<real> &p, const real *X0, int grid)
{
#define CALL(N_) PFDR_LAUNCH((quad_dia_init_kernel<real, N_>), grid, 256, 0, p, X0)
    PFDR_ND_SWITCH(CALL);
#undef CALL
}
template <typename real>
void quad_dia_launch_recond(const QP<real> &p, int grid)
{
#define CALL(N_) PFDR_LAUNCH((quad_dia_recond_kernel<real, N_>), grid, 256, 0, p)
    PFDR_ND_SWITCH(CALL);
#undef CALL
}
template <typename real>
void quad_dia_launch_iter(const QP<real> &p, int grid)
{
    constexpr int MINB = sizeof(real) == 4 ? 4 : 3;
    constexpr int G = sizeof(real) == 4 ? 6 : 3;         // half-edges requested at a time (5 loads each)
    const int g = std::min(grid, num_sms() * MINB);      // one wave of resident CTAs
#define CALL(N_)                                                                                         \
    if (p.slab.on && p.thc) PFDR_LAUNCH((quad_dia_iter_kernel<real, true, N_, G, MINB, true>), g, 256, 0, p);   \
    else if (p.slab.on) PFDR_LAUNCH((quad_dia_iter_kernel<real, false, N_, G, MINB, true>), g, 256, 0, p);    \
    else if (p.thc) PFDR_LAUNCH((quad_dia_iter_kernel<real, true, N_, G, MINB>), g, 256, 0, p);          \
    else PFDR_LAUNCH((quad_dia_iter_kernel<real, false, N_, G, MINB>), g, 256, 0, p)
    PFDR_ND_SWITCH(CALL);
#undef CALL
}

template void quad_dia_launch_init<float>(const QP<float> &, const float *, int);
template void quad_dia_launch_init<double>(const QP<double> &, const double *, int);
template void quad_dia_launch_recond<float>(const QP<float> &, int);
template void quad_dia_launch_recond<double>(const QP<double> &, int);
template void quad_dia_launch_iter<float>(const QP<float> &, int);
template void quad_dia_launch_iter<double>(const QP<double> &, int);

}  // namespace pfdr
