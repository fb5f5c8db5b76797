// simplex.cu - PFDR for  loss(P;Q) + sum_e La_d1 sum_k |P_uk - P_vk|  with every P_v on the simplex,
// the B200-native replacement of src/PFDR_graph_loss_d1_simplex.cpp (+ src/proj_simplex_metric.cpp).
//
// One fused kernel per iteration.  A lane owns LPT consecutive labels of one vertex (LPT = 4 when K % 4 == 0:
// every access to a K-vector is then a 16/32-byte vector load; LPT = 1 otherwise), TPV = next power of two
// >= K/LPT lanes cooperate on a vertex and C = 32/TPV vertices share a warp and one SELL-C slice of the
// half-edge table (graph.cuh).  U half-edge rows are in flight per lane (loads of the whole chunk first,
// then the gathers, then the arithmetic).  Per trip and per (half-edge, label):
//   - the explicit step is already folded into the stored forward point FP = 2P - Ga.grad (ref :569-587),
//     which is the only vertex value a neighbour gathers (P and FP are separate arrays for that reason);
//   - d1 prox + relaxation (:590-634) for the lane's own half-edge copy, ping/pong buffers;
//   - weighted average over the vertex's rows in edge order (:637-648);
//   - projection on the simplex in the diagonal metric (:651): the lanes of a vertex run Michelot's
//     active-set iteration in parallel (two group reductions per pass) - same projection as the reference's
//     sequential recursion (proj_simplex_metric.cpp:40-81), which cost ~60 % of the v0 kernel's instructions;
//   - next (P, FP) and the evolution sums (:654-691).
// Edge constants are recomputed from T = Ga.S as in quad.cu:  W = w/S[u], W_d1u = T[v]/(T[u]+T[v]),
// Th_d1 = (La_d1/w)(T[u]+T[v]); before any reconditioning w = La_d1[e] for all labels (:207-208), so
// La_d1/w = 1, S is per vertex, and T doubles as the projection metric (the projection is invariant to a
// positive rescaling of the metric, and the reference's normalisation Ga/max Ga (:362-369) is one).
#include <type_traits>
#include "graph.cuh"
#include "solvers.cuh"
#include <chrono>

#include "simplex.cuh"
#include "dia.cuh"
#include "simplex_dia.cuh"

namespace pfdr {

// ---------------------------------------------------------------------------------------------
// first preconditioning (recond = 0, ref :160-370 with Zu == NULL) and reconditioning (recond = 1):
// everything is local to a vertex row, so one kernel does it all.
// ---------------------------------------------------------------------------------------------
template <typename real, int LPT>
__global__ void __launch_bounds__(256) simplex_precond_kernel(SP<real> p, const real *__restrict__ Pinit, int recond)
{
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slice >= ((long long) p.V + p.C - 1) >> p.logC) return;     // whole warp leaves together
    const int K = p.K, Kv = p.Kv, TPV = p.TPV;
    const int vl = lane >> p.logTPV, k0 = (lane & (TPV - 1)) * LPT;
    const int v = (int) (slice << p.logC) + vl;
    const bool on = (v < p.V) && (k0 < Kv);
    const size_t vk = (size_t) (on ? v : 0) * K + (on ? k0 : 0);
    const int cur = recond ? (p.ctl->it & 1) : 0;
    real *Pb = cur ? p.P1 : p.P0, *Fb = cur ? p.F1 : p.F0;
    real *Z = cur ? p.Z1 : p.Z0;
    const bool perk_old = recond && (p.ctl->scratch[0] != 0.0);   // weights already per label before this call
    const long long base = p.slice_off[slice];
    const int width = (int) ((p.slice_off[slice + 1] - base) >> p.logC);
    real P[LPT], q[LPT], S[LPT], ga_old[LPT], gaq_old[LPT], sinv_old[LPT];
#pragma unroll
    for (int i = 0; i < LPT; i++) {
        const bool li = on && k0 + i < Kv;
        P[i] = li ? (recond ? Pb[vk + i] : Pinit[vk + i]) : (real) 0;
        q[i] = li ? p.Q[vk + i] : (real) 0;
        S[i] = 0;
        ga_old[i] = (recond && li) ? p.GaU[vk + i] : (real) 1;
        gaq_old[i] = (recond && li) ? p.GaQ[vk + i] : (real) 0;
        sinv_old[i] = (recond && li) ? (perk_old ? p.Sinvk[vk + i] : p.Sinv[v]) : (real) 0;
    }
    for (int j = 0; j < width; j++) {
        const long long s = base + ((long long) j << p.logC) + vl;
        const int n = p.nbr[s];
        if (n >= 0 && on) {
            const real la = p.La_d1[p.hkey[s] >> 1];
            if (!recond && k0 == 0) p.w[s] = la;                               // ref :207-208
#pragma unroll
            for (int i = 0; i < LPT; i++) if (k0 + i < Kv) {
                const size_t sk = (size_t) s * K + k0 + i;
                if (!recond) { S[i] += la; Z[sk] = P[i]; }                       // ref :419-430
                else {
                    const real Wold = (perk_old ? p.wk[sk] : p.w[s]) * sinv_old[i];
                    Z[sk] = (Wold / ga_old[i]) * (explicit_point(p.loss, p.aK1, P[i], gaq_old[i], q[i]) - Z[sk]);   // :137-156
                    real d = sabs(P[i] - Pb[(size_t) n * K + k0 + i]);                                             // :210-213
                    if (d < p.condMin) d = p.condMin;
                    const real wv = la / d;
                    p.wk[sk] = wv; p.thk[sk] = d; S[i] += wv;
                }
            }
        }
    }
    real ga[LPT], gloc = 0;
#pragma unroll
    for (int i = 0; i < LPT; i++) {
        const bool li = on && k0 + i < Kv;
        ga[i] = li ? metric(p, v, P[i], q[i], S[i]) : (real) 0;
        gloc = max(gloc, ga[i]);
    }
    const real gmax = group_max(gloc, TPV);                                    // ref :362-369
    if (on) {
        if (!recond && k0 == 0) p.Sinv[v] = (real) 1 / S[0];
#pragma unroll
        for (int i = 0; i < LPT; i++) if (k0 + i < Kv) {
            const real sinv = (real) 1 / S[i], gaq = metric_times_info(p, v, ga[i], q[i]);
            if (p.GaU) p.GaU[vk + i] = ga[i];
            p.T[vk + i] = ga[i] * S[i]; p.GaQ[vk + i] = gaq;
            if (recond) {
                p.Sinvk[vk + i] = sinv; p.Gn[vk + i] = ga[i] / gmax;
                for (int j = 0; j < width; j++) {
                    const long long s = base + ((long long) j << p.logC) + vl;
                    if (p.nbr[s] >= 0) {
                        const size_t sk = (size_t) s * K + k0 + i;
                        const real W = p.wk[sk] * sinv;
                        if (p.loss == LOSS_QUAD) Z[sk] = P[i] - gaq * (P[i] - q[i] + Z[sk] / W);                    // :348-350
                        else Z[sk] = explicit_point(p.loss, p.aK1, P[i], gaq, q[i]) - (ga[i] / W) * Z[sk];        // :345-353
                    }
                }
            }
            Pb[vk + i] = P[i];
            Fb[vk + i] = forward_point(p.loss, p.aK1, P[i], gaq, q[i]);
        }
    }
    if (!recond && p.lab) {                                                   // ref :447-462
        VecN<real, LPT> pv;
#pragma unroll
        for (int i = 0; i < LPT; i++) pv.v[i] = P[i];
        const int i = group_argmax<real, LPT>(pv, on ? k0 : Kv, Kv, TPV);
        if (on && k0 == 0) p.lab[v] = i;
    }
}

template <typename real>
__global__ void simplex_resume_kernel(SP<real> p)
{
    p.ctl->rcd *= 0.1; p.ctl->state = ST_RUN; p.ctl->hold = 0; p.ctl->scratch[0] = 1.0;     // ref :563
    p.ctl->scratch[1] = (double) (p.ctl->it + 1);            // slab-sharded: this trip's reconditioning test has been made
}

template <typename real>
__global__ void simplex_finalize_kernel(SP<real> p)
{
    if (p.ctl->hold != 0 || p.ctl->state != ST_RUN) return;
    double s = 0.0;
    for (int r = 0; r < p.world; r++) s += p.sums[r];            // rank order: same bits everywhere
    simplex_finalize(p, p.ctl->it, s);
}

constexpr int SIMPLEX_HUB_ROWS = 32;   // rows of a slice walked by one lane group per vertex; past that the warp shares a hub's rows

template <typename real, int LPT, bool PERK, int LOSS, int MINB, bool IDX32, bool HUB>
__global__ void __launch_bounds__(256, MINB) simplex_iter_kernel(SP<real> p)
{
    typedef VecN<real, LPT> vec;
    __shared__ double smem[32];
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
    const real *__restrict__ Fc = cur ? p.F1 : p.F0;
    real *__restrict__ Pn = cur ? p.P0 : p.P1;
    real *__restrict__ Fn = cur ? p.F0 : p.F1;
    const real *__restrict__ Tg = p.T;
    const int *__restrict__ nbr = p.nbr;
    const int *__restrict__ twin = p.twin;
    const int K = p.K, TPV = p.TPV, C = p.C, lane = threadIdx.x & 31;
    const long long wslot = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nslices = ((long long) p.Vown + C - 1) >> p.logC;
    const real rho = p.rho;
    real red[1] = {(real) 0};
    if (wslot < nslices) {                                                    // warp-uniform
        const long long slice = p.order ? (long long) p.order[wslot] : wslot;
        const int vl = lane >> p.logTPV, k0 = (lane & (TPV - 1)) * LPT;
        const int v = (int) (slice << p.logC) + vl;
        const bool on = (v < p.Vown) && (k0 < K);    // LPT == 4 requires K % 4 == 0: all LPT labels valid
        const size_t vk = (size_t) (on ? v : 0) * K + (on ? k0 : 0);
        const vec pold = ldv<real, LPT>(Pc + vk), fown = ldv<real, LPT>(Fc + vk), Tv = ldv<real, LPT>(Tg + vk);
        vec Sv;
        if (PERK) Sv = ldv<real, LPT>(p.Sinvk + vk); else Sv = splat<real, LPT>(p.Sinv[on ? v : 0]);
        const long long base = p.slice_off[slice];
        const int width = (int) ((p.slice_off[slice + 1] - base) >> p.logC);
        const long long s0 = base + vl;
        vec acc = splat<real, LPT>((real) 0);
        // Software pipeline over the vertex's rows: the indices (and the lane's own copy) of row j+1 are
        // requested while row j's gathers are in flight, so a row costs one memory round trip, not two.
        // Element offsets are 32-bit when every array has < 2^32 elements (IDX32): one IMAD per gather address.
        typedef typename std::conditional<IDX32, unsigned, size_t>::type idx_t;
        const idx_t rowK = (idx_t) C * (idx_t) K;
        int s = (int) s0;                                                // slot of this lane's vertex in row j
        idx_t zoff = (idx_t) s0 * (idx_t) K + (idx_t) k0;               // its K-vector of edge copies
        int n_c = -1, tw_c = 0;
        real w_c = 0;
        vec zo_c = splat<real, LPT>((real) 0);
        const int wl = HUB ? min(width, SIMPLEX_HUB_ROWS) : width;       // HUB: later rows are shared by the warp, below
        if (on && wl > 0) {
            n_c = nbr[s]; tw_c = twin[s];
            if (!PERK) w_c = p.w[s];
            zo_c = ldv<real, LPT>(Zin + zoff);
        }
        for (int j = 0; j < wl; j++) {
            const bool live = n_c >= 0;
            vec zt, fn, tn, wvk, th;
            if (live) {
                const idx_t no = (idx_t) n_c * (idx_t) K + (idx_t) k0;
                zt = ldv<real, LPT>(Zin + ((idx_t) tw_c * (idx_t) K + (idx_t) k0));
                fn = ldv<real, LPT>(Fc + no);
                if (LOSS != LOSS_LIN) tn = ldv<real, LPT>(Tg + no);
                if (PERK) { wvk = ldv<real, LPT>(p.wk + zoff); th = ldv<real, LPT>(p.thk + zoff); }
            }
            int n_n = -1, tw_n = 0;
            real w_n = 0;
            vec zo_n = zo_c;
            if (on && j + 1 < wl) {
                n_n = nbr[s + C]; tw_n = twin[s + C];
                if (!PERK) w_n = p.w[s + C];
                zo_n = ldv<real, LPT>(Zin + (zoff + rowK));
            }
            if (live) {
                const vec zn = prox_row<real, LPT, PERK, LOSS>(rho, fown, pold, Tv, Sv, zo_c, zt, fn, tn, wvk, th, w_c, acc);
                if (p.stream_z) stv_cs<real, LPT>(Zout + zoff, zn); else stv<real, LPT>(Zout + zoff, zn);
            }
            n_c = n_n; tw_c = tw_n; w_c = w_n; zo_c = zo_n;
            s += C; zoff += rowK;
        }
        if (HUB && width > SIMPLEX_HUB_ROWS) {
            // A hub vertex (a cut-pursuit reduced graph's background component): the C vertex groups of the warp
            // share the rows past SIMPLEX_HUB_ROWS of every hub of the slice, each group doing the full prox of one
            // row for its labels; the hub's state is broadcast by shuffles, the partial sums folded across groups.
            const unsigned full = 0xffffffffu;
            const int kl = lane & (TPV - 1);
            const bool kok = k0 < K;
            const unsigned m = __ballot_sync(full, v < p.Vown && nbr[(int) base + SIMPLEX_HUB_ROWS * C + vl] >= 0);
            for (int c = 0; c < C; c++) {
                if (!((m >> (c * TPV)) & 1u)) continue;                          // warp-uniform
                const int src = c * TPV + kl;
                vec pold_c, fown_c, Tv_c, Sv_c, part = splat<real, LPT>((real) 0);
#pragma unroll
                for (int i = 0; i < LPT; i++) {
                    pold_c.v[i] = __shfl_sync(full, pold.v[i], src); fown_c.v[i] = __shfl_sync(full, fown.v[i], src);
                    Tv_c.v[i] = __shfl_sync(full, Tv.v[i], src); Sv_c.v[i] = __shfl_sync(full, Sv.v[i], src);
                }
                constexpr int HU = 2;                                             // rows in flight per lane group
                for (int j0 = SIMPLEX_HUB_ROWS + vl; j0 < width && kok; j0 += HU * C) {
                    int sh[HU], n[HU], tw[HU];
                    real wh[HU];
                    vec zo[HU], zt[HU], fn[HU], tn[HU], wvk[HU], th[HU];
#pragma unroll
                    for (int u = 0; u < HU; u++) {
                        const int j = j0 + u * C;
                        sh[u] = (int) base + j * C + c;
                        n[u] = j < width ? nbr[sh[u]] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < HU; u++) {
                        if (n[u] >= 0) {
                            const idx_t zo_off = (idx_t) sh[u] * (idx_t) K + (idx_t) k0, no = (idx_t) n[u] * (idx_t) K + (idx_t) k0;
                            tw[u] = twin[sh[u]];
                            wh[u] = PERK ? (real) 0 : p.w[sh[u]];
                            zo[u] = ldv<real, LPT>(Zin + zo_off);
                            fn[u] = ldv<real, LPT>(Fc + no);
                            if (LOSS != LOSS_LIN) tn[u] = ldv<real, LPT>(Tg + no);
                            if (PERK) { wvk[u] = ldv<real, LPT>(p.wk + zo_off); th[u] = ldv<real, LPT>(p.thk + zo_off); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < HU; u++)
                        if (n[u] >= 0) zt[u] = ldv<real, LPT>(Zin + ((idx_t) tw[u] * (idx_t) K + (idx_t) k0));
#pragma unroll
                    for (int u = 0; u < HU; u++) {
                        if (n[u] >= 0) {
                            const vec zn = prox_row<real, LPT, PERK, LOSS>(rho, fown_c, pold_c, Tv_c, Sv_c, zo[u], zt[u], fn[u], tn[u], wvk[u], th[u], wh[u], part);
                            stv<real, LPT>(Zout + ((idx_t) sh[u] * (idx_t) K + (idx_t) k0), zn);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < LPT; i++) {
                    for (int o = TPV; o < 32; o <<= 1) part.v[i] += __shfl_xor_sync(full, part.v[i], o);
                    if (vl == c) acc.v[i] += part.v[i];
                }
            }
        }
        red[0] = (real) simplex_epilogue<real, LPT, PERK, LOSS>(p, ctl, on, v, k0, vk, acc, Tv, pold, Pn, Fn);
    }
    double tot[1];
    if (grid_sum_fast<1, real>(ctl, p.partial, red, tot, smem) && threadIdx.x == 0) {
        if (p.sharded) p.sums[p.rank] = tot[0];
        else simplex_finalize(p, it, tot[0]);
    }
}


// objective (ref :477-544)
template <typename real>
__global__ void __launch_bounds__(256) simplex_objective_kernel(SP<real> p)
{
    __shared__ double smem[32];
    Ctl *ctl = p.ctl;
    const CtlView cv = load_ctl(ctl);
    if (cv.hold) return;
    const int it = cv.it, K = p.K;
    const real *__restrict__ P = (it & 1) ? p.P1 : p.P0;
    double r[1] = {0.0};
    const size_t stride = (size_t) gridDim.x * blockDim.x, t0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const bool slab = p.slab.on != 0;                        // slab-sharded: owned vertices, edges whose first end is owned
    if (slab) {                                              // ... whose other end may be a ghost: the neighbours' stores have landed
        if (threadIdx.x == 0) slab_wait_neighbours(p.slab, it);
        __syncthreads();
    }
    const size_t i0 = slab ? (size_t) p.vbeg * K : 0, i1 = slab ? (size_t) p.vend * K : (size_t) p.V * K;
    for (size_t i = i0 + t0; i < i1; i += stride) {
        const real Pv = P[i], q = p.Q[i];
        const double lf = p.La_f ? (double) p.La_f[i / K] : 1.0;
        if (p.loss == LOSS_LIN) r[0] -= (double) Pv * (double) q;
        else if (p.loss == LOSS_QUAD) { const double c = (double) Pv - (double) q; r[0] += 0.5 * lf * c * c; }
        else { const real c = p.aK + p.a1 * q; r[0] += lf * (double) c * log((double) (c / (p.aK + p.a1 * Pv))); }
    }
    for (size_t i = t0; i < (size_t) p.E * K; i += stride) {
        const size_t e = i / K, k = i % K;
        const int u = p.Eu[e];
        if (slab && (u < p.vbeg || u >= p.vend)) continue;
        r[0] += (double) p.La_d1[e] * (double) sabs(P[(size_t) u * K + k] - P[(size_t) p.Ev[e] * K + k]);
    }
    if (grid_sum_s<1>(ctl, p.partial, r, smem) && threadIdx.x == 0) {
        if (slab) *p.obj_part = r[0];                            // travels with the trip's stopping sums (simplex_slab_decide)
        else p.dObj[it] = (real) r[0];
    }
}

template <typename real>
__global__ void simplex_obj_final_kernel(SP<real> p) { p.dObj[p.ctl->it] = (real) *p.obj_part; }

template <typename real>
__global__ void simplex_extract_kernel(SP<real> p, real *__restrict__ out)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) p.Vown * p.Kv) return;
    const size_t v = i / p.Kv, k = i - v * p.Kv;                               // out has the caller's row stride Kv
    out[i] = ((p.ctl->it & 1) ? p.P1 : p.P0)[(v + (size_t) (p.dia_nd ? p.vbeg : 0)) * p.K + k];
}

// K-vectors of the caller (row stride Kv) into rows of stride K >= Kv, the extra labels zero
template <typename real>
__global__ void simplex_pack_kernel(const real *__restrict__ in, real *__restrict__ out, size_t V, int Kv, int K)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V * (size_t) K) return;
    const size_t v = i / K, k = i - v * K;
    out[i] = k < (size_t) Kv ? in[v * Kv + k] : (real) 0;
}
template <typename real>
__global__ void simplex_fill_kernel(real *__restrict__ out, size_t n, real x)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename real>
struct SimplexSolver : SolverBase {
    const SimplexProblem<real> q;
    bool keep_p0 = false;
    SellGraph g;
    DiaGraph dg;
    bool dia = false;               // constant-offset layout (dia.cuh) instead of the half-edge table
    long long nhalf = 0;            // half-edge copies held per label (slots of the table / 2.ND.V)
    SP<real> p{};
    DevBuf<Ctl> ctl;
    DevBuf<real> Wrow;
    DevIn<real> Q, La_f, La_d1, P0;
    DevIn<int> Eu, Ev;
    DevBuf<real> w, wk, thk, Z0, Z1, Pa, Pb, Fa, Fb, T, GaQ, Sinv, Sinvk, Gn, GaU, dObj, dDif, Pout, Qal;
    DevBuf<int> lab, order;
    DevBuf<double> partial, sums, obj_part;
    Halo<real> halo;
    SlabLink link;                  // slab-sharded solve (slab.cuh)
    bool slab = false;
    int slab_lowB = 0, slab_highB = 0, slab_nbchunks = 0;
    int hctl_it_parity = 0;         // parity of the trip a reconditioning interrupts
    bool sharded = false;
    Ctl *hctl = nullptr;
    cudaEvent_t ev[2]{}, ev_a{}, ev_b{}, ev_c{};
    cudaGraphExec_t gexec = nullptr;
    ProgressPrinter progress;
    int graph_trips = 0, grid = 0, ogrid = 0, LPT = 1, Vglobal = 0;

    SimplexSolver(const SimplexProblem<real> &q_, bool keep) : q(q_), keep_p0(keep) {}
    ~SimplexSolver() override
    {
        if (gexec) cudaGraphExecDestroy(gexec);
        pinned_put(hctl);
        for (auto e : {ev[0], ev[1], ev_a, ev_b, ev_c}) event_put(e);
    }

    void setup()
    {
        const int Kv = q.K, V = q.V, E = q.E;
        int K = Kv;                                           // row stride of every K-vector on the device
        PFDR_REQUIRE(K >= 1 && V > 0 && E >= 0 && q.itMax >= 0, "K, V must be positive, E and itMax non-negative");
        PFDR_REQUIRE(q.al >= (real) 0 && q.al <= (real) 1, "al must lie in [0, 1] (the reference is undefined beyond, SURVEY 5.1-10)");
        sharded = q.shard != nullptr;
        slab = q.slab != nullptr;
        if (slab) {
            PFDR_REQUIRE(!sharded, "a solve is sharded either by slabs or by a shard plan");
            PFDR_REQUIRE(q.slab->world >= 1 && q.slab->lo <= q.slab->v0 && q.slab->v0 < q.slab->v1 && q.slab->v1 <= q.slab->hi &&
                         q.slab->hi - q.slab->lo == V, "bad slab: need lo <= v0 < v1 <= hi and Vloc = hi - lo");
        }
        if (sharded) {
            PFDR_REQUIRE(q.Obj == nullptr && q.difRcd == (real) 0, "sharded solves do not record Obj nor recondition in this build");
            PFDR_REQUIRE(q.shard->Vown > 0 && q.shard->Vown <= V, "bad shard: Vown");
        }
        // Labels per lane: 4 in float32 (one 16-byte access per piece of a K-vector), 2 in float64 when that fits 32 lanes
        // (half the registers of a lane, twice the resident warps), else 4.  When K is not a multiple of it, rows are
        // padded on the device: the extra labels stay exactly 0 and never enter a projection, an arg-max or a sum
        // (masks on Kv).  Measured on a 64^3 grid in float32: K = 5 / 13 / 21 run 2.0 / 2.5 / 2.7 x faster than with one
        // label per lane (profiles/r01_notes.md).  PFDR_B200_LPT=1 forces one label per lane, =4 four (float64).
        const char *fl = getenv("PFDR_B200_LPT");
        LPT = 4;
        // (two labels per lane win on mid-size problems - 64^3 x 16 labels: 16.4 vs 16.9 ms per 100 iterations - and lose on
        //  the full-size tiled sweep - 252^3: 10.4 vs 10.0 ms per iteration: chosen by size, PFDR_B200_LPT=2 / 4 force)
        if (sizeof(real) == 8 && ((Kv + 1) & ~1) <= 64 && !(fl && fl[0] == '4') &&
            ((size_t) V * (size_t) Kv <= ((size_t) 1 << 23) || (fl && fl[0] == '2'))) LPT = 2;
        if (Kv <= 2 || (fl && fl[0] == '1' && Kv <= 32)) LPT = 1;
        // more than 128 labels: 8 or 16 labels per lane (a vertex still fits one warp; registers spill - a correct, slow path
        // for a size the reference accepts)
        if (Kv > 128) LPT = Kv <= 256 ? 8 : 16;
        cudaStream_t st = ctx().stream;
        for (auto *e : {&ev[0], &ev[1], &ev_a, &ev_b, &ev_c}) *e = event_get();
        PFDR_CUDA(cudaEventRecord(ev_a, st));
        Eu.bind(q.Eu, E); Ev.bind(q.Ev, E);
        {   // constant-offset graph (grids): index-free layout, PFDR_B200_DIA=0 keeps the half-edge table (tests compare the two)
            const char *de = getenv("PFDR_B200_DIA");
            constexpr int DL = SimplexDiaLanes<real>::LPT;
            const bool lanes_ok = Kv >= 3 && Kv <= 128 && !(fl && fl[0] != '0' + DL);
            if (!sharded && lanes_ok && !(de && de[0] == '0') && (size_t) V * (size_t) (((Kv + DL - 1) / DL) * DL) * 6 < ((size_t) 1 << 40))
                dia = build_dia(dg, Eu, Ev, V, E);
            if (dia) LPT = DL;
            PFDR_REQUIRE(dia || !slab, "slab sharding needs a constant-offset (grid) graph with its edges in vertex-major order, "
                                       "K >= 3 labels and a plain PFDR_B200_DIA / PFDR_B200_LPT environment");
        }
        K = ((Kv + LPT - 1) / LPT) * LPT;
        const bool padded = K != Kv;
        PFDR_REQUIRE((K + LPT - 1) / LPT <= 32, "this build supports K <= 512 labels");
        static_assert(2 * sizeof(Ctl) <= PINNED_BLOCK, "control block mirror");
        hctl = (Ctl *) pinned_get();
        int TPV = 1, logTPV = 0;
        while (TPV * LPT < K) { TPV <<= 1; logTPV++; }
        const int C = 32 / TPV;
        int logC = 0; while ((1 << logC) < C) logC++;
        const size_t VK = (size_t) V * K;
        Q.bind(q.Q, (size_t) V * Kv); La_d1.bind(q.La_d1, E);
        La_f.bind(q.al > (real) 0 ? q.La_f : nullptr, V); P0.bind(q.P, (size_t) V * Kv);
        if (padded) {
            DevBuf<real> Qp(VK), Pp(VK);
            PFDR_LAUNCH(simplex_pack_kernel<real>, cdiv(VK, 256), 256, 0, Q.p, Qp.p, (size_t) V, Kv, K);
            PFDR_LAUNCH(simplex_pack_kernel<real>, cdiv(VK, 256), 256, 0, P0.p, Pp.p, (size_t) V, Kv, K);
            Q.own = std::move(Qp); Q.p = Q.own.p;
            P0.own = std::move(Pp); P0.p = P0.own.p;
        }
        if (LPT > 1 && ((uintptr_t) Q.p) % 32 != 0) {           // vector loads of Q (quadratic loss) need alignment
            Qal.alloc(VK);
            PFDR_CUDA(cudaMemcpyAsync(Qal.p, Q.p, VK * sizeof(real), cudaMemcpyDeviceToDevice, st));
            Q.own = std::move(Qal); Q.p = Q.own.p;
        }
        if (dia) {
            nhalf = 2LL * dg.ND * V;
            Wrow.alloc((size_t) nhalf); Wrow.zero();
            ctx().stats.slots = 2LL * E; ctx().stats.layout = 1;
        } else {
            build_sell(g, Eu, Ev, V, E, C);
            nhalf = g.nslots;
            w.alloc((size_t) std::max<long long>(nhalf, 1));
        }
        const size_t ns = (size_t) std::max<long long>(nhalf, 1);
        if (slab) {          // what the neighbouring ranks write / copy into lives in the peer-visible arena
            link.reserve(SLAB_Z0, ns * K * sizeof(real)); link.reserve(SLAB_Z1, ns * K * sizeof(real));
            link.reserve(SLAB_F0, VK * sizeof(real)); link.reserve(SLAB_F1, VK * sizeof(real)); link.reserve(SLAB_T, VK * sizeof(real));
            link.reserve(SLAB_X0, VK * sizeof(real)); link.reserve(SLAB_X1, VK * sizeof(real));
            link.commit(q.slab->world, q.slab->rank, q.slab->lo, q.slab->hi, q.slab->v0, q.slab->v1, K);
            Z0.borrow(link.ptr<real>(SLAB_Z0), ns * K); Z1.borrow(link.ptr<real>(SLAB_Z1), ns * K);
            Fa.borrow(link.ptr<real>(SLAB_F0), VK); Fb.borrow(link.ptr<real>(SLAB_F1), VK); T.borrow(link.ptr<real>(SLAB_T), VK);
            Pa.borrow(link.ptr<real>(SLAB_X0), VK); Pb.borrow(link.ptr<real>(SLAB_X1), VK);
            obj_part.alloc(1); obj_part.zero();
        } else {
            Z0.alloc(ns * K); Z1.alloc(ns * K); Fa.alloc(VK); Fb.alloc(VK); T.alloc(VK); Pa.alloc(VK); Pb.alloc(VK);
        }
        GaQ.alloc(VK); Sinv.alloc(V);
        if (q.difRcd > (real) 0) GaU.alloc(VK);
        Z0.zero(); Z1.zero(); w.zero();
        if (padded) {           // what the extra labels read: T = 1 (a denominator), everything else 0
            Pa.zero(); Pb.zero(); Fa.zero(); Fb.zero(); GaQ.zero();
            PFDR_LAUNCH(simplex_fill_kernel<real>, cdiv(VK, 256), 256, 0, T.p, VK, (real) 1);
        }
        const bool by_label = q.difTol >= (real) 1;
        if (by_label) lab.alloc(V);
        if (q.Obj) dObj.alloc((size_t) q.itMax + 1);
        if (q.Dif || q.verbose > 0) dDif.alloc((size_t) std::max(q.itMax, 1));
        ctl.alloc(1); sums.alloc((size_t) std::max(1, comm().world)); sums.zero();
        const long long nslices = ((long long) V + C - 1) / C;
        grid = (int) cdiv((size_t) nslices * 32, 256);
        const int Vown = sharded ? q.shard->Vown : (slab ? q.slab->v1 - q.slab->v0 : V);
        ogrid = (int) cdiv((size_t) (((long long) Vown + C - 1) / C) * 32, 256);
        partial.alloc((size_t) std::max(grid, num_sms() * 8) * 2);

        p.Vown = Vown; p.sharded = sharded ? 1 : 0; p.sums = sums.p; p.Vglobal = sharded ? q.shard->Vglobal : (slab ? q.slab->Vglobal : V);
        p.world = sharded ? comm().world : 1; p.rank = sharded ? comm().rank : 0;
        {   // banded graph whose +-band reuse window would not fit the L2: sweep (x,y)-tiles along z instead
            const int width = dia ? 2 * dg.ND : std::max(g.max_width, 1), band = dia ? (dg.ND > 1 ? dg.band() : 0) : g.band;
            const double bytes_per_vertex = (double) sizeof(real) * K * (2.0 * width + 6.0);
            const char *e = getenv("PFDR_B200_TILE");
            const long long tile_env = e ? atoll(e) : -1;
            p.order = nullptr;
            const int nchunks = (int) (((long long) Vown + C - 1) / C);
            // slab-sharded: the chunks within the band of a neighbouring rank are worked on first (slab.cuh)
            slab_lowB = (slab && q.slab->rank > 0) ? std::min(dg.band(), Vown) : 0;
            slab_highB = (slab && q.slab->rank + 1 < q.slab->world) ? std::min(dg.band(), Vown - slab_lowB) : 0;
            {
                const long long lowc = ((long long) slab_lowB + C - 1) / C, firsthigh = slab_highB ? ((long long) Vown - slab_highB) / C : nchunks;
                slab_nbchunks = (int) std::min<long long>(nchunks, lowc + (nchunks - firsthigh) - std::max<long long>(0, lowc - firsthigh));
            }
            const bool tiled = tile_env != 0 && band > 0 && ((double) band * bytes_per_vertex > 24e6 || tile_env > 0);
            if (tiled || slab_lowB + slab_highB > 0) {
                long long T = tile_env > 0 ? tile_env : (long long) (9.3e6 / bytes_per_vertex);   // C4 252^3 f32: 32 rows of a plane (profiles/r02_notes.md)
                T = std::max<long long>(C, std::min<long long>(T, std::max(band, C)));
                T = tiled ? (T / C) * C : 0;
                const int nx = (dia && dg.ND == 3 && dg.off[0] == 1 && dg.off[2] % dg.off[1] == 0) ? dg.off[1] : 0;
                build_chunk_order(nchunks, C, band, (int) T, order, nx, slab ? q.slab->v0 : 0, slab_lowB, slab_highB, Vown);
                p.order = order.p;
            }
        }
        p.ctl = ctl.p; p.K = K; p.Kv = Kv; p.V = V; p.E = E; p.TPV = TPV; p.C = C; p.logC = logC; p.logTPV = logTPV; p.LPT = LPT;
        p.loss = (q.al == (real) 0) ? LOSS_LIN : (q.al == (real) 1 ? LOSS_QUAD : LOSS_KLA);
        p.by_label = by_label;
        p.aK = q.al / (real) Kv; p.a1 = (real) 1 - q.al; p.aK1 = p.aK / p.a1;                     // ref :386-391
        p.rho = q.rho; p.condMin = q.condMin;
        p.slice_off = g.slice_off.p; p.nbr = g.nbr.p; p.twin = g.twin.p; p.hkey = g.hkey.p;
        p.dia_nd = dia ? dg.ND : 0;
        for (int i = 0; i < 3; i++) p.dia_off[i] = dia ? dg.off[i] : 0;
        p.vbeg = slab ? q.slab->v0 - q.slab->lo : 0; p.vend = p.vbeg + Vown; p.dmask = dg.mask.p;
        p.slab = SlabDev{}; p.plo = SlabPeer<real>{}; p.phi = SlabPeer<real>{};
        if (slab) { p.slab = link.dev(); p.slab.band = dg.band(); p.slab.lowB = slab_lowB; p.slab.highB = slab_highB; p.slab.nbchunks = slab_nbchunks; } p.eoff = dg.eoff.p; p.Wrow = Wrow.p;
        p.w = w.p; p.wk = nullptr; p.thk = nullptr; p.Z0 = Z0.p; p.Z1 = Z1.p;
        p.P0 = Pa.p; p.P1 = Pb.p; p.F0 = Fa.p; p.F1 = Fb.p; p.T = T.p; p.GaQ = GaQ.p; p.Sinv = Sinv.p;
        p.Sinvk = nullptr; p.Gn = nullptr; p.GaU = GaU.p; p.lab = lab.p;
        p.Q = Q.p; p.La_f = La_f.p; p.La_d1 = La_d1.p; p.Eu = Eu.p; p.Ev = Ev.p;
        p.partial = partial.p; p.dObj = dObj.p; p.dDif = dDif.p; p.obj_part = obj_part.p;
        p.stream_z = 0;

        if (sharded) { halo.setup(*q.shard, g.hkey.p, g.nslots, 2 * (size_t) E, K); halo.build(HALO_VA0 | HALO_VA1); halo.build(HALO_VA0 | HALO_Z); }
        if (!(slab && link.world > 1)) reset();            // a slab-sharded solve is initialised by slab_connect (collective)
        if (!keep_p0) P0.own.release();
        PFDR_CUDA(cudaEventRecord(ev_b, st));
    }

    void launch_precond(const real *Pinit, int recond)
    {
        if (dia) { simplex_dia_launch_precond<real>(p, Pinit, recond, grid); return; }
        if (LPT == 16) PFDR_LAUNCH((simplex_precond_kernel<real, 16>), grid, 256, 0, p, Pinit, recond);
        else if (LPT == 8) PFDR_LAUNCH((simplex_precond_kernel<real, 8>), grid, 256, 0, p, Pinit, recond);
        else if (LPT == 4) PFDR_LAUNCH((simplex_precond_kernel<real, 4>), grid, 256, 0, p, Pinit, recond);
        else if (LPT == 2) { if constexpr (sizeof(real) == 8) PFDR_LAUNCH((simplex_precond_kernel<real, 2>), grid, 256, 0, p, Pinit, recond); }
        else PFDR_LAUNCH((simplex_precond_kernel<real, 1>), grid, 256, 0, p, Pinit, recond);
    }

    void reset() override
    {
        cudaStream_t st = ctx().stream;
        PFDR_REQUIRE(P0.p != nullptr, "reset: the initial iterate was not kept on the device");
        if (gexec && p.wk) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        p.wk = nullptr; p.thk = nullptr; p.Sinvk = nullptr; p.Gn = nullptr;
        Ctl h{};                                                                                  // ref :445-446
        h.it = 0; h.state = (q.itMax == 0) ? ST_FINAL : ST_RUN; h.itMax = q.itMax;
        h.track = (q.difTol > (real) 0 || q.difRcd > (real) 0 || q.Dif != nullptr);
        h.tol = (double) q.difTol; h.rcd = (double) q.difRcd;
        h.dif = (double) (q.difTol > q.difRcd ? q.difTol : q.difRcd);
        hctl[0] = h;
        PFDR_CUDA(cudaMemcpyAsync(ctl.p, &hctl[0], sizeof(Ctl), cudaMemcpyHostToDevice, st));
        if (slab && link.world > 1) {                          // nobody still iterates on the buffers about to be rewritten
            PFDR_REQUIRE(link.connected, "slab-sharded solver: reset before slab_connect");
            link.barrier();
            link.run_id++;
            p.slab.run_id = link.run_id;
            if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        }
        launch_precond(P0.p, 0);
        if (sharded) halo.exchange(HALO_VA0 | HALO_VA1, PingPong<real>(T.p), PingPong<real>(Fa.p), PingPong<real>(), nullptr, PARITY_ZERO);
        if (slab && link.connected && link.world > 1) slab_refresh();
        PFDR_CUDA(cudaStreamSynchronize(st));
        if (slab && link.world > 1) slab_check();
    }

    // ghost zones of the neighbours <- this rank's boundary: metric T, forward point and cut-edge copies of the CURRENT
    // (parity 0) buffers; afterwards every rank has everything the first trip reads
    void slab_refresh(int par = 0)
    {
        const size_t VK = (size_t) q.V * p.K;
        const int K = p.K, ND = dg.ND;
        for (int side = 0; side < 2; side++) {
            const int nb = side ? link.nb_hi : link.nb_lo;
            if (nb < 0) continue;
            const SlabBlob &b = link.blobs[nb];
            const long long depth = side ? b.v0 - b.lo : b.hi - b.v1;          // its ghost zone on our side
            const long long first = side ? p.vend - depth : p.vbeg;
            const long long dshift = (link.lo - b.lo) * K;
            const size_t VKn = (size_t) (b.hi - b.lo) * K;
            link.push<real>(SLAB_T, nb, (size_t) first * K, (size_t) depth * K, dshift);
            link.push<real>(par ? SLAB_F1 : SLAB_F0, nb, (size_t) first * K, (size_t) depth * K, dshift);
            link.push<real>(par ? SLAB_X1 : SLAB_X0, nb, (size_t) first * K, (size_t) depth * K, dshift);
            for (int i = 0; i < ND; i++) {                                       // copies on the cut edges of offset d_i
                const long long d = std::min<long long>(dg.off[i], depth), f = side ? p.vend - d : p.vbeg;
                const int plane = side ? i : ND + i;                             // our out_i copies go up, our in_i copies down
                link.push<real>(par ? SLAB_Z1 : SLAB_Z0, nb, (size_t) plane * VK + (size_t) f * K, (size_t) d * K,
                                (long long) plane * ((long long) VKn - (long long) VK) + dshift);
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
        p.slab = link.dev(); p.slab.band = dg.band(); p.slab.lowB = slab_lowB; p.slab.highB = slab_highB; p.slab.nbchunks = slab_nbchunks;
        for (int side = 0; side < 2; side++) {
            const int nb = side ? link.nb_hi : link.nb_lo;
            if (nb < 0) continue;
            SlabPeer<real> &pe = side ? p.phi : p.plo;
            pe.Z[0] = link.ptr<real>(SLAB_Z0, nb); pe.Z[1] = link.ptr<real>(SLAB_Z1, nb);
            pe.F[0] = link.ptr<real>(SLAB_F0, nb); pe.F[1] = link.ptr<real>(SLAB_F1, nb);
            pe.X[0] = link.ptr<real>(SLAB_X0, nb); pe.X[1] = link.ptr<real>(SLAB_X1, nb);
            pe.plane = (size_t) (link.blobs[nb].hi - link.blobs[nb].lo) * p.K;
            pe.shift = (link.lo - link.blobs[nb].lo) * p.K;
        }
        if (link.world > 1) reset();                          // collective: first preconditioning, ghost zones, barrier
    }

    template <int L, bool PERK, int LOSS>
    void launch_iter_u()
    {
        constexpr int MBD = (sizeof(real) == 4 || L == 2) ? 4 : 2;
        const bool small = (size_t) std::max<long long>(g.nslots, q.V) * (size_t) p.K < 0xFFFFFFFFull;
        const char *ie = getenv("PFDR_B200_IDX32");              // read per call: the tests switch it (0: 64-bit element offsets)
        const bool i32 = !(ie && ie[0] == '0');
        // a slice wider than SIMPLEX_HUB_ROWS (a hub vertex): variant whose warps share the hub's rows
        if (g.max_width > SIMPLEX_HUB_ROWS) PFDR_LAUNCH((simplex_iter_kernel<real, L, PERK, LOSS, 1, false, true>), ogrid, 256, 0, p);
        else if (L > 1 && !PERK && small && i32) {
            if constexpr (L > 1) PFDR_LAUNCH((simplex_iter_kernel<real, L, false, LOSS, MBD, true, false>), ogrid, 256, 0, p);
        }
        else PFDR_LAUNCH((simplex_iter_kernel<real, L, PERK, LOSS, MBD, false, false>), ogrid, 256, 0, p);
    }
    // K > 128: one generic variant per (labels per lane, loss): 64-bit offsets, one CTA per SM's worth of registers
    template <int L, bool PERK>
    void launch_iter_wide()
    {
        if (p.loss == LOSS_LIN) PFDR_LAUNCH((simplex_iter_kernel<real, L, PERK, LOSS_LIN, 1, false, false>), ogrid, 256, 0, p);
        else if (p.loss == LOSS_QUAD) PFDR_LAUNCH((simplex_iter_kernel<real, L, PERK, LOSS_QUAD, 1, false, false>), ogrid, 256, 0, p);
        else PFDR_LAUNCH((simplex_iter_kernel<real, L, PERK, LOSS_KLA, 1, false, false>), ogrid, 256, 0, p);
    }
    template <int L, bool PERK>
    void launch_iter_loss()
    {
        if (p.loss == LOSS_LIN) launch_iter_u<L, PERK, LOSS_LIN>();
        else if (p.loss == LOSS_QUAD) launch_iter_u<L, PERK, LOSS_QUAD>();
        else launch_iter_u<L, PERK, LOSS_KLA>();
    }

    void enqueue_trip()
    {
        if (q.Obj) {
            const size_t work = (size_t) std::max(q.E, q.V) * p.K;
            PFDR_LAUNCH(simplex_objective_kernel<real>, std::min<unsigned>(cdiv(work, 256), num_sms() * 8), 256, 0, p);
        }
        if (dia) simplex_dia_launch_iter<real>(p, ogrid);
        else if (LPT == 16) { if (p.wk) launch_iter_wide<16, true>(); else launch_iter_wide<16, false>(); }
        else if (LPT == 8) { if (p.wk) launch_iter_wide<8, true>(); else launch_iter_wide<8, false>(); }
        else if (LPT == 4) { if (p.wk) launch_iter_loss<4, true>(); else launch_iter_loss<4, false>(); }
        else if (LPT == 2) { if constexpr (sizeof(real) == 8) { if (p.wk) launch_iter_loss<2, true>(); else launch_iter_loss<2, false>(); } }
        else { if (p.wk) launch_iter_loss<1, true>(); else launch_iter_loss<1, false>(); }
        if (sharded) {     // evolution sum over all ranks, then the new boundary forward points and cut-edge copies
            halo.exchange(HALO_VA0 | HALO_Z, PingPong<real>(Fa.p, Fb.p), PingPong<real>(), PingPong<real>(Z0.p, Z1.p), ctl.p,
                          PARITY_NEXT_IF_RAN, sums.p, 1);
            PFDR_LAUNCH(simplex_finalize_kernel<real>, 1, 1, 0, p);
        }
    }
    int launches_per_trip() const { return (q.Obj ? 1 : 0) + 1 + (sharded ? 4 : 0); }

    void enqueue_batch(int batch)
    {
        cudaStream_t st = ctx().stream;
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
    }

    void recondition()
    {
        if (!wk.p) {
            const size_t n = (size_t) std::max<long long>(nhalf, 1) * p.K, VK = (size_t) q.V * p.K;
            wk.alloc(n); thk.alloc(n); wk.zero(); thk.zero();
            Sinvk.alloc(VK); Gn.alloc(VK);
            if (p.K != p.Kv) { Sinvk.zero(); Gn.zero(); }                       // padded labels: defined values
        }
        p.wk = wk.p; p.thk = thk.p; p.Sinvk = Sinvk.p; p.Gn = Gn.p;
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        launch_precond(nullptr, 1);
        if (slab && link.world > 1) slab_refresh(hctl_it_parity);   // new metric, forward point and copies of the boundary -> neighbours' ghost zones
        PFDR_LAUNCH(simplex_resume_kernel<real>, 1, 1, 0, p);
    }

    int inspect(int b)
    {
        PFDR_CUDA(cudaEventSynchronize(ev[b]));
        const Ctl &h = hctl[b];
        if (!h.hold) return 0;
        return h.state == ST_FINAL ? 1 : (h.state == ST_RECOND ? 2 : 0);
    }

    void run() override
    {
        cudaStream_t st = ctx().stream;
        const size_t work = (size_t) nhalf * p.K;
        int batch = work > (1u << 24) ? 4 : (work > (1u << 20) ? 16 : 64);
        batch = std::max(1, std::min(batch, q.itMax + 1));
        ctx().stats.layout = dia ? 1 : 0; ctx().stats.gpus = slab ? link.world : 1;
        ctx().in_loop = true;
        PFDR_CUDA(cudaEventRecord(ev_b, st));
        progress.start(q.verbose);
        long long enqueued = 0, budget = (long long) q.itMax + 1 + 2LL * batch;
        int fresh_from = 0;
        for (int k = 0;; k++) {
            enqueue_batch(batch);
            enqueued += batch;
            PFDR_CUDA(cudaMemcpyAsync(&hctl[k & 1], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
            PFDR_CUDA(cudaEventRecord(ev[k & 1], st));
            int verdict = 0;
            if (k >= 1 && k - 1 >= fresh_from) verdict = inspect((k - 1) & 1);
            if (verdict == 0 && enqueued >= budget) verdict = (k >= fresh_from) ? inspect(k & 1) : 0;
            if (progress.every && k >= 1 && k - 1 >= fresh_from)
                progress.upto(hctl[(k - 1) & 1].it, q.itMax, dDif.p, hctl[(k - 1) & 1].track != 0, hctl[(k - 1) & 1].tol, hctl[(k - 1) & 1].rcd);
            if (verdict == 1) break;
            if (verdict == 2) {
                PFDR_CUDA(cudaStreamSynchronize(st));
                hctl_it_parity = hctl[k & 1].it & 1;            // (the device froze at the trip that asked for it)
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
                PFDR_LAUNCH(simplex_obj_final_kernel<real>, 1, 1, 0, p);
                PFDR_CUDA(cudaStreamSynchronize(st));
            }
            slab_check();
        }
    }

    void fetch(void *Po, int *ito, void *Objo, void *Difo) override
    {
        cudaStream_t st = ctx().stream;
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        const Ctl h = hctl[0];
        const size_t VK = (size_t) p.Vown * p.Kv;
        if (Po) {
            if (is_device_ptr(Po)) PFDR_LAUNCH(simplex_extract_kernel<real>, cdiv(VK, 256), 256, 0, p, (real *) Po);
            else {
                Pout.alloc(VK);
                PFDR_LAUNCH(simplex_extract_kernel<real>, cdiv(VK, 256), 256, 0, p, Pout.p);
                copy_out((real *) Po, Pout.p, VK);
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
        if (q.verbose) {                                                       // final information, as the reference's
            printf("iteration %d (max. %d)\n", h.it, q.itMax);
            if (h.track) printf("iterate evolution %g (recond. %g; tol. %g)\n", h.dif, h.rcd, h.tol);
            printf("PFDR on B200: loop %.3f ms, %lld kernel launches\n", ctx().stats.loop_ms, ctx().stats.loop_launches);
            fflush(stdout);
        }
    }
};

template <typename real>
SolverBase *make_simplex(const SimplexProblem<real> &q)
{
    SimplexSolver<real> *s = new SimplexSolver<real>(q, true);
    try { s->setup(); } catch (...) { delete s; throw; }
    return s;
}
template SolverBase *make_simplex<float>(const SimplexProblem<float> &);
template SolverBase *make_simplex<double>(const SimplexProblem<double> &);

template <typename real>
void solve_simplex(const SimplexProblem<real> &q)
{
    const bool trace = getenv("PFDR_B200_TRACE") != nullptr;      // one line per call on stderr
    const auto t0 = std::chrono::steady_clock::now();
    SimplexSolver<real> s(q, false);
    s.setup();
    const auto t1 = std::chrono::steady_clock::now();
    s.run();
    const auto t2 = std::chrono::steady_clock::now();
    s.fetch(q.P, q.it, q.Obj, q.Dif);
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        fprintf(stderr, "[pfdr_b200] simplex V=%d E=%d K=%d f%d: %d it, setup %.2f ms, run %.2f ms (device loop %.2f ms), fetch %.2f ms\n",
                q.V, q.E, q.K, (int) sizeof(real) * 8, (int) ctx().stats.iterations, ms(t0, t1), ms(t1, t2), ctx().stats.loop_ms, ms(t2, t3));
    }
}
template void solve_simplex<float>(const SimplexProblem<float> &);
template void solve_simplex<double>(const SimplexProblem<double> &);

// ---------------------------------------------------------------------------------------------
// stand-alone proj_simplex_metric (ref src/proj_simplex_metric.cpp:18-83).  Here the lanes of a column run the
// reference's *sequential* recursion redundantly through shuffles, so the result tracks the reference to
// the last bits (the fused iteration above uses the cheaper parallel active-set form).
// ---------------------------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ real simplex_project_seq(real x, real m, real target, int k, int K, int TPV)
{
    const unsigned full = 0xffffffffu;
    const real xs = x / m;
    const real x0 = __shfl_sync(full, x, 0, TPV), m0 = __shfl_sync(full, m, 0, TPV);
    real la = (x0 - target) / m0, s = m0;
    unsigned act = 1u;
    for (int d = 1; d < K; d++) {                                  // first pass  :49-58
        const real xd = __shfl_sync(full, xs, d, TPV), md = __shfl_sync(full, m, d, TPV);
        if (xd > la) { act |= 1u << d; s += md; la += md * (xd - la) / s; }
    }
    for (;;) {                                                     // pruning passes :60-73
        bool changed = false;
        for (int d = 0; d < K; d++) {
            const real xd = __shfl_sync(full, xs, d, TPV), md = __shfl_sync(full, m, d, TPV);
            if (((act >> d) & 1u) && xd < la) { act &= ~(1u << d); s -= md; la += md * (la - xd) / s; changed = true; }
        }
        if (!__any_sync(full, changed)) break;
    }
    return ((act >> k) & 1u) ? (xs - la) * m : (real) 0;           // :75-81
}

template <typename real>
__global__ void __launch_bounds__(256) proj_small_kernel(real *__restrict__ X, const real *__restrict__ M, int D, int N,
                                                         int nm, const real *__restrict__ A, int na, int TPV, int logTPV)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int C = 32 >> logTPV;
    if (warp * C >= N) return;
    const long long n = warp * C + (lane >> logTPV);
    const int d = lane & (TPV - 1);
    const bool on = n < N && d < D;
    const size_t xi = (size_t) (on ? n : 0) * D + (on ? d : 0);
    const size_t mi = (size_t) (on ? (n < nm ? n : nm - 1) : 0) * D + (on ? d : 0);
    const real x = on ? X[xi] : (real) 0, m = on ? M[mi] : (real) 1;
    const real a = A[on ? (n < na ? n : na - 1) : 0];
    const real r = simplex_project_seq<real>(x, m, a, d, D, TPV);
    if (on) X[xi] = r;
}

// any D: one thread per column, the reference's loops verbatim on global memory, flags in a bit set
// WORDS > 0: the flags of up to 32 * WORDS coordinates in local memory; WORDS == 0: in the scratch array `flags`
// (words_per_col words per column) - no limit on D
template <typename real, int WORDS>
__global__ void __launch_bounds__(128) proj_serial_kernel(real *__restrict__ X, const real *__restrict__ M, int D, int N,
                                                          int nm, const real *__restrict__ A, int na, unsigned *__restrict__ flags,
                                                          int words_per_col)
{
    const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    real *x = X + (size_t) n * D;
    const real *m = M + (size_t) (n < nm ? n : nm - 1) * D;
    const real a = A[n < na ? n : na - 1];
    unsigned local[WORDS > 0 ? WORDS : 1];
    unsigned *act = WORDS > 0 ? local : flags + (size_t) n * words_per_col;
    const int words = WORDS > 0 ? WORDS : words_per_col;
    for (int i = 0; i < words; i++) act[i] = 0;
    real la = (x[0] - a) / m[0], s = m[0];
    x[0] = x[0] / m[0]; act[0] = 1u;
    for (int d = 1; d < D; d++) {
        const real xd = x[d] / m[d];
        x[d] = xd;
        if (xd > la) { act[d >> 5] |= 1u << (d & 31); s += m[d]; la += m[d] * (xd - la) / s; }
    }
    for (bool again = true; again;) {
        again = false;
        for (int d = 0; d < D; d++)
            if (((act[d >> 5] >> (d & 31)) & 1u) && x[d] < la) {
                act[d >> 5] &= ~(1u << (d & 31)); s -= m[d]; la += m[d] * (la - x[d]) / s; again = true;
            }
    }
    for (int d = 0; d < D; d++) x[d] = ((act[d >> 5] >> (d & 31)) & 1u) ? (x[d] - la) * m[d] : (real) 0;
}

template <typename real>
void proj_simplex_metric(real *X, const real *M, int D, int N, int nm, const real *A, int na)
{
    PFDR_REQUIRE(D >= 1 && N >= 0 && nm >= 1 && na >= 1, "proj_simplex_metric: D, nm, na must be positive");
    if (N == 0) return;
    const size_t n = (size_t) D * N;
    const bool inplace = is_device_ptr(X);
    if (inplace) order_after_caller();
    DevBuf<real> xbuf;
    real *dX = X;
    if (!inplace) {
        xbuf.alloc(n);
        PFDR_CUDA(cudaMemcpyAsync(xbuf.p, X, n * sizeof(real), cudaMemcpyHostToDevice, ctx().stream));
        ctx().stats.h2d_bytes += (long long) (n * sizeof(real));
        dX = xbuf.p;
    }
    DevIn<real> dM, dA;
    dM.bind(M, (size_t) D * nm); dA.bind(A, na);
    if (D <= 32) {
        int TPV = 1, logTPV = 0;
        while (TPV < D) { TPV <<= 1; logTPV++; }
        const size_t warps = ((size_t) N + (32 / TPV) - 1) / (32 / TPV);
        PFDR_LAUNCH(proj_small_kernel<real>, cdiv(warps * 32, 256), 256, 0, dX, dM.p, D, N, nm, dA.p, na, TPV, logTPV);
    } else if (D <= 4096) {
        PFDR_LAUNCH((proj_serial_kernel<real, 128>), cdiv(N, 128), 128, 0, dX, dM.p, D, N, nm, dA.p, na, (unsigned *) nullptr, 0);
    } else {
        const int words = (D + 31) / 32;
        DevBuf<unsigned> flags((size_t) N * words);
        PFDR_LAUNCH((proj_serial_kernel<real, 0>), cdiv(N, 128), 128, 0, dX, dM.p, D, N, nm, dA.p, na, flags.p, words);
        if (!inplace) copy_out(X, dX, n);
        PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
        return;
    }
    if (!inplace) copy_out(X, dX, n);
    PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
}
template void proj_simplex_metric<float>(float *, const float *, int, int, int, const float *, int);
template void proj_simplex_metric<double>(double *, const double *, int, int, int, const double *, int);

}  // namespace pfdr
