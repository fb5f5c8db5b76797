// simplex_dia.cu - the simplex-family kernels on constant-offset graphs (dia.cuh): same arithmetic, in the same
// order, as the half-edge-table kernels of simplex.cu (results are bit-identical), but every address is computed
// from the vertex number: no index stream, no index -> gather dependency, 24 B per edge less HBM traffic.
//
// Work split as in simplex.cu: a lane owns LPT consecutive labels of a vertex (one 16-byte access per K-vector
// piece: LPT = 4 in float32, 2 in float64), TPV lanes per vertex, C = 32 / TPV consecutive vertices per warp.
// The 2.ND half-edges of a vertex are taken G at a time: all 4.G vector loads of a group (own copy, twin copy,
// neighbour's forward point and metric) are independent and issued back to back.
// ref src/PFDR_graph_loss_d1_simplex.cpp:474-694 (iteration), :64-370 (preconditioning).
#include "dia.cuh"
#include "simplex.cuh"
#include "simplex_dia.cuh"

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

// ---------------------------------------------------------------------------------------------
// first preconditioning (recond = 0) and reconditioning (recond = 1): as simplex_precond_kernel of simplex.cu
// ---------------------------------------------------------------------------------------------
template <typename real, int LPT, int ND>
__global__ void __launch_bounds__(256) simplex_dia_precond_kernel(SP<real> p, const real *__restrict__ Pinit, int recond)
{
    typedef DiaHalf<ND> H;
    const int lane = threadIdx.x & 31;
    const long long chunk = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chunk >= ((long long) (p.vend - p.vbeg) + p.C - 1) >> p.logC) return;     // whole warp leaves together
    const int K = p.K, Kv = p.Kv, TPV = p.TPV;
    const int vl = lane >> p.logTPV, k0 = (lane & (TPV - 1)) * LPT;
    const int v = p.vbeg + (int) (chunk << p.logC) + vl;          // owned vertices only (slab-sharded: ghosts come from their owners)
    const bool on = (v < p.vend) && (k0 < Kv);
    const int vs = on ? v : p.vbeg;
    const size_t VK = (size_t) p.V * K, Vp = (size_t) p.V;
    const size_t vk = (size_t) vs * K + (on ? k0 : 0);
    const int cur = recond ? (p.ctl->it & 1) : 0;
    real *Pb = cur ? p.P1 : p.P0, *Fb = cur ? p.F1 : p.F0;
    real *Z = cur ? p.Z1 : p.Z0;
    const bool perk_old = recond && (p.ctl->scratch[0] != 0.0);   // weights already per label before this call
    const int off[3] = {p.dia_off[0], p.dia_off[1], p.dia_off[2]};
    const unsigned m = on ? (unsigned) p.dmask[vs] : 0u;
    real P[LPT], q[LPT], S[LPT], ga_old[LPT], gaq_old[LPT], sinv_old[LPT];
#pragma unroll
    for (int i = 0; i < LPT; i++) {
        const bool li = on && k0 + i < Kv;
        P[i] = li ? (recond ? Pb[vk + i] : Pinit[vk + i]) : (real) 0;
        q[i] = li ? p.Q[vk + i] : (real) 0;
        S[i] = 0;
        ga_old[i] = (recond && li) ? p.GaU[vk + i] : (real) 1;
        gaq_old[i] = (recond && li) ? p.GaQ[vk + i] : (real) 0;
        sinv_old[i] = (recond && li && perk_old) ? p.Sinvk[vk + i] : (real) 0;
    }
    real la_h[2 * ND], wnew[2 * ND][LPT];
#pragma unroll
    for (int h = 0; h < 2 * ND; h++) {
        la_h[h] = 0;
        if (!((m >> H::mask_bit(h)) & 1u)) continue;
        const int i_off = H::idx(h);
        const int nb = H::out(h) ? v + off[i_off] : v - off[i_off];
        const real la = p.La_d1[dia_edge_id<ND>(p.eoff, p.dmask, v, m, h, off)];
        la_h[h] = la;
        const size_t zk = (size_t) H::own_plane(h) * VK + vk;
#pragma unroll
        for (int i = 0; i < LPT; i++) if (k0 + i < Kv) {
            if (!recond) { S[i] += la; Z[zk + i] = P[i]; }                       // ref :419-430
            else {
                const real Wold = perk_old ? p.wk[zk + i] * sinv_old[i] : p.Wrow[(size_t) h * Vp + v];
                Z[zk + i] = (Wold / ga_old[i]) * (explicit_point(p.loss, p.aK1, P[i], gaq_old[i], q[i]) - Z[zk + i]);   // :137-156
                real d = sabs(P[i] - Pb[(size_t) nb * K + k0 + i]);                                                  // :210-213
                if (d < p.condMin) d = p.condMin;
                const real wv = la / d;
                p.wk[zk + i] = wv; p.thk[zk + i] = d; S[i] += wv;
                wnew[h][i] = wv;
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
        if (!recond && k0 == 0) {
            const real sinv0 = (real) 1 / S[0];
            p.Sinv[v] = sinv0;
#pragma unroll
            for (int h = 0; h < 2 * ND; h++) p.Wrow[(size_t) h * Vp + v] = la_h[h] * sinv0;     // row weight w/S (ref :222-241)
        }
#pragma unroll
        for (int i = 0; i < LPT; i++) if (k0 + i < Kv) {
            const real sinv = (real) 1 / S[i], gaq = metric_times_info(p, v, ga[i], q[i]);
            if (p.GaU) p.GaU[vk + i] = ga[i];
            p.T[vk + i] = ga[i] * S[i]; p.GaQ[vk + i] = gaq;
            if (recond) {
                p.Sinvk[vk + i] = sinv; p.Gn[vk + i] = ga[i] / gmax;
#pragma unroll
                for (int h = 0; h < 2 * ND; h++) {
                    if (!((m >> H::mask_bit(h)) & 1u)) continue;
                    const size_t zk = (size_t) H::own_plane(h) * VK + vk + i;
                    const real W = wnew[h][i] * sinv;
                    if (p.loss == LOSS_QUAD) Z[zk] = P[i] - gaq * (P[i] - q[i] + Z[zk] / W);                    // :348-350
                    else Z[zk] = explicit_point(p.loss, p.aK1, P[i], gaq, q[i]) - (ga[i] / W) * Z[zk];        // :345-353
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

// ---------------------------------------------------------------------------------------------
// fused iteration (ref :569-691)
// ---------------------------------------------------------------------------------------------
// Persistent warps: the grid is one wave of resident CTAs and every warp walks its own sequence of chunks (C consecutive
// vertices) with no block-level synchronisation inside the loop - the one-CTA-per-64-vertices form spent 18 % of its
// issue slots' stall time at the CTA barriers of the control-block broadcast and of the final reduction (ncu, r01), and
// a finished warp kept its registers idle until the slowest warp of its CTA was done.  Loads never depend on the edge
// mask: the neighbour index is clamped into [0, V) and a half-edge that does not exist is dropped at compute time, so
// the whole group of loads is issued as soon as the chunk number is known (which is fetched one chunk ahead).
template <typename real, int LPT, bool PERK, int LOSS, int ND, int G, int MINB, bool SLAB = false>
__global__ void __launch_bounds__(256, MINB) simplex_dia_iter_kernel(SP<real> p)
{
    typedef VecN<real, LPT> vec;
    typedef DiaHalf<ND> H;
    __shared__ double smem[32];
    Ctl *ctl = p.ctl;
    const CtlView cv = load_ctl(ctl);
    if (cv.hold != 0 || cv.state != ST_RUN) {
        if (cv.hold == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctl->hold = 1;
        return;
    }
    const int it = cv.it, cur = it & 1;
    if (SLAB) {             // neighbours' pushes of the previous trip have landed; stop / recondition decision over all ranks
        __shared__ int s_state;
        if (threadIdx.x == 0) s_state = simplex_slab_decide(p, it);
        __syncthreads();
        if (s_state != ST_RUN) return;
    }
    const real *__restrict__ Zin = cur ? p.Z1 : p.Z0;
    real *__restrict__ Zout = cur ? p.Z0 : p.Z1;
    const real *__restrict__ Pc = cur ? p.P1 : p.P0;
    const real *__restrict__ Fc = cur ? p.F1 : p.F0;
    real *__restrict__ Pn = cur ? p.P0 : p.P1;
    real *__restrict__ Fn = cur ? p.F0 : p.F1;
    const real *__restrict__ Tg = p.T;
    const real *__restrict__ Wrow = p.Wrow;
    const int *__restrict__ order = p.order;
    const int K = p.K, TPV = p.TPV, lane = threadIdx.x & 31;
    const int vl = lane >> p.logTPV, k0 = (lane & (TPV - 1)) * LPT;
    const int nchunks = (int) (((long long) (p.vend - p.vbeg) + p.C - 1) >> p.logC);
    const int nwarps = (int) gridDim.x * (int) (blockDim.x >> 5);        // both far below 2^31
    const size_t VK = (size_t) p.V * K, Vp = (size_t) p.V;
    const int Vm1 = p.V - 1;
    const real rho = p.rho;
    const int off[3] = {p.dia_off[0], p.dia_off[1], p.dia_off[2]};
    double red = 0.0;
    int bcount = 0;                 // slab-sharded: boundary chunks come first in the work order (slab.cuh)
    int idx = (int) blockIdx.x * (int) (blockDim.x >> 5) + (int) (threadIdx.x >> 5);
    int chunk_next = idx < nchunks ? (order ? __ldg(order + idx) : idx) : 0;
    for (; idx < nchunks; idx += nwarps) {                                    // warp-uniform trip count
        const int chunk = chunk_next;
        if (idx + nwarps < nchunks) chunk_next = order ? __ldg(order + (idx + nwarps)) : idx + nwarps;
        if (SLAB) {
            const bool bchunk = idx < p.slab.nbchunks;
            if (!bchunk && bcount) { slab_boundary_done(p.slab, &ctl->bdone, bcount, it); bcount = 0; }
            if (bchunk) bcount++;
        }
        const int v = p.vbeg + (chunk << p.logC) + vl;
        const bool on = (v < p.vend) && (k0 < K);
        const int vs = on ? v : p.vbeg;
        const size_t vk = (size_t) vs * K + (on ? k0 : 0);
        const unsigned m = on ? (unsigned) __ldg(p.dmask + vs) : 0u;
        const vec pold = ldv<real, LPT>(Pc + vk), fown = ldv<real, LPT>(Fc + vk), Tv = ldv<real, LPT>(Tg + vk);
        vec Sv;
        if (PERK) Sv = ldv<real, LPT>(p.Sinvk + vk); else Sv = splat<real, LPT>((real) 1);   // row weights hold w/S already
        vec acc = splat<real, LPT>((real) 0), gaq;
#pragma unroll
        for (int h0 = 0; h0 < 2 * ND; h0 += G) {
            vec zo[G], zt[G], fn[G], tn[G], wvk[G], th[G];
            real wr[G];
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int h = h0 + u;
                if (h < 2 * ND) {
                    const int i_off = H::idx(h);
                    const int nb = H::out(h) ? min(vs + off[i_off], Vm1) : max(vs - off[i_off], 0);
                    const size_t nk = (size_t) nb * K + (on ? k0 : 0), zk = (size_t) H::own_plane(h) * VK + vk;
                    zo[u] = ldv<real, LPT>(Zin + zk);
                    zt[u] = ldv<real, LPT>(Zin + ((size_t) H::twin_plane(h) * VK + nk));
                    fn[u] = ldv<real, LPT>(Fc + nk);
                    if (LOSS != LOSS_LIN) tn[u] = ldv<real, LPT>(Tg + nk);
                    if (PERK) { wvk[u] = ldv<real, LPT>(p.wk + zk); th[u] = ldv<real, LPT>(p.thk + zk); }
                    else wr[u] = __ldg(Wrow + ((size_t) h * Vp + vs));
                }
            }
            if (h0 + G >= 2 * ND) gaq = ldv<real, LPT>(p.GaQ + vk);           // for the epilogue, behind the last group's loads
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int h = h0 + u;
                if (h < 2 * ND && ((m >> H::mask_bit(h)) & 1u)) {
                    const vec zn = prox_row<real, LPT, PERK, LOSS>(rho, fown, pold, Tv, Sv, zo[u], zt[u], fn[u], tn[u], wvk[u], th[u], wr[u], acc);
                    stv<real, LPT>(Zout + ((size_t) H::own_plane(h) * VK + vk), zn);
                    if (SLAB) {     // the other end belongs to the rank above / below: its twin copy lives there too
                        if (H::out(h)) { if (v + off[H::idx(h)] >= p.vend) stv<real, LPT>(p.phi.Z[cur ^ 1] + ((size_t) H::own_plane(h) * p.phi.plane + (size_t) ((long long) vk + p.phi.shift)), zn); }
                        else if (v - off[H::idx(h)] < p.vbeg) stv<real, LPT>(p.plo.Z[cur ^ 1] + ((size_t) H::own_plane(h) * p.plo.plane + (size_t) ((long long) vk + p.plo.shift)), zn);
                    }
                }
            }
        }
        real *flo = nullptr, *fhi = nullptr, *xlo = nullptr, *xhi = nullptr;
        if (SLAB && on) {
            if (p.plo.F[0] && v < p.vbeg + p.slab.band) { flo = p.plo.F[cur ^ 1] + ((long long) vk + p.plo.shift); xlo = p.plo.X[cur ^ 1] + ((long long) vk + p.plo.shift); }
            if (p.phi.F[0] && v >= p.vend - p.slab.band) { fhi = p.phi.F[cur ^ 1] + ((long long) vk + p.phi.shift); xhi = p.phi.X[cur ^ 1] + ((long long) vk + p.phi.shift); }
        }
        red += simplex_epilogue<real, LPT, PERK, LOSS>(p, ctl, on, v, k0, vk, acc, Tv, pold, Pn, Fn, &gaq, flo, fhi, xlo, xhi);
    }
    if (SLAB && bcount) slab_boundary_done(p.slab, &ctl->bdone, bcount, it);
    double tot[1] = {red};
    if (grid_sum_s<1>(ctl, p.partial, tot, smem) && threadIdx.x == 0) {
        if (SLAB) {
            ctl->it = it + 1;
            const double v4[SLAB_NSUM] = {tot[0], 0.0, p.obj_part ? *p.obj_part : 0.0, 0.0};
            if (ctl->track || p.dObj) slab_post(p.slab, v4, (unsigned long long) (it + 1));   // only when someone waits for them
        } else if (p.sharded) p.sums[p.rank] = tot[0];
        else simplex_finalize(p, it, tot[0]);
    }
}

// CTAs of `kernel` that are resident at once on the current device (persistent grids are exactly one wave)
template <typename Kernel>
int resident_ctas(Kernel kernel, int threads, int expect_per_sm)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = expect_per_sm;
    }
    return num_sms() * per_sm;
}

template <typename real, int LPT, bool PERK, int LOSS, int ND>
void launch_iter_nd(const SP<real> &p, int grid)
{
    // half-edges per group / resident CTAs per SM: 2 x 4 vector loads of 16 bytes in flight per lane at 3 CTAs per SM
    // (85 registers); 1-D chains take their two half-edges at once
    // float32: 2 half-edges (8 vector loads of 16 bytes per lane) in flight at 3 CTAs per SM; float64 (4 labels = 32 bytes
    // per lane and vector): one half-edge at a time at 2 CTAs per SM
    constexpr int G = sizeof(real) == 4 ? 2 : 1;
    constexpr int MINB = (PERK || sizeof(real) == 8) ? 2 : 3;
    // one wave of resident CTAs (persistent warps); small problems: one warp per chunk
    const int wave = resident_ctas(simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, G, MINB>, 256, MINB);
#ifdef PFDR_TUNE
    // tuning build only (-DPFDR_TUNE): other group sizes / occupancies of the C4 kernel, chosen by PFDR_B200_DIAV
    if constexpr (sizeof(real) == 4 && !PERK && LOSS == LOSS_KLA && ND == 3) {
        const char *e = getenv("PFDR_B200_DIAV");
        const int var = e ? atoi(e) : 0;
        const int ws = getenv("PFDR_B200_WAVES") ? atoi(getenv("PFDR_B200_WAVES")) : 1;
        if (var == 1) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, 3, 2>), std::min(grid, num_sms() * 2 * ws), 256, 0, p); return; }
        if (var == 2) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, 2, 2>), std::min(grid, num_sms() * 2 * ws), 256, 0, p); return; }
        if (var == 3) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, 1, 4>), std::min(grid, num_sms() * 4 * ws), 256, 0, p); return; }
        if (var == 4) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, 6, 2>), std::min(grid, num_sms() * 2 * ws), 256, 0, p); return; }
        if (var == 5) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, 2, 4>), std::min(grid, num_sms() * 4 * ws), 256, 0, p); return; }
        if (ws > 1) { PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, G, MINB>), std::min(grid, wave * ws), 256, 0, p); return; }
    }
#endif
    if (p.slab.on) {
        const int w2 = resident_ctas(simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, G, MINB, true>, 256, MINB);
        PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, G, MINB, true>), std::min(grid, w2), 256, 0, p);
        return;
    }
    PFDR_LAUNCH((simplex_dia_iter_kernel<real, LPT, PERK, LOSS, ND, G, MINB>), std::min(grid, wave), 256, 0, p);
}
template <typename real, int LPT, bool PERK, int LOSS>
void launch_iter_loss(const SP<real> &p, int grid)
{
    if (p.dia_nd == 1) launch_iter_nd<real, LPT, PERK, LOSS, 1>(p, grid);
    else if (p.dia_nd == 2) launch_iter_nd<real, LPT, PERK, LOSS, 2>(p, grid);
    else launch_iter_nd<real, LPT, PERK, LOSS, 3>(p, grid);
}
template <typename real, int LPT, bool PERK>
void launch_iter_perk(const SP<real> &p, int grid)
{
    if (p.loss == LOSS_LIN) launch_iter_loss<real, LPT, PERK, LOSS_LIN>(p, grid);
    else if (p.loss == LOSS_QUAD) launch_iter_loss<real, LPT, PERK, LOSS_QUAD>(p, grid);
    else launch_iter_loss<real, LPT, PERK, LOSS_KLA>(p, grid);
}

}  // namespace

template <typename real>
void simplex_dia_launch_iter(const SP<real> &p, int grid)
{
    constexpr int LPT = SimplexDiaLanes<real>::LPT;
    if (p.wk) launch_iter_perk<real, LPT, true>(p, grid);
    else launch_iter_perk<real, LPT, false>(p, grid);
}
template void simplex_dia_launch_iter<float>(const SP<float> &, int);
template void simplex_dia_launch_iter<double>(const SP<double> &, int);

template <typename real>
void simplex_dia_launch_precond(const SP<real> &p, const real *Pinit, int recond, int grid)
{
    constexpr int LPT = SimplexDiaLanes<real>::LPT;
    if (p.dia_nd == 1) PFDR_LAUNCH((simplex_dia_precond_kernel<real, LPT, 1>), grid, 256, 0, p, Pinit, recond);
    else if (p.dia_nd == 2) PFDR_LAUNCH((simplex_dia_precond_kernel<real, LPT, 2>), grid, 256, 0, p, Pinit, recond);
    else PFDR_LAUNCH((simplex_dia_precond_kernel<real, LPT, 3>), grid, 256, 0, p, Pinit, recond);
}
template void simplex_dia_launch_precond<float>(const SP<float> &, const float *, int, int);
template void simplex_dia_launch_precond<double>(const SP<double> &, const double *, int, int);

}  // namespace pfdr
