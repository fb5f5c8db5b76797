// simplex.cuh - kernel parameter block and device helpers shared by the two layouts of the simplex family
// (simplex.cu: half-edge table; simplex_dia.cu: constant-offset graphs).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "slab.cuh"

namespace pfdr {

enum : int { LOSS_LIN = 0, LOSS_QUAD = 1, LOSS_KLA = 2 };

template <typename real>
struct SP {
    Ctl *ctl;
    int K, V, E, TPV, C, logC, logTPV, loss, by_label, LPT, stream_z;
    int Kv;                           // labels that exist (K is the row stride: Kv rounded up to a multiple of 4 when padded, else Kv)
    int Vown, sharded, Vglobal;   // vertices this rank owns (== V unless sharded: ghosts follow), global count
    int world, rank;              // sharded: sums holds one value per rank, sums[rank] is ours
    double *sums;                 // [1] reduction scratch of the sharded path
    real aK, a1, aK1, rho, condMin;
    const long long *slice_off; const int *nbr, *twin, *hkey;
    const int *order;             // [owned slices] processing order for banded graphs (graph.cuh), NULL = natural
    real *w;                      // [slots]   raw weight before any reconditioning (= La_d1[e] for every label)
    real *wk, *thk;               // [slots*K] raw weight / La_d1 over it, per label, after a reconditioning
    real *Z0, *Z1;                // [slots*K] edge copies, ping/pong
    real *P0, *P1, *F0, *F1;      // [V*K] iterate and forward point, ping/pong
    real *T, *GaQ;                // [V*K] Ga.S  and  metric x first-order information (:308-335)
    real *Sinv;                   // [V]   1/S before any reconditioning
    real *Sinvk, *Gn, *GaU;       // [V*K] 1/S per label, normalised metric (after a reconditioning), raw metric
    int *lab;                     // [V] previous arg-max label (difTol >= 1)
    const real *Q, *La_f, *La_d1;
    const int *Eu, *Ev;
    double *partial;
    real *dObj, *dDif;
    // constant-offset layout (dia.cuh; dia_nd > 0): Z0/Z1 are 2.ND planes of V*K edge copies, w / wk / thk are ND
    // planes indexed by the out-end of the edge, Wrow holds the row weight w/S of every half-edge (2.ND planes of V)
    int dia_nd, dia_off[3];
    int vbeg, vend;               // vertices processed by the iteration kernel: [vbeg, vend)
    const unsigned char *dmask;   // [V]
    const int *eoff;              // [V]
    real *Wrow;                   // [2.ND * V]
    // slab-sharded solve (slab.cuh): V counts the vertices held (owned + ghost zones), [vbeg, vend) are the owned ones
    SlabDev slab;
    SlabPeer<real> plo, phi;      // the ghost zones this rank writes: in the rank below / above
    double *obj_part;             // [1] this rank's part of the objective of the current iterate (Obj requested)
};

template <typename real> __device__ __forceinline__ real sabs(real a) { return a < (real) 0 ? -a : a; }

// ---- LPT-wide vectors of labels ----------------------------------------------------------------
template <typename real, int L> struct VecN { real v[L]; };

template <typename real, int L>
__device__ __forceinline__ VecN<real, L> ldv(const real *q)
{
    VecN<real, L> r;
    if constexpr (L == 1) r.v[0] = *q;
    else if constexpr (L == 2) {
        if constexpr (sizeof(real) == 4) { const float2 t = *reinterpret_cast<const float2 *>(q); r.v[0] = t.x; r.v[1] = t.y; }
        else { const double2 t = *reinterpret_cast<const double2 *>(q); r.v[0] = t.x; r.v[1] = t.y; }
    } else {                                         // L = 4, 8, 16: pieces of four labels (16 / 2 x 16 bytes)
        static_assert(L % 4 == 0, "labels per lane: 1, 2 or a multiple of 4");
#pragma unroll
        for (int j = 0; j < L; j += 4) {
            if constexpr (sizeof(real) == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(q + j);
                r.v[j] = t.x; r.v[j + 1] = t.y; r.v[j + 2] = t.z; r.v[j + 3] = t.w;
            } else {
                const double2 a = *reinterpret_cast<const double2 *>(q + j), b = *reinterpret_cast<const double2 *>(q + j + 2);
                r.v[j] = a.x; r.v[j + 1] = a.y; r.v[j + 2] = b.x; r.v[j + 3] = b.y;
            }
        }
    }
    return r;
}
template <typename real, int L>
__device__ __forceinline__ void stv(real *q, const VecN<real, L> &r)
{
    if constexpr (L == 1) *q = r.v[0];
    else if constexpr (L == 2) {
        if constexpr (sizeof(real) == 4) *reinterpret_cast<float2 *>(q) = make_float2(r.v[0], r.v[1]);
        else *reinterpret_cast<double2 *>(q) = make_double2(r.v[0], r.v[1]);
    } else {
#pragma unroll
        for (int j = 0; j < L; j += 4) {
            if constexpr (sizeof(real) == 4) *reinterpret_cast<float4 *>(q + j) = make_float4(r.v[j], r.v[j + 1], r.v[j + 2], r.v[j + 3]);
            else {
                *reinterpret_cast<double2 *>(q + j) = make_double2(r.v[j], r.v[j + 1]);
                *reinterpret_cast<double2 *>(q + j + 2) = make_double2(r.v[j + 2], r.v[j + 3]);
            }
        }
    }
}
// streaming (evict-first) store: the new edge copies are not touched again during this launch
template <typename real, int L>
__device__ __forceinline__ void stv_cs(real *q, const VecN<real, L> &r)
{
    if constexpr (L == 1) __stcs(q, r.v[0]);
    else if constexpr (L == 2) {
        if constexpr (sizeof(real) == 4) __stcs(reinterpret_cast<float2 *>(q), make_float2(r.v[0], r.v[1]));
        else __stcs(reinterpret_cast<double2 *>(q), make_double2(r.v[0], r.v[1]));
    } else {
#pragma unroll
        for (int j = 0; j < L; j += 4) {
            if constexpr (sizeof(real) == 4) __stcs(reinterpret_cast<float4 *>(q + j), make_float4(r.v[j], r.v[j + 1], r.v[j + 2], r.v[j + 3]));
            else {
                __stcs(reinterpret_cast<double2 *>(q + j), make_double2(r.v[j], r.v[j + 1]));
                __stcs(reinterpret_cast<double2 *>(q + j + 2), make_double2(r.v[j + 2], r.v[j + 3]));
            }
        }
    }
}
template <typename real, int L>
__device__ __forceinline__ VecN<real, L> splat(real x)
{
    VecN<real, L> r;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = x;
    return r;
}

template <typename real>
__device__ __forceinline__ real group_sum(real x, int TPV)
{
    for (int o = TPV >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <typename real>
__device__ __forceinline__ real group_max(real x, int TPV)
{
    for (int o = TPV >> 1; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

template <int NV>
__device__ bool grid_sum_s(Ctl *ctl, double *partial, double (&v)[NV], double *smem)
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

// forward point 2P - Ga.grad(P)  (ref :569-587)
template <typename real>
__device__ __forceinline__ real forward_point(int loss, real aK1, real P, real gaq, real q)
{
    if (loss == LOSS_LIN) return (real) 2 * P + gaq;
    if (loss == LOSS_QUAD) return (real) 2 * P - gaq * (P - q);
    return (real) 2 * P + gaq / (aK1 + P);
}
// P - Ga.grad(P), as the reference's subgradient expressions need it (:137-156, :337-358)
template <typename real>
__device__ __forceinline__ real explicit_point(int loss, real aK1, real P, real gaq, real q)
{
    if (loss == LOSS_LIN) return P + gaq;
    if (loss == LOSS_QUAD) return P - gaq * (P - q);
    return P + gaq / (aK1 + P);
}

// metric of one (vertex,label) from the loss Hessian and the weight sum S  (ref :160-190, :222-285)
template <typename real>
__device__ __forceinline__ real metric(const SP<real> &p, int v, real P, real q, real S)
{
    const real bound = (real) 1.9 * ((real) 2 - p.rho);
    if (p.loss == LOSS_LIN) return (real) 1 / S;
    if (p.loss == LOSS_QUAD) {
        real ga = (real) 1 / ((p.La_f ? p.La_f[v] : (real) 1) + S);
        if (p.La_f) { const real c = bound / p.La_f[v]; if (ga > c) ga = c; }
        else if (bound < (real) 1) { if (ga > bound) ga = bound; }
        return ga;
    }
    const real a = p.aK1 + P, cq = p.aK + p.a1 * q;
    const real h = p.La_f ? p.La_f[v] * cq / (a * a) : cq / (a * a);
    real ga = (real) 1 / (h + S);
    const real b = p.La_f ? p.La_f[v] * (real) 1 / (p.aK1 * p.aK1) : (real) 1 / (p.aK1 * p.aK1);
    const real c = bound / (cq * b);
    if (ga > c) ga = c;
    return ga;
}
template <typename real>
__device__ __forceinline__ real metric_times_info(const SP<real> &p, int v, real ga, real q)
{   // GaQ, ref :308-335
    if (p.loss == LOSS_LIN) return ga * q;
    if (p.loss == LOSS_QUAD) return p.La_f ? p.La_f[v] * ga : ga;
    return p.La_f ? p.La_f[v] * ga * (p.aK + p.a1 * q) : ga * (p.aK + p.a1 * q);
}

// first arg-max label over the lanes of a vertex (ref :662-669): ties go to the lowest label
template <typename real, int LPT>
__device__ __forceinline__ int group_argmax(const VecN<real, LPT> &x, int k0, int K, int TPV)
{
    real a = -HUGE_VAL; int i = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < LPT; t++) if (k0 + t < K && x.v[t] > a) { a = x.v[t]; i = k0 + t; }
    for (int o = TPV >> 1; o > 0; o >>= 1) {
        const real b = __shfl_xor_sync(0xffffffffu, a, o);
        const int j = __shfl_xor_sync(0xffffffffu, i, o);
        if (b > a || (b == a && j < i)) { a = b; i = j; }
    }
    return i;
}

// ---------------------------------------------------------------------------------------------
// fused iteration
// ---------------------------------------------------------------------------------------------
// d1 prox + relaxation of one half-edge row for the lane's LPT labels (ref :590-634); accumulates W.z into acc
template <typename real, int LPT, bool PERK, int LOSS>
__device__ __forceinline__ VecN<real, LPT> prox_row(real rho, const VecN<real, LPT> &fown, const VecN<real, LPT> &pold,
    const VecN<real, LPT> &Tv, const VecN<real, LPT> &Sv, const VecN<real, LPT> &zo, const VecN<real, LPT> &zt,
    const VecN<real, LPT> &fn, const VecN<real, LPT> &tn, const VecN<real, LPT> &wvk, const VecN<real, LPT> &th, real w,
    VecN<real, LPT> &acc)
{
    VecN<real, LPT> zn;
    const real ws = w * Sv.v[0];                                              // row weight W = w/S (per vertex before recond.)
#pragma unroll
    for (int i = 0; i < LPT; i++) {
        const real po = fown.v[i] - zo.v[i], pt = fn.v[i] - zt.v[i];
        const real d0 = po - pt;
        if (LOSS == LOSS_LIN) {                                               // ref :599-614 (weights 1/2, threshold 2)
            const real ds = d0 - min(max(d0, (real) -2), (real) 2);
            zn.v[i] = zo.v[i] + rho * ((real) 0.5 * (po + pt) + (real) 0.5 * ds - pold.v[i]);
        } else {                                                              // ref :615-631
            const real tsum = Tv.v[i] + tn.v[i], inv = fast_rcp(tsum);
            const real t = PERK ? th.v[i] * tsum : tsum;                      // La_d1/w = 1 before a reconditioning
            const real wo = tn.v[i] * inv, wt = Tv.v[i] * inv;                // W_d1 own / twin side
            const real ds = d0 - min(max(d0, -t), t);                         // soft threshold, exact 0 inside
            // W_d1u.pu + W_d1v.pv = pt + wo.(po - pt) since the two weights sum to one
            zn.v[i] = zo.v[i] + rho * (fma(wo, d0, pt) + wt * ds - pold.v[i]);
        }
        acc.v[i] += (PERK ? wvk.v[i] * Sv.v[i] : ws) * zn.v[i];
    }
    return zn;
}

// projection on the simplex in the metric m (Michelot's active-set iteration, all lanes of the vertex in parallel,
// ref :651 / proj_simplex_metric.cpp:40-81), next (P, FP) (:569-587) and this lane's share of the evolution (:654-691)
template <typename real, int LPT, bool PERK, int LOSS>
__device__ __forceinline__ double simplex_epilogue(const SP<real> &p, Ctl *ctl, bool on, int v, int k0, size_t vk,
    VecN<real, LPT> acc, const VecN<real, LPT> &Tv, const VecN<real, LPT> &pold, real *__restrict__ Pn, real *__restrict__ Fn,
    const VecN<real, LPT> *gaq_pre = nullptr,       // GaQ of the lane's labels when the caller fetched it ahead of time
    real *Fpush_lo = nullptr, real *Fpush_hi = nullptr,   // slab-sharded: where a neighbouring rank keeps this vertex's forward point
    real *Ppush_lo = nullptr, real *Ppush_hi = nullptr)   // ... and its iterate (read there by the objective and by a reconditioning)
{
    typedef VecN<real, LPT> vec;
    const int K = p.K, TPV = p.TPV;
    vec m;
    if (PERK) { if (on) m = ldv<real, LPT>(p.Gn + vk); else m = splat<real, LPT>((real) 1); }
    else if (LOSS == LOSS_LIN) m = splat<real, LPT>((real) 1);     // Ga = 1/S is the same for all labels
    else m = Tv;
    vec xs;
    unsigned act = 0;
#pragma unroll
    for (int i = 0; i < LPT; i++) {
        if (!on) { m.v[i] = (real) 1; acc.v[i] = 0; }
        xs.v[i] = acc.v[i] * fast_rcp(m.v[i]);
        if (on && k0 + i < p.Kv) act |= 1u << i;
    }
    real la = 0;
    for (;;) {
        real sx = 0, sm = 0;
#pragma unroll
        for (int i = 0; i < LPT; i++) if ((act >> i) & 1u) { sx += acc.v[i]; sm += m.v[i]; }
        sx = group_sum(sx, TPV); sm = group_sum(sm, TPV);
        la = (sx - (real) 1) * fast_rcp(sm);
        unsigned keep = act;
#pragma unroll
        for (int i = 0; i < LPT; i++) if (((act >> i) & 1u) && xs.v[i] < la) keep &= ~(1u << i);
        const bool changed = keep != act;
        act = keep;
        if (!__any_sync(0xffffffffu, changed)) break;
    }
    vec pnew;
#pragma unroll
    for (int i = 0; i < LPT; i++) pnew.v[i] = ((act >> i) & 1u) ? (xs.v[i] - la) * m.v[i] : (real) 0;
    if (on) {
        const vec gaq = gaq_pre ? *gaq_pre : ldv<real, LPT>(p.GaQ + vk);
        vec qv = splat<real, LPT>((real) 0), f;
        if (LOSS == LOSS_QUAD) qv = ldv<real, LPT>(p.Q + vk);
#pragma unroll
        for (int i = 0; i < LPT; i++) {
            if (LOSS == LOSS_KLA) f.v[i] = (real) 2 * pnew.v[i] + gaq.v[i] * fast_rcp(p.aK1 + pnew.v[i]);
            else f.v[i] = forward_point(LOSS, p.aK1, pnew.v[i], gaq.v[i], qv.v[i]);
        }
        stv<real, LPT>(Pn + vk, pnew);
        stv<real, LPT>(Fn + vk, f);
        if (Fpush_lo) { stv<real, LPT>(Fpush_lo, f); stv<real, LPT>(Ppush_lo, pnew); }
        if (Fpush_hi) { stv<real, LPT>(Fpush_hi, f); stv<real, LPT>(Ppush_hi, pnew); }
    }
    double red = 0.0;
    if (p.by_label) {                                                         // ref :657-676
        const int i = group_argmax<real, LPT>(pnew, on ? k0 : p.Kv, p.Kv, TPV);
        if (on && k0 == 0 && ctl->track) { if (i != p.lab[v]) { red = 1.0; p.lab[v] = i; } }
    } else if (on) {                                                          // ref :681-687
        real d = 0;
#pragma unroll
        for (int i = 0; i < LPT; i++) d += sabs(pold.v[i] - pnew.v[i]);
        red = (double) d;
    }
    return red;
}

// end of an iteration: evolution measure (ref :654-691), iteration counter, stop / recondition decision (:551, :554)
template <typename real>
__device__ void simplex_finalize(const SP<real> &p, int it, double s)
{
    Ctl *ctl = p.ctl;
    double dif = ctl->dif;
    if (ctl->track) {
        dif = p.by_label ? s : s / (double) p.Vglobal;              // ref :688 divides by the (global) vertex count
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
__device__ int simplex_slab_decide(const SP<real> &p, int it)
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
    double dif = ctl->dif, s = 0.0, obj = 0.0;
    if (ok && all) {
        for (int q = 0; q < p.slab.world; q++) {                    // rank order: same bits everywhere
            ok = slab_wait(&me->sum_flag[q], target, me) && ok;
            s += me->sums[(it - 1) & 1][q][0]; obj += me->sums[(it - 1) & 1][q][2];
        }
        if (track) dif = p.by_label ? s : s / (double) p.Vglobal;   // ref :688
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
