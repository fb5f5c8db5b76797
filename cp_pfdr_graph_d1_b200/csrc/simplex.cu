// simplex.cu - PFDR for  loss(P;Q) + sum_e La_d1 sum_k |P_uk - P_vk|  with every P_v on the simplex,
// the B200-native replacement of src/PFDR_graph_loss_d1_simplex.cpp (+ src/proj_simplex_metric.cpp).
//
// One fused kernel per iteration.  TPV = next power of two >= K lanes cooperate on one vertex (lane = label),
// C = 32/TPV vertices share a warp and one SELL-C slice of the half-edge table (graph.cuh); the K labels of
// a half-edge are contiguous, so every load of edge copies is one coalesced row.  Per trip and per
// (half-edge, label): explicit step on both ends is already folded into the stored forward point FP
// (ref :569-587), d1 prox + relaxation (:590-634) for the lane's own half, weighted average in edge order
// (:637-648); then the K lanes of a vertex run proj_simplex_metric (:651, proj_simplex_metric.cpp:40-81)
// redundantly through warp shuffles, write (P, FP) of the next trip and add up the evolution (:654-691).
// Edge constants are recomputed from per-(vertex,label) S and T = Ga.S exactly as in quad.cu.
#include "graph.cuh"
#include "solvers.cuh"

namespace pfdr {

template <typename real> using pair_t = typename Pair<real>::type;

enum : int { LOSS_LIN = 0, LOSS_QUAD = 1, LOSS_KLA = 2 };

template <typename real>
struct SP {
    Ctl *ctl;
    int K, V, E, TPV, C, logC, logTPV, loss, by_label;
    real aK, a1, aK1, rho, condMin;
    const long long *slice_off; const int *nbr, *twin, *hkey;
    real *w;                      // [slots]   raw weight (first preconditioning: La_d1[e] for every label)
    real *wk, *thk;               // [slots*K] raw weight / its La_d1 ratio per label after a reconditioning
    real *Z0, *Z1;                // [slots*K] edge copies
    pair_t<real> *PF0, *PF1;      // [V*K] (P, forward point)
    real *T, *Sinv, *Gn, *GaQ, *GaU;   // [V*K] Ga.S, 1/S, normalised metric, metric x first order info, raw metric
    int *lab;                     // [V] previous arg-max label (difTol >= 1)
    const real *Q, *La_f, *La_d1;
    const int *Eu, *Ev;
    double *partial;
    real *dObj, *dDif;
};

template <typename real> __device__ __forceinline__ real sabs(real a) { return a < (real) 0 ? -a : a; }

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
__device__ __forceinline__ real forward_point(const SP<real> &p, real P, real gaq, real q)
{
    if (p.loss == LOSS_LIN) return (real) 2 * P + gaq;
    if (p.loss == LOSS_QUAD) return (real) 2 * P - gaq * (P - q);
    return (real) 2 * P + gaq / (p.aK1 + P);
}

// proj_simplex_metric on the TPV lanes of one vertex (lane k holds x_k, m_k); every lane runs the
// reference's sequential active-set recursion (proj_simplex_metric.cpp:43-81) on shuffled values, so the
// result is the reference's up to the rounding of the inputs.  All 32 lanes must call.
template <typename real>
__device__ __forceinline__ real simplex_project(real x, real m, real target, int k, int K, int TPV)
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

// first arg-max label of the K lanes of a vertex (ref :662-669); all 32 lanes must call
template <typename real>
__device__ __forceinline__ int argmax_label(real x, int K, int TPV)
{
    real a = __shfl_sync(0xffffffffu, x, 0, TPV);
    int i = 0;
    for (int d = 1; d < K; d++) {
        const real xd = __shfl_sync(0xffffffffu, x, d, TPV);
        if (xd > a) { a = xd; i = d; }
    }
    return i;
}

template <typename real>
__device__ __forceinline__ real group_max(real x, int TPV)
{
    for (int o = TPV >> 1; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
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
    real h = p.La_f ? p.La_f[v] * cq / (a * a) : cq / (a * a);
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

// gradient step folded as the reference's subgradient expressions need it:  P - Ga.grad(P)
template <typename real>
__device__ __forceinline__ real explicit_point(const SP<real> &p, real P, real gaq, real q)
{
    if (p.loss == LOSS_LIN) return P + gaq;
    if (p.loss == LOSS_QUAD) return P - gaq * (P - q);
    return P + gaq / (p.aK1 + P);
}

// ---------------------------------------------------------------------------------------------
// first preconditioning (recond = 0, ref :160-370 with Zu == NULL) and reconditioning (recond = 1)
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) simplex_precond_kernel(SP<real> p, const real *__restrict__ P0, int recond)
{
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slice >= ((long long) p.V + p.C - 1) >> p.logC) return;     // whole warp leaves together
    const int K = p.K, TPV = p.TPV;
    const int vl = lane >> p.logTPV, k = lane & (TPV - 1);
    const int v = (int) (slice << p.logC) + vl;
    const bool on = (v < p.V) && (k < K);
    const size_t vk = (size_t) (on ? v : 0) * K + (on ? k : 0);
    const int cur = recond ? (p.ctl->it & 1) : 0;
    pair_t<real> *PF = cur ? p.PF1 : p.PF0;
    real *Z = cur ? p.Z1 : p.Z0;
    const real P = recond ? PF[vk].x : P0[vk];
    const real q = p.Q[vk];
    const long long base = p.slice_off[slice];
    const int width = (int) ((p.slice_off[slice + 1] - base) >> p.logC);
    const real ga_old = recond ? p.GaU[vk] : (real) 0, gaq_old = recond ? p.GaQ[vk] : (real) 0;
    const real sinv_old = recond ? p.Sinv[vk] : (real) 0;
    const bool perk_old = recond && (p.ctl->scratch[0] != 0.0);   // weights already per label before this call
    real S = 0;
    for (int j = 0; j < width; j++) {
        const long long s = base + ((long long) j << p.logC) + vl;
        const int n = p.nbr[s];
        if (n >= 0 && on) {
            const size_t sk = (size_t) s * K + k;
            if (!recond) {
                const real wv = p.La_d1[p.hkey[s] >> 1];                        // ref :207-208
                if (k == 0) p.w[s] = wv;
                S += wv; Z[sk] = P;                                             // ref :419-430
            } else {
                const real Wold = (perk_old ? p.wk[sk] : p.w[s]) * sinv_old;
                Z[sk] = (Wold / ga_old) * (explicit_point(p, P, gaq_old, q) - Z[sk]);      // ref :137-156
                real d = sabs(P - PF[(size_t) n * K + k].x);                              // ref :210-213
                if (d < p.condMin) d = p.condMin;
                const real wv = p.La_d1[p.hkey[s] >> 1] / d;
                p.wk[sk] = wv; p.thk[sk] = d; S += wv;
            }
        }
    }
    const real ga = on ? metric(p, v, P, q, S) : (real) 1;
    const real gmax = group_max(on ? ga : (real) 0, TPV);                     // ref :362-369
    if (on) {
        const real sinv = (real) 1 / S, gaq = metric_times_info(p, v, ga, q);
        p.GaU[vk] = ga; p.T[vk] = ga * S; p.Sinv[vk] = sinv; p.GaQ[vk] = gaq; p.Gn[vk] = ga / gmax;
        if (recond) {
            for (int j = 0; j < width; j++) {
                const long long s = base + ((long long) j << p.logC) + vl;
                if (p.nbr[s] >= 0) {
                    const size_t sk = (size_t) s * K + k;
                    const real W = p.wk[sk] * sinv;
                    if (p.loss == LOSS_QUAD) Z[sk] = P - gaq * (P - q + Z[sk] / W);       // ref :348-350
                    else Z[sk] = explicit_point(p, P, gaq, q) - (ga / W) * Z[sk];          // ref :345-347,:351-353
                }
            }
        }
        pair_t<real> o; o.x = P; o.y = forward_point(p, P, gaq, q);
        PF[vk] = o;
    }
    if (!recond && p.lab) {                                                   // ref :447-462
        const int i = argmax_label(on ? P : (real) -HUGE_VAL, K, TPV);
        if (on && k == 0) p.lab[v] = i;
    }
}

template <typename real>
__global__ void simplex_resume_kernel(SP<real> p)
{
    p.ctl->rcd *= 0.1; p.ctl->state = ST_RUN; p.ctl->hold = 0; p.ctl->scratch[0] = 1.0;     // ref :563
}

// ---------------------------------------------------------------------------------------------
// fused iteration
// ---------------------------------------------------------------------------------------------
template <typename real, bool PERK>
__global__ void __launch_bounds__(256) simplex_iter_kernel(SP<real> p)
{
    __shared__ double smem[32];
    Ctl *ctl = p.ctl;
    {
        const int hold = ld_flag(&ctl->hold), state = ld_flag(&ctl->state);
        if (hold != 0 || state != ST_RUN) {
            if (hold == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctl->hold = 1;
            return;
        }
    }
    const int it = ld_flag(&ctl->it), cur = it & 1;
    const real *__restrict__ Zin = cur ? p.Z1 : p.Z0;
    real *__restrict__ Zout = cur ? p.Z0 : p.Z1;
    const pair_t<real> *__restrict__ PFc = cur ? p.PF1 : p.PF0;
    pair_t<real> *__restrict__ PFn = cur ? p.PF0 : p.PF1;
    const int K = p.K, TPV = p.TPV, lane = threadIdx.x & 31;
    const long long slice = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nslices = ((long long) p.V + p.C - 1) >> p.logC;
    const real rho = p.rho;
    double red[1] = {0.0};
    if (slice < nslices) {                                                    // warp-uniform
        const int vl = lane >> p.logTPV, k = lane & (TPV - 1);
        const int v = (int) (slice << p.logC) + vl;
        const bool on = (v < p.V) && (k < K);
        const size_t vk = (size_t) (on ? v : 0) * K + (on ? k : 0);
        const pair_t<real> pf = PFc[vk];
        const real Tv = p.T[vk], Sv = p.Sinv[vk];
        const long long base = p.slice_off[slice];
        const int width = (int) ((p.slice_off[slice + 1] - base) >> p.logC);
        real acc = 0;
#pragma unroll 2
        for (int j = 0; j < width; j++) {
            const long long s = base + ((long long) j << p.logC) + vl;
            const int n = p.nbr[s];
            if (n >= 0 && on) {
                const int tw = p.twin[s];
                const size_t sk = (size_t) s * K + k, nk = (size_t) n * K + k;
                const real zo = Zin[sk], zt = Zin[(size_t) tw * K + k];
                const pair_t<real> pn = PFc[nk];
                const real wv = PERK ? p.wk[sk] : p.w[s];
                const real po = pf.y - zo, pt = pn.y - zt;
                real zn;
                if (p.loss == LOSS_LIN) {                                     // ref :599-614
                    const real c = (real) 0.5 * (po + pt);
                    real d = po - pt;
                    if (d > (real) 2) zn = zo + rho * (c + (real) 0.5 * (d - (real) 2) - pf.x);
                    else if (d < (real) -2) zn = zo + rho * (c + (real) 0.5 * (d + (real) 2) - pf.x);
                    else zn = zo + rho * (c - pf.x);
                } else {                                                      // ref :615-631
                    const real Tn = p.T[nk];
                    const real tsum = Tv + Tn, inv = (real) 1 / tsum;
                    const real th = PERK ? p.thk[sk] * tsum : tsum;           // La_d1/w = 1 before a reconditioning
                    const real wo = Tn * inv, wt = Tv * inv;
                    const real c = wo * po + wt * pt;
                    real d = po - pt;
                    if (d > th) zn = zo + rho * (c + wt * (d - th) - pf.x);
                    else if (d < -th) zn = zo + rho * (c + wt * (d + th) - pf.x);
                    else zn = zo + rho * (c - pf.x);
                }
                Zout[sk] = zn;
                acc += (wv * Sv) * zn;
            }
        }
        const real m = on ? p.Gn[vk] : (real) 1;
        const real pnew = simplex_project<real>(on ? acc : (real) 0, m, (real) 1, k, K, TPV);
        if (on) {
            pair_t<real> o; o.x = pnew; o.y = forward_point(p, pnew, p.GaQ[vk], p.loss == LOSS_QUAD ? p.Q[vk] : (real) 0);
            PFn[vk] = o;
        }
        if (p.by_label) {                                                     // ref :657-676
            const int i = argmax_label(on ? pnew : (real) -HUGE_VAL, K, TPV);
            if (on && k == 0 && ctl->track) { if (i != p.lab[v]) { red[0] = 1.0; p.lab[v] = i; } }
        } else if (on) red[0] = (double) sabs(pf.x - pnew);                   // ref :681-687
    }
    if (grid_sum_s<1>(ctl, p.partial, red, smem) && threadIdx.x == 0) {
        double dif = ctl->dif;
        if (ctl->track) {
            dif = p.by_label ? red[0] : red[0] / (double) p.V;
            ctl->dif = dif;
            if (p.dDif) p.dDif[it] = (real) dif;
        }
        const int itn = it + 1;
        ctl->it = itn;
        if (itn == ctl->itMax || (real) dif < (real) ctl->tol) ctl->state = ST_FINAL;   // ref :551
        else if ((real) dif < (real) ctl->rcd) ctl->state = ST_RECOND;                  // ref :554
    }
}

// objective (ref :477-544)
template <typename real>
__global__ void __launch_bounds__(256) simplex_objective_kernel(SP<real> p)
{
    __shared__ double smem[32];
    Ctl *ctl = p.ctl;
    if (ld_flag(&ctl->hold)) return;
    const int it = ld_flag(&ctl->it), K = p.K;
    const pair_t<real> *__restrict__ PF = (it & 1) ? p.PF1 : p.PF0;
    double r[1] = {0.0};
    const size_t stride = (size_t) gridDim.x * blockDim.x, t0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t0; i < (size_t) p.V * K; i += stride) {
        const real P = PF[i].x, q = p.Q[i];
        const double lf = p.La_f ? (double) p.La_f[i / K] : 1.0;
        if (p.loss == LOSS_LIN) r[0] -= (double) P * (double) q;
        else if (p.loss == LOSS_QUAD) { const double c = (double) P - (double) q; r[0] += 0.5 * lf * c * c; }
        else { const real c = p.aK + p.a1 * q; r[0] += lf * (double) c * log((double) (c / (p.aK + p.a1 * P))); }
    }
    for (size_t i = t0; i < (size_t) p.E * K; i += stride) {
        const size_t e = i / K, k = i % K;
        r[0] += (double) p.La_d1[e] * (double) sabs(PF[(size_t) p.Eu[e] * K + k].x - PF[(size_t) p.Ev[e] * K + k].x);
    }
    if (grid_sum_s<1>(ctl, p.partial, r, smem) && threadIdx.x == 0) p.dObj[it] = (real) r[0];
}

template <typename real>
__global__ void simplex_extract_kernel(SP<real> p, real *__restrict__ out)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) p.V * p.K) return;
    const pair_t<real> *PF = (p.ctl->it & 1) ? p.PF1 : p.PF0;
    out[i] = PF[i].x;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename real>
struct SimplexSolver : SolverBase {
    const SimplexProblem<real> q;
    bool keep_p0 = false;
    SellGraph g;
    SP<real> p{};
    DevBuf<Ctl> ctl;
    DevIn<real> Q, La_f, La_d1, P0;
    DevIn<int> Eu, Ev;
    DevBuf<real> w, wk, thk, Z0, Z1, T, Sinv, Gn, GaQ, GaU, dObj, dDif, Pout;
    DevBuf<pair_t<real>> PF0, PF1;
    DevBuf<int> lab;
    DevBuf<double> partial;
    Ctl *hctl = nullptr;
    cudaEvent_t ev[2]{}, ev_a{}, ev_b{}, ev_c{};
    cudaGraphExec_t gexec = nullptr;
    int graph_trips = 0, grid = 0;

    SimplexSolver(const SimplexProblem<real> &q_, bool keep) : q(q_), keep_p0(keep) {}
    ~SimplexSolver() override
    {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (hctl) cudaFreeHost(hctl);
        for (auto e : {ev[0], ev[1], ev_a, ev_b, ev_c}) if (e) cudaEventDestroy(e);
    }

    void setup()
    {
        const int K = q.K, V = q.V, E = q.E;
        PFDR_REQUIRE(K >= 1 && V > 0 && E >= 0 && q.itMax >= 0, "K, V must be positive, E and itMax non-negative");
        PFDR_REQUIRE(K <= 32, "K > 32 labels is not supported by this build (lanes of one warp hold the labels of a vertex)");
        PFDR_REQUIRE(q.al >= (real) 0 && q.al <= (real) 1, "al must lie in [0, 1] (the reference is undefined beyond, SURVEY 5.1-10)");
        cudaStream_t st = ctx().stream;
        for (auto *e : {&ev[0], &ev[1], &ev_a, &ev_b, &ev_c}) PFDR_CUDA(cudaEventCreate(e));
        PFDR_CUDA(cudaEventRecord(ev_a, st));
        PFDR_CUDA(cudaMallocHost((void **) &hctl, 2 * sizeof(Ctl)));
        int TPV = 1, logTPV = 0;
        while (TPV < K) { TPV <<= 1; logTPV++; }
        const int C = 32 / TPV;
        int logC = 0; while ((1 << logC) < C) logC++;
        const size_t VK = (size_t) V * K;
        Eu.bind(q.Eu, E); Ev.bind(q.Ev, E); Q.bind(q.Q, VK); La_d1.bind(q.La_d1, E);
        La_f.bind(q.al > (real) 0 ? q.La_f : nullptr, V); P0.bind(q.P, VK);
        build_sell(g, Eu, Ev, V, E, C);
        const size_t ns = (size_t) std::max<long long>(g.nslots, 1);
        w.alloc(ns); Z0.alloc(ns * K); Z1.alloc(ns * K);
        PF0.alloc(VK); PF1.alloc(VK); T.alloc(VK); Sinv.alloc(VK); Gn.alloc(VK); GaQ.alloc(VK); GaU.alloc(VK);
        Z0.zero(); Z1.zero(); w.zero();
        const bool by_label = q.difTol >= (real) 1;
        if (by_label) lab.alloc(V);
        if (q.Obj) dObj.alloc((size_t) q.itMax + 1);
        if (q.Dif) dDif.alloc((size_t) std::max(q.itMax, 1));
        ctl.alloc(1);
        const long long nslices = ((long long) V + C - 1) / C;
        grid = (int) cdiv((size_t) nslices * 32, 256);
        partial.alloc((size_t) std::max(grid, num_sms() * 8) * 2);

        p.ctl = ctl.p; p.K = K; p.V = V; p.E = E; p.TPV = TPV; p.C = C; p.logC = logC; p.logTPV = logTPV;
        p.loss = (q.al == (real) 0) ? LOSS_LIN : (q.al == (real) 1 ? LOSS_QUAD : LOSS_KLA);
        p.by_label = by_label;
        p.aK = q.al / (real) K; p.a1 = (real) 1 - q.al; p.aK1 = p.aK / p.a1;                     // ref :386-391
        p.rho = q.rho; p.condMin = q.condMin;
        p.slice_off = g.slice_off.p; p.nbr = g.nbr.p; p.twin = g.twin.p; p.hkey = g.hkey.p;
        p.w = w.p; p.wk = nullptr; p.thk = nullptr; p.Z0 = Z0.p; p.Z1 = Z1.p; p.PF0 = PF0.p; p.PF1 = PF1.p;
        p.T = T.p; p.Sinv = Sinv.p; p.Gn = Gn.p; p.GaQ = GaQ.p; p.GaU = GaU.p; p.lab = lab.p;
        p.Q = Q.p; p.La_f = La_f.p; p.La_d1 = La_d1.p; p.Eu = Eu.p; p.Ev = Ev.p;
        p.partial = partial.p; p.dObj = dObj.p; p.dDif = dDif.p;

        reset();
        if (!keep_p0) P0.own.release();
        PFDR_CUDA(cudaEventRecord(ev_b, st));
    }

    void reset() override
    {
        cudaStream_t st = ctx().stream;
        PFDR_REQUIRE(P0.p != nullptr, "reset: the initial iterate was not kept on the device");
        if (gexec && p.wk) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        p.wk = nullptr; p.thk = nullptr;
        Ctl h{};                                                                                  // ref :445-446
        h.it = 0; h.state = (q.itMax == 0) ? ST_FINAL : ST_RUN; h.itMax = q.itMax;
        h.track = (q.difTol > (real) 0 || q.difRcd > (real) 0 || q.Dif != nullptr);
        h.tol = (double) q.difTol; h.rcd = (double) q.difRcd;
        h.dif = (double) (q.difTol > q.difRcd ? q.difTol : q.difRcd);
        hctl[0] = h;
        PFDR_CUDA(cudaMemcpyAsync(ctl.p, &hctl[0], sizeof(Ctl), cudaMemcpyHostToDevice, st));
        PFDR_LAUNCH(simplex_precond_kernel<real>, grid, 256, 0, p, P0.p, 0);
        PFDR_CUDA(cudaStreamSynchronize(st));
    }

    void enqueue_trip()
    {
        if (q.Obj) {
            const size_t work = (size_t) std::max(q.E, q.V) * q.K;
            PFDR_LAUNCH(simplex_objective_kernel<real>, std::min<unsigned>(cdiv(work, 256), num_sms() * 8), 256, 0, p);
        }
        if (p.wk) PFDR_LAUNCH((simplex_iter_kernel<real, true>), grid, 256, 0, p);
        else PFDR_LAUNCH((simplex_iter_kernel<real, false>), grid, 256, 0, p);
    }
    int launches_per_trip() const { return (q.Obj ? 1 : 0) + 1; }

    void enqueue_batch(int batch)
    {
        cudaStream_t st = ctx().stream;
        static int use_graph = -1;
        if (use_graph < 0) { const char *e = getenv("PFDR_B200_GRAPH"); use_graph = (e && e[0] == '0') ? 0 : 1; }
        if (!use_graph) { for (int i = 0; i < batch; i++) enqueue_trip(); return; }
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
            const size_t n = (size_t) std::max<long long>(g.nslots, 1) * q.K;
            wk.alloc(n); thk.alloc(n); wk.zero(); thk.zero();
            p.wk = wk.p; p.thk = thk.p;
        }
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        PFDR_LAUNCH(simplex_precond_kernel<real>, grid, 256, 0, p, (const real *) nullptr, 1);
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
        const size_t work = (size_t) g.nslots * q.K;
        int batch = work > (1u << 24) ? 4 : (work > (1u << 20) ? 16 : 64);
        batch = std::max(1, std::min(batch, q.itMax + 1));
        ctx().in_loop = true;
        PFDR_CUDA(cudaEventRecord(ev_b, st));
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
            if (verdict == 1) break;
            if (verdict == 2) {
                PFDR_CUDA(cudaStreamSynchronize(st));
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
    }

    void fetch(void *Po, int *ito, void *Objo, void *Difo) override
    {
        cudaStream_t st = ctx().stream;
        PFDR_CUDA(cudaMemcpyAsync(&hctl[0], ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        const Ctl h = hctl[0];
        const size_t VK = (size_t) q.V * q.K;
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
        if (q.verbose) {
            printf("PFDR simplex (B200): %d iterations (max. %d), evolution %g, loop %.3f ms\n", h.it, q.itMax, h.dif,
                   ctx().stats.loop_ms);
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
    SimplexSolver<real> s(q, false);
    s.setup();
    s.run();
    s.fetch(q.P, q.it, q.Obj, q.Dif);
}
template void solve_simplex<float>(const SimplexProblem<float> &);
template void solve_simplex<double>(const SimplexProblem<double> &);

// ---------------------------------------------------------------------------------------------
// stand-alone proj_simplex_metric (ref src/proj_simplex_metric.cpp:18-83)
// ---------------------------------------------------------------------------------------------
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
    const real r = simplex_project<real>(x, m, a, d, D, TPV);
    if (on) X[xi] = r;
}

// any D: one thread per column, the reference's loops verbatim on global memory, flags in a bit set
template <typename real, int WORDS>
__global__ void __launch_bounds__(128) proj_serial_kernel(real *__restrict__ X, const real *__restrict__ M, int D, int N,
                                                          int nm, const real *__restrict__ A, int na)
{
    const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    real *x = X + (size_t) n * D;
    const real *m = M + (size_t) (n < nm ? n : nm - 1) * D;
    const real a = A[n < na ? n : na - 1];
    unsigned act[WORDS];
    for (int i = 0; i < WORDS; i++) act[i] = 0;
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
    PFDR_REQUIRE(D <= 4096, "proj_simplex_metric: D > 4096 is not supported");
    if (N == 0) return;
    const size_t n = (size_t) D * N;
    const bool inplace = is_device_ptr(X);
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
    } else {
        PFDR_LAUNCH((proj_serial_kernel<real, 128>), cdiv(N, 128), 128, 0, dX, dM.p, D, N, nm, dA.p, na);
    }
    if (!inplace) copy_out(X, dX, n);
    PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
}
template void proj_simplex_metric<float>(float *, const float *, int, int, int, const float *, int);
template void proj_simplex_metric<double>(double *, const double *, int, int, int, const double *, int);

}  // namespace pfdr
