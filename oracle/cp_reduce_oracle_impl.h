/* TEST INFRASTRUCTURE ONLY - serial CPU restatement of the reduced-problem construction inside the reference's
 * cut-pursuit.  Included by pfdr_oracle.c once per floating type.
 *
 * The reference does not expose these steps as functions: they are inline blocks of
 * CP_PFDR_graph_quadratic_d1_l1() (src/CP_PFDR_graph_quadratic_d1_l1.cpp).  Each function below restates one block,
 * with the block's own accumulation type (REAL) and member order (the component list Vc / rVc).  The restatement is
 * pinned through the only observable the reference offers - the components and reduced values its cut-pursuit returns
 * (stopped after a few outer iterations; tests/test_oracle_vs_ref.py::test_cp_reduce_chain, tests/test_golden.py):
 * reference Cv -> these functions -> oracle PFDR -> the reference's rX.
 *
 * Components: rV of them; Vc[rVc[c] .. rVc[c+1]) lists the vertices of component c (reference :574-600). */

/* gradient of the quadratic term at the current reduced iterate, before the d1 part is added (reference :340-374)
 *   N > 0: DfS = -A^T R                       (R = residual Y - A x of the previous outer iteration, N values)
 *   N < 0: DfS = (AtA) x - AtY, x piecewise constant: per column u, sum_c rX[c] * (sum of AtA[v,u], v in c), components with rX == 0 skipped
 *   N = 0: DfS = a_v x_v - y_v                (a = diagonal of A^T A, or 1 when A == NULL) */
void FN(oracle_cp_gradient)(int V, int N, int rV, const REAL *A, const REAL *YorR, const REAL *rX, const int *Cv,
                            const int *Vc, const int *rVc, REAL *DfS)
{
    if (N > 0) {
        for (int v = 0; v < V; v++) {
            const REAL *col = A + (size_t) N * v;
            REAL acc = 0;
            for (int n = 0; n < N; n++) acc += col[n] * YorR[n];
            DfS[v] = -acc;
        }
    } else if (N < 0) {
        for (int u = 0; u < V; u++) {
            const REAL *col = A + (size_t) V * u;
            REAL total = 0;
            for (int c = 0; c < rV; c++) {
                if (rX[c] == 0) continue;
                REAL inner = 0;
                for (int s = rVc[c]; s < rVc[c + 1]; s++) inner += col[Vc[s]];
                total += inner * rX[c];
            }
            DfS[u] = total - YorR[u];
        }
    } else {
        for (int v = 0; v < V; v++) DfS[v] = (A ? A[v] * rX[Cv[v]] : rX[Cv[v]]) - YorR[v];
    }
}

/* reduced operator (reference :664-772).
 *   N > 0: rA  (N x rV, column major) = per component, the sum of its columns of A            (:677-688)
 *          preAt: rAA (rV x rV) = rA^T rA - upper triangle computed, then mirrored             (:689-702, :759-771)
 *                 rY  (rV)      = rA^T Y                                                      (:703-711)
 *   N < 0: rY = per component sum of Y (= A^T Y);  rAA[ru][rv] = sum over u in ru, v in rv of AtA[v + V u],
 *          rv <= ru computed and mirrored                                                     (:713-744, :759-771)
 *   N = 0: rY as above;  rAA (rV) = per component sum of the diagonal A, or the component size (:745-757)
 * rA may be NULL when N <= 0, rAA / rY may be NULL when N > 0 and !preAt. */
void FN(oracle_cp_reduce)(int V, int N, int rV, const REAL *A, const REAL *Y, const int *Vc, const int *rVc, int preAt,
                          REAL *rA, REAL *rAA, REAL *rY)
{
    (void) V;
    if (N > 0) {
        memset(rA, 0, sizeof(REAL) * (size_t) N * rV);
        for (int c = 0; c < rV; c++) {
            REAL *out = rA + (size_t) N * c;
            for (int s = rVc[c]; s < rVc[c + 1]; s++) {
                const REAL *col = A + (size_t) N * Vc[s];
                for (int n = 0; n < N; n++) out[n] += col[n];
            }
        }
        if (!preAt) return;
        for (int cu = 0; cu < rV; cu++) {
            const REAL *a = rA + (size_t) N * cu;
            for (int cv = 0; cv <= cu; cv++) {
                const REAL *b = rA + (size_t) N * cv;
                REAL acc = 0;
                for (int n = 0; n < N; n++) acc += b[n] * a[n];
                rAA[(size_t) rV * cu + cv] = acc;
            }
        }
        for (int c = 0; c < rV; c++) {
            const REAL *a = rA + (size_t) N * c;
            REAL acc = 0;
            for (int n = 0; n < N; n++) acc += a[n] * Y[n];
            rY[c] = acc;
        }
    } else {
        for (int c = 0; c < rV; c++) {
            REAL acc = 0;
            for (int s = rVc[c]; s < rVc[c + 1]; s++) acc += Y[Vc[s]];
            rY[c] = acc;
        }
        if (N == 0) {
            for (int c = 0; c < rV; c++) {
                if (A) {
                    REAL acc = 0;
                    for (int s = rVc[c]; s < rVc[c + 1]; s++) acc += A[Vc[s]];
                    rAA[c] = acc;
                } else {
                    rAA[c] = (REAL) (rVc[c + 1] - rVc[c]);
                }
            }
            return;
        }
        for (int cu = 0; cu < rV; cu++)
            for (int cv = 0; cv <= cu; cv++) {
                REAL acc = 0;
                for (int s = rVc[cu]; s < rVc[cu + 1]; s++) {
                    const REAL *col = A + (size_t) V * Vc[s];
                    for (int t = rVc[cv]; t < rVc[cv + 1]; t++) acc += col[Vc[t]];
                }
                rAA[(size_t) rV * cu + cv] = acc;
            }
    }
    /* mirror: entry (cu, cv), cv > cu, takes the value computed at (cv, cu) */
    for (int cu = 0; cu + 1 < rV; cu++)
        for (int cv = cu + 1; cv < rV; cv++) rAA[(size_t) rV * cu + cv] = rAA[(size_t) rV * cv + cu];
}

/* residual of the direct matricial case after the reduced solve, R = Y - rA rX, components in ascending order (reference :889-903) */
void FN(oracle_cp_residual)(int N, int rV, const REAL *rA, const REAL *rX, const REAL *Y, REAL *R)
{
    for (int n = 0; n < N; n++) {
        REAL acc = 0;
        for (int c = 0; c < rV; c++) acc += rA[n + (size_t) N * c] * rX[c];
        R[n] = Y[n] - acc;
    }
}

/* Jacobi-equilibrated copy of the reduced operator handed to operator_norm_matrix, and the scale vector (reference :779-789, :801-818):
 *   preAt : d[c] = sqrt(rAA[c][c]);  S = rAA / (d d^T)   (rV x rV)
 *   !preAt: d[c] = ||rA[:, c]||;     S = rA / d          (N x rV)
 * The Lipschitz metric is then L[c] = d[c]^2 * ||S||^2-estimate (:831-835); the caller multiplies. */
void FN(oracle_cp_equilibrate)(int N, int rV, int preAt, const REAL *M, REAL *S, REAL *d)
{
    if (preAt) {
        for (int c = 0; c < rV; c++) d[c] = (REAL) sqrt(M[(size_t) c * (rV + 1)]);
        for (int cu = 0; cu < rV; cu++)
            for (int cv = 0; cv < rV; cv++) S[(size_t) rV * cu + cv] = M[(size_t) rV * cu + cv] / (d[cu] * d[cv]);
    } else {
        for (int c = 0; c < rV; c++) {
            REAL acc = 0;
            for (int n = 0; n < N; n++) { const REAL b = M[n + (size_t) N * c]; acc += b * b; }
            d[c] = (REAL) sqrt(acc);
            for (int n = 0; n < N; n++) S[n + (size_t) N * c] = M[n + (size_t) N * c] / d[c];
        }
    }
}

#ifndef ORACLE_CP_GRAPH_ONCE
#define ORACLE_CP_GRAPH_ONCE
/* connected components of the graph restricted to its NON-active edges (reference :570-600): vertices are visited in
 * ascending order, an unassigned vertex opens component number rV and a breadth-first sweep over non-active edges assigns
 * everything reachable - so components are numbered by their smallest vertex.  Vc / rVc receive the visiting order.
 * The reference walks its own adjacency lists (src/graph.cpp); the sweep below walks a CSR built from the edge list, which
 * changes the order inside Vc but neither Cv nor rVc.  Returns rV. */
int oracle_cp_components(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, int *Cv, int *Vc, int *rVc)
{
    int *deg = (int *) calloc((size_t) V + 1, sizeof(int));
    for (int e = 0; e < E; e++) if (!active[e]) { deg[Eu[e] + 1]++; deg[Ev[e] + 1]++; }
    for (int v = 0; v < V; v++) deg[v + 1] += deg[v];
    int *adj = (int *) malloc(sizeof(int) * (size_t) (deg[V] > 0 ? deg[V] : 1));
    int *fill = (int *) malloc(sizeof(int) * (size_t) V);
    for (int v = 0; v < V; v++) fill[v] = deg[v];
    for (int e = 0; e < E; e++) if (!active[e]) { adj[fill[Eu[e]]++] = Ev[e]; adj[fill[Ev[e]]++] = Eu[e]; }
    for (int v = 0; v < V; v++) Cv[v] = -1;
    int rV = 0, assigned = 0, head = 0;
    rVc[0] = 0;
    for (int u = 0; u < V; u++) {
        if (Cv[u] != -1) continue;
        Cv[u] = rV;
        Vc[assigned++] = u;
        while (head < assigned) {
            const int v = Vc[head++];
            for (int k = deg[v]; k < deg[v + 1]; k++) {
                const int w = adj[k];
                if (Cv[w] != -1) continue;
                Cv[w] = rV;
                Vc[assigned++] = w;
            }
        }
        rVc[++rV] = assigned;
    }
    free(deg); free(adj); free(fill);
    return rV;
}
#endif

/* reduced connectivity (reference :603-661), in canonical form: one reduced edge per unordered pair of components joined
 * by at least one ACTIVE edge of nonzero weight, weights added up; an active edge whose two ends lie in the same component
 * yields the loop (c, c) and is met from both of its ends, so it counts twice (:622-640 walks every member's arcs);
 * components touched by no such edge get the loop (c, c) of weight eps (:645-648).
 * Output sorted by (smaller component, larger component) - the reference emits first-encounter order, which only matters
 * for the rounding of sums downstream.  acc: double so that the comparison is about the rule, not summation order.
 * rEu / rEv / rLa need room for E + rV entries.  Returns rE. */
int FN(oracle_cp_reduced_edges)(int V, int E, const int *Eu, const int *Ev, const unsigned char *active, const REAL *La_d1,
                                const int *Cv, int rV, REAL eps, int *rEu, int *rEv, REAL *rLa)
{
    (void) V;
    /* bucket the active nonzero edges by smaller component, then sort each bucket by larger component (insertion into a
     * per-bucket list is enough at test sizes) */
    int *cnt = (int *) calloc((size_t) rV + 1, sizeof(int));
    for (int e = 0; e < E; e++) if (active[e] && La_d1[e] != 0) { const int a = Cv[Eu[e]], b = Cv[Ev[e]]; cnt[(a < b ? a : b) + 1]++; }
    for (int c = 0; c < rV; c++) cnt[c + 1] += cnt[c];
    const int m = cnt[rV];
    int *hi = (int *) malloc(sizeof(int) * (size_t) (m > 0 ? m : 1));
    double *w = (double *) malloc(sizeof(double) * (size_t) (m > 0 ? m : 1));
    int *fill = (int *) malloc(sizeof(int) * (size_t) rV);
    for (int c = 0; c < rV; c++) fill[c] = cnt[c];
    unsigned char *touched = (unsigned char *) calloc((size_t) rV, 1);
    for (int e = 0; e < E; e++) if (active[e] && La_d1[e] != 0) {
        const int a = Cv[Eu[e]], b = Cv[Ev[e]], lo = a < b ? a : b, up = a < b ? b : a;
        hi[fill[lo]] = up; w[fill[lo]] = (a == b ? 2.0 : 1.0) * (double) La_d1[e]; fill[lo]++;
        touched[a] = touched[b] = 1;
    }
    int rE = 0;
    double *sum = (double *) calloc((size_t) rV, sizeof(double));
    unsigned char *seen = (unsigned char *) calloc((size_t) rV, 1);
    for (int c = 0; c < rV; c++) {
        for (int k = cnt[c]; k < cnt[c + 1]; k++) { sum[hi[k]] += w[k]; seen[hi[k]] = 1; }
        if (!touched[c]) { rEu[rE] = c; rEv[rE] = c; rLa[rE++] = eps; continue; }
        for (int d = c; d < rV; d++) if (seen[d]) { rEu[rE] = c; rEv[rE] = d; rLa[rE++] = (REAL) sum[d]; seen[d] = 0; sum[d] = 0.0; }
    }
    free(cnt); free(hi); free(w); free(fill); free(touched); free(sum); free(seen);
    return rE;
}

/* merge step (reference :863-886): an active edge whose two components took (relatively) equal values is deactivated:
 * d = |a - b| / max(|a|, |b|, eps) <= CP_difTol.  Returns the number of edges deactivated. */
int FN(oracle_cp_merge)(int E, const int *Eu, const int *Ev, const int *Cv, const REAL *rX, REAL CP_difTol, REAL eps, unsigned char *active)
{
    int n = 0;
    for (int e = 0; e < E; e++) {
        if (!active[e]) continue;
        REAL a = rX[Cv[Eu[e]]], b = rX[Cv[Ev[e]]], d = a - b;
        if (a < 0) a = -a;
        if (b < 0) b = -b;
        if (d < 0) d = -d;
        if (a < b) a = b;
        d = (a > eps) ? d / a : d / eps;
        if (d <= CP_difTol) { active[e] = 0; n++; }
    }
    return n;
}
