/* TEST INFRASTRUCTURE ONLY - serial CPU restatement of the reduced-problem construction inside the reference's
 * cut-pursuit.  Included by pfdr_oracle.c once per floating type.
 *
 * The reference does not expose these steps as functions: they are inline blocks of
 * CP_PFDR_graph_quadratic_d1_l1() (src/CP_PFDR_graph_quadratic_d1_l1.cpp).  Each function below restates one block,
 * with the block's own accumulation type (REAL) and member order (the component list Vc / rVc).  The restatement is
 * pinned through the only observable the reference offers - the reduced values its cut-pursuit returns after ONE outer
 * iteration (tests/test_oracle_vs_ref.py::test_cp_reduce_chain_*): reference Cv -> these functions -> oracle PFDR ->
 * the reference's rX.
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
