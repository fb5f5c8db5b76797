// cpreduce.cu - the dense passes of the cut-pursuit's outer iteration on the device:
// the operator A (N x V), A^tA (V x V) or its diagonal stays resident across outer iterations; per iteration
//   gradient  DfS = -A^t R | (A^tA) x - A^tY | a x - y       reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:340-374
//   reduce    rA = A S, rAA = S^t (A^tA) S, rY = S^t A^tY    reference :664-772      (S = V x rV indicator of the components)
//   residual  R = Y - rA rX                                  reference :889-903
// (the bounds twin has the same blocks: src/CP_PFDR_graph_quadratic_d1_bounds.cpp:334-366, :640-748, :866-880).
//
// All three are single passes over the resident operator, bound by HBM: the operator is read ONCE per call.  Sums over a
// component are taken in a fixed order (members by ascending vertex number, split into pieces of 64 members whose partial
// sums are added in piece order), in double whatever the storage type - results do not depend on the launch shape and are
// closer to the exact sums than the reference's running sums in `real`; tests compare against the oracle restatement
// of the reference's blocks within 1e-10 (float64) / 1e-3 (float32).
#include "solvers.cuh"
#include <cub/cub.cuh>

namespace pfdr {
namespace {

constexpr int PIECE = 64;                  // component members per piece
constexpr int LOADS = 8;                   // column loads a thread has in flight

// ---------------------------------------------------------------------------------------------
// component lists
// ---------------------------------------------------------------------------------------------
__global__ void cp_iota_check_kernel(const int *__restrict__ Cv, int V, int rV, int *__restrict__ idx, int *__restrict__ bad)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    idx[v] = v;
    if (Cv[v] < 0 || Cv[v] >= rV) atomicOr(bad, 1);
}

// rVc[c] = first position of the sorted assignment holding a value >= c  (c = 0 .. rV)
__global__ void cp_bounds_kernel(const int *__restrict__ sorted, int V, int rV, int *__restrict__ rVc, int *__restrict__ npieces)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > rV) return;
    auto lower = [&](int key) { int lo = 0, hi = V; while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; } return lo; };
    const int b = lower(c);
    rVc[c] = b;
    if (c < rV) { const int n = lower(c + 1) - b; npieces[c] = n > 0 ? (n + PIECE - 1) / PIECE : 1; }
    else npieces[c] = 0;
}

// piece p -> component c and the positions [s0, s1) of its members in Vc
__device__ __forceinline__ void piece_range(int p, const int *__restrict__ first, const int *__restrict__ rVc, int rV, int &c, int &s0, int &s1)
{
    int lo = 0, hi = rV;                       // first[lo] <= p < first[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (first[mid] <= p) lo = mid; else hi = mid; }
    c = lo;
    s0 = rVc[c] + (p - first[c]) * PIECE;
    s1 = min(s0 + PIECE, rVc[c + 1]);
}

// ---------------------------------------------------------------------------------------------
// Pp[p][r] = sum over the members s of piece p of M[r + ld * Vc[s]]        (columns of a column-major matrix)
// A thread owns VEC consecutive rows and adds the (up to 64) columns of the piece in member order, eight vector loads in
// flight at a time; a CTA of W warps covers W * 32 * VEC consecutive rows, i.e. reads W * 512 B (float) of every column.
// ---------------------------------------------------------------------------------------------
template <typename real, int VEC> struct RowVec;
template <> struct RowVec<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct RowVec<double, 4> {
    static __device__ __forceinline__ void load(const double *p, double (&v)[4])
    {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <typename real> struct RowVec<real, 1> {
    static __device__ __forceinline__ void load(const real *p, real (&v)[1]) { v[0] = __ldg(p); }
};

template <typename real, int VEC>
__global__ void __launch_bounds__(256) seg_colsum_kernel(const real *__restrict__ M, size_t ld, int rows, const int *__restrict__ Vc,
                                                        const int *__restrict__ rVc, const int *__restrict__ first, int rV,
                                                        double *__restrict__ Pp)
{
    const int p = blockIdx.x;
    if (p >= first[rV]) return;
    __shared__ int s_rng[2];
    __shared__ int s_idx[PIECE];
    if (threadIdx.x == 0) { int c, s0, s1; piece_range(p, first, rVc, rV, c, s0, s1); s_rng[0] = s0; s_rng[1] = s1 - s0; }
    __syncthreads();
    const int n = s_rng[1];
    for (int j = threadIdx.x; j < n; j += blockDim.x) s_idx[j] = Vc[s_rng[0] + j];
    __syncthreads();
    const int r0 = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
    if (r0 >= rows) return;
    double acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i++) acc[i] = 0.0;
    for (int j0 = 0; j0 < n; j0 += LOADS) {
        real v[LOADS][VEC];
#pragma unroll
        for (int j = 0; j < LOADS; j++) {
            if (j0 + j < n) RowVec<real, VEC>::load(M + ld * (size_t) s_idx[j0 + j] + r0, v[j]);
            else {
#pragma unroll
                for (int i = 0; i < VEC; i++) v[j][i] = (real) 0;
            }
        }
#pragma unroll
        for (int j = 0; j < LOADS; j++)
#pragma unroll
            for (int i = 0; i < VEC; i++) acc[i] += (double) v[j][i];
    }
#pragma unroll
    for (int i = 0; i < VEC; i++) Pp[(size_t) p * rows + r0 + i] = acc[i];
}

// out[c][r] = sum of the pieces of component c, in piece order; stored in `real`
template <typename real>
__global__ void piece_combine_kernel(const double *__restrict__ Pp, const int *__restrict__ first, int rV, int rows, real *__restrict__ out)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) rV * rows) return;
    const int c = (int) (i / rows), r = (int) (i % rows);
    double t = 0.0;
    for (int p = first[c]; p < first[c + 1]; p++) t += Pp[(size_t) p * rows + r];
    out[i] = (real) t;
}

// component sizes as the reduced diagonal of the identity (reference :753-755)
template <typename real>
__global__ void comp_size_kernel(const int *__restrict__ rVc, int rV, real *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < rV) out[c] = (real) (rVc[c + 1] - rVc[c]);
}

// ---------------------------------------------------------------------------------------------
// N > 0: rAA = rA^t rA (lower triangle computed, mirrored), rY = rA^t Y
// ---------------------------------------------------------------------------------------------
// 64 x 64 tile of rAA per CTA, 4 x 4 per thread, 16 rows of rA staged per step as [row][column] so that a thread's four
// operands of either side are 32 contiguous bytes; products and sums in double, n ascending
template <typename real>
__global__ void __launch_bounds__(256) syrk_kernel(const real *__restrict__ rA, int N, int rV, real *__restrict__ rAA)
{
    const int bu = blockIdx.y, bv = blockIdx.x;
    if (bv > bu) return;
    constexpr int T = 64, KC = 16, LD = T + 2;
    __shared__ __align__(16) double a[KC][LD], b[KC][LD];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    for (int n0 = 0; n0 < N; n0 += KC) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = threadIdx.x + 256 * q, k = e & (KC - 1), col = e >> 4, n = n0 + k;
            const int cu = bu * T + col, cv = bv * T + col;
            a[k][col] = (cu < rV && n < N) ? (double) rA[(size_t) N * cu + n] : 0.0;
            b[k][col] = (cv < rV && n < N) ? (double) rA[(size_t) N * cv + n] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; k++) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&a[k][ty * 4]), a23 = *reinterpret_cast<const double2 *>(&a[k][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2 *>(&b[k][tx * 4]), b23 = *reinterpret_cast<const double2 *>(&b[k][tx * 4 + 2]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bw[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += av[i] * bw[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int cu = bu * T + ty * 4 + i, cv = bv * T + tx * 4 + j;
            if (cu < rV && cv <= cu) {
                const real t = (real) acc[i][j];
                rAA[(size_t) rV * cu + cv] = t;
                rAA[(size_t) rV * cv + cu] = t;
            }
        }
}

// out[c] = sgn * <M[:, c], x>  (- y[c]);  G threads per column, fixed reduction tree
template <typename real, int VEC, int G>
__global__ void __launch_bounds__(256) col_dot_kernel(const real *__restrict__ M, size_t ld, int rows, int cols, const real *__restrict__ x,
                                                     const real *__restrict__ y, int negate, real *__restrict__ out)
{
    __shared__ double sm[32];
    const int per_cta = 256 / G;
    const int c = blockIdx.x * per_cta + threadIdx.x / G;
    const int t = threadIdx.x % G;
    double acc = 0.0;
    if (c < cols) {
        const real *col = M + ld * (size_t) c;
        for (int r = t * VEC; r < rows; r += G * VEC) {
            real a[VEC], b[VEC];
            RowVec<real, VEC>::load(col + r, a);
            RowVec<real, VEC>::load(x + r, b);
#pragma unroll
            for (int i = 0; i < VEC; i++) acc += (double) a[i] * (double) b[i];
        }
    }
    if (G == 32) {
        acc = warp_sum(acc);
    } else {
        double v[1] = {acc};
        block_sum<1>(v, sm);
        acc = v[0];
    }
    if (t == 0 && c < cols) {
        double r = negate ? -acc : acc;
        if (y) r -= (double) y[c];
        out[c] = (real) r;
    }
}

template <typename real>
__global__ void expand_kernel(const real *__restrict__ rX, const int *__restrict__ Cv, int V, real *__restrict__ x)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) x[v] = rX[Cv[v]];
}

template <typename real>
__global__ void diag_gradient_kernel(const real *__restrict__ A, const real *__restrict__ rX, const int *__restrict__ Cv, const real *__restrict__ Y,
                                     int V, real *__restrict__ out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const real x = rX[Cv[v]];
    out[v] = (A ? A[v] * x : x) - Y[v];
}

template <typename real>
__global__ void residual_kernel(const real *__restrict__ rA, const real *__restrict__ rX, const real *__restrict__ Y, int N, int rV, real *__restrict__ R)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double acc = 0.0;
    for (int c = 0; c < rV; c++) acc += (double) rA[n + (size_t) N * c] * (double) rX[c];
    R[n] = (real) ((double) Y[n] - acc);
}

// ---------------------------------------------------------------------------------------------
// N < 0, second and third stage: U[p][c] = sum over the members s of component c of Pp[p][Vc[s]] (one warp per (p, c)),
// rAA[cu][cv] = sum over the pieces p of cu of U[p][cv], cv <= cu, mirrored
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_rowsum_kernel(const double *__restrict__ Pp, int V, const int *__restrict__ Vc, const int *__restrict__ rVc,
                                                        const int *__restrict__ first, int rV, double *__restrict__ U)
{
    const int p = blockIdx.x;
    if (p >= first[rV]) return;
    const int c = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= rV) return;
    const double *row = Pp + (size_t) p * V;
    double acc = 0.0;
    for (int s = rVc[c] + lane; s < rVc[c + 1]; s += 32) acc += row[Vc[s]];
    acc = warp_sum(acc);
    if (lane == 0) U[(size_t) p * rV + c] = acc;
}

template <typename real>
__global__ void piece_fold_kernel(const double *__restrict__ U, const int *__restrict__ first, int rV, real *__restrict__ rAA)
{
    const int cv = blockIdx.x * blockDim.x + threadIdx.x, cu = blockIdx.y;
    if (cv > cu) return;
    double t = 0.0;
    for (int p = first[cu]; p < first[cu + 1]; p++) t += U[(size_t) p * rV + cv];
    const real r = (real) t;
    rAA[(size_t) rV * cu + cv] = r;
    rAA[(size_t) rV * cv + cu] = r;
}

template <typename real>
bool vec4_ok(const real *p, size_t ld, int rows) { return rows % 4 == 0 && ld % 4 == 0 && ((uintptr_t) p % (4 * sizeof(real))) == 0; }

}  // namespace

// ---------------------------------------------------------------------------------------------
template <typename real>
struct CpOperator : CpOperatorBase {
    int V, N;
    DevIn<real> A;
    DevBuf<int> Cv, Vc, rVc, first;
    int rV = 0, pieces_max = 0;
    DevBuf<real> rA, R;                 // N > 0: reduced operator of the last reduce(), residual of the last residual()
    int rA_rV = 0; bool have_R = false;

    CpOperator(int V_, int N_, const real *A_) : V(V_), N(N_)
    {
        PFDR_REQUIRE(V >= 1, "cp_operator: V must be positive");
        PFDR_REQUIRE(N >= 0 || N == -V, "cp_operator: N < 0 announces the V x V matrix A^tA and must equal -V");
        PFDR_REQUIRE(N == 0 || A_ != nullptr, "cp_operator: A must not be NULL unless N == 0 (identity)");
        const size_t count = N > 0 ? (size_t) N * V : (N < 0 ? (size_t) V * V : (size_t) V);
        A.bind(A_, A_ ? count : 0);
    }

    void set_components(int rV_, const int *Cv_host) override
    {
        PFDR_REQUIRE(rV_ >= 1 && rV_ <= V && Cv_host != nullptr, "cp_set_components: need 1 <= rV <= V and the assignment Cv");
        cudaStream_t st = ctx().stream;
        rV = rV_;
        DevIn<int> in; in.bind(Cv_host, V);
        Cv.alloc(V); Vc.alloc(V); rVc.alloc(rV + 1); first.alloc(rV + 1);
        PFDR_CUDA(cudaMemcpyAsync(Cv.p, in.p, sizeof(int) * V, cudaMemcpyDeviceToDevice, st));
        DevBuf<int> idx(V), sorted(V), np(rV + 1), bad(1);
        bad.zero();
        PFDR_LAUNCH(cp_iota_check_kernel, cdiv(V, 256), 256, 0, Cv.p, V, rV, idx.p, bad.p);
        int bits = 1; while (((long long) 1 << bits) < (long long) rV) bits++;
        size_t t1 = 0, t2 = 0;
        PFDR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, Cv.p, sorted.p, idx.p, Vc.p, V, 0, bits, st));
        PFDR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, np.p, first.p, rV + 1, st));
        DevBuf<unsigned char> tmp(std::max(t1, t2) + 16);
        size_t tb = tmp.n;
        PFDR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, Cv.p, sorted.p, idx.p, Vc.p, V, 0, bits, st));   // stable: ascending vertex inside a component
        PFDR_LAUNCH(cp_bounds_kernel, cdiv(rV + 1, 256), 256, 0, sorted.p, V, rV, rVc.p, np.p);
        tb = tmp.n;
        PFDR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, np.p, first.p, rV + 1, st));
        ctx().stats.launches += 2;
        int hbad = 0;
        PFDR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
        if (hbad) { rV = 0; throw Error(PFDR_B200_ERR_ARG, "cp_set_components: Cv holds a value outside [0, rV)"); }
        pieces_max = V / PIECE + rV;               // >= sum over components of max(1, ceil(size / PIECE))
        rA_rV = 0; have_R = false;
    }

    // Pp[p][r] over the pieces of the current components
    void colsum(const real *M, size_t ld, int rows, double *Pp)
    {
        const bool v4 = vec4_ok(M, ld, rows);
        const int per_warp = v4 ? 128 : 32;
        const int warps = (int) std::min<size_t>(8, cdiv(rows, per_warp));
        dim3 grid(pieces_max, cdiv(rows, (size_t) per_warp * warps));
        if (v4) { auto kern = seg_colsum_kernel<real, 4>; PFDR_LAUNCH(kern, grid, 32 * warps, 0, M, ld, rows, Vc.p, rVc.p, first.p, rV, Pp); }
        else { auto kern = seg_colsum_kernel<real, 1>; PFDR_LAUNCH(kern, grid, 32 * warps, 0, M, ld, rows, Vc.p, rVc.p, first.p, rV, Pp); }
    }

    // out[c] = sum of vec over component c (vec indexed by vertex)
    void vecsum(const real *vec, real *out)
    {
        DevBuf<double> Pp((size_t) pieces_max);
        colsum(vec, 1, 1, Pp.p);
        PFDR_LAUNCH(piece_combine_kernel<real>, cdiv(rV, 256), 256, 0, Pp.p, first.p, rV, 1, out);
    }

    void reduce(const void *Y_, int preAt, void *rA_out, void *rAA_out, void *rY_out) override
    {
        PFDR_REQUIRE(rV > 0, "cp_reduce: call cp_set_components first");
        const real *Yh = static_cast<const real *>(Y_);
        const bool pre = preAt != 0 || N <= 0;
        PFDR_REQUIRE(!pre || (rAA_out && rY_out && Yh), "cp_reduce: rAA, rY and Y are required (N <= 0 or preAt)");
        DevIn<real> Y; Y.bind(Yh, pre ? (size_t) (N > 0 ? N : V) : 0);
        DevBuf<real> rY(pre ? rV : 0);
        if (N > 0) {
            rA.alloc((size_t) N * rV);
            {
                DevBuf<double> Pp((size_t) pieces_max * N);
                colsum(A.p, (size_t) N, N, Pp.p);
                PFDR_LAUNCH(piece_combine_kernel<real>, cdiv((size_t) rV * N, 256), 256, 0, Pp.p, first.p, rV, N, rA.p);
            }
            rA_rV = rV;
            copy_out(static_cast<real *>(rA_out), rA.p, (size_t) N * rV);
            if (pre) {
                DevBuf<real> rAA((size_t) rV * rV);
                dim3 grid(cdiv(rV, 64), cdiv(rV, 64));
                PFDR_LAUNCH(syrk_kernel<real>, grid, 256, 0, rA.p, N, rV, rAA.p);
                if (vec4_ok(rA.p, (size_t) N, N) && vec4_ok(Y.p, 4, N))
                    { auto kern = col_dot_kernel<real, 4, 32>; PFDR_LAUNCH(kern, cdiv(rV, 8), 256, 0, rA.p, (size_t) N, N, rV, Y.p, (const real *) nullptr, 0, rY.p); }
                else
                    { auto kern = col_dot_kernel<real, 1, 32>; PFDR_LAUNCH(kern, cdiv(rV, 8), 256, 0, rA.p, (size_t) N, N, rV, Y.p, (const real *) nullptr, 0, rY.p); }
                copy_out(static_cast<real *>(rAA_out), rAA.p, (size_t) rV * rV);
                copy_out(static_cast<real *>(rY_out), rY.p, (size_t) rV);
                PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
            }
        } else {
            vecsum(Y.p, rY.p);
            copy_out(static_cast<real *>(rY_out), rY.p, (size_t) rV);
            if (N == 0) {
                DevBuf<real> rAA(rV);
                if (A.p) vecsum(A.p, rAA.p);
                else PFDR_LAUNCH(comp_size_kernel<real>, cdiv(rV, 256), 256, 0, rVc.p, rV, rAA.p);
                copy_out(static_cast<real *>(rAA_out), rAA.p, (size_t) rV);
                PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
            } else {
                DevBuf<real> rAA((size_t) rV * rV);
                DevBuf<double> Pp((size_t) pieces_max * V), U((size_t) pieces_max * rV);
                colsum(A.p, (size_t) V, V, Pp.p);
                dim3 g2(pieces_max, cdiv(rV, 8));
                PFDR_LAUNCH(seg_rowsum_kernel, g2, 256, 0, Pp.p, V, Vc.p, rVc.p, first.p, rV, U.p);
                dim3 g3(cdiv(rV, 128), rV);
                PFDR_LAUNCH(piece_fold_kernel<real>, g3, 128, 0, U.p, first.p, rV, rAA.p);
                copy_out(static_cast<real *>(rAA_out), rAA.p, (size_t) rV * rV);
                PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
            }
        }
        PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
    }

    void gradient(const void *rX_, const void *YorR_, void *DfS_out) override
    {
        PFDR_REQUIRE(rV > 0, "cp_gradient: call cp_set_components first");
        PFDR_REQUIRE(rX_ && DfS_out, "cp_gradient: rX and DfS must not be NULL");
        DevIn<real> rX; rX.bind(static_cast<const real *>(rX_), rV);
        DevBuf<real> out(V);
        if (N > 0) {
            DevIn<real> Rin;
            const real *Rp;
            if (YorR_) { Rin.bind(static_cast<const real *>(YorR_), N); Rp = Rin.p; }
            else { PFDR_REQUIRE(have_R, "cp_gradient: no residual given and none kept from cp_residual"); Rp = R.p; }
            launch_col_dot(A.p, (size_t) N, N, V, Rp, nullptr, 1, out.p);
        } else {
            PFDR_REQUIRE(YorR_ != nullptr, "cp_gradient: the observation A^tY must not be NULL");
            DevIn<real> Y; Y.bind(static_cast<const real *>(YorR_), V);
            if (N < 0) {
                DevBuf<real> x(V);
                PFDR_LAUNCH(expand_kernel<real>, cdiv(V, 256), 256, 0, rX.p, Cv.p, V, x.p);
                launch_col_dot(A.p, (size_t) V, V, V, x.p, Y.p, 0, out.p);
            } else {
                PFDR_LAUNCH(diag_gradient_kernel<real>, cdiv(V, 256), 256, 0, A.p, rX.p, Cv.p, Y.p, V, out.p);
            }
            copy_out(static_cast<real *>(DfS_out), out.p, (size_t) V);
            PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
            return;
        }
        copy_out(static_cast<real *>(DfS_out), out.p, (size_t) V);
        PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
    }

    void launch_col_dot(const real *M, size_t ld, int rows, int cols, const real *x, const real *y, int negate, real *out)
    {
        const bool v4 = vec4_ok(M, ld, rows) && vec4_ok(x, 4, rows);
        if (rows >= 2048) {
            if (v4) { auto kern = col_dot_kernel<real, 4, 256>; PFDR_LAUNCH(kern, cols, 256, 0, M, ld, rows, cols, x, y, negate, out); }
            else { auto kern = col_dot_kernel<real, 1, 256>; PFDR_LAUNCH(kern, cols, 256, 0, M, ld, rows, cols, x, y, negate, out); }
        } else {
            if (v4) { auto kern = col_dot_kernel<real, 4, 32>; PFDR_LAUNCH(kern, cdiv(cols, 8), 256, 0, M, ld, rows, cols, x, y, negate, out); }
            else { auto kern = col_dot_kernel<real, 1, 32>; PFDR_LAUNCH(kern, cdiv(cols, 8), 256, 0, M, ld, rows, cols, x, y, negate, out); }
        }
    }

    void residual(const void *rX_, const void *Y_, void *R_out) override
    {
        PFDR_REQUIRE(N > 0, "cp_residual: only the direct matricial case (N > 0) has a residual");
        PFDR_REQUIRE(rA_rV == rV && rV > 0, "cp_residual: call cp_reduce for the current components first");
        PFDR_REQUIRE(rX_ && Y_, "cp_residual: rX and Y must not be NULL");
        DevIn<real> rX, Y;
        rX.bind(static_cast<const real *>(rX_), rV);
        Y.bind(static_cast<const real *>(Y_), N);
        R.alloc(N);
        PFDR_LAUNCH(residual_kernel<real>, cdiv(N, 128), 128, 0, rA.p, rX.p, Y.p, N, rV, R.p);
        have_R = true;
        copy_out(static_cast<real *>(R_out), R.p, (size_t) N);
        PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
};

template <typename real> CpOperatorBase *make_cp_operator(int V, int N, const real *A) { return new CpOperator<real>(V, N, A); }
template CpOperatorBase *make_cp_operator<float>(int, int, const float *);
template CpOperatorBase *make_cp_operator<double>(int, int, const double *);

}  // namespace pfdr
