// opnorm.cu - squared operator norm of a matrix by the power method, the B200-native replacement of
// src/operator_norm_matrix.cpp:84-212 of the reference.
//
// Same estimator as the reference, start for start: nbInit is rounded up to a multiple of the host's core count
// (:171-172), host thread t draws its starts from glibc's rand_r seeded with time(NULL) + t (:182-191, static
// schedule), every start iterates x <- S (x / |x|) until the relative change of the norm falls below nTol (:192-203)
// and the largest norm wins (:204).  The starts are generated on the host exactly as the reference's threads would
// (the draws are a few thousand integers) and uploaded, so that cut-pursuit runs on this library take the same
// Lipschitz metric - hence the same trajectory and the same components - as on the reference.  The power
// iterations themselves run on the device:
//   * the matrix is symmetrised (A^tA / AA^t, whichever side the reference iterates on, :111-168) once;
//   * small systems (the reduced problems of cut-pursuit): one CTA per start, the whole iteration in one launch;
//   * large systems: all starts advance together, S streamed ONCE per iteration for all of them (a thin
//     matrix-matrix product: lanes along the contiguous index of S, one accumulator per start), the work spread
//     over every SM; the per-start stopping rule is evaluated on the device, no host round trip per iteration.
// Tensor cores: the only dense contraction of the hot path, but thin (<= 32 right-hand sides), HBM-bound at that
// width in float32 and needed in float64 for parity - plain FMA.
// PFDR_B200_OPNORM_TIME=<t> replaces time(NULL) (reproducible runs; the tests pin the reference's clock the same way).
#include "solvers.cuh"
#include <sched.h>
#include <time.h>
#include <unistd.h>

namespace pfdr {

namespace {

// S = B B^t with B (n x L):  rows_mode: B[r,l] = A[r + M l]  (A A^t, n = M, L = N)
//                            else     : B[r,l] = A[l + M r]  (A^t A, n = N, L = M)
template <typename real>
__global__ void __launch_bounds__(256) gram_kernel(const real *__restrict__ A, int M, int n, int L, int rows_mode,
                                                   real *__restrict__ S)
{
    __shared__ real Bp[16][17], Bq[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int p0 = blockIdx.y * 16, q0 = blockIdx.x * 16;
    if (q0 > p0) return;                                  // lower triangle, mirrored below
    real acc = 0;
    for (int l0 = 0; l0 < L; l0 += 16) {
        {   // stage the two 16x16 panels; in rows_mode r is the contiguous index of A, else l is
            const int a = rows_mode ? tx : ty, b = rows_mode ? ty : tx;   // a: row offset, b: l offset
            const int l = l0 + b;
            const int rp = p0 + a, rq = q0 + a;
            Bp[a][b] = (rp < n && l < L) ? (rows_mode ? A[(size_t) rp + (size_t) M * l] : A[(size_t) l + (size_t) M * rp]) : (real) 0;
            Bq[a][b] = (rq < n && l < L) ? (rows_mode ? A[(size_t) rq + (size_t) M * l] : A[(size_t) l + (size_t) M * rq]) : (real) 0;
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 16; l++) acc += Bp[ty][l] * Bq[tx][l];
        __syncthreads();
    }
    const int pr = p0 + ty, qc = q0 + tx;
    if (pr < n && qc < n) {
        S[(size_t) pr + (size_t) n * qc] = acc;
        S[(size_t) qc + (size_t) n * pr] = acc;
    }
}

template <typename real>
__device__ real block_norm(const real *x, int n, double *sm)
{
    double v[1] = {0.0};
    for (int j = threadIdx.x; j < n; j += blockDim.x) v[0] += (double) x[j] * (double) x[j];
    block_sum<1>(v, sm);
    return (real) sqrt(v[0]);
}

// x <- S (x / b)   (ref normalize_and_apply_matrix :52-81, symmetrised branch); warp per column of S
template <typename real>
__device__ void apply_sym(const real *__restrict__ S, int n, real *x, real *t, real b)
{
    for (int j = threadIdx.x; j < n; j += blockDim.x) t[j] = x[j] / b;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = wid; j < n; j += nw) {
        const real *col = S + (size_t) n * j;
        real a = 0;
        for (int i = lane; i < n; i += 32) a += col[i] * t[i];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) x[j] = a;
    }
    __syncthreads();
}

// small systems: one CTA per start, x holds the start on entry
template <typename real>
__global__ void __launch_bounds__(256) power_kernel(const real *__restrict__ S, int n, real nTol, int itMax,
                                                    real *xbuf, real *tbuf, real *out)
{
    __shared__ double sm[32];
    real *x = xbuf + (size_t) blockIdx.x * n, *t = tbuf + (size_t) blockIdx.x * n;
    real b = block_norm(x, n, sm);
    apply_sym(S, n, x, t, b);
    b = block_norm(x, n, sm);
    if (b > (real) 0) {
        for (int it = 0; it < itMax; it++) {                                                            // ref :196-203
            apply_sym(S, n, x, t, b);
            const real a = block_norm(x, n, sm);
            if ((a - b) / b < nTol) break;
            b = a;
        }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = b;
}

// ---- large systems: all starts together ---------------------------------------------------------------------
// state of start s: b[s] current norm, phase[s]: 0 = before the first product, 1 = iterating, 2 = stopped; trip[s]
constexpr int NBMAX = 32;             // starts per batch
constexpr int ITILE = 64;             // rows of x staged per step

// part[seg][s][j] = sum over the segment's i of S[j + n i] . x[s][i] / b[s]     (S symmetric: row access is coalesced)
template <typename real, int NB>
__global__ void __launch_bounds__(256) symm_batched_kernel(const real *__restrict__ S, int n, int nb, int nseg,
                                                           const real *__restrict__ x, const real *__restrict__ bnorm,
                                                           const int *__restrict__ nactive, real *__restrict__ part)
{
    __shared__ __align__(16) real xs[ITILE][NB];
    if (*nactive == 0) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y;
    const int per = (n + nseg - 1) / nseg, i0 = seg * per, i1 = min(n, i0 + per);
    real acc[NB];
#pragma unroll
    for (int s = 0; s < NB; s++) acc[s] = 0;
    for (int it0 = i0; it0 < i1; it0 += ITILE) {
        __syncthreads();
        for (int q = threadIdx.x; q < ITILE * NB; q += blockDim.x) {
            const int ii = q / NB, s = q % NB, i = it0 + ii;
            xs[ii][s] = (s < nb && i < i1) ? x[(size_t) s * n + i] / bnorm[s] : (real) 0;
        }
        __syncthreads();
        if (j < n) {
            const int lim = min(ITILE, i1 - it0);
#pragma unroll 4
            for (int ii = 0; ii < lim; ii++) {
                const real sv = S[(size_t) j + (size_t) n * (it0 + ii)];
#pragma unroll
                for (int s = 0; s < NB; s++) acc[s] += sv * xs[ii][s];
            }
        }
    }
    if (j < n) {
#pragma unroll
        for (int s = 0; s < NB; s++) if (s < nb) part[((size_t) seg * nb + s) * n + j] = acc[s];
    }
}

// one CTA per start: x[s] <- sum of the segment partials, its norm, the reference's stopping rule (:192-203)
template <typename real>
__global__ void __launch_bounds__(256) power_step_kernel(int n, int nb, int nseg, const real *__restrict__ part, real *x,
                                                         real *bnorm, int *phase, int *trips, real nTol, int itMax, int *nactive)
{
    __shared__ double sm[32];
    const int s = blockIdx.x;
    if (phase[s] == 2) return;
    double v[1] = {0.0};
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        real y = 0;
        for (int g = 0; g < nseg; g++) y += part[((size_t) g * nb + s) * n + j];
        x[(size_t) s * n + j] = y;
        v[0] += (double) y * (double) y;
    }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) {
        const real a = (real) sqrt(v[0]), b = bnorm[s];
        int ph = phase[s];
        if (ph == 0) { bnorm[s] = a; ph = (a > (real) 0 && itMax > 0) ? 1 : 2; }       // the product before the loop (:193-194)
        else {
            if ((a - b) / b < nTol) ph = 2;                                           // b keeps the previous norm (:200)
            else { bnorm[s] = a; if (++trips[s] >= itMax) ph = 2; }
        }
        phase[s] = ph;
        if (ph == 2) atomicSub(nactive, 1);
    }
}

template <typename real>
__global__ void norm_init_kernel(int n, const real *__restrict__ x, real *bnorm)
{
    __shared__ double sm[32];
    const real b = block_norm(x + (size_t) blockIdx.x * n, n, sm);
    if (threadIdx.x == 0) bnorm[blockIdx.x] = b;
}

// glibc's rand_r (stdlib/rand_r.c): the generator behind the reference's starts
inline int glibc_rand_r(unsigned int *seed)
{
    unsigned int next = *seed;
    int result;
    next *= 1103515245u; next += 12345u;
    result = (int) ((next / 65536u) % 2048u);
    next *= 1103515245u; next += 12345u;
    result <<= 10; result ^= (int) ((next / 65536u) % 1024u);
    next *= 1103515245u; next += 12345u;
    result <<= 10; result ^= (int) ((next / 65536u) % 1024u);
    *seed = next;
    return result;
}

inline int host_cores()
{   // omp_get_num_procs(): the CPUs this process may run on
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) return c; }
    const long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int) c : 1;
}

}  // namespace

template <typename real>
real operator_norm_matrix(int M, int N, const real *A, real nTol, int itMax, int nbInit, int verbose)
{
    PFDR_REQUIRE(M >= 0 && N >= 0 && (M > 0 || N > 0), "operator_norm_matrix: bad dimensions");
    PFDR_REQUIRE(nbInit >= 1 && itMax >= 0, "operator_norm_matrix: nbInit must be positive");
    cudaStream_t st = ctx().stream;
    const int Pm = std::min(M, N);
    // which system the reference iterates on (:111-168): the given symmetric matrix; else the explicit Gram matrix
    // of the smaller side when that is cheaper than two products with A per iteration; else A^tA applied as A^t(A x)
    // with starts in R^N - formed explicitly here as well (one pass over A instead of two per iteration).
    int n;
    DevIn<real> dA;
    DevBuf<real> S;
    const real *Sym;
    if (Pm == 0) {
        n = std::max(M, N);
        dA.bind(A, (size_t) n * n);
        Sym = dA.p;
    } else {
        const long long i = (long long) nbInit * itMax;
        const bool small_side = 2LL * M * N * i > (long long) M * N * Pm + (long long) Pm * Pm * i;
        const int rows_mode = small_side ? (Pm == M) : 0;           // A A^t (M x M) or A^t A (N x N)
        n = rows_mode ? M : N;
        dA.bind(A, (size_t) M * N);
        S.alloc((size_t) n * n);
        const int L = rows_mode ? N : M;
        dim3 grid(cdiv(n, 16), cdiv(n, 16));
        PFDR_LAUNCH(gram_kernel<real>, grid, 256, 0, dA.p, M, n, L, rows_mode, S.p);
        Sym = S.p;
    }
    // the reference's starts
    const int cores = host_cores();
    const int nb = (1 + (nbInit - 1) / cores) * cores, per = nb / cores;
    const char *te = getenv("PFDR_B200_OPNORM_TIME");
    const unsigned int t0 = te ? (unsigned int) strtoul(te, nullptr, 10) : (unsigned int) time(nullptr);
    constexpr int HALF_RAND_MAX = RAND_MAX / 2 + 1;
    std::vector<real> hx((size_t) nb * n);
    for (int t = 0; t < cores; t++) {
        unsigned int seed = t0 + (unsigned int) t;
        for (int k = 0; k < per; k++) {
            real *xs = hx.data() + (size_t) (t * per + k) * n;
            for (int j = 0; j < n; j++) xs[j] = ((real) (glibc_rand_r(&seed) - HALF_RAND_MAX)) / ((real) HALF_RAND_MAX);
        }
    }
    DevBuf<real> x((size_t) n * nb), out(nb);
    PFDR_CUDA(cudaMemcpyAsync(x.p, hx.data(), hx.size() * sizeof(real), cudaMemcpyHostToDevice, st));
    ctx().stats.h2d_bytes += (long long) (hx.size() * sizeof(real));
    std::vector<real> h(nb);
    if (n <= 1024) {
        DevBuf<real> t((size_t) n * nb);
        PFDR_LAUNCH(power_kernel<real>, nb, 256, 0, Sym, n, nTol, itMax, x.p, t.p, out.p);
        PFDR_CUDA(cudaMemcpyAsync(h.data(), out.p, nb * sizeof(real), cudaMemcpyDeviceToHost, st));
        PFDR_CUDA(cudaStreamSynchronize(st));
    } else {
        const int nseg = std::max(1, std::min(16, (4 * num_sms()) / (int) cdiv(n, 256)));
        for (int s0 = 0; s0 < nb; s0 += NBMAX) {
            const int bsz = std::min(NBMAX, nb - s0);
            DevBuf<real> part((size_t) nseg * bsz * n), bn(bsz);
            DevBuf<int> state(2 * (size_t) bsz + 1);
            state.zero();
            int *phase = state.p, *trips = state.p + bsz, *nact = state.p + 2 * bsz;
            PFDR_CUDA(cudaMemcpyAsync(nact, &bsz, sizeof(int), cudaMemcpyHostToDevice, st));
            real *xb = x.p + (size_t) s0 * n;
            PFDR_LAUNCH(norm_init_kernel<real>, bsz, 256, 0, n, xb, bn.p);
            dim3 grid(cdiv(n, 256), nseg);
            for (int it = 0; it <= itMax; it++) {                  // the product before the loop + at most itMax more
                if (bsz <= 8) PFDR_LAUNCH((symm_batched_kernel<real, 8>), grid, 256, 0, Sym, n, bsz, nseg, xb, bn.p, nact, part.p);
                else if (bsz <= 16) PFDR_LAUNCH((symm_batched_kernel<real, 16>), grid, 256, 0, Sym, n, bsz, nseg, xb, bn.p, nact, part.p);
                else PFDR_LAUNCH((symm_batched_kernel<real, NBMAX>), grid, 256, 0, Sym, n, bsz, nseg, xb, bn.p, nact, part.p);
                PFDR_LAUNCH(power_step_kernel<real>, bsz, 256, 0, n, bsz, nseg, part.p, xb, bn.p, phase, trips, nTol, itMax, nact);
                if ((it & 15) == 15) {                                 // every 16 trips: has every start stopped?
                    int left = 0;
                    PFDR_CUDA(cudaMemcpyAsync(&left, nact, sizeof(int), cudaMemcpyDeviceToHost, st));
                    PFDR_CUDA(cudaStreamSynchronize(st));
                    if (left == 0) break;
                }
            }
            PFDR_CUDA(cudaMemcpyAsync(h.data() + s0, bn.p, bsz * sizeof(real), cudaMemcpyDeviceToHost, st));
            PFDR_CUDA(cudaStreamSynchronize(st));
        }
    }
    ctx().stats.d2h_bytes += (long long) (nb * sizeof(real));
    real best = 0;
    for (real v : h) if (v > best) best = v;                                                            // ref :204
    if (verbose) { printf("operator norm (B200, %d starts over %d host cores' seeds): %g\n", nb, cores, (double) best); fflush(stdout); }
    return best;
}
template float operator_norm_matrix<float>(int, int, const float *, float, int, int, int);
template double operator_norm_matrix<double>(int, int, const double *, double, int, int, int);

}  // namespace pfdr
