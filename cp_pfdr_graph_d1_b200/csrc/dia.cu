// dia.cu - recognises constant-offset graphs and builds the stencil layout of dia.cuh on the device.
#include "dia.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace pfdr {

namespace {

// flags[0]: an edge refers to a vertex outside [0, V);  flags[1]: not a vertex-major constant-offset edge list
__global__ void __launch_bounds__(256) dia_scan_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, int E, int V,
                                                       int *__restrict__ table, int *__restrict__ flags)
{
    __shared__ int s_tab[DIA_MAXD];
    if (threadIdx.x < DIA_MAXD) s_tab[threadIdx.x] = ld_flag(table + threadIdx.x);
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int u = Eu[e], v = Ev[e];
    if (u < 0 || u >= V || v < 0 || v >= V) { flags[0] = 1; return; }
    const int d = v - u;
    if (d <= 0) { flags[1] = 1; return; }
    if (e > 0) {                                          // sorted by (Eu, d), strictly: vertex-major, no duplicate edge
        const int pu = Eu[e - 1], pd = Ev[e - 1] - pu;
        if (!(pu < u || (pu == u && pd < d))) { flags[1] = 1; return; }
    }
    for (int i = 0; i < DIA_MAXD; i++) if (s_tab[i] == d) return;
    for (int i = 0; i < DIA_MAXD; i++) {                  // not in the block's snapshot: look at / extend the global table
        int t = ld_flag(table + i);
        if (t == 0) t = atomicCAS(table + i, 0, d);
        if (t == 0 || t == d) return;
    }
    flags[1] = 1;                                         // more than DIA_MAXD distinct offsets
}

struct Offs { int ND; int d[4]; };

__device__ __forceinline__ int offset_index(const Offs &o, int d)
{
    int i = 0;
#pragma unroll
    for (int k = 1; k < DIA_MAXD; k++) if (k < o.ND && d == o.d[k]) i = k;
    return i;
}

// the thread of a vertex's first out-edge records the vertex's out-mask and the id of that edge
__global__ void __launch_bounds__(256) dia_out_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, int E, Offs o,
                                                      unsigned char *__restrict__ mout, int *__restrict__ eoff)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int u = Eu[e];
    if (e > 0 && Eu[e - 1] == u) return;
    unsigned m = 0;
    for (int j = e; j < E && j < e + o.ND && Eu[j] == u; j++) m |= 1u << offset_index(o, Ev[j] - u);
    mout[u] = (unsigned char) m;
    eoff[u] = e;
}

__global__ void __launch_bounds__(256) dia_in_kernel(int V, Offs o, const unsigned char *__restrict__ mout,
                                                     unsigned char *__restrict__ mask)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    unsigned m = mout[v];
#pragma unroll
    for (int i = 0; i < DIA_MAXD; i++)
        if (i < o.ND && v - o.d[i] >= 0 && ((mout[v - o.d[i]] >> i) & 1u)) m |= 16u << i;
    mask[v] = (unsigned char) m;
}

__global__ void chunk_key_kernel(int nchunks, int C, int band, int T, int nx, long long v0, int lowB, int highB, int nown,
                                 unsigned long long *__restrict__ key, int *__restrict__ val)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nchunks) return;
    // slab-sharded: chunks holding a vertex within the band of a neighbouring rank come first (slab.cuh)
    const bool interior = !((long long) s * C < lowB || ((long long) s + 1) * C > (long long) nown - highB);
    if (T <= 0) { key[s] = ((unsigned long long) interior << 62) | (unsigned long long) s; val[s] = s; return; }
    const long long v = v0 + (long long) s * C;                     // global number of the chunk's first vertex
    const long long z = v / band, r = v % band;
    const long long nz = (v0 + (long long) nchunks * C + band - 1) / band + 1;
    if (nx > 0) {
        // 3-D grid (rows of nx vertices): tiles of TY rows swept along z, and inside a tile patches of 8 rows x one chunk,
        // the 8 chunks of a patch consecutive (= the 8 warps of a CTA round): +-nx neighbours are loaded by the same SM
        const long long TY = max(8, T / nx / 8 * 8), x = r % nx, y = r / nx, npx = nx / C + 1;
        key[s] = (unsigned long long) (((((y / TY) * nz + z) * (TY / 8) + (y % TY) / 8) * npx + x / C) * 8 + y % 8);
    } else
        key[s] = (unsigned long long) (((r / T) * nz + z) * T + (r % T));
    key[s] |= (unsigned long long) interior << 62;
    val[s] = s;
}

}  // namespace

void build_chunk_order(int nchunks, int C, int band, int T, DevBuf<int> &order, int nx, long long v0, int lowB, int highB, int nown)
{
    cudaStream_t st = ctx().stream;
    DevBuf<unsigned long long> key(nchunks), skey(nchunks);
    DevBuf<int> val(nchunks);
    order.alloc(nchunks);
    PFDR_LAUNCH(chunk_key_kernel, cdiv(nchunks, 256), 256, 0, nchunks, C, std::max(band, 1), T, nx, v0, lowB, highB, nown > 0 ? nown : nchunks * C, key.p, val.p);
    size_t tb = 0;
    PFDR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, skey.p, val.p, order.p, nchunks, 0, 64, st));
    DevBuf<unsigned char> tmp(tb + 16);
    PFDR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, skey.p, val.p, order.p, nchunks, 0, 64, st));
    ctx().stats.launches += 1;
    PFDR_CUDA(cudaStreamSynchronize(st));       // temporaries die here
}

bool build_dia(DiaGraph &g, const int *dEu, const int *dEv, int V, int E)
{
    if (V <= 0 || E <= 0) return false;
    cudaStream_t st = ctx().stream;
    DevBuf<int> scratch(8);                     // [0..3] offset table, [4..5] flags
    scratch.zero();
    PFDR_LAUNCH(dia_scan_kernel, cdiv(E, 256), 256, 0, dEu, dEv, E, V, scratch.p, scratch.p + 4);
    int h[8];
    PFDR_CUDA(cudaMemcpyAsync(h, scratch.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaStreamSynchronize(st));
    PFDR_REQUIRE(!h[4], "edge list refers to a vertex outside [0, V)");
    if (h[5]) return false;
    Offs o{};
    for (int i = 0; i < DIA_MAXD; i++) if (h[i] > 0) o.d[o.ND++] = h[i];
    if (o.ND == 0) return false;
    std::sort(o.d, o.d + o.ND);
    g.V = V; g.E = E; g.ND = o.ND;
    for (int i = 0; i < 4; i++) g.off[i] = i < o.ND ? o.d[i] : 0;
    DevBuf<unsigned char> mout(V);
    mout.zero();
    g.mask.alloc(V); g.eoff.alloc(V);
    g.eoff.zero();
    PFDR_LAUNCH(dia_out_kernel, cdiv(E, 256), 256, 0, dEu, dEv, E, o, mout.p, g.eoff.p);
    PFDR_LAUNCH(dia_in_kernel, cdiv(V, 256), 256, 0, V, o, mout.p, g.mask.p);
    PFDR_CUDA(cudaStreamSynchronize(st));       // mout dies here
    return true;
}

}  // namespace pfdr
