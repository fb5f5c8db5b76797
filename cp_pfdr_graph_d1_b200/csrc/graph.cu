// graph.cu - builds the SELL-C half-edge table on the device (see graph.cuh).
// One-time setup; CUB supplies the stable radix sort and the scans, everything else is ours.
#include "graph.cuh"
#include <cub/cub.cuh>

namespace pfdr {

namespace {

__global__ void degree_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, int E, int V,
                              int *__restrict__ deg, int *__restrict__ vkey, int *__restrict__ hid,
                              int *__restrict__ bad)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;   // half-edge id 2e+side
    if (i >= 2 * (size_t) E) return;
    const int e = (int) (i >> 1);
    const int v = (i & 1) ? Ev[e] : Eu[e];
    if (v < 0 || v >= V) { *bad = 1; vkey[i] = 0; hid[i] = (int) i; return; }
    atomicAdd(&deg[v], 1);      // integer count: order-independent
    vkey[i] = v;
    hid[i] = (int) i;
}

__global__ void slice_width_kernel(const int *__restrict__ deg, int V, int C, int nslices,
                                   long long *__restrict__ slots_of_slice)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int w = 0;
    for (int k = 0; k < C; k++) {
        const int v = s * C + k;
        if (v < V) w = max(w, deg[v]);
    }
    slots_of_slice[s] = (long long) w * C;
}

// p = position in the list sorted by (vertex, half-edge id)
__global__ void place_kernel(const int *__restrict__ svkey, const int *__restrict__ shid,
                             const int *__restrict__ vptr, const long long *__restrict__ slice_off,
                             size_t n, int C, int logC, int *__restrict__ pos, int *__restrict__ hkey)
{
    const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int v = svkey[p], h = shid[p];
    const int j = (int) (p - (size_t) vptr[v]);
    const long long slot = slice_off[v >> logC] + (long long) j * C + (v & (C - 1));
    pos[h] = (int) slot;
    hkey[slot] = h;
}

__global__ void link_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, size_t n,
                            const int *__restrict__ pos, int *__restrict__ nbr, int *__restrict__ twin)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = (int) (i >> 1);
    const int slot = pos[i];
    nbr[slot] = (i & 1) ? Eu[e] : Ev[e];
    twin[slot] = pos[i ^ 1];
}

__global__ void max_offset_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, int E, int *__restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (e < E) d = abs(Eu[e] - Ev[e]);
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(out, d);
}
__global__ void count_offset_kernel(const int *__restrict__ Eu, const int *__restrict__ Ev, int E, const int *__restrict__ band,
                                    int *__restrict__ count)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = e < E && abs(Eu[e] - Ev[e]) == *band;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}
}  // namespace

void build_sell(SellGraph &g, const int *dEu, const int *dEv, int V, int E, int C)
{
    PFDR_REQUIRE(V > 0 && E >= 0, "build_sell: V must be positive, E non-negative");
    PFDR_REQUIRE(C >= 1 && C <= 32 && (C & (C - 1)) == 0, "build_sell: C must be a power of two <= 32");
    PFDR_REQUIRE((long long) E * 2 < 2147483647LL, "build_sell: 2E must fit a 32-bit int");
    cudaStream_t st = ctx().stream;
    int logC = 0; while ((1 << logC) < C) logC++;
    g.V = V; g.E = E; g.C = C;
    g.nslices = (V + C - 1) / C;
    const size_t nh = 2 * (size_t) E;

    g.deg.alloc((size_t) V + 1);
    g.deg.zero();
    DevBuf<int> vkey(nh), hid(nh), svkey(nh), shid(nh), bad(1), vptr((size_t) V + 1);
    bad.zero();
    if (nh) PFDR_LAUNCH(degree_kernel, cdiv(nh, 256), 256, 0, dEu, dEv, E, V, g.deg.p, vkey.p, hid.p, bad.p);

    // vertex pointer (exclusive scan of the degrees) and slice offsets
    DevBuf<long long> sl((size_t) g.nslices + 1);
    sl.zero();
    g.slice_off.alloc((size_t) g.nslices + 1);
    PFDR_LAUNCH(slice_width_kernel, cdiv(g.nslices, 256), 256, 0, g.deg.p, V, C, g.nslices, sl.p);
    size_t tmp1 = 0, tmp2 = 0, tmp3 = 0;
    int vbits = 1; while (((long long) 1 << vbits) < (long long) V) vbits++;
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp1, g.deg.p, vptr.p, V + 1, st));
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, sl.p, g.slice_off.p, g.nslices + 1, st));
    if (nh) PFDR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp3, vkey.p, svkey.p, hid.p, shid.p,
                                                      (int) nh, 0, vbits, st));
    DevBuf<unsigned char> tmp(std::max(tmp1, std::max(tmp2, tmp3)) + 16);
    size_t tb = tmp.n;
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, g.deg.p, vptr.p, V + 1, st));
    tb = tmp.n;
    PFDR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, sl.p, g.slice_off.p, g.nslices + 1, st));
    ctx().stats.launches += 2;

    // widest slice = max degree
    DevBuf<int> dmax(1);
    size_t tmp4 = 0;
    PFDR_CUDA(cub::DeviceReduce::Max(nullptr, tmp4, g.deg.p, dmax.p, V, st));
    DevBuf<unsigned char> tmpm(tmp4 + 16);
    PFDR_CUDA(cub::DeviceReduce::Max(tmpm.p, tmp4, g.deg.p, dmax.p, V, st));
    ctx().stats.launches += 1;
    PFDR_CUDA(cudaMemcpyAsync(&g.max_width, dmax.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    // band detection (see build_slice_order)
    DevBuf<int> dband(2);
    dband.zero();
    if (E > 0) {
        PFDR_LAUNCH(max_offset_kernel, cdiv(E, 256), 256, 0, dEu, dEv, E, dband.p);
        PFDR_LAUNCH(count_offset_kernel, cdiv(E, 256), 256, 0, dEu, dEv, E, dband.p, dband.p + 1);
    }
    int hband[2] = {0, 0};
    PFDR_CUDA(cudaMemcpyAsync(hband, dband.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    long long nslots = 0; int hbad = 0;
    PFDR_CUDA(cudaMemcpyAsync(&nslots, g.slice_off.p + g.nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PFDR_CUDA(cudaStreamSynchronize(st));
    PFDR_REQUIRE(!hbad, "edge list refers to a vertex outside [0, V)");
    g.band = (hband[0] > 0 && (long long) hband[1] * 5 >= (long long) E) ? hband[0] : 0;
    PFDR_REQUIRE(nslots < 2147483647LL, "half-edge table exceeds 2^31 slots (very irregular degrees?)");
    g.nslots = nslots;
    ctx().stats.slots = nslots;

    const size_t ns = (size_t) std::max<long long>(nslots, 1);
    g.nbr.alloc(ns); g.twin.alloc(ns); g.hkey.alloc(ns);
    PFDR_CUDA(cudaMemsetAsync(g.nbr.p, 0xFF, ns * sizeof(int), st));
    PFDR_CUDA(cudaMemsetAsync(g.twin.p, 0xFF, ns * sizeof(int), st));
    PFDR_CUDA(cudaMemsetAsync(g.hkey.p, 0xFF, ns * sizeof(int), st));
    if (nh) {
        // stable sort by vertex keeps each vertex's half-edges in (edge id, side) order
        tb = tmp.n;
        PFDR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, vkey.p, svkey.p, hid.p, shid.p, (int) nh, 0, vbits, st));
        ctx().stats.launches += 1;
        DevBuf<int> pos(nh);
        PFDR_LAUNCH(place_kernel, cdiv(nh, 256), 256, 0, svkey.p, shid.p, vptr.p, g.slice_off.p, nh, C, logC,
                    pos.p, g.hkey.p);
        PFDR_LAUNCH(link_kernel, cdiv(nh, 256), 256, 0, dEu, dEv, nh, pos.p, g.nbr.p, g.twin.p);
        PFDR_CUDA(cudaStreamSynchronize(st));   // temporaries die here
    }
}

}  // namespace pfdr
