// dia.cuh - constant-offset ("stencil" / DIA) layout of the graph in HBM.
//
// Every BASELINE.json workload but C3 lives on a regular grid: all edges are (v, v + d) for a handful of offsets d
// (1, n, n^2 on an n^3 grid) and the edge list is emitted vertex-major.  For such graphs the half-edge table of
// graph.cuh (neighbour + twin-slot indices, 8 bytes per half-edge, and a dependent index -> gather chain per row)
// is replaced by arithmetic:
//
//   out_i(v)  = edge (v, v + d_i)   exists iff bit i       of mask[v]        i < ND <= DIA_MAXD
//   in_i(v)   = edge (v - d_i, v)   exists iff bit (4 + i) of mask[v]
//   copy of vertex v on out_i(v)    Z[ i      ][v]        its twin:  Z[ND + i][v + d_i]
//   copy of vertex v on in_i(v)     Z[ND + i  ][v]        its twin:  Z[i     ][v - d_i]
//
// so the edge copies are 2.ND vertex-indexed planes, every access of the iteration kernels is a unit-stride (coalesced)
// access at a computed address, no index is streamed and all loads of a vertex are independent of each other.
// The reference adds a vertex's half-edges in edge-id order (src/PFDR_graph_quadratic_d1_l1.cpp:492-497).  The layout
// is only used when the edge list is sorted by (Eu, d) strictly ascending with Eu < Ev (the numbering every grid
// generator produces), where that order is  in_{ND-1}, ..., in_0, out_0, ..., out_{ND-1} : the per-vertex sums are
// taken in exactly this order, so results are bit-identical to the half-edge-table path.  Anything else (irregular
// graphs, self-loops, duplicate edges, other numberings, more than DIA_MAXD offsets) keeps the half-edge table.
#pragma once
#include "common.cuh"

namespace pfdr {

constexpr int DIA_MAXD = 3;

struct DiaGraph {
    int V = 0, E = 0, ND = 0;
    int off[4] = {0, 0, 0, 0};        // offsets, ascending
    DevBuf<unsigned char> mask;       // [V] bits 0..ND-1: out_i exists, bits 4..4+ND-1: in_i exists
    DevBuf<int> eoff;                 // [V] edge id of the first out-edge of v (valid where v has one);
                                      //     edge id of out_i(v) = eoff[v] + popc(mask[v] & ((1 << i) - 1))
    int band() const { return ND ? off[ND - 1] : 0; }
};

// Eu, Ev: device arrays of length E.  Returns false (g untouched apart from scratch) when the graph does not have the
// structure above; throws on edges outside [0, V).
bool build_dia(DiaGraph &g, const int *dEu, const int *dEv, int V, int E);

// Processing order of chunks of C consecutive vertices for banded graphs (see graph.cuh build_slice_order): vertex
// v = z*band + r is visited in the order (r / T, z, r % T).  nx > 0 (3-D grid with rows of nx vertices): the tile is a
// band of rows and is walked in patches of 8 rows x one chunk, so that the 8 warps of a CTA hold vertical neighbours.
// v0: global number of the first vertex of chunk 0 (slab-sharded solves order their owned range); lowB / highB > 0:
// chunks holding one of the first lowB / last highB of the nown vertices come first (slab.cuh); T <= 0: no tiling.
void build_chunk_order(int nchunks, int C, int band, int T, DevBuf<int> &order, int nx = 0, long long v0 = 0,
                       int lowB = 0, int highB = 0, int nown = 0);

// device helpers --------------------------------------------------------------------------------
// half-edge h = 0 .. 2ND-1 of a vertex in the reference's summation order
template <int ND> struct DiaHalf {
    __host__ __device__ static constexpr int idx(int h) { return h < ND ? ND - 1 - h : h - ND; }     // offset index i
    __host__ __device__ static constexpr bool out(int h) { return h >= ND; }                          // out_i or in_i
    __host__ __device__ static constexpr int own_plane(int h) { return out(h) ? idx(h) : ND + idx(h); }
    __host__ __device__ static constexpr int twin_plane(int h) { return out(h) ? ND + idx(h) : idx(h); }
    __host__ __device__ static constexpr int mask_bit(int h) { return out(h) ? idx(h) : 4 + idx(h); }
};

}  // namespace pfdr
