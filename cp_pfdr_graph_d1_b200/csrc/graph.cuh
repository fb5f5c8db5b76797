// graph.cuh - half-edge layout of the graph in HBM.
//
// The reference keeps one record per edge and scatters edge copies to the vertices in a serial loop
// (src/PFDR_graph_quadratic_d1_l1.cpp:492-497).  Here every edge e=(u,v) becomes two half-edges,
// (2e) owned by u and (2e+1) owned by v, stored in a sliced-ELL ("SELL-C") table keyed by vertex:
//
//   slice s      = vertices [s*C, (s+1)*C)                      (C = vertices per warp row)
//   slot(v, j)   = slice_off[s] + j*C + (v % C)                 j-th half-edge of v, j < width(s)
//   width(s)     = max degree in the slice; missing entries are padding (nbr = -1)
//
// A vertex's half-edges are ordered by (edge id, side), i.e. exactly the order in which the
// reference's serial scatter adds them, so per-vertex sums are taken in the reference's order
// and are bit-reproducible from run to run (no float atomics anywhere).
// Lanes of a warp walk consecutive vertices, so slot j of a slice is one coalesced row.
#pragma once
#include "common.cuh"

namespace pfdr {

struct SellGraph {
    int V = 0, E = 0, C = 32;
    int nslices = 0;
    long long nslots = 0;
    int max_width = 0;             // widest slice (rows)
    int band = 0;                  // |u - v| shared by >= 20 % of the edges and maximal among all edges (n^2 on an
                                   // n^3 grid), 0 if the graph has no such band (-> dia.cuh build_chunk_order)
    DevBuf<long long> slice_off;   // [nslices+1] slot offset of each slice
    DevBuf<int> nbr;               // [nslots] vertex at the other end (-1: padding)
    DevBuf<int> twin;              // [nslots] slot of the opposite half-edge
    DevBuf<int> hkey;              // [nslots] 2*e+side (-1: padding)
    DevBuf<int> deg;               // [V]
};

// Eu, Ev: device arrays of length E.  C must be a power of two <= 32.
void build_sell(SellGraph &g, const int *dEu, const int *dEv, int V, int E, int C);

}  // namespace pfdr
