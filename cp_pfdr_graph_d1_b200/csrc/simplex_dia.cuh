// simplex_dia.cuh - launchers of the constant-offset-layout kernels of the simplex family (simplex_dia.cu)
#pragma once
#include "simplex.cuh"

namespace pfdr {

// labels per lane on this path: four (16 bytes in float32, 2 x 16 bytes in float64: with two, the float64 kernel
// issues twice the instructions per vertex and is issue bound - 14.3 vs 9.9 ms per iteration on C4 252^3)
template <typename real> struct SimplexDiaLanes { static constexpr int LPT = 4; };

template <typename real> void simplex_dia_launch_iter(const SP<real> &p, int grid);
template <typename real> void simplex_dia_launch_precond(const SP<real> &p, const real *Pinit, int recond, int grid);

}  // namespace pfdr
