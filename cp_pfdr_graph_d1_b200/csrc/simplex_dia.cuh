// simplex_dia.cuh - launchers of the constant-offset-layout kernels of the simplex family (simplex_dia.cu)
#pragma once
#include "simplex.cuh"

namespace pfdr {

// labels per lane on this path: one 16-byte access per piece of a K-vector
template <typename real> struct SimplexDiaLanes { static constexpr int LPT = sizeof(real) == 4 ? 4 : 2; };

template <typename real> void simplex_dia_launch_iter(const SP<real> &p, int grid);
template <typename real> void simplex_dia_launch_precond(const SP<real> &p, const real *Pinit, int recond, int grid);

}  // namespace pfdr
