// quad_dia.cuh - launchers of the constant-offset-layout kernels of the quadratic family (quad_dia.cu)
#pragma once
#include "quad.cuh"

namespace pfdr {

template <typename real> void quad_dia_launch_init(const QP<real> &p, const real *X0, int grid);
template <typename real> void quad_dia_launch_recond(const QP<real> &p, int grid);
template <typename real> void quad_dia_launch_iter(const QP<real> &p, int grid);

}  // namespace pfdr
