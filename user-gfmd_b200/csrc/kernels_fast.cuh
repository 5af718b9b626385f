// Specialised kernels for large power-of-two grids (selected by fast_plan when
// the grid qualifies; otherwise the generic kernels of kernels_generic.cuh run).
#pragma once

#include "fft_engine.cuh"
#include "kernels_generic.cuh"

namespace gfmd {

inline int fast_plan(const GridDesc &g, int &fast_rows, int &fast_cols)
{
  (void) g;
  fast_rows = 0;
  fast_cols = 0;
  return 0;
}

inline int fast_rows_fwd(int, const double *, double2 *, const GridDesc &, const double2 *, const FftDesc &,
                         cudaStream_t, long long *)
{
  return 1;
}

inline int fast_rows_inv(int, const double2 *, double *, const GridDesc &, const double2 *, const FftDesc &,
                         cudaStream_t, long long *)
{
  return 1;
}

inline int fast_cols_fused(int, const double2 *, double2 *, const GridDesc &, const FftDesc &, const double *,
                           const double *, double *, StepResults *, cudaStream_t, long long *)
{
  return 1;
}

}  // namespace gfmd
