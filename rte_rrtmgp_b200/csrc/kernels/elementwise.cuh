// elementwise.cuh - launch helper for the pure-bandwidth kernels (fills, optical-property
// arithmetic, frontend glue).  Every (ncol,nlay,ngpt) array is Fortran-ordered with columns
// innermost, so a flat index is already the coalesced order: consecutive threads touch
// consecutive addresses.  Grid: a multiple of the SM count (148 on B200), grid-stride loop, so
// the same launch shape serves a 32-element unit test and a 1.2e9-element plane.
#pragma once
#include "../common.cuh"

namespace rrtmgpb {

constexpr int kEltThreads = 256;
constexpr int kSMs = 148;

template <typename F>
__global__ void __launch_bounds__(kEltThreads) elementwise_kernel(size_t n, F f) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <typename F>
inline void launch_elementwise(size_t n, F f, const char* name = nullptr) {
  if (n == 0) return;
  KernelTimer timer(name ? name : (tl_op_name ? tl_op_name : "elementwise"));
  size_t blocks = (n + kEltThreads - 1) / kEltThreads;
  const size_t cap = (size_t)kSMs * 16;  // 8 resident CTAs/SM x 2 waves
  if (blocks > cap) blocks = cap;
  elementwise_kernel<<<(unsigned)blocks, kEltThreads, 0, stream()>>>(n, f);
  RB_LAUNCH_CHECK();
}

}  // namespace rrtmgpb
