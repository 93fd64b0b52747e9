// tma.cuh - Tensor Memory Accelerator plumbing for the solver kernels (sm_90+/sm_100a PTX, no CUTLASS).
//
// The register solvers consume, per g-point, one (16 columns x nlay layers) tile of every input plane
// (ncol, nlay, ngpt).  That tile is a dense 2-D box of a 3-D tensor, i.e. exactly what cp.async.bulk.tensor moves
// with ONE instruction issued by one thread: no per-lane address arithmetic, no 8-byte LDGSTS requests that touch
// eight 128-byte lines each (measured: the LW solver sat at 87% L1/TEX throughput on them,
// profiles/r1_v10_solvers.txt), completion signalled on an mbarrier.
//
// Shared-memory tile layout: [row = layer][16 columns] doubles = 128-byte rows, SWIZZLE_128B (the 16-byte chunk index
// is XORed with row & 7), so the solver's access pattern - 4 neighbouring columns x 8 rows that are CL layers apart
// per warp request - spreads over the banks instead of hitting 8 rows of the same bank group.
#pragma once
#include <cuda.h>
#include <cstdint>
#include "../common.cuh"

namespace rrtmgpb {

constexpr int kTmaCols = 16;  // columns per tile = one 128-byte row of doubles

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box (kTmaCols, box_rows, 1) of a (ncol, nrows, ngpt) plane -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int col0, int row0, int g) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(col0), "r"(row0), "r"(g)
      : "memory");
}
// element offset of (row, col) inside a swizzled tile whose base is 1024-byte aligned
__host__ __device__ __forceinline__ int tile_off(int row, int col) {
  static_assert(sizeof(Float) == 8 || sizeof(Float) == 4, "");
  if (sizeof(Float) == 8) {
    const int chunk = (col >> 1) ^ (row & 7);
    return row * kTmaCols + chunk * 2 + (col & 1);
  } else {  // single precision: 16 columns = 64-byte rows, SWIZZLE_64B (chunk index ^ ((row >> 1) & 3))
    const int chunk = (col >> 2) ^ ((row >> 1) & 3);
    return row * kTmaCols + chunk * 4 + (col & 3);
  }
}
__device__ __forceinline__ const Float* tile_at(const Float* tile, int row, int col) { return tile + tile_off(row, col); }
// bytes of one tile in shared memory, rounded up to the swizzle atom (1024 B)
__host__ __device__ inline size_t tile_bytes(int rows) { return ((size_t)rows * kTmaCols * sizeof(Float) + 1023) & ~(size_t)1023; }

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Tensor map of a Fortran-ordered (ncol, nrows, ngpt) plane with box (kTmaCols, box_rows, 1); box_rows = 0: nrows.  A
// box taller than the plane (or started at a negative row) is allowed: rows outside the tensor arrive as zeros
// (CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = zero fill), which the solvers use as exact pass-through padding layers.
// Single precision: 64-byte rows, SWIZZLE_64B (tile_off), tested by tests/test_single_precision.py.
// Returns false when the plane cannot be described (odd ncol -> strides not multiples of 16 B, misaligned base, more
// than 256 rows, no driver).
inline bool make_plane_tmap(CUtensorMap* tm, const Float* base, int ncol, int nrows, int ngpt, int box_rows = 0) {
  if (box_rows <= 0) box_rows = nrows;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || !base || box_rows > 256 || ((size_t)ncol * sizeof(Float)) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0)
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)ncol, (cuuint64_t)nrows, (cuuint64_t)ngpt};
  const cuuint64_t strides[2] = {(cuuint64_t)ncol * sizeof(Float), (cuuint64_t)ncol * nrows * sizeof(Float)};
  const cuuint32_t box[3] = {(cuuint32_t)kTmaCols, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, sizeof(Float) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                         const_cast<Float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sizeof(Float) == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace rrtmgpb
