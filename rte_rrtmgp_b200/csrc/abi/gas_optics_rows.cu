// gas_optics_rows.cu - the ROWS instantiations of gas_tau_g_kernel (kernels/gas_optics_gfast.cuh: tau_band_rows, the
// lanes-along-g-points mapping for warps of unrelated neighbouring columns), in a translation unit - and therefore a cubin -
// of their own, so that the default instantiations in gas_optics_fused.cu keep their code placement (see the note at
// launch_tau_rows' declaration).  Selected by rrtmgpb_set_gas_optics_rows_path(1) / RRTMGPB_TAU_ROWS=1.
#include "../kernels/gas_optics_gfast.cuh"

namespace rrtmgpb {

namespace {
template <typename K>
void launch(K kern, const FusedParams& p, const TablesT& tt, unsigned grid, size_t smem) {
  if (smem > 48 * 1024) RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kGThreads, smem, stream()>>>(p, tt);
}
}  // namespace

void launch_tau_rows(const FusedParams& p, const TablesT& tt, unsigned grid, size_t smem, bool sw, int kind, bool abi) {
#ifdef RTE_USE_SP
  (void)p; (void)tt; (void)grid; (void)smem; (void)sw; (void)kind; (void)abi;   // double-precision table layout only; never selected
#else
  if (abi) launch(gas_tau_g_kernel<false, 2, false, 0, true, true, true>, p, tt, grid, smem);
  else if (sw) {
    if (kind) launch(gas_tau_g_kernel<true, 2, false, 1, true, false, true>, p, tt, grid, smem);
    else launch(gas_tau_g_kernel<true, 2, false, 0, true, false, true>, p, tt, grid, smem);
  } else {
    if (kind) launch(gas_tau_g_kernel<false, 2, false, 1, true, false, true>, p, tt, grid, smem);
    else launch(gas_tau_g_kernel<false, 2, false, 0, true, false, true>, p, tt, grid, smem);
  }
#endif
}

}  // namespace rrtmgpb
