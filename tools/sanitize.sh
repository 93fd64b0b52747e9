# compute-sanitizer memcheck over the round-2 kernels at test sizes (run under gpurun)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest -x -q -m gpu \
  "tests/test_gas_optics_rows_path.py::test_rows_path_distinct_columns[True-False-True]" \
  "tests/test_gas_optics_rows_path.py::test_rows_path_mixed_and_ragged" \
  "tests/test_sw_ws_kernel.py::test_ws_kernel_matches_register_kernel_and_oracle[38-72-True-False-True]" \
  "tests/test_sw_ws_kernel.py::test_ws_kernel_matches_register_kernel_and_oracle[22-60-False-True-False]" \
  "tests/test_rrtmgp_symbols_parity.py::test_tau_and_planck_symbols_on_the_gfast_kernels" \
  "tests/test_single_precision.py::test_sp_lw_noscat" \
  > gpurun_out/r2_sanitizer.txt 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/r2_sanitizer.txt | head -20
