# A/B of the warp-specialised SW kernel: parity tests, then per-kernel times with RRTMGPB_SW_WS=0/1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sw_ws_kernel.py -x -q -m gpu 2>&1 | tail -3
for ws in 0 1; do
  RRTMGPB_SW_WS=$ws timeout 300 python tools/kbench.py --sw-only --steps 5 --tag "ws=$ws" 2>&1 | tail -1 | tee -a gpurun_out/r2_ws_ab.jsonl
done
