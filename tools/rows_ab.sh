mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gas_optics_rows_path.py -x -q -m gpu 2>&1 | tail -3
for r in 0 1; do
  timeout 300 python tools/kbench.py --distinct --nlay 60 --steps 3 --rows $r --tag "distinct rows=$r" 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/r2_rows_ab.jsonl
done
timeout 300 python tools/kbench.py --steps 3 --tag "replicated rows=default" 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/r2_rows_ab.jsonl
timeout 300 python tools/kbench.py --steps 3 --rows 1 --tag "replicated rows=1" 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/r2_rows_ab.jsonl
