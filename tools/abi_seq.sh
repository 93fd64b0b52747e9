timeout 600 python -m pytest tests/test_rrtmgp_symbols_parity.py tests/test_allsky_parity.py tests/test_threads.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/seq_bench.json 2>gpurun_out/seq_bench.err; tail -2 gpurun_out/seq_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/seq_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','value_reference_call_sequence']})
PY
