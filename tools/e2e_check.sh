timeout 300 python -m pytest tests/test_streaming.py -x -q -m gpu 2>&1 | tail -3
RRTMGPB_STREAM_TRACE=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-seq --no-extras > gpurun_out/e2e_check.json 2> gpurun_out/e2e_check.err
grep "stream trace" gpurun_out/e2e_check.err | tail -7
python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_check.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
