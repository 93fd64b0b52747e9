# quick GPU check: parity tests touching the changed path + a 3-step bench; usage: gpu_quick.sh [pytest -k expr]
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu ${1:+-k "$1"} 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu --no-seq > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'])
for k in d['kernels'][:12]: print('  ',k['kernel'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['share'],3), round(k.get('frac') or 0,3))
PY
