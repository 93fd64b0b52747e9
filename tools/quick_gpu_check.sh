python -m pytest tests -x -q -m gpu -k "solver or allsky or golden" 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_scan.json 2>gpurun_out/bench_scan.err; tail -2 gpurun_out/bench_scan.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_scan.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'])
for k in d['kernels'][:8]: print('  ',k['kernel'], round(k['ms_per_step'],3), round(k['share'],3))
"
