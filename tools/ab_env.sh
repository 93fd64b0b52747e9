# usage: ab_env.sh VAR val1 val2 ... : quick A/B of the all-sky step under an environment switch
VAR=$1; shift
for v in "$@"; do
env $VAR=$v timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu --no-seq > gpurun_out/ab_$v.json 2>gpurun_out/ab.err; tail -2 gpurun_out/ab.err
python -c "
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('$VAR=$v', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])
for k in d['kernels'][:5]: print('  ',k['kernel'], round(k['ms_per_step'],3), round(k['share'],3))
"
done
