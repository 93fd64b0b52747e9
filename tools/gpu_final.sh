# final verification of the tree: all GPU tests, smoke, the bench line, the reference arm
TAG=${1:-r2g}
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu 2>&1 | tail -4) 2>&1 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','gpu_launches','max_abs_flux_err_vs_oracle_Wm2','value_express','value_reference_call_sequence','value_host_pointer_call_sequence']}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print(d['roofline']['frac'], d['step_roofline'], d['clocks'])
print({k:(v.get('value'), v.get('frac_of_unfused_abi_roofline')) for k,v in (d.get('other_configs') or {}).items()})
for k in d['kernels'][:6]: print('  ',k['kernel'], round(k['ms_per_step'],3), round(k.get('frac') or 0,3))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>>gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
