# Full GPU round (round 2): parity tests, bench line, reference arm, ncu launch list, one --set full capture of the five
# dominant kernels.  usage (from the repo root, under gpurun): bash tools/gpu_round2.sh TAG
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(time python -m pytest tests -q -m gpu 2>&1 | tail -6) 2>&1 | tee gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','gpu_launches','max_abs_flux_err_vs_oracle_Wm2','value_express','value_reference_call_sequence','value_host_pointer_call_sequence']}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print(d['roofline']); print(d['step_roofline']); print(d['clocks']); print(d.get('single_precision_solvers'))
print({k:(v.get('value'), v.get('frac_of_unfused_abi_roofline')) for k,v in (d.get('other_configs') or {}).items()})
for k in d['kernels']: print('  ',k['kernel'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['share'],3), round(k.get('frac') or 0,3))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>>gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
bash tools/ncu_capture.sh ${TAG} 16384 > gpurun_out/${TAG}_ncu_capture.log 2>&1
rm -f gpurun_out/${TAG}_full.ncu-rep
tail -4 gpurun_out/${TAG}_ncu_capture.log
