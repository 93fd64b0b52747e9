# Full GPU round: parity tests, bench line, reference arm, ncu launch list, one full ncu capture per top kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(time python -m pytest tests -q -m gpu 2>&1 | tail -6) 2>&1 | tee gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','max_abs_flux_err_vs_oracle_Wm2']}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print(d['roofline']); print(d['step_roofline']); print(d['clocks'])
for k in d['kernels']: print('  ',k['kernel'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['share'],3), round(k.get('frac') or 0,3))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>>gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
# launch list (same command as the bench; times under ncu are cold-cache and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-seq > gpurun_out/${TAG}_launches.log 2>&1
bash tools/gpu_ncu.sh ${TAG} sw_2stream_reg_kernel lw_noscat_reg_kernel gas_tau_g_kernel planck_g_kernel
SKIP=1 bash tools/gpu_ncu.sh ${TAG}sw gas_tau_g_kernel
