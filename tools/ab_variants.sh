# A/B of compile-time variants: swaps pre-built variant libraries (build/variants/*.so) into place one at a time
# usage (under gpurun): bash tools/ab_variants.sh
mkdir -p gpurun_out
L=rte_rrtmgp_b200/lib/librte_rrtmgp_b200.so
cp $L /tmp/lib_orig.so
for v in build/variants/*.so; do
  cp $v $L
  timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu --no-seq > gpurun_out/abv.json 2>gpurun_out/abv.err || tail -3 gpurun_out/abv.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/abv.json').read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],2), [(k['kernel'], round(k['ms_per_step'],2)) for k in d['kernels'][:5]])
PY
done
cp /tmp/lib_orig.so $L
