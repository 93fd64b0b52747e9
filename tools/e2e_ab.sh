# e2e (host-buffer entry) under the chunk-schedule switches; run under torchrun for N > 1.  usage: e2e_ab.sh NGPU
N=${1:-1}
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  export RRTMGPB_STREAM_RAMP=$1 RRTMGPB_STREAM_TAIL=$2 RRTMGPB_STREAM_TRACE=1
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-seq --no-extras > gpurun_out/e2e_ab.json 2> gpurun_out/e2e_ab.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-seq --no-extras > gpurun_out/e2e_ab.json 2> gpurun_out/e2e_ab.err
  fi
  python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_ab.json').read().strip().splitlines()[-1])
print('N=$N ramp=$1 tail=$2', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))
PY
  grep "stream trace" gpurun_out/e2e_ab.err | tail -8 | cut -c1-150 > gpurun_out/e2e_ab_trace_N${N}_r$1_t$2.txt
done
