# one full ncu capture per kernel regex at a reduced column count; usage: gpu_ncu.sh TAG kernel-regex ...
TAG=$1; shift
mkdir -p gpurun_out
NCOL=${NCOL:-16384}
for k in "$@"; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s ${SKIP:-0} -c 1 -f -o gpurun_out/${TAG}_ncu_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-seq --ncol $NCOL > gpurun_out/${TAG}_ncu_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_ncu_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_ncu_$k.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_${k}_src.csv 2>/dev/null
  [ -n "$KEEP_REP" ] || rm -f gpurun_out/${TAG}_ncu_$k.ncu-rep   # gpurun copies back at most 64 MiB
done
ls -la gpurun_out | grep ${TAG}_ncu
