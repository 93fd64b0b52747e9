# one full ncu capture per register solver kernel at a reduced column count (ncu replays each launch ~40x)
set -x
NCOL=${NCOL:-8192}
for k in sw_2stream_reg_kernel lw_noscat_reg_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/ncu_$k \
    python bench.py --steps 1 --warmup 1 --no-cpu --ncol $NCOL > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
