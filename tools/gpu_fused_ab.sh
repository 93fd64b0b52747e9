# A/B of the fused gas-optics path + one full ncu capture of its two kernels
mkdir -p gpurun_out
bash tools/ab_env.sh RRTMGPB_FUSED 1
NCOL=${NCOL:-16384}
for k in gas_tau_fused_kernel planck_fused_kernel; do
  RRTMGPB_FUSED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/fused_ncu_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu --ncol $NCOL > gpurun_out/fused_ncu_$k.log 2>&1
  ncu -i gpurun_out/fused_ncu_$k.ncu-rep --page raw --csv > gpurun_out/fused_ncu_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
