#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu evidence for the headline step.  usage: tools/ncu_capture.sh TAG [NCOL]
#   1. launch list of the bench command (gpu__time_duration.sum per launch, --clock-control none)
#   2. one `--set full` capture of each dominant kernel (one launch each, NCOL columns, source import on)
# Outputs under gpurun_out/: ${TAG}_launches.csv, ${TAG}_full.ncu-rep, ${TAG}_full_raw.csv
TAG=${1:-r2}; NCOL=${2:-16384}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-seq --no-extras --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1
# kbench: warm-up step (checks on), step (checks off), then the profiled steps; the 5 kernels launch once per step
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k 'regex:sw_2stream_reg_kernel|lw_noscat_reg_kernel|gas_tau_g_kernel|planck_g_kernel' --launch-skip 10 --launch-count 5 \
    -f -o gpurun_out/${TAG}_full python tools/kbench.py --ncol ${NCOL} --steps 2 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --kernel-name regex:sw_2stream_reg_kernel > gpurun_out/${TAG}_sw_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
