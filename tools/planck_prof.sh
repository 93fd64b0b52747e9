mkdir -p gpurun_out
timeout 200 python tools/kbench.py --lw-only --steps 5 --tag hoist 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:planck_g_kernel -s 2 -c 1 -f -o gpurun_out/r2p_planck python tools/kbench.py --lw-only --ncol 16384 --steps 1 > gpurun_out/r2p_planck.log 2>&1
ncu -i gpurun_out/r2p_planck.ncu-rep --page source --csv > gpurun_out/r2p_planck_src.csv 2>/dev/null
ncu -i gpurun_out/r2p_planck.ncu-rep --page raw --csv > gpurun_out/r2p_planck_raw.csv 2>/dev/null
rm -f gpurun_out/r2p_planck.ncu-rep
ls -la gpurun_out/r2p_*
