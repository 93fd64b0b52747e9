# ncu --set full of the two tau kernels (one launch each); usage: tau_prof.sh TAG [kbench args]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gas_tau_g_kernel -s 4 -c 2 -f -o gpurun_out/${TAG} python tools/kbench.py --ncol 16384 --steps 1 "$@" > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}.ncu-rep
ls -la gpurun_out/${TAG}*
