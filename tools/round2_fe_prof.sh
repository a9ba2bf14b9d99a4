# ncu source-level capture of K1 (frontend_kernel) at the res8 bench configuration
set -x
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"frontend_kernel" -s 4 -c 1 -o gpurun_out/r02_fe -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --no-e2e --no-module-path > gpurun_out/r02_fe.log 2>&1
ncu -i gpurun_out/r02_fe.ncu-rep --page raw --csv > gpurun_out/r02_fe_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_fe.ncu-rep --page source --csv > gpurun_out/r02_fe_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
