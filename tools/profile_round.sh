set -x
timeout 300 python -m pytest tests/test_gpu_las.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --model las --steps 10 --warmup 3 > gpurun_out/bench_las.json 2>gpurun_out/bench_las.err
python tools/launch_times.py > gpurun_out/r02_launch_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-library-baseline > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3|frontend_kernel|conv0_|bn_bwd_head" -s 69 -c 23 -o gpurun_out/r02_prof -f python tools/launch_times.py > gpurun_out/r02_prof.log 2>&1
ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_prof_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd|lstm_bwd|lstm_atb|ctc_kernel" -s 12 -c 6 -o gpurun_out/r02_lstm -f python bench.py --model seq-lstm --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --no-e2e > gpurun_out/r02_lstm.log 2>&1
ncu -i gpurun_out/r02_lstm.ncu-rep --page raw --csv > gpurun_out/r02_lstm_raw.csv 2>/dev/null
ls -la gpurun_out | tail -15
