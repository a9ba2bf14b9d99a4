set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for m in res8 mobilenet seq-lstm lstm las; do
  timeout 600 python bench.py --model $m > gpurun_out/r02_bench_$m.json 2> gpurun_out/r02_bench_$m.err
  tail -c 200 gpurun_out/r02_bench_$m.json
done
timeout 300 python bench.py --seconds 0.5 --batch 64 --steps 50 --warmup 5 > gpurun_out/r02_bench_res8_config1.json 2>/dev/null
