# gpurun call 2 (last of the round): LSTM engine 2 (tests + A/B bench), the driver's default bench invocation, smoke, ncu of the
# recurrences, then as much of the remaining GPU suite as the budget allows.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_api.py -q -x 2>&1 | tail -25 > gpurun_out/t_lstm2.txt; tail -4 gpurun_out/t_lstm2.txt
B="--no-cpu-baseline --no-gpu-library-baseline --steps 30 --warmup 5"
for m in seq-lstm lstm; do
  HOWL_LSTM_ENGINE=1 timeout 100 python bench.py --model $m $B > gpurun_out/bench2_${m}_e1.json 2> gpurun_out/bench2_${m}_e1.err
  timeout 100 python bench.py --model $m $B > gpurun_out/bench2_${m}_e2.json 2> gpurun_out/bench2_${m}_e2.err
  python - <<PY
import json
for e in (1, 2):
    try:
        d = json.load(open("gpurun_out/bench2_${m}_e%d.json" % e)); g = d["groups_ms"]
        print("${m}", "engine", e, round(d["ms_per_step"], 4), "ms", round(d["value"]), "utt/s  fwd", g.get("lstm_fwd"), "bwd", g.get("lstm_bwd"), "e2e", round(d["e2e"]["value"]))
    except Exception as exc:
        print("${m}", e, "FAILED", exc)
PY
done
timeout 200 python bench.py > gpurun_out/r02_bench_res8_final.json 2> gpurun_out/r02_bench_res8_final.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_res8_final.json"))
    print("res8", round(d["ms_per_step"], 4), "ms", round(d["value"]), "utt/s e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "modules", d.get("drop_in_modules"))
except Exception as exc:
    print("res8 FAILED", exc)
PY
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd|lstm_bwd" -s 8 -c 4 -o gpurun_out/r02_lstm_pipe -f python bench.py --model seq-lstm --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --no-e2e > gpurun_out/r02_lstm_pipe.log 2>&1
ncu -i gpurun_out/r02_lstm_pipe.ncu-rep --page raw --csv > gpurun_out/r02_lstm_pipe_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
timeout ${REST_TIMEOUT:-300} python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_lstm.py --ignore=tests/test_gpu_api.py --durations=12 > gpurun_out/t_rest.txt 2>&1; tail -22 gpurun_out/t_rest.txt | cut -c1-180
