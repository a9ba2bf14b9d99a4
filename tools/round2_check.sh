# gpurun call 1: LSTM engine A/B (tests + bench) and the tests touching this session's other changes (deltas entry point, frontend.cu).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 400 python -m pytest tests/test_gpu_lstm.py -q -x 2>&1 | tail -25 > gpurun_out/t_lstm.txt; tail -5 gpurun_out/t_lstm.txt
B="--no-cpu-baseline --no-gpu-library-baseline --steps 30 --warmup 5"
for m in seq-lstm lstm; do
  HOWL_LSTM_ENGINE=0 timeout 200 python bench.py --model $m $B > gpurun_out/bench_${m}_e0.json 2> gpurun_out/bench_${m}_e0.err
  timeout 200 python bench.py --model $m $B > gpurun_out/bench_${m}_e1.json 2> gpurun_out/bench_${m}_e1.err
  python - <<PY
import json
for e in (0, 1):
    try:
        d = json.load(open("gpurun_out/bench_${m}_e%d.json" % e)); g = d["groups_ms"]
        print("${m}", "engine", e, round(d["ms_per_step"], 4), "ms", round(d["value"]), "utt/s  fwd", g.get("lstm_fwd"), "bwd", g.get("lstm_bwd"), "e2e", round(d["e2e"]["value"]))
    except Exception as exc:
        print("${m}", e, "FAILED", exc)
PY
done
timeout 300 python -m pytest tests/test_gpu_api.py -q 2>&1 | tail -15 > gpurun_out/t_api.txt; tail -4 gpurun_out/t_api.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k frontend 2>&1 | tail -8 > gpurun_out/t_fe.txt; tail -3 gpurun_out/t_fe.txt
