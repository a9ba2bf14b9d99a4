# gpurun: MobileNetV2 stem backward with the im2col-gradient tile staged in shared memory -- parity tests + config-3 bench
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mobilenet.py -q -x 2>&1 | tail -6 > gpurun_out/t_mbn.txt; tail -3 gpurun_out/t_mbn.txt
timeout 200 python bench.py --model mobilenet --no-cpu-baseline --no-gpu-library-baseline --steps 10 --warmup 3 > gpurun_out/bench4_mobilenet.json 2> gpurun_out/bench4_mobilenet.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench4_mobilenet.json")); g = d["groups_ms"]
    print("mobilenet", round(d["ms_per_step"], 3), "ms", round(d["value"]), "utt/s e2e", round(d["e2e"]["value"]), {k: v for k, v in g.items() if "stem" in k})
except Exception as exc:
    print("mobilenet FAILED", exc)
PY
