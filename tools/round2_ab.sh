# gpurun call 3: A/B of the L2 prefetch ahead of the conv TMA rings (res8), with the res8 parity tests on the new default.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -6 > gpurun_out/t_parity3.txt; tail -3 gpurun_out/t_parity3.txt
B="--no-cpu-baseline --no-gpu-library-baseline --no-module-path --steps 40 --warmup 5"
for e in 0 1 0 1; do
  HOWL_TC_L2_PREFETCH=$e timeout 100 python bench.py $B > gpurun_out/bench3_res8_p$e.json 2> gpurun_out/bench3_res8_p$e.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench3_res8_p$e.json")); g = d["groups_ms"]
    print("prefetch", $e, round(d["ms_per_step"], 4), "ms", round(d["value"]), "utt/s  fwd", g.get("conv3x3_fwd_tc"), "dgrad", g.get("conv3x3_dgrad_tc"), "wgrad", g.get("conv3x3_wgrad_tc"), "e2e", round(d["e2e"]["value"]))
except Exception as exc:
    print("prefetch", $e, "FAILED", exc)
PY
done
