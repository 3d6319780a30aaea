set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8
timeout 900 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v11.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
