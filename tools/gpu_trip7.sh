set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -15
timeout 600 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -2 gpurun_out/bench_v6.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v6.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v6.csv python tools/run_decoder_once.py 100 2>&1 | tail -2
