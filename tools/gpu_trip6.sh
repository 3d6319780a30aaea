set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "umma or persistent or philox or full_size" 2>&1 | tail -15
T2_L2_PIN_FRAC=0.5 timeout 600 python bench.py --decoder-impl persistent --steps 2 --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; tail -2 gpurun_out/bench_v5.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v5.json") if l.startswith("{")][-1])
print("value", d["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
