set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "umma or persistent or philox or full_size" 2>&1 | tail -15
for cfg in "8 0.5" "1 0.5" "8 1.0"; do
  set -- $cfg
  T2_CLUSTER=$1 T2_L2_PIN_FRAC=$2 timeout 600 python bench.py --decoder-impl persistent --steps 2 --no-cpu-baseline > gpurun_out/bench_v4_c$1_p$2.json 2> gpurun_out/bench_v4_c$1_p$2.err; tail -2 gpurun_out/bench_v4_c$1_p$2.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_v4_c$1_p$2.json") if l.startswith("{")][-1])
    print("CLUSTER $1 PIN $2 value", d["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"])
    for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
except Exception as e: print("no result", e)
PY
done
