set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8
for hb in 1 0; do
T2_HIER_BARRIER=$hb T2_VERBOSE=1 timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_v10_hb$hb.json 2> gpurun_out/bench_v10_hb$hb.err; grep t2b200 gpurun_out/bench_v10_hb$hb.err | tail -1
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v10_hb$hb.json") if l.startswith("{")][-1])
print("HIER $hb value", d["value"], "e2e", d["e2e"]["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
done
T2_CLUSTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_persistent -c 1 -o gpurun_out/prof_dec_v10 python tools/run_decoder_once.py 100 2>&1 | tail -3
