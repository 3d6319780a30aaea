set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -25
timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -2 gpurun_out/bench_v8.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v8.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
for c in 1 2 4; do
T2_CONV_CLUSTER=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc -c 30 --csv --log-file gpurun_out/conv_c$c.csv python tools/run_decoder_once.py 4 2>&1 | tail -1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/conv_c$c.csv")))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; vi = h.index('Metric Value'); ki = h.index('Kernel Name')
print("CONV CLUSTER $c:", [ (r[ki][28:52], r[vi]) for r in rows[hi+1:hi+10] ])
PY
done
