set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -25
timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -2 gpurun_out/bench_v7.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v7.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dec_step_us", d["decoder_step_us"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v7.csv python tools/run_decoder_once.py 50 2>&1 | tail -2
