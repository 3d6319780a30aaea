cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "4 0.5" "4 0.25" "4 0.75" "4 0.6" "2 0.5" "1 0.5"; do
  set -- $cfg
  T2_CLUSTER=$1 T2_L2_PIN_FRAC=$2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/sw.json") if l.startswith("{")][-1])
print("SWEEP cluster $1 pin $2 value %.0f dec_step_us %.2f ms %.2f" % (d["value"], d["decoder_step_us"], d["ms_per_step"]))
PY
done
