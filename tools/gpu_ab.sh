# Same-box A/B of decoder-kernel variants (run through gpurun):  bash tools/gpu_ab.sh "<name>|<env assignments>" ...
# e.g.  bash tools/gpu_ab.sh "r1|T2B200_LIB=$PWD/tacotron2_b200/libt2b200_r1.so" "new|" "new_s4|T2_STAGES=4"
# Prints value / decoder step and the in-kernel phase profile of CTA 100 for every variant, 2 repetitions each.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for rep in 1 2; do
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}
  env $envs timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/ab_$name$rep.json 2> gpurun_out/ab_$name$rep.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/ab_$name$rep.json") if l.startswith("{")][-1])
    pp = d.get("decoder_phase_profile") or {}
    print("VARIANT $name rep $rep value %.0f dec_step_us %.2f" % (d["value"], d["decoder_step_us"]),
          {k[:8]: v["us_per_step_cta0_60_100"][2] for k, v in pp.items() if isinstance(v, dict) and "us_per_step_cta0_60_100" in v})
except Exception as e:
    print("VARIANT $name rep $rep FAILED", e, open("gpurun_out/ab_$name$rep.err").read()[-600:])
PY
done
done
