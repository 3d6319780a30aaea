set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in A B; do
  if [ $v = A ]; then export T2B200_LIB=$GRAFT_REPO_ROOT/tacotron2_b200/libt2b200_A.so; else unset T2B200_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/ab_$v$rep.json 2> gpurun_out/ab_$v$rep.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/ab_$v$rep.json") if l.startswith("{")][-1])
pp = d["decoder_phase_profile"]
print("VARIANT $v rep $rep value %.0f dec_step_us %.2f" % (d["value"], d["decoder_step_us"]), {k[:6]: v["us_per_step_cta0_60_100"][2] for k, v in pp.items()})
PY
done
done
