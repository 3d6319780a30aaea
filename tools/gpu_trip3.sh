set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "SMs", p.multi_processor_count, "smem optin", p.shared_memory_per_block_optin)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "umma or persistent or philox or full_size" 2>&1 | tail -15
for f in 1.0 0.75 0.5; do
  T2_L2_PIN_FRAC=$f timeout 600 python bench.py --decoder-impl persistent --steps 2 --no-cpu-baseline > gpurun_out/bench_v2_pin$f.json 2> gpurun_out/bench_v2_pin$f.err; tail -2 gpurun_out/bench_v2_pin$f.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_v2_pin$f.json") if l.startswith("{")][-1])
print("PIN $f value", d["value"], "dec_step_us", d["decoder_step_us"])
for k, v in d["decoder_phase_profile"].items(): print("  %-26s" % k, v["us_per_step_cta0_60_100"])
PY
done
