# compute-sanitizer over the small-shape driver (run through gpurun); logs -> gpurun_out/sanitize_<tool>_<what>.log
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  for what in infer train; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_small.py $what > gpurun_out/sanitize_${tool}_${what}.log 2>&1
    echo "== $tool $what: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' gpurun_out/sanitize_${tool}_${what}.log | tr '\n' ' ')"
  done
done
