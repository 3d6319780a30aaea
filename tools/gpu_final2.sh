set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -1 gpurun_out/bench_final2.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; tail -1 gpurun_out/bench_ref2.json | cut -c1-400
