set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "native or umma" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "stepwise or teacher or encoder" 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "persistent or philox or full_size" 2>&1 | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 600 python bench.py --decoder-impl stepwise --steps 3 --no-cpu-baseline > gpurun_out/bench_stepwise.json 2> gpurun_out/bench_stepwise.err; tail -3 gpurun_out/bench_stepwise.err; cat gpurun_out/bench_stepwise.json
timeout 600 python bench.py --steps 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; tail -3 gpurun_out/bench_auto.err; cat gpurun_out/bench_auto.json
