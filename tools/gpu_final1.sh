set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_persistent -c 1 -o gpurun_out/prof_dec_final python tools/run_decoder_once.py 100 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
