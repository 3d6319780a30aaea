set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -2 gpurun_out/bench_v9.err; cat gpurun_out/bench_v9.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-1200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_persistent -c 1 -o gpurun_out/prof_dec_v8 python tools/run_decoder_once.py 100 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 12 -c 1 -o gpurun_out/prof_conv_v8 python tools/run_decoder_once.py 4 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-300
