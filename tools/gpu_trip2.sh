set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --decoder-impl persistent --steps 2 --no-cpu-baseline > gpurun_out/bench_v1_prof.json 2> gpurun_out/bench_v1_prof.err; tail -3 gpurun_out/bench_v1_prof.err; cat gpurun_out/bench_v1_prof.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_persistent -c 1 -o gpurun_out/prof_dec_v1 python tools/run_decoder_once.py 100 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_v1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
ls -la gpurun_out
