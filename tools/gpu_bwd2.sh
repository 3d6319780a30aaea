set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bwd.csv python tools/decoder_train_timing.py 64 150 40 > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log
