set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -q -s -k "full_size" 2>&1 | grep "worst\|passed\|failed\|Error" | tail
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1500 --csv --log-file gpurun_out/launches_train.csv python tools/train_step_timing.py 64 150 40 2 > gpurun_out/ncu_train.log 2>&1
for k in bwd_gemm_kernel att_bwd_kernel lstm_bwd_row_kernel; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 30 -c 1 -o gpurun_out/ncu2_$k -f python tools/decoder_train_timing.py 64 150 24 > gpurun_out/ncu2_$k.log 2>&1
done
ls gpurun_out/*.ncu-rep
