set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -q 2>&1 | tail -3
timeout 400 python tools/train_step_timing.py 64 150 800 3 2>&1 | tail -4
for k in att_bwd_kernel skinny_nn_kernel lstm_bwd_kernel; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 20 -c 1 -o gpurun_out/ncu_$k -f python tools/decoder_train_timing.py 64 150 24 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
