set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s 2>&1 | tail -40 > gpurun_out/bwd_tests.log
cat gpurun_out/bwd_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "teacher or golden" 2>&1 | tail -5
timeout 600 python tools/decoder_train_timing.py 2>&1 | tail -8 | tee gpurun_out/dec_train_timing.log
