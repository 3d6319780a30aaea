set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s 2>&1 | tail -60 > gpurun_out/bwd_tests.log
cat gpurun_out/bwd_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5
