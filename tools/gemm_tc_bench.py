"""Times gemm_tc.cu (the training path's tensor-core GEMM) on the shapes the backward pass uses, next to torch.matmul fp32
(cuBLAS SIMT sgemm, TF32 off) for context.   python tools/gemm_tc_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacotron2_b200 import _capi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
L = _capi.selftest_lib()
SHAPES = [  # name, ta, tb, M, N, K
    ("conv fwd tap (postnet 512->512)", 0, 1, 51456, 512, 512),
    ("q = ha . Wq^T", 0, 1, 51200, 128, 1024),
    ("gproj = dy . [Wp;Wg]", 0, 0, 51200, 1536, 81),
    ("dW_proj = dy^T . hd", 1, 0, 80, 1024, 51200),
    ("dW_query = dq^T . ha", 1, 0, 128, 1024, 51200),
    ("dWeff chunk = gs^T . cols", 1, 0, 128, 64, 1 << 20),
    ("prenet dz1 = dz2 . W2", 0, 0, 51264, 256, 256),
    ("BiLSTM dW_ih = dG^T . x", 1, 0, 1024, 512, 9600),
    ("BiLSTM dx = dG . W_ih", 0, 0, 9600, 512, 1024),
    ("d_memory = dpm . Wm", 0, 0, 9600, 512, 128),
]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, ta, tb, M, N, K in SHAPES:
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((N, K) if tb else (K, N), device="cuda")
    Cc = torch.empty(M, N, device="cuda")

    def ours():
        _capi.check_selftest(L.t2_selftest_gemm_tc(ta, tb, M, N, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], Cc.data_ptr(), N,
                                                   0.0, 1, 0, 0, 0, st))

    def ref():
        return (A.t() if ta else A) @ (B.t() if tb else B)
    res = []
    for fn in (ours, ref):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); e1.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    err = float((Cc - ref()).abs().max() / ref().abs().max())
    print("%-34s M=%6d N=%5d K=%8d  gemm_tc %8.3f ms (%6.1f TFLOP/s)   cuBLAS fp32 %8.3f ms   rel diff %.1e" % (
        name, M, N, K, res[0], 2.0 * M * N * K / res[0] / 1e9, res[1], err), flush=True)
