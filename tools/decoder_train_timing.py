"""Times the teacher-forced decoder forward (persistent kernel + stash) and its backward on one GPU.
    python tools/decoder_train_timing.py [B] [T_enc] [T_mel]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Te = int(sys.argv[2]) if len(sys.argv) > 2 else 150
T = int(sys.argv[3]) if len(sys.argv) > 3 else 800
torch.manual_seed(1234)
model = t2.Tacotron2(t2.create_hparams()).cuda().train()
dec = model.decoder
memory = torch.randn(B, Te, 512, device="cuda", requires_grad=True)
mels = torch.randn(B, 80, T, device="cuda")
lens = torch.full((B,), Te, device="cuda", dtype=torch.long)
for it in range(4):
    for p in dec.parameters():
        p.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    mel, gate, align = dec(memory, mels, lens)
    loss = mel.square().mean() + gate.square().mean()
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    print("iter %d: forward %.2f ms, backward %.2f ms  (B=%d T_enc=%d T_mel=%d) loss %.4f" % (
        it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), B, Te, T, float(loss)), flush=True)
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
