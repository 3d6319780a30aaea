"""Times one teacher-forced training step (config 3 of BASELINE.json: B=64, T_text=150, T_mel=800, fp32 here) through the
public API: Tacotron2.forward + Tacotron2Loss + backward + grad-norm clip + Adam (train.py:209-236).
    python tools/train_step_timing.py [B] [T_text] [T_mel] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Tt = int(sys.argv[2]) if len(sys.argv) > 2 else 150
Tm = int(sys.argv[3]) if len(sys.argv) > 3 else 800
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
torch.manual_seed(1234)
hp = t2.create_hparams()
model = t2.Tacotron2(hp).cuda().train()
fused = os.environ.get("T2_TORCH_ADAM") is None
opt = (t2.FusedClipAdam if fused else torch.optim.Adam)(model.parameters(), lr=hp.learning_rate, weight_decay=hp.weight_decay)
crit = t2.Tacotron2Loss()
g = torch.Generator().manual_seed(0)
text = torch.randint(0, 148, (B, Tt), generator=g).cuda()
tl = torch.sort(torch.randint(Tt // 2, Tt + 1, (B,), generator=g), descending=True)[0]
tl[0] = Tt
ol = torch.randint(Tm // 2, Tm + 1, (B,), generator=g)
ol[0] = Tm
mels = torch.randn(B, 80, Tm, generator=g)
gt = torch.zeros(B, Tm)
for i, n in enumerate(ol.tolist()):
    mels[i, :, n:] = 0
    gt[i, n - 1:] = 1
tl, ol, mels, gt = tl.cuda(), ol.cuda(), mels.cuda(), gt.cuda()
x = (text, tl, mels, int(tl.max()), ol)
for it in range(iters):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    model.zero_grad()
    out = model(x)
    loss = crit(out, (mels, gt))
    ev[1].record()
    loss.backward()
    ev[2].record()
    if fused:
        gn = opt.step(max_norm=hp.grad_clip_thresh)
    else:
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), hp.grad_clip_thresh)
        opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    fw, bw, up = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    print("iter %d: forward+loss %.1f ms, backward %.1f ms, clip+Adam %.1f ms, total %.1f ms -> %.0f frames/s; loss %.4f grad norm %.3f" % (
        it, fw, bw, up, fw + bw + up, B * Tm / (fw + bw + up) * 1e3, loss.item(), float(gn)), flush=True)
print("max mem GB %.2f" % (torch.cuda.max_memory_allocated() / 2**30))
