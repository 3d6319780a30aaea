"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck) over the kernels built on hand-rolled grid
barriers and cross-proxy fences: the persistent decoder (inference + teacher mode with stash), the persistent encoder
BiLSTM, the backward skinny GEMMs, the tensor-core GEMM / conv / weight-gradient engines.
    compute-sanitizer --tool racecheck python tools/sanitize_small.py [infer|train|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2  # noqa: E402
from tests.common import keep_mask, rand_text, synth_state_dict  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
sd = synth_state_dict(5, gate_bias=-10.0, scale=2.0)
model = t2.Tacotron2(t2.create_hparams())
model.load_state_dict(sd)
model = model.cuda()
if what in ("infer", "all"):
    model.eval()
    model.decoder.max_decoder_steps = 4
    with torch.no_grad(), t2.dropout_masks(prenet=keep_mask((4, 2, 2, 256), 0.5, 4)):
        out = model.inference(rand_text(2, 19, 3).cuda())
    torch.cuda.synchronize()
    print("inference ok", [tuple(o.shape) for o in out], flush=True)
if what in ("train", "all"):
    model.train()
    B, Tt, Tm = 2, 11, 3
    g = torch.Generator().manual_seed(0)
    text = rand_text(B, Tt, 1).cuda()
    tl = torch.tensor([Tt, Tt - 3]).cuda()
    ol = torch.tensor([Tm, Tm - 1]).cuda()
    mels = torch.randn(B, 80, Tm, generator=g).cuda()
    gt = torch.zeros(B, Tm).cuda(); gt[0, -1] = 1; gt[1, -2:] = 1
    opt = t2.FusedClipAdam(model.parameters(), lr=1e-3)
    out = model((text, tl, mels, Tt, ol))
    loss = t2.Tacotron2Loss()(out, (mels, gt))
    loss.backward()
    opt.step(max_norm=1.0)
    torch.cuda.synchronize()
    print("train step ok, loss %.5f" % float(loss), flush=True)
