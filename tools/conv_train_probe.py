"""Probe: training-mode Encoder / Postnet forward outputs vs the oracle: convs as row-shifted products on gemm_tc
(T2_CONV_TRAIN=cublas, the historical name of "not on conv_tc"; add T2_GEMM=cublas for real cuBLAS) vs the conv_tc engine (fwd)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2
from oracle import tacotron2_oracle as O
from tests.common import keep_mask, rel_err, synth_state_dict

for (B, T) in [(3, 21), (5, 64), (2, 150), (7, 300)]:
    sd = synth_state_dict(seed=33, scale=2.0)
    g = torch.Generator().manual_seed(B * 100 + T)
    x = torch.randn(B, 80, T, generator=g)
    emb = torch.randn(B, 512, T, generator=g)
    lens = torch.sort(torch.randint(max(1, T // 2), T + 1, (B,), generator=g), descending=True)[0]; lens[0] = T
    post_keep = [keep_mask((B, 512, T), 0.5, 7 + i) for i in range(4)] + [keep_mask((B, 80, T), 0.5, 11)]
    ek = keep_mask((3, B, 512, T), 0.5, 12)
    with torch.no_grad():
        ref_post = O.postnet(sd, x, True, post_keep)
        ref_enc = O.encoder(sd, emb, lens, True, ek)
    for mode in ("cublas", "fwd"):
        os.environ["T2_CONV_TRAIN"] = mode
        model = t2.Tacotron2(t2.create_hparams()); model.load_state_dict(sd); model = model.cuda().train()
        xe, ee = x.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
        with t2.dropout_masks(enc=ek, post=post_keep):
            po = model.postnet(xe)
            eo = model.encoder(ee, lens.cuda())
        torch.cuda.synchronize()
        d = (eo.detach().cpu() - ref_enc).abs()
        print("B=%d T=%d %-6s postnet fwd rel err %.2e  encoder fwd rel err %.2e (max abs %.2e at %s)" % (
            B, T, mode, rel_err(po, ref_post), rel_err(eo, ref_enc), float(d.max()), tuple(int(i) for i in (d == d.max()).nonzero()[0])), flush=True)
