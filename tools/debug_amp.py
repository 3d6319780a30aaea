import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2
from oracle import tacotron2_oracle as O
from tests.common import rel_err
from tests.test_oracle_golden import grad_inputs, load
from tests.test_gpu_amp import _Round16
gfix = load("grad_train_b4")
sd, text, tl, ol, mels, gt, m = grad_inputs(gfix)
S = 4096.0
smv = float(torch.finfo(torch.float16).min)
model = t2.Tacotron2(t2.create_hparams("fp16_run=True")); model.load_state_dict(sd); model = model.cuda().train()
model.decoder.attention_layer.score_mask_value = smv
model = t2.amp.initialize(model, None, opt_level="O2")
is_bn = lambda k: ".1." in k and ("encoder.convolutions" in k or "postnet.convolutions" in k)
w16 = {k: (v if (not v.dtype.is_floating_point or is_bn(k)) else v.half().float()) for k, v in sd.items()}
R = _Round16.apply
# --- decoder alone: same fp16 memory on both sides
g = torch.Generator().manual_seed(1)
mem = (torch.randn(4, 24, 512, generator=g) * 0.5).half()
d_mel = torch.randn(4, 80, 12, generator=g); d_gate = torch.randn(4, 12, generator=g)
mem_e = mem.clone().cuda().requires_grad_(True)
with t2.dropout_masks(prenet=m["pk"], att=m["ak"], dec=m["dk"]):
    mel, gate, align = model.decoder(mem_e, mels.cuda(), tl.cuda())
    loss = ((mel.float() * d_mel.cuda()).sum() + (gate.float() * d_gate.cuda()).sum()) * S
    loss.backward()
mem_o = mem.float().clone().requires_grad_(True)
sdg = {k: v.clone() for k, v in w16.items()}
mel_o, gate_o, align_o = O.decoder_forward(sdg, mem_o, mels, tl, m["pk"], m["ak"], m["dk"], True, smv)
((R(mel_o) * d_mel).sum() * S + (R(gate_o) * d_gate).sum() * S).backward()
print("decoder alone: mel", rel_err(mel.float(), mel_o), "d_memory engine(fp16) vs oracle fp32:", rel_err(mem_e.grad.float(), mem_o.grad),
      "vs oracle rounded:", rel_err(mem_e.grad.float(), mem_o.grad.half().float()), "max", float(mem_o.grad.abs().max()))
# --- encoder alone: same d_memory
emb = w16["embedding.weight"][text].transpose(1, 2)
dm = torch.randn(4, 24, 512, generator=g)
names = [k for k in w16 if k.startswith("encoder.") and w16[k].dtype.is_floating_point and "running" not in k]
sdg = dict(w16)
for k in names: sdg[k] = w16[k].clone().requires_grad_(True)
mem_o = O.encoder(sdg, emb, tl, True, m["ek"])
(mem_o * dm).sum().backward()
emb_e = emb.half().cuda()
with t2.dropout_masks(enc=m["ek"]):
    mem_e = model.encoder(emb_e, tl.cuda())
    (mem_e.float() * dm.cuda()).sum().backward()
print("encoder alone: memory", rel_err(mem_e.float(), mem_o))
for k in names:
    p = dict(model.named_parameters())[k]
    print("   %-50s %.2e (dtype %s)" % (k, rel_err(p.grad.float(), sdg[k].grad), p.grad.dtype))
