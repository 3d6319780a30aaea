"""Probe: where does the teacher-forced persistent decoder first differ run to run?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2
from tacotron2_b200 import _capi
from tests.common import keep_mask

torch.manual_seed(7)
model = t2.Tacotron2(t2.create_hparams()).cuda().eval()
dec = model.decoder
B, Te, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 150, 24
g = torch.Generator().manual_seed(3)
memory = torch.randn(B, Te, 512, generator=g).cuda()
mels = torch.randn(B, 80, T, generator=g).cuda()
lens = torch.full((B,), Te, dtype=torch.long)
pk = keep_mask((T + 1, 2, B, 256), 0.5, 1)


def first_diff(a, b, dim_t):
    d = (a != b)
    if not bool(d.any()):
        return None
    dims = [i for i in range(a.dim()) if i != dim_t]
    per_t = d.sum(dim=dims)
    t0 = int((per_t > 0).nonzero()[0])
    return t0, int(per_t[t0]), float((a - b).abs().max())


for impl, name in ((_capi.IMPL_PERSISTENT, "persistent"), (_capi.IMPL_STEPWISE, "stepwise")):
    model._t2_engine().impl = impl
    outs = []
    for rep in range(3):
        with torch.no_grad(), t2.dropout_masks(prenet=pk):
            mel, gate, align, sv = dec._teacher_forward(memory, mels, lens.cuda(), impl == _capi.IMPL_PERSISTENT)
        torch.cuda.synchronize()
        st = None
        if sv["stash"] is not None:
            f = sv["stash"].view(torch.float32)
            n = T * B * 4096
            ga, gd = f[:n].view(T, B, 4096), f[n:2 * n].view(T, B, 4096)
            o = 2 * n
            m1 = (T + 1) * B * 1024
            ca, ha, cd, hd = [f[o + i * m1:o + (i + 1) * m1].view(T + 1, B, 1024)[1:] for i in range(4)]
            ctx = f[o + 4 * m1:o + 4 * m1 + (T + 1) * B * 512].view(T + 1, B, 512)[1:]
            st = dict(ga=ga.clone(), ha=ha.clone(), ctx=ctx.clone(), gd=gd.clone(), hd=hd.clone())
        outs.append((mel.clone(), gate.clone(), align.clone(), st))
    for rep in (1, 2):
        print(name, "rep", rep, "mel", first_diff(outs[0][0], outs[rep][0], 1), "gate", first_diff(outs[0][1], outs[rep][1], 1),
              "align", first_diff(outs[0][2], outs[rep][2], 1))
        if outs[0][3] is not None:
            print("   stash (first differing step, count, max abs):", {k: first_diff(outs[0][3][k], outs[rep][3][k], 0) for k in outs[0][3]})
# inference mode
model._t2_engine().impl = _capi.IMPL_PERSISTENT
model.decoder.max_decoder_steps = 24
model.decoder.gate_threshold = 1.0
pk2 = keep_mask((24, 2, B, 256), 0.5, 5)
res = []
for rep in range(3):
    with torch.no_grad(), t2.dropout_masks(prenet=pk2):
        o = dec.inference(memory)
    torch.cuda.synchronize()
    res.append([x.clone() for x in o])
for rep in (1, 2):
    print("inference rep", rep, "mel", first_diff(res[0][0], res[rep][0], 2), "align", first_diff(res[0][2], res[rep][2], 1))
