"""GPU (B200): gradients of the CUDA training path vs torch autograd through the CPU oracle (the oracle's
autograd is pinned against the reference's own autograd in tests/test_oracle_vs_reference.py and by
tests/golden/grad_*.npz).  Same weights, inputs and dropout masks on both sides.

Tolerance: max|g - g_ref| / max|g_ref| < 1e-3 for every parameter gradient and for d_memory."""
import pytest
import torch

import tacotron2_b200 as t2
from oracle import tacotron2_oracle as O
from tests.common import keep_mask, rel_err, synth_state_dict

pytestmark = pytest.mark.gpu
TOL = 1e-3


def decoder_case(B, Te, T, seed, ragged=True):
    g = torch.Generator().manual_seed(seed)
    memory = torch.randn(B, Te, 512, generator=g)
    mels = torch.randn(B, 80, T, generator=g)
    lens = torch.full((B,), Te, dtype=torch.long)
    if ragged and B > 1:
        lens[1:] = torch.randint(max(1, Te // 2), Te + 1, (B - 1,), generator=g)
        lens, _ = torch.sort(lens, descending=True)
    pk = keep_mask((T + 1, 2, B, 256), 0.5, seed + 1)
    ak = keep_mask((T, B, 1024), 0.1, seed + 2)
    dk = keep_mask((T, B, 1024), 0.1, seed + 3)
    d_mel = torch.randn(B, 80, T, generator=g)
    d_gate = torch.randn(B, T, generator=g)
    d_align = torch.randn(B, T, Te, generator=g)
    return memory, mels, lens, pk, ak, dk, d_mel, d_gate, d_align


def oracle_decoder_grads(sd, memory, mels, lens, pk, ak, dk, d_mel, d_gate, d_align, training):
    names = [k for k in sd if k.startswith("decoder.")]
    sdg = dict(sd)
    for k in names:
        sdg[k] = sd[k].clone().requires_grad_(True)
    mem = memory.clone().requires_grad_(True)
    mel, gate, align = O.decoder_forward(sdg, mem, mels, lens, pk.float(), ak.float(), dk.float(), training=training)
    loss = (mel * d_mel).sum() + (gate * d_gate).sum()
    if d_align is not None:
        loss = loss + (align * d_align).sum()
    loss.backward()
    return (mel.detach(), gate.detach(), align.detach()), {k: sdg[k].grad for k in names}, mem.grad


@pytest.mark.parametrize("B,Te,T,training,use_align", [(3, 19, 7, True, False), (5, 40, 12, True, True),
                                                       (4, 150, 9, False, False), (64, 33, 5, True, False)])
def test_decoder_backward_vs_oracle_autograd(B, Te, T, training, use_align):
    sd = synth_state_dict(seed=21, scale=2.0)
    memory, mels, lens, pk, ak, dk, d_mel, d_gate, d_align = decoder_case(B, Te, T, seed=100 + B)
    if not use_align:
        d_align = None
    ref_out, ref_g, ref_dmem = oracle_decoder_grads(sd, memory, mels, lens, pk, ak, dk, d_mel, d_gate, d_align, training)

    model = t2.Tacotron2(t2.create_hparams())
    model.load_state_dict(sd)
    model = model.cuda().train(training)
    dec = model.decoder
    mem = memory.cuda().requires_grad_(True)
    with t2.dropout_masks(prenet=pk, att=ak, dec=dk):
        mel, gate, align = dec(mem, mels.cuda(), lens.cuda())
        loss = (mel * d_mel.cuda()).sum() + (gate * d_gate.cuda()).sum()
        if d_align is not None:
            loss = loss + (align * d_align.cuda()).sum()
        loss.backward()
    torch.cuda.synchronize()
    assert rel_err(mel, ref_out[0]) < TOL and rel_err(gate, ref_out[1]) < TOL and rel_err(align, ref_out[2]) < TOL
    errs = {"d_memory": rel_err(mem.grad, ref_dmem)}
    for k, p in dec.named_parameters():
        assert p.grad is not None, k
        errs[k] = rel_err(p.grad, ref_g["decoder." + k])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    print("decoder backward B=%d Te=%d T=%d: worst %.2e" % (B, Te, T, max(errs.values())))
    assert not bad, bad
