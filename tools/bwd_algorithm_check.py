"""Hand-derived reverse-time recurrence of the decoder backward (the decomposition the CUDA kernels in
tacotron2_b200/csrc/decoder_backward.cu implement), checked on CPU against torch autograd through the
oracle's Decoder.forward restatement (model.py:381-416).

    python tools/bwd_algorithm_check.py

Per step t (reverse): carries g_ah, g_ac, g_dh, g_dc, g_ctx (grad wrt the step-t states coming from step
t+1), A0 (grad wrt aw_t from the location conv of step t+1) and C (grad wrt awc_t).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tacotron2_oracle as O  # noqa: E402
from tests import common  # noqa: E402

P = "decoder."


def lstm_bwd(g_h, g_c, gates, c, c_prev):
    i, f, g, o = gates.chunk(4, -1)
    tc = torch.tanh(c)
    d_o = g_h * tc
    d_c = g_c + g_h * o * (1 - tc * tc)
    d_i, d_g, d_f = d_c * g, d_c * i, d_c * c_prev
    dG = torch.cat((d_i * i * (1 - i), d_f * f * (1 - f), d_g * (1 - g * g), d_o * o * (1 - o)), -1)
    return dG, d_c * f


def manual(sd, memory, mels_in, mem_len, pk, ak, dk, d_mel, d_gate, p=0.1):
    """Forward with stash, then the reverse recurrence.  Returns grads dict + d_memory."""
    B, n_mel, T = mels_in.shape
    Te = memory.shape[1]
    w = lambda n: sd[P + n]
    Wia, Wha = w("attention_rnn.weight_ih"), w("attention_rnn.weight_hh")
    ba = w("attention_rnn.bias_ih") + w("attention_rnn.bias_hh")
    Wid, Whd = w("decoder_rnn.weight_ih"), w("decoder_rnn.weight_hh")
    bd = w("decoder_rnn.bias_ih") + w("decoder_rnn.bias_hh")
    Wq, Wm = w("attention_layer.query_layer.linear_layer.weight"), w("attention_layer.memory_layer.linear_layer.weight")
    v = w("attention_layer.v.linear_layer.weight")[0]
    Wloc = w("attention_layer.location_layer.location_conv.conv.weight")     # (32, 2, 31)
    Wld = w("attention_layer.location_layer.location_dense.linear_layer.weight")  # (128, 32)
    Weff = torch.einsum("af,fck->ack", Wld, Wloc)                            # (128, 2, 31)
    Wp = torch.cat((w("linear_projection.linear_layer.weight"), w("gate_layer.linear_layer.weight")), 0)  # (81, 1536)
    bp = torch.cat((w("linear_projection.linear_layer.bias"), w("gate_layer.linear_layer.bias")))
    W1, W2 = w("prenet.layers.0.linear_layer.weight"), w("prenet.layers.1.linear_layer.weight")
    # ---- prenet (hoisted) ----
    frames = torch.cat((memory.new_zeros(1, B, n_mel), mels_in.permute(2, 0, 1)), 0)[:T]   # (T, B, 80)
    x1 = torch.relu(frames @ W1.t()) * pk[:T, 0] * 2
    x2 = torch.relu(x1 @ W2.t()) * pk[:T, 1] * 2
    pm = memory @ Wm.t()
    valid = O.get_mask_from_lengths(mem_len, Te)
    # ---- forward with stash ----
    z = lambda n: memory.new_zeros(B, n)
    HA, CA, HD, CD, CTX = [z(1024)], [z(1024)], [z(1024)], [z(1024)], [z(512)]
    AW, AWC = [z(Te)], [z(Te)]
    GA, GD, Y = [], [], []
    sc = 1.0 / (1.0 - p)

    def act(g):
        i, f, gg, o = g.chunk(4, -1)
        return torch.cat((torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)), -1)

    def pad_cat(aw, awc):
        return torch.nn.functional.pad(torch.stack((aw, awc), 1), (15, 15))   # (B, 2, Te+30)

    def loc(aw, awc):
        pc = pad_cat(aw, awc)
        cols = pc.unfold(2, 31, 1)                                             # (B, 2, Te, 31)
        return torch.einsum("bcjk,ack->bja", cols, Weff), cols

    for t in range(T):
        ga = act(torch.cat((x2[t], CTX[-1]), -1) @ Wia.t() + HA[-1] @ Wha.t() + ba)
        i, f, g, o = ga.chunk(4, -1)
        c = f * CA[-1] + i * g
        h = o * torch.tanh(c) * ak[t] * sc
        GA.append(ga); CA.append(c); HA.append(h)
        pa, _ = loc(AW[-1], AWC[-1])
        s = (h @ Wq.t())[:, None, :] + pa + pm
        e = torch.tanh(s) @ v
        e = e.masked_fill(~valid, -float("inf"))
        aw = torch.softmax(e, 1)
        ctx = torch.einsum("bj,bjc->bc", aw, memory)
        AW.append(aw); AWC.append(AWC[-1] + aw); CTX.append(ctx)
        gd = act(torch.cat((h, ctx), -1) @ Wid.t() + HD[-1] @ Whd.t() + bd)
        i, f, g, o = gd.chunk(4, -1)
        c = f * CD[-1] + i * g
        hd = o * torch.tanh(c) * dk[t] * sc
        GD.append(gd); CD.append(c); HD.append(hd)
        Y.append(torch.cat((hd, ctx), -1) @ Wp.t() + bp)
    Y = torch.stack(Y)                                                          # (T, B, 81)
    # ---- reverse recurrence ----
    dY = torch.cat((d_mel.permute(2, 0, 1), d_gate.t()[:, :, None]), -1)        # (T, B, 81)
    g_ah, g_ac, g_dh, g_dc, g_ctx = z(1024), z(1024), z(1024), z(1024), z(512)
    A0, C = z(Te), z(Te)
    DGA, DGD, DQ, DCTX, GS, DX2 = [None] * T, [None] * T, [None] * T, [None] * T, [None] * T, [None] * T
    dv = torch.zeros(128)
    for t in reversed(range(T)):
        g_dhc = dY[t] @ Wp
        g_dh = g_dh + g_dhc[:, :1024]
        g_ctx = g_ctx + g_dhc[:, 1024:]
        dGd, g_dc = lstm_bwd(g_dh * dk[t] * sc, g_dc, GD[t], CD[t + 1], CD[t])
        DGD[t] = dGd
        r = dGd @ Wid
        g_ah = g_ah + r[:, :1024]
        g_ctx = g_ctx + r[:, 1024:]
        g_dh = dGd @ Whd
        # attention
        DCTX[t] = g_ctx
        aw = AW[t + 1]
        g_aw = A0 + C + torch.einsum("bjc,bc->bj", memory, g_ctx)
        g_e = aw * (g_aw - (aw * g_aw).sum(1, keepdim=True))
        pa, cols = loc(AW[t], AWC[t])
        th = torch.tanh((HA[t + 1] @ Wq.t())[:, None, :] + pa + pm)
        g_s = g_e[:, :, None] * v * (1 - th * th)                                # (B, Te, 128)
        dv = dv + torch.einsum("bj,bja->a", g_e, th)
        GS[t] = (g_s, cols)
        g_q = g_s.sum(1)
        DQ[t] = g_q
        U = torch.einsum("bja,ack->bjck", g_s, Weff)                             # (B, Te, 2, 31)
        g_cat = memory.new_zeros(B, 2, Te + 30)
        for k in range(31):
            g_cat[:, :, k:k + Te] += U[:, :, :, k].transpose(1, 2)
        g_cat = g_cat[:, :, 15:15 + Te]
        A0, C = g_cat[:, 0], g_cat[:, 1] + C
        g_ah = g_ah + g_q @ Wq
        dGa, g_ac = lstm_bwd(g_ah * ak[t] * sc, g_ac, GA[t], CA[t + 1], CA[t])
        DGA[t] = dGa
        r = dGa @ Wia
        DX2[t] = r[:, :256]
        g_ctx = r[:, 256:]
        g_ah = dGa @ Wha
    # ---- time-batched weight gradients ----
    S = lambda xs: torch.stack(xs).reshape(-1, xs[0].shape[-1])
    DGA_, DGD_ = S(DGA), S(DGD)
    grads = {}
    XA = torch.cat((x2.reshape(-1, 256), S(CTX[:T])), -1)
    grads["attention_rnn.weight_ih"] = DGA_.t() @ XA
    grads["attention_rnn.weight_hh"] = DGA_.t() @ S(HA[:T])
    grads["attention_rnn.bias_ih"] = grads["attention_rnn.bias_hh"] = DGA_.sum(0)
    XD = torch.cat((S(HA[1:]), S(CTX[1:])), -1)
    grads["decoder_rnn.weight_ih"] = DGD_.t() @ XD
    grads["decoder_rnn.weight_hh"] = DGD_.t() @ S(HD[:T])
    grads["decoder_rnn.bias_ih"] = grads["decoder_rnn.bias_hh"] = DGD_.sum(0)
    dWp = dY.reshape(-1, 81).t() @ torch.cat((S(HD[1:]), S(CTX[1:])), -1)
    grads["linear_projection.linear_layer.weight"] = dWp[:80]
    grads["gate_layer.linear_layer.weight"] = dWp[80:]
    grads["linear_projection.linear_layer.bias"] = dY.reshape(-1, 81).sum(0)[:80]
    grads["gate_layer.linear_layer.bias"] = dY.reshape(-1, 81).sum(0)[80:]
    grads["attention_layer.query_layer.linear_layer.weight"] = S(DQ).t() @ S(HA[1:])
    grads["attention_layer.v.linear_layer.weight"] = dv[None]
    gs = torch.stack([g for g, _ in GS])                                        # (T, B, Te, 128)
    cols = torch.stack([c for _, c in GS])                                      # (T, B, 2, Te, 31)
    dWeff = torch.einsum("tbja,tbcjk->ack", gs, cols)
    grads["attention_layer.location_layer.location_dense.linear_layer.weight"] = torch.einsum("ack,fck->af", dWeff, Wloc)
    grads["attention_layer.location_layer.location_conv.conv.weight"] = torch.einsum("af,ack->fck", Wld, dWeff)
    d_pm = gs.sum(0)                                                            # (B, Te, 128)
    grads["attention_layer.memory_layer.linear_layer.weight"] = d_pm.reshape(-1, 128).t() @ memory.reshape(-1, 512)
    d_memory = d_pm @ Wm + torch.einsum("tbj,tbc->bjc", torch.stack(AW[1:]), torch.stack(DCTX))
    # prenet
    dz2 = torch.stack(DX2) * 2 * (x2 > 0)
    grads["prenet.layers.1.linear_layer.weight"] = dz2.reshape(-1, 256).t() @ x1.reshape(-1, 256)
    dz1 = (dz2 @ W2) * 2 * (x1 > 0)
    grads["prenet.layers.0.linear_layer.weight"] = dz1.reshape(-1, 256).t() @ frames.reshape(-1, 80)
    return Y, grads, d_memory


def main():
    torch.manual_seed(0)
    B, Te, T = 3, 19, 7
    sd = {k: v.double() if v.is_floating_point() else v for k, v in common.synth_state_dict(seed=5, scale=2.0).items()}
    memory = torch.randn(B, Te, 512, dtype=torch.double)
    mels = torch.randn(B, 80, T, dtype=torch.double)
    mem_len = torch.tensor([19, 14, 9])
    g = torch.Generator().manual_seed(1)
    pk = (torch.rand(T + 1, 2, B, 256, generator=g) < 0.5).double()
    ak = (torch.rand(T, B, 1024, generator=g) < 0.9).double()
    dk = (torch.rand(T, B, 1024, generator=g) < 0.9).double()
    d_mel = torch.randn(B, 80, T, dtype=torch.double)
    d_gate = torch.randn(B, T, dtype=torch.double)
    # autograd through the oracle
    names = [k for k in sd if k.startswith(P)]
    sdg = dict(sd)
    for k in names:
        sdg[k] = sd[k].clone().requires_grad_(True)
    mem_g = memory.clone().requires_grad_(True)
    mel_o, gate_o, _ = O.decoder_forward(sdg, mem_g, mels, mem_len, pk, ak, dk, training=True)
    loss = (mel_o * d_mel).sum() + (gate_o * d_gate).sum()
    loss.backward()
    Y, grads, d_memory = manual(sd, memory, mels, mem_len, pk, ak, dk, d_mel, d_gate)
    print("fwd mel err", (Y[:, :, :80].permute(1, 2, 0) - mel_o).abs().max().item())
    worst = 0.0
    for k in names:
        ref = sdg[k].grad
        got = grads[k[len(P):]]
        err = (ref - got).abs().max().item() / (ref.abs().max().item() + 1e-30)
        worst = max(worst, err)
        print("%-70s rel err %.2e  (|ref| %.3e)" % (k, err, ref.abs().max().item()))
    err = (mem_g.grad - d_memory).abs().max().item() / mem_g.grad.abs().max().item()
    print("d_memory rel err %.2e" % err)
    worst = max(worst, err)
    assert worst < 1e-9, worst
    print("OK")


if __name__ == "__main__":
    main()
