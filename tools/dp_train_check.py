"""Data-parallel training step on N GPUs (config 4 of BASELINE.json), one process per GPU:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py [B] [T_text] [T_mel]
Checks that the gradients after apply_gradient_allreduce are identical on every rank and equal the mean of the
ranks' local gradients (same dropout masks in both passes), then times steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacotron2_b200 as t2  # noqa: E402
from tacotron2_b200.distributed import apply_gradient_allreduce  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Tt = int(sys.argv[2]) if len(sys.argv) > 2 else 40
Tm = int(sys.argv[3]) if len(sys.argv) > 3 else 60
hp = t2.create_hparams()
torch.manual_seed(1234 + rank)                     # different initial weights: the wrapper broadcasts rank 0's
model = t2.Tacotron2(hp).cuda().train()
crit = t2.Tacotron2Loss()
g = torch.Generator().manual_seed(100 + rank)      # different data per rank
text = torch.randint(0, 148, (B, Tt), generator=g).cuda()
tl = torch.sort(torch.randint(Tt // 2, Tt + 1, (B,), generator=g), descending=True)[0]; tl[0] = Tt
ol = torch.randint(Tm // 2, Tm + 1, (B,), generator=g); ol[0] = Tm
mels = torch.randn(B, 80, Tm, generator=g)
gt = torch.zeros(B, Tm)
for i, n in enumerate(ol.tolist()):
    mels[i, :, n:] = 0; gt[i, n - 1:] = 1
tl, ol, mels, gt = tl.cuda(), ol.cuda(), mels.cuda(), gt.cuda()
x = (text, tl, mels, int(tl.max()), ol)
km = lambda shape, p, s: (torch.rand(shape, generator=torch.Generator().manual_seed(s + 17 * rank)) >= p).to(torch.uint8)
masks = dict(prenet=km((Tm + 1, 2, B, 256), 0.5, 1), att=km((Tm, B, 1024), 0.1, 2), dec=km((Tm, B, 1024), 0.1, 3),
             enc=km((3, B, 512, Tt), 0.5, 4), post=[km((B, 512, Tm), 0.5, 5 + i) for i in range(4)] + [km((B, 80, Tm), 0.5, 9)])

apply_gradient_allreduce(model)
bn_state = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}


def step(reduce):
    model.zero_grad()
    model.needs_reduction = False
    with t2.dropout_masks(**masks):
        if reduce:
            out = model(x)                          # forward hook arms the reduction
        else:
            out = model.forward(x)                  # bypasses the hooks: purely local gradients
        loss = crit(out, (mels, gt))
        loss.backward()
    torch.cuda.synchronize()
    model.load_state_dict(bn_state, strict=False)  # same BatchNorm running stats for both passes
    return [p.grad.detach().clone() for p in model.parameters()], float(loss)


local, l0 = step(False)
reduced, l1 = step(True)
worst_avg, worst_rank, worst_name = 0.0, 0.0, ""
gmax = max(float(t.abs().max()) for t in local)
for (name, _), lg, rg in zip(model.named_parameters(), local, reduced):
    mean = lg.clone()
    dist.all_reduce(mean)
    mean /= world
    # tolerance 1e-5 of the parameter's own gradient + 1e-6 of the largest gradient (conv biases in front of a
    # training-mode BatchNorm hold rounding noise only, ~1e-8)
    den = float(mean.abs().max()) + 0.1 * gmax
    e = float((rg - mean).abs().max()) / den
    if e > worst_avg:
        worst_avg, worst_name = e, "%s |mean| %.2e |local| %.2e |reduced| %.2e" % (name, float(mean.abs().max()), float(lg.abs().max()), float(rg.abs().max()))
    other = rg.clone()
    dist.broadcast(other, 0)
    worst_rank = max(worst_rank, float((rg - other).abs().max()) / den)
print("rank %d/%d: loss %.5f; reduced grads vs mean of local grads: %.2e (%s); vs rank 0: %.2e" % (rank, world, l1, worst_avg, worst_name, worst_rank), flush=True)
assert worst_avg < 1e-5 and worst_rank == 0.0
# timing
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.zero_grad()
    out = model(x)
    crit(out, (mels, gt)).backward()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("step %d: %.1f ms (max over ranks) -> %.0f frames/s over %d GPUs" % (it, float(ms), world * B * Tm / float(ms) * 1e3, world), flush=True)
dist.destroy_process_group()
