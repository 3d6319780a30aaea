"""The three Apex AMP entry points the reference's train.py uses (train.py:173-176, 222-231), over libt2b200's fused
mixed-precision optimizer step -- ``from tacotron2_b200 import amp`` in place of ``from apex import amp``:

    model, optimizer = amp.initialize(model, optimizer, opt_level='O2')          # train.py:174-175
    with amp.scale_loss(loss, optimizer) as scaled_loss: scaled_loss.backward()   # train.py:223-224
    grad_norm = optimizer.step(max_norm=hparams.grad_clip_thresh)                 # replaces train.py:229-236:
                                                                                  #   clip_grad_norm_(amp.master_params(optimizer), ...)
                                                                                  #   optimizer.step()

O2 semantics: parameters are stored in fp16 except BatchNorm (keep_batchnorm_fp32), the model's outputs are cast to fp32
for the loss (cast_model_outputs), fp32 master weights and a dynamic loss scale live in the optimizer.  The kernels
compute fp32-grade from the fp16-rounded weights (split-fp16 tensor-core operands), i.e. never narrower than the reference.
"""
import contextlib

import torch
from torch import nn

from .optim import AmpFusedClipAdam


def initialize(model, optimizer=None, opt_level="O2", loss_scale="dynamic", **unused):
    if opt_level not in ("O0", "O2"):
        raise ValueError("tacotron2_b200.amp: opt_level O2 (the reference's) or O0, got %r" % (opt_level,))
    if opt_level == "O0":
        return (model, optimizer) if optimizer is not None else model
    for mod in model.modules():
        if isinstance(mod, nn.modules.batchnorm._BatchNorm):
            continue
        for p in mod.parameters(recurse=False):
            p.data = p.data.half()
    model.__dict__["_t2_cast_outputs"] = torch.float32
    if hasattr(model, "invalidate_weights"):
        model.invalidate_weights()
    if optimizer is None:
        return model
    g = optimizer.param_groups[0]
    kw = dict(lr=g["lr"], betas=g.get("betas", (0.9, 0.999)), eps=g.get("eps", 1e-8), weight_decay=g.get("weight_decay", 0.0))
    if loss_scale != "dynamic":
        kw.update(init_scale=float(loss_scale), growth_interval=0)
    new_opt = AmpFusedClipAdam([p for grp in optimizer.param_groups for p in grp["params"]], **kw)
    return model, new_opt


def master_params(optimizer):
    return optimizer.master_params()


@contextlib.contextmanager
def scale_loss(loss, optimizer):
    yield optimizer.scale_loss(loss)
