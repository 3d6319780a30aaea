"""Fused gradient clipping + Adam for train.py:229-236.

    grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), hparams.grad_clip_thresh)    # train.py:232-233
    optimizer.step()                                                                             # train.py:236

become ONE call, ``grad_norm = optimizer.step(max_norm=hparams.grad_clip_thresh)``: three multi-tensor launches in
libt2b200 (sum of squares, norm / clip coefficient, update) instead of torch's per-operation foreach kernels.  Same
arithmetic as ``clip_grad_norm_`` + ``torch.optim.Adam`` (L2 weight decay, bias correction); state_dict layout is
torch.optim.Adam's (``exp_avg``, ``exp_avg_sq``, ``step``), so checkpoints interchange (train.py:99-113)."""
import ctypes as C

import torch

from . import _capi
from ._engine import bump_weights_generation


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._ws = None

    @torch.no_grad()
    def step(self, closure=None, max_norm=None):
        """Clips the gradients of ALL parameters to ``max_norm`` (None / <= 0: no clipping), then one Adam update per
        param group.  Returns the total gradient norm before clipping (0-dim CUDA tensor; .item() syncs)."""
        if closure is not None:
            raise RuntimeError("FusedClipAdam does not support closures")
        L = _capi.lib()
        groups = [(g, [p for p in g["params"] if p.grad is not None]) for g in self.param_groups]
        all_p = [p for _, ps in groups for p in ps]
        if not all_p:
            return None
        dev = all_p[0].device
        for p in all_p:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("FusedClipAdam: fp32 CUDA parameters and gradients only")
            if not p.is_contiguous():
                raise RuntimeError("FusedClipAdam: parameters must be contiguous")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            st = self.state[p]
            if not st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        norm = torch.zeros((), device=dev, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        total = sum(p.numel() for p in all_p)
        nbytes = L.t2_clip_adam_workspace_bytes(total, len(all_p))
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        first = True
        with torch.cuda.device(dev):
            for g, ps in groups:
                if not ps:
                    continue
                # one group = one launch set; the clip coefficient is computed over ALL parameters (first launch set);
                # further groups get max_norm = 0 after their gradients were pre-scaled -> keep it simple: single group
                if not first:
                    raise RuntimeError("FusedClipAdam: one param group only (train.py uses one)")
                first = False
                n = len(ps)
                arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
                steps = {int(self.state[p]["step"].item()) for p in ps}
                if len(steps) != 1:
                    raise RuntimeError("FusedClipAdam: parameters with different step counts")
                step = steps.pop() + 1
                a = _capi.T2AdamArgs()
                a.n = n
                pa, ga = arr(ps), arr([p.grad for p in ps])
                ma, va = arr([self.state[p]["exp_avg"] for p in ps]), arr([self.state[p]["exp_avg_sq"] for p in ps])
                ne = (C.c_int64 * n)(*[p.numel() for p in ps])
                a.params, a.grads, a.exp_avg, a.exp_avg_sq, a.numel = pa, ga, ma, va, ne
                a.lr, (a.beta1, a.beta2) = float(g["lr"]), [float(b) for b in g["betas"]]
                a.eps, a.weight_decay = float(g["eps"]), float(g["weight_decay"])
                a.max_norm = float(max_norm) if max_norm else 0.0
                a.step = step
                a.grad_norm = norm.data_ptr()
                a.ws, a.ws_bytes = self._ws.data_ptr(), self._ws.numel()
                _capi.check(L.t2_clip_adam_step(C.byref(a), stream))
                for p in ps:
                    self.state[p]["step"] += 1
        bump_weights_generation()       # parameters changed underneath torch's version counters
        return norm
