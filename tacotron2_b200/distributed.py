"""Data parallelism for the two multi-GPU configurations of BASELINE.json.

* Inference (configs 2 / 5) shards the BATCH: utterances are independent, weights are replicated, there
  is NO collective on the data path -- ``shard_rows`` / ``gather_rows`` are the whole protocol.
* Training (config 4) is the reference's scheme (distributed.py:126-173): broadcast rank 0's state when
  the model is wrapped, and after every backward all-reduce the gradients and divide by the world size.
  The one exchange step is a NCCL all-reduce (NVLink 5 / NVSwitch; in-switch reduction when NVLS is
  available).  Differences from the reference, none visible to train.py:
    - gradients are flattened into size-bounded buckets in reverse registration order (postnet ->
      decoder -> encoder, the order in which backward produces them) instead of one 112.8 MB buffer,
      and a bucket's all-reduce is launched asynchronously as soon as its last gradient exists, so the
      exchange overlaps the backward kernels of the earlier layers;
    - the hook is registered once per parameter even when apply_gradient_allreduce is called twice
      (train.py wraps at :79 and :179).
"""
import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world):
    """Contiguous, balanced slice of ``n_rows`` batch rows for ``rank`` (first ``n % world`` ranks get
    one extra row).  Returns (start, stop)."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local, n_rows, group=None):
    """All-gather variable-size row shards (dim 0) back into the full batch order."""
    world = dist.get_world_size(group)
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    max_rows = max(b - a for a, b in sizes)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def _buckets(params, bucket_bytes):
    cur, cur_bytes, cur_dtype = [], 0, None
    for p in params:
        nb = p.numel() * p.element_size()
        if cur and (cur_bytes + nb > bucket_bytes or p.dtype != cur_dtype):
            yield cur
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nb
        cur_dtype = p.dtype
    if cur:
        yield cur


def allreduce_gradients(module, bucket_bytes=32 << 20, group=None):
    """Average ``.grad`` of every parameter over the process group (distributed.py:141-161)."""
    world = dist.get_world_size(group)
    params = [p for p in reversed(list(module.parameters())) if p.requires_grad and p.grad is not None]
    pending = []
    for bucket in _buckets(params, bucket_bytes):
        flat = torch.cat([p.grad.data.reshape(-1) for p in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        pending.append((work, flat, bucket))
    for work, flat, bucket in pending:
        work.wait()
        flat /= world
        off = 0
        for p in bucket:
            n = p.numel()
            p.grad.data.copy_(flat[off:off + n].view_as(p.grad))
            off += n


def apply_gradient_allreduce(module, bucket_bytes=32 << 20):
    """Drop-in for the reference's ``apply_gradient_allreduce`` (distributed.py:126-173)."""
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("apply_gradient_allreduce: torch.distributed is not initialised")
    if getattr(module, "_t2_dp_wrapped", False):
        return module
    for t in module.state_dict().values():                        # distributed.py:132-135
        if torch.is_tensor(t):
            dist.broadcast(t, 0)
    module.needs_reduction = False
    # Buckets in the order backward produces the gradients (postnet -> decoder -> encoder).  A bucket's all-reduce is
    # launched (async, NCCL's own stream) as soon as its last gradient has been accumulated, so the exchange of the
    # postnet / decoder gradients overlaps the backward kernels of the layers before them; the callback at the end of
    # backward only waits and scatters the averaged values back (the reference reduces everything after backward).
    params = [p for p in reversed(list(module.parameters())) if p.requires_grad]
    buckets = list(_buckets(params, bucket_bytes))
    where = {id(p): bi for bi, bucket in enumerate(buckets) for p in bucket}
    state = {"remaining": [len(b_) for b_ in buckets], "pending": {}}

    def launch(bi):
        bucket = buckets[bi]
        flat = torch.cat([p.grad.data.reshape(-1) for p in bucket])
        state["pending"][bi] = (dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat)

    def reduce_now():
        if not module.needs_reduction:
            return
        module.needs_reduction = False
        world = dist.get_world_size()
        late = []
        for bi, bucket in enumerate(buckets):
            if bi not in state["pending"]:
                if all(p.grad is not None for p in bucket):
                    launch(bi)
                else:                                              # parameters that took no part in this backward
                    late += [p for p in bucket if p.grad is not None]
        for bi, (work, flat) in sorted(state["pending"].items()):
            work.wait()
            flat /= world
            off = 0
            for p in buckets[bi]:
                n = p.numel()
                p.grad.data.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        state["pending"] = {}
        if late:
            flat = torch.cat([p.grad.data.reshape(-1) for p in late])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= world
            off = 0
            for p in late:
                n = p.numel()
                p.grad.data.copy_(flat[off:off + n].view_as(p.grad))
                off += n

    def grad_hook(*unused):                                       # distributed.py:163-167
        torch.autograd.Variable._execution_engine.queue_callback(reduce_now)

    def grad_ready(p):
        if not module.needs_reduction:
            return
        bi = where[id(p)]
        state["remaining"][bi] -= 1
        if state["remaining"][bi] == 0:
            launch(bi)

    for p in params:
        p.register_hook(grad_hook)
        p.register_post_accumulate_grad_hook(grad_ready)

    def set_needs_reduction(self, inputs, output):                # distributed.py:169-172
        self.needs_reduction = True
        state["remaining"] = [len(b_) for b_ in buckets]
        state["pending"] = {}

    module.register_forward_hook(set_needs_reduction)
    module._t2_dp_wrapped = True
    return module
