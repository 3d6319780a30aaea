"""utils.py of the reference, minus the wav/filelist helpers (out of scope): device-agnostic
``get_mask_from_lengths`` (the reference hard-codes torch.cuda.LongTensor, utils.py:8) and ``to_gpu``."""
import torch


def get_mask_from_lengths(lengths, max_len=None):
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device, dtype=torch.long)
    return (ids < lengths.unsqueeze(1)).bool()


def to_gpu(x):
    x = x.contiguous()
    if torch.cuda.is_available():
        x = x.cuda(non_blocking=True)
    return x
