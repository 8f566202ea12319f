"""Trainer-side loss glue used by bench.py (host code that stays the reference's: trainer.py:912-914 ->
models/loss_fn.py:7-21 `proxy_loss`).  Plain torch on [B, D] / [classes, D] tensors; NOT part of the oracle and not on
the kernel path."""
import torch
import torch.nn.functional as F


def proxy_loss(proxies: torch.Tensor, emb: torch.Tensor, gt: torch.Tensor, scale: float) -> torch.Tensor:
    p = scale * F.normalize(proxies, p=2, dim=-1)
    e = scale * F.normalize(emb, p=2, dim=-1)
    dist = torch.cdist(e, p, p=2) ** 2
    return F.cross_entropy(-dist, gt, reduction="mean")
