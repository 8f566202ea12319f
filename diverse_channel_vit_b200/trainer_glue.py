"""Trainer-side loss glue (host code that stays the reference's own torch code): the CHAMMI main loss
`proxy_loss` (models/loss_fn.py:7-21) with the fixed or learnable temperature (trainer.py:876-883), and the way the
trainer combines it with the module's extra loss (trainer.py:912-914, :986-995).  Plain torch on [B, cls] / [B, D]
tensors; not part of the oracle, not on the kernel path.  Used by bench.py and the tests so that they read like the
reference's training loop."""
import torch
import torch.nn.functional as F


def proxy_loss(proxies: torch.Tensor, emb: torch.Tensor, gt: torch.Tensor, scale) -> torch.Tensor:
    p = scale * F.normalize(proxies, p=2, dim=-1)
    e = scale * F.normalize(emb, p=2, dim=-1)
    dist = torch.cdist(e, p, p=2) ** 2
    return F.cross_entropy(-dist, gt, reduction="mean")


def model_scale(model):
    """trainer.py:876-883: exp(logit_scale) when the temperature is learnable, else the fixed sqrt(1 / T)."""
    if hasattr(model, "logit_scale"):
        return model.logit_scale.exp()
    return model.scale


def training_loss(model, out: torch.Tensor, extra, y: torch.Tensor, has_head: bool, extra_loss_lambda: float = 1.0):
    """trainer.py:986-995 (cross entropy on the classifier head: JUMP-CP / So2Sat) or :912-914 (proxy loss: CHAMMI)."""
    main = F.cross_entropy(out, y) if has_head else proxy_loss(model.proxies, out, y, model_scale(model))
    return main + extra * extra_loss_lambda
