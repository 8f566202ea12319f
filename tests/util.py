"""Shared helpers of the parity tests: the golden cases (same definitions as oracle/make_golden.py),
config objects for the drop-in module, CUDA-vs-oracle comparison of one training step."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import dichavit_oracle as O  # noqa: E402
from oracle.make_golden import CHAMMI_MAPPER, Cfg, cases, make_inputs, ref_cfg  # noqa: E402,F401

GOLD = ROOT / "tests" / "golden"


def load_golden(name: str):
    return np.load(GOLD / f"{name}.npz")


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_cuda_model(oc: O.OracleConfig, mapper, weights, device="cuda"):
    from diverse_channel_vit_b200.dichavit import dichavit

    model = dichavit(ref_cfg(oc), mapper=mapper)
    sd = model.state_dict()
    assert set(sd) == set(weights), set(sd) ^ set(weights)
    model.load_state_dict({k: weights[k].clone() for k in sd}, strict=True)
    return model.to(device)


def train_loss_torch(out, extra, y, proxies, scale, has_head, xlam):
    """trainer.py:986-995 / :876-914 loss glue, plain torch on whatever device `out` lives on."""
    if has_head:
        main = F.cross_entropy(out, y)
    else:
        main = O.proxy_loss(proxies, out, y, scale)
    return main + extra * xlam


def cuda_step(model, x, y, chunk, has_head, xlam, indices=None):
    """One fwd+loss+bwd of the drop-in module. `indices`: force the DCS result (host list)."""
    model.train()
    model.zero_grad(set_to_none=True)
    if indices is not None:
        pe = model.feature_extractor.patch_embed
        chan = pe.chunk_channels(chunk, x.device)
        idx = torch.tensor(list(indices), dtype=torch.int32, device=x.device)
        gid = chan[idx.long()].to(torch.int32)
        pe.select_channels = lambda *_a, **_k: (len(indices), idx, gid)  # test hook
    out, extra = model(x, chunk)
    loss = train_loss_torch(out, extra, y, model.proxies, model.scale, has_head, xlam)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    return out, extra, loss, grads
