"""CPU oracle for the DiChaViT training hot path -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (fp32 or fp64, CPU or any device) functional restatement of the
reference algorithm, used exclusively as the checker by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
The product path (diverse_channel_vit_b200/) never imports this file.

Parity pinning: the reference ships no tests or golden vectors.  This
restatement is pinned by executing the *imported, unmodified reference module* on identical
inputs / weights / RNG state in the build container (oracle/make_golden.py, which asserts
agreement and writes tests/golden/*.npz); tests/test_oracle_golden.py re-checks the oracle
against those committed vectors wherever the suite runs.

Each function cites the reference file:line it follows (paths relative to the reference
repository root).
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# configuration
# ---------------------------------------------------------------------------------------------

_SIZES = {  # models/dichavit.py:676-745 (depth 12, mlp_ratio 4, qkv_bias, LN eps 1e-6)
    "tiny": (192, 3),
    "small": (384, 6),
    "distill": (384, 6),
    "base": (768, 12),
}


@dataclass
class OracleConfig:
    """The keys of configs/model/dichavit.yaml the hot path reads (+ trainer-filled ones)."""

    pretrained_model_name: str = "small"
    img_size: int = 224
    patch_size: int = 16
    in_channel_names: Sequence[str] = field(default_factory=lambda: ["c0", "c1", "c2"])
    num_classes: int = 14
    temperature: float = 0.11111
    enable_sample: bool = False
    hcs_sampling: str = "none"
    hcs_sampling_temp: float = 0.1
    proxy_loss_lambda: float = 0.0
    ortho_loss_v1_lambda: float = 0.0
    gamma_s: float = 1.0
    gamma_d: float = 0.5
    reverse_pos_pairs: bool = False
    use_square: bool = False
    depth: int = 12

    @property
    def dim(self) -> int:
        return _SIZES[self.pretrained_model_name][0]

    @property
    def heads(self) -> int:
        return _SIZES[self.pretrained_model_name][1]

    @property
    def n_patches(self) -> int:
        return (self.img_size // self.patch_size) ** 2

    @property
    def total_channels(self) -> int:
        return len(self.in_channel_names)


def param_shapes(cfg: OracleConfig, has_head: bool) -> Dict[str, Tuple[int, ...]]:
    """state_dict names/shapes of the reference DiChaViT (models/dichavit.py:42-96,423-507,749-812)."""
    D, P, C, N = cfg.dim, cfg.patch_size, cfg.total_channels, cfg.n_patches
    s: Dict[str, Tuple[int, ...]] = {}
    fe = "feature_extractor."
    s["proxies"] = (cfg.num_classes, D)
    s[fe + "cls_token"] = (1, 1, D)
    s[fe + "pos_embed"] = (1, N + 1, D)
    if cfg.proxy_loss_lambda > 0:
        s[fe + "patch_embed.channel_emb_proxies"] = (C, D)
    s[fe + "patch_embed.proj.weight"] = (D, 1, 1, P, P)
    s[fe + "patch_embed.proj.bias"] = (D,)
    s[fe + "patch_embed.channel_embed.weight"] = (C, D)
    for i in range(cfg.depth):
        b = f"{fe}blocks.{i}."
        s[b + "norm1.weight"] = (D,)
        s[b + "norm1.bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D)
        s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D)
        s[b + "attn.proj.bias"] = (D,)
        s[b + "norm2.weight"] = (D,)
        s[b + "norm2.bias"] = (D,)
        s[b + "mlp.fc1.weight"] = (4 * D, D)
        s[b + "mlp.fc1.bias"] = (4 * D,)
        s[b + "mlp.fc2.weight"] = (D, 4 * D)
        s[b + "mlp.fc2.bias"] = (D,)
    s[fe + "norm.weight"] = (D,)
    s[fe + "norm.bias"] = (D,)
    if has_head:
        s["classifer_head.weight"] = (cfg.num_classes, D)
        s["classifer_head.bias"] = (cfg.num_classes,)
    s["adaptive_interface.0"] = (cfg.num_classes, D)  # alias of proxies (dichavit.py:812)
    return s


def make_weights(cfg: OracleConfig, has_head: bool, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, non-degenerate ("trained-looking") weights shared by reference, oracle and
    the CUDA module in parity tests.  CPU mt19937 stream -> identical on every machine."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg, has_head).items():
        if name == "adaptive_interface.0":
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float64)
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("norm.weight"):
            t = 1.0 + 0.1 * r
        elif name.endswith(".bias"):
            t = 0.05 * r
        elif name.endswith("proj.weight") and "patch_embed" in name:
            t = r / math.sqrt(shape[-1] * shape[-2]) * 0.6
        elif name.endswith("channel_embed.weight"):
            t = r / math.sqrt(shape[-1]) + 0.05
        elif name.endswith("channel_emb_proxies") or name == "proxies":
            t = r / 8.0
        elif name.endswith("pos_embed") or name.endswith("cls_token"):
            t = 0.05 * r
        elif name.endswith(".weight"):
            t = r * (1.0 / math.sqrt(shape[-1])) * 0.7
        else:
            t = 0.02 * r
        out[name] = t.to(dtype)
    out["adaptive_interface.0"] = out["proxies"]
    return out


# ---------------------------------------------------------------------------------------------
# DCS -- Diverse Channel Sampling (models/dichavit.py:127-216, mode "lowest_cosine_prob")
# ---------------------------------------------------------------------------------------------

def dcs_probabilities(channel_embed: torch.Tensor, anchor: int, temp: float) -> torch.Tensor:
    """dichavit.py:169-174,194-196: p = softmax((1 - cos(e_anchor, e_j)) / temp)."""
    e = F.normalize(channel_embed, p=2, dim=-1)
    cos = torch.einsum("c d, e d -> c e", e, e)[anchor]
    return F.softmax((1 - cos) / temp, dim=-1)


def dcs_select(channel_embed: torch.Tensor, temp: float, mode: str = "lowest_cosine_prob") -> Tuple[int, int, List[int]]:
    """Consumes RNG exactly like the reference: python random.randint(1,C) (dichavit.py:128),
    random.randint(0,C-1) (:154), then torch.multinomial on the default generator of
    channel_embed's device (:199).  Returns (C', anchor, indices into the chunk's channel list)."""
    C = channel_embed.shape[0]
    c_new = random.randint(1, C)
    if mode in ("none", None):
        idx = random.sample(list(range(C)), k=c_new)  # dichavit.py:131 (sampling positions == sampling ids)
        return c_new, -1, idx
    anchor = random.randint(0, C - 1)
    with torch.no_grad():
        if mode == "lowest_cosine_prob":
            prob = dcs_probabilities(channel_embed, anchor, temp)
            idx = torch.multinomial(prob, c_new, replacement=False).cpu().numpy().tolist()
        elif mode in ("lowest_cosine", "highest_cosine"):
            e = F.normalize(channel_embed, p=2, dim=-1)
            cos = (e @ e.t())[anchor]
            idx = torch.topk(cos, k=c_new, largest=(mode == "highest_cosine"))[1].cpu().numpy().tolist()
        else:
            raise ValueError(f"Invalid hcs_sampling: '{mode}'")
    if anchor not in idx:  # dichavit.py:201-202
        idx[-1] = anchor
    return c_new, anchor, idx


def leave_one_out_channel_tokens(weight: torch.Tensor, mapper: Dict[str, Sequence[int]], chunk_name: str,
                                 training_chunks: str, new_channel_init: str,
                                 channel_map: Optional[Dict[int, int]] = None) -> torch.Tensor:
    """dichavit.py:219-374 (eval only, `training_chunks` given): channel tokens [C_in, D] of the chunk where every
    channel that was not seen in training gets a token synthesised from training channels.  Modes that need the
    `bank` attribute (dynamic_input_corr_*) are not restated: nothing in the reference ever sets it."""
    training_channels = [c for ch in training_chunks.split("_") for c in mapper[ch]]
    chs_not_seen = [c for c in training_channels if c not in mapper[chunk_name]]
    ch_banks = chs_not_seen if "not_in_chunk" in new_channel_init else training_channels
    rows, cur = [], 0
    for c in mapper[chunk_name]:
        if c in training_channels:
            rows.append(weight[c][None])
            continue
        n = len(ch_banks)
        if new_channel_init in ("avg_2", "avg_2_not_in_chunk"):
            param = weight[[ch_banks[cur], ch_banks[(cur + 1) % n]]].mean(dim=0, keepdim=True)
        elif new_channel_init in ("avg_3", "avg_3_not_in_chunk"):
            param = weight[[ch_banks[cur], ch_banks[(cur + 1) % n], ch_banks[(cur + 2) % n]]].mean(dim=0, keepdim=True)
        elif new_channel_init == "replicate":
            param = weight[ch_banks[cur]][None]
        elif new_channel_init == "zero":
            param = torch.zeros_like(weight[0])[None]
        elif new_channel_init == "random":
            param = weight[c][None]
        elif new_channel_init == "fixed_input_corr":
            if channel_map is None:
                raise ValueError("provide a channel_map (dict)!")
            param = weight[channel_map[c]][None]
        else:
            raise ValueError(f"Invalid new_channel_init: '{new_channel_init}'")
        cur = (cur + 1) % n
        rows.append(param)
    return torch.cat(rows, dim=0)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------

def proxy_loss(proxies: torch.Tensor, emb: torch.Tensor, gt: torch.Tensor, scale: float) -> torch.Tensor:
    """models/loss_fn.py:7-21 + utils.py:461-465: CE over -||s*e_hat - s*p_hat||^2."""
    p = scale * F.normalize(proxies, p=2, dim=-1)
    e = scale * F.normalize(emb, p=2, dim=-1)
    dist = torch.cdist(e, p, p=2) ** 2
    return F.cross_entropy(-dist, gt, reduction="mean")


def tdl_loss(features: torch.Tensor, labels: torch.Tensor, gamma_s: float, gamma_d: float,
             reverse_pos_pairs: bool, use_square: bool) -> torch.Tensor:
    """models/loss_fn.py:24-59 (ortho_proj_loss_fn_v2), Gram-matrix form exactly as the reference."""
    f = F.normalize(features, p=2, dim=-1)
    lab = labels[None, :, None]
    mask = torch.eq(lab, lab.transpose(-2, -1))
    eye = torch.eye(mask.shape[-2], mask.shape[-1], dtype=torch.bool, device=f.device).unsqueeze(0)
    mask_pos = mask.masked_fill(eye, False).to(f.dtype)
    mask_neg = (~mask).to(f.dtype)
    dot = torch.matmul(f, f.transpose(-2, -1))
    pos_n = mask_pos.sum(dim=(-2, -1)) + 1e-6
    neg_n = mask_neg.sum(dim=(-2, -1)) + 1e-6
    pos = (mask_pos * dot).sum(dim=(-2, -1)) / pos_n
    neg = (mask_neg * dot).sum(dim=(-2, -1)) / neg_n
    if use_square:
        neg = neg ** 2
    if reverse_pos_pairs:
        if use_square:
            pos = pos ** 2
        loss = gamma_s * pos + gamma_d * neg
    else:
        loss = gamma_s * (1.0 - pos) + gamma_d * neg
    return loss.mean()


# ---------------------------------------------------------------------------------------------
# token preparation
# ---------------------------------------------------------------------------------------------

def patch_project(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, P: int) -> torch.Tensor:
    """dichavit.py:77-82,377: Conv3d(1,D,(1,P,P),stride (1,P,P)) == per-channel patch GEMM.
    x [B,C,H,W] -> [B,C,N,D], token order (c, hp, wp), k = ph*P + pw."""
    B, C, H, W = x.shape
    hp, wp = H // P, W // P
    patches = x.reshape(B, C, hp, P, wp, P).permute(0, 1, 2, 4, 3, 5).reshape(B, C, hp * wp, P * P)
    return patches @ w.reshape(w.shape[0], P * P).t() + b


def interpolate_pos(pos_embed: torch.Tensor, n_tokens_wo_cls: int, w: int, h: int, P: int, nc: int) -> torch.Tensor:
    """dichavit.py:518-552: raw pos_embed iff C'==1 (and w==h), else bicubic resample with
    scale (w//P + 0.1)/sqrt(N), tiled over the nc channels; CLS keeps pos_embed[:, :1]."""
    N = pos_embed.shape[1] - 1
    if n_tokens_wo_cls == N and w == h:
        return pos_embed
    dim = pos_embed.shape[-1]
    cls_pos, patch_pos = pos_embed[:, :1], pos_embed[:, 1:]
    w0, h0 = w // P + 0.1, h // P + 0.1
    s = int(math.sqrt(N))
    patch_pos = F.interpolate(
        patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
        scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)),
        mode="bicubic",
    )
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).view(1, 1, -1, dim)
    patch_pos = patch_pos.expand(1, nc, -1, dim).reshape(1, -1, dim)
    return torch.cat((cls_pos, patch_pos), dim=1)


# ---------------------------------------------------------------------------------------------
# transformer
# ---------------------------------------------------------------------------------------------

def block_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, heads: int) -> torch.Tensor:
    """models/vit.py:383-399 (Block), :121-144 (Attention), :76-82 (Mlp). Pre-LN, eps 1e-6."""
    B, L, D = x.shape
    hd = D // heads
    u = F.layer_norm(x, (D,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], eps=1e-6)
    qkv = F.linear(u, p[prefix + "attn.qkv.weight"], p[prefix + "attn.qkv.bias"])
    qkv = qkv.reshape(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * hd ** -0.5
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, L, D)
    x = x + F.linear(o, p[prefix + "attn.proj.weight"], p[prefix + "attn.proj.bias"])
    v2 = F.layer_norm(x, (D,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], eps=1e-6)
    h = F.gelu(F.linear(v2, p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"]))
    x = x + F.linear(h, p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])
    return x


@dataclass
class OracleOutput:
    out: torch.Tensor                 # logits [B,cls] (head present) or CLS feature [B,D]
    extra_loss: torch.Tensor          # lambda_tdl*TDL + lambda_cdl*CDL (dichavit.py:406-408)
    tdl: torch.Tensor
    cdl: torch.Tensor
    tokens: torch.Tensor              # [B,L,D] entering block 0
    indices: List[int]                # positions (in the chunk's channel list) of the channels used
    blocks: Optional[List[torch.Tensor]] = None


def forward(x: torch.Tensor, p: Dict[str, torch.Tensor], cfg: OracleConfig, channels: Sequence[int],
            training: bool, has_head: bool, indices: Optional[Sequence[int]] = None,
            keep_blocks: bool = False, channel_embed_override: Optional[torch.Tensor] = None) -> OracleOutput:
    """DiChaViT.forward (dichavit.py:844-861) -> ChannelVisionTransformer.forward (:631-652) ->
    prepare_tokens (:554-629) -> PatchEmbedPerChannel.forward (:110-417).

    `channels` = mapper[chunk_name] (global channel ids of x's channels).  `indices`: if given,
    the DCS result to use; if None and sampling is enabled in training, it is drawn with
    dcs_select (consuming RNG like the reference)."""
    fe = "feature_extractor."
    D, P = cfg.dim, cfg.patch_size
    B, C, H, W = x.shape
    chan_ids = torch.tensor(list(channels), device=x.device)
    channel_embed = p[fe + "patch_embed.channel_embed.weight"][chan_ids]  # dichavit.py:122
    if channel_embed_override is not None:  # eval-time leave-one-out synthesis, dichavit.py:219-374
        channel_embed = channel_embed_override
    cur_channels = list(channels)
    if training and cfg.enable_sample:
        if indices is None:
            _, _, indices = dcs_select(channel_embed.detach(), cfg.hcs_sampling_temp, cfg.hcs_sampling)
    elif indices is None:
        indices = list(range(C))
    indices = list(indices)
    if indices != list(range(C)):
        cur_channels = [cur_channels[i] for i in indices]  # dichavit.py:203,208-212
        x = x[:, indices]
        channel_embed = channel_embed[indices]
    Cn = len(indices)

    y = patch_project(x, p[fe + "patch_embed.proj.weight"], p[fe + "patch_embed.proj.bias"], P)  # [B,Cn,N,D]
    N = y.shape[2]
    zero = torch.zeros((), dtype=y.dtype, device=y.device)
    tdl = zero
    if cfg.ortho_loss_v1_lambda > 0:  # dichavit.py:378-389, on the pre-add projection
        labels = torch.arange(Cn, device=x.device).repeat_interleave(N)
        tdl = tdl_loss(y.reshape(B, Cn * N, D), labels, cfg.gamma_s, cfg.gamma_d, cfg.reverse_pos_pairs,
                       cfg.use_square)
    cdl = zero
    if cfg.proxy_loss_lambda > 0:  # dichavit.py:399-402
        prox = p[fe + "patch_embed.channel_emb_proxies"][torch.tensor(cur_channels, device=x.device)]
        gt = torch.eye(Cn, device=x.device, dtype=y.dtype)
        cdl = proxy_loss(prox, channel_embed, gt, math.sqrt(1.0 / cfg.temperature))
    extra = tdl * cfg.ortho_loss_v1_lambda + cdl * cfg.proxy_loss_lambda  # dichavit.py:406-408

    tok = (y + channel_embed[None, :, None, :]).reshape(B, Cn * N, D)  # dichavit.py:409-415
    tok = torch.cat((p[fe + "cls_token"].expand(B, -1, -1), tok), dim=1)  # :561-562
    tok = tok + interpolate_pos(p[fe + "pos_embed"], Cn * N, H, W, P, Cn)  # :565

    xx = tok
    blocks = [] if keep_blocks else None
    for i in range(cfg.depth):  # dichavit.py:645-649
        xx = block_forward(xx, p, f"{fe}blocks.{i}.", cfg.heads)
        if keep_blocks:
            blocks.append(xx)
    xx = F.layer_norm(xx, (D,), p[fe + "norm.weight"], p[fe + "norm.bias"], eps=1e-6)  # :651
    feat = xx[:, 0]  # :652
    out = F.linear(feat, p["classifer_head.weight"], p["classifer_head.bias"]) if has_head else feat  # :855
    return OracleOutput(out=out, extra_loss=extra, tdl=tdl, cdl=cdl, tokens=tok, indices=indices, blocks=blocks)


# ---------------------------------------------------------------------------------------------
# trainer loss glue
# ---------------------------------------------------------------------------------------------

def train_loss(o: OracleOutput, y: torch.Tensor, p: Dict[str, torch.Tensor], cfg: OracleConfig, has_head: bool,
               extra_loss_lambda: float = 1.0) -> torch.Tensor:
    """trainer.py:986-995 (CE; JUMP-CP / So2Sat) and :876-914 (proxy_loss vs model.proxies; CHAMMI)."""
    if has_head:
        main = F.cross_entropy(o.out, y)
    else:
        main = proxy_loss(p["proxies"], o.out, y, math.sqrt(1.0 / cfg.temperature))
    return main + o.extra_loss * extra_loss_lambda


def loss_and_grads(x, y, params, cfg, channels, has_head, indices=None, extra_loss_lambda=1.0, dtype=torch.float32):
    """One training step's loss and parameter gradients (trainer.py:978-1001) by autograd."""
    p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in params.items() if k != "adaptive_interface.0"}
    o = forward(x.to(dtype), p, cfg, channels, training=True, has_head=has_head, indices=indices)
    loss = train_loss(o, y, p, cfg, has_head, extra_loss_lambda)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else None) for k, v in p.items()}
    return loss.detach(), o, grads


# ---------------------------------------------------------------------------------------------
# optimiser schedules
# ---------------------------------------------------------------------------------------------

def timm_cosine_lr(t: int, base_lr: float, t_initial: int, lr_min: float = 0.0, warmup_t: int = 0,
                   warmup_lr_init: float = 0.0, warmup_prefix: bool = False, cycle_mul: float = 1.0,
                   cycle_decay: float = 1.0, cycle_limit: int = 1, k_decay: float = 1.0) -> float:
    """CosineLRScheduler._get_lr(t) for one parameter group.  Third-party algorithm: timm (pinned 0.8.3.dev0 in the
    reference's requirements.txt:18, absent from /root/reference and from this image), restated from its published
    source timm/scheduler/cosine_lr.py; built by reference lr_schedulers.py:6-9 with configs/scheduler/cosine.yaml
    and stepped at trainer.py:344-348 (per epoch, t_in_epochs) or trainer.py:1009-1010 (per update)."""
    if t < warmup_t:
        return warmup_lr_init + t * ((base_lr - warmup_lr_init) / warmup_t)
    if warmup_prefix:
        t = t - warmup_t
    if cycle_mul != 1:
        i = math.floor(math.log(1 - t / t_initial * (1 - cycle_mul), cycle_mul))
        t_i = cycle_mul ** i * t_initial
        t_curr = t - (1 - cycle_mul ** i) / (1 - cycle_mul) * t_initial
    else:
        i = t // t_initial
        t_i = t_initial
        t_curr = t - (t_initial * i)
    lr_max = base_lr * (cycle_decay ** i)
    if i < cycle_limit:
        return lr_min + 0.5 * (lr_max - lr_min) * (1 + math.cos(math.pi * t_curr ** k_decay / t_i ** k_decay))
    return lr_min


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """utils.py:563-574, verbatim semantics (numpy table of epochs * niter_per_ep values)."""
    import numpy as np

    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


def trainer_lr_wd_sequence(n_updates: int, updates_per_epoch: int, epochs: int, base_lr: float, wd: float,
                           wd_end: Optional[float], sched: Optional[dict], t_in_epochs: bool) -> List[Tuple[float, float]]:
    """(lr, weight_decay) the optimiser sees at update 1..n_updates under the reference's training loop: the epoch
    loop calls scheduler.step(epoch) before each epoch (trainer.py:344-348; a no-op unless t_in_epochs), every batch
    calls optimizer.step() and THEN scheduler.step_update(num_updates) (a no-op if t_in_epochs) and overwrites
    param_group["weight_decay"] with wd_schedule[num_updates - 1] (trainer.py:1006-1019).  timm's constructor
    starts the groups at warmup_lr_init when warmup_t > 0 (Scheduler.__init__ / CosineLRScheduler.__init__)."""
    table = cosine_scheduler(wd, wd_end, epochs, updates_per_epoch) if wd_end is not None and wd_end > -1 else None
    lr = base_lr
    if sched is not None and sched.get("warmup_t", 0):
        lr = sched.get("warmup_lr_init", 0.0)
    cur_wd = wd
    out = []
    for u in range(1, n_updates + 1):
        epoch = 1 + (u - 1) // updates_per_epoch
        if sched is not None and t_in_epochs and (u - 1) % updates_per_epoch == 0:
            lr = timm_cosine_lr(epoch, base_lr, **sched)
        out.append((lr, cur_wd))
        if sched is not None and not t_in_epochs:
            lr = timm_cosine_lr(u, base_lr, **sched)
        if sched is not None and table is not None:
            cur_wd = float(table[min(u - 1, len(table) - 1)])
    return out
