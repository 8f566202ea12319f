"""Drop-in `DiChaViT` module (reference models/dichavit.py:748-865) whose forward and backward run
entirely in the hand-written sm_100a kernels of libdcvit.so (C ABI: include/dcvit.h).

Same constructor (`DiChaViT(config, mapper=...)`, factory `dichavit(cfg, **kw)`), same parameter
names / shapes / state_dict keys (154 for ViT-S), same forward signature and return convention
(train: `(out, extra_loss)`, eval: `out`), same host-visible attributes (`proxies`, `scale`,
`feature_extractor.patch_embed.counter`, `.mapper`).  There is no CPU / PyTorch fallback: calling
the module with a non-CUDA tensor or without the built library raises.

PyTorch is used for: parameter storage (one flat fp32 buffer, nn.Parameters are views into it),
the caching allocator (activation arena), streams, the DCS probability / `torch.multinomial` draw
(kept as the reference's own ATen ops so the sampled indices are bit-identical for a given RNG
state) and `torch.distributed` for the data-parallel gradient all-reduce.
"""
from __future__ import annotations

import ctypes
import math
import random
from collections import Counter
from ctypes import POINTER, Structure, byref, c_float, c_int, c_longlong, c_void_p
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import DcvError, check

_SIZES = {  # reference models/dichavit.py:676-745
    "tiny": (192, 3),
    "small": (384, 6),
    "distill": (384, 6),
    "base": (768, 12),
}
_DEPTH = 12
_MLP_RATIO = 4


# ---------------------------------------------------------------------------------------------
# ctypes mirrors of the structs in include/dcvit.h
# ---------------------------------------------------------------------------------------------
class _Dims(Structure):
    _fields_ = [(n, c_int) for n in ("B", "L", "D", "H", "F")]


class _BlockParams(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_b", "proj_b", "ln2_w", "ln2_b", "fc1_b", "fc2_b",
                                        "qkv_w", "proj_w", "fc1_w", "fc2_w")]


class _BlockGrads(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class _BlockActs(Structure):
    _fields_ = [(n, c_void_p) for n in ("x_in", "u", "mean1", "rstd1", "qkv", "o", "lse2", "x_mid", "v", "mean2",
                                        "rstd2", "h", "g", "x_out")]


class _BlockWs(Structure):
    _fields_ = [(n, c_void_p) for n in ("dh", "dv", "d_o", "dqkv", "delta", "dq_acc")]


class _EmbedDims(Structure):
    _fields_ = [(n, c_int) for n in ("B", "C", "Cs", "H", "W", "P", "D")]


class _EmbedCfg(Structure):
    _fields_ = [("lambda_tdl", c_float), ("lambda_cdl", c_float), ("gamma_s", c_float), ("gamma_d", c_float),
                ("cdl_scale", c_float), ("reverse_pos_pairs", c_int), ("use_square", c_int), ("x_is_u8", c_int)]


class _EmbedParams(Structure):
    _fields_ = [(n, c_void_p) for n in ("proj_w", "proj_b", "chan_embed", "proxies", "cls", "pos", "pos_map", "pix_mean",
                                        "pix_inv_std")]


class _EmbedGrads(Structure):
    _fields_ = [(n, c_void_p) for n in ("proj_w", "proj_b", "chan_embed", "proxies", "cls", "pos")]


class _EmbedActs(Structure):
    _fields_ = [(n, c_void_p) for n in ("patches", "wsplit", "pos_patch", "addend", "tokens", "S", "Q", "rnorm", "S_all",
                                        "loss_b", "coef_pos", "coef_neg", "cdl_dE", "cdl_dP", "tdl", "cdl", "extra")]


class _EmbedWs(Structure):
    _fields_ = [(n, c_void_p) for n in ("dY", "R", "dpos_patch")]


def _trunc_normal_(t: torch.Tensor, std: float = 0.02, a: float = -2.0, b: float = 2.0) -> torch.Tensor:
    """Truncated normal init, same RNG consumption as reference utils.py:477-517 (uniform -> erfinv)."""
    def cdf(v):
        return (1.0 + math.erf(v / math.sqrt(2.0))) / 2.0

    with torch.no_grad():
        lo, hi = cdf(a / std), cdf(b / std)
        t.uniform_(2 * lo - 1, 2 * hi - 1)
        t.erfinv_()
        t.mul_(std * math.sqrt(2.0))
        t.clamp_(min=a, max=b)
    return t


# ---------------------------------------------------------------------------------------------
# parameter containers (never executed; they only own parameters under the reference's names)
# ---------------------------------------------------------------------------------------------
class _Attention(nn.Module):  # reference models/vit.py:101-119
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):  # reference models/vit.py:59-74
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):  # reference models/vit.py:346-381
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)


class _LazyCounter:
    """`patch_embed.counter` (reference dichavit.py:66-67,214-216; read by trainer.py:796-804): how often
    each global channel id was sampled.  Counts accumulate on the device; they are copied to the host only
    when somebody reads them, so the training step has no device->host sync."""

    def __init__(self, n_channels: int):
        self.n = n_channels
        self._dev: Optional[torch.Tensor] = None
        self._host = np.zeros(n_channels, dtype=np.int64)

    def add(self, gid: torch.Tensor) -> None:
        if self._dev is None or self._dev.device != gid.device:
            self._flush()
            self._dev = torch.zeros(self.n, dtype=torch.int64, device=gid.device)
        self._dev.index_add_(0, gid.long(), torch.ones_like(gid, dtype=torch.int64))

    def add_host(self, ids: Sequence[int]) -> None:
        for k, v in Counter(ids).items():
            self._host[k] += v

    def _flush(self) -> None:
        if self._dev is not None:
            self._host += self._dev.cpu().numpy()
            self._dev.zero_()

    def as_dict(self) -> Dict[int, int]:
        self._flush()
        return {int(i): int(v) for i, v in enumerate(self._host) if v > 0}

    def items(self):
        return self.as_dict().items()

    def keys(self):
        return self.as_dict().keys()

    def values(self):
        return self.as_dict().values()

    def __getitem__(self, k):
        return self.as_dict().get(k, 0)

    def __iter__(self):
        return iter(self.as_dict())

    def __len__(self):
        return len(self.as_dict())


class PatchEmbedPerChannel(nn.Module):
    """Parameters + DCS index selection of reference models/dichavit.py:39-216."""

    def __init__(self, config, img_size: int, patch_size: int, in_chans: int, mapper, embed_dim: int,
                 enable_sample: bool, use_channelvit_channels: bool = True):
        super().__init__()
        self.cfg = config
        self.img_size = img_size
        self.mapper = mapper
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) * (img_size // patch_size) * in_chans
        self.channel_scale = np.sqrt(1.0 / config.temperature)
        if config.proxy_loss_lambda > 0:
            self.channel_emb_proxies = nn.Parameter(torch.randn(in_chans, embed_dim) / 8)
            if config.get("proxy_orthogonal_init", False):
                nn.init.orthogonal_(self.channel_emb_proxies)
        if config.hcs_sampling != "none" and config.hcs_sampling is not None:
            self.counter = _LazyCounter(in_chans)
        if not isinstance(config.hcs_sampling, str) and config.hcs_sampling is not None:
            raise ValueError("hcs_sampling must be a string")
        self.proj = nn.Conv3d(1, embed_dim, kernel_size=(1, patch_size, patch_size),
                              stride=(1, patch_size, patch_size))
        if not use_channelvit_channels:
            raise NotImplementedError("use_channelvit_channels=False is outside the DiChaViT hot path")
        self.channel_embed = nn.Embedding(in_chans, embed_dim)
        if config.orthogonal_channel_emb_init:
            nn.init.orthogonal_(self.channel_embed.weight)
        else:
            _trunc_normal_(self.channel_embed.weight, std=0.02)
        if config.freeze_channel_emb:
            self.channel_embed.weight.requires_grad = False
        self.use_channelvit_channels = use_channelvit_channels
        self.enable_sample = enable_sample
        self._chan_cache: Dict[tuple, torch.Tensor] = {}

    def chunk_channels(self, chunk_name: str, device) -> torch.Tensor:
        key = (chunk_name, tuple(self.mapper[chunk_name]), str(device))
        t = self._chan_cache.get(key)
        if t is None:
            t = torch.tensor(list(self.mapper[chunk_name]), dtype=torch.int64, device=device)
            self._chan_cache[key] = t
        return t

    def select_channels(self, chunk_name: str, n_in: int, device):
        """DCS (reference dichavit.py:127-216).  Returns (C', idx int32 [C'] or None, gid int32 [C']):
        positions of the kept channels inside x, and their global channel ids, both on `device`.
        RNG consumption is the reference's: random.randint(1,C), random.randint(0,C-1), then
        torch.multinomial on the device generator.  Nothing is copied back to the host."""
        pre = getattr(self, "_prefetched", None)
        if pre is not None:
            self._prefetched = None
            if pre[0] == (chunk_name, n_in, str(device), self.training, self.enable_sample):
                return pre[1]
        chan = self.chunk_channels(chunk_name, device)
        if chan.numel() != n_in:
            raise ValueError(f"x has {n_in} channels but mapper['{chunk_name}'] lists {chan.numel()}")
        draw = self.draw_host(chunk_name, n_in)
        if draw is None:
            return n_in, None, chan.to(torch.int32)
        return self.select_device(chunk_name, device, draw)

    def draw_host(self, chunk_name: str, n_in: int) -> Optional[dict]:
        """The host half of DCS (reference dichavit.py:127-174): the python `random` draws, in the reference's order.
        None: no sampling (eval mode / enable_sample off).  The device half (select_device) is what a captured CUDA
        graph replays; the host half decides which (C') graph to launch."""
        if not (self.training and self.enable_sample):
            return None
        mode = self.cfg.hcs_sampling
        c_new = random.randint(1, n_in)
        if mode == "none" or mode is None:
            cur = random.sample(list(self.mapper[chunk_name]), k=c_new)
            pos = [list(self.mapper[chunk_name]).index(c) for c in cur]
            return dict(mode="none", c_new=c_new, pos=pos, cur=cur)
        if mode == "hcs_per_sample":
            raise ValueError("hcs_per_sample not implemented!")  # as the reference, dichavit.py:394-395
        if mode not in ("lowest_cosine_prob", "lowest_cosine", "highest_cosine"):
            raise NotImplementedError(f"hcs_sampling='{mode}' is outside the DiChaViT hot path")
        return dict(mode=mode, c_new=c_new, anchor=random.randint(0, n_in - 1))

    def select_device(self, chunk_name: str, device, draw: dict, anchor_dev: Optional[torch.Tensor] = None,
                      pos_dev: Optional[torch.Tensor] = None):
        """The device half of DCS (reference dichavit.py:176-216): cosine to the anchor, softmax, torch.multinomial on
        the device generator (the reference's own ATen calls: bit-identical indices for a given RNG state), anchor
        fix-up, counter.  `anchor_dev` (int64 [1]) / `pos_dev` (int32 [C']): read the host draw from device memory
        instead of baking it into the launches -- what the captured-graph step does."""
        chan = self.chunk_channels(chunk_name, device)
        c_new, mode = draw["c_new"], draw["mode"]
        if mode == "none":
            idx = pos_dev if pos_dev is not None else torch.tensor(draw["pos"], dtype=torch.int32, device=device)
            gid = chan[idx.long()].to(torch.int32) if pos_dev is not None else \
                torch.tensor(draw["cur"], dtype=torch.int32, device=device)
            return c_new, idx, gid
        with torch.no_grad():
            emb = self.channel_embed.weight[chan]
            emb_n = F.normalize(emb, p=2, dim=-1)
            cos_all = torch.einsum("c d, e d -> c e", emb_n, emb_n)
            cosine = cos_all.index_select(0, anchor_dev)[0] if anchor_dev is not None else cos_all[draw["anchor"]]
            if mode == "lowest_cosine_prob":
                prob = F.softmax((1 - cosine) / self.cfg.hcs_sampling_temp, dim=-1)
                indices = torch.multinomial(prob, c_new, replacement=False)
            else:
                indices = torch.topk(cosine, k=c_new, largest=(mode == "highest_cosine"))[1]
            # `if anchor not in indices: indices[-1] = anchor` (dichavit.py:201-202), on the device
            anchor_t = anchor_dev[0] if anchor_dev is not None else \
                torch.full((), draw["anchor"], dtype=indices.dtype, device=device)
            last = torch.where((indices == anchor_t).any(), indices[-1], anchor_t)
            indices = torch.cat((indices[:-1], last[None]))
            gid = chan[indices]
            self.counter.add(gid)
        return c_new, indices.to(torch.int32), gid.to(torch.int32)

    def prefetch(self, chunk_name: str, n_in: int, device) -> None:
        """Draw the NEXT forward's DCS selection now (same RNG calls, just earlier) so that the ~20 tiny sampling
        launches overlap the GPU work still queued from the current step instead of sitting between the host's
        per-step synchronisation and the first heavy kernel.  Opt-in: the next forward uses the draw iff it asks for
        the same (chunk, channel count, device, mode); nothing else may consume the RNGs in between if bit-exact
        agreement with an un-prefetched run is required."""
        self._prefetched = None
        key = (chunk_name, n_in, str(device), self.training, self.enable_sample)
        self._prefetched = (key, self.select_channels(chunk_name, n_in, device))

    def leave_one_out_tokens(self, chunk_name: str, training_chunks: str, new_channel_init) -> Optional[torch.Tensor]:
        """Eval-time channel tokens for chunks with channels unseen in training (reference dichavit.py:219-374).
        Returns None when every channel of the chunk was seen (the branch degenerates to the plain lookup), else a
        [C_in, D] tensor where unseen channels get a token synthesised from the training channels.  The
        dynamic_input_corr_* modes need a `bank` attribute that nothing in the reference ever sets: like the
        reference they raise ValueError("provide a channel_map (dict)!")."""
        mapper = self.mapper
        training_channels = [c for ch in str(training_chunks).split("_") for c in mapper[ch]]
        chunk = list(mapper[chunk_name])
        if all(c in training_channels for c in chunk):
            return None
        mode = getattr(new_channel_init, "value", new_channel_init)
        if not isinstance(mode, str):
            raise TypeError("new_channel_init must be given when the chunk has channels unseen in training")
        chs_not_seen = [c for c in training_channels if c not in chunk]
        bank = chs_not_seen if "not_in_chunk" in mode else training_channels
        w = self.channel_embed.weight.detach()
        rows, cur = [], 0
        for c in chunk:
            if c in training_channels:
                rows.append(w[c][None])
                continue
            n = len(bank)
            if mode in ("avg_2", "avg_2_not_in_chunk"):
                param = w[[bank[cur], bank[(cur + 1) % n]]].mean(dim=0, keepdim=True)
            elif mode in ("avg_3", "avg_3_not_in_chunk"):
                param = w[[bank[cur], bank[(cur + 1) % n], bank[(cur + 2) % n]]].mean(dim=0, keepdim=True)
            elif mode == "replicate":
                param = w[bank[cur]][None]
            elif mode == "zero":
                param = torch.zeros_like(w[0])[None]
            elif mode == "random":
                param = w[c][None]
            elif mode == "fixed_input_corr":
                if not hasattr(self, "channel_map"):
                    raise ValueError("provide a channel_map (dict)!")
                param = w[self.channel_map[c]][None]
            elif mode == "random_input_corr":
                param = w[int(np.random.choice(training_channels))][None]
            elif mode.startswith("dynamic_input_corr"):
                if not hasattr(self, "bank"):
                    raise ValueError("provide a channel_map (dict)!")
                raise NotImplementedError("dynamic_input_corr_* needs an image bank the reference never provides")
            else:
                raise ValueError(f"Invalid new_channel_init: '{mode}'")
            cur = (cur + 1) % n
            rows.append(param)
        return torch.cat(rows, dim=0).contiguous().float()


class ChannelVisionTransformer(nn.Module):
    """Parameter layout of reference models/dichavit.py:420-516."""

    def __init__(self, config, img_size, patch_size: int, in_chans: int, mapper, embed_dim: int, depth: int,
                 num_heads: int, mlp_ratio: int, enable_sample: bool, use_channelvit_channels: bool = True):
        super().__init__()
        self.cfg = config
        if config.drop_path_rate != 0:
            raise NotImplementedError("drop_path_rate != 0 is outside the DiChaViT hot path (reference default 0.0)")
        if config.block_type != "block":
            if config.block_type == "block_v2":
                raise NotImplementedError("block_type=block_v2 (token pruning) is outside the DiChaViT hot path")
            raise ValueError(f"Unknown block type: {config.block_type}")
        if config.dropout_tokens_hcs not in ("none", None):
            raise NotImplementedError("dropout_tokens_hcs != none is outside the DiChaViT hot path")
        self.num_features = self.embed_dim = self.out_dim = embed_dim
        self.in_chans = in_chans
        self.num_heads = num_heads
        img = img_size[0] if isinstance(img_size, (list, tuple)) or hasattr(img_size, "__getitem__") else img_size
        self.patch_embed = PatchEmbedPerChannel(config, img, patch_size, in_chans, mapper, embed_dim, enable_sample,
                                                use_channelvit_channels)
        num_patches = self.patch_embed.num_patches
        self.patch_size = patch_size
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.num_extra_tokens = 1
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches // in_chans + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, int(embed_dim * mlp_ratio)) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Identity()
        _trunc_normal_(self.pos_embed, std=0.02)
        _trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):  # reference dichavit.py:509-516
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


def bicubic_pos_matrix(grid: int, w: int, h: int, patch: int) -> torch.Tensor:
    """The fixed linear map of reference dichavit.py:531-549: F.interpolate(mode="bicubic") of the
    [grid, grid] positional grid with scale ((w//P + 0.1)/grid, (h//P + 0.1)/grid), as an explicit
    [N_out, N_in] fp32 matrix (identity pushed through the same ATen call, once, on the host)."""
    n_in = grid * grid
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    eye = torch.eye(n_in, dtype=torch.float32).reshape(1, grid, grid, n_in).permute(0, 3, 1, 2)
    out = F.interpolate(eye, scale_factor=(w0 / grid, h0 / grid), mode="bicubic")
    if int(w0) != out.shape[-2] or int(h0) != out.shape[-1]:
        raise ValueError("bicubic positional resample produced an unexpected grid")
    return out.permute(0, 2, 3, 1).reshape(-1, n_in).contiguous()


def bicubic_pos_backward_matrix(grid: int, w: int, h: int, patch: int) -> Optional[torch.Tensor]:
    """What the reference's BACKWARD applies to the gradient of the resampled positional grid -- which is not the
    transpose of `bicubic_pos_matrix`: ATen's upsample_bicubic2d backward, reached through F.interpolate(scale_factor=...),
    derives its sampling scale from the tensor sizes (out / in), not from the (w//P + 0.1)/grid the forward was given.
    For every training shape (w//P == grid) that scale is 1 and the backward is the IDENTITY, while the forward matrix
    has ~1.8 % off-diagonal mass: the true adjoint differs from what the reference trains with by 9 % on the patch rows
    of d pos_embed (So2Sat shape, found with tools/grad_diag.py).  Parity means the reference's gradient, so this is the
    matrix autograd itself applies, read out with one backward pass over one-hot output gradients.  Returns
    [N_out, N_in] (d pos_in = M^T d pos_out), or None when it is the identity (no map in the backward)."""
    n_in = grid * grid
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    wo, ho = int(w0), int(h0)
    n_out = wo * ho
    with torch.enable_grad():
        x = torch.zeros(1, n_out, grid, grid, dtype=torch.float32, requires_grad=True)
        y = F.interpolate(x, scale_factor=(w0 / grid, h0 / grid), mode="bicubic")
        gy = torch.zeros_like(y)
        j = torch.arange(n_out)
        gy[0, j, j // ho, j % ho] = 1.0
        y.backward(gy)
    m = x.grad[0].reshape(n_out, n_in).contiguous()
    if n_out == n_in and torch.equal(m, torch.eye(n_in)):
        return None
    return m


# ---------------------------------------------------------------------------------------------
# flat parameter store
# ---------------------------------------------------------------------------------------------
_ALIGN = 64  # elements: 256 B in fp32, 128 B in bf16 (TMA needs 16 B)


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class _Arena:
    """Carves one uint8 torch allocation into 256-byte aligned raw pointers."""

    def __init__(self):
        self.off = 0
        self.slots: Dict[str, int] = {}

    def add(self, name: str, nbytes: int) -> None:
        self.slots[name] = self.off
        self.off += _round_up(max(int(nbytes), 4), 256)

    def alloc(self, device, at_least: int = 0) -> torch.Tensor:
        return torch.empty(max(self.off, at_least, 256), dtype=torch.uint8, device=device)


class DiChaViT(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.cfg = config
        mapper = kwargs["mapper"]
        name = config.pretrained_model_name
        if name not in _SIZES:
            raise ValueError("Unknown model name")
        dim, heads = _SIZES[name]
        total_in_channels = len(config.in_channel_names)
        self.feature_extractor = ChannelVisionTransformer(
            config, img_size=config.img_size, patch_size=config.patch_size, in_chans=total_in_channels, mapper=mapper,
            embed_dim=dim, depth=_DEPTH, num_heads=heads, mlp_ratio=_MLP_RATIO, enable_sample=config.enable_sample,
            use_channelvit_channels=config.use_channelvit_channels)
        self.classifer_head = nn.Identity()
        if "Allen" not in mapper:
            self.classifer_head = nn.Linear(dim, config.num_classes)
        self.dim = dim
        self.proxies = nn.Parameter(torch.randn(config.num_classes, dim) / 8)
        if config.learnable_temp:
            self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / config.temperature))
        else:
            self.scale = np.sqrt(1.0 / config.temperature)
        self.adaptive_interface = nn.ParameterList([self.proxies])

        # engine state
        self._flat: Optional[torch.Tensor] = None      # fp32 master parameters
        self._bflat: Optional[torch.Tensor] = None     # bf16 operand copy of the parameters
        self._bflat_version = -1                       # _flat._version the bf16 copy corresponds to (-1: stale)
        self._layout: List[tuple] = []                 # (param, offset, numel)
        self._groups: List[tuple] = []                 # (name, start, end) gradient buckets in backward order
        self._pos_maps: Dict[tuple, torch.Tensor] = {}
        self.grad_allreduce = False                    # set by enable_data_parallel()
        self._pg = None
        self._comm_stream = None
        self.last_losses: Dict[str, torch.Tensor] = {}
        self._plan_cache: Dict[tuple, dict] = {}
        self._bp_cache = None
        self.direct_grad = False  # see _DiChaViTFn.backward
        self._static_gflat: Optional[torch.Tensor] = None  # set by graphs.GraphedTrainStep
        self.grad_sync = True     # data parallel: all-reduce in this backward (False inside no_sync())
        self._grad_anchor: Optional[torch.Tensor] = None
        self._last_gflat: Optional[torch.Tensor] = None
        self._dp_synced_ptr = 0   # data_ptr of the flat buffer whose contents were broadcast from rank 0
        self._arena_bytes: Dict[tuple, int] = {}   # forward arena size of the full-channel plan per input shape
        self._ws_bytes: Dict[tuple, int] = {}      # backward workspace high-water mark per input shape

    # ------------------------------------------------------------------ parameters
    def _ordered_params(self):
        fe = self.feature_extractor
        pe = fe.patch_embed
        embed = [fe.cls_token, fe.pos_embed]
        if hasattr(pe, "channel_emb_proxies"):
            embed.append(pe.channel_emb_proxies)
        embed += [pe.proj.weight, pe.proj.bias, pe.channel_embed.weight]
        groups = [("embed", embed)]
        for i, b in enumerate(fe.blocks):
            groups.append((f"block{i}", [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.qkv.bias,
                                         b.attn.proj.weight, b.attn.proj.bias, b.norm2.weight, b.norm2.bias,
                                         b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]))
        tail = [fe.norm.weight, fe.norm.bias]
        if isinstance(self.classifer_head, nn.Linear):
            tail += [self.classifer_head.weight, self.classifer_head.bias]
        tail.append(self.proxies)
        if hasattr(self, "logit_scale"):
            tail.append(self.logit_scale)
        groups.append(("tail", tail))
        return groups

    @property
    def _external_ids(self):
        """Parameters the module owns for the trainer but never reads in forward (`proxies`, `logit_scale`: used by the
        trainer's loss glue, trainer.py:876-914).  Their gradients come from torch autograd, not from the kernels."""
        ids = {id(self.proxies)}
        if hasattr(self, "logit_scale"):
            ids.add(id(self.logit_scale))
        return ids

    # engine caches hold raw device pointers (ctypes structs) and buffers tied to this instance: copies / pickles
    # (AveragedModel for SWA at trainer.py:243, checkpointing of the module object) rebuild them lazily
    def __getstate__(self):
        st = dict(self.__dict__)
        st.update(_flat=None, _bflat=None, _bflat_version=-1, _layout=[], _groups=[], _pos_maps={}, _plan_cache={},
                  _bp_cache=None, _grad_anchor=None, _last_gflat=None, _static_gflat=None, _comm_stream=None, _pg=None, _arena_bytes={},
                  _ws_bytes={}, last_losses={}, _dp_synced_ptr=0, _off={})
        return st

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        memo[id(self.cfg)] = self.cfg  # the config is shared, not copied (read-only by convention; attr-dict configs
        #                                whose __getattr__ raises KeyError cannot be deep-copied at all)
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)  # Parameter.__deepcopy__ clones: no view of our flat buffer survives
        return new

    def _param_version(self) -> int:
        """Sum of the autograd version counters of every parameter (+ the flat buffer's): changes whenever torch
        writes a parameter in place."""
        try:
            v = self._flat._version
            for p, _, _ in self._layout:
                v += p._version
        except RuntimeError:  # inference tensors carry no version counter: never trust the copy
            return -1
        return v

    def prefetch_dcs(self, chunk_name: str, n_in: int) -> None:
        """See PatchEmbedPerChannel.prefetch: overlap the next step's channel sampling with the current step."""
        self.feature_extractor.patch_embed.prefetch(chunk_name, n_in, self._flat.device if self._flat is not None else "cuda")

    def mark_params_dirty(self) -> None:
        """Force the next forward to rebuild the bf16 operand copy (needed after writes through `p.data`)."""
        self._bflat_version = -1

    def _ensure_flat(self, device) -> None:
        ok = self._flat is not None and self._flat.device == device
        if ok:
            base = self._flat.data_ptr()
            for p, off, n in self._layout:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            return
        groups = self._ordered_params()
        layout, bounds, off = [], [], 0
        for gname, plist in groups:
            start = off
            for p in plist:
                layout.append((p, off, p.numel()))
                off += _round_up(p.numel(), _ALIGN)
            bounds.append((gname, start, off))
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, o, n in layout:
                if p.device != device:
                    raise DcvError(f"parameter on {p.device}, input on {device}: move the module with .to(device)")
                flat[o:o + n].copy_(p.detach().reshape(-1).float())
                p.data = flat[o:o + n].view(p.shape)
        self._flat, self._layout, self._groups = flat, layout, bounds
        self._bflat = torch.empty(off, dtype=torch.bfloat16, device=device)
        self._bflat_version = -1
        self._off = {id(p): o for p, o, _ in layout}
        self._last_gflat = None
        self._sync_params_from_rank0()

    def _sync_params_from_rank0(self) -> None:
        """DDP broadcasts rank 0's parameters at construction (trainer.py:1185); the reference seeds every process
        differently by default (trainer.py:81), so without this the replicas would start -- and stay -- different."""
        if not self.grad_allreduce or self._flat is None or self._dp_synced_ptr == self._flat.data_ptr():
            return
        import torch.distributed as dist

        dist.broadcast(self._flat, src=dist.get_global_rank(self._pg, 0) if self._pg is not None else 0, group=self._pg)
        self._dp_synced_ptr = self._flat.data_ptr()
        self._bflat_version = -1

    def _allreduce_external(self, flat_slice: torch.Tensor) -> None:
        """average a trainer-owned parameter's gradient (already copied into the flat buffer) over the ranks"""
        import torch.distributed as dist

        flat_slice.mul_(1.0 / dist.get_world_size(group=self._pg))
        dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self._pg)

    def allreduce_external_grads(self) -> None:
        """Data parallel with a torch optimizer instead of FusedAdamW: the kernels' gradients are averaged by the
        bucketed reducer during backward, those of `proxies` / `logit_scale` (produced by torch autograd from the
        trainer's proxy loss) are averaged here -- call it after the last backward of the step."""
        if not self.grad_allreduce:
            return
        for p in (self.proxies, getattr(self, "logit_scale", None)):
            if p is not None and p.grad is not None:
                self._allreduce_external(p.grad)

    def no_sync(self):
        """Context manager like DDP.no_sync(): backwards inside it do not all-reduce (gradient accumulation over
        several forward/backward passes, e.g. the three CHAMMI chunks of one optimiser step, trainer.py:846-931);
        the first backward outside it reduces the accumulated flat gradient."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old = self.grad_sync
            self.grad_sync = False
            try:
                yield
            finally:
                self.grad_sync = old

        return ctx()

    def _fptr(self, p) -> int:
        return self._flat.data_ptr() + 4 * self._off[id(p)]

    def _bptr(self, p) -> int:
        return self._bflat.data_ptr() + 2 * self._off[id(p)]

    def _pos_map(self, w: int, h: int, device) -> torch.Tensor:
        fe = self.feature_extractor
        n = fe.pos_embed.shape[1] - 1
        key = (w, h, str(device))
        m = self._pos_maps.get(key)
        if m is None:
            m = bicubic_pos_matrix(int(math.sqrt(n)), w, h, fe.patch_size).to(device)
            self._pos_maps[key] = m
        return m

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None, overlap: bool = True) -> "DiChaViT":
        """Batch-sharded data parallelism (replaces DDP at reference trainer.py:1185): gradient buckets
        are all-reduced (average) with NCCL on a side stream as soon as the backward of the layers they
        cover has been enqueued, overlapping the remaining backward."""
        import torch.distributed as dist

        if not dist.is_initialized():
            raise DcvError("torch.distributed is not initialised")
        self.grad_allreduce = True
        self._pg = process_group
        self._overlap = overlap
        self._dp_synced_ptr = 0
        self._sync_params_from_rank0()  # if the flat buffer exists already; else at its creation
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, chunk_name: str, training_chunks: Optional[str] = None, init_first_layer=None,
                new_channel_init=None, **kwargs):
        if not x.is_cuda:
            raise DcvError("DiChaViT (B200-native) needs a CUDA tensor: there is no CPU fallback")
        if x.dim() != 4:
            raise ValueError("x must be [B, C, H, W]")
        # uint8 input: the loader's per-channel standardisation runs inside the patch-gather kernel (8(f) #3);
        # kwargs pixel_mean / pixel_std: [C] tensors or sequences in the order of x's channels
        x_is_u8 = x.dtype == torch.uint8
        pix_norm = None
        if x_is_u8:
            x = x.contiguous()
            if kwargs.get("pixel_mean") is not None:
                mean = torch.as_tensor(kwargs["pixel_mean"], dtype=torch.float32, device=x.device).contiguous()
                std = torch.as_tensor(kwargs["pixel_std"], dtype=torch.float32, device=x.device).contiguous()
                if mean.numel() != x.shape[1] or std.numel() != x.shape[1]:
                    raise ValueError("pixel_mean / pixel_std must have one entry per input channel")
                pix_norm = (mean, (1.0 / std).contiguous())
        else:
            x = x.contiguous().float()
        self._ensure_flat(x.device)
        pe = self.feature_extractor.patch_embed
        cs, idx, gid = pe.select_channels(chunk_name, x.shape[1], x.device)
        ce_override = None
        if (not self.training) and training_chunks is not None:  # reference dichavit.py:219
            ce_override = pe.leave_one_out_tokens(chunk_name, training_chunks, new_channel_init)
            if ce_override is not None:
                gid = torch.arange(cs, dtype=torch.int32, device=x.device)  # rows of the synthesised token matrix
        # everything the backward needs to know about THIS call travels with the autograd node, not on the module:
        # a second forward, or a train()/eval() toggle, before backward() must not change the gradients
        call = dict(training=self.training, x_is_u8=x_is_u8, pix_norm=pix_norm, ce_override=ce_override)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p, _, _ in self._layout)
        if self.direct_grad and need_grad:
            if self._grad_anchor is None or self._grad_anchor.device != x.device:
                self._grad_anchor = torch.zeros((), device=x.device, requires_grad=True)
            params = [self._grad_anchor]  # one differentiable input keeps the node in the graph
        else:
            ext = self._external_ids
            params = [p for p, _, _ in self._layout if id(p) not in ext]
        out, extra = _DiChaViTFn.apply(self, x, cs, idx, gid, need_grad, call, *params)
        if self.training:
            return out, extra
        return out

    # ------------------------------------------------------------------ engine
    def _plan(self, B: int, cs: int, H: int, W: int, keep: bool):
        key = (B, cs, H, W, keep)
        hit = self._plan_cache.get(key)
        if hit is not None:
            return hit
        pl = self._plan_build(B, cs, H, W, keep)
        self._plan_cache[key] = pl
        return pl

    def _plan_build(self, B: int, cs: int, H: int, W: int, keep: bool):
        fe = self.feature_extractor
        P, D, heads = fe.patch_size, self.dim, fe.num_heads
        N = (H // P) * (W // P)
        T = cs * N
        L = T + 1
        M = B * L
        Fh = D * _MLP_RATIO
        Lp = _round_up(L, 128)
        depth = len(fe.blocks)
        ar = _Arena()
        ar.add("patches", B * T * 3 * P * P * 2)
        ar.add("wsplit", D * 3 * P * P * 2)
        ar.add("pos_patch", N * D * 4)
        ar.add("addend", T * D * 4)
        for nm, sz in (("S", B * cs * D), ("Q", B * cs), ("rnorm", B * T), ("S_all", B * D), ("loss_b", B),
                       ("coef_pos", B), ("coef_neg", B), ("cdl_dE", cs * D), ("cdl_dP", cs * D)):
            ar.add(nm, sz * 4)
        ar.add("head_mean", B * 4)
        ar.add("head_rstd", B * 4)
        nsets = depth if keep else 1
        nx = depth + 1 if keep else 2
        for i in range(nx):
            # the last block only produces the CLS row of every image (dcv_block_fwd_cls)
            ar.add(f"x{i}", (B if keep and i == depth else M) * D * 4)
        for i in range(nsets):
            rows = B if keep and i == depth - 1 else M  # compact tail of the last block
            ar.add(f"u{i}", M * D * 2)
            ar.add(f"mean1_{i}", M * 4)
            ar.add(f"rstd1_{i}", M * 4)
            ar.add(f"qkv{i}", M * 3 * D * 2)
            ar.add(f"o{i}", M * D * 2)
            ar.add(f"lse{i}", B * heads * Lp * 4)
            ar.add(f"xmid{i}", rows * D * 4)
            ar.add(f"v{i}", rows * D * 2)
            ar.add(f"mean2_{i}", rows * 4)
            ar.add(f"rstd2_{i}", rows * 4)
            ar.add(f"h{i}", rows * Fh * 2)
            ar.add(f"g{i}", rows * Fh * 2)
        return dict(B=B, cs=cs, H=H, W=W, P=P, D=D, heads=heads, N=N, T=T, L=L, M=M, F=Fh, Lp=Lp, depth=depth,
                    arena=ar, keep=keep)

    def _pos_map_bwd(self, w: int, h: int, device) -> Optional[torch.Tensor]:
        fe = self.feature_extractor
        n = fe.pos_embed.shape[1] - 1
        key = ("bwd", w, h, str(device))
        if key not in self._pos_maps:
            m = bicubic_pos_backward_matrix(int(math.sqrt(n)), w, h, fe.patch_size)
            self._pos_maps[key] = m.to(device) if m is not None else None
        return self._pos_maps[key]

    def _embed_structs(self, pl, base: int, scal: torch.Tensor, C_in: int, use_map: bool, device, call: dict,
                       backward: bool = False):
        fe = self.feature_extractor
        pe = fe.patch_embed
        cfg = self.cfg
        s = pl["arena"].slots
        dims = _EmbedDims(pl["B"], C_in, pl["cs"], pl["H"], pl["W"], pl["P"], pl["D"])
        # TDL / CDL only enter the training output (dichavit.py:856-861); the reference still evaluates them in eval
        # mode and throws the value away -- here the kernels are simply not launched
        l_tdl = float(cfg.ortho_loss_v1_lambda) if call["training"] else 0.0
        l_cdl = float(cfg.proxy_loss_lambda) if call["training"] else 0.0
        norm = call["pix_norm"]
        ecfg = _EmbedCfg(l_tdl, l_cdl, float(cfg.gamma_s), float(cfg.gamma_d), float(pe.channel_scale),
                         int(bool(cfg.reverse_pos_pairs)), int(bool(cfg.use_square)), int(call["x_is_u8"]))
        has_prox = hasattr(pe, "channel_emb_proxies")
        # forward: the bicubic resample matrix; backward: what the reference's autograd applies instead of its transpose
        # (bicubic_pos_backward_matrix; None = identity, the kernels then accumulate d pos_embed directly)
        if not use_map:
            pos_map = None
        elif backward:
            pos_map = self._pos_map_bwd(pl["W"], pl["H"], device)
        else:
            pos_map = self._pos_map(pl["W"], pl["H"], device)
        ce_ptr = call["ce_override"].data_ptr() if call["ce_override"] is not None else \
            self._fptr(pe.channel_embed.weight)
        ep = _EmbedParams(self._fptr(pe.proj.weight), self._fptr(pe.proj.bias), ce_ptr,
                          self._fptr(pe.channel_emb_proxies) if has_prox else None, self._fptr(fe.cls_token),
                          self._fptr(fe.pos_embed), pos_map.data_ptr() if pos_map is not None else None,
                          norm[0].data_ptr() if norm is not None else None,
                          norm[1].data_ptr() if norm is not None else None)
        sp = scal.data_ptr()
        acts = _EmbedActs(base + s["patches"], base + s["wsplit"], base + s["pos_patch"], base + s["addend"], base + s["x0"],
                          base + s["S"], base + s["Q"], base + s["rnorm"], base + s["S_all"], base + s["loss_b"],
                          base + s["coef_pos"], base + s["coef_neg"], base + s["cdl_dE"], base + s["cdl_dP"],
                          sp, sp + 4, sp + 8)
        return dims, ecfg, ep, acts, pos_map

    def _block_param_structs(self):
        """ctypes parameter structs of every block (pointers into the flat fp32 / bf16 buffers), and the element
        offsets of every block's gradient tensors inside a flat gradient buffer; rebuilt only after a re-flatten."""
        key = (self._flat.data_ptr(), self._bflat.data_ptr())
        if self._bp_cache is None or self._bp_cache[0] != key:
            structs, goffs = [], []
            for b in self.feature_extractor.blocks:
                structs.append(_BlockParams(
                    self._fptr(b.norm1.weight), self._fptr(b.norm1.bias), self._fptr(b.attn.qkv.bias),
                    self._fptr(b.attn.proj.bias), self._fptr(b.norm2.weight), self._fptr(b.norm2.bias),
                    self._fptr(b.mlp.fc1.bias), self._fptr(b.mlp.fc2.bias), self._bptr(b.attn.qkv.weight),
                    self._bptr(b.attn.proj.weight), self._bptr(b.mlp.fc1.weight), self._bptr(b.mlp.fc2.weight)))
                goffs.append([4 * self._off[id(t)] for t in (
                    b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.qkv.bias, b.attn.proj.weight,
                    b.attn.proj.bias, b.norm2.weight, b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight,
                    b.mlp.fc2.bias)])
            self._bp_cache = (key, structs, goffs)
        return self._bp_cache[1], self._bp_cache[2]

    def _block_structs(self, pl, base: int, i: int):
        s = pl["arena"].slots
        keep = pl["keep"]
        k = i if keep else 0
        xin = f"x{i}" if keep else f"x{i % 2}"
        xout = f"x{i + 1}" if keep else f"x{(i + 1) % 2}"
        bp = self._block_param_structs()[0][i]
        cache = pl.setdefault("acts_cache", {})
        ba = cache.get((base, i))
        if ba is None:
            ba = _BlockActs(base + s[xin], base + s[f"u{k}"], base + s[f"mean1_{k}"], base + s[f"rstd1_{k}"],
                            base + s[f"qkv{k}"], base + s[f"o{k}"], base + s[f"lse{k}"], base + s[f"xmid{k}"],
                            base + s[f"v{k}"], base + s[f"mean2_{k}"], base + s[f"rstd2_{k}"], base + s[f"h{k}"],
                            base + s[f"g{k}"], base + s[xout])
            if len(cache) < 256:
                cache[(base, i)] = ba
        return bp, ba, xout

    def _run_forward(self, x: torch.Tensor, cs: int, idx, gid, keep: bool, call: dict):
        lib = _lib.lib()
        st = _lib.stream_ptr()
        dev = x.device
        B, C_in, H, W = x.shape
        fe = self.feature_extractor
        n_pos = fe.pos_embed.shape[1] - 1
        P = fe.patch_size
        if (H // P) * (W // P) != n_pos:
            raise ValueError(f"input {H}x{W} does not match the {n_pos}-patch positional grid")
        pl = self._plan(B, cs, H, W, keep)
        # DCS changes the token count every step: always request the size of the full-channel plan so that the
        # caching allocator hands back the same block instead of growing a new one per (B, C') shape
        key = (B, C_in, H, W, keep)
        if key not in self._arena_bytes:
            self._arena_bytes[key] = self._plan(B, C_in, H, W, keep)["arena"].off
        arena_min = self._arena_bytes[key]
        # fp32 master -> bf16 operand copy of every parameter (one launch) -- skipped while the copy is current: the
        # fused AdamW step writes it itself, and any torch-side in-place write to a parameter (optimizer,
        # load_state_dict, init) bumps that parameter's version counter.  Writes through `.data` are invisible to
        # the counters: call mark_params_dirty() after them.
        ver = self._param_version()
        if ver < 0 or self._bflat_version != ver:
            check(lib.dcv_cast_f32_bf16(c_void_p(self._flat.data_ptr()), c_void_p(self._bflat.data_ptr()),
                                        c_longlong(self._flat.numel()), st), "dcv_cast_f32_bf16")
            self._bflat_version = ver
        arena = pl["arena"].alloc(dev, arena_min)
        base = arena.data_ptr()
        scal = torch.zeros(4, dtype=torch.float32, device=dev)  # tdl, cdl, extra
        # reference dichavit.py:529-530: raw pos_embed iff the token count equals the grid and w == h
        use_map = not (cs * pl["N"] == n_pos and W == H)
        dims, ecfg, ep, eacts, pos_map = self._embed_structs(pl, base, scal, C_in, use_map, dev, call)
        check(lib.dcv_embed_fwd(byref(dims), byref(ecfg), byref(ep), c_void_p(x.data_ptr()),
                                c_void_p(idx.data_ptr()) if idx is not None else None, c_void_p(gid.data_ptr()),
                                byref(eacts), st), "dcv_embed_fwd")
        bd = _Dims(B, pl["L"], pl["D"], pl["heads"], pl["F"])
        last = "x0"
        for i in range(pl["depth"]):
            bp, ba, last = self._block_structs(pl, base, i)
            if i == pl["depth"] - 1:  # only the CLS row of the last block's output is consumed
                check(lib.dcv_block_fwd_cls(byref(bd), byref(bp), byref(ba), st), "dcv_block_fwd_cls")
            else:
                check(lib.dcv_block_fwd(byref(bd), byref(bp), byref(ba), st), "dcv_block_fwd")
        D = pl["D"]
        s = pl["arena"].slots
        feat = torch.empty((B, D), dtype=torch.float32, device=dev)
        head = self.classifer_head if isinstance(self.classifer_head, nn.Linear) else None
        ncls = head.out_features if head is not None else 0
        logits = torch.empty((B, ncls), dtype=torch.float32, device=dev) if head is not None else None
        check(lib.dcv_head_fwd(c_void_p(base + s[last]), B, 1, D, c_void_p(self._fptr(fe.norm.weight)),
                               c_void_p(self._fptr(fe.norm.bias)), c_void_p(feat.data_ptr()),
                               c_void_p(base + s["head_mean"]), c_void_p(base + s["head_rstd"]),
                               c_void_p(self._fptr(head.weight)) if head is not None else None,
                               c_void_p(self._fptr(head.bias)) if head is not None else None,
                               c_void_p(logits.data_ptr()) if logits is not None else None, ncls, st), "dcv_head_fwd")
        self.last_losses = {"tdl": scal[0], "cdl": scal[1], "extra": scal[2]}
        state = dict(pl=pl, arena=arena, scal=scal, x=x, idx=idx, gid=gid, feat=feat, last=last, C_in=C_in,
                     use_map=use_map, pos_map=pos_map, call=call) if keep else None
        out = logits if head is not None else feat
        return out, scal[2], state

    def _run_backward(self, state, d_out: Optional[torch.Tensor], d_extra: Optional[torch.Tensor]):
        lib = _lib.lib()
        st = _lib.stream_ptr()
        pl = state["pl"]
        dev = state["x"].device
        B, D, L, M, Fh, heads, Lp, T = pl["B"], pl["D"], pl["L"], pl["M"], pl["F"], pl["heads"], pl["Lp"], pl["T"]
        fe = self.feature_extractor
        pe = fe.patch_embed
        base = state["arena"].data_ptr()
        s = pl["arena"].slots
        # direct_grad + a live accumulation buffer (every .grad still is the view of the flat gradient of the previous
        # backward, nothing reduced yet): the kernels accumulate on top of it -- gradient accumulation over several
        # forward/backward passes (CHAMMI: three chunks per optimiser step) costs no extra pass over the buffer
        static = getattr(self, "_static_gflat", None)  # captured-graph step: one module-owned buffer, zeroed by the graph
        in_place = static is not None or (self.direct_grad and self._accumulation_buffer_live())
        gflat = static if static is not None else (self._last_gflat if in_place else torch.zeros_like(self._flat))
        gb = gflat.data_ptr()

        def gp(p):
            return gb + 4 * self._off[id(p)]

        ws = _Arena()
        ws.add("dres", M * D * 4)
        ws.add("dres_b", M * D * 2)
        ws.add("dh", max(M * Fh * 2, B * T * D * 2))
        ws.add("dv", M * D * 2)
        ws.add("d_o", M * D * 2)
        ws.add("dqkv", M * 3 * D * 2)
        ws.add("delta", B * heads * Lp * 4)
        ws.add("dq_acc", B * heads * L * 64 * 4)
        ws.add("R", L * D * 4)
        ws.add("dpos_patch", pl["N"] * D * 4)
        ws.add("dfeat", B * D * 4)
        ws.add("dres_c", B * D * 4)
        ws.add("dres_c_b", B * D * 2)
        wkey = (B, state["C_in"], pl["H"], pl["W"])
        self._ws_bytes[wkey] = max(self._ws_bytes.get(wkey, 0), ws.off)
        wbuf = ws.alloc(dev, self._ws_bytes[wkey])
        wb = wbuf.data_ptr()
        w = ws.slots
        head = self.classifer_head if isinstance(self.classifer_head, nn.Linear) else None
        ncls = head.out_features if head is not None else 0
        if d_out is None:
            d_out = torch.zeros((B, ncls if head is not None else D), dtype=torch.float32, device=dev)
        d_out = d_out.contiguous().float()
        last_blk = fe.blocks[-1]
        check(lib.dcv_head_bwd(c_void_p(d_out.data_ptr()), c_void_p(base + s[state["last"]]), B, 1, D,
                               c_void_p(self._fptr(fe.norm.weight)), c_void_p(state["feat"].data_ptr()),
                               c_void_p(base + s["head_mean"]), c_void_p(base + s["head_rstd"]),
                               c_void_p(self._fptr(head.weight)) if head is not None else None, ncls,
                               c_void_p(wb + w["dfeat"]), c_void_p(wb + w["dres_c"]), c_void_p(wb + w["dres_c_b"]),
                               c_void_p(gp(fe.norm.weight)), c_void_p(gp(fe.norm.bias)),
                               c_void_p(gp(head.weight)) if head is not None else None,
                               c_void_p(gp(head.bias)) if head is not None else None,
                               c_void_p(gp(last_blk.mlp.fc2.bias)), st), "dcv_head_bwd")
        reducer = _GradReducer(self, gflat) if (self.grad_allreduce and self.grad_sync) else None
        if reducer:
            reducer.ready("tail", flush=True)
        goffs = self._block_param_structs()[1]
        bd = _Dims(B, L, D, heads, Fh)
        bws = _BlockWs(wb + w["dh"], wb + w["dv"], wb + w["d_o"], wb + w["dqkv"], wb + w["delta"], wb + w["dq_acc"])
        for i in reversed(range(pl["depth"])):
            bp, ba, _ = self._block_structs(pl, base, i)
            bg = _BlockGrads(*[gb + o for o in goffs[i]])
            prev_bias = gb + goffs[i - 1][11] if i > 0 else None
            if i == pl["depth"] - 1:
                check(lib.dcv_block_bwd_cls(byref(bd), byref(bp), byref(ba), byref(bg), byref(bws),
                                            c_void_p(wb + w["dres_c"]), c_void_p(wb + w["dres_c_b"]),
                                            c_void_p(wb + w["dres"]), c_void_p(wb + w["dres_b"]),
                                            c_void_p(prev_bias) if prev_bias else None, st), "dcv_block_bwd_cls")
            else:
                check(lib.dcv_block_bwd(byref(bd), byref(bp), byref(ba), byref(bg), byref(bws),
                                        c_void_p(wb + w["dres"]), c_void_p(wb + w["dres_b"]),
                                        c_void_p(prev_bias) if prev_bias else None, st), "dcv_block_bwd")
            if reducer:
                reducer.ready(f"block{i}")
        call = state["call"]
        dims, ecfg, ep, eacts, _ = self._embed_structs(pl, base, state["scal"], state["C_in"], state["use_map"], dev, call,
                                                       backward=True)
        has_prox = hasattr(pe, "channel_emb_proxies")
        # no channel-token gradient when the tokens are frozen (freeze_channel_emb) or were synthesised for unseen
        # channels (eval-time leave-one-out: gid indexes the synthesised matrix, not channel_embed.weight)
        ce_grad = gp(pe.channel_embed.weight) if (pe.channel_embed.weight.requires_grad and call["ce_override"] is None) \
            else None
        eg = _EmbedGrads(gp(pe.proj.weight), gp(pe.proj.bias), ce_grad,
                         gp(pe.channel_emb_proxies) if has_prox else None, gp(fe.cls_token), gp(fe.pos_embed))
        ews = _EmbedWs(wb + w["dh"], wb + w["R"], wb + w["dpos_patch"])
        if d_extra is not None:
            d_extra = d_extra.contiguous().float()
        check(lib.dcv_embed_bwd(byref(dims), byref(ecfg), byref(ep), c_void_p(state["gid"].data_ptr()), byref(eacts),
                                byref(eg), byref(ews), c_void_p(wb + w["dres"]),
                                c_void_p(d_extra.data_ptr()) if d_extra is not None else None, st), "dcv_embed_bwd")
        if getattr(self, "_debug_keep_token_grad", False):  # diagnostics (tools/grad_diag.py): dLoss/d tokens, fp32 [B, L, D]
            off = w["dres"]
            self._last_token_grad = wbuf[off:off + M * D * 4].view(torch.float32).view(B, L, D).clone()
        if reducer:
            reducer.ready("embed", flush=True)
            reducer.finish()
        synced = reducer is not None
        ext = self._external_ids
        if static is not None:  # every .grad already is its view of the static buffer (graphs.py)
            self._last_gflat = gflat
            self._gflat_synced = False
            return None
        if self.direct_grad:
            # skip autograd's 150 AccumulateGrad nodes: .grad of every parameter the kernels own becomes (or accumulates
            # into) a view of the flat buffer.  Tensor hooks / DDP reducer hooks on the parameters do NOT fire in this
            # mode.  `proxies` / `logit_scale` are left to torch autograd (the trainer's loss glue produces them).
            if not in_place:
                prev = self._last_gflat
                if prev is not None and self._grads_alias(prev):
                    prev.add_(gflat)  # a reduced buffer cannot be accumulated into by the kernels: one flat add
                    gflat = prev
                else:
                    views = self._grad_views(gflat)
                    fresh = True
                    for (p, _, _), v in zip(self._layout, views):
                        if not p.requires_grad or id(p) in ext:
                            continue
                        if p.grad is None:
                            p.grad = v
                        else:
                            p.grad.add_(v)
                            fresh = False
                    if not fresh:
                        gflat = None  # gradients live in tensors we do not own: FusedAdamW gathers them
            self._last_gflat = gflat
            self._gflat_synced = synced or (in_place and getattr(self, "_gflat_synced", False))
            return None
        self._last_gflat = gflat  # FusedAdamW consumes the flat buffer directly when .grad still aliases it
        self._gflat_synced = synced
        return [v if (p.requires_grad and id(p) not in ext) else None
                for (p, _, _), v in zip(self._layout, self._grad_views(gflat)) if id(p) not in ext]

    def _grads_alias(self, g: torch.Tensor) -> bool:
        """every trainable parameter the kernels own has .grad == its view of the flat buffer g"""
        base = g.data_ptr()
        ext = self._external_ids
        for p, off, _ in self._layout:
            if not p.requires_grad or id(p) in ext:
                continue
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                return False
        return True

    def _accumulation_buffer_live(self) -> bool:
        g = self._last_gflat
        if g is None or g.numel() != self._flat.numel() or g.device != self._flat.device:
            return False
        if self.grad_allreduce and getattr(self, "_gflat_synced", False):
            return False  # already averaged over the ranks: adding raw local gradients would mix scales
        return self._grads_alias(g)

    def _grad_views(self, gflat: torch.Tensor):
        return [gflat[off:off + n].view(p.shape) for p, off, n in self._layout]


class _GradReducer:
    """Bucketed all-reduce(average) of the flat gradient buffer on a side stream, overlapped with the rest of
    backward (NCCL over NVLink on the GPUs; the same logic runs over gloo on CPU tensors in the tests).
    The flat buffer is laid out embed | block0 .. block11 | tail and backward completes it from the top down,
    so the finished-but-unreduced region is always one contiguous range."""

    BUCKET_BLOCKS = 3

    def __init__(self, module, gflat: torch.Tensor):
        import torch.distributed as dist

        self.dist = dist
        self.m = module
        self.g = gflat
        self.bounds = {n: (a, b) for n, a, b in module._groups}
        self.overlap = bool(getattr(module, "_overlap", True))
        self.world = dist.get_world_size(group=module._pg)
        self.cuda = gflat.is_cuda
        self.stream = None
        if self.cuda:
            if module._comm_stream is None:
                module._comm_stream = torch.cuda.Stream(device=gflat.device)
            self.stream = module._comm_stream
        self.works = []
        self.ranges = []  # (lo, hi) launched, for inspection by tests
        self.lo: Optional[int] = None
        self.hi: Optional[int] = None
        self.n = 0

    def _launch(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        self.ranges.append((lo, hi))
        chunk = self.g[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                chunk.mul_(1.0 / self.world)  # pre-scale: SUM of pre-scaled shards == AVG, NVLS-friendly
                self.works.append(self.dist.all_reduce(chunk, op=self.dist.ReduceOp.SUM, group=self.m._pg,
                                                       async_op=True))
        else:
            chunk.mul_(1.0 / self.world)
            self.works.append(self.dist.all_reduce(chunk, op=self.dist.ReduceOp.SUM, group=self.m._pg, async_op=True))

    def ready(self, name: str, flush: bool = False) -> None:
        """The gradients of parameter group `name` are complete on the current stream."""
        if not self.overlap:
            return
        a, b = self.bounds[name]
        self.lo = a if self.lo is None else min(self.lo, a)
        self.hi = b if self.hi is None else max(self.hi, b)
        self.n += 1
        if flush or self.n >= self.BUCKET_BLOCKS:
            self._launch(self.lo, self.hi)
            self.lo = self.hi = None
            self.n = 0

    def finish(self) -> None:
        if not self.overlap:
            self._launch(0, self.g.numel())
        elif self.lo is not None:
            self._launch(self.lo, self.hi)
        for wk in self.works:
            wk.wait()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)


class _DiChaViTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: DiChaViT, x, cs, idx, gid, need_grad, call, *params):
        out, extra, state = module._run_forward(x, cs, idx, gid, keep=need_grad, call=call)
        ctx.module = module
        ctx.state = state
        ctx.set_materialize_grads(False)
        return out, extra

    @staticmethod
    def backward(ctx, d_out, d_extra):
        if ctx.state is None:
            raise DcvError("backward called on a forward that ran without gradient tracking")
        grads = ctx.module._run_backward(ctx.state, d_out, d_extra)
        ctx.state = None
        if grads is None:  # direct_grad: gradients were written to .grad, the only input is the anchor scalar
            return (None, None, None, None, None, None, None, None)
        return (None, None, None, None, None, None, None, *grads)


class ChannelViTAdapt(DiChaViT):
    """Sibling baseline of the reference (models/channel_vit_adapt.py:ChannelViTAdapt, factory `channelvit_adapt`,
    SURVEY 8(f) #4): the same channel-adaptive ViT without CDL / TDL and with uniform hierarchical channel sampling
    (`random.sample`) instead of DCS; forward returns the logits / features only, in train and eval mode alike."""

    def __init__(self, config, **kwargs):
        class _View(dict):  # the sibling's config has no DiChaViT-specific keys: supply their neutral values
            __getattr__ = dict.__getitem__

        cfg = _View(config)
        cfg.update(proxy_loss_lambda=0, ortho_loss_v1_lambda=0, hcs_sampling="none")
        for k, v in (("hcs_sampling_temp", 0.1), ("gamma_s", 1.0), ("gamma_d", 0.5), ("reverse_pos_pairs", False),
                     ("use_square", False), ("dropout_tokens_hcs", "none"), ("block_type", "block"),
                     ("freeze_channel_emb", False), ("orthogonal_channel_emb_init", False), ("learnable_temp", False)):
            cfg.setdefault(k, v)
        super().__init__(cfg, **kwargs)

    def forward(self, x, chunk_name, training_chunks=None, init_first_layer=None, new_channel_init=None, **kwargs):
        out = super().forward(x, chunk_name, training_chunks, init_first_layer, new_channel_init, **kwargs)
        return out[0] if isinstance(out, tuple) else out


def channelvit_adapt(cfg, **kwargs) -> ChannelViTAdapt:
    return ChannelViTAdapt(config=cfg, **kwargs)


def dichavit(cfg, **kwargs) -> DiChaViT:
    """Factory registered under the reference's name (models/dichavit.py:864-865, models/__init__.py:9)."""
    return DiChaViT(config=cfg, **kwargs)
