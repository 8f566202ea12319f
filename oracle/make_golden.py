"""Pins the oracle: runs the UNMODIFIED reference module (imported from /root/reference with three
import stubs) and the oracle restatement (oracle/dichavit_oracle.py) on identical
inputs / weights / RNG state, asserts that they agree, and writes small golden vectors to
tests/golden/*.npz.  TEST INFRASTRUCTURE ONLY; runs in the build container (the reference checkout does
not exist on the GPU box -- the committed .npz files travel instead).

    python oracle/make_golden.py            # regenerate every fixture
"""
from __future__ import annotations

import contextlib
import importlib
import io
import math
import os
import random
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import dichavit_oracle as O  # noqa: E402

REF = Path(os.environ.get("DCV_REFERENCE", "/root/reference"))
GOLD = ROOT / "tests" / "golden"


def import_reference():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # utils.py:10
    om = types.ModuleType("omegaconf")
    om.MISSING = "???"
    sys.modules.setdefault("omegaconf", om)  # config.py:6
    sys.path.insert(0, str(REF))
    pkg = types.ModuleType("models")
    pkg.__path__ = [str(REF / "models")]
    sys.modules["models"] = pkg  # skip models/__init__.py (needs timm)
    with contextlib.redirect_stdout(io.StringIO()):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dichavit = importlib.import_module("models.dichavit")
    return dichavit


class Cfg(dict):
    __getattr__ = dict.__getitem__


def ref_cfg(oc: O.OracleConfig) -> Cfg:
    """configs/model/dichavit.yaml + the keys trainer.py:1138-1143 fills in."""
    return Cfg(name="dichavit", pretrained=False, pretrained_model_name=oc.pretrained_model_name, in_dim=None,
               num_classes=oc.num_classes, pooling="avg", temperature=oc.temperature, learnable_temp=False,
               unfreeze_last_n_layers=-1, unfreeze_first_layer=True, init_first_layer=None,
               reset_last_n_unfrozen_layers=False, enable_sample=oc.enable_sample,
               in_channel_names=list(oc.in_channel_names), new_channel_inits=None, use_channelvit_channels=True,
               patch_size=oc.patch_size, orthogonal_channel_emb_init=False, dropout_tokens_hcs="none",
               freeze_channel_emb=False, keep_rate=None, block_type="block", hcs_sampling=oc.hcs_sampling,
               hcs_sampling_temp=oc.hcs_sampling_temp, proxy_loss_lambda=oc.proxy_loss_lambda,
               ortho_loss_v1_lambda=oc.ortho_loss_v1_lambda, drop_path_rate=0.0, gamma_s=oc.gamma_s,
               gamma_d=oc.gamma_d, reverse_pos_pairs=oc.reverse_pos_pairs, use_square=oc.use_square,
               img_size=[oc.img_size])


def build_reference(ref, oc: O.OracleConfig, mapper, weights):
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.dichavit(ref_cfg(oc), mapper=mapper)
    sd = model.state_dict()
    missing = set(sd) - set(weights)
    assert not missing, missing
    model.load_state_dict({k: weights[k].clone() for k in sd}, strict=True)
    return model


def make_inputs(oc: O.OracleConfig, B: int, C: int, n_cls: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, oc.img_size, oc.img_size, generator=g)
    y = torch.randint(0, n_cls, (B,), generator=g)
    return x, y


CHAMMI_MAPPER = {"Allen": [0, 1, 2], "HPA": [3, 4, 5, 6], "CP": [7, 8, 9, 10, 11]}


def cases():
    """name -> (OracleConfig, mapper, chunk, has_head, B, weight seed, input seed, extra_loss_lambda)"""
    names12 = [f"c{i}" for i in range(12)]
    tiny = dict(pretrained_model_name="tiny", img_size=32, patch_size=8)
    out = {}
    out["tiny_chammi_hpa"] = (O.OracleConfig(**tiny, in_channel_names=names12, num_classes=14, temperature=0.07,
                                             proxy_loss_lambda=0.1, ortho_loss_v1_lambda=1.0, gamma_s=0.5, gamma_d=2.0,
                                             reverse_pos_pairs=True), CHAMMI_MAPPER, "HPA", False, 4, 11, 12, 1.0)
    out["tiny_chammi_allen_nolosses"] = (O.OracleConfig(**tiny, in_channel_names=names12, num_classes=14,
                                                        temperature=0.07), CHAMMI_MAPPER, "Allen", False, 3, 13, 14, 1.0)
    out["tiny_jumpcp"] = (O.OracleConfig(**tiny, in_channel_names=[f"c{i}" for i in range(8)], num_classes=10,
                                         proxy_loss_lambda=0.001, ortho_loss_v1_lambda=0.001, gamma_s=1.0, gamma_d=4.0,
                                         reverse_pos_pairs=True), {"train": list(range(8))}, "train", True, 3, 21, 22, 1.0)
    for rp in (False, True):
        for sq in (False, True):
            out[f"tiny_flags_rp{int(rp)}_sq{int(sq)}"] = (
                O.OracleConfig(**tiny, in_channel_names=[f"c{i}" for i in range(5)], num_classes=7,
                               proxy_loss_lambda=0.5, ortho_loss_v1_lambda=0.7, gamma_s=0.5, gamma_d=4.0,
                               reverse_pos_pairs=rp, use_square=sq), {"train": list(range(5))}, "train", True, 2, 31, 32,
                0.5)
    out["small_c1"] = (O.OracleConfig(pretrained_model_name="small", img_size=224, patch_size=16,
                                      in_channel_names=names12, num_classes=14, temperature=0.07,
                                      proxy_loss_lambda=0.1, ortho_loss_v1_lambda=1.0, gamma_s=0.5, gamma_d=2.0,
                                      reverse_pos_pairs=True), CHAMMI_MAPPER, "Allen", False, 8, 41, 42, 1.0)
    # the BENCHED configurations at full model / image size (BASELINE.json configs[2], [3], [4]), small batch
    ch8 = [f"c{i}" for i in range(8)]
    out["full_c3"] = (O.OracleConfig(pretrained_model_name="small", img_size=224, patch_size=16, in_channel_names=ch8,
                                     num_classes=161, proxy_loss_lambda=0.001, ortho_loss_v1_lambda=0.001, gamma_s=1.0,
                                     gamma_d=4.0, reverse_pos_pairs=True, hcs_sampling_temp=1000.0),
                      {"train": list(range(8))}, "train", True, 2, 71, 72, 1.0)
    out["full_c4"] = (O.OracleConfig(pretrained_model_name="small", img_size=32, patch_size=8,
                                     in_channel_names=[f"c{i}" for i in range(18)], num_classes=17,
                                     proxy_loss_lambda=0.001, ortho_loss_v1_lambda=0.1, gamma_s=0.5, gamma_d=4.0,
                                     reverse_pos_pairs=True, hcs_sampling_temp=0.01),
                      {"train": list(range(18))}, "train", True, 8, 73, 74, 1.0)
    out["full_c5"] = (O.OracleConfig(pretrained_model_name="base", img_size=224, patch_size=16, in_channel_names=ch8,
                                     num_classes=161), {"train": list(range(8))}, "train", True, 2, 75, 76, 1.0)
    return out


def summarize_grad(g: torch.Tensor, k: int = 16):
    flat = g.detach().reshape(-1).double()
    gen = torch.Generator().manual_seed(flat.numel())
    pick = torch.randint(0, flat.numel(), (k,), generator=gen)
    return np.array([flat.norm().item(), flat.sum().item()]), flat[pick].numpy()


def run_case(ref, name, spec):
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = spec
    channels = mapper[chunk]
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(channels), oc.num_classes, iseed)
    model = build_reference(ref, oc, mapper, weights)
    model.train()
    with contextlib.redirect_stdout(io.StringIO()):
        out, extra = model(x, chunk)
    import torch.nn.functional as F

    loss_fn = importlib.import_module("models.loss_fn")
    if has_head:
        main = F.cross_entropy(out, y)  # trainer.py:994
    else:
        main = loss_fn.proxy_loss(model.proxies, out, y, model.scale)  # trainer.py:912-914
    loss = main + extra * xlam
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}

    # oracle on the same inputs
    o_loss, o_out, o_grads = O.loss_and_grads(x, y, weights, oc, channels, has_head, extra_loss_lambda=xlam)
    err_out = (o_out.out - out).abs().max().item()
    err_extra = abs(o_out.extra_loss.item() - float(extra))
    err_loss = abs(o_loss.item() - loss.item())
    assert err_out < 5e-5 and err_extra < 1e-6 * max(1, abs(float(extra))) + 1e-6 and err_loss < 2e-5, (
        name, err_out, err_extra, err_loss)
    worst = 0.0
    for k, g in ref_grads.items():
        og = o_grads.get(k)
        if g is None:
            assert og is None or og.abs().max() == 0, k
            continue
        rel = (og - g).norm().item() / max(g.norm().item(), 1e-12)
        worst = max(worst, rel)
        assert rel < 2e-3, (name, k, rel)
    print(f"{name}: oracle vs reference  out {err_out:.2e}  extra {err_extra:.2e}  loss {err_loss:.2e}  "
          f"worst grad rel {worst:.2e}")

    # eval mode: bare tensor, no sampling
    model.eval()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out_eval = model(x, chunk)
    assert isinstance(out_eval, torch.Tensor)

    rec = dict(out=out.detach().numpy(), extra=np.float64(float(extra)), loss=np.float64(loss.item()),
               out_eval=out_eval.numpy(), B=B, wseed=wseed, iseed=iseed, xlam=xlam)
    with torch.no_grad():
        oo = O.forward(x, weights, oc, channels, training=True, has_head=has_head)
    rec["tdl"] = np.float64(oo.tdl.item())
    rec["cdl"] = np.float64(oo.cdl.item())
    rec["tokens_sample"] = oo.tokens[:, :: max(1, oo.tokens.shape[1] // 8), ::16].numpy()
    for k, g in ref_grads.items():
        if g is None:
            continue
        stats, samp = summarize_grad(g)
        rec["gstat:" + k] = stats
        rec["gsamp:" + k] = samp
    np.savez_compressed(GOLD / f"{name}.npz", **rec)


def run_dcs(ref):
    """DCS index selection of the reference (dichavit.py:127-216), CPU generator: for (chunk, seed) record
    C', the final ordered channel positions, and the counter.  The reference converts indices to host lists
    internally; they are recovered by tagging every channel of x with a constant."""
    names12 = [f"c{i}" for i in range(12)]
    recs = {}
    for temp, tname in ((0.1, "t01"), (1000.0, "t1000"), (0.01, "t001")):
        oc = O.OracleConfig(pretrained_model_name="tiny", img_size=16, patch_size=8, in_channel_names=names12,
                            num_classes=14, enable_sample=True, hcs_sampling="lowest_cosine_prob",
                            hcs_sampling_temp=temp, proxy_loss_lambda=0.1)
        weights = O.make_weights(oc, False, 51)
        model = build_reference(ref, oc, CHAMMI_MAPPER, weights)
        model.train()
        pe = model.feature_extractor.patch_embed
        for chunk, chans in CHAMMI_MAPPER.items():
            rows = []
            for seed in range(40):
                random.seed(seed)
                torch.manual_seed(seed + 2)
                C = len(chans)
                x = torch.zeros(1, C, 16, 16)
                for c in range(C):
                    x[:, c] = float(c + 1)
                # capture the gathered x through the conv: hook the proj input
                got = {}

                def hook(mod, inp):
                    got["x"] = inp[0]

                h = pe.proj.register_forward_pre_hook(hook)
                with contextlib.redirect_stdout(io.StringIO()):
                    pe(x, chunk, None, None)
                h.remove()
                idx = [int(round(v)) - 1 for v in got["x"][0, 0, :, 0, 0].tolist()]
                # oracle with the same RNG state
                random.seed(seed)
                torch.manual_seed(seed + 2)
                ce = weights["feature_extractor.patch_embed.channel_embed.weight"][torch.tensor(chans)]
                c_new, anchor, oidx = O.dcs_select(ce, temp, "lowest_cosine_prob")
                assert oidx == idx and c_new == len(idx), (chunk, seed, idx, oidx)
                rows.append([seed, c_new, anchor] + idx + [-1] * (5 - len(idx)))
            recs[f"{tname}:{chunk}"] = np.array(rows, dtype=np.int64)
    np.savez_compressed(GOLD / "dcs_indices.npz", **recs)
    print("dcs: oracle == reference on", sum(len(v) for v in recs.values()), "draws")


LOO_MAPPER = {"train": [0, 1, 2, 3, 4], "test": [2, 5, 3, 6]}
LOO_MODES = ["avg_2", "avg_2_not_in_chunk", "avg_3", "avg_3_not_in_chunk", "replicate", "zero", "random"]


def run_leave_one_out(ref):
    """Eval forward with channels unseen in training (dichavit.py:219-374): reference vs oracle, all modes that do
    not need the never-set `bank` attribute."""
    oc = O.OracleConfig(pretrained_model_name="tiny", img_size=32, patch_size=8,
                        in_channel_names=[f"c{i}" for i in range(7)], num_classes=6, proxy_loss_lambda=0.1,
                        ortho_loss_v1_lambda=0.5)
    weights = O.make_weights(oc, True, 61)
    model = build_reference(ref, oc, LOO_MAPPER, weights)
    model.eval()
    x, _ = make_inputs(oc, 3, 4, oc.num_classes, 62)
    first = importlib.import_module("helper_classes.first_layer_init")
    rec = {}
    for mode in LOO_MODES:
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out = model(x, "test", training_chunks="train", new_channel_init=first.NewChannelLeaveOneOut(mode))
        ce = O.leave_one_out_channel_tokens(weights["feature_extractor.patch_embed.channel_embed.weight"], LOO_MAPPER,
                                            "test", "train", mode)
        with torch.no_grad():
            oo = O.forward(x, weights, oc, LOO_MAPPER["test"], training=False, has_head=True, channel_embed_override=ce)
        assert (oo.out - out).abs().max().item() < 1e-5, mode
        rec[mode] = out.numpy()
    np.savez_compressed(GOLD / "leave_one_out.npz", **rec)
    print("leave-one-out: oracle == reference for", LOO_MODES)


def run_pos_matrix():
    """bicubic positional resample (dichavit.py:518-552) for the grids in use: oracle == reference."""
    ref_mod = sys.modules["models.dichavit"]
    recs = {}
    for grid, img, P in ((14, 224, 16), (4, 32, 8), (2, 16, 8)):
        D = 8
        pos = torch.randn(1, grid * grid + 1, D, generator=torch.Generator().manual_seed(grid))

        class Dummy:
            num_extra_tokens = 1
            patch_embed = types.SimpleNamespace(patch_size=P)
            pos_embed = pos

        xx = torch.zeros(1, 1 + 3 * grid * grid, D)
        got = ref_mod.ChannelVisionTransformer.interpolate_pos_encoding(Dummy(), xx, img, img, 3)
        mine = O.interpolate_pos(pos, 3 * grid * grid, img, img, P, 3)
        assert torch.equal(got, mine)
        recs[f"pos_{grid}"] = pos.numpy()
        recs[f"out_{grid}"] = got[:, : 1 + grid * grid].numpy()
    np.savez_compressed(GOLD / "pos_interp.npz", **recs)
    print("pos interpolation: oracle == reference")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    GOLD.mkdir(parents=True, exist_ok=True)
    ref = import_reference()
    only = sys.argv[1:]
    for name, spec in cases().items():
        if only and name not in only:
            continue
        run_case(ref, name, spec)
    if not only or "dcs" in only:
        run_dcs(ref)
    if not only or "pos" in only:
        run_pos_matrix()
    if not only or "loo" in only:
        run_leave_one_out(ref)


if __name__ == "__main__":
    main()
