"""CPU: the C-ABI library loads and exports every symbol of include/dcvit.h; the host-side module mirrors the
reference's interface (names, shapes, init RNG consumption, error behaviour).  No compute calls."""
import ctypes
import random
import re

import numpy as np
import pytest
import torch

from tests.util import CHAMMI_MAPPER, O, cases, ref_cfg

from diverse_channel_vit_b200 import _lib
from diverse_channel_vit_b200 import dichavit as D


def test_library_loads_and_exports_all_symbols():
    lib = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dcv_version() >= 1
    assert lib.dcv_launch_count() >= 0
    assert lib.dcv_profile_num_tags() > 10


def test_invalid_arguments_are_rejected_without_gpu():
    lib = _lib.lib()
    assert lib.dcv_gemm_nt(None, 8, None, 8, 1, 64, 8, 0, None, None, None, None, None, 64, None) == -1
    assert b"null" in lib.dcv_last_error()
    assert lib.dcv_block_fwd(None, None, None, None) == -1
    assert lib.dcv_embed_fwd(None, None, None, None, None, None, None, None) == -1


def _struct_fields(name):
    text = _lib.HEADER_PATH.read_text()
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_0-9]+", part)[-1])
    return fields


@pytest.mark.parametrize("cname,cls", [("dcv_dims", D._Dims), ("dcv_block_params", D._BlockParams),
                                       ("dcv_block_grads", D._BlockGrads), ("dcv_block_acts", D._BlockActs),
                                       ("dcv_block_ws", D._BlockWs), ("dcv_embed_dims", D._EmbedDims),
                                       ("dcv_embed_cfg", D._EmbedCfg), ("dcv_embed_params", D._EmbedParams),
                                       ("dcv_embed_grads", D._EmbedGrads), ("dcv_embed_acts", D._EmbedActs),
                                       ("dcv_embed_ws", D._EmbedWs)])
def test_ctypes_structs_mirror_header(cname, cls):
    assert _struct_fields(cname) == [f[0] for f in cls._fields_]


def test_state_dict_matches_reference_layout():
    for name in ("tiny_chammi_hpa", "tiny_jumpcp", "small_c1"):
        oc, mapper, chunk, has_head, *_ = cases()[name]
        m = D.dichavit(ref_cfg(oc), mapper=mapper)
        sd = m.state_dict()
        want = O.param_shapes(oc, has_head)
        assert set(sd) == set(want)
        for k, shape in want.items():
            assert tuple(sd[k].shape) == tuple(shape), k
        if name == "small_c1":
            assert len(sd) == 154
            assert sum(p.numel() for p in m.parameters()) == 21_483_648  # SURVEY appendix A, 12ch / 14 classes
        assert sd["adaptive_interface.0"].data_ptr() == sd["proxies"].data_ptr()


def test_init_statistics():
    """Init follows the reference: LN = (1, 0), Linear weights trunc-normal(0.02) with zero bias, cls/pos
    trunc-normal(0.02), orthogonal channel tokens when asked, proxies ~ N(0, 1/64)."""
    oc, mapper, *_ = cases()["small_c1"]
    cfg = ref_cfg(oc)
    cfg["orthogonal_channel_emb_init"] = True
    torch.manual_seed(0)
    m = D.dichavit(cfg, mapper=mapper)
    fe = m.feature_extractor
    b = fe.blocks[3]
    assert torch.all(b.norm1.weight == 1) and torch.all(b.norm1.bias == 0)
    assert torch.all(b.attn.qkv.bias == 0) and torch.all(b.mlp.fc2.bias == 0)
    assert abs(b.mlp.fc1.weight.std().item() - 0.02) < 2e-3 and b.mlp.fc1.weight.abs().max() <= 2.0
    assert abs(fe.pos_embed.std().item() - 0.02) < 2e-3
    e = fe.patch_embed.channel_embed.weight
    torch.testing.assert_close(e @ e.t(), torch.eye(12), atol=1e-5, rtol=0)
    assert abs(m.proxies.std().item() - 0.125) < 0.01
    assert m.scale == pytest.approx(np.sqrt(1 / 0.07))


def test_forward_rejects_cpu_tensors_and_unsupported_configs():
    oc, mapper, chunk, *_ = cases()["tiny_chammi_hpa"]
    m = D.dichavit(ref_cfg(oc), mapper=mapper)
    with pytest.raises(_lib.DcvError):
        m(torch.zeros(1, 4, 32, 32), chunk)
    for key, val, exc in (("block_type", "block_v2", NotImplementedError), ("block_type", "nope", ValueError),
                          ("dropout_tokens_hcs", "random", NotImplementedError),
                          ("pretrained_model_name", "huge", ValueError)):
        cfg = ref_cfg(oc)
        cfg[key] = val
        with pytest.raises(exc):
            D.dichavit(cfg, mapper=mapper)


def test_dcs_host_rng_consumption():
    """select_channels draws random.randint(1, C) then random.randint(0, C-1) like the reference
    (dichavit.py:128,154); with sampling off nothing is consumed."""
    oc, mapper, chunk, *_ = cases()["tiny_chammi_hpa"]
    cfg = ref_cfg(oc)
    cfg["enable_sample"] = True
    cfg["hcs_sampling"] = "lowest_cosine_prob"
    m = D.dichavit(cfg, mapper=mapper)
    pe = m.feature_extractor.patch_embed
    m.train()
    random.seed(5)
    torch.manual_seed(7)
    c_new, idx, gid = pe.select_channels("HPA", 4, torch.device("cpu"))
    state_after = random.getstate()
    random.seed(5)
    want_c = random.randint(1, 4)
    want_anchor = random.randint(0, 3)
    assert random.getstate() == state_after
    assert c_new == want_c and idx.numel() == c_new and want_anchor in idx.tolist()
    assert gid.tolist() == [mapper["HPA"][i] for i in idx.tolist()]
    # same draw as the oracle / reference for the same RNG state
    random.seed(5)
    torch.manual_seed(7)
    o = O.dcs_select(pe.channel_embed.weight[torch.tensor(mapper["HPA"])].detach(), cfg.hcs_sampling_temp)
    assert o[2] == idx.tolist()
    assert pe.counter.as_dict() == {g: 1 for g in gid.tolist()}
    m.eval()
    st = random.getstate()
    assert pe.select_channels("HPA", 4, torch.device("cpu"))[1] is None
    assert random.getstate() == st


def test_dcs_prefetch_keeps_the_draw_sequence_cpu():
    """prefetch() moves the RNG calls of the next forward earlier; the sequence of draws is unchanged, a stale or
    mismatching prefetch is dropped."""
    oc, mapper, chunk, *_ = cases()["tiny_chammi_hpa"]
    cfg = ref_cfg(oc)
    cfg["enable_sample"] = True
    cfg["hcs_sampling"] = "lowest_cosine_prob"
    m = D.dichavit(cfg, mapper=mapper).train()
    pe = m.feature_extractor.patch_embed
    dev = torch.device("cpu")

    def draws(prefetch):
        random.seed(3)
        torch.manual_seed(4)
        pe._prefetched = None
        out = []
        for _ in range(10):
            c, idx, gid = pe.select_channels("HPA", 4, dev)
            out.append((c, idx.tolist(), gid.tolist()))
            if prefetch:
                pe.prefetch("HPA", 4, dev)
        return out

    assert draws(False) == draws(True)
    pe.prefetch("HPA", 4, dev)
    assert pe._prefetched is not None
    m.eval()
    assert pe.select_channels("HPA", 4, dev)[1] is None and pe._prefetched is None  # mode changed: dropped
