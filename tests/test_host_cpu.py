"""CPU-side checks: C-ABI exports, host logic, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib, build
from depth_image_captioning_pub_b200.engine import batch_sizes_from_lengths

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(build.LIB)
    header = open(os.path.join(ROOT, "include", "dic.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dic_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().dic_version() == 100


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    d = _lib.Dims(196, 2048, 128, 128, 128, 10000)
    assert lib.dic_pack_bytes(ctypes.byref(d), _lib.DIC_BF16) > 0
    assert lib.dic_train_workspace_bytes(ctypes.byref(d), _lib.DIC_BF16, 256, 20) > 0
    assert lib.dic_decode_workspace_bytes(ctypes.byref(d), _lib.DIC_F32, 128, 5) > 0
    bad = _lib.Dims(196, 2044, 128, 128, 128, 10000)     # D % 8 != 0
    assert lib.dic_pack_bytes(ctypes.byref(bad), _lib.DIC_F32) == 0
    assert b"D%8" in lib.dic_last_error()


def test_batch_sizes_from_lengths():
    assert batch_sizes_from_lengths([7, 5, 4]) == [3, 3, 3, 2, 1, 1]
    assert batch_sizes_from_lengths([21] * 4) == [4] * 20
    with pytest.raises(ValueError):
        batch_sizes_from_lengths([4, 5])
    with pytest.raises(ValueError):
        batch_sizes_from_lengths([3, 1])


def test_state_dict_matches_reference_layout():
    from conftest import load_golden
    _, w, _ = load_golden("depth_soft")
    A, D = w["attention.encoder_att.weight"].shape
    V, E = w["embed.weight"].shape
    H = w["decode_step.weight_hh"].shape[1]
    for cls, extra in ((P.CD_RNNDecoderWithSoftAttention, ()), (P.RNNDecoderWithSoftAttention, ()),
                       (P.CD_RNNDecoderWithHardAttention, ("cuda:0",)), (P.RNNDecoderWithHardAttention, ("cuda:0",))):
        m = cls(A, E, D, H, V, *extra)
        assert list(m.state_dict().keys()) == list(w.keys())
        m.load_state_dict(w)          # strict: shapes and names must agree
        assert list(_lib.PARAM_KEYS) == list(w.keys())


def test_no_cpu_fallback():
    m = P.CD_RNNDecoderWithSoftAttention(32, 16, 32, 32, 53)
    f = torch.rand(2, 196, 32)
    caps = torch.zeros(2, 4, dtype=torch.int64)
    with pytest.raises(P.DicError):
        m(f, f, caps, [4, 3])
    with pytest.raises(P.DicError):
        m.batch_sample(f, f, {"<start>": 49, "<end>": 50}, max_length=3)
    att = P.Soft_Attention(32, 32, 32)
    with pytest.raises(P.DicError):
        att(f, torch.rand(2, 32))


def test_hard_attention_noise_matches_reference_draw_order():
    # one torch.rand(sum(bs), k) == the reference's per-step torch.rand(bs_valid, k) concatenated
    torch.manual_seed(11)
    per_step = torch.cat([torch.rand(n, 196) for n in (3, 3, 2, 1)])
    torch.manual_seed(11)
    assert torch.equal(per_step, torch.rand(9, 196))


def test_depth_encoder_state_dict_keys_match_reference():
    """Depth_CNN_endoder keeps the reference's state_dict keys (depth_models.py:12-47, incl. the features.N aliases):
    checkpoints of depth_train.py:306-322 load unchanged."""
    import os
    import numpy as np
    import depth_image_captioning_pub_b200 as P
    rec = np.load(os.path.join(os.path.dirname(__file__), "golden", "depth_encoder.npz"))
    m = P.Depth_CNN_endoder(14)
    assert sorted(m.state_dict().keys()) == [str(k) for k in rec["keys"]]
    with __import__("pytest").raises(Exception):
        m(__import__("torch").zeros(1, 1, 224, 224))          # CPU tensors: no fallback
