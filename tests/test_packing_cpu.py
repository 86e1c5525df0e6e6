"""Host logic without a GPU: state_dict compatibility, eval-mode folding (checked by emulating the kernel
sequence on the packed tensors, tests/packed_emulator.py), C-ABI symbol table, error behaviour."""
import ctypes
import os

import numpy as np
import pytest
import torch

import packed_emulator
from oracle import stgcn_port as sp, synth
from temporal_inverse_kinematics_b200 import _lib, engine
from temporal_inverse_kinematics_b200.pose_regressor import IKPoseTrainer, PoseRegressor, default_hparams
from temporal_inverse_kinematics_b200.st_gcn import StgConfig, StgGcn18, StgLayerConfig


def _model(seed=0):
    m = PoseRegressor(default_hparams()).eval()
    sd = synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=seed)
    m.load_state_dict(sd, strict=True)
    return m, sd


def test_state_dict_layout_matches_reference():
    m, sd = _model()
    assert list(m.state_dict().keys()) == list(sd.keys())          # SURVEY.md section 5 key list, reference order
    assert sum(p.numel() for p in m.parameters()) == 3171824
    for i in (0, 2, 5, 6, 7):
        assert f"backbone.st_gcn_networks.{i}.residual.0.weight" in sd
    for i in (1, 3, 4):
        assert f"backbone.st_gcn_networks.{i}.residual.0.weight" not in sd
    t = IKPoseTrainer()
    assert all(k.startswith("regressor.") for k in t.state_dict())


@pytest.mark.parametrize("n,t", [(2, 9), (2, 64), (1, 13)])
def test_fp32_folding_matches_oracle(n, t):
    m, sd = _model()
    packed = engine.PackedNet(m.backbone, m._head(), "fp32")
    x = synth.make_clips(n, t, seed=99)
    got = packed_emulator.forward(packed, x)
    want = sp.regressor_forward(sd, x)["poses"]
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-4


def test_bf16_folding_tolerance():
    m, sd = _model()
    packed = engine.PackedNet(m.backbone, m._head(), "bf16")
    x = synth.make_clips(2, 32, seed=98)
    got = packed_emulator.forward(packed, x)
    want = sp.regressor_forward(sd, x)["poses"]
    err = float((got - want).abs().max())
    assert err < 0.08, err                                            # stated bf16 tolerance, see DESIGN.md


@pytest.mark.parametrize("strategy,max_hop,kt,layers", [
    ("distance", 2, 3, [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)]),
    ("spatial", 1, 3, [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)]),
    ("uniform", 2, 5, [(3, 64, 1, False), (64, 64, 1, True), (64, 128, 2, True)]),
])
def test_backbone_variants_fold(strategy, max_hop, kt, layers):
    graph_cfg = dict(layout="coco", strategy=strategy, max_hop=max_hop, dilation=1)
    cfg = StgConfig(layers=[StgLayerConfig(*l) for l in layers], temporal_kernel_size=kt)
    bb = StgGcn18(cfg, graph_cfg).eval()
    sd = synth.make_backbone_state(sp.build_adjacency("coco", strategy, max_hop, 1), layers, kt=kt, seed=5, prefix="")
    bb.load_state_dict(sd, strict=True)
    packed = engine.PackedNet(bb, None, "fp32")
    x = synth.make_clips(2, 10, seed=77)
    got = packed_emulator.forward(packed, x)
    with torch.no_grad():
        want = sp.backbone_forward(sd, x, layers, prefix="")
    assert float((got - want).abs().max()) < 1e-4


def test_no_cpu_fallback_and_train_mode_raise():
    m, _ = _model()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        m(torch.zeros(1, 9, 17, 3))
    with pytest.raises(NotImplementedError, match="eval"):
        m.train()(torch.zeros(1, 9, 17, 3))
    with pytest.raises(TypeError):
        m.eval()(np.zeros((1, 9, 17, 3)))


def test_c_abi_exports_every_declared_symbol():
    """include/tik.h <-> libtik.so: every declared entry point is exported (no compute calls here)."""
    assert os.path.exists(_lib.LIB_PATH), "build libtik.so first: python -m temporal_inverse_kinematics_b200.build"
    header = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include", "tik.h")).read()
    import re
    declared = sorted(set(re.findall(r"\b(tik_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.exported_symbols()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name
    assert _lib.lib().tik_version() >= 100
    assert ctypes.sizeof(_lib.TikSlab) == 24 and ctypes.sizeof(_lib.TikBlock) == 80
    # static_assert'ed to the same numbers in csrc/pack.cu
    assert (ctypes.sizeof(_lib.TikRawBN), ctypes.sizeof(_lib.TikRawBlock), ctypes.sizeof(_lib.TikPackBuffers)) == (40, 216, 48)


def test_pack_block_sizes_and_argument_errors():
    """tik_pack_block_bytes is host-only: the buffer sizes of the first block (fp32 stem weights, per-node bias, stem
    residual) and of an inner block, and the error codes for shapes the packer rejects."""
    lib = _lib.lib()
    r = _lib.TikRawBlock()
    r.c_in, r.c_out, r.stride, r.kt, r.K, r.V, r.residual = 3, 64, 1, 3, 1, 17, _lib.RES_CONV
    nb = (_lib.i64 * 6)()
    assert lib.tik_pack_block_bytes(ctypes.byref(r), _lib.TIK_BF16, 1, nb) == 0
    assert list(nb) == [17 * 17 * 4, 64 * 3 * 4, 17 * 64 * 4, 64 * (3 * 64 + 64) * 2, 17 * 64 * 4, 17 * 64 * 3 * 4]
    r.c_in = 64
    assert lib.tik_pack_block_bytes(ctypes.byref(r), _lib.TIK_F32, 0, nb) == 0
    assert list(nb) == [17 * 17 * 4, 64 * 64 * 4, 17 * 64 * 4, 64 * (3 * 64 + 64) * 4, 64 * 4, 0]
    assert lib.tik_pack_block_bytes(ctypes.byref(r), _lib.TIK_F32, 1, nb) == _lib._H["TIK_ERR_UNSUPPORTED"]   # c_in > 8 on block 0
    r.residual = 7
    assert lib.tik_pack_block_bytes(ctypes.byref(r), _lib.TIK_F32, 0, nb) == _lib._H["TIK_ERR_INVALID"]
    assert lib.tik_pack_block_bytes(None, _lib.TIK_F32, 0, nb) == _lib._H["TIK_ERR_INVALID"]


def test_keypoint_preprocessing_matches_reference(golden):
    """moveai -> COCO-17 remap + axis swap (reference inference.py:121-133) against the reference-generated fixture."""
    from temporal_inverse_kinematics_b200 import keypoints_util as ku
    g = golden("dance.npz")
    names = [str(s) for s in g["joint_3d_names"]]
    assert np.array_equal(ku.moveai_to_coco(g["joints_3d"], names), g["coco_seq"])
    smplx_names = ["pelvis", "spine1", "spine2"] + ku.COCO17
    assert ku.generate_smplx_to_coco_mappings(smplx_names) == list(range(3, 20))
    assert len(ku.COCO_BONES) == 15
