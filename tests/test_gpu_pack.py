"""tik_pack_bn / tik_pack_block (weight folding through the C ABI) against the host-side packer: the
folded tensors must be BIT-identical, and a network packed by the library gives the same poses."""
import pytest
import torch

from oracle import stgcn_port as sp, synth

pytestmark = pytest.mark.gpu


def _backbone(layers, kt, strategy="uniform", max_hop=2, seed=5, data_bn=True, importance=True):
    from temporal_inverse_kinematics_b200.st_gcn import StgConfig, StgGcn18, StgLayerConfig
    cfg = StgConfig(layers=[StgLayerConfig(*l) for l in layers], temporal_kernel_size=kt)
    bb = StgGcn18(cfg, dict(layout="coco", strategy=strategy, max_hop=max_hop, dilation=1), data_bn=data_bn,
                  edge_importance_weighting=importance).eval()
    sd = synth.make_backbone_state(sp.build_adjacency("coco", strategy, max_hop, 1), layers, kt=kt, seed=seed, prefix="")
    sd = {k: v for k, v in sd.items() if k in bb.state_dict()}
    bb.load_state_dict(sd, strict=True)
    return bb.cuda()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_pose_regressor_packs_bit_identically(dtype):
    from temporal_inverse_kinematics_b200 import engine
    from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams
    m = PoseRegressor(default_hparams()).eval()
    m.load_state_dict(synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=3), strict=True)
    m = m.cuda()
    host = engine.PackedNet(m.backbone, m._head(), dtype, packer="host")
    dev = engine.PackedNet(m.backbone, m._head(), dtype, packer="device")
    torch.cuda.synchronize()
    assert sorted(host.named) == sorted(dev.named)
    for k, v in host.named.items():
        assert v.dtype == dev.named[k].dtype and v.numel() == dev.named[k].numel(), k
        assert torch.equal(v.reshape(-1), dev.named[k].reshape(-1)), k
    for i in range(host.net.n_blocks):
        a, b = host.net.blocks[i], dev.net.blocks[i]
        assert (a.c_in, a.c_out, a.stride, a.kt, a.res_kind, a.res_as_slab) == (b.c_in, b.c_out, b.stride, b.kt, b.res_kind, b.res_as_slab)


@pytest.mark.parametrize("strategy,max_hop,kt,layers,data_bn", [
    ("distance", 2, 3, [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)], True),
    ("spatial", 1, 3, [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)], False),
    ("uniform", 2, 5, [(3, 64, 1, False), (64, 64, 1, True), (64, 128, 2, False)], True),
])
def test_backbone_variants_pack_bit_identically(strategy, max_hop, kt, layers, data_bn):
    """K = 1..3 adjacency partitions, kt = 5, blocks without a residual, no data_bn."""
    from temporal_inverse_kinematics_b200 import engine
    bb = _backbone(layers, kt, strategy, max_hop, data_bn=data_bn)
    host = engine.PackedNet(bb, None, "fp32", packer="host")
    dev = engine.PackedNet(bb, None, "fp32", packer="device")
    torch.cuda.synchronize()
    for k, v in host.named.items():
        assert torch.equal(v.reshape(-1), dev.named[k].reshape(-1)), k


def test_network_packed_by_the_library_gives_the_same_poses(monkeypatch):
    from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams
    sd = synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=0)
    x = synth.make_clips(6, 64, seed=4).cuda()
    outs = {}
    for packer in ("host", "device"):
        monkeypatch.setenv("TIK_PACKER", packer)
        for dtype in ("fp32", "bf16"):
            m = PoseRegressor(default_hparams()).eval()
            m.load_state_dict(sd, strict=True)
            m = m.cuda().set_compute_dtype(dtype)
            outs[packer, dtype] = m(x)["poses"].clone()
            assert m._engine._packed[dtype].packer == packer
    for dtype in ("fp32", "bf16"):
        assert torch.equal(outs["host", dtype], outs["device", dtype])
    want = sp.regressor_forward(sd, x.cpu())["poses"]
    assert float((outs["device", "fp32"].cpu() - want).abs().max()) < 1e-4


def test_pack_block_rejects_what_the_host_packer_rejects():
    import ctypes as C
    from temporal_inverse_kinematics_b200 import _lib as L, engine
    bb = _backbone([(3, 64, 1, True), (64, 64, 1, True)], 3)
    raw = engine.raw_block(bb, 1)
    nbytes = (L.i64 * 6)()
    assert L.lib().tik_pack_block_bytes(C.byref(raw), L.TIK_BF16, 0, nbytes) == 0
    assert list(nbytes) == [17 * 17 * 4, 64 * 64 * 2, 17 * 64 * 4, 64 * (3 * 64 + 64) * 2, 64 * 4, 0]
    assert L.lib().tik_pack_block_bytes(C.byref(raw), L.TIK_BF16, 1, nbytes) == L._H["TIK_ERR_UNSUPPORTED"]     # identity residual on block 0
    raw.kt = 4
    assert L.lib().tik_pack_block_bytes(C.byref(raw), L.TIK_F32, 0, nbytes) == L._H["TIK_ERR_INVALID"]
    assert b"kt 4" in L.lib().tik_last_error()
