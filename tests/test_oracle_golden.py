"""The oracle (CPU restatement) against fixtures produced by the REAL reference
(oracle/make_golden.py) and the reference's own known-answer vectors."""
import numpy as np
import pytest
import torch

from oracle import fk_port, geometry_port as gp, stgcn_port as sp, synth


def test_graph_bit_identical(golden):
    g = golden("graph.npz")
    assert len(g.files) == 48
    for key in g.files:
        layout, strategy, max_hop, dilation = key.split("|")
        A = sp.build_adjacency(layout, strategy, int(max_hop), int(dilation))
        assert A.shape == g[key].shape, key
        assert np.array_equal(A, g[key]), key
    A = sp.build_adjacency("coco", "uniform", 2, 1)
    assert A.shape == (1, 17, 17) and int((A != 0).sum()) == 107


def _sd():
    return synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=0)


@pytest.mark.parametrize("tag", ["t9", "t64", "t13", "t1", "t128"])
def test_regressor_matches_reference(golden, tag):
    g = golden("stgcn.npz")
    sd = _sd()
    assert synth.state_checksum(sd) == float(g["state_checksum"]), "synthetic weight stream drifted"
    n, t, seed = [int(v) for v in g[f"{tag}_shape"]]
    x = synth.make_clips(n, t, seed=seed)
    blocks = []
    with torch.no_grad():
        feat = sp.backbone_forward(sd, x, collect=blocks)
    y = sp.regressor_forward(sd, x)["poses"].numpy()
    assert y.shape == g[f"{tag}_poses"].shape
    np.testing.assert_allclose(y, g[f"{tag}_poses"], rtol=0, atol=2e-5)
    stats = np.array([[float(b.double().mean()), float(b.double().abs().mean()), float(b.double().abs().max())]
                      for b in blocks])
    np.testing.assert_allclose(stats, g[f"{tag}_block_stats"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(feat.numpy().reshape(-1)[:2048], g[f"{tag}_feat_head"], rtol=0, atol=2e-5)


def test_backbone_variants_match_reference(golden):
    g = golden("stgcn_variants.npz")
    layers = [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)]
    for strategy, max_hop in [("distance", 2), ("spatial", 1), ("spatial", 2)]:
        A = sp.build_adjacency("coco", strategy, max_hop, 1)
        sd = synth.make_backbone_state(A, layers, kt=3, seed=5, prefix="")
        assert synth.state_checksum(sd) == float(g[f"{strategy}{max_hop}_checksum"])
        with torch.no_grad():
            f = sp.backbone_forward(sd, synth.make_clips(2, 10, seed=77), layers, prefix="").numpy()
        np.testing.assert_allclose(f, g[f"{strategy}{max_hop}_feat"], rtol=0, atol=2e-5)
    layers5 = [(3, 64, 1, False), (64, 64, 1, True), (64, 128, 2, True)]
    sd = synth.make_backbone_state(sp.build_adjacency("coco", "uniform", 2, 1), layers5, kt=5, seed=6, prefix="")
    assert synth.state_checksum(sd) == float(g["kt5_checksum"])
    with torch.no_grad():
        f = sp.backbone_forward(sd, synth.make_clips(2, 11, seed=78), layers5, prefix="").numpy()
    np.testing.assert_allclose(f, g["kt5_feat"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("K", [1, 3])
def test_graph_conv_matches_reference(golden, K):
    g = golden("stgcn_variants.npz")
    t = lambda k: torch.from_numpy(g[f"gconv{K}_{k}"])
    y = sp.graph_conv(t("x"), t("A"), t("w"), t("b")).numpy()
    np.testing.assert_allclose(y, g[f"gconv{K}_y"], rtol=0, atol=1e-5)


def test_geometry_matches_reference(golden):
    g = golden("geometry.npz")
    np.testing.assert_allclose(gp.rot6d_to_rotmat(g["rot6d_in"]), g["rot6d_out"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.rot6d_to_rotmat(g["rot6d_in"][3:]), g["rot6d_spin_out"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.angle_axis_to_rotation_matrix(g["aa_in"]), g["aa_kornia_R"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.batch_rodrigues(g["aa_in"]), g["aa_rodrigues_R9"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.rotation_matrix_to_quaternion(g["R_in"]), g["R_to_quat_wxyz"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.rotation_matrix_to_angle_axis(g["R_in"]), g["R_to_aa"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(gp.rotation_matrix_to_quaternion_kornia_xyzw(g["R_in"]), g["R_to_quat_kornia_xyzw"],
                               rtol=0, atol=1e-6)
    np.testing.assert_allclose(gp.rotation_matrix_to_angle_axis_kornia_quirk(g["R_in"]), g["R_to_aa_kornia_quirk"],
                               rtol=0, atol=2e-6)


def test_kornia_known_answers():
    """The only golden vectors in the reference:
    common/kornia_geometry_conversion.py:322-324 and :350-354."""
    np.testing.assert_allclose(gp.normalize_quaternion(np.array([1., 0., 1., 0.], np.float32)),
                               [0.70710678, 0, 0.70710678, 0], atol=1e-6)
    np.testing.assert_allclose(gp.quaternion_to_rotation_matrix_xyzw(np.array([0., 0., 1., 0.], np.float32))[0],
                               np.diag([-1., -1., 1.]), atol=1e-7)


def test_dance_preprocessing_and_windows(golden):
    g = golden("dance.npz")
    names = [str(s) for s in g["joint_3d_names"]]
    seq = sp.moveai_to_coco(g["joints_3d"], names)
    assert np.array_equal(seq, g["coco_seq"])
    wins = sp.inference_windows(seq, 9)
    assert wins.shape == (231, 9, 17, 3)
    for k, i in enumerate(g["win_idx"]):
        np.testing.assert_allclose(wins[i], g["windows"][k], rtol=0, atol=1e-7)
    sd = _sd()
    y = sp.regressor_forward(sd, torch.from_numpy(wins.astype(np.float32)))["poses"].numpy()
    np.testing.assert_allclose(y, g["poses"], rtol=0, atol=2e-5)


def test_fk_properties():
    """FK parity is unpinned (third-party smplx); property tests per SURVEY.md section 8c."""
    parents = synth.SMPLX_BODY_PARENTS
    rest = synth.make_rest_skeleton()
    F = 5
    zero = np.zeros((F, 22, 3), np.float64)
    j, R, gR = fk_port.fk_from_axis_angle(zero, rest.astype(np.float64), parents)
    np.testing.assert_allclose(j, np.broadcast_to(rest, (F, 22, 3)), atol=1e-6)      # zero pose -> rest joints
    aa = synth.make_axis_angles(F).astype(np.float64)
    j, R, gR = fk_port.fk_from_axis_angle(aa, rest.astype(np.float64), parents)
    for i, p in enumerate(parents):                                                  # bone lengths invariant
        if p >= 0:
            np.testing.assert_allclose(np.linalg.norm(j[:, i] - j[:, p], axis=1),
                                       np.linalg.norm(rest[i] - rest[p]), rtol=1e-6)
    np.testing.assert_allclose(np.einsum("fjab,fjcb->fjac", gR, gR), np.broadcast_to(np.eye(3), (F, 22, 3, 3)),
                               atol=1e-6)
    root_only = np.zeros((F, 22, 3))
    root_only[:, 0] = aa[:, 0]
    j2, R2, _ = fk_port.fk_from_axis_angle(root_only, rest.astype(np.float64), parents)
    expect = np.einsum("fab,jb->fja", R2[:, 0], rest - rest[0]) + rest[0]
    np.testing.assert_allclose(j2, expect, atol=1e-6)                               # rigid rotation about the root
    assert fk_port.tree_levels(parents) == [0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7]
