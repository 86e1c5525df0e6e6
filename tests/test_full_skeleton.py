"""Full 55-joint SMPL-X skeleton + face landmarks: tree definition, pose layout, FK on the GPU vs the
oracle's generic chain, COCO re-projection.  FK parity is *unpinned* (third-party smplx, see oracle/fk_port.py), so the
checks are the oracle restatement plus tree properties."""
import numpy as np
import pytest
import torch

from oracle import fk_port
from temporal_inverse_kinematics_b200 import keypoints_util, smpl_util


def _model():
    return smpl_util.SyntheticBodyModel("neutral", skeleton="full")


def test_full_tree_is_a_valid_smplx_tree():
    m = _model()
    par, names = m.parents, m.joint_names
    assert len(par) == len(names) == 60 and len(set(names)) == 60
    assert par[:22] == smpl_util.SMPLX_BODY_PARENTS
    assert all(p < i for i, p in enumerate(par)) and par.count(-1) == 1
    assert [names[p] for p in par[22:25]] == ["head"] * 3                       # jaw and eyeballs hang off the head
    # every finger chain is wrist -> 1 -> 2 -> 3
    for side, wrist in (("left", 20), ("right", 21)):
        for f in ("index", "middle", "pinky", "ring", "thumb"):
            i1, i2, i3 = (names.index(f"{side}_{f}{k}") for k in (1, 2, 3))
            assert (par[i1], par[i2], par[i3]) == (wrist, i1, i2)
    assert [names[p] for p in par[55:]] == ["head"] * 5 and max(fk_port.tree_levels(par)) == 10
    coco = keypoints_util.generate_smplx_to_coco_mappings(names)
    assert coco[:5] == [55, 57, 56, 59, 58] and coco[9:11] == [20, 21]        # landmarks, then body joints by name


def test_pose_layout_matches_reference_slices():
    rs = np.random.RandomState(0)
    poses = rs.standard_normal((3, 156)).astype(np.float32)
    full = smpl_util.full_pose_from_amass(poses, 60)
    assert np.array_equal(full[:, :22].reshape(3, 66), poses[:, :66])          # root + body   (smpl_util.py:61-62)
    assert np.array_equal(full[:, 25:40].reshape(3, 45), poses[:, 66:111])     # left hand     (:63)
    assert np.array_equal(full[:, 40:55].reshape(3, 45), poses[:, 111:156])    # right hand    (:64)
    assert not full[:, 22:25].any() and not full[:, 55:].any()                 # jaw / eyes / landmarks: identity
    assert not smpl_util.full_pose_from_amass(poses[:, :66], 60)[:, 22:].any()


def test_oracle_landmarks_are_rigid_to_the_head():
    m = _model()
    rs = np.random.RandomState(1)
    aa = (rs.standard_normal((5, 60, 3)) * 0.4).astype(np.float64)
    aa[:, 22:25] = 0
    aa[:, 55:] = 0
    joints, _, gR = fk_port.fk_from_axis_angle(aa, m.rest_joints.astype(np.float64), m.parents)
    rest = m.rest_joints.astype(np.float64)
    for lm in range(55, 60):
        want = joints[:, 15] + np.einsum("fab,b->fa", gR[:, 15], rest[lm] - rest[15])
        assert np.abs(joints[:, lm] - want).max() < 1e-12
    zero, _, _ = fk_port.fk_from_axis_angle(np.zeros((1, 60, 3)), rest, m.parents)
    assert np.abs(zero[0] - rest).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("F", [1, 7, 1000])
def test_gpu_full_skeleton_fk_matches_oracle(F):
    m = _model()
    rs = np.random.RandomState(F)
    aa = (rs.standard_normal((F, 60, 3)) * 0.5).astype(np.float32)
    tr = rs.standard_normal((F, 3)).astype(np.float32)
    want_j, want_l, want_g = fk_port.fk_from_axis_angle(aa.astype(np.float64), m.rest_joints.astype(np.float64), m.parents, tr)
    j, l, g = smpl_util.fk_body(torch.from_numpy(aa).cuda(), m.rest_joints, m.parents, torch.from_numpy(tr).cuda(),
                                want_local=True, want_global=True)
    assert np.abs(j.cpu().numpy() - want_j).max() < 1e-4
    assert np.abs(l.cpu().numpy() - want_l).max() < 1e-5
    assert np.abs(g.cpu().numpy() - want_g).max() < 1e-4
    # joints only, axis-angle in: the thread-per-frame specialisation for exactly this tree
    j1 = smpl_util.fk_body(torch.from_numpy(aa).cuda(), m.rest_joints, m.parents, torch.from_numpy(tr).cuda())
    assert np.abs(j1.cpu().numpy() - want_j).max() < 1e-4
    # rotation-matrix input path, no translation (level-order tree kernel)
    j2 = smpl_util.fk_body(torch.from_numpy(want_l.astype(np.float32)).cuda(), m.rest_joints, m.parents)
    assert np.abs(j2.cpu().numpy() - (want_j - tr[:, None, :])).max() < 1e-4
    # a 33-joint chain (deepest tree the level table has to hold is 24; this one is a star + chain mix)
    par = [-1] + [0] * 16 + list(range(17, 33))
    par = [p if i < 17 else i - 1 for i, p in enumerate(par)]
    par[17] = 3
    assert max(fk_port.tree_levels(par)) <= 23
    rest = rs.standard_normal((33, 3)).astype(np.float32)
    aa3 = (rs.standard_normal((F, 33, 3)) * 0.3).astype(np.float32)
    w3, _, _ = fk_port.fk_from_axis_angle(aa3.astype(np.float64), rest.astype(np.float64), par)
    j3 = smpl_util.fk_body(torch.from_numpy(aa3).cuda(), rest, par)
    assert np.abs(j3.cpu().numpy() - w3).max() < 2e-4


@pytest.mark.gpu
def test_gpu_run_smpl_inference_full_skeleton_and_coco():
    models = smpl_util.load_smplx_models(None, "cuda", 8, skeleton="full")
    m = models["male"]
    rs = np.random.RandomState(3)
    data = {"poses": (rs.standard_normal((50, 156)) * 0.3).astype(np.float32), "gender": "male",
            "trans": rs.standard_normal((50, 3)).astype(np.float32), "betas": rs.standard_normal(16).astype(np.float32)}
    joints = smpl_util.run_smpl_inference(data, models, "cuda")
    assert joints.shape == (50, 60, 3)
    aa = smpl_util.full_pose_from_amass(data["poses"], 60)
    want, _, _ = fk_port.fk_from_axis_angle(aa.astype(np.float64), m.rest(data["betas"][:10]).astype(np.float64), m.parents,
                                            data["trans"])
    assert np.abs(joints - want).max() < 1e-4
    coco = smpl_util.smplx_joints_to_coco(joints, m.joint_names)
    assert coco.shape == (50, 17, 3)
    assert np.array_equal(coco[:, 9], joints[:, 20]) and np.array_equal(coco[:, 0], joints[:, 55])
    body = smpl_util.run_smpl_inference(data, smpl_util.load_smplx_models(None, "cuda", 8), "cuda")
    assert body.shape == (50, 22, 3)
