"""CPU forward kinematics from the published SMPL/SMPL-X kinematic chain (TEST INFRASTRUCTURE).

**parity unpinned**: the reference delegates FK to the third-party ``smplx`` package
(common/smpl_util.py:13-18,67-69 -> smplx.SMPLX.forward -> smplx.lbs.batch_rigid_transform), which is
neither vendored nor version-pinned, and its model files are licensed.  This port restates the public
definition:

    G_0 = [R_0 | j_0],   G_i = G_parent(i) . [R_i | j_i - j_parent(i)]
    posed joint i = translation column of G_i (+ transl)

with rotations from axis-angle by the quaternion Rodrigues of common/geometry.py:22-65
(theta = ||aa + 1e-8||), which is also how smplx.lbs.batch_rodrigues behaves.
"""
import numpy as np

from .geometry_port import batch_rodrigues


def tree_levels(parents):
    depth = []
    for i, p in enumerate(parents):
        depth.append(0 if p < 0 else depth[p] + 1)
    return depth


def forward_kinematics(rotmats, rest_joints, parents, transl=None):
    """rotmats (F,J,3,3) local rotations; rest_joints (J,3) or (F,J,3); parents list (parents[i] < i).
    Returns (joints (F,J,3), global_R (F,J,3,3)) in the dtype of ``rotmats``."""
    R = np.asarray(rotmats)
    F, J = R.shape[:2]
    dt = R.dtype
    rest = np.broadcast_to(np.asarray(rest_joints, dtype=dt), (F, J, 3))
    gR = np.zeros((F, J, 3, 3), dtype=dt)
    gt = np.zeros((F, J, 3), dtype=dt)
    for i, p in enumerate(parents):
        if p < 0:
            gR[:, i] = R[:, i]
            gt[:, i] = rest[:, i]
        else:
            gR[:, i] = np.einsum("fab,fbc->fac", gR[:, p], R[:, i])
            gt[:, i] = np.einsum("fab,fb->fa", gR[:, p], rest[:, i] - rest[:, p]) + gt[:, p]
    if transl is not None:
        gt = gt + np.asarray(transl, dtype=dt)[:, None, :]
    return gt, gR


def fk_from_axis_angle(aa, rest_joints, parents, transl=None):
    """aa (F,J,3) -> (joints (F,J,3), local rotmats (F,J,3,3), global rotmats (F,J,3,3))."""
    aa = np.asarray(aa)
    F, J = aa.shape[:2]
    R = batch_rodrigues(aa.reshape(-1, 3)).reshape(F, J, 3, 3)
    joints, gR = forward_kinematics(R, rest_joints, parents, transl)
    return joints, R, gR
