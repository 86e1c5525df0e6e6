"""Independent cross-check of ``fk_port`` (TEST INFRASTRUCTURE; parity with ``smplx`` itself stays unpinned).

``fk_port.forward_kinematics`` carries rotation and translation separately and gets its rotations from the
quaternion Rodrigues of common/geometry.py:22-65.  This file evaluates the same kinematic chain two other ways so
that a slip in either the chain or the Rodrigues formula cannot hide:

* ``fk_homogeneous``: the literal published SMPL formulation (what ``smplx.lbs.batch_rigid_transform`` documents):
  4x4 matrices ``T_i = [[R_i, j_i - j_parent(i)], [0, 1]]``, ``G_i = G_parent(i) @ T_i``, posed joint = ``G_i[:3, 3]``;
  rotations from ``scipy.spatial.transform.Rotation.from_rotvec`` (scipy 1.18 in this image), float64.
* ``rodrigues_skew``: the matrix form ``R = I + sin(t) K + (1 - cos(t)) K^2`` with ``t = ||aa + 1e-8||`` and
  ``K = skew(aa / t)`` -- the form the published ``smplx.lbs.batch_rodrigues`` uses -- against which the quaternion
  form of the reference (``geometry_port.batch_rodrigues``) is compared.

The reference call sites these stand behind: common/smpl_util.py:61-70 (``smplx_model(global_orient, body_pose, ...)``
-> ``body.joints``).
"""
import numpy as np


def rodrigues_skew(aa):
    aa = np.asarray(aa, dtype=np.float64).reshape(-1, 3)
    t = np.linalg.norm(aa + 1e-8, axis=1, keepdims=True)
    d = aa / t
    K = np.zeros((aa.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -d[:, 2], d[:, 1]
    K[:, 1, 0], K[:, 1, 2] = d[:, 2], -d[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -d[:, 1], d[:, 0]
    s, c = np.sin(t)[:, :, None], np.cos(t)[:, :, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def fk_homogeneous(aa, rest_joints, parents, transl=None):
    """aa (F,J,3) axis-angle -> (joints (F,J,3), global rotations (F,J,3,3)), float64, scipy rotations."""
    from scipy.spatial.transform import Rotation
    aa = np.asarray(aa, dtype=np.float64)
    F, J = aa.shape[:2]
    rest = np.asarray(rest_joints, dtype=np.float64).reshape(J, 3)
    R = Rotation.from_rotvec(aa.reshape(-1, 3)).as_matrix().reshape(F, J, 3, 3)
    G = np.zeros((F, J, 4, 4))
    for i, p in enumerate(parents):
        T = np.zeros((F, 4, 4))
        T[:, :3, :3] = R[:, i]
        T[:, :3, 3] = rest[i] if p < 0 else rest[i] - rest[p]
        T[:, 3, 3] = 1.0
        G[:, i] = T if p < 0 else G[:, p] @ T
    joints = G[:, :, :3, 3].copy()
    if transl is not None:
        joints += np.asarray(transl, dtype=np.float64)[:, None, :]
    return joints, G[:, :, :3, :3].copy()
