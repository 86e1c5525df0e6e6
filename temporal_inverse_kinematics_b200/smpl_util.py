"""Body-model forward kinematics behind the reference's ``run_smpl_inference`` signature.

Mirrors reference common/smpl_util.py:8-82.  The reference delegates to the third-party ``smplx``
package and licensed SMPL-X model files; neither is available offline, so ``load_smplx_models`` builds
*synthetic* body models (seeded kinematic tree + rest joints + linear joint shape directions) and the
joints come from the warp-per-frame FK kernel in libtik.so.  Vertices (``return_mesh=True``) need the
licensed mesh template and are out of scope.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

# SMPL-X body joints 0..21 (names: reference bld/syn_motion_videos.py:48-69)
SMPLX_BODY_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]
SMPLX_BODY_JOINT_NAMES = ["pelvis", "left_hip", "right_hip", "spine1", "left_knee", "right_knee", "spine2",
                          "left_ankle", "right_ankle", "spine3", "left_foot", "right_foot", "neck", "left_collar",
                          "right_collar", "head", "left_shoulder", "right_shoulder", "left_elbow", "right_elbow",
                          "left_wrist", "right_wrist"]


def fk_body(pose, rest_joints, parents, transl=None, want_local=False, want_global=False):
    """pose (F,J,3) axis-angle or (F,J,3,3) rotation matrices, CUDA fp32; rest_joints (J,3) host array.
    Returns joints (F,J,3) [, local R (F,J,3,3)] [, global R (F,J,3,3)]."""
    if not (torch.is_tensor(pose) and pose.is_cuda):
        raise RuntimeError("fk_body: pose must be a CUDA tensor -- there is no CPU fallback")
    is_rot = pose.dim() == 4
    pose = pose.detach().float().contiguous()
    F, J = pose.shape[0], pose.shape[1]
    rest = np.ascontiguousarray(np.asarray(rest_joints, dtype=np.float32).reshape(J, 3))
    par = np.ascontiguousarray(np.asarray(parents, dtype=np.int32).reshape(J))
    joints = torch.empty((F, J, 3), dtype=torch.float32, device=pose.device)
    loc = torch.empty((F, J, 3, 3), dtype=torch.float32, device=pose.device) if want_local else None
    glo = torch.empty((F, J, 3, 3), dtype=torch.float32, device=pose.device) if want_global else None
    if transl is not None:
        transl = transl.detach().float().contiguous()
    L.check(L.lib().tik_fk_body(L.ptr(pose), int(is_rot), rest.ctypes.data_as(C.POINTER(C.c_float)),
                                par.ctypes.data_as(C.POINTER(C.c_int32)), J, L.ptr(transl), L.ptr(joints), L.ptr(loc),
                                L.ptr(glo), F, L.stream_ptr(pose.device)))
    out = (joints,) + ((loc,) if want_local else ()) + ((glo,) if want_global else ())
    return out[0] if len(out) == 1 else out


class SyntheticBodyModel:
    """Stand-in for an ``smplx`` body model: 22-joint SMPL-X body tree, seeded rest joints, and joint shape
    directions so that ``betas`` move the rest skeleton linearly (as J_regressor . shapedirs does in SMPL)."""

    def __init__(self, gender="neutral", batch_size=1, device="cuda", seed=7, num_betas=10):
        rs = np.random.RandomState(seed + {"male": 0, "female": 1, "neutral": 2}[gender])
        self.gender, self.batch_size, self.device = gender, batch_size, device
        self.parents = list(SMPLX_BODY_PARENTS)
        J = len(self.parents)
        rest = np.zeros((J, 3))
        for i, p in enumerate(self.parents):
            off = rs.standard_normal(3) * 0.1
            rest[i] = off if p < 0 else rest[p] + off
        self.rest_joints = rest.astype(np.float32)
        self.joint_shapedirs = (rs.standard_normal((J, 3, num_betas)) * 0.01).astype(np.float32)
        self.faces = None

    def to(self, device):
        self.device = device
        return self

    def rest(self, betas=None):
        if betas is None:
            return self.rest_joints
        return self.rest_joints + self.joint_shapedirs @ np.asarray(betas, dtype=np.float32)[: self.joint_shapedirs.shape[2]]


def load_smplx_models(smplx_dir, device, batch_size):
    """Same signature as reference common/smpl_util.py:8-19; ``smplx_dir`` is unused (synthetic models)."""
    return {g: SyntheticBodyModel(g, batch_size, device) for g in ("male", "female", "neutral")}


def run_smpl_inference(data, smplx_models, device, apply_trans=True, apply_root_rot=True, apply_shape=True,
                       return_mesh=False):
    """data: {'poses': (F, >=66) axis-angle, 'gender', ['trans' (F,3)], ['betas']} -> joints (F, 22, 3) numpy.
    Reference common/smpl_util.py:22-82; the fixed-batch padding loop there is unnecessary here (one launch)."""
    if return_mesh:
        raise NotImplementedError("SMPL-X vertices need the licensed mesh template; only joints are produced")
    model = smplx_models[str(data["gender"])]
    poses = torch.as_tensor(np.asarray(data["poses"], dtype=np.float32)[:, :66]).to(device).view(-1, 22, 3).contiguous()
    if not apply_root_rot:
        poses = poses.clone()
        poses[:, 0] = 0
    transl = torch.as_tensor(np.asarray(data["trans"], dtype=np.float32)).to(device) if apply_trans else None
    rest = model.rest(np.asarray(data["betas"])[:10] if apply_shape else None)
    return fk_body(poses, rest, model.parents, transl).cpu().numpy()
