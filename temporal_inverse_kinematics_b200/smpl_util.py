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


# Full SMPL-X skeleton: 22 body joints, jaw, eyes, 2 x 15 finger joints (smplx.joint_names.JOINT_NAMES[:55];
# kinematic tree of the published SMPL-X model), followed by the five face landmarks COCO needs.  In smplx those
# landmarks are mesh vertices; without the licensed mesh they are rigid offsets from the head joint, i.e. extra leaf
# joints with identity rotation ("face landmarks as rigid offsets from head joint when no mesh").
_HAND = ["index1", "index2", "index3", "middle1", "middle2", "middle3", "pinky1", "pinky2", "pinky3",
         "ring1", "ring2", "ring3", "thumb1", "thumb2", "thumb3"]
SMPLX_FULL_JOINT_NAMES = (SMPLX_BODY_JOINT_NAMES + ["jaw", "left_eye_smplhf", "right_eye_smplhf"]
                          + ["left_" + h for h in _HAND] + ["right_" + h for h in _HAND])
SMPLX_FULL_PARENTS = (SMPLX_BODY_PARENTS + [15, 15, 15]
                      + [20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38]
                      + [21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53])
SMPLX_LANDMARK_NAMES = ["nose", "right_eye", "left_eye", "right_ear", "left_ear"]      # smplx JOINT_NAMES[55:60]
SMPLX_LANDMARK_PARENT = 15                                                              # head


def fk_body(pose, rest_joints, parents, transl=None, want_local=False, want_global=False):
    """pose (F,J,3) axis-angle or (F,J,3,3) rotation matrices, CUDA fp32; rest_joints (J,3) host array.
    Returns joints (F,J,3) [, local R (F,J,3,3)] [, global R (F,J,3,3)]."""
    if not (torch.is_tensor(pose) and pose.is_cuda):
        raise RuntimeError("fk_body: pose must be a CUDA tensor -- there is no CPU fallback")
    is_rot = pose.dim() == 4
    pose = pose.detach().float().contiguous()
    if pose.data_ptr() % 16:
        pose = pose.clone()
    F, J = pose.shape[0], pose.shape[1]
    if J > L.MAX_JOINTS:
        raise ValueError(f"fk_body supports at most {L.MAX_JOINTS} joints, got {J}")
    rest = np.ascontiguousarray(np.asarray(rest_joints, dtype=np.float32).reshape(J, 3))
    par = np.ascontiguousarray(np.asarray(parents, dtype=np.int32).reshape(J))
    joints = torch.empty((F, J, 3), dtype=torch.float32, device=pose.device)
    loc = torch.empty((F, J, 3, 3), dtype=torch.float32, device=pose.device) if want_local else None
    glo = torch.empty((F, J, 3, 3), dtype=torch.float32, device=pose.device) if want_global else None
    if transl is not None:
        transl = transl.detach().float().contiguous()
    with L.on_device(pose):
        L.check(L.lib().tik_fk_body(L.ptr(pose), int(is_rot), rest.ctypes.data_as(C.POINTER(C.c_float)),
                                    par.ctypes.data_as(C.POINTER(C.c_int32)), J, L.ptr(transl), L.ptr(joints), L.ptr(loc),
                                    L.ptr(glo), F, L.stream_ptr(pose.device)))
    out = (joints,) + ((loc,) if want_local else ()) + ((glo,) if want_global else ())
    return out[0] if len(out) == 1 else out


class SyntheticBodyModel:
    """Stand-in for an ``smplx`` body model: SMPL-X kinematic tree, seeded rest joints, and joint shape directions so
    that ``betas`` move the rest skeleton linearly (as J_regressor . shapedirs does in SMPL).

    skeleton='body' is the 22-joint tree on the IK path; skeleton='full' is the 55-joint tree plus the five face
    landmarks as rigid children of the head (60 joints), what ``smplx`` returns in ``body.joints[:, :60]``."""

    def __init__(self, gender="neutral", batch_size=1, device="cuda", seed=7, num_betas=10, skeleton="body"):
        rs = np.random.RandomState(seed + {"male": 0, "female": 1, "neutral": 2}[gender])
        self.gender, self.batch_size, self.device = gender, batch_size, device
        if skeleton == "body":
            self.parents, self.joint_names = list(SMPLX_BODY_PARENTS), list(SMPLX_BODY_JOINT_NAMES)
        elif skeleton == "full":
            self.parents = list(SMPLX_FULL_PARENTS) + [SMPLX_LANDMARK_PARENT] * len(SMPLX_LANDMARK_NAMES)
            self.joint_names = list(SMPLX_FULL_JOINT_NAMES) + list(SMPLX_LANDMARK_NAMES)
        else:
            raise ValueError("skeleton must be 'body' or 'full'")
        self.skeleton = skeleton
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


def load_smplx_models(smplx_dir, device, batch_size, skeleton="body"):
    """Same signature as reference common/smpl_util.py:8-19; ``smplx_dir`` is unused (synthetic models).
    skeleton='full' builds the 55-joint + landmark tree (hands, jaw, eyes)."""
    return {g: SyntheticBodyModel(g, batch_size, device, skeleton=skeleton) for g in ("male", "female", "neutral")}


def full_pose_from_amass(poses, n_joints):
    """AMASS / reference pose rows (F, 66 | 156 | 165): root + body [0:66], left hand [66:111], right hand [111:156]
    (common/smpl_util.py:61-64) -> (F, n_joints, 3) in SMPL-X joint order; jaw, eyes and landmark pseudo-joints get a
    zero (identity) rotation exactly as ``smplx`` defaults them."""
    poses = np.asarray(poses, dtype=np.float32)
    F = poses.shape[0]
    out = np.zeros((F, n_joints, 3), dtype=np.float32)
    out[:, :22] = poses[:, :66].reshape(F, 22, 3)
    if n_joints >= 55 and poses.shape[1] >= 156:
        out[:, 25:40] = poses[:, 66:111].reshape(F, 15, 3)
        out[:, 40:55] = poses[:, 111:156].reshape(F, 15, 3)
    return out


def run_smpl_inference(data, smplx_models, device, apply_trans=True, apply_root_rot=True, apply_shape=True,
                       return_mesh=False):
    """data: {'poses': (F, >=66) axis-angle, 'gender', ['trans' (F,3)], ['betas']} -> joints (F, J, 3) numpy
    (J = 22 for body models, 60 for skeleton='full' models: 55 SMPL-X joints + 5 face landmarks).
    Reference common/smpl_util.py:22-82; the fixed-batch padding loop there is unnecessary here (one launch)."""
    if return_mesh:
        raise NotImplementedError("SMPL-X vertices need the licensed mesh template; only joints are produced")
    model = smplx_models[str(data["gender"])]
    J = len(model.parents)
    if J == 22:
        poses = torch.as_tensor(np.asarray(data["poses"], dtype=np.float32)[:, :66]).to(device).view(-1, 22, 3).contiguous()
    else:                                                         # full skeleton: body + hands (+ identity jaw / eyes / landmarks)
        poses = torch.as_tensor(full_pose_from_amass(data["poses"], J)).to(device)
    if not apply_root_rot:
        poses = poses.clone()
        poses[:, 0] = 0
    transl = torch.as_tensor(np.asarray(data["trans"], dtype=np.float32)).to(device) if apply_trans else None
    rest = model.rest(np.asarray(data["betas"])[:10] if apply_shape else None)
    return fk_body(poses, rest, model.parents, transl).cpu().numpy()


def smplx_joints_to_coco(joints, joint_names):
    """(F, J, 3) joints of the full skeleton -> (F, 17, 3) COCO keypoints, the gather of
    common/keypoints_util.py:27-60 applied to FK output."""
    from .keypoints_util import convert_seq_keypoints, generate_smplx_to_coco_mappings
    return convert_seq_keypoints(joints, generate_smplx_to_coco_mappings(list(joint_names)))
