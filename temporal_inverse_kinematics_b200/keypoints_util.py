"""Keypoint index conventions and format conversions (host-side numpy, mirrors reference common/keypoints_util.py:5-60
and the COCO-17 order of common/pose_def.py:109-145)."""
import numpy as np

COCO17 = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder", "left_elbow",
          "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip", "left_knee", "right_knee", "left_ankle",
          "right_ankle"]
# 15 COCO bones as index pairs (pose_def.py:137-145), used for drawing / bone-length checks
COCO_BONES = [(0, 1), (1, 3), (0, 2), (2, 4), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10), (5, 11), (11, 13), (13, 15), (6, 12),
              (12, 14), (14, 16)]
COCO_ROOT_PAIR = (11, 12)      # mid-hip = 0.5 * (left_hip + right_hip), data_amass.py:232-235

_MOVEAI = {"left_ear": "L_Ear", "right_ear": "R_Ear", "left_shoulder": "L_Shoulder", "right_shoulder": "R_Shoulder",
           "left_elbow": "L_Elbow", "right_elbow": "R_Elbow", "left_wrist": "L_Wrist", "right_wrist": "R_Wrist",
           "left_hip": "L_Hip", "right_hip": "R_Hip", "left_knee": "L_Knee", "right_knee": "R_Knee",
           "left_ankle": "L_Ankle", "right_ankle": "R_Ankle"}


def generate_smplx_to_coco_mappings(smplx_kps_names):
    """index of every COCO-17 keypoint in an SMPL-X joint-name list."""
    return [smplx_kps_names.index(n) for n in COCO17]


def generate_moveai3d_to_coco_mappings(mvai_3d_joint_names):
    """moveai 3-D joints -> COCO-17; nose and eyes have no counterpart (-1) and are filled by the caller."""
    return [mvai_3d_joint_names.index(_MOVEAI[n]) if n in _MOVEAI else -1 for n in COCO17]


def convert_seq_keypoints(in_seq_kps, mappings, do_copy=False):
    """(B, J, C) -> (B, len(mappings), C) float32 gather; unmapped (-1) targets stay zero."""
    src = np.asarray(in_seq_kps)
    out = np.zeros((src.shape[0], len(mappings), src.shape[2]), dtype=np.float32)
    for tgt, idx in enumerate(mappings):
        if idx >= 0:
            out[:, tgt, :] = src[:, idx, :]
    return out


def moveai_to_coco(joints_3d, joint_names):
    """The pre-processing of reference inference.py:121-133: remap, nose = mean of the ears, eyes = ears, y <- z, z <- -y."""
    j = np.asarray(joints_3d, dtype=np.float32)
    seq = convert_seq_keypoints(j, generate_moveai3d_to_coco_mappings(list(joint_names)))
    seq[:, 0] = 0.5 * (j[:, -1] + j[:, -2])
    seq[:, 1] = j[:, -2]
    seq[:, 2] = j[:, -1]
    y = seq[:, :, 1].copy()
    seq[:, :, 1] = seq[:, :, 2]
    seq[:, :, 2] = -y
    return seq


def moveai_to_coco_device(joints_3d, joint_names):
    """`moveai_to_coco` for a sequence that already lives on the GPU (optional part): the same
    gather / nose = mean of the ears / eyes = ears / axis swap as inference.py:121-133, as ONE gather-and-blend on the
    tensor's device (two index vectors and two weights per COCO keypoint), bit-identical to the numpy version.
    joints_3d: torch tensor (F, J, 3) -> (F, 17, 3) float32 on the same device."""
    import torch
    j = joints_3d.to(torch.float32)
    J = j.shape[1]
    maps = generate_moveai3d_to_coco_mappings(list(joint_names))
    ia, ib, wa, wb = [], [], [], []
    for tgt, idx in enumerate(maps):
        if tgt == 0:        # nose = 0.5 * (joint[-1] + joint[-2])
            ia.append(J - 1); ib.append(J - 2); wa.append(0.5); wb.append(0.5)
        elif tgt == 1:      # left eye = joint[-2]
            ia.append(J - 2); ib.append(J - 2); wa.append(1.0); wb.append(0.0)
        elif tgt == 2:      # right eye = joint[-1]
            ia.append(J - 1); ib.append(J - 1); wa.append(1.0); wb.append(0.0)
        elif idx >= 0:
            ia.append(idx); ib.append(idx); wa.append(1.0); wb.append(0.0)
        else:
            ia.append(0); ib.append(0); wa.append(0.0); wb.append(0.0)
    dev = j.device
    ia, ib = torch.tensor(ia, device=dev), torch.tensor(ib, device=dev)
    wa = torch.tensor(wa, dtype=torch.float32, device=dev)[None, :, None]
    wb = torch.tensor(wb, dtype=torch.float32, device=dev)[None, :, None]
    seq = wa * j[:, ia] + wb * j[:, ib]
    return torch.stack([seq[:, :, 0], seq[:, :, 2], -seq[:, :, 1]], dim=2).contiguous()      # y <- z, z <- -y
