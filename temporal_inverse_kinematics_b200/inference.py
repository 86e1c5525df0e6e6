"""Sequence-level inference with the reference's call signature (reference inference.py:37-67), entirely on the GPU.

The reference builds one edge-padded, root-centred window per frame on the host (InferenceDataset, numpy per item),
collates them with a DataLoader (batch 64), runs the model batch by batch with a D2H copy each, and averages the
window predictions per frame in Python lists.  Here the sequence is uploaded once, the windows are *never
materialised* (the stem kernel gathers frames with clamped indices and subtracts the mid-hip root on the fly), all
windows go through the network in one pass, and the per-frame aggregation runs on the device.
"""
import numpy as np
import torch

from .keypoints_util import COCO_ROOT_PAIR


def _window_sum(vals, h_w_size):
    """sum and count over the windows that cover each frame: window i contributes vals[i, o + h] to frame i + o."""
    F = vals.shape[0]
    acc = torch.zeros((F,) + tuple(vals.shape[2:]), dtype=vals.dtype, device=vals.device)
    cnt = torch.zeros((F,) + (1,) * (vals.dim() - 2), dtype=vals.dtype, device=vals.device)
    for o in range(-h_w_size, h_w_size + 1):
        lo, hi = max(o, 0), F + min(o, 0)                 # destination frames i + o
        if hi <= lo:
            continue
        acc[lo:hi] += vals[lo - o:hi - o, o + h_w_size]
        cnt[lo:hi] += 1
    return acc, cnt


def window_mean(preds, h_w_size=0, mode="euclidean"):
    """Per-frame mean of overlapping window predictions (reference inference.py:56-67).
    preds (F, T', D): window i contributes preds[i, o + h_w_size] to frame i + o for o in [-h_w_size, h_w_size].

    mode="euclidean" is the reference's ``np.mean`` of the axis-angle vectors.  mode="rotation" 
    averages ROTATIONS instead: every prediction goes axis-angle -> rotation matrix (`batch_rodrigues` kernel), the
    matrices are averaged per frame and joint, the mean is projected back onto SO(3) by Gram-Schmidt of its first two
    columns (the 6-D representation's mean; `rot6d_to_rotmat` kernel) and returned as axis-angle
    (`rotation_matrix_to_angle_axis` kernel).  Axis-angle vectors near +-pi of the same rotation average to ~0 in
    euclidean mode and to the right rotation here."""
    if mode not in ("euclidean", "rotation"):
        raise ValueError("mode must be 'euclidean' or 'rotation'")
    if h_w_size == 0:
        return preds[:, 0]
    if preds.shape[1] < 2 * h_w_size + 1:
        raise ValueError(f"need {2 * h_w_size + 1} predictions per window, the model emits {preds.shape[1]}")
    if mode == "euclidean":
        acc, cnt = _window_sum(preds, h_w_size)
        return acc / cnt
    from . import geometry
    F, Tp, D = preds.shape
    if D % 3:
        raise ValueError("rotation mode needs axis-angle predictions (D divisible by 3)")
    R = geometry.batch_rodrigues(preds.reshape(-1, 3)).view(F, Tp, D // 3, 9)
    acc, cnt = _window_sum(R, h_w_size)
    M = (acc / cnt).view(-1, 3, 3)
    x6 = M[:, :, :2].reshape(-1, 6)                        # (a1, a2) interleaved as rot6d_to_rotmat reads them
    return geometry.rotation_matrix_to_angle_axis(geometry.rot6d_to_rotmat(x6)).view(F, D)


def run_inference(model, seq_3d_kps, h_w_size=0, relative_pose=True, mean_mode="euclidean"):
    """model: IKPoseTrainer (or anything with .hparams.win_size, .regressor, .device); seq_3d_kps (F, 17, 3).
    Returns (F, 66) axis-angle poses as numpy, like the reference (`h_w_size` / `mean_mode`: see window_mean)."""
    seq = torch.as_tensor(np.asarray(seq_3d_kps, dtype=np.float32)).to(model.device).contiguous()
    half = model.hparams.win_size // 2
    poses = model.regressor.forward_windows(seq, 2 * half + 1, offset=-half, stride=1,
                                            root=COCO_ROOT_PAIR if relative_pose else None)["poses"]
    return window_mean(poses, h_w_size, mean_mode).detach().cpu().numpy()
