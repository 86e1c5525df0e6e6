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


def window_mean(preds, h_w_size=0):
    """Per-frame mean of overlapping window predictions (reference inference.py:56-67).
    preds (F, T', D): window i contributes preds[i, o + h_w_size] to frame i + o for o in [-h_w_size, h_w_size]."""
    F = preds.shape[0]
    if h_w_size == 0:
        return preds[:, 0]
    if preds.shape[1] < 2 * h_w_size + 1:
        raise ValueError(f"need {2 * h_w_size + 1} predictions per window, the model emits {preds.shape[1]}")
    acc = torch.zeros((F, preds.shape[2]), dtype=preds.dtype, device=preds.device)
    cnt = torch.zeros((F, 1), dtype=preds.dtype, device=preds.device)
    for o in range(-h_w_size, h_w_size + 1):
        lo, hi = max(o, 0), F + min(o, 0)                 # destination frames i + o
        if hi <= lo:
            continue
        acc[lo:hi] += preds[lo - o:hi - o, o + h_w_size]
        cnt[lo:hi] += 1
    return acc / cnt


def run_inference(model, seq_3d_kps, h_w_size=0, relative_pose=True):
    """model: IKPoseTrainer (or anything with .hparams.win_size, .regressor, .device); seq_3d_kps (F, 17, 3).
    Returns (F, 66) axis-angle poses as numpy, like the reference."""
    seq = torch.as_tensor(np.asarray(seq_3d_kps, dtype=np.float32)).to(model.device).contiguous()
    half = model.hparams.win_size // 2
    poses = model.regressor.forward_windows(seq, 2 * half + 1, offset=-half, stride=1,
                                            root=COCO_ROOT_PAIR if relative_pose else None)["poses"]
    return window_mean(poses, h_w_size).detach().cpu().numpy()
