"""Deterministic synthetic weights / inputs / kinematic tree (no datasets or checkpoints are available offline).

Everything here is generated from ``numpy.random.RandomState`` (a frozen,
bit-stable stream), so the dev container (where goldens are produced with the
real reference) and the GPU box (where only the fixtures travel) build identical
tensors.  Random-init would leave BatchNorm ~identity and ``edge_importance``
== 1 and therefore hide folding bugs, so BN running
statistics, BN affine terms, biases and edge importances are all randomised.

State-dict key names and shapes follow the reference exactly
(SURVEY.md section 5 "Checkpoint"; pose_trainer.py:66-92,
mmskeleton/models/backbones/st_gcn_aaai18.py:52-206).
"""
from collections import OrderedDict

import numpy as np
import torch

# (in_channels, out_channels, temporal_stride, is_residual) -- pose_trainer.py:76-83
POSE_REGRESSOR_LAYERS = [(3, 64, 1, True), (64, 64, 1, True), (64, 128, 2, True), (128, 128, 1, True),
                         (128, 128, 1, True), (128, 128, 2, True), (128, 256, 2, True), (256, 256, 2, True)]

# SMPL-X body kinematic tree, 22 joints (names: bld/syn_motion_videos.py:48-69)
SMPLX_BODY_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _bn(rs, c, prefix, sd):
    sd[prefix + ".weight"] = _t(rs.uniform(0.6, 1.4, c))
    sd[prefix + ".bias"] = _t(rs.standard_normal(c) * 0.1)
    sd[prefix + ".running_mean"] = _t(rs.standard_normal(c) * 0.2)
    sd[prefix + ".running_var"] = _t(rs.uniform(0.5, 1.5, c))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(7, dtype=torch.long)


def make_backbone_state(A, layers=POSE_REGRESSOR_LAYERS, kt=3, seed=0, prefix="backbone."):
    """A: (K,V,V) float array (the Graph adjacency).  Returns an OrderedDict in reference key order."""
    rs = np.random.RandomState(seed)
    K, V, _ = A.shape
    sd = OrderedDict()
    sd[prefix + "A"] = _t(A)
    _bn(rs, layers[0][0] * V, prefix + "data_bn", sd)
    for i, (cin, cout, stride, residual) in enumerate(layers):
        p = f"{prefix}st_gcn_networks.{i}."
        sd[p + "gcn.conv.weight"] = _t(rs.standard_normal((K * cout, cin, 1, 1)) * np.sqrt(2.0 / (cin * K)))
        sd[p + "gcn.conv.bias"] = _t(rs.standard_normal(K * cout) * 0.1)
        _bn(rs, cout, p + "tcn.0", sd)
        sd[p + "tcn.2.weight"] = _t(rs.standard_normal((cout, cout, kt, 1)) * np.sqrt(1.0 / (cout * kt)))
        sd[p + "tcn.2.bias"] = _t(rs.standard_normal(cout) * 0.1)
        _bn(rs, cout, p + "tcn.3", sd)
        if residual and not (cin == cout and stride == 1):
            sd[p + "residual.0.weight"] = _t(rs.standard_normal((cout, cin, 1, 1)) * np.sqrt(1.0 / cin))
            sd[p + "residual.0.bias"] = _t(rs.standard_normal(cout) * 0.1)
            _bn(rs, cout, p + "residual.1", sd)
    for i in range(len(layers)):
        sd[f"{prefix}edge_importance.{i}"] = _t(rs.uniform(0.5, 1.5, (K, V, V)))
    return sd


def make_regressor_state(A, layers=POSE_REGRESSOR_LAYERS, kt=3, seed=0, hidden=512, pose_dim=66):
    """Full PoseRegressor state dict (pose_trainer.py:66-92)."""
    sd = make_backbone_state(A, layers, kt, seed)
    rs = np.random.RandomState(seed + 1000003)
    feat = A.shape[1] * layers[-1][1]
    sd["pose_regressor.0.weight"] = _t(rs.standard_normal((hidden, feat)) * np.sqrt(1.0 / feat))
    sd["pose_regressor.0.bias"] = _t(rs.standard_normal(hidden) * 0.1)
    sd["pose_regressor.3.weight"] = _t(rs.standard_normal((pose_dim, hidden)) * np.sqrt(1.0 / hidden))
    sd["pose_regressor.3.bias"] = _t(rs.standard_normal(pose_dim) * 0.1)
    return sd


def make_iterative_state(A, layers=POSE_REGRESSOR_LAYERS, kt=3, seed=0, hidden=512, njoints=22):
    """State dict of IterativePoseRegressor (the reference's commented-out 6-D head, pose_trainer.py:53-64): backbone,
    fc1 (17*256 + 132 -> 512), fc2 (512 -> 512), decpose (512 -> 132, small init as xavier gain 0.01) and init_pose =
    the 6-D encoding of the identity rotation for every joint (the reference loads SPIN's mean pose from a file)."""
    sd = make_backbone_state(A, layers, kt, seed)
    rs = np.random.RandomState(seed + 2000003)
    feat = A.shape[1] * layers[-1][1]
    npose = njoints * 6
    sd["fc1.weight"] = _t(rs.standard_normal((hidden, feat + npose)) * np.sqrt(1.0 / (feat + npose)))
    sd["fc1.bias"] = _t(rs.standard_normal(hidden) * 0.1)
    sd["fc2.weight"] = _t(rs.standard_normal((hidden, hidden)) * np.sqrt(1.0 / hidden))
    sd["fc2.bias"] = _t(rs.standard_normal(hidden) * 0.1)
    sd["decpose.weight"] = _t(rs.standard_normal((npose, hidden)) * 0.3 * np.sqrt(1.0 / hidden))
    sd["decpose.bias"] = _t(rs.standard_normal(npose) * 0.05)
    sd["init_pose"] = _t(np.tile(np.array([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]), njoints)[None, :])
    return sd


def make_clips(n, t, v=17, c=3, seed=1234, scale=0.3):
    """Root-relative COCO-17 clips (N,T,V,C): mid-hip = 0.5*(kp11+kp12) subtracted
    (mmskeleton/datasets/data_amass.py:232-235)."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, t, v, c)).astype(np.float32) * np.float32(scale)
    if v > 12:
        root = np.float32(0.5) * (x[:, :, 11] + x[:, :, 12])
        x = x - root[:, :, None, :]
    return torch.from_numpy(np.ascontiguousarray(x))


def make_rest_skeleton(parents=SMPLX_BODY_PARENTS, seed=7, sigma=0.1):
    """Synthetic rest joints (J,3): child = parent + N(0, sigma^2) metres."""
    rs = np.random.RandomState(seed)
    j = np.zeros((len(parents), 3), dtype=np.float64)
    for i, p in enumerate(parents):
        off = rs.standard_normal(3) * sigma
        j[i] = off if p < 0 else j[p] + off
    return j.astype(np.float32)


def make_axis_angles(f, j=22, seed=11, scale=0.6):
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((f, j, 3)) * scale).astype(np.float32)


def state_checksum(sd):
    """Order-sensitive fp64 checksum used to detect RNG drift between machines."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += (i + 1) * float(v.double().sum()) + float(v.double().abs().sum())
    return acc
