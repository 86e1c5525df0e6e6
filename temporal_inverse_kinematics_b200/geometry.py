"""Rotation conversions with the reference's function names, executed by libtik.so.

Mirrors reference common/geometry.py (batch_rodrigues :22-34, rotation_matrix_to_angle_axis :68-97,
rot6d_to_rotmat_spin :308-327, rot6d_to_rotmat :330-344).  fp32 CUDA tensors only; no CPU fallback.
"""
import torch

from . import _lib as L


def _prep(x, last, what):
    if not torch.is_tensor(x):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(x)))
    if not x.is_cuda:
        raise RuntimeError(f"{what}: input must be a CUDA tensor -- there is no CPU fallback")
    if x.numel() % last:
        raise ValueError(f"{what}: input with {x.numel()} elements is not a multiple of {last}")
    x = x.detach().to(torch.float32).contiguous().view(-1, last)
    if x.data_ptr() % 16:            # a contiguous slice (aa[1:]) keeps its storage offset; the kernels use 16-byte vectors
        x = x.clone()
    return x


def _run(fn, x, out_shape, *extra):
    out = torch.empty((x.shape[0],) + out_shape, dtype=torch.float32, device=x.device)
    with L.on_device(x):
        L.check(fn(L.ptr(x), L.ptr(out), x.shape[0], *extra, L.stream_ptr(x.device)))
    return out


def rot6d_to_rotmat(x):
    """(..., 6) -> (M, 3, 3); a1 = x[0::2], a2 = x[1::2]; columns b1, b2, b3."""
    return _run(L.lib().tik_rot6d_to_rotmat, _prep(x, 6, "rot6d_to_rotmat"), (3, 3))


rot6d_to_rotmat_spin = rot6d_to_rotmat   # identical outputs


def batch_rodrigues(axisang):
    """(M, 3) axis-angle -> (M, 9) flat rotation matrices (quaternion Rodrigues, theta = ||aa + 1e-8||)."""
    return _run(L.lib().tik_batch_rodrigues, _prep(axisang, 3, "batch_rodrigues"), (9,))


def rotation_matrix_to_angle_axis(rotation_matrix):
    """(M, 3, 3) [or (M, 3, 4), last column ignored] -> (M, 3), the self-consistent (w,x,y,z) path, NaN -> 0."""
    if torch.is_tensor(rotation_matrix) and rotation_matrix.dim() == 3 and rotation_matrix.shape[1:] == (3, 4):
        rotation_matrix = rotation_matrix[:, :, :3]
    return _run(L.lib().tik_rotmat_to_aa, _prep(rotation_matrix, 9, "rotation_matrix_to_angle_axis"), (3,), 0)


def quat2mat(quat):
    """(M, 4) quaternions (w,x,y,z), normalised first -> (M, 3, 3) (reference common/geometry.py:37-65)."""
    return _run(L.lib().tik_quat_to_rotmat, _prep(quat, 4, "quat2mat"), (3, 3))


def rotation_matrix_to_quaternion(rotation_matrix, eps=1e-6):
    """(M, 3, 3) [or the reference's (M, 3, 4), last column ignored] -> (M, 4) quaternions (w,x,y,z)
    (reference common/geometry.py:153-233; ``eps`` is the reference's 1e-6 threshold on R[2,2], fixed in the kernel)."""
    if eps != 1e-6:
        raise NotImplementedError("rotation_matrix_to_quaternion: only the reference's default eps=1e-6 is built")
    if torch.is_tensor(rotation_matrix) and rotation_matrix.dim() == 3 and rotation_matrix.shape[1:] == (3, 4):
        rotation_matrix = rotation_matrix[:, :, :3]
    return _run(L.lib().tik_rotmat_to_quat, _prep(rotation_matrix, 9, "rotation_matrix_to_quaternion"), (4,))


def quaternion_to_angle_axis(quaternion):
    """(M, 4) quaternions (w,x,y,z) -> (M, 3) axis-angle (reference common/geometry.py:100-150)."""
    return _run(L.lib().tik_quat_to_aa, _prep(quaternion, 4, "quaternion_to_angle_axis"), (3,))
