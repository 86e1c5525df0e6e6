"""Vendored-kornia conversions of the reference (common/kornia_geometry_conversion.py) on libtik.so.

Only the functions on the IK hot path are provided: angle_axis_to_rotation_matrix (:125-201) and
rotation_matrix_to_angle_axis (:204-227).  The latter is internally inconsistent in the reference
(SURVEY.md section 0.6); ``kornia_quirk=True`` (default, to stay drop-in) reproduces it, ``False`` gives
the self-consistent result of common/geometry.py.
"""
import torch

from . import _lib as L
from .geometry import _prep, _run


def angle_axis_to_rotation_matrix(angle_axis):
    if not isinstance(angle_axis, torch.Tensor):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(angle_axis)))
    if not angle_axis.shape[-1] == 3:
        raise ValueError("Input size must be a (*, 3) tensor. Got {}".format(angle_axis.shape))
    return _run(L.lib().tik_aa_to_rotmat, _prep(angle_axis, 3, "angle_axis_to_rotation_matrix"), (3, 3))


def rotation_matrix_to_angle_axis(rotation_matrix, kornia_quirk=True):
    if not isinstance(rotation_matrix, torch.Tensor):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(rotation_matrix)))
    if not rotation_matrix.shape[-2:] == (3, 3):
        raise ValueError("Input size must be a (*, 3, 3) tensor. Got {}".format(rotation_matrix.shape))
    return _run(L.lib().tik_rotmat_to_aa, _prep(rotation_matrix, 9, "rotation_matrix_to_angle_axis"), (3,),
                int(bool(kornia_quirk)))
