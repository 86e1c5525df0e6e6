"""Synthetic weights / inputs shared with the product package (see temporal_inverse_kinematics_b200/synthetic.py).
The oracle may import the product's data generators; the product never imports the oracle."""
from temporal_inverse_kinematics_b200.synthetic import *  # noqa: F401,F403
from temporal_inverse_kinematics_b200.synthetic import POSE_REGRESSOR_LAYERS, SMPLX_BODY_PARENTS  # noqa: F401
