"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the reference's learned-IK inference hot path
(khanhha/temporal_inverse_kinematics), used as the checker for the CUDA kernels
in ``temporal_inverse_kinematics_b200``.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
anything from here.  The product package never imports ``oracle`` and fails
loudly when its CUDA library is missing.

Modules
-------
stgcn_port      functional fp32 restatement of Graph / ConvTemporalGraphical /
                StGcnBlock / StgGcn18 / PoseRegressor forward (torch CPU ops).
geometry_port   numpy restatement of the rotation conversions.
fk_port         numpy forward kinematics from the published SMPL formula.
ref_import      imports the *real* reference modules from /root/reference with
                namespace stubs (dev container only; never on the GPU box).
make_golden     runs the real reference to write tests/golden/*.npz.

Parity pinning
--------------
* ST-GCN / head / conversions: PINNED -- ``tests/golden/*.npz`` were produced by
  the reference's own modules (``make_golden.py``) and the port is checked
  against them, plus the four kornia docstring known-answer vectors.
* FK: **parity unpinned**.  The arithmetic lives in the third-party ``smplx``
  package (un-vendored, version un-pinned, model files licensed); the port
  restates the published SMPL kinematic chain and is property-tested only.
"""
