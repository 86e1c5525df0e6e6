"""Small end-to-end run of every CUDA path (all tensor-core kernels in bf16, the fp32 path, window mode, the FK
variants, conversions): a quick smoke run, and the input for `compute-sanitizer --tool memcheck` where the pool
allows it (the gpurun pool used in round 1 does not)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import geometry as G, smpl_util as SU, synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402

g = Graph(layout="coco", strategy="uniform", max_hop=2, dilation=1)
sd = synth.make_regressor_state(g.A, seed=0)
for dtype in ("bf16", "fp32"):
    m = PoseRegressor(default_hparams()).eval()
    m.load_state_dict(sd)
    m = m.cuda().set_compute_dtype(dtype)
    for n, t in ((3, 64), (2, 9), (5, 32)):
        y = m(synth.make_clips(n, t, seed=n).cuda())["poses"]
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
    seq = synth.make_clips(1, 40, seed=9)[0].cuda()
    y = m.forward_windows(seq, 16, offset=-8, stride=1, root=(11, 12), n_windows=40)["poses"]
    torch.cuda.synchronize()
aa = torch.randn(77, 22, 3, device="cuda")
SU.fk_body(aa, synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS, want_local=True)
mf = SU.SyntheticBodyModel("neutral", skeleton="full")
SU.fk_body(torch.randn(33, 60, 3, device="cuda"), mf.rest_joints, mf.parents)
SU.fk_body(torch.randn(33, 60, 3, device="cuda"), mf.rest_joints, mf.parents, want_global=True)
G.rotation_matrix_to_angle_axis(G.batch_rodrigues(aa.view(-1, 3)).view(-1, 3, 3))
G.rot6d_to_rotmat(torch.randn(100, 6, device="cuda"))
# round 2: bulk-async pipelined kernels (need >= 148 whole tiles of 256 rotations / >= 1 tile of 128 (64) frames)
from temporal_inverse_kinematics_b200 import kornia_geometry_conversion as KG  # noqa: E402
big = torch.randn(148 * 256 * 2 + 77, 3, device="cuda") * 0.7
R = G.batch_rodrigues(big).view(-1, 3, 3)
KG.angle_axis_to_rotation_matrix(big)
G.rotation_matrix_to_angle_axis(R)
G.rot6d_to_rotmat(torch.randn(148 * 256 + 5, 6, device="cuda"))
SU.fk_body(torch.randn(128 * 5 + 9, 22, 3, device="cuda"), synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS,
           transl=torch.randn(128 * 5 + 9, 3, device="cuda"))
SU.fk_body(torch.randn(64 * 7 + 3, 60, 3, device="cuda"), mf.rest_joints, mf.parents)
# latency plan (one persistent cooperative kernel), time-segmented halo conv (T = 128), 15-frame stem tiles
m = PoseRegressor(default_hparams()).eval()
m.load_state_dict(sd)
m = m.cuda().set_compute_dtype("bf16")
m(synth.make_clips(2, 128, seed=4).cuda())
m.low_latency = True
for n, t in ((1, 64), (3, 9), (2, 128)):
    y = m(synth.make_clips(n, t, seed=n).cuda())["poses"]
    assert torch.isfinite(y).all()
y = m.forward_windows(seq, 16, offset=-8, stride=1, root=(11, 12), n_windows=4)["poses"]
torch.cuda.synchronize()
print("sanitize_small: done")
