"""Size-independent properties at benchmark sizes (the CPU oracle cannot run millions of frames): the conversion and FK
kernels at 2,097,152 frames x 22 joints (46 M rotations) -- orthonormality, det = +1, round trips, agreement between the
two axis-angle -> matrix formulas, bone-length invariance and the rigid-root property of FK -- and the network at the
benchmarked batch: clips are independent, so any slice of a 4096-clip batch must reproduce the full-batch result."""
import numpy as np
import pytest
import torch

from oracle import stgcn_port as sp, synth

pytestmark = pytest.mark.gpu

F, J = 1 << 21, 22


def test_conversions_at_full_size():
    from temporal_inverse_kinematics_b200 import geometry as G, kornia_geometry_conversion as KG
    g = torch.Generator(device="cuda").manual_seed(1)
    aa = torch.randn(F * J, 3, device="cuda", generator=g)
    aa = aa / aa.norm(dim=1, keepdim=True) * (torch.rand(F * J, 1, device="cuda", generator=g) * 3.0 + 0.01)   # |theta| in (0.01, 3.01)
    R = G.batch_rodrigues(aa).view(-1, 3, 3)
    eye = torch.eye(3, device="cuda")
    assert float((R @ R.transpose(1, 2) - eye).abs().max()) < 2e-6
    assert float((torch.linalg.det(R[:: 97]) - 1).abs().max()) < 2e-6
    Rk = KG.angle_axis_to_rotation_matrix(aa)
    assert float((Rk - R).abs().max()) < 5e-6                      # kornia's Rodrigues vs the quaternion form
    back = G.rotation_matrix_to_angle_axis(R)
    assert float((back - aa).abs().max()) < 2e-3                   # round trip (conditioning grows towards pi)
    small = aa.norm(dim=1) < 2.5
    assert float((back[small] - aa[small]).abs().max()) < 2e-4
    x6 = torch.randn(F * J // 4, 6, device="cuda", generator=g)
    R6 = G.rot6d_to_rotmat(x6)
    # Gram-Schmidt is ill-conditioned where a2 is nearly parallel to a1 (the reference normalises rounding noise there
    # too): the orthonormality bound is for the well-conditioned rows, finiteness for all
    a1, a2 = x6[:, 0::2], x6[:, 1::2]
    b1 = a1 / a1.norm(dim=1, keepdim=True)
    u = a2 - (b1 * a2).sum(1, keepdim=True) * b1
    ok = (u.norm(dim=1) > 0.05 * a2.norm(dim=1)) & (a1.norm(dim=1) > 0.05)
    assert float(ok.float().mean()) > 0.99 and torch.isfinite(R6).all()
    assert float((R6[ok] @ R6[ok].transpose(1, 2) - eye).abs().max()) < 1e-5
    again = G.rot6d_to_rotmat(R6[:, :, :2].reshape(-1, 6))         # a rotation's own first two columns reproduce it
    assert float((again[ok] - R6[ok]).abs().max()) < 1e-5


def test_fk_at_full_size():
    from temporal_inverse_kinematics_b200 import geometry as G, smpl_util as SU
    g = torch.Generator(device="cuda").manual_seed(2)
    pose = torch.randn(F, J, 3, device="cuda", generator=g) * 0.6
    rest, parents = synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS
    joints = SU.fk_body(pose, rest, parents)
    assert torch.isfinite(joints).all()
    rest_t = torch.from_numpy(rest).cuda()
    for i, p in enumerate(parents):                                 # bone lengths are pose-invariant
        if p >= 0:
            d = (joints[:, i] - joints[:, p]).norm(dim=1)
            assert float((d - float((rest_t[i] - rest_t[p]).norm())).abs().max()) < 2e-5, i
    assert float((joints[:, 0] - rest_t[0]).abs().max()) < 1e-6     # the root joint stays at its rest position
    # rigid root: rotating only the root rotates the whole posed skeleton about the root joint
    sub = pose[:: 64].clone()
    body = sub.clone()
    body[:, 0] = 0
    j_body = SU.fk_body(body, rest, parents)
    R0 = G.batch_rodrigues(sub[:, 0].contiguous()).view(-1, 3, 3)
    want = torch.einsum("fab,fjb->fja", R0, j_body - rest_t[0]) + rest_t[0]
    assert float((SU.fk_body(sub, rest, parents) - want).abs().max()) < 2e-5
    # local rotations returned next to the joints are the Rodrigues matrices of the pose
    j2, Rl = SU.fk_body(sub, rest, parents, want_local=True)
    assert float((Rl.view(-1, 9) - G.batch_rodrigues(sub.view(-1, 3))).abs().max()) == 0.0


@pytest.mark.parametrize("n,t", [(4096, 64), (2048, 128)])
def test_network_batch_slices_reproduce_the_full_batch(n, t):
    """Eval-mode clips are independent (st_gcn_aaai18.py:113-133): head, middle and tail slices of the benchmarked batch,
    run on their own (other tile counts, other CTA assignments, other chunking), give the full-batch poses."""
    from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams
    m = PoseRegressor(default_hparams()).eval()
    m.load_state_dict(synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=0))
    m = m.cuda().set_compute_dtype("bf16")
    x = synth.make_clips(n, t, seed=77).cuda()
    full = m(x)["poses"]
    assert torch.isfinite(full).all()
    for lo, hi in ((0, 5), (n // 2 - 3, n // 2 + 4), (n - 9, n), (100, 1124)):
        part = m(x[lo:hi].contiguous())["poses"]
        assert torch.equal(part, full[lo:hi]), (lo, hi)
