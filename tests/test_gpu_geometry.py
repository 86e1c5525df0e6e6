"""CUDA rotation conversions / FK through the C ABI vs the oracle and the reference-generated goldens."""
import numpy as np
import pytest
import torch

from oracle import fk_port, geometry_port as gp, synth

pytestmark = pytest.mark.gpu

TOL = 1e-4      # north_star: <= 1e-4 max-abs on rotation matrices and joint positions in fp32


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_conversions_match_reference_goldens(golden):
    from temporal_inverse_kinematics_b200 import geometry as G, kornia_geometry_conversion as KG
    g = golden("geometry.npz")
    R = G.rot6d_to_rotmat(_cuda(g["rot6d_in"])).cpu().numpy()
    assert np.isfinite(R).all()
    ok = np.ones(len(R), bool)
    ok[1] = False    # a2 parallel to a1: b2 = normalize(rounding noise), ill-conditioned in the reference itself
    np.testing.assert_allclose(R[ok], g["rot6d_out"][ok], rtol=0, atol=1e-5)
    np.testing.assert_allclose(G.rot6d_to_rotmat_spin(_cuda(g["rot6d_in"][3:])).cpu().numpy(), g["rot6d_spin_out"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(KG.angle_axis_to_rotation_matrix(_cuda(g["aa_in"])).cpu().numpy(), g["aa_kornia_R"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(G.batch_rodrigues(_cuda(g["aa_in"])).cpu().numpy(), g["aa_rodrigues_R9"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(G.rotation_matrix_to_angle_axis(_cuda(g["R_in"])).cpu().numpy(), g["R_to_aa"], rtol=0, atol=TOL)
    np.testing.assert_allclose(KG.rotation_matrix_to_angle_axis(_cuda(g["R_in"])).cpu().numpy(), g["R_to_aa_kornia_quirk"], rtol=0, atol=TOL)


@pytest.mark.parametrize("M", [1, 255, 256, 257, 16384 * 22])
def test_conversions_match_oracle(M):
    from temporal_inverse_kinematics_b200 import geometry as G, kornia_geometry_conversion as KG
    rs = np.random.RandomState(M % 1000)
    x6 = rs.standard_normal((M, 6)).astype(np.float32)
    aa = (rs.standard_normal((M, 3)) * 1.0).astype(np.float32)
    R = G.rot6d_to_rotmat(_cuda(x6)).cpu().numpy()
    assert np.abs(R - gp.rot6d_to_rotmat(x6)).max() < TOL   # ill-conditioned samples (a2 nearly parallel to a1) amplify rounding
    Rk = KG.angle_axis_to_rotation_matrix(_cuda(aa)).cpu().numpy()
    assert np.abs(Rk - gp.angle_axis_to_rotation_matrix(aa)).max() < 1e-5
    Rr = G.batch_rodrigues(_cuda(aa)).cpu().numpy()
    assert np.abs(Rr - gp.batch_rodrigues(aa)).max() < 1e-5
    back = G.rotation_matrix_to_angle_axis(_cuda(Rr.reshape(-1, 3, 3))).cpu().numpy()
    assert np.abs(back - gp.rotation_matrix_to_angle_axis(Rr.reshape(-1, 3, 3))).max() < TOL
    # size-independent properties: orthonormal, det +1, aa -> R -> aa round trip for |aa| < pi
    RtR = np.einsum("mab,mac->mbc", R, R)
    assert np.abs(RtR - np.eye(3)).max() < TOL
    assert np.abs(np.linalg.det(R.astype(np.float64)) - 1).max() < TOL
    small = np.linalg.norm(aa, axis=1) < 3.0
    assert np.abs(back[small] - aa[small]).max() < 2e-3


def test_conversion_edge_cases():
    from temporal_inverse_kinematics_b200 import geometry as G
    assert G.rot6d_to_rotmat(torch.zeros(0, 6).cuda()).shape == (0, 3, 3)
    with pytest.raises(RuntimeError):
        G.rot6d_to_rotmat(torch.zeros(4, 6))                          # CPU tensor: no fallback
    with pytest.raises(TypeError):
        G.batch_rodrigues(np.zeros((4, 3), np.float32))


@pytest.mark.parametrize("F", [1, 7, 63, 64, 65, 16384])
def test_fk_matches_oracle(F):
    from temporal_inverse_kinematics_b200 import smpl_util as SU
    parents = synth.SMPLX_BODY_PARENTS
    rest = synth.make_rest_skeleton()
    aa = synth.make_axis_angles(F, seed=F)
    transl = np.random.RandomState(3).standard_normal((F, 3)).astype(np.float32)
    j, lr, gr = SU.fk_body(_cuda(aa), rest, parents, _cuda(transl), want_local=True, want_global=True)
    ej, eR, egR = fk_port.fk_from_axis_angle(aa.astype(np.float64), rest.astype(np.float64), parents, transl.astype(np.float64))
    assert np.abs(j.cpu().numpy() - ej).max() < TOL
    assert np.abs(lr.cpu().numpy() - eR).max() < TOL
    assert np.abs(gr.cpu().numpy() - egR).max() < TOL
    # SMPL-X body fast path (thread per frame): joints only, and joints + local rotations
    j3, lr3 = SU.fk_body(_cuda(aa), rest, parents, _cuda(transl), want_local=True)
    assert np.abs(j3.cpu().numpy() - ej).max() < TOL and np.abs(lr3.cpu().numpy() - eR).max() < TOL
    assert np.abs(SU.fk_body(_cuda(aa), rest, parents).cpu().numpy() - (ej - transl[:, None, :])).max() < TOL
    # rotation-matrix input path (the rot6d head variant feeds matrices)
    j2 = SU.fk_body(lr, rest, parents, _cuda(transl))
    assert np.abs(j2.cpu().numpy() - ej).max() < TOL


def test_fk_properties_large():
    from temporal_inverse_kinematics_b200 import smpl_util as SU
    parents = synth.SMPLX_BODY_PARENTS
    rest = synth.make_rest_skeleton()
    F = 1 << 18
    aa = torch.randn(F, 22, 3, device="cuda") * 0.7
    j = SU.fk_body(aa, rest, parents)
    for i, p in enumerate(parents):                                   # bone lengths are pose-invariant
        if p >= 0:
            d = (j[:, i] - j[:, p]).norm(dim=1)
            assert float((d - float(np.linalg.norm(rest[i] - rest[p]))).abs().max()) < 1e-5
    z = SU.fk_body(torch.zeros(5, 22, 3, device="cuda"), rest, parents)
    assert float((z - torch.from_numpy(rest).cuda()).abs().max()) < 1e-6   # zero pose -> rest joints
    # a chain (depth 21) and a star (depth 1) exercise the pointer-jumping round count
    for par in ([-1] + list(range(21)), [-1] + [0] * 21):
        jj = SU.fk_body(aa[:64], rest, par).cpu().numpy()
        ej, _, _ = fk_port.fk_from_axis_angle(aa[:64].cpu().numpy().astype(np.float64), rest.astype(np.float64), par)
        assert np.abs(jj - ej).max() < TOL


def test_run_smpl_inference_signature():
    from temporal_inverse_kinematics_b200 import smpl_util as SU
    models = SU.load_smplx_models(None, "cuda", 9)
    F = 13
    data = {"poses": np.random.RandomState(0).standard_normal((F, 156)).astype(np.float32) * 0.3, "gender": "male",
            "trans": np.ones((F, 3), np.float32), "betas": np.linspace(-1, 1, 16)}
    j = SU.run_smpl_inference(data, models, "cuda")
    assert j.shape == (F, 22, 3)
    m = models["male"]
    ej, _, _ = fk_port.fk_from_axis_angle(data["poses"][:, :66].reshape(F, 22, 3).astype(np.float64),
                                          m.rest(data["betas"][:10]).astype(np.float64), m.parents, data["trans"].astype(np.float64))
    assert np.abs(j - ej).max() < TOL
    j0 = SU.run_smpl_inference(data, models, "cuda", apply_trans=False, apply_shape=False, apply_root_rot=False)
    assert np.abs(j0[:, 0] - m.rest_joints[0]).max() < 1e-6
    with pytest.raises(NotImplementedError):
        SU.run_smpl_inference(data, models, "cuda", return_mesh=True)


def test_bulk_kernels_write_only_their_outputs():
    """compute-sanitizer is closed on the GPU pool, so the bulk-async (cp.async.bulk) kernels are checked with guard
    regions: outputs are carved out of a larger sentinel-filled buffer through the raw C ABI, for sizes around the
    bulk / remainder split (whole 256-rotation and 128- / 64-frame tiles, plus ragged tails)."""
    import ctypes as C
    from temporal_inverse_kinematics_b200 import _lib as L, geometry as G, smpl_util as SU
    lib = L.lib()
    stream = L.stream_ptr(torch.device("cuda"))
    GUARD, SENT = 4096, 12345.0

    def guarded(n_out):
        buf = torch.full((GUARD + n_out + GUARD,), SENT, dtype=torch.float32, device="cuda")
        return buf, buf[GUARD:GUARD + n_out]

    def check(buf, n_out, what):
        assert bool((buf[:GUARD] == SENT).all()) and bool((buf[GUARD + n_out:] == SENT).all()), what
        assert bool(torch.isfinite(buf[GUARD:GUARD + n_out]).all()) and not bool((buf[GUARD:GUARD + n_out] == SENT).any()), what

    for M in (148 * 256, 148 * 256 + 1, 148 * 256 * 3 + 255, 37):
        x3 = torch.randn(M, 3, device="cuda") * 0.7
        x6 = torch.randn(M, 6, device="cuda")
        for name, fn, x, per in (("rodrigues", lib.tik_batch_rodrigues, x3, 9), ("aa_kornia", lib.tik_aa_to_rotmat, x3, 9),
                                 ("rot6d", lib.tik_rot6d_to_rotmat, x6, 9)):
            buf, out = guarded(M * per)
            L.check(fn(L.ptr(x), L.ptr(out), M, stream))
            torch.cuda.synchronize()
            check(buf, M * per, (name, M))
        R = G.batch_rodrigues(x3)
        buf, out = guarded(M * 3)
        L.check(lib.tik_rotmat_to_aa(L.ptr(R), L.ptr(out), M, 0, stream))
        torch.cuda.synchronize()
        check(buf, M * 3, ("rotmat_to_aa", M))
    rest, parents = synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS
    full = SU.SyntheticBodyModel(skeleton="full")
    for J, rj, par, frames in ((22, rest, parents, (128, 129, 128 * 7 + 5, 127)), (60, full.rest_joints, full.parents, (64, 65, 64 * 5 + 3, 63))):
        rj = np.ascontiguousarray(np.asarray(rj, dtype=np.float32))
        pa = np.ascontiguousarray(np.asarray(par, dtype=np.int32))
        for F in frames:
            pose = torch.randn(F, J, 3, device="cuda") * 0.5
            buf, out = guarded(F * J * 3)
            L.check(lib.tik_fk_body(L.ptr(pose), 0, rj.ctypes.data_as(C.POINTER(C.c_float)), pa.ctypes.data_as(C.POINTER(C.c_int32)), J,
                                    None, L.ptr(out), None, None, F, stream))
            torch.cuda.synchronize()
            check(buf, F * J * 3, ("fk", J, F))
            assert torch.equal(out.view(F, J, 3), SU.fk_body(pose, rj, par))


def test_quaternion_helpers_match_reference(golden):
    """quat2mat / rotation_matrix_to_quaternion / quaternion_to_angle_axis (common/geometry.py:37-65,100-233): the
    building blocks of batch_rodrigues and rotation_matrix_to_angle_axis, exposed as the reference exposes them."""
    from temporal_inverse_kinematics_b200 import geometry as G
    g = golden("geometry.npz")
    R = g["R_in"]
    q = G.rotation_matrix_to_quaternion(_cuda(R)).cpu().numpy()
    np.testing.assert_allclose(q, g["R_to_quat_wxyz"], rtol=0, atol=1e-5)
    aa = G.quaternion_to_angle_axis(_cuda(g["R_to_quat_wxyz"])).cpu().numpy()
    want = gp.quaternion_to_angle_axis(g["R_to_quat_wxyz"])
    ok = np.isfinite(want).all(axis=1)
    np.testing.assert_allclose(aa[ok], want[ok], rtol=0, atol=2e-5)
    # the composition is the R -> aa kernel (NaN rows become 0 there)
    np.testing.assert_allclose(np.nan_to_num(aa[ok]), G.rotation_matrix_to_angle_axis(_cuda(R)).cpu().numpy()[ok], rtol=0, atol=2e-5)
    # quat2mat: un-normalised quaternions, against the oracle's Rodrigues (which is quat2mat of (cos, sin * axis))
    a = synth.make_axis_angles(500, 1, seed=9).reshape(-1, 3)
    th = np.linalg.norm(a + 1e-8, axis=1, keepdims=True)
    quat = np.concatenate([np.cos(th / 2), np.sin(th / 2) * a / th], axis=1).astype(np.float32) * 3.7
    Rq = G.quat2mat(_cuda(quat)).cpu().numpy()
    np.testing.assert_allclose(Rq.reshape(-1, 9), gp.batch_rodrigues(a), rtol=0, atol=1e-5)
    q34 = np.concatenate([R, np.zeros((R.shape[0], 3, 1), dtype=np.float32)], axis=2)   # the reference's (N,3,4) input form
    np.testing.assert_allclose(G.rotation_matrix_to_quaternion(_cuda(q34)).cpu().numpy(), q, rtol=0, atol=0)
