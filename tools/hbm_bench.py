"""Stand-alone bandwidth numbers for the HBM-bound kernels (conversions, FK) at F = 8,388,608 frames x 22 joints
(SURVEY.md section 8d): achieved GB/s = algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import geometry as G, kornia_geometry_conversion as KG, smpl_util as SU, synthetic  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e-3


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
    once = "--once" in sys.argv          # one launch per kernel after a warm-up pass: the shape an ncu capture wants
    global timed
    if once:
        def timed(fn, reps=1):
            fn()
            torch.cuda.synchronize()
            return 1.0
    J = 22
    peak = 6534.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    M = F * J
    x6 = torch.randn(M, 6, device="cuda")
    aa = torch.randn(M, 3, device="cuda") * 0.7
    R = G.batch_rodrigues(aa).view(-1, 3, 3)
    rest, parents = synthetic.make_rest_skeleton(), synthetic.SMPLX_BODY_PARENTS
    pose = aa.view(F, J, 3)
    out = {}
    for name, fn, byts in [
        ("rot6d_to_rotmat", lambda: G.rot6d_to_rotmat(x6), M * 60),
        ("angle_axis_to_rotation_matrix", lambda: KG.angle_axis_to_rotation_matrix(aa), M * 48),
        ("batch_rodrigues", lambda: G.batch_rodrigues(aa), M * 48),
        ("rotation_matrix_to_angle_axis", lambda: G.rotation_matrix_to_angle_axis(R), M * 48),
        ("fk_body joints", lambda: SU.fk_body(pose, rest, parents), F * 528),
        ("fk_body joints+local_R", lambda: SU.fk_body(pose, rest, parents, want_local=True), F * 1320),
    ]:
        t = timed(fn)
        gbs = byts / t / 1e9
        out[name] = {"ms": t * 1e3, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak, "frames_per_s": F / t}
        print(f"{name:34s} {t * 1e3:8.3f} ms  {gbs:8.1f} GB/s  frac {gbs / peak:5.3f}  {F / t / 1e9:6.3f} G frames/s", flush=True)
    # full SMPL-X skeleton (55 joints + 5 landmarks, level-order tree kernel), F/4 frames: 720 B in + 720 B out per frame
    del x6, R
    m = SU.SyntheticBodyModel("neutral", skeleton="full")
    F2 = F // 4
    pose60 = torch.randn(F2, 60, 3, device="cuda") * 0.5
    t = timed(lambda: SU.fk_body(pose60, m.rest_joints, m.parents))
    gbs = F2 * 1440 / t / 1e9
    out["fk_body full skeleton (60 joints)"] = {"ms": t * 1e3, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak, "frames_per_s": F2 / t}
    print(f"{'fk_body full skeleton (60 joints)':34s} {t * 1e3:8.3f} ms  {gbs:8.1f} GB/s  frac {gbs / peak:5.3f}  {F2 / t / 1e9:6.3f} G frames/s", flush=True)
    json.dump({"F": F, "J": J, "hbm_peak_gbs": peak, "kernels": out}, open(os.path.join(ROOT, "gpurun_out", "hbm_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
