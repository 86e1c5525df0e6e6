"""Edge-shape sweep of the bf16 path against the oracle (stated tolerance 0.08 max-abs on the poses): odd clip lengths,
single clips, tile-boundary cases of the fused first block (12 output frames per tile), the halo temporal conv (G row
groups per tile), the paired 64-channel graph conv (odd tile counts) and chunked plans."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import stgcn_port as sp, synth  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402

sd = synth.make_regressor_state(sp.build_adjacency("coco", "uniform", 2, 1), seed=0)
worst = 0.0
cases = [(1, 1), (1, 2), (3, 5), (2, 11), (1, 12), (2, 13), (5, 23), (1, 24), (3, 25), (2, 31), (7, 33), (1, 47), (2, 61),
         (3, 63), (1, 65), (2, 95), (1, 100), (2, 126), (1, 127), (1, 129), (1, 190), (1, 200), (9, 64), (33, 16), (150, 9)]
for n, t in cases:
    for chunk in (None, 2):
        if chunk and n <= chunk:
            continue
        m = PoseRegressor(default_hparams()).eval()
        m.load_state_dict(sd)
        m = m.cuda().set_compute_dtype("bf16")
        if chunk:
            m.chunk_clips = chunk
        x = synth.make_clips(n, t, seed=1000 * n + t)
        want = sp.regressor_forward(sd, x)["poses"]
        got = m(x.cuda())["poses"].cpu()
        err = float((got - want).abs().max())
        ok = got.shape == want.shape and torch.isfinite(got).all() and err < 0.08
        worst = max(worst, err)
        print(f"n={n:4d} T={t:4d} chunk={chunk}: max err {err:.4f} {'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            sys.exit(1)
print(f"shape sweep: PASS (worst {worst:.4f})")
