"""Per-kernel times of the fp32 plan at B=256, T=64 (TIK_PLAN_TRACE), and the forward time with / without the L2 flush."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402

m = PoseRegressor(default_hparams()).eval()
m.load_state_dict(synth.make_regressor_state(Graph("coco", "uniform", 2, 1).A, seed=0))
m = m.cuda().set_compute_dtype("fp32")
x = synth.make_clips(256, 64, seed=1).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m(x); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
print(f"fp32 forward B=256 T=64: {sorted(ts)[5]:.3f} ms")
os.environ["TIK_PLAN_TRACE"] = "1"
kinds, flops = m.plan_for(256, 64).profile(x)
print(kinds, f"{flops / (kinds['gemm'][0] * 1e-3) / 1e12:.1f} TFLOP/s on the GEMM launches")
