"""Where the host time of one low-latency forward goes (microseconds per call, 3000 calls each)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import engine, synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402

m = PoseRegressor(default_hparams()).eval()
m.load_state_dict(synth.make_regressor_state(Graph("coco", "uniform", 2, 1).A, seed=0))
m = m.cuda()
m.low_latency = True
x = synth.make_clips(1, 64, seed=3).cuda()
m(x)
torch.cuda.synchronize()
plan = m.plan_for(1, 64)


def t(fn, n=3000):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


e = m._engine
print(f"require_cuda_eval        {t(lambda: engine.require_cuda_eval(m, x, 'x')):7.1f} us")
print(f"_check_input             {t(lambda: m.backbone._check_input(x)):7.1f} us")
print(f"stamp (_refresh)         {t(lambda: e._refresh(m.backbone, m._head())):7.1f} us")
print(f"plan_for                 {t(lambda: m.plan_for(1, 64)):7.1f} us")
print(f"torch.empty              {t(lambda: torch.empty((1, 4, 66), dtype=torch.float32, device=x.device)):7.1f} us")
print(f"plan.run (launch incl.)  {t(lambda: plan.run(x)):7.1f} us   <- GPU-bound when the kernel (119 us) is the longer side")
print(f"model(x)                 {t(lambda: m(x)):7.1f} us")
