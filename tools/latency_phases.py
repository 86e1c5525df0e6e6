"""Per-phase timeline of the latency plan (csrc/latency.cu): CTA 0 stamps %globaltimer at every phase start.
    python tools/latency_phases.py [T=64] [N=1]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import _lib as L, synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = PoseRegressor(default_hparams()).eval()
m.load_state_dict(synth.make_regressor_state(Graph("coco", "uniform", 2, 1).A, seed=0))
m = m.cuda()
m.low_latency = True
x = synth.make_clips(N, T, seed=3).cuda()
plan = m.plan_for(N, T)
grid = 148 * int(os.environ.get("TIK_LAT_CTAS_PER_SM", "1"))
buf = torch.zeros(grid * 2 * plan.phases, dtype=torch.int64, device="cuda")
L.check(L.lib().tik_debug_latency_times(plan.handle, C.c_void_p(buf.data_ptr())))
rows = []
for i in range(60):
    m(x)
    torch.cuda.synchronize()
    rows.append(buf.cpu().numpy().reshape(grid, plan.phases, 2).copy())
L.check(L.lib().tik_debug_latency_times(plan.handle, None))
t = np.array(rows[10:]).astype(np.float64) / 1e3                    # (runs, grid, phases, 2) us
start, end = t[..., 0], t[..., 1]
work_max = np.median((end - start).max(axis=1), axis=0)              # slowest CTA's work per phase
work_med = np.median(np.median(end - start, axis=1), axis=0)
first_start = start.min(axis=1)                                       # (runs, phases)
last_end = end.max(axis=1)
span = np.median(np.diff(np.concatenate([first_start, last_end[:, -1:]], axis=1), axis=1), axis=0)
barrier = np.median(first_start[:, 1:] - last_end[:, :-1], axis=0)    # last arrival -> first CTA released
names = ["stem", "b0.tcn"] + [f"b{i}.{k}" for i in range(1, 8) for k in ("gcn", "tcn")] + ["head1", "head2"]
print(f"{'phase':8s} {'span':>7s} {'work max':>9s} {'work med':>9s} {'barrier':>8s}   (us)")
for i, n in enumerate(names):
    print(f"{n:8s} {span[i]:7.1f} {work_max[i]:9.1f} {work_med[i]:9.1f} {barrier[i] if i < len(barrier) else 0.0:8.1f}")
print(f"total    {span.sum():7.1f} us  work(max) {work_max.sum():.1f}  barriers {barrier.sum():.1f}  ({plan.phases} phases, N={N}, T={T})")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    m(x)
e1.record()
e1.synchronize()
print(f"back-to-back launches: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per window")
