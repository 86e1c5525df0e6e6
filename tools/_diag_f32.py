import sys, torch, time
sys.path.insert(0, '/root/repo')
from temporal_inverse_kinematics_b200 import synthetic as synth
from temporal_inverse_kinematics_b200.graph import Graph
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams
m = PoseRegressor(default_hparams()).eval()
m.load_state_dict(synth.make_regressor_state(Graph("coco","uniform",2,1).A, seed=0))
m = m.cuda().set_compute_dtype("fp32")
x = synth.make_clips(256, 64, seed=1).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for graph in (False, True):
    m.use_cuda_graph = graph
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    for fl in (False, True):
        ts = []
        for _ in range(10):
            if fl: flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); m(x); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"graph={graph} flush={fl}: {sorted(ts)[len(ts)//2]:.3f} ms  (min {min(ts):.3f} max {max(ts):.3f})")
import os
os.environ["TIK_PLAN_TRACE"] = "1"
plan = m.plan_for(256, 64)
print(plan.profile(x))
