import os, sys, torch
sys.path.insert(0, "/root/repo")
from temporal_inverse_kinematics_b200 import synthetic as synth
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams
from temporal_inverse_kinematics_b200.graph import Graph
m = PoseRegressor(default_hparams()).eval()
g = Graph(layout="coco", strategy="uniform", max_hop=2, dilation=1)
m.load_state_dict(synth.make_regressor_state(g.A, seed=0))
m = m.cuda().set_compute_dtype("bf16")
x = synth.make_clips(1, 64, seed=3).cuda()
for _ in range(3): m(x)
torch.cuda.synchronize()
plan = m.plan_for(1, 64)
os.environ["TIK_PLAN_TRACE"] = "1"
plan.profile(x); plan.profile(x)
