"""clock64 timeline of the fused first-block kernel (stem_block.cu), CTA 0, tile iteration 4.  Needs a TIK_PROBE build:
TIK_PROBE=1 python -m temporal_inverse_kinematics_b200.build"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import _lib, synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402

NAMES = {0: "setup", 1: "B.top", 2: "B.landed", 3: "B.a0_empty", 4: "B.published", 5: "B.end", 6: "M.A.top", 7: "M.a0_full",
         8: "M.tmem_empty", 9: "M.3.top", 10: "M.h_full", 11: "M.3.issued", 12: "E.mid.top", 13: "E.da_full", 14: "E.ld_done",
         15: "E.h_empty", 16: "E.mid.done", 17: "E.fin.top", 18: "E.d3_full", 19: "E.fin.done", 20: "S.stage_full",
         21: "S.read_done", 22: "S.next_read_done", 23: "E.mid.stored", 24: "E.mid.fenced", 25: "B.aggregated",
         26: "B.stored", 27: "B.fenced", 28: "B.prefetched", 29: "B.copies_done"}

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    hp = default_hparams()
    m = PoseRegressor(hp).eval()
    g = Graph(layout="coco", strategy="uniform", max_hop=2, dilation=1)
    m.load_state_dict(synth.make_regressor_state(g.A, seed=0))
    m = m.cuda().set_compute_dtype("bf16")
    x = synth.make_clips(n, 64, seed=3).cuda()
    lib = _lib.lib()
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    tb = torch.zeros(64, dtype=torch.int64, device="cuda")
    lib.tik_debug_set_umma_times(_lib.ptr(tb))
    m(x)
    torch.cuda.synchronize()
    lib.tik_debug_set_umma_times(None)
    tl = tb.cpu().tolist()[32:]        # slots 0-31 belong to the rowgemm kernel's probes
    ev = sorted((tl[i] - tl[0], NAMES[i]) for i in NAMES if tl[i])
    for t, name in ev:
        print(f"{t:9d}  {name}")
