"""BASELINE.json configs[4]: one long sequence (T=8192), sliding windows of 64 frames, stride 1.

  bulk      all windows in one pass (no window tensor is materialised: the stem kernel gathers frames)
  ref-win   the reference's own semantics: 65-frame edge-padded, root-centred windows, one per frame
  stream    one window per step through the public API with CUDA-graph replay: p50 / p99 latency
Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from temporal_inverse_kinematics_b200 import synthetic as synth  # noqa: E402
from temporal_inverse_kinematics_b200.graph import Graph  # noqa: E402
from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams  # noqa: E402


def ev_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    F, W = 8192, 64
    model = PoseRegressor(default_hparams()).eval()
    model.load_state_dict(synth.make_regressor_state(Graph("coco", "uniform", 2, 1).A, seed=0))
    model = model.cuda().set_compute_dtype(dtype)
    seq = synth.make_clips(1, F, seed=5)[0].cuda()
    out = {"sequence_frames": F, "window": W, "dtype": dtype}

    n_win = F - W + 1
    t = ev_time(lambda: model.forward_windows(seq, W, offset=0, stride=1, root=(11, 12)), 10)
    out["bulk"] = {"windows": n_win, "ms": t * 1e3, "windows_per_s": n_win / t, "window_frames_per_s": n_win * W / t}
    t = ev_time(lambda: model.forward_windows(seq, 65, offset=-32, stride=1, root=(11, 12)), 10)
    out["reference_windows_65_edge_padded"] = {"windows": F, "ms": t * 1e3, "windows_per_s": F / t}

    # streaming: one window per step, public API, graph replay; latency on the device (events) and on the host clock
    model.use_cuda_graph = True
    xs = [synth.make_clips(1, W, seed=100 + i).cuda() for i in range(8)]
    for x in xs:
        model(x)
    torch.cuda.synchronize()
    dev_us, host_us = [], []
    for i in range(2000):
        x = xs[i % 8]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        y = model(x)["poses"]
        e1.record()
        e1.synchronize()
        host_us.append((time.perf_counter() - t0) * 1e6)
        dev_us.append(e0.elapsed_time(e1) * 1e3)
    q = lambda a, p: float(np.percentile(np.array(a), p))
    out["stream_batch1"] = {"steps": 2000, "device_us_p50": q(dev_us, 50), "device_us_p99": q(dev_us, 99),
                            "host_us_p50": q(host_us, 50), "host_us_p99": q(host_us, 99)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
