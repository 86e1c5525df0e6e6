#!/usr/bin/env python
"""Headline benchmark: IK-solved frames/s (ST-GCN forward + SMPL-X body FK) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

Workload (BASELINE.json configs[2]): B=4096 clips of T=64 root-relative COCO-17 frames per GPU, bf16
tensor-core ST-GCN forward (PoseRegressor) + 22-joint body FK on the B*T/16 solved poses.  A "step" is one
pass of that path over one synthetic batch.  One process per GPU; with N>1 every rank works on its own shard
(weak scaling) and the solved poses are all-gathered over NCCL once per step.  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port of its PyTorch
modules, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T_FRAMES, V, C_IN, J = 4096, 64, 17, 3, 22
METRIC = "IK-solved frames/sec (ST-GCN fwd + SMPL-X FK)"
UNIT = "frames/s"
CPU_SAMPLE_CLIPS = 256


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there when NCCL_DEBUG is
    set) must not interleave with it: keep a private handle on the real stdout and point fd 1 at stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def _ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per tensor-core launch, averaged over one step of the committed
    ncu capture (profiles/r1_launches_final.csv); None if the capture is absent."""
    import csv
    p = os.path.join(ROOT, "profiles", "r1_launches_final.csv")
    if not os.path.exists(p):
        return None
    per = {}
    with open(p) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    for r in rows[1:]:
        if any(k in r[ki] for k in ("rowgemm_umma", "rowgemm_ts", "tcn_halo", "gcn_fused", "stem_block")) and r[mi].startswith("dram__bytes"):
            per[r[ii]] = per.get(r[ii], 0.0) + float(r[vi].replace(",", ""))
    return sum(per.values()) / len(per) if per else None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.i0 = 0

    def mark(self):
        """Everything sampled from now on is inside the timed region (the process is started earlier: nvidia-smi
        needs tens of ms to produce its first row)."""
        self.i0 = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = self.rows[self.i0:] or self.rows
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _oracle_model():
    from oracle import stgcn_port as sp, synth
    A = sp.build_adjacency("coco", "uniform", 2, 1)
    return sp, synth, synth.make_regressor_state(A, seed=0)


def cpu_reference_step(sd, sp, fk_port, x, rest, parents):
    """The reference's CPU path for one batch: PoseRegressor forward (oracle port of the PyTorch modules) + FK."""
    poses = sp.regressor_forward(sd, x)["poses"]
    aa = poses.reshape(-1, J, 3).numpy()
    joints, _, _ = fk_port.fk_from_axis_angle(aa, rest, parents)
    return poses, joints


def time_cpu(clips, iters, warm=1):
    import torch
    from oracle import fk_port
    sp, synth, sd = _oracle_model()
    x = synth.make_clips(clips, T_FRAMES, seed=1234)
    rest, parents = synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS
    for _ in range(warm):
        cpu_reference_step(sd, sp, fk_port, x, rest, parents)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        cpu_reference_step(sd, sp, fk_port, x, rest, parents)
        ts.append(time.perf_counter() - t0)
    return clips * T_FRAMES / (sum(ts) / len(ts)), sum(ts) / len(ts), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    fps, sec, threads = time_cpu(CPU_SAMPLE_CLIPS, max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    sample = f"{CPU_SAMPLE_CLIPS} clips x T={T_FRAMES} per step (bounded sample of the B={B_PER_GPU} batch), fp32, torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: B={B_PER_GPU} clips/GPU, T={T_FRAMES}, ST-GCN fwd + 22-joint FK", "sample": sample},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from temporal_inverse_kinematics_b200 import _lib, smpl_util, synthetic as synth
    from temporal_inverse_kinematics_b200.graph import Graph
    from temporal_inverse_kinematics_b200.pose_regressor import PoseRegressor, default_hparams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().tik_check_device())

    B, T = args.batch, T_FRAMES
    model = PoseRegressor(default_hparams()).eval()
    model.load_state_dict(synth.make_regressor_state(Graph("coco", "uniform", 2, 1).A, seed=0))
    model = model.to(dev).set_compute_dtype(args.dtype)
    if args.chunk:
        model.chunk_clips = args.chunk
    model.use_cuda_graph = not args.no_graph
    rest, parents = synth.make_rest_skeleton(), synth.SMPLX_BODY_PARENTS

    x_host = synth.make_clips(B, T, seed=1234 + rank).pin_memory()
    x_dev = x_host.to(dev)
    T_out = model.backbone.out_frames(T)
    gathered = torch.empty((world * B, T_out, 66), device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(x):
        poses = model(x)["poses"]                                # (B, T', 66) axis-angle
        joints = smpl_util.fk_body(poses.view(-1, J, 3), rest, parents)
        if world > 1:
            dist.all_gather_into_tensor(gathered, poses)
        return poses, joints

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                          # running before the warm-up; rows are kept from mark() on
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    barrier()

    # ---- device-resident timing: CUDA events per step on the launching stream, L2 flushed between steps
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step(x_dev)
        b.record()
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps

    # ---- end to end through the public API: pinned host input -> H2D -> forward + FK -> D2H of the results
    out_p = torch.empty((B, T_out, 66), dtype=torch.float32).pin_memory()
    out_j = torch.empty((B * T_out, J, 3), dtype=torch.float32).pin_memory()

    # Every step copies its input from pinned host memory and reads poses + joints back.  The copies are pipelined the
    # way a serving loop would: the H2D copy of step i+1 runs on a copy stream while step i computes (two device
    # input buffers), the D2H read-back of step i runs on that stream while step i+1 computes.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    x_bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the buffer's previous reader has finished
            x_bufs[i % 2].copy_(x_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(n_steps):
        for ev in consumed:
            ev.record(main_stream)
        upload(0)
        for i in range(n_steps):
            main_stream.wait_event(ready[i % 2])
            poses, joints = step(x_bufs[i % 2])
            consumed[i % 2].record(main_stream)
            done[i % 2].record(main_stream)
            if i + 1 < n_steps:
                upload(i + 1)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i % 2])
                out_p.copy_(poses, non_blocking=True)
                out_j.copy_(joints, non_blocking=True)
                poses.record_stream(copy_stream)
                joints.record_stream(copy_stream)
        main_stream.wait_stream(copy_stream)

    e2e_run(3)                                                   # warm-up of the pipelined loop (pinned buffers, copy stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks, peak_src = _peaks()
        plan = model.plan_for(B, T)
        kinds, gemm_flops = plan.profile(x_dev)                  # CUDA events around every kernel (one extra run)
        g_ms, g_n = kinds["gemm"]
        achieved = gemm_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"] if args.dtype == "bf16" else 74.0
        launches_per_step = plan.launches(B) + 1                 # + FK
        roofline = {"bound": "tensor", "kernel": "tcgen05 family: rowgemm_umma_kernel / rowgemm_ts_kernel / tcn_halo_kernel / gcn_fused_kernel / stem_block_kernel" if args.dtype == "bf16" else "rowgemm_f32_kernel",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": (peak_src + " bf16_tflops_sustained (kernel timed inside a long step)") if args.dtype == "bf16" else "nominal fp32 SIMT 148 SM x 128 FMA x 2 x 1.965 GHz",
                    "launches": g_n, "avg_launch_us": g_ms * 1e3 / max(g_n, 1), "flops_per_step": gemm_flops,
                    "kernel_ms_per_step": {k: v[0] for k, v in kinds.items()}, "traffic": _ncu_traffic_per_launch()}
        if roofline["traffic"] and g_ms > 0:
            # context: the same launches against the HBM roofline (DRAM bytes from the committed ncu capture / live time)
            gbs = roofline["traffic"] / (roofline["avg_launch_us"] * 1e-6) / 1e9
            roofline["dram"] = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]}
        line = {"metric": METRIC, "value": world * B * T / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"configs[2]: B={B} clips/GPU, T={T}, {args.dtype} ST-GCN fwd (PoseRegressor, random-init "
                                       f"weights) + 22-joint body FK on {B * T_out} solved poses/GPU",
                           "global_batch": world * B, "frames_per_step": world * B * T, "chunk_clips": plan.n_chunk, "cuda_graph": bool(model.use_cuda_graph),
                           "l2": "flushed between timed steps (256 MiB memset)", "parallelism": f"dp{world}",
                           "gather": "NCCL all_gather of poses each step" if world > 1 else "none",
                           "e2e_pipeline": "pinned H2D of step i+1 and D2H of step i overlap compute on a copy stream"},
                "e2e": {"value": world * B * T / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": world * x_host.numel() * 4,
                        "d2h_bytes_per_step": world * (out_p.numel() + out_j.numel()) * 4},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline,
                "clips_per_s": world * B / (ms * 1e-3), "output_poses_per_s": world * B * T_out / (ms * 1e-3)}
        if world == 1 and not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            fps, sec, threads = time_cpu(CPU_SAMPLE_CLIPS, 3)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{CPU_SAMPLE_CLIPS} of the {B} clips (T={T}), 3 iterations, oracle port of the "
                                              f"reference PyTorch modules + numpy FK; {sec:.2f} s/iteration"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying a CUDA graph")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
