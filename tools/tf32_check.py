"""Accuracy / speed of the opt-in 3xTF32 implicit GEMM (csrc/rowgemm_tf32.cu) against the SIMT fp32 kernel, one GEMM
at a time: max relative error vs an fp64 matmul and CUDA-event time.   python tools/tf32_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_inverse_kinematics_b200 import ops  # noqa: E402


def run(rows, k, n, mode):
    os.environ.pop("TIK_NO_TF32", None)
    if mode == "simt":
        os.environ["TIK_NO_TF32"] = "1"
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(rows, 1, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.zeros(1, n, device="cuda")
    out = ops.rowgemm([(a, 1, 0)], w, b, rows, 1, 1)
    ref = a.double().reshape(rows, k) @ w.double().t()
    err = ((out.reshape(rows, n).double() - ref).abs().max() / ref.abs().max()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.rowgemm([(a, 1, 0)], w, b, rows, 1, 1)
    e0.record()
    for _ in range(10):
        ops.rowgemm([(a, 1, 0)], w, b, rows, 1, 1)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    return err, us, 2.0 * rows * k * n / us / 1e6


if __name__ == "__main__":
    if "--once" in sys.argv:      # one launch per shape on the tensor-core kernel: the input of an ncu capture
        os.environ.pop("TIK_NO_TF32", None)
        for rows, k, n in ((65536, 64, 64), (65536, 256, 128), (65536, 1024, 256), (16384, 4352, 512)):
            a = torch.randn(rows, 1, k, device="cuda")
            w = torch.randn(n, k, device="cuda") / k ** 0.5
            ops.rowgemm([(a, 1, 0)], w, torch.zeros(1, n, device="cuda"), rows, 1, 1)
        torch.cuda.synchronize()
        sys.exit(0)
    for rows, k, n in ((65536, 64, 64), (65536, 256, 128), (65536, 1024, 256), (16384, 4352, 512)):
        for mode in ("simt", "tf32x3"):
            err, us, tf = run(rows, k, n, mode)
            print(f"rows={rows} K={k} N={n} {mode:7s}: max rel err {err:.2e}  {us:.1f} us  {tf:.1f} TFLOP/s")
