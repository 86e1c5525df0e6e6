"""Thin Python wrappers over the libtik.so primitives (include/tik.h).  CUDA tensors only.

Activation tensors are *node-major* (N, V, T, C) contiguous; ``to_node_major`` / ``from_node_major`` convert
from / to the reference's NCHW (N, C, T, V) module-boundary layout (plumbing, done with torch).
"""
import ctypes as C

import torch

from . import _lib as L


def _code(t):
    if t.dtype == torch.float32:
        return L.TIK_F32
    if t.dtype == torch.bfloat16:
        return L.TIK_BF16
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def _dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libtik ops need CUDA tensors -- there is no CPU fallback")
        if t is not None and not t.is_contiguous():
            raise ValueError("libtik ops need contiguous tensors")


def to_node_major(x_nctv, dtype):
    """(N, C, T, V) -> (N, V, T, C) contiguous in `dtype`."""
    return x_nctv.permute(0, 3, 2, 1).to(dtype).contiguous()


def from_node_major(x_nvtc):
    """(N, V, T, C) -> (N, C, T, V) contiguous fp32 (what the reference modules return)."""
    return x_nvtc.permute(0, 3, 2, 1).float().contiguous()


def aggregate(x, A_hat):
    """x (N,V,T,C), A_hat (K,V,V) fp32 -> (K,N,V,T,C):  out[k,n,w] = sum_v A_hat[k,v,w] x[n,v]."""
    _dev(x, A_hat)
    N, V, T, Cc = x.shape
    K = A_hat.shape[0]
    out = torch.empty((K, N, V, T, Cc), dtype=x.dtype, device=x.device)
    with L.on_device(x):
        L.check(L.lib().tik_aggregate(_code(x), L.ptr(x), L.ptr(A_hat), L.ptr(out), N, T, V, Cc, K, L.stream_ptr(x.device)))
    return out


def stem_gcn(x, in_scale, in_shift, A_hat, w, bias, out_dtype, relu=True, res_w=None, res_stride=1):
    """x (N,T,V,Cin) fp32 -> (N,V,T,Cout) out_dtype (data_bn + aggregate + channel mix + bias [+ReLU]);
    with res_w (V,Cout,Cin) also returns the residual branch (N,V,T',Cout)."""
    _dev(x, in_scale, in_shift, A_hat, w, bias, res_w)
    N, T, V, Cin = x.shape
    K = A_hat.shape[0]
    Cout = w.shape[0]
    out = torch.empty((N, V, T, Cout), dtype=out_dtype, device=x.device)
    res = None if res_w is None else torch.empty((N, V, (T - 1) // res_stride + 1, Cout), dtype=out_dtype, device=x.device)
    with L.on_device(x):
        L.check(L.lib().tik_stem_gcn(_code(out), L.ptr(x), L.ptr(in_scale), L.ptr(in_shift), L.ptr(A_hat), L.ptr(w), L.ptr(bias),
                                     L.ptr(out), L.ptr(res_w), L.ptr(res), res_stride, N, T, V, Cin, K, Cout, int(relu),
                                     L.stream_ptr(x.device)))
    return out if res_w is None else (out, res)


def build_abd(A_hat, T, cin=None):
    """(128,128) bf16 block-structured aggregation operand of the fused graph conv: rows (w, t), columns (v, t'),
    entry A_hat[v, w] * (t == t'), f = ceil(T / ceil(T/7)) frames per tile (tik.h; at most 5 frames for the 256-channel
    kernel: f = ceil(T / ceil(T/5))), zero padding."""
    V = A_hat.shape[-1]
    fmax = 5 if cin == 256 else 7
    tiles = (int(T) + fmax - 1) // fmax
    f = (int(T) + tiles - 1) // tiles
    a = A_hat.reshape(V, V).float()
    abd = torch.zeros(128, 128, dtype=torch.float32, device=A_hat.device)
    abd[: V * f, : V * f] = torch.kron(a.t().contiguous(), torch.eye(f, device=A_hat.device))
    return abd.to(torch.bfloat16).contiguous()


def gcn_fused(x, abd, w, bias, relu=True):
    """x (N,V,T,Cin) bf16 -> (N,V,T,Cout) bf16: relu(W . aggregate(x) + bias[node]) in one tensor-core kernel."""
    _dev(x, abd, w, bias)
    N, V, T, Cin = x.shape
    Cout = w.shape[0]
    out = torch.empty((N, V, T, Cout), dtype=torch.bfloat16, device=x.device)
    with L.on_device(x):
        L.check(L.lib().tik_gcn_fused(L.ptr(x), L.ptr(abd), L.ptr(w), L.ptr(bias), L.ptr(out), N, T, V, Cin, Cout, int(relu),
                                      L.stream_ptr(x.device)))
    return out


def rowgemm(slabs, w, bias, nv, v, t_out, act="none", slope=0.01, residual=None, out_layout="node", c_out_valid=None,
            stem_residual=None, out=None):
    """Generic implicit GEMM (TikRowGemm).

    slabs: list of (tensor (NV, t_in, c) node-major, t_mul, t_off); w (c_out, sum c); bias (1|V, c_out) fp32.
    residual: tensor (NV, t_out, c_out) of the activation dtype (identity residual).
    stem_residual: (x (N,T,V,cin) fp32, res_w (V,c_out,cin) fp32, t_mul).
    out_layout: 'node' (NV,t_out,c_out) | 'time' (N,t_out,V,c_out) | 'rows_f32' (NV*t_out, c_out_valid) fp32.
    out: optional preallocated result of that shape / dtype (stable addresses let libtik re-use the prepared launch
    and make the call CUDA-graph capturable without allocator traffic).
    """
    a0 = slabs[0][0]
    _dev(w, bias, residual, *[s[0] for s in slabs])
    code = _code(a0)
    g = L.TikRowGemm()
    g.n_slabs = len(slabs)
    for i, (a, t_mul, t_off) in enumerate(slabs):
        if a.dtype != a0.dtype:
            raise TypeError("all slabs must share one dtype")
        g.slabs[i].a_dev = a.data_ptr()
        g.slabs[i].c = a.shape[-1]
        g.slabs[i].t_in = a.shape[-2]
        g.slabs[i].t_mul, g.slabs[i].t_off = t_mul, t_off
    c_out = w.shape[0]
    c_valid = c_out if c_out_valid is None else c_out_valid
    if w.dtype != a0.dtype or w.shape[1] != sum(s[0].shape[-1] for s in slabs):
        raise ValueError("weight dtype/shape does not match the slabs")
    g.w_dev, g.bias_dev = w.data_ptr(), bias.data_ptr()
    g.bias_per_node = int(bias.dim() == 2 and bias.shape[0] > 1)
    g.nv, g.v, g.t_out, g.c_out, g.c_out_valid = nv, v, t_out, c_out, c_valid
    g.act = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "leaky": L.ACT_LEAKY}[act]
    g.slope = slope
    keep = []
    if residual is not None:
        g.res_kind, g.res_dev = L.RES_IDENTITY, residual.data_ptr()
    elif stem_residual is not None:
        xr, rw, t_mul = stem_residual
        _dev(xr, rw)
        g.res_kind, g.res_dev, g.res_w_dev = L.RES_STEM, xr.data_ptr(), rw.data_ptr()
        g.res_cin, g.res_t_mul, g.res_t_in = xr.shape[-1], t_mul, xr.shape[1]
        keep += [xr, rw]
    else:
        g.res_kind = L.RES_NONE
    if out_layout == "node":
        shape, odt, g.out_layout = (nv, t_out, c_out), a0.dtype, L.OUT_NODE_MAJOR
    elif out_layout == "time":
        shape, odt, g.out_layout = (nv // v, t_out, v, c_out), a0.dtype, L.OUT_TIME_MAJOR
    elif out_layout == "rows_f32":
        shape, odt, g.out_layout = (nv * t_out, c_valid), torch.float32, L.OUT_ROWS_F32
    else:
        raise ValueError(out_layout)
    if out is None:
        out = torch.empty(shape, dtype=odt, device=a0.device)
    elif tuple(out.shape) != shape or out.dtype != odt or not out.is_contiguous() or out.device != a0.device:
        raise ValueError(f"out must be a contiguous {odt} tensor of shape {shape} on {a0.device}")
    g.out_dev = out.data_ptr()
    with L.on_device(a0):
        L.check(L.lib().tik_rowgemm(code, C.byref(g), L.stream_ptr(a0.device)))
    return out
