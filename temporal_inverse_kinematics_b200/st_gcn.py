"""ST-GCN modules with the reference's constructor arguments, parameter names and state_dict layout,
whose forward passes run on libtik.so (hand-written sm_100a kernels).

Mirrors
  * ConvTemporalGraphical ... reference mmskeleton/ops/st_gcn/gconv_origin.py:7-65
  * StGcnBlock .............. reference mmskeleton/models/backbones/st_gcn_aaai18.py:136-214
  * StgLayerConfig/StgConfig  reference mmskeleton/models/backbones/st_gcn_aaai18.py:17-29
  * StgGcn18 ................ reference mmskeleton/models/backbones/st_gcn_aaai18.py:32-133

``load_state_dict(reference_model.state_dict(), strict=True)`` works in both directions.  The torch
sub-modules below (Conv2d / BatchNorm) only *hold* parameters; they are never called.  Forward supports
eval mode + CUDA tensors only and raises otherwise (no CPU / eager fallback).
"""
import os
from dataclasses import dataclass
from typing import List

import torch
import torch.nn as nn

from . import engine, ops
from .graph import Graph


def _default_dtype():
    return os.environ.get("TIK_COMPUTE_DTYPE", "fp32")


class _ComputeDtypeMixin:
    """compute_dtype: 'fp32' (SIMT, <=1e-4 parity with the reference) or 'bf16' (tcgen05 tensor cores)."""

    def set_compute_dtype(self, name):
        engine.resolve_dtype(name)
        for m in self.modules():
            if isinstance(m, _ComputeDtypeMixin):
                m.compute_dtype = name
        return self


def _identity(x):
    """data_bn=False stand-in (module level so the model pickles; the reference uses a lambda there)."""
    return x


def _pad_channels(t, mult):
    c = t.shape[-1]
    pad = (-c) % mult
    return t if pad == 0 else torch.nn.functional.pad(t, (0, pad))


class ConvTemporalGraphical(nn.Module, _ComputeDtypeMixin):
    """Graph convolution: 1x1 conv to K*Cout channels, then einsum('nkctv,kvw->nctw') with A."""

    def __init__(self, in_channels, out_channels, kernel_size, t_kernel_size=1, t_stride=1, t_padding=0,
                 t_dilation=1, bias=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(t_kernel_size, 1),
                              padding=(t_padding, 0), stride=(t_stride, 1), dilation=(t_dilation, 1), bias=bias)
        self._temporal = (t_kernel_size, t_stride, t_padding, t_dilation)
        self.compute_dtype = _default_dtype()

    def forward(self, x, A):
        assert A.size(0) == self.kernel_size
        engine.require_cuda_eval(self, x, "ConvTemporalGraphical")
        if self._temporal != (1, 1, 0, 1):
            raise NotImplementedError("the CUDA path implements t_kernel_size=1 graph convolutions (all the IK model uses)")
        _, tdt = engine.resolve_dtype(self.compute_dtype)
        mult = 4 if tdt == torch.float32 else 64
        N, Cin, T, V = x.shape
        K, cout = self.kernel_size, self.out_channels
        w, b = engine.fold_gcn(self.conv.weight, self.conv.bias, A, K, None)
        cpad = (-Cin) % mult
        if cpad:
            w = torch.nn.functional.pad(w.view(cout, K, Cin), (0, cpad)).reshape(cout, K * (Cin + cpad))
        opad = (-cout) % mult
        if opad:
            w = torch.nn.functional.pad(w, (0, 0, 0, opad))
            b = torch.nn.functional.pad(b, (0, opad))
        xa = ops.aggregate(_pad_channels(ops.to_node_major(x, tdt), mult), A.detach().float().contiguous())
        slabs = [(xa[k].view(N * V, T, Cin + cpad), 1, 0) for k in range(K)]
        y = ops.rowgemm(slabs, w.to(tdt).contiguous(), b.float().contiguous(), N * V, V, T, act="none")
        y = y.view(N, V, T, cout + opad)[..., :cout]
        return ops.from_node_major(y), A


class StGcnBlock(nn.Module, _ComputeDtypeMixin):
    """residual + (graph conv -> BN -> ReLU -> (kt x 1) temporal conv -> BN -> Dropout) -> ReLU."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        kt, K = kernel_size
        self.in_channels, self.out_channels, self.stride, self.temporal_kernel = in_channels, out_channels, stride, kt
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, K)
        self.tcn = nn.Sequential(
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, (kt, 1), (stride, 1), ((kt - 1) // 2, 0)),
            nn.BatchNorm2d(out_channels),
            nn.Dropout(dropout, inplace=True),
        )
        if not residual:
            self.residual_kind = "none"
        elif in_channels == out_channels and stride == 1:
            self.residual_kind = "identity"
        else:
            self.residual_kind = "conv"
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                                          nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(inplace=True)
        self.compute_dtype = _default_dtype()

    def forward(self, x, A):
        """x (N, Cin, T, V) NCHW, A (K, V, V) -> (y (N, Cout, T', V), A) -- stand-alone block call."""
        engine.require_cuda_eval(self, x, "StGcnBlock")
        _, tdt = engine.resolve_dtype(self.compute_dtype)
        mult = 4 if tdt == torch.float32 else 64
        N, Cin, T, V = x.shape
        K, cout, s, kt = self.gcn.kernel_size, self.out_channels, self.stride, self.temporal_kernel
        if cout % mult:
            raise NotImplementedError(f"out_channels must be a multiple of {mult} for compute dtype {self.compute_dtype}")
        cpad = (-Cin) % mult
        wg, bg = engine.fold_gcn(self.gcn.conv.weight, self.gcn.conv.bias, A, K, self.tcn[0])
        if cpad:
            wg = torch.nn.functional.pad(wg.view(cout, K, Cin), (0, cpad)).reshape(cout, -1)
        conv_res = self.residual_kind == "conv"
        wt, bt, wr = engine.fold_tcn(self.tcn[2], self.tcn[3], self.residual[0] if conv_res else None,
                                     self.residual[1] if conv_res else None)
        xn = _pad_channels(ops.to_node_major(x, tdt), mult)
        xa = ops.aggregate(xn, A.detach().float().contiguous())
        h = ops.rowgemm([(xa[k].view(N * V, T, Cin + cpad), 1, 0) for k in range(K)], wg.to(tdt).contiguous(),
                        bg.float().contiguous(), N * V, V, T, act="relu")
        T_out = (T - 1) // s + 1
        pad = (kt - 1) // 2
        slabs = [(h, s, dt - pad) for dt in range(kt)]
        residual = None
        if conv_res:
            slabs.append((xn.view(N * V, T, Cin + cpad), s, 0))
            wt = torch.cat([wt, torch.nn.functional.pad(wr, (0, cpad))], dim=1)
        elif self.residual_kind == "identity":
            residual = xn.view(N * V, T, Cin)
        y = ops.rowgemm(slabs, wt.to(tdt).contiguous(), bt.float().contiguous()[None], N * V, V, T_out, act="relu",
                        residual=residual)
        return ops.from_node_major(y.view(N, V, T_out, cout)), A


@dataclass
class StgLayerConfig:
    in_channels: int
    out_channels: int
    temporal_stride: int
    is_residual: True


@dataclass
class StgConfig:
    layers: List[StgLayerConfig]
    temporal_kernel_size: int


class StgGcn18(engine.EngineHolder, nn.Module, _ComputeDtypeMixin):
    """ST-GCN backbone: (N, T, V, C) keypoint windows -> (N, T', V * C_last) features."""

    def __init__(self, config: StgConfig, graph_cfg, edge_importance_weighting=True, data_bn=True, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_cfg)
        self.register_buffer("A", torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False))
        self.n_in_keypoints = self.A.size(1)
        kernel_size = (config.temporal_kernel_size, self.A.size(0))
        c0 = config.layers[0].in_channels
        self.data_bn = nn.BatchNorm1d(c0 * self.A.size(1)) if data_bn else _identity
        block_kw = {k: v for k, v in kwargs.items() if k != "dropout"}
        self.st_gcn_networks = nn.ModuleList([
            StGcnBlock(l.in_channels, l.out_channels, kernel_size, stride=l.temporal_stride, residual=l.is_residual,
                       **block_kw) for l in config.layers])
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList(
                [nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks])
        else:
            self.edge_importance = [1] * len(self.st_gcn_networks)
        self.compute_dtype = _default_dtype()
        self.chunk_clips = None            # None = engine.default_chunk(T)
        self.weight_check = None           # None = engine default ('version'); 'content' also checksums the tensors
        self._engine = None

    def out_frames(self, T):
        for b in self.st_gcn_networks:
            T = (T - 1) // b.stride + 1
        return T

    def _check_input(self, x):
        if x.dim() != 4 or x.shape[2] != self.A.size(1) or x.shape[3] != self.st_gcn_networks[0].in_channels:
            raise ValueError(f"expected (N, T, {self.A.size(1)}, {self.st_gcn_networks[0].in_channels}) input, "
                             f"got {tuple(x.shape)}")
        return x.detach().float().contiguous()

    def forward(self, x):
        engine.require_cuda_eval(self, x, "StgGcn18")
        x = self._check_input(x)
        if self._engine is None:
            self._engine = engine.Engine(self)
        self._engine.weight_check = self.weight_check or self._engine.weight_check
        N, T = x.shape[0], x.shape[1]
        if N == 0:
            return x.new_zeros((0, self.out_frames(T), self.A.size(1) * self.st_gcn_networks[-1].out_channels))
        plan = self._engine.plan(self.compute_dtype, N, T, self.chunk_clips)
        _, feat = plan.run(x, want_feat=True)
        return feat.float()
