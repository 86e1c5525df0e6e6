"""CPU restatement of the reference ST-GCN IK forward (TEST INFRASTRUCTURE).

Functional, state-dict driven, fp32, torch CPU ops in the reference's own order
of operations.  Nothing here is used by the product path.

Follows (all paths relative to /root/reference):
  * skeleton graph ............. mmskeleton/ops/st_gcn/graph.py:27-133,136-159
  * graph convolution .......... mmskeleton/ops/st_gcn/gconv_origin.py:56-65
  * ST-GCN block ............... mmskeleton/models/backbones/st_gcn_aaai18.py:161-214
  * backbone ................... mmskeleton/models/backbones/st_gcn_aaai18.py:113-133
  * pose regressor head ........ pose_trainer.py:89-106,128-133
"""
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F

from .synth import POSE_REGRESSOR_LAYERS

BN_EPS = 1e-5  # torch.nn.BatchNorm default, used by every BN in the reference


# --------------------------------------------------------------------------- graph
def skeleton_edges(layout):
    """(num_node, neighbour links, centre) -- graph.py:43-88."""
    if layout == "openpose":
        links = [(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2),
                 (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14)]
        return 18, links, 1
    if layout == "ntu-rgb+d":
        one = [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10),
               (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 23),
               (23, 8), (24, 25), (25, 12)]
        return 25, [(a - 1, b - 1) for a, b in one], 20
    if layout == "ntu_edge":
        one = [(1, 2), (3, 2), (4, 3), (5, 2), (6, 5), (7, 6), (8, 7), (9, 2), (10, 9), (11, 10), (12, 11),
               (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (21, 22), (22, 8),
               (23, 24), (24, 12)]
        return 24, [(a - 1, b - 1) for a, b in one], 2
    if layout == "coco":
        one = [(16, 14), (14, 12), (17, 15), (15, 13), (12, 13), (6, 12), (7, 13), (6, 7), (8, 6), (9, 7),
               (10, 8), (11, 9), (2, 3), (2, 1), (3, 1), (4, 2), (5, 3), (4, 6), (5, 7)]
        return 17, [(a - 1, b - 1) for a, b in one], 0
    raise ValueError("Do Not Exist This Layout.")


def hop_distance(num_node, links, max_hop):
    """Shortest-path hop counts, inf beyond max_hop -- graph.py:136-148 (matrix powers there,
    breadth-first search here; identical integers)."""
    nbr = [[] for _ in range(num_node)]
    for a, b in links:
        nbr[a].append(b)
        nbr[b].append(a)
    dist = np.full((num_node, num_node), np.inf)
    for s in range(num_node):
        dist[s, s] = 0
        q = deque([s])
        while q:
            u = q.popleft()
            if dist[s, u] >= max_hop:
                continue
            for w in nbr[u]:
                if dist[s, w] == np.inf:
                    dist[s, w] = dist[s, u] + 1
                    q.append(w)
    return dist


def build_adjacency(layout="openpose", strategy="uniform", max_hop=1, dilation=1):
    """Normalised adjacency stack A (K,V,V) float64 -- graph.py:91-133,151-159."""
    V, links, centre = skeleton_edges(layout)
    hop = hop_distance(V, links, max_hop)
    hops = list(range(0, max_hop + 1, dilation))
    binary = np.zeros((V, V))
    for h in hops:
        binary[hop == h] = 1
    colsum = binary.sum(0)
    inv = np.zeros(V)
    inv[colsum > 0] = colsum[colsum > 0] ** (-1)
    norm = binary * inv[None, :]          # A . D^-1 (column normalisation)
    if strategy == "uniform":
        return norm[None].copy()
    if strategy == "distance":
        A = np.zeros((len(hops), V, V))
        for i, h in enumerate(hops):
            A[i][hop == h] = norm[hop == h]
        return A
    if strategy == "spatial":
        planes = []
        for h in hops:
            root, close, further = np.zeros((V, V)), np.zeros((V, V)), np.zeros((V, V))
            for i in range(V):
                for j in range(V):
                    if hop[j, i] == h:
                        if hop[j, centre] == hop[i, centre]:
                            root[j, i] = norm[j, i]
                        elif hop[j, centre] > hop[i, centre]:
                            close[j, i] = norm[j, i]
                        else:
                            further[j, i] = norm[j, i]
            if h == 0:
                planes.append(root)
            else:
                planes.append(root + close)
                planes.append(further)
        return np.stack(planes)
    raise ValueError("Do Not Exist This Strategy")


# --------------------------------------------------------------------------- layers
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def graph_conv(x, A, weight, bias):
    """ConvTemporalGraphical.forward with t_kernel_size=1 -- gconv_origin.py:56-65.
    x (N,Cin,T,V), A (K,V,V), weight (K*Cout,Cin,1,1) -> (N,Cout,T,V)."""
    K = A.shape[0]
    y = F.conv2d(x, weight, bias)
    n, kc, t, v = y.shape
    y = y.view(n, K, kc // K, t, v)
    return torch.einsum("nkctv,kvw->nctw", y, A).contiguous()


def st_gcn_block(x, A, sd, p, cin, cout, stride, residual=True):
    """StGcnBlock.forward in eval mode -- st_gcn_aaai18.py:191-214."""
    if not residual:
        res = 0
    elif cin == cout and stride == 1:
        res = x
    else:
        res = _bn(F.conv2d(x, sd[p + "residual.0.weight"], sd[p + "residual.0.bias"], stride=(stride, 1)),
                  sd, p + "residual.1")
    y = graph_conv(x, A, sd[p + "gcn.conv.weight"], sd[p + "gcn.conv.bias"])
    y = F.relu(_bn(y, sd, p + "tcn.0"))
    wt = sd[p + "tcn.2.weight"]
    y = F.conv2d(y, wt, sd[p + "tcn.2.bias"], stride=(stride, 1), padding=((wt.shape[2] - 1) // 2, 0))
    y = _bn(y, sd, p + "tcn.3")           # Dropout(p=0) in eval is the identity
    return F.relu(y + res)


def backbone_forward(sd, x, layers=POSE_REGRESSOR_LAYERS, prefix="backbone.", collect=None):
    """StgGcn18.forward -- st_gcn_aaai18.py:113-133.  x (N,T,V,C) -> (N,T',V*C_last)."""
    N, T, V, C = x.shape
    A = sd[prefix + "A"]
    h = x.permute(0, 2, 3, 1).contiguous().view(N, V * C, T)
    if prefix + "data_bn.weight" in sd:
        h = _bn(h, sd, prefix + "data_bn")
    h = h.view(N, V, C, T).permute(0, 2, 3, 1).contiguous()      # (N,C,T,V)
    for i, (cin, cout, stride, residual) in enumerate(layers):
        imp = sd.get(f"{prefix}edge_importance.{i}", 1)
        h = st_gcn_block(h, A * imp, sd, f"{prefix}st_gcn_networks.{i}.", cin, cout, stride, residual)
        if collect is not None:
            collect.append(h)
    h = h.permute(0, 2, 3, 1).contiguous()
    return h.view(h.shape[0], h.shape[1], -1)


def regressor_forward(sd, x, layers=POSE_REGRESSOR_LAYERS):
    """PoseRegressor.forward (live 66-d axis-angle head) -- pose_trainer.py:94-133."""
    with torch.no_grad():
        f = backbone_forward(sd, x, layers)
        n, t, c = f.shape
        h = F.linear(f.view(n * t, c), sd["pose_regressor.0.weight"], sd["pose_regressor.0.bias"])
        h = F.leaky_relu(h, 0.01)          # Dropout(0.7) is the identity in eval
        h = F.linear(h, sd["pose_regressor.3.weight"], sd["pose_regressor.3.bias"])
        return {"poses": h.view(n, t, -1)}


def iterative_regressor_forward(sd, x, n_iter=3, layers=POSE_REGRESSOR_LAYERS):
    """The HMR-style iterative 6-D head that pose_trainer.py keeps commented out (:53-64 definition, :108-126 forward):
    pred = init_pose; repeat n_iter times: xc = fc2(fc1(cat[x, pred])) (dropouts are identities in eval, there is no
    activation between the layers in the reference text), pred = decpose(xc) + pred; rot6d -> rotmat -> axis-angle.
    **parity unpinned**: the code is commented out in the reference, so there is nothing to run for goldens."""
    from . import geometry_port as gp
    with torch.no_grad():
        f = backbone_forward(sd, x, layers)
        n, t, c = f.shape
        feats = f.reshape(n * t, c)
        pred = sd["init_pose"].expand(n * t, -1)
        for _ in range(n_iter):
            xc = torch.cat([feats, pred], 1)
            xc = F.linear(xc, sd["fc1.weight"], sd["fc1.bias"])
            xc = F.linear(xc, sd["fc2.weight"], sd["fc2.bias"])
            pred = F.linear(xc, sd["decpose.weight"], sd["decpose.bias"]) + pred
        rot = torch.from_numpy(gp.rot6d_to_rotmat(pred.numpy().reshape(-1, 6))).view(n * t, 22, 3, 3)
        aa = torch.from_numpy(gp.rotation_matrix_to_angle_axis(rot.numpy().reshape(-1, 3, 3))).view(n, t, 66)
        return {"poses": aa, "rotmats": rot, "rot6d": pred}


# --------------------------------------------------------------------------- callers
def sample_window(arr, idx, half):
    """Edge-padded window of 2*half+1 frames centred on idx -- data_amass.py:18-42."""
    n = arr.shape[0]
    if n < 2 * half + 1:
        # the reference pads only one side (elif chain) and returns ragged windows / raises here
        raise ValueError(f"sequence of {n} frames is shorter than the {2 * half + 1}-frame window")
    ids = np.clip(np.arange(idx - half, idx + half + 1), 0, n - 1)
    return arr[ids]


def inference_windows(seq, win_size, relative=True):
    """InferenceDataset over a whole sequence -- data_amass.py:221-236. (F,17,3)->(F,2h+1,17,3)."""
    half = win_size // 2
    out = np.stack([sample_window(seq, i, half) for i in range(seq.shape[0])])
    if relative:
        root = 0.5 * (out[:, :, 11, :] + out[:, :, 12, :])
        out = out - root[:, :, None, :]
    return out


def moveai_to_coco(joints_3d, names):
    """inference.py:121-133: moveai 22-joint -> COCO-17 remap, nose/eyes from ears, y<-z, z<- -y."""
    want = ["L_Ear", "R_Ear", "L_Shoulder", "R_Shoulder", "L_Elbow", "R_Elbow", "L_Wrist", "R_Wrist",
            "L_Hip", "R_Hip", "L_Knee", "R_Knee", "L_Ankle", "R_Ankle"]
    out = np.zeros((joints_3d.shape[0], 17, joints_3d.shape[2]), dtype=np.float32)
    for k, nm in enumerate(want):
        out[:, 3 + k] = joints_3d[:, names.index(nm)]
    out[:, 0] = 0.5 * (joints_3d[:, -1] + joints_3d[:, -2])
    out[:, 1] = joints_3d[:, -2]
    out[:, 2] = joints_3d[:, -1]
    y = out[:, :, 1].copy()
    z = out[:, :, 2].copy()
    out[:, :, 1] = z
    out[:, :, 2] = -y
    return out
