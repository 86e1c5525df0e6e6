"""Generate tests/golden/*.npz by running the REAL reference (dev container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  Imports the reference's own modules from /root/reference
(oracle/ref_import.py), feeds them the deterministic synthetic weights/inputs of
oracle/synth.py and stores inputs + reference outputs as small fixtures.  The
fixtures (not the reference) travel to the GPU box.
"""
import os
import sys

import numpy as np
import torch

from . import ref_import, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_regressor(ref, sd, **hp):
    m = ref.pose_trainer.PoseRegressor(ref_import.default_hparams(**hp)).eval()
    m.load_state_dict(sd, strict=True)
    return m


def golden_graph(ref):
    out = {}
    for layout in ["coco", "openpose", "ntu-rgb+d", "ntu_edge"]:
        for strategy in ["uniform", "distance", "spatial"]:
            for max_hop, dilation in [(1, 1), (2, 1), (2, 2), (3, 1)]:
                g = ref.graph.Graph(layout=layout, strategy=strategy, max_hop=max_hop, dilation=dilation)
                out[f"{layout}|{strategy}|{max_hop}|{dilation}"] = g.A
    np.savez_compressed(os.path.join(OUT, "graph.npz"), **out)


def golden_stgcn(ref):
    A = ref.graph.Graph(layout="coco", strategy="uniform", max_hop=2, dilation=1).A
    sd = synth.make_regressor_state(A, seed=0)
    m = _ref_regressor(ref, sd)
    out = {"state_checksum": np.float64(synth.state_checksum(sd))}
    for tag, (n, t, seed) in {"t9": (3, 9, 1234), "t64": (2, 64, 1235), "t13": (2, 13, 1236), "t1": (2, 1, 1237),
                              "t128": (1, 128, 1238)}.items():
        x = synth.make_clips(n, t, seed=seed)
        with torch.no_grad():
            feats = []
            h = x
            # per-block outputs through the reference's own modules (NCHW), stored as checksums
            bb = m.backbone
            N, T, V, C = x.shape
            h = x.permute(0, 2, 3, 1).contiguous().view(N, V * C, T)
            h = bb.data_bn(h).view(N, V, C, T).permute(0, 2, 3, 1).contiguous().view(N, C, T, V)
            for gcn, imp in zip(bb.st_gcn_networks, bb.edge_importance):
                h, _ = gcn(h, bb.A * imp)
                feats.append([float(h.double().mean()), float(h.double().abs().mean()), float(h.double().abs().max())])
            y = m(x)["poses"]
            f = m.backbone(x)
        out[f"{tag}_shape"] = np.array([n, t, seed])
        out[f"{tag}_poses"] = y.numpy()
        out[f"{tag}_block_stats"] = np.array(feats)
        out[f"{tag}_feat_head"] = f.numpy().reshape(-1)[:2048].copy()
    np.savez_compressed(os.path.join(OUT, "stgcn.npz"), **out)

    # multi-partition graphs (K=3 distance, K=3 spatial@hop1) on a small backbone
    out = {}
    layers = [(3, 64, 1, True), (64, 64, 2, True), (64, 128, 1, True)]
    for strategy, max_hop in [("distance", 2), ("spatial", 1), ("spatial", 2)]:
        graph_cfg = dict(layout="coco", strategy=strategy, max_hop=max_hop, dilation=1)
        A = ref.graph.Graph(**graph_cfg).A
        sd = synth.make_backbone_state(A, layers, kt=3, seed=5, prefix="")
        cfg = ref.st_gcn.StgConfig(layers=[ref.st_gcn.StgLayerConfig(*l) for l in layers], temporal_kernel_size=3)
        bb = ref.st_gcn.StgGcn18(config=cfg, graph_cfg=graph_cfg).eval()
        bb.load_state_dict(sd, strict=True)
        x = synth.make_clips(2, 10, seed=77)
        with torch.no_grad():
            out[f"{strategy}{max_hop}_feat"] = bb(x).numpy()
        out[f"{strategy}{max_hop}_checksum"] = np.float64(synth.state_checksum(sd))
    # temporal kernel 5, no residual on one layer, no data_bn
    layers5 = [(3, 64, 1, False), (64, 64, 1, True), (64, 128, 2, True)]
    graph_cfg = dict(layout="coco", strategy="uniform", max_hop=2, dilation=1)
    A = ref.graph.Graph(**graph_cfg).A
    sd = synth.make_backbone_state(A, layers5, kt=5, seed=6, prefix="")
    cfg = ref.st_gcn.StgConfig(layers=[ref.st_gcn.StgLayerConfig(*l) for l in layers5], temporal_kernel_size=5)
    bb = ref.st_gcn.StgGcn18(config=cfg, graph_cfg=graph_cfg).eval()
    bb.load_state_dict(sd, strict=True)
    x = synth.make_clips(2, 11, seed=78)
    with torch.no_grad():
        out["kt5_feat"] = bb(x).numpy()
    out["kt5_checksum"] = np.float64(synth.state_checksum(sd))

    # single ConvTemporalGraphical call (gconv_origin.py:56-65), K=1 and K=3
    rs = np.random.RandomState(3)
    for K in (1, 3):
        g = ref.gconv.ConvTemporalGraphical(8, 16, K).eval()
        w = torch.from_numpy(rs.standard_normal((16 * K, 8, 1, 1)).astype(np.float32))
        b = torch.from_numpy(rs.standard_normal(16 * K).astype(np.float32))
        g.load_state_dict({"conv.weight": w, "conv.bias": b})
        xx = torch.from_numpy(rs.standard_normal((2, 8, 5, 17)).astype(np.float32))
        AA = torch.from_numpy(rs.uniform(0, 1, (K, 17, 17)).astype(np.float32))
        with torch.no_grad():
            yy, _ = g(xx, AA)
        out[f"gconv{K}_w"], out[f"gconv{K}_b"] = w.numpy(), b.numpy()
        out[f"gconv{K}_x"], out[f"gconv{K}_A"], out[f"gconv{K}_y"] = xx.numpy(), AA.numpy(), yy.numpy()
    np.savez_compressed(os.path.join(OUT, "stgcn_variants.npz"), **out)


def golden_geometry(ref):
    rs = np.random.RandomState(21)
    g, k = ref.geometry, ref.kornia
    out = {}
    x6 = rs.standard_normal((64, 6)).astype(np.float32)
    x6[0] = 0                       # degenerate: both vectors zero
    x6[1, 1::2] = x6[1, 0::2] * 2   # a2 parallel to a1
    x6[2] *= 1e-7
    out["rot6d_in"] = x6
    out["rot6d_out"] = g.rot6d_to_rotmat(torch.from_numpy(x6)).numpy()
    out["rot6d_spin_out"] = g.rot6d_to_rotmat_spin(torch.from_numpy(x6[3:].copy())).numpy()

    aa = (rs.standard_normal((96, 3)) * 1.2).astype(np.float32)
    aa[0] = 0
    aa[1] = [1e-4, -2e-4, 3e-4]
    aa[2] = [1e-3, 0, 0]
    aa[3] = [np.pi, 0, 0]
    aa[4] = [0, np.pi - 1e-3, 0]
    aa[5] = [2.0, -2.0, 1.0]
    aa[6] = [5e-4, 5e-4, 5e-4]
    out["aa_in"] = aa
    t = torch.from_numpy(aa)
    out["aa_kornia_R"] = k.angle_axis_to_rotation_matrix(t).numpy()
    out["aa_rodrigues_R9"] = g.batch_rodrigues(t).numpy()

    R = g.batch_rodrigues(t).view(-1, 3, 3)
    extra = torch.tensor([[[1., 0, 0], [0, 1, 0], [0, 0, 1]],
                          [[-1., 0, 0], [0, -1, 0], [0, 0, 1]],
                          [[1., 0, 0], [0, -1, 0], [0, 0, -1]],
                          [[-1., 0, 0], [0, 1, 0], [0, 0, -1]],
                          [[0., -1, 0], [1, 0, 0], [0, 0, 1]]])
    R = torch.cat([R, extra], 0).contiguous()
    out["R_in"] = R.numpy()
    out["R_to_aa"] = g.rotation_matrix_to_angle_axis(R.clone()).numpy()
    out["R_to_quat_wxyz"] = g.rotation_matrix_to_quaternion(
        torch.cat([R, torch.tensor([0., 0, 1]).view(1, 3, 1).expand(R.shape[0], -1, -1)], -1)).numpy()
    out["R_to_aa_kornia_quirk"] = k.rotation_matrix_to_angle_axis(R.clone()).numpy()
    out["R_to_quat_kornia_xyzw"] = k.rotation_matrix_to_quaternion(R.clone()).numpy()
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **out)


def golden_dance(ref):
    """Config 1 (BASELINE.json configs[0]): data/sample_3d_poses/dance_contemporary.npz through the
    preprocessing of inference.py:121-133 and InferenceDataset(win_size=9)."""
    d = np.load(os.path.join(ref_import.REF_ROOT, "data/sample_3d_poses/dance_contemporary.npz"), allow_pickle=True)
    j3d = d["joints_3d"].astype(np.float32)
    names = d["joint_3d_names"].tolist()
    ku = ref.keypoints_util
    maps = ku.generate_moveai3d_to_coco_mappings(names)
    seq = ku.convert_seq_keypoints(j3d, maps)
    seq[:, 0] = 0.5 * (j3d[:, -1] + j3d[:, -2])
    seq[:, 1] = j3d[:, -2]
    seq[:, 2] = j3d[:, -1]
    y = seq[:, :, 1].copy()
    z = seq[:, :, 2].copy()
    seq[:, :, 1] = z
    seq[:, :, 2] = -y
    ds = ref.data_amass.InferenceDataset(seq, win_size=9, relative_pose=True)
    idxs = [0, 1, 3, 4, 115, 226, 227, 230]
    wins = np.stack([ds[i][0] for i in idxs])
    A = ref.graph.Graph(layout="coco", strategy="uniform", max_hop=2, dilation=1).A
    sd = synth.make_regressor_state(A, seed=0)
    m = _ref_regressor(ref, sd)
    allw = np.stack([ds[i][0] for i in range(len(ds))]).astype(np.float32)
    with torch.no_grad():
        poses = m(torch.from_numpy(allw))["poses"].numpy()      # (231,1,66)
    np.savez_compressed(os.path.join(OUT, "dance.npz"), joints_3d=j3d, joint_3d_names=np.array(names),
                        coco_seq=seq, win_idx=np.array(idxs), windows=wins, poses=poses,
                        state_checksum=np.float64(synth.state_checksum(sd)))


def main():
    if not ref_import.available():
        sys.exit("reference not mounted; goldens can only be regenerated in the dev container")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_import.load()
    golden_graph(ref)
    golden_stgcn(ref)
    golden_geometry(ref)
    golden_dance(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
