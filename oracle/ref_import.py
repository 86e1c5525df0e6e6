"""Import the REAL reference modules from /root/reference (dev container only).

TEST INFRASTRUCTURE.  Used by ``make_golden.py`` and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent,
e.g. on the GPU box).  Follows the recipe in SURVEY.md Appendix A: the package
``__init__`` files of the reference drag in mmcv / lazy_import / pycocotools,
which are not installed, so we pre-register empty namespace modules whose
``__path__`` points at the real directories and stub the few third-party
packages that the hot-path files import at module scope but never call on the
inference path (pytorch_lightning, smplx, matplotlib).
"""
import argparse
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


_snapshot = False


def _find_root():
    """The reference checkout (dev container), else the archive oracle/build_ref.py made of its hot-path files
    (``oracle/_ref/hot_path.tar.gz``, which travels to the GPU box), unpacked into a temporary directory outside the
    repository for the life of the process."""
    global _snapshot
    for cand in (os.environ.get("TIK_REFERENCE_ROOT"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "mmskeleton")):
            return cand
    archive = os.path.join(_HERE, "_ref", "hot_path.tar.gz")
    if os.path.exists(archive):
        import atexit
        import shutil
        import tempfile
        from . import build_ref
        dst = tempfile.mkdtemp(prefix="tik_ref_")
        atexit.register(shutil.rmtree, dst, ignore_errors=True)
        build_ref.extract(dst)
        _snapshot = True
        return dst
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mmskeleton"))


def is_snapshot() -> bool:
    return _snapshot


def _namespace(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    import torch.nn as nn

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name, rel in [("mmskeleton", "mmskeleton"),
                      ("mmskeleton.ops", "mmskeleton/ops"),
                      ("mmskeleton.models", "mmskeleton/models"),
                      ("mmskeleton.models.backbones", "mmskeleton/models/backbones"),
                      ("mmskeleton.datasets", "mmskeleton/datasets")]:
        _namespace(name, os.path.join(REF_ROOT, rel))

    class _LightningModule(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

    if "pytorch_lightning" not in sys.modules:
        cb = _stub("pytorch_lightning.callbacks", ModelCheckpoint=object)
        core = _stub("pytorch_lightning.core", LightningModule=_LightningModule)
        _stub("pytorch_lightning", _logger=None, Trainer=object, callbacks=cb, core=core,
              LightningModule=_LightningModule)
    if "smplx" not in sys.modules:
        jn = _stub("smplx.joint_names", JOINT_NAMES=[])
        _stub("smplx", create=None, joint_names=jn)
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot")
        _stub("matplotlib", pyplot=plt)
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=object)

    st = importlib.import_module("mmskeleton.models.backbones.st_gcn_aaai18")
    models = sys.modules["mmskeleton.models"]
    models.StgGcn18, models.StgLayerConfig, models.StgConfig = st.StgGcn18, st.StgLayerConfig, st.StgConfig
    gconv = importlib.import_module("mmskeleton.ops.st_gcn.gconv_origin")
    graph = importlib.import_module("mmskeleton.ops.st_gcn.graph")
    data_amass = importlib.import_module("mmskeleton.datasets.data_amass")
    sys.modules["mmskeleton.datasets"].AmassDataset = data_amass.AmassDataset
    pose_trainer = importlib.import_module("pose_trainer")
    geometry = importlib.import_module("common.geometry")
    kornia = importlib.import_module("common.kornia_geometry_conversion")
    kps_util = importlib.import_module("common.keypoints_util")

    _loaded = types.SimpleNamespace(st_gcn=st, gconv=gconv, graph=graph, data_amass=data_amass,
                                    pose_trainer=pose_trainer, geometry=geometry, kornia=kornia,
                                    keypoints_util=kps_util)
    return _loaded


def default_hparams(**over):
    hp = dict(graph_layout="coco", max_hop=2, dilation=1, kps_channel=3, win_size=9, lr=1e-4)
    hp.update(over)
    return argparse.Namespace(**hp)
