"""Recipe for ``oracle/_ref``: a runnable snapshot of the reference's own hot-path modules.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package.

The reference is pure Python (no setup.py, nothing to compile), and ``/root/reference`` does not exist on the
GPU box.  So that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the UNMODIFIED reference
modules there (kind "reference") instead of the oracle restatement (kind "port"), this recipe copies the import
closure of the hot path -- found by importing it with the stubs of ``oracle/ref_import.py`` -- into ONE archive,
``oracle/_ref/hot_path.tar.gz`` (git-ignored, NOT gpurun-ignored: it travels to the box like a built .so, and stays
out of the repository history; no reference source file is ever unpacked inside the repository --
``ref_import`` extracts the archive into a temporary directory at run time).  ``__graft_entry__.build()`` runs it
whenever ``/root/reference`` is present.

    python -m oracle.build_ref            # packs, prints the manifest
"""
import io
import os
import shutil
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("TIK_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

# import closure of pose_trainer.PoseRegressor + conversions + InferenceDataset under the ref_import stubs
HOT_PATH_FILES = [
    "pose_trainer.py",
    "mmskeleton/models/backbones/st_gcn_aaai18.py",
    "mmskeleton/ops/st_gcn/__init__.py",
    "mmskeleton/ops/st_gcn/gconv_origin.py",
    "mmskeleton/ops/st_gcn/graph.py",
    "mmskeleton/datasets/data_amass.py",
    "common/geometry.py",
    "common/kornia_geometry_conversion.py",
    "common/keypoints_util.py",
    "common/pose_def.py",
    "common/draw_util.py",
    "common/smpl_util.py",
]


ARCHIVE = os.path.join(REF_DST, "hot_path.tar.gz")


def build(verbose=False):
    """Packs the files if the reference checkout is present; returns the archive path or None."""
    if not os.path.isdir(os.path.join(REF_SRC, "mmskeleton")):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    if os.path.isdir(REF_DST):                      # drop anything an older recipe left unpacked here
        for name in os.listdir(REF_DST):
            path = os.path.join(REF_DST, name)
            if os.path.isdir(path):
                shutil.rmtree(path)
            elif name != os.path.basename(ARCHIVE):
                os.remove(path)
    os.makedirs(REF_DST, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for rel in HOT_PATH_FILES:
            info = tar.gettarinfo(os.path.join(REF_SRC, rel), arcname=rel)
            info.mtime, info.uid, info.gid, info.uname, info.gname, info.mode = 0, 0, 0, "", "", 0o644   # reproducible
            with open(os.path.join(REF_SRC, rel), "rb") as f:
                tar.addfile(info, f)
            if verbose:
                print(rel)
    data = buf.getvalue()
    if not (os.path.exists(ARCHIVE) and open(ARCHIVE, "rb").read() == data):
        with open(ARCHIVE, "wb") as f:
            f.write(data)
    return ARCHIVE


def extract(dst):
    """Unpacks the archive into `dst` (a directory OUTSIDE the repository); returns dst."""
    with tarfile.open(ARCHIVE, "r:gz") as tar:
        for m in tar.getmembers():
            if m.name not in HOT_PATH_FILES:
                raise RuntimeError(f"unexpected member {m.name!r} in {ARCHIVE}")
        tar.extractall(dst, filter="data")
    return dst


if __name__ == "__main__":
    print(build(verbose=True))
