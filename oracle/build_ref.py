"""Recipe for ``oracle/_ref``: a runnable snapshot of the reference's own hot-path modules.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package.

The reference is pure Python (no setup.py, nothing to compile), and ``/root/reference`` does not exist on the
GPU box.  So that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the UNMODIFIED reference
modules there (kind "reference") instead of the oracle restatement (kind "port"), this recipe copies the import
closure of the hot path -- found by importing it with the stubs of ``oracle/ref_import.py`` -- verbatim into
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the box like a built .so, and stays out of the
repository history).  ``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.

    python -m oracle.build_ref            # copies, prints the manifest
"""
import filecmp
import os
import shutil

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


def build(verbose=False):
    """Copies the files if the reference checkout is present; returns the destination or None."""
    if not os.path.isdir(os.path.join(REF_SRC, "mmskeleton")):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "mmskeleton")) else None
    for rel in HOT_PATH_FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
        if verbose:
            print(rel)
    with open(os.path.join(REF_DST, "README"), "w") as f:
        f.write("Verbatim copies of reference files made by oracle/build_ref.py (git-ignored; never edit, never commit).\n")
    return REF_DST


if __name__ == "__main__":
    print(build(verbose=True))
