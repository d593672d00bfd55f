"""Recipe for ``oracle/_ref``: the UNMODIFIED reference package, staged for the GPU box.

TEST INFRASTRUCTURE -- nothing under ``ocelot_b200/`` may import ``oracle/``.

The reference (ocelot-collab/ocelot v26.06.1) is pure Python, so "building" it is a byte-for-byte
copy of the package directory ``/root/reference/ocelot`` into ``oracle/_ref/ocelot``.  The copy is
git-ignored (it never enters the history) but is NOT gpurun-ignored, so it travels to the GPU box
with the snapshot, where

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's own
    ``SpaceCharge.apply`` (ocelot/cpbd/sc.py:208-251) on the box's host cores (kind "reference"),
  * ``tests/test_gpu_reference_dropin.py`` runs the CUDA class under the reference's own
    ``Navigator`` / ``track()`` (navi.py:63-98, track.py:431-504) beside the reference class.

Only ``optics/data`` (25 MB of X-ray tables, unrelated to tracking) and byte-code caches are left
out.  ``/root/reference`` exists in the build container only: on the GPU box the prebuilt copy is
used as is, and ``available()`` says whether it is there.

    python -m oracle.build_ref          # (re)stage the copy
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
SRC = "/root/reference/ocelot"
_SKIP = {"__pycache__", "data"}          # optics/data is the only directory called "data"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ocelot", "cpbd", "sc.py"))


def build(force: bool = False) -> str | None:
    """Stage the copy when the reference checkout is present; return its root (or None)."""
    if not os.path.isdir(SRC):
        return REF_ROOT if available() else None
    dst = os.path.join(REF_ROOT, "ocelot")
    if force and os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(REF_ROOT, exist_ok=True)
    shutil.copytree(SRC, dst, ignore=lambda d, names: [n for n in names if n in _SKIP], dirs_exist_ok=True)
    # the staged hot-path sources must be the reference's, byte for byte
    for rel in ("cpbd/sc.py", "cpbd/coord_transform.py", "cpbd/physics_proc.py", "cpbd/navi.py", "cpbd/track.py"):
        if not filecmp.cmp(os.path.join(SRC, rel), os.path.join(dst, rel), shallow=False):
            raise RuntimeError(f"oracle/_ref/ocelot/{rel} differs from the reference")
    return REF_ROOT


def import_reference():
    """Import the staged reference package (``import ocelot``) and return the module.

    The staged copy is put on ``sys.path`` only when no ``ocelot`` is importable already (in the build
    container ``/root/reference`` may be on PYTHONPATH: same files)."""
    if "ocelot" in sys.modules:
        return sys.modules["ocelot"]
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python -m oracle.build_ref` in the build container")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import ocelot  # noqa: F401
    return sys.modules["ocelot"]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
