"""Stage the UNMODIFIED reference package for the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (MarieRoald/matcouply) is pure Python: "building" it means
making its package importable.  This script copies `/root/reference/src/matcouply` (read-only, present in the build
container only) byte for byte into the git-ignored `oracle/_ref/matcouply`, which travels to the GPU box with the
snapshot like the built `.so` files do.  Nothing under `oracle/_ref/` is tracked, edited or imported by the product.
The reference's third-party dependency `tensorly` is not installable offline; `oracle/_tl_standin/tensorly` (1:1 NumPy
aliases) stands in for it at import time.

    python oracle/stage_ref.py            # no-op (exit 0) when /root/reference is absent
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/matcouply"
DST = os.path.join(HERE, "_ref", "matcouply")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"stage_ref: {SRC} not present (GPU box): using whatever is staged in {DST}")
        return os.path.isdir(DST)
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    cmp = filecmp.dircmp(SRC, DST, ignore=["__pycache__"])
    assert not cmp.left_only and not cmp.diff_files, (cmp.left_only, cmp.diff_files)
    with open(os.path.join(HERE, "_ref", "STAGED_FROM"), "w") as f:
        f.write(f"{SRC} (unmodified copy; see oracle/stage_ref.py)\n")
    if verbose:
        print(f"stage_ref: staged {SRC} -> {DST}")
    return True


def load_reference():
    """`matcouply.decomposition` of the staged reference, or None when nothing is staged.  Puts the tensorly stand-in
    and oracle/_ref on sys.path (only bench.py's CPU arm and the tests call this)."""
    if not os.path.isfile(os.path.join(DST, "decomposition.py")):
        return None
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", "b2_numba_cache"))
    for p in (os.path.join(HERE, "_ref"), os.path.join(HERE, "_tl_standin")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib

    return importlib.import_module("matcouply.decomposition")


if __name__ == "__main__":
    stage()
