#!/usr/bin/env python
"""Vendor the UNMODIFIED reference Python package into oracle/_ref (git-ignored, not gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (koskja/linalg-solver) is pure Python plus a Rust pyo3 module
that cannot be built in this image (no cargo), so "building" it is a byte-for-byte copy of
/root/reference/linalg_solver next to oracle/standin/linalg_helper.py, the stand-in for the Rust module (only what
`import linalg_solver` and the elimination path touch).  Nothing is copied into tracked files: oracle/_ref/ is listed
in .gitignore and travels to the GPU box with the snapshot, where `bench.py` times it on a small sample
(`cpu_baseline.reference_unmodified`, kind "reference") beside the pinned port.  Run by __graft_entry__.build() when
/root/reference exists; a no-op elsewhere (the GPU box uses the copy made here).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/linalg_solver"
DST = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/build_ref.py: %s not present, keeping oracle/_ref as it is" % SRC)
        return os.path.isdir(os.path.join(DST, "linalg_solver"))
    os.makedirs(DST, exist_ok=True)
    pkg = os.path.join(DST, "linalg_solver")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copyfile(os.path.join(HERE, "standin", "linalg_helper.py"), os.path.join(DST, "linalg_helper.py"))
    # the copy must be the reference, byte for byte
    cmp = filecmp.dircmp(SRC, pkg, ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, (cmp.diff_files, cmp.left_only)
    if verbose:
        print("oracle/build_ref.py: %d files of the unmodified reference under oracle/_ref" % len(cmp.same_files))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
