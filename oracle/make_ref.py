#!/usr/bin/env python3
"""Recipe for oracle/_ref/: an installed copy of the UNMODIFIED reference (Mutation-Simulator 3.0.2).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python; `pip install --target oracle/_ref` of
/root/reference (from a scratch copy, the tree is read-only; --no-deps because its `pyfaidx` dependency is not
installable here — oracle/pyfaidx_standin supplies the twelve calls the reference makes).  oracle/_ref/ is
git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot, so that
`bench.py` can time the real reference on the box's host cores (`cpu_baseline.kind = "reference"`) and
`bench.py --impl reference` can run it as the reference arm.  Nothing in the product imports it.

    python oracle/make_ref.py          (only where /root/reference exists; __graft_entry__.build() calls it)
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
DEST = HERE / "_ref"


def available() -> bool:
    return (DEST / "mutation_simulator" / "mutator.py").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards."""
    if available() and not force:
        return True
    if not (REF_SRC / "mutation_simulator" / "mutator.py").exists():
        return False
    DEST.mkdir(exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "*.pdf", "*.png", "data"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", str(DEST), "--upgrade", str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not available():
            # flit_core missing etc.: the package is a flat directory of .py files, an install is a copy
            shutil.copytree(src / "mutation_simulator", DEST / "mutation_simulator", dirs_exist_ok=True)
            (DEST / "INSTALL_NOTE.txt").write_text("pip install failed, package directory copied instead:\n" + r.stderr[-2000:])
    return available()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference here)")
