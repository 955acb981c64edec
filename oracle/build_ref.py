"""Recipe for ``oracle/_ref``: an installed copy of the UNMODIFIED reference package, so that the
reference arm of ``bench.py`` (``--impl reference``) can time the reference's own fock backend on the
GPU box's host cores (``/root/reference`` does not exist there; ``oracle/_ref`` is git-ignored but
travels with the repository snapshot).  TEST / MEASUREMENT INFRASTRUCTURE: nothing under
``strawberryfields_b200`` imports it.

    python -m oracle.build_ref            # run in the build container (needs /root/reference)

The reference is pure Python (``setup.py``, no native code): the install is
``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of the tree>``
(from a copy under /tmp because /root/reference is read-only); if pip cannot build it offline the
package directory is copied as it is -- which is all an install of a pure-Python package does.  Its
third-party imports that are absent from the image (thewalrus, blackbird, xir, xcc) are stubbed at
import time by ``oracle/ref_shim.py``, which also binds the five ``thewalrus.fock_gradients`` functions
to the restated recursions of ``oracle/gates.py``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"


def build(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SOURCE, "strawberryfields")):
        if os.path.isdir(os.path.join(TARGET, "strawberryfields")):
            return TARGET  # the GPU box: use what travelled
        raise RuntimeError("%s not present and %s not built" % (SOURCE, TARGET))
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET)
    how = "pip"
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns("doc", "tests", ".git", "examples"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not os.path.isdir(os.path.join(TARGET, "strawberryfields")):
            how = "copy (pip could not build offline: %s)" % (res.stderr.strip().splitlines() or ["?"])[-1][:200]
            shutil.rmtree(TARGET, ignore_errors=True)
            os.makedirs(TARGET)
            shutil.copytree(os.path.join(src, "strawberryfields"), os.path.join(TARGET, "strawberryfields"))
    with open(os.path.join(TARGET, "HOW"), "w") as f:
        f.write("installed from %s by oracle/build_ref.py: %s\n" % (SOURCE, how))
    if verbose:
        print("oracle/_ref:", how)
    return TARGET


if __name__ == "__main__":
    build()
