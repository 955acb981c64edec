"""pytest plugin (build container only): run the REFERENCE's own Blackbird I/O tests on ``strawberryfields_b200.io``.

    cd /tmp && PYTHONPATH=/root/repo:/root/repo/tests python -m pytest -p b200_ref_io_plugin -p no:cacheprovider \
        /root/reference/tests/frontend/io -q

Installs the import shim for the absent third-party packages (oracle/ref_shim.py) and then replaces the inert
``blackbird`` / ``xir`` stand-ins by ``tests/blackbird_facade.py`` / ``tests/xir_facade.py``: the reference's ``sf.load`` / ``sf.save`` /
``io.to_blackbird`` / ``io.to_program`` run unmodified, every script they parse or write goes through our io."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_shim  # noqa: E402

ref_shim.install()
import blackbird_facade  # noqa: E402
import xir_facade  # noqa: E402

blackbird_facade.install()
xir_facade.install()
