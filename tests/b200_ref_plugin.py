"""pytest plugin (build container only): run the REFERENCE's own fock-marked test-suite against
`b200fock`.

    PYTHONPATH=/root/repo:/root/repo/tests python -m pytest -p b200_ref_plugin -p no:cacheprovider \
        -m fock /root/reference/tests/backend -q

Before the reference's conftest is imported the plugin (1) installs the import shim for the absent
third-party packages (oracle/ref_shim.py), (2) swaps `FockBackend` for `B200FockBackend` in
`strawberryfields.backends` (class attribute and the "fock" entry of the backend registry), so every
test parametrised over the fock backend drives our plugin instead, and (3) points the plugin at the
numpy double of the C ABI (no GPU here; /root/reference does not exist on the GPU box).  What this
checks is the drop-in contract: API, error behaviour, mode bookkeeping and numerics of the host logic."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_shim  # noqa: E402

sf = ref_shim.install()
import strawberryfields_b200 as plugin  # noqa: E402
from strawberryfields_b200 import circuit, lib  # noqa: E402
from fake_lib import FakeLib  # noqa: E402

lib._lib = FakeLib()
circuit._TEST_HOST_MODE = True

import strawberryfields.backends as sfb  # noqa: E402
import strawberryfields.backends.fockbackend as fb  # noqa: E402


class B200AsFock(plugin.B200FockBackend):
    """what the reference's tests see under the name `FockBackend`"""

    short_name = "fock"

    def begin_circuit(self, num_subsystems, **kwargs):
        kwargs.pop("batch_size", None)        # the reference fock backend ignores it
        kwargs.setdefault("strict_purity", True)  # reproduce the reference's pure/mixed representation (F7)
        # the plugin default (lazy vacuum + deferred program); B200_REF_SUITE_LAZY=0 runs the eager path
        kwargs.setdefault("lazy_vacuum", os.environ.get("B200_REF_SUITE_LAZY", "1") != "0")
        return super().begin_circuit(num_subsystems, **kwargs)


fb.FockBackend = B200AsFock
sfb.FockBackend = B200AsFock
sfb.local_backends["fock"] = B200AsFock
if hasattr(sfb, "supported_backends"):
    sfb.supported_backends["fock"] = B200AsFock
