"""b200fock -- a Fock-basis simulator backend for NVIDIA B200 (sm_100a).

Drop-in for the Strawberry Fields ``fock`` backend on the hot path
``strawberryfields/backends/fockbackend``:

    import strawberryfields as sf, strawberryfields_b200
    strawberryfields_b200.register()
    eng = sf.Engine("b200fock", backend_options={"cutoff_dim": 10})

or stand-alone through the same backend API (``B200FockBackend``).  All state updates
are hand-written CUDA kernels in ``libb200fock.so`` (C ABI: ``include/b200fock.h``);
there is no CPU fallback.
"""
from .autodiff import TorchCircuit  # noqa: F401
from .backend import B200FockBackend, register  # noqa: F401
from .circuit import DeviceCircuit, DeviceParams  # noqa: F401
from .states import B200FockState  # noqa: F401
from . import io  # noqa: F401  (Blackbird / XIR program I/O, state checkpoints)

__version__ = "0.2.0"

try:  # make sf.Engine("b200fock") work as soon as the package is imported next to SF
    register()
except Exception:  # Strawberry Fields not installed (e.g. on the GPU box): stand-alone use only
    pass
