"""ctypes binding of ``libb200fock.so`` (C ABI in ``include/b200fock.h``).

There is no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  ``load()`` builds the library with nvcc when it is absent and
a compiler is available (the build container); on the GPU box the prebuilt ``.so``
travels with the repository snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200fock.so")

MAX_AXES = 24
MAX_CUTOFF = 64        # single-mode and diagonal gates, reductions
MAX_PAIR_CUTOFF = 27   # two-mode gates / loss channel: the packed table must fit 227 KB of shared memory
MAX_BATCH = 65535      # the batch axis is gridDim.z
GRAM_MAX_CUTOFF = 12    # b200_gram1 keeps D (D + 1) accumulators in registers

# gate kinds / rules (mirror include/b200fock.h)
GATE_DISPLACEMENT, GATE_SQUEEZE = 1, 2
DIAG_ROTATION, DIAG_KERR, DIAG_CROSS_KERR = 10, 11, 12
GATE_BEAMSPLITTER, GATE_MZ, GATE_S2 = 20, 21, 22
CHANNEL_LOSS = 30
RULE_SINGLE, RULE_SUM, RULE_DIFF = 0, 1, 2

FLAG_CONJ_B, FLAG_REAL_OUT, FLAG_REAL_IN = 1, 2, 4


class B200Error(RuntimeError):
    pass


class GatherDesc(C.Structure):
    _fields_ = [
        ("n_out_axes", C.c_int),
        ("n_red_axes", C.c_int),
        ("out_ext", C.c_int32 * MAX_AXES),
        ("out_sa", C.c_int64 * MAX_AXES),
        ("out_sb", C.c_int64 * MAX_AXES),
        ("out_sc", C.c_int64 * MAX_AXES),
        ("red_ext", C.c_int32 * MAX_AXES),
        ("red_ta", C.c_int64 * MAX_AXES),
        ("red_tb", C.c_int64 * MAX_AXES),
        ("base_a", C.c_int64),
        ("base_b", C.c_int64),
        ("base_c", C.c_int64),
    ]


class TileOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("axis1", C.c_int),
        ("axis2", C.c_int),
        ("conj", C.c_int),
        ("coef_offset", C.c_int64),
    ]


XCHG_MAX_AXES = 12
XCHG_MAX_PEERS = 32


class XchgDesc(C.Structure):
    _fields_ = [
        ("n_axes", C.c_int),
        ("n_src", C.c_int),
        ("first_src", C.c_int),
        ("ext", C.c_int32 * XCHG_MAX_AXES),
        ("ss", C.c_int64 * XCHG_MAX_AXES),
        ("ds", C.c_int64 * XCHG_MAX_AXES),
        ("run", C.c_int64),
        ("src", C.c_void_p * XCHG_MAX_PEERS),
        ("dst", C.c_void_p * XCHG_MAX_PEERS),
        ("src_base", C.c_int64 * XCHG_MAX_PEERS),
        ("dst_base", C.c_int64 * XCHG_MAX_PEERS),
    ]


class PeerFlags(C.Structure):
    _fields_ = [
        ("n_ranks", C.c_int),
        ("rank", C.c_int),
        ("flags", C.c_void_p * XCHG_MAX_PEERS),
    ]


TILE_DIAG = 3
TILE_MAX_OPS = 16
MAX_FAST_CUTOFF = 16
SMEM_LIMIT = 227 * 1024

_P, _I, _L, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "b200_version": [],
    "b200_last_error": [],
    "b200_packed_size": [_I],
    "b200_launch_count": [],
    "b200_reset_launch_count": [],
    "b200_enable_peer_access": [_I],
    "b200_gen_gate1": [_I, _I, _I, _D, _D, _P, _P, _P],
    "b200_gen_diag": [_I, _I, _I, _D, _P, _P, _P],
    "b200_gen_gate2": [_I, _I, _I, _D, _D, _P, _P, _P],
    "b200_compose_gate1": [_I, _I, _P, _P, _P, _P],
    "b200_fold_diag_gate1": [_I, _I, _P, _P, _P, _P],
    "b200_fold_diag_gate2": [_I, _I, _I, _P, _P, _P, _P, _P, _P],
    "b200_unpack_gate2": [_I, _I, _P, _P, _P],
    "b200_mul_tables": [_L, _P, _P, _I, _P, _P],
    "b200_apply_gate1": [_P, _L, _I, _L, _P, _I, _I, _L, _L, _P],
    "b200_apply_gate2": [_P, _L, _I, _L, _L, _I, _P, _I, _I, _L, _L, _P],
    "b200_apply_diag": [_P, _L, _I, _L, _L, _P, _I, _I, _L, _L, _P],
    "b200_apply_diag_multi": [_P, _L, _I, _I, _P, _P, _P, _I, _L, _L, _P],
    "b200_tile_groups": [_I],
    "b200_tile_smem_bytes": [_I, _L],
    "b200_apply_tile_pass": [_P, _L, _I, _L, _L, C.POINTER(TileOp), _I, C.POINTER(C.c_int), _P, _L, _I, _L, _L, _P],
    "b200_gather_reduce": [C.POINTER(GatherDesc), _P, _P, _P, _I, _P, _P],
    "b200_outer_axis": [_P, _P, _P, _L, _I, _I, _L, _L, _P],
    "b200_fill_zero": [_P, _L, _P],
    "b200_set_element": [_P, _L, _D, _D, _P],
    "b200_abs2": [_P, _P, _L, _P],
    "b200_norm2": [_P, _L, _P, _P, _P],
    "b200_scale": [_P, _L, _D, _D, _P, _I, _P],
    "b200_gram1_part_doubles": [_I, _I],
    "b200_gram1": [_P, _L, _I, _L, _I, _P, _P, _I, _L, _P],
    "b200_exchange_copy": [C.POINTER(XchgDesc), _I, _I, _P],
    "b200_peer_barrier": [C.POINTER(PeerFlags), C.c_uint64, _D, _P],
}
_RESTYPES = {
    "b200_last_error": C.c_char_p,
    "b200_packed_size": C.c_int64,
    "b200_launch_count": C.c_int64,
    "b200_tile_smem_bytes": C.c_int64,
    "b200_gram1_part_doubles": C.c_int64,
    "b200_reset_launch_count": None,
}

_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if needed and possible) and return the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.path.exists(LIB_PATH):
        # only a MISSING library is built here (single process, rank 0 of a multi-rank job must build
        # beforehand); a stale one is rebuilt explicitly with `python -m strawberryfields_b200.build`
        try:
            from . import build as _build

            if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) and \
                    int(os.environ.get("WORLD_SIZE", "1")) == 1:
                _build.build()
        except Exception as exc:
            raise B200Error(f"libb200fock.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build it with `python -m strawberryfields_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export the symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what or 'b200fock'} failed (code {rc}): {msg}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)


def tile_smem_bytes(D: int, coef_count: int) -> int:
    """shared memory of one b200_apply_tile_pass CTA (mirrors b200_tile_smem_bytes)"""
    s1 = D | 1
    s0 = (D * s1) | 1
    groups = 1 if D >= 14 else max(1, 32 // D)
    return (2 * groups * ((D * s0) | 1) + coef_count) * 16  # two buffers: the kernel double-buffers


def packed_size(D: int) -> int:
    return D * D + (D - 1) * D * (2 * D - 1) // 3
