"""Per-geometry rate of the gate kernels that touch the innermost axis (1e8 amplitudes, cutoff 10): a pair
gate on (axis a, innermost) for every a -- adjacent axes, short mid, rows -- and a one-mode gate on the
innermost axis, each timed with CUDA events through the C ABI.  One JSON line."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from strawberryfields_b200 import lib as L

    D, n = 10, 8
    total = D ** n
    st = torch.randn(total, dtype=torch.complex128, device="cuda")
    st /= st.norm()
    P = L.packed_size(D)
    G = torch.empty(P, dtype=torch.complex128, device="cuda")
    L.call("b200_gen_gate2", L.GATE_BEAMSPLITTER, D, 1, 0.7, 0.3, None, C.c_void_p(G.data_ptr()), None)
    U = torch.empty(D * D, dtype=torch.complex128, device="cuda")
    L.call("b200_gen_gate1", L.GATE_DISPLACEMENT, D, 1, 0.3, 0.2, None, C.c_void_p(U.data_ptr()), None)
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("B200_")}}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 32 * total / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9

    for a in range(n - 1):
        s1 = D ** (n - 1 - a)
        out["pair_axis%d_inner" % (n - 1 - a)] = round(timed(lambda: L.call(
            "b200_apply_gate2", C.c_void_p(st.data_ptr()), total, D, s1, 1, L.RULE_SUM, C.c_void_p(G.data_ptr()), 0, 1,
            total, 0, None)))
    out["pair_inner_axis1"] = round(timed(lambda: L.call(
        "b200_apply_gate2", C.c_void_p(st.data_ptr()), total, D, 1, D, L.RULE_SUM, C.c_void_p(G.data_ptr()), 0, 1,
        total, 0, None)))
    out["one_mode_inner"] = round(timed(lambda: L.call(
        "b200_apply_gate1", C.c_void_p(st.data_ptr()), total // D, D, 1, C.c_void_p(U.data_ptr()), 0, 1, total, 0, None)))
    out["one_mode_axis1"] = round(timed(lambda: L.call(
        "b200_apply_gate1", C.c_void_p(st.data_ptr()), total // (D * D), D, D, C.c_void_p(U.data_ptr()), 0, 1, total, 0, None)))
    out["pair_axis2_axis1"] = round(timed(lambda: L.call(
        "b200_apply_gate2", C.c_void_p(st.data_ptr()), total, D, D * D, D, L.RULE_SUM, C.c_void_p(G.data_ptr()), 0, 1,
        total, 0, None)))
    out["pair_axis4_axis3"] = round(timed(lambda: L.call(
        "b200_apply_gate2", C.c_void_p(st.data_ptr()), total, D, D ** 4, D ** 3, L.RULE_SUM, C.c_void_p(G.data_ptr()), 0, 1,
        total, 0, None)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
