"""Where the end-to-end step of bench.py goes (config 2 through the plugin API with device-resident parameters):
wall time of the three phases of a step -- gate calls (recorded by the deferred program), ``state()`` (replay:
table generation, folds, gate passes) and the result reads -- next to the GPU time of the whole step, and the same
step with host-float parameters (table cache hits).  One GPU, seconds.

    python tools/e2e_breakdown.py [--steps 10] > gpurun_out/e2e_breakdown.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strawberryfields_b200 import B200FockBackend, DeviceParams, lib  # noqa: E402
from strawberryfields_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--modes", type=int, default=8)
    ap.add_argument("--cutoff", type=int, default=10)
    args = ap.parse_args()
    n, D = args.modes, args.cutoff
    calls = W.config2_circuit(n)
    two_param = ("squeeze", "displacement", "beamsplitter", "mzgate", "two_mode_squeeze")

    def params_of(c):
        vals = [np.asarray([x], dtype=np.float64) for x in c[1:] if not isinstance(x, (int, np.integer))]
        vals += [np.zeros(1)] * (2 - len(vals))
        return np.stack(vals[:2])

    ptab = np.ascontiguousarray(np.stack([params_of(c) for c in calls]))
    pinned = torch.from_numpy(ptab).pin_memory()
    outcomes = [[0] * n] + [[k if j == m else 0 for j in range(n)] for m in range(n) for k in (1, 2)]
    be = B200FockBackend()
    be.begin_circuit(n, cutoff_dim=D)
    handle = lib.load(build_if_missing=False)

    def step(device_params, t):
        t0 = time.perf_counter()
        dev = pinned.to("cuda", non_blocking=True) if device_params else None
        be.reset(pure=True)
        for i, c in enumerate(calls):
            modes = [x for x in c[1:] if isinstance(x, (int, np.integer))]
            if device_params:
                getattr(be, c[0])(*([DeviceParams(dev[i])] + ([None] if c[0] in two_param else []) + modes))
            else:
                getattr(be, c[0])(*c[1:])
        t1 = time.perf_counter()
        st = be.state()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        res = np.array([st.trace()] + [st.fock_prob(o) for o in outcomes])
        t4 = time.perf_counter()
        for k, v in (("calls_ms", t1 - t0), ("state_host_ms", t2 - t1), ("state_gpu_tail_ms", t3 - t2), ("reads_ms", t4 - t3)):
            t[k] = t.get(k, 0.0) + v * 1e3
        return res

    out = {"modes": n, "cutoff": D, "steps": args.steps, "outcomes_read": len(outcomes) + 1}
    for label, dp in (("device_params", True), ("host_params_cached_tables", False)):
        for _ in range(3):
            step(dp, {})
        torch.cuda.synchronize()
        t = {}
        l0 = int(handle.b200_launch_count())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step(dp, t)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) * 1e3 / args.steps
        out[label] = dict({k: v / args.steps for k, v in t.items()}, wall_ms_per_step=wall,
                          event_ms_per_step=e0.elapsed_time(e1) / args.steps,
                          launches_per_step=(int(handle.b200_launch_count()) - l0) / args.steps)
    # the same step without the mid-step synchronize (what bench.py times)
    for label, dp in (("device_params_no_sync", True),):
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            dev = pinned.to("cuda", non_blocking=True)
            be.reset(pure=True)
            for i, c in enumerate(calls):
                modes = [x for x in c[1:] if isinstance(x, (int, np.integer))]
                getattr(be, c[0])(*([DeviceParams(dev[i])] + ([None] if c[0] in two_param else []) + modes))
            st = be.state()
            np.array([st.trace()] + [st.fock_prob(o) for o in outcomes])
        torch.cuda.synchronize()
        out[label] = {"wall_ms_per_step": (time.perf_counter() - w0) * 1e3 / args.steps}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
