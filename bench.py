#!/usr/bin/env python
"""Benchmark of the Fock-backend hot path (BASELINE.json metric: Fock amp-gate updates/s +
achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c1|c2|c3|c4] [--modes M] [--fuse fold|tile|off]
                    [--exchange auto|p2p|push|nccl] [--exchange-overlap G] [--from-vacuum]
                    [--no-parity] [--no-ten-mode] [--no-cpu-baseline]

Workload (N = 1): BASELINE config 2 -- 8-mode pure state, cutoff 10 (1e8 complex128
amplitudes, 1.6 GB), Sgate + Dgate on every mode and a random 8-mode rectangular
interferometer = 80 gates (8 S + 8 D + 36 R + 28 BS).  A "step" is one pass of the whole
circuit over the resident state.  One amp-gate update = one stored amplitude passing
through one gate of the call list (fused or not), so a step is 80 x 1e8 updates.
Under torchrun (N > 1) the workload is BASELINE config 5: ONE 9-mode state (1e9 amplitudes)
sharded over the N GPUs ("scaling": "strong"; `--modes 10` gives the 160 GB state on 8 GPUs).
`--workload c1|c3|c4` run the other BASELINE configs on one GPU (boson sampling, mixed state +
loss + MeasureFock, batched QNN layer).

* ``value``: updates/s with the state resident in HBM, CUDA-event timed, max over ranks.
* ``e2e``: the same metric through the reference-facing plugin API with its default options
  (``B200FockBackend``: ``reset`` .. gate calls .. ``state()``) and host buffers: every step
  copies the step's gate-parameter table from pinned host memory to the device and reads the
  requested result (trace + a list of Fock probabilities) back to the host.
* ``roofline``: the dominant kernel (k_apply_blocks) -- algorithmic bytes (32 B per
  amplitude per pass) / CUDA-event time per launch, against MEASURED_PEAKS.json; ``by_pass``
  gives every pass family, ``hbm_GBps_per_gpu_whole_step`` the whole step.
* ``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference fock backend (oracle/_ref,
  installed by oracle/build_ref.py) on the host cores, on a full-size gate prefix of the same
  workload (one BSgate / the first S + D + R + BS of the circuit); the oracle port on a reduced
  size only if oracle/_ref is missing.
* multi-rank lines additionally carry ``parity`` (single-photon transfer on the sharded state,
  the sharded state against an unsharded run of the same circuit on rank 0, a seeded
  MeasureFock), ``single_gpu_same_workload_ms`` + ``strong_efficiency``, ``exchange`` (NVLink
  GB/s per direction) and, at N = 8, ``ten_mode`` (the 10-mode / 160 GB state).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fock_amp_gate_updates_per_s"
UNIT = "updates/s"
DEVICE = "cuda"  # where the timing scalars and the e2e parameter table live (the CPU dry-run test overrides it)


# ------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, elements_per_launch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full captures
    (profiles/r0*_ncu_traffic.json, taken on config 2: 1e8 amplitudes per launch; the streaming kernel's
    capture is round 1's, the kernel is unchanged); None for any other launch size -- never extrapolated."""
    if elements_per_launch != 10 ** 8:
        return None
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            table = json.load(f)["bytes_per_launch"]
        vals = [val for key, val in table.items() if kernel in key]
        if vals:
            return vals[-1]
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                smax = float(f[2])
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": smax, "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def cpu_reference_run(n_modes, D, steps, warmup):
    """The oracle port (reference loop structure, numba selection-rule kernels) on the host:
    one step = the config-2 circuit generator at ``n_modes`` modes.  Only used when the installed copy of
    the reference (oracle/_ref) is missing."""
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200 import workloads as W

    calls = W.config2_circuit(n_modes, seed=42)
    be = OracleBackend(style="reference")
    be.begin_circuit(n_modes, cutoff_dim=D)
    # numba JIT warm-up for this ndim (the reference pays 9-15 s per new signature)
    W.run_calls(be, [c for c in calls if c[0] == "beamsplitter"][:2])
    for _ in range(warmup):
        be.reset()
        W.run_calls(be, calls)
    t0 = time.perf_counter()
    for _ in range(steps):
        be.reset()
        W.run_calls(be, calls)
    dt = time.perf_counter() - t0
    updates = len(calls) * D ** n_modes * steps
    return updates / dt, dt / steps, len(calls)


# The prefix of the config-2 call list the reference is timed on at FULL size (8 modes, cutoff 10, 1e8
# amplitudes): one gate of every kind on the path.  The whole circuit takes the reference ~40 min
# (25-35 s per gate, measured in the build container), BASELINE.md section 4.3 prescribes a fixed prefix.
REF_PREFIX = ("squeeze", "displacement", "rotation", "beamsplitter")


def reference_prefix_calls(n_modes):
    from strawberryfields_b200 import workloads as W

    calls, out = W.config2_circuit(n_modes, seed=42), []
    for kind in REF_PREFIX:
        out.append(next(c for c in calls if c[0] == kind))
    return out


def unmodified_reference_run(n_modes, D, kinds, steps):
    """Time the UNMODIFIED reference fock backend (``strawberryfields.backends.fockbackend.FockBackend`` from
    oracle/_ref, installed by oracle/build_ref.py; absent third-party imports stubbed by oracle/ref_shim.py)
    on the host: ``steps`` times the gates ``kinds`` of the config-2 prefix on a full-size state.  The numba
    kernels are compiled beforehand on a cutoff-2 state of the same rank (the JIT signature depends on the
    rank only).  Returns (updates/s, seconds per step, gates per step) or None when the installed
    reference is missing."""
    try:
        from oracle import ref_shim

        if not ref_shim.available():
            return None
        ref_shim.install()
        from strawberryfields.backends.fockbackend import FockBackend
    except Exception as exc:  # pragma: no cover - reported by the caller
        sys.stderr.write("reference arm: the installed reference is not importable (%r)\n" % (exc,))
        return None
    from strawberryfields_b200 import workloads as W

    calls = [c for c in reference_prefix_calls(n_modes) if c[0] in kinds]
    warm = FockBackend()
    warm.begin_circuit(n_modes, cutoff_dim=2)
    W.run_calls(warm, calls)
    be = FockBackend()
    be.begin_circuit(n_modes, cutoff_dim=D)
    W.run_calls(be, [c for c in reference_prefix_calls(n_modes) if c[0] == "displacement"])  # leave the vacuum
    t0 = time.perf_counter()
    for _ in range(steps):
        W.run_calls(be, calls)
    dt = time.perf_counter() - t0
    return len(calls) * D ** n_modes * steps / dt, dt / steps, len(calls)


def reference_arm(args):
    """``--impl reference``: the reference's own CPU implementation of the path on this box's host cores --
    the unmodified reference (oracle/_ref) on a full-size gate prefix of the SAME workload (config 2: 8 modes,
    cutoff 10); the oracle port on a reduced size only if the installed reference is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_modes, D = 8, 10
    steps, warmup = 1, 0   # one step is ~2 minutes of host time; the JIT warm-up runs on a cutoff-2 state
    res = unmodified_reference_run(n_modes, D, REF_PREFIX, steps)
    if res is not None:
        val, per_step, ngates = res
        kind, same = "reference", True
        sample = ("the first %s of the config-2 call list (%d gates) on the full-size state: %d modes, cutoff %d, "
                  "%d amplitudes; %.0f s per step" % (" + ".join(REF_PREFIX), ngates, n_modes, D, D ** n_modes, per_step))
        note = ("unmodified strawberryfields FockBackend (oracle/_ref, installed by oracle/build_ref.py; thewalrus "
                "gate recursions bound to oracle/gates.py); single-threaded like the reference (numba kernels "
                "without parallel=True, D x D numpy)")
    else:
        n_modes = 6
        val, per_step, ngates = cpu_reference_run(n_modes, D, 1, 0)
        kind, same = "port", False
        sample = "config-2 generator at %d modes, cutoff %d (%d amplitudes, %d gates) per step" % (
            n_modes, D, D ** n_modes, ngates)
        note = "oracle/_ref missing: port = oracle/fock_oracle.py style='reference' on a reduced size"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: 8-mode pure state, Sgate+Dgate per mode + random rectangular "
                               "interferometer; cutoff 10, 1e+08 stored complex128 elements; the reference arm "
                               "times a gate prefix: " + sample,
                   "same_config": same and args.gpus == 1, "requested_steps": args.steps,
                   "requested_warmup": args.warmup},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "note": note},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.gpus > 1:
        # the sharded arm runs config 5 (the same circuit family on 9 modes, 1e9 amplitudes): a 9-mode prefix
        # would take the single-threaded reference 16 GB and ~7 minutes per step, so its rate is taken on the
        # 8-mode sample -- updates/s, the metric, does not depend on the mode count for the reference
        line["config"]["note"] = ("--gpus %d: the b200 arm runs config 5 (9 modes, one state sharded); the reference "
                                  "rate is measured on the 8-mode sample of the same circuit family" % args.gpus)
    print(json.dumps(line))


def c1_arm(args):
    """BASELINE config 1 (examples/boson_sampling.py, 4 modes, cutoff 7): a launch-latency-bound
    circuit.  One step = begin_circuit + 4 preparations + 12 gates + all_fock_probs() on the host.
    The reference turns the state into a 7^8-element density tensor at the first Fock preparation
     and needs ~15 s; b200fock keeps the 7^4-amplitude ket."""
    import torch

    from strawberryfields_b200 import B200FockBackend, lib
    from strawberryfields_b200 import workloads as W

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    handle = lib.load(build_if_missing=False)
    calls = W.config1_circuit()
    n_gates = len([c for c in calls if not c[0].startswith("prepare")])
    be = B200FockBackend()

    def step():
        be.begin_circuit(4, cutoff_dim=7)
        W.run_calls(be, calls)
        return be.state().all_fock_probs()

    for _ in range(max(args.warmup, 3)):
        probs = step()
    torch.cuda.synchronize()
    handle.b200_reset_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        probs = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    launches = int(handle.b200_launch_count())
    ok = abs(probs[1, 1, 0, 1] - 0.174689160486) < 1e-10 and abs(probs[2, 0, 0, 1] - 0.106441927246) < 1e-10
    line = {
        "metric": METRIC, "value": n_gates * 7 ** 4 / dt, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": "BASELINE config 1: examples/boson_sampling.py, 4 modes, cutoff 7, pure ket of 2401 "
                               "amplitudes (the reference holds a 7^8 density tensor), 12 gates + all_fock_probs()",
                   "bound": "kernel-launch / host latency, not HBM"},
        "e2e": {"value": n_gates * 7 ** 4 / dt, "unit": UNIT, "h2d_bytes_per_step": 4 * 7 * 16,
                "d2h_bytes_per_step": 8 * 7 ** 4},
        "gpu_launches": launches, "circuit_ms": dt * 1e3,
        "golden_probabilities_ok": bool(ok),
    }
    if not args.no_cpu_baseline:
        from oracle.fock_oracle import OracleBackend

        ob = OracleBackend(style="reference")
        t0 = time.perf_counter()
        ob.begin_circuit(4, cutoff_dim=7)
        W.run_calls(ob, calls)
        ob.state().all_fock_probs()
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_gates * 7 ** 8 / tc, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "the whole config-1 circuit on the reference's 7^8-element mixed "
                                          "representation, incl. numba JIT: %.1f s" % tc,
                                "circuit_ms": tc * 1e3}
    print(json.dumps(line))



# ------------------------------------------------------------------------------ sharded parity
def sharded_parity(be, calls, n_modes, D, rank, world, lazy, exchange, n_probs=1000, time_steps=3):
    """Correctness of the sharded GPU path, recorded in the bench line itself (every rank calls this).

    (a) single-photon transfer: |1_k> through the passive part of the circuit (the interferometer) on
        the SHARDED state gives <1_j|psi> = U[j, k] -- the analytic N x N mode transformation;
    (b) the state the sharded circuit ends in is compared with an UNSHARDED run of the same circuit on
        rank 0's GPU: trace, ``n_probs`` sampled Fock probabilities, every single-mode marginal
        (tolerance of the north star: 1e-12 absolute);
    (c) a seeded MeasureFock gives the same outcome, and the same post-measurement probabilities.
    Also times the unsharded circuit on rank 0 (the same-workload single-GPU reference of the strong
    scaling figure).  Returns (parity dict, single-GPU ms per step or None)."""
    import torch
    import torch.distributed as dist

    from strawberryfields_b200 import B200FockBackend
    from strawberryfields_b200 import workloads as W

    # (a) ------------------------------------------------------------------------------------------
    passive = [c for c in calls if c[0] in ("rotation", "beamsplitter")]
    U = W.interferometer_unitary(n_modes, passive)
    k = n_modes // 2
    be.reset()
    be.prepare_fock_state(1, k)
    W.run_calls(be, passive)
    sp_err = 0.0
    for j in range(n_modes):
        o = [0] * n_modes
        o[j] = 1
        sp_err = max(sp_err, abs(complex(be.circuit.element(o)[0]) - U[j, k]))
    sp_norm = abs(be.state().trace() - 1.0)

    # (b) ------------------------------------------------------------------------------------------
    rng = np.random.RandomState(123)
    outcomes = [[0] * n_modes]
    while len(outcomes) < n_probs:  # low photon numbers: where the probability mass is
        o = rng.choice(3, size=n_modes, p=[0.7, 0.2, 0.1])
        outcomes.append([int(x) for x in o])
    be.reset()
    W.run_calls(be, calls)
    st = be.state()
    got = {"trace": st.trace(), "probs": np.array([st.fock_prob(o) for o in outcomes]),
           "marg": np.array([np.real(np.diag(st.reduced_dm([m]))) for m in range(n_modes)])}
    np.random.seed(7)
    got["outcome"] = be.measure_fock([0, n_modes - 1])
    st2 = be.state()
    got["post"] = np.array([st2.fock_prob(o) for o in outcomes[:64]])

    res = torch.zeros(8, dtype=torch.float64, device=DEVICE)
    if rank == 0:
        ref = B200FockBackend()
        ref.begin_circuit(n_modes, cutoff_dim=D)
        W.run_calls(ref, calls)
        ref.circuit._flush()
        torch.cuda.synchronize() if DEVICE == "cuda" else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(time_steps):   # steady state, like the timed region of the sharded arm
            W.run_calls(ref, calls)
            ref.circuit._flush()
        e1.record()
        torch.cuda.synchronize() if DEVICE == "cuda" else None
        single_ms = e0.elapsed_time(e1) / time_steps
        ref.reset()
        W.run_calls(ref, calls)
        rst = ref.state()
        want_probs = np.array([rst.fock_prob(o) for o in outcomes])
        want_marg = np.array([np.real(np.diag(rst.reduced_dm([m]))) for m in range(n_modes)])
        np.random.seed(7)
        want_out = ref.measure_fock([0, n_modes - 1])
        rst2 = ref.state()
        want_post = np.array([rst2.fock_prob(o) for o in outcomes[:64]])
        res[0] = abs(got["trace"] - rst.trace())
        res[1] = np.abs(got["probs"] - want_probs).max()
        res[2] = np.abs(got["marg"] - want_marg).max()
        res[3] = 1.0 if np.array_equal(got["outcome"], want_out) else 0.0
        res[4] = np.abs(got["post"] - want_post).max()
        res[5] = single_ms
        res[6] = float(np.sum(want_probs))
        del ref, rst, rst2
        if DEVICE == "cuda":
            torch.cuda.empty_cache()
    if world > 1:
        dist.broadcast(res, src=0)
    r = res.cpu().numpy()
    parity = {
        "max_abs_err": float(max(sp_err, r[0], r[1], r[2], r[4])),
        "tolerance": 1e-12,
        "single_photon_transfer_max_abs_err": float(sp_err), "single_photon_norm_err": float(sp_norm),
        "vs_unsharded_run_on_rank0": {"trace_abs_err": float(r[0]), "fock_probs_compared": len(outcomes),
                                      "fock_prob_max_abs_err": float(r[1]), "probability_mass_compared": float(r[6]),
                                      "single_mode_marginals_max_abs_err": float(r[2]),
                                      "post_measurement_prob_max_abs_err": float(r[4])},
        "measure_equal": bool(r[3] == 1.0), "measure_seed": 7, "measured_modes": [0, n_modes - 1],
        "measure_outcome": [int(x) for x in np.asarray(got["outcome"]).reshape(-1)],
    }
    return parity, float(r[5])


def ten_mode_block(world, rank, D, exchange, steps=3):
    """BASELINE config 5, second half: the 10-mode cutoff-10 pure state (1e10 amplitudes, 160 GB) sharded
    over 8 GPUs -- ms per circuit, exchanges, and the single-photon transfer check on the sharded state
    (the reference cannot hold 160 GB)."""
    import torch
    import torch.distributed as dist

    from strawberryfields_b200 import B200FockBackend
    from strawberryfields_b200 import workloads as W

    n = 10
    calls = W.config2_circuit(n, seed=42)
    be = B200FockBackend()
    be.begin_circuit(n, cutoff_dim=D, shard=True, exchange=exchange)
    W.run_calls(be, calls)
    be.circuit._flush()
    dist.barrier()
    torch.cuda.synchronize()
    x0 = be.circuit.exchanges
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        W.run_calls(be, calls)
        be.circuit._flush()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=DEVICE)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ex = (be.circuit.exchanges - x0) / steps
    passive = [c for c in calls if c[0] in ("rotation", "beamsplitter")]
    U = W.interferometer_unitary(n, passive)
    be.reset()
    be.prepare_fock_state(1, 4)
    W.run_calls(be, passive)
    err = 0.0
    for j in range(n):
        o = [0] * n
        o[j] = 1
        err = max(err, abs(complex(be.circuit.element(o)[0]) - U[j, 4]))
    norm_err = abs(be.state().trace() - 1.0)
    be.circuit._bufs = be.circuit._buf = be.circuit._backing = be.circuit._peers = None
    del be
    torch.cuda.empty_cache()
    t = float(ms.item())
    return {"modes": n, "amplitudes": 10 ** n, "state_GB": 16 * 10 ** n / 1e9, "gates": len(calls), "steps": steps,
            "ms_per_step": t, "updates_per_s": len(calls) * 10 ** n / (t * 1e-3), "exchanges_per_step": ex,
            "parity": {"single_photon_transfer_max_abs_err": float(err), "norm_err": float(norm_err),
                       "tolerance": 1e-12}}


# ------------------------------------------------------------------------------ b200 arm
def b200_arm(args):
    import torch
    import torch.distributed as dist

    from strawberryfields_b200 import B200FockBackend, lib
    from strawberryfields_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        if DEVICE == "cuda":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:  # CPU dry-run of the control flow (tests/test_bench_dryrun.py)
            dist.init_process_group("gloo")
    handle = lib.load(build_if_missing=False)

    # N = 1: BASELINE config 2 (8 modes).  N > 1: BASELINE config 5 -- ONE 9-mode state sharded over
    # the N GPUs (strong scaling), gates on sharded modes served by the all-to-all axis exchange.
    sharded = world > 1
    D = args.cutoff
    shard_kw = {"shard": True, "exchange": args.exchange, "exchange_overlap": args.exchange_overlap} if sharded else {}
    if args.workload == "c3":      # BASELINE config 3: 4-mode MIXED state, S/BS layers + LossChannel(0.9)
        n_modes = args.modes or 4
        calls = W.config3_circuit(n_modes, seed=42)
        elements = D ** (2 * n_modes)
        shard_kw = {"pure": False}
        wl_name = ("3: %d-mode mixed state (density matrix), Sgate/BSgate layers + LossChannel(0.9) per mode, "
                   "then MeasureFock on all modes" % n_modes)
    elif args.workload == "c4":    # BASELINE config 4: QNN layer, 6 modes, batch of 64 pure states
        n_modes = args.modes or 6
        calls = W.config4_circuit(n_modes, batch=args.batch, seed=42)
        elements = args.batch * D ** n_modes
        shard_kw = {"batch_size": args.batch}
        wl_name = "4: CV-QNN layer, %d modes, batch of %d pure states, per-entry weights" % (n_modes, args.batch)
    else:
        n_modes = args.modes or (9 if sharded else 8)
        calls = W.config2_circuit(n_modes, seed=42)
        elements = D ** n_modes                    # whole-job amplitudes
        wl_name = ("5 (one state sharded over %d GPUs)" % world if sharded else "2") + \
            ": %d-mode pure state, Sgate+Dgate per mode + random rectangular interferometer" % n_modes
    local_elements = elements // world if sharded else elements
    updates_per_step = len(calls) * elements

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: `value` ---------------------------------------------------
    fuse = {"tile": "tile", "fold": "fold", "off": False}[args.fuse]
    # --from-vacuum: every step is a whole program run -- reset to |0..0>, then the gates -- with the
    # lazy-vacuum option (modes stay product factors until a two-mode gate needs them).
    # Default: the gates are applied to whatever state the previous step left (a generic dense state),
    # i.e. the steady-state cost of the kernels alone.
    value_kw = dict(shard_kw, lazy_vacuum=bool(args.from_vacuum))

    measured = {}

    def one_step(b):
        # config 3 ends with MeasureFock on every mode (BASELINE config 3): the measurement is part of the step,
        # and because it collapses the state every step starts from a fresh density matrix
        if args.from_vacuum or args.workload == "c3":
            b.reset(pure=args.workload != "c3")
        W.run_calls(b, calls)
        if args.workload == "c3":
            np.random.seed(7)
            measured["outcome"] = np.asarray(b.measure_fock(list(range(n_modes)))).reshape(-1).tolist()
        b.circuit._flush()

    be = B200FockBackend()
    be.begin_circuit(n_modes, cutoff_dim=D, fuse=fuse, **value_kw)
    for _ in range(args.warmup):
        one_step(be)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    handle.b200_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step(be)
    e1.record()
    barrier()
    launches = int(handle.b200_launch_count())
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=DEVICE)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = updates_per_step * args.steps / (ms_total * 1e-3)   # whole job: one state, all GPUs
    exchanges = getattr(be.circuit, "exchanges", 0)

    # ---- roofline leg: per-launch CUDA events on one more step ------------------------------
    prof = []
    be.circuit.profile = prof
    one_step(be)
    torch.cuda.synchronize()
    be.circuit.profile = None
    by_tag = {}
    for tag, nbytes, a, b in prof:
        if tag.startswith("exchange"):  # reported separately (NVLink, not HBM)
            continue
        t = a.elapsed_time(b) * 1e-3
        d = by_tag.setdefault(tag, [0, 0.0, 0])
        d[0] += nbytes
        d[1] += t
        d[2] += 1
    kernels = {"k_tile_pass": "tile", "k_apply_blocks": "gate", "k_apply_inner": "inner", "k_apply_diag": "diag"}
    groups = {}
    for kname, prefix in kernels.items():
        sel = [v for tag, v in by_tag.items() if tag.startswith(prefix)]
        if sel:
            groups[kname] = (sum(v[0] for v in sel), sum(v[1] for v in sel), sum(v[2] for v in sel))
    dom_kernel = max(groups, key=lambda k: groups[k][1])
    dom_bytes, dom_time, dom_n = groups[dom_kernel]
    peak, peak_src = measured_peaks()
    achieved = dom_bytes / dom_time / 1e9 if dom_time > 0 else 0.0
    roofline = {
        "kernel": dom_kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "share_of_step_time": dom_time / max(sum(g[1] for g in groups.values()), 1e-12),
        "frac": achieved / peak, "traffic": ncu_traffic(dom_kernel, local_elements), "peak_source": peak_src,
        "launches_per_step": dom_n, "avg_launch_ms": dom_time / max(dom_n, 1) * 1e3,
        "algorithmic_bytes_per_launch": dom_bytes / max(dom_n, 1),
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "by_pass": {tag: {"launches": v[2], "GBps": v[0] / v[1] / 1e9} for tag, v in sorted(by_tag.items())},
    }

    # ---- end-to-end leg through the plugin API with host buffers: `e2e` -----------------------
    # inputs: the step's gate-parameter table in pinned host memory, copied to the device and
    # consumed there by the gate-table generators; result: trace + a list of Fock probabilities
    # read back to the host.
    from strawberryfields_b200 import DeviceParams

    nb = args.batch if args.workload == "c4" else 1
    two_param = ("squeeze", "displacement", "beamsplitter", "mzgate", "two_mode_squeeze")

    def params_of(c):  # -> [2, nb]
        vals = [np.broadcast_to(np.asarray(x, dtype=np.float64), (nb,)) for x in c[1:]
                if not isinstance(x, (int, np.integer))]
        vals += [np.zeros(nb)] * (2 - len(vals))
        return np.stack(vals[:2])

    ptab = np.ascontiguousarray(np.stack([params_of(c) for c in calls]))  # [ncalls, 2, nb]
    pinned = torch.from_numpy(ptab).pin_memory()
    outcomes = [[0] * n_modes]
    for m in range(n_modes):
        for k in (1, 2):
            o = [0] * n_modes
            o[m] = k
            outcomes.append(o)

    # the plugin with its default options: every step is a whole program run from vacuum (lazy vacuum, gate
    # calls deferred until state() is asked for -- DESIGN 4.7)
    be2 = B200FockBackend()
    be2.begin_circuit(n_modes, cutoff_dim=D, fuse=fuse, **shard_kw)

    def e2e_step():
        dev_params = pinned.to(DEVICE, non_blocking=True)  # H2D of the step's inputs
        be2.reset(pure=args.workload != "c3")  # what LocalEngine.reset() does between runs (engine.py:413-417)
        for i, c in enumerate(calls):
            modes = [x for x in c[1:] if isinstance(x, (int, np.integer))]
            getattr(be2, c[0])(*([DeviceParams(dev_params[i])] + ([None] if c[0] in two_param else []) + modes))
        st = be2.state()
        return np.array([st.trace()] + [st.fock_prob(o) for o in outcomes])  # D2H reads

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device=DEVICE)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = updates_per_step * args.steps / (float(ms2.item()) * 1e-3)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ptab.nbytes),
           "d2h_bytes_per_step": int(8 * (len(outcomes) + 1) * nb),
           "api": "B200FockBackend.begin_circuit/gates/state().trace()/fock_prob()",
           "state_at_step_start": "vacuum (backend.reset() every step, plugin defaults: lazy vacuum)"}

    # ---- sharded arm: parity of the sharded GPU path + the same workload on ONE GPU -------------
    parity = single_ms = ten = None
    if sharded and args.workload == "c2" and not args.no_parity:
        parity, single_ms = sharded_parity(be2, calls, n_modes, D, rank, world, args.from_vacuum, args.exchange,
                                            n_probs=args.parity_probs)
        if world == 8 and n_modes == 9 and not args.no_ten_mode:
            be.circuit._bufs = be.circuit._buf = be.circuit._backing = be.circuit._peers = None
            be2.circuit._bufs = be2.circuit._buf = be2.circuit._backing = be2.circuit._peers = None
            torch.cuda.empty_cache()
            ten = ten_mode_block(world, rank, D, args.exchange)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = unmodified_reference_run(8, 10, ("beamsplitter",), 1) if args.workload == "c2" else None
        if res is not None:
            v, per_step, ngates = res
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "one BSgate of the config-2 circuit on the full-size state (8 modes, cutoff 10, 1e8 "
                             "amplitudes) with the unmodified reference FockBackend (oracle/_ref): %.1f s" % per_step,
                   "host_cores_available": os.cpu_count()}
        else:
            v, per_step, ngates = cpu_reference_run(6, 10, 1, 0)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "config-2 generator at 6 modes, cutoff 10 (1e6 amplitudes, %d gates), 1 step = %.1f s"
                             % (ngates, per_step),
                   "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic",
            "config": {
                "workload": "BASELINE config %s; cutoff %d, %.3g stored complex128 elements, %d gates per step"
                            % (wl_name, D, elements, len(calls)),
                "parallelism": ("state sharded on its leading axes over %d ranks, all-to-all axis exchange over "
                                "NVLink (see exchange.mode)" % world) if sharded else "single GPU",
                "l2": "state %.2f GB per GPU > 126 MB L2: every pass streams from HBM"
                      % (local_elements * 16 / 1e9),
                "passes_per_step": sum(v[2] for v in by_tag.values()),
                "gate_queue": args.fuse,
                "state_at_step_start": ("vacuum (reset inside the timed step, lazy-vacuum factors)" if args.from_vacuum
                                        else "vacuum density matrix (reset inside the timed step; the step ends with "
                                             "MeasureFock on all modes, seed 7)" if args.workload == "c3"
                                        else "dense state left by the previous step"),
            },
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "hbm_GBps_per_gpu_whole_step": sum(v[0] for v in by_tag.values()) * args.steps
                                           / (ms_total * 1e-3) / 1e9,
            "circuit_ms": ms_total / args.steps,
        }
        if sharded:
            ex = [(nb, a.elapsed_time(b) * 1e-3) for tag, nb, a, b in prof if tag == "exchange"]
            line["exchange"] = {
                "all_to_all_per_step": len(ex),
                "bytes_sent_per_gpu_per_exchange": int(16 * local_elements * (world - 1) // world),
                "GBps_per_gpu_per_direction_incl_pack_unpack":
                    (sum(nb for nb, _ in ex) / sum(t for _, t in ex) / 1e9) if ex else None,
                "exchanges_in_timed_region": int(exchanges),
            }
            if not getattr(be.circuit, "_p2p", False):
                line["exchange"]["mode"] = "pack + NCCL all_to_all + unpack"
            elif args.exchange == "push":
                line["exchange"]["mode"] = "peer-memory push kernel (NVLink P2P stores, no staging)"
            else:
                line["exchange"]["mode"] = "peer-memory pull kernel (NVLink P2P loads, no staging)"
            for part in ("p2p_pull", "p2p_push", "pack", "all_to_all", "unpack"):
                sel = [(nb, a.elapsed_time(b) * 1e-3) for tag, nb, a, b in prof if tag == "exchange/" + part]
                if sel:
                    line["exchange"][part] = {"ms": sum(t for _, t in sel) / len(sel) * 1e3,
                                              "GBps": sum(nb for nb, _ in sel) / sum(t for _, t in sel) / 1e9}
        if measured:
            line["measure_fock_outcome"] = measured["outcome"]   # tests/golden/ref_config3_full.npz holds the oracle's
        if parity is not None:
            line["parity"] = parity
            line["single_gpu_same_workload_ms"] = single_ms
            line["strong_efficiency"] = single_ms / (world * ms_total / args.steps)
        if ten is not None:
            line["ten_mode"] = ten
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--modes", type=int, default=0, help="default: 8 on one GPU, 9 sharded")
    ap.add_argument("--cutoff", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4"],
                    help="c2 (default; c5 when sharded over N > 1 GPUs), c3 = mixed state + loss, c4 = batched QNN layer")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "push", "nccl"],
                    help="multi-GPU axis exchange: peer-memory pull kernel (auto / p2p), peer-memory push kernel "
                         "(experimental), or pack + NCCL all-to-all + unpack")
    ap.add_argument("--fuse", default="fold", choices=["tile", "fold", "off"],
                    help="gate queue: diagonal / same-mode folding (default), + multi-gate tile passes, "
                         "or one pass per gate")
    ap.add_argument("--exchange-overlap", type=int, default=4,
                    help="sharded arm: gates behind an exchange that run part by part while the rest of the shard is "
                         "still crossing NVLink (0 = off)")
    ap.add_argument("--no-parity", action="store_true", help="sharded arm: skip the parity block")
    ap.add_argument("--parity-probs", type=int, default=1000, help="sharded arm: Fock probabilities compared")
    ap.add_argument("--no-ten-mode", action="store_true", help="8-GPU sharded arm: skip the 10-mode / 160 GB block")
    ap.add_argument("--from-vacuum", action="store_true",
                    help="the device-timed step starts from vacuum with the plugin's default lazy vacuum (reset + "
                         "gates + flush) instead of re-applying the gates to the dense state of the previous step")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "c1":
        c1_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
