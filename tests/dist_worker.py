"""Worker of tests/test_sharding.py: run under torchrun with the gloo backend.  Each rank
drives a ShardedCircuit (host logic + numpy double of the C ABI) through a circuit and
checks the gathered ket against the single-process oracle."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    if os.environ.get("B200_DUMP_AFTER"):  # debugging aid: where is a rank stuck?
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["B200_DUMP_AFTER"]), exit=True)
    mode = sys.argv[1]          # "host" (gloo + numpy double) or "gpu" (nccl + CUDA kernels)
    n, D = int(sys.argv[2]), int(sys.argv[3])
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200 import workloads as W
    from strawberryfields_b200.backend import B200FockBackend

    if mode == "host":
        from fake_lib import FakeLib

        lib._lib = FakeLib()
        circuit._TEST_HOST_MODE = True
        dist.init_process_group("gloo")
    else:
        import torch

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    from oracle.fock_oracle import OracleBackend

    calls = W.config2_circuit(n, seed=5)
    extra = [("mzgate", 0.4, 1.1, 0, n - 1), ("two_mode_squeeze", 0.2, 0.3, n - 1, 1), ("kerr_interaction", 0.1, 0),
             ("cross_kerr_interaction", 0.3, 0, n - 1), ("beamsplitter", 0.7, 0.2, 1, 0)]
    be = B200FockBackend()
    exchange = sys.argv[4] if len(sys.argv) > 4 else "auto"
    flags = sys.argv[5:]
    if "window7" in flags:  # plan the queue in windows of 7 entries (very long programs do that at 256)
        from strawberryfields_b200 import sharding as _sh

        _sh._PLAN_WINDOW = 7
    lazy = "lazy" in flags
    # "mixed": the circuit starts as a density matrix; "loss": it starts pure and a LossChannel in the
    # middle turns the sharded ket into a sharded density matrix
    pure = "mixed" not in flags
    if "mixed" in flags or "loss" in flags:
        half = len(calls) // 2
        calls = calls[:half] + [("loss", 0.8, 0), ("loss", 0.6, n - 1)] + calls[half:] + [("loss", 0.9, 1)]
    be.begin_circuit(n, cutoff_dim=D, shard=True, exchange=exchange, lazy_vacuum=lazy, pure=pure)
    ob = OracleBackend()
    ob.begin_circuit(n, cutoff_dim=D, pure=pure)
    if "fock" in flags:
        # boson-sampling inputs: single photons in the first and the last mode.  The oracle prepares the
        # whole product ket at once (its single-mode preparations would mix the state)
        be.prepare_fock_state(1, 0)
        be.prepare_fock_state(1, n - 1)
        ket = np.zeros([D] * n, dtype=complex)
        ket[(1,) + (0,) * (n - 2) + (1,)] = 1.0
        ob.prepare_ket_state(ket, list(range(n)))
    W.run_calls(be, calls + extra)
    st = be.state()
    if "reset" in flags:
        # a state object handed out BEFORE reset() keeps its state (the reference's reset allocates a new
        # array; ADVICE round 1: the NCCL mode used to zero the buffer the snapshot shared)
        p_before = st.fock_prob([0] * n)
        tr_before = st.trace()
        be.reset()
        ok = bool(abs(st.fock_prob([0] * n) - p_before) < 1e-15 and abs(st.trace() - tr_before) < 1e-15
                  and abs(be.state().fock_prob([0] * n) - 1.0) < 1e-15 and p_before < 0.999)
        for bad in ({"pure": 1}, {"cutoff_dim": 0}, {"num_subsystems": 2.5}):
            try:
                be.circuit.reset(**bad)
                ok = False
            except ValueError:
                pass
        print(json.dumps({"rank": dist.get_rank(), "world": dist.get_world_size(), "ok": ok}))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    if "addmode" in flags:
        # add_mode on a sharded state: the new mode is a whole (local) axis in the vacuum; gates on it work
        W.run_calls(ob, calls + extra)
        be.add_mode(1)
        ob.add_mode(1)
        be.beamsplitter(0.4, 0.3, 1, n)
        ob.beamsplitter(0.4, 0.3, 1, n)
        be.displacement(0.2, 0.1, n)
        ob.displacement(0.2, 0.1, n)
        st2, ost2 = be.state(), ob.state()
        ok = bool(st2.num_modes == n + 1 and np.abs(st2.data - ost2.data).max() < 1e-12
                  and abs(st2.trace() - ost2.trace()) < 1e-12)
        print(json.dumps({"rank": dist.get_rank(), "world": dist.get_world_size(), "ok": ok}))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    if "ckpt" in flags:
        # per-rank checkpoint: save the shards, wipe the circuit, load them back -- same state
        import tempfile

        from strawberryfields_b200 import sharding

        want = st.data.copy()
        path = [tempfile.mkdtemp() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(path, src=0)
        be.circuit.save_shard(path[0])
        be.reset()
        be.circuit.load_shard(path[0])
        got = be.state().data
        ok = bool(np.abs(got - want).max() == 0 and abs(be.state().trace() - st.trace()) < 1e-15)
        be.rotation(0.3, 0)   # the restored circuit keeps working
        ok = ok and abs(be.state().trace() - st.trace()) < 1e-12
        print(json.dumps({"rank": dist.get_rank(), "world": dist.get_world_size(), "ok": ok}))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    from strawberryfields_b200 import sharding

    first_layout = list(next(iter(sharding._PLANS.values()))[0])  # layout the planner chose for |0..0>
    W.run_calls(ob, calls + extra)
    ost = ob.state()
    want = ost.data
    assert st.is_pure == ost.is_pure
    err = float(np.abs(st.data - want).max())
    tr_err = abs(st.trace() - ost.trace())
    idx = [1] + [0] * (n - 2) + [2 % D]
    fp_err = abs(st.fock_prob(idx) - ost.fock_prob(idx))
    # reductions: all_fock_probs (local reduce + all-reduce), then a seeded MeasureFock on a sharded and
    # a local mode: same outcome and same post-measurement state as the oracle
    probs_before = np.array(ost.all_fock_probs(), copy=True)
    probs_err = float(np.abs(st.all_fock_probs() - probs_before).max())
    # reduced density matrices (kept axes are exchanged in if they are sharded) and what builds on them
    probs_err = max(probs_err, float(np.abs(st.reduced_dm([0, n - 1]) - ost.reduced_dm([0, n - 1])).max()),
                    float(np.abs(np.array(st.mean_photon(1)) - np.array(ost.mean_photon(1))).max()),
                    float(np.abs(np.array(st.quad_expectation(0, 0.3)) - np.array(ost.quad_expectation(0, 0.3))).max()),
                    abs(st.fidelity_coherent([0.2 * np.exp(0.5j * m) for m in range(n)])
                        - ost.fidelity_coherent([0.2 * np.exp(0.5j * m) for m in range(n)])),
                    abs(st.parity_expectation([0, n - 1]) - ost.parity_expectation([0, n - 1])))
    # state(modes=[...]): the reduced state in the requested (here: unsorted) mode order, on every rank
    probs_err = max(probs_err, float(np.abs(be.state(modes=[n - 1, 1]).dm() - ob.state(modes=[n - 1, 1]).data).max()))
    np.random.seed(5)
    got_out = be.measure_fock([0, n - 1])
    np.random.seed(5)
    want_out = ob.measure_fock([0, n - 1])
    post_err = float(np.abs(be.state().data - ob.state().data).max())
    # the state object taken BEFORE the measurement still holds the pre-measurement state
    post_err = max(post_err, float(np.abs(st.all_fock_probs() - probs_before).max()))
    # homodyne on a mode that is entangled with the rest: same sample (rank 0's draw) and same conditional
    # state as the UNSHARDED backend on the same double (the oracle has no homodyne; the single-process path
    # is pinned by the reference's own homodyne tests, profiles/r01_reference_suite.md)
    if mode == "host" and "fock" not in flags:
        ref = B200FockBackend()
        ref.begin_circuit(n, cutoff_dim=D, pure=pure)
        W.run_calls(ref, calls + extra)
        np.random.seed(5)
        ref.measure_fock([0, n - 1])
        for b in (be, ref):
            b.beamsplitter(0.6, 0.4, 0, 1)
        np.random.seed(9)
        got_x = be.measure_homodyne(0.3, 1, num_bins=2000)
        np.random.seed(9)
        want_x = ref.measure_homodyne(0.3, 1, num_bins=2000)
        post_err = max(post_err, float(np.abs(np.asarray(got_x) - np.asarray(want_x)).max()),
                       float(np.abs(be.state().data - ref.state().data).max()))
    ok = bool(err < 1e-12 and tr_err < 1e-12 and fp_err < 1e-12 and probs_err < 1e-12
              and np.array_equal(got_out, want_out) and post_err < 1e-12)
    print(json.dumps({"rank": dist.get_rank(), "world": dist.get_world_size(), "err": err,
                      "trace_err": float(tr_err), "fock_prob_err": float(fp_err),
                      "exchanges": int(be.circuit.exchanges), "p2p": bool(be.circuit._p2p),
                      "free_layout": bool(first_layout != list(range(n))), "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
