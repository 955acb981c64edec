"""Sharded states: world_size-2 and -4 runs on the CPU with the gloo backend,
driving the host logic (rank/axis geometry, exchange planning, pack / all-to-all / unpack)
against the numpy double of the C ABI; the gathered ket must equal the oracle's to 1e-12.
The same worker runs on GPUs with NCCL (``-m gpu``, needs >= 2 devices)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(mode, world, n, D, exchange="auto", timeout=600, lazy=False, flags=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py"), mode, str(n), str(D), exchange]
    cmd += (["lazy"] if lazy else []) + list(flags)
    # start_new_session: on a timeout the whole torchrun process group is killed, nothing is left behind
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, 9)
        proc.communicate()
        raise AssertionError("sharded worker did not finish within %d s" % timeout)
    res = subprocess.CompletedProcess(cmd, proc.returncode, out, err)
    import re

    lines = [json.loads(m) for m in re.findall(r"\{[^{}]*\}", res.stdout)]  # ranks may share a line
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert len(lines) == world and all(l["ok"] for l in lines), lines
    return lines


def test_factor_world():
    from strawberryfields_b200.sharding import factor_world

    assert factor_world(2, 10) == [2]
    assert factor_world(4, 10) == [2, 2]
    assert factor_world(8, 10) == [2, 2, 2]
    assert factor_world(6, 6) == [6]
    assert factor_world(4, 6) == [2, 2]
    with pytest.raises(ValueError):
        factor_world(2, 7)


@pytest.mark.parametrize("world,n,D,exchange", [(2, 4, 4, "auto"), (3, 4, 6, "auto"), (4, 5, 6, "p2p"),
                                                (8, 8, 2, "p2p"), (4, 5, 6, "push")])
def test_sharded_circuit_matches_oracle_gloo(world, n, D, exchange):
    """exchange="auto" on the CPU is pack -> all_to_all_single -> unpack; "p2p" runs the peer-memory pull
    path (one strided gather per source rank, ping-pong buffers) with POSIX shared memory standing in
    for the NVLink-mapped peer buffers."""
    lines = _run("host", world, n, D, exchange)
    assert lines[0]["exchanges"] >= 1  # the circuit touches sharded modes: at least one all-to-all
    assert all(l["p2p"] == (exchange in ("p2p", "push")) for l in lines)
    if world == 8:  # three sharded axes: the planner starts the vacuum from a layout of its own choice
        assert lines[0]["free_layout"]


@pytest.mark.parametrize("world,n,D,exchange", [(8, 8, 2, "auto")])
def test_sharded_lazy_vacuum_matches_oracle_gloo(world, n, D, exchange):
    """lazy_vacuum=True on a sharded circuit: the prefix of the program that fits one rank runs replicated
    as a small lazy-vacuum circuit, then every rank writes its shard once; same ket, same
    measurement outcome (asserted by the worker on every rank)."""
    lines = _run("host", world, n, D, exchange, lazy=True)
    assert all(l["p2p"] == (exchange == "p2p") for l in lines)


def test_sharded_long_program_is_planned_in_windows_gloo():
    """Queues longer than the planning window (256 entries; 7 here) are planned window by window, each
    window starting from the layout the previous one ended in."""
    _run("host", 4, 5, 6, "p2p", flags=["window7"])


def test_sharded_fock_inputs_gloo():
    """Single-photon inputs (prepare_fock_state on untouched modes of a sharded ket) are queued as rank-one
    single-mode operators; with lazy_vacuum they are product factors of the replicated prefix."""
    _run("host", 4, 5, 6, "p2p", flags=["fock"])
    _run("host", 2, 4, 4, "auto", flags=["lazy", "fock"])


@pytest.mark.parametrize("world,n,D,exchange,flag", [(2, 3, 4, "auto", "loss"), (4, 4, 6, "p2p", "loss"),
                                                     (8, 5, 2, "auto", "mixed")])
def test_sharded_density_matrices_match_oracle_gloo(world, n, D, exchange, flag):
    """Sharded MIXED states: "mixed" starts as a density matrix (2n tensor axes, the leading ones sharded),
    "loss" starts as a sharded ket that LossChannels turn into a sharded density matrix (all-gather of the
    small ket, local outer product).  Gates become a ket-axis and a bra-axis entry of the exchange plan;
    trace / probabilities / MeasureFock walk the ket = bra diagonal across the ranks.  Same density
    matrix, reductions and measurement outcome as the oracle (asserted by the worker on every rank)."""
    _run("host", world, n, D, exchange, flags=[flag])


def test_sharded_add_mode_gloo():
    """add_mode on a sharded ket (circuit.py:373-382): no communication, the shard grows by a whole axis"""
    _run("host", 2, 3, 4, "p2p", flags=["addmode"])


def test_sharded_checkpoint_round_trip_gloo():
    """ShardedCircuit.save_shard / load_shard (on-disk checkpoint of a sharded ket)"""
    _run("host", 4, 5, 6, "p2p", flags=["ckpt"])


@pytest.mark.parametrize("exchange", ["auto"])   # the NCCL-mode snapshot shares the buffer; p2p snapshots clone
def test_state_object_survives_reset_gloo(exchange):
    """eng.run() -> eng.reset() -> result.state: the snapshot keeps its buffer in every exchange mode, and
    ShardedCircuit.reset validates its arguments like the single-GPU reset."""
    _run("host", 2, 3, 4, exchange, flags=["reset"])


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_sharded_circuit_matches_oracle_on_gpus(exchange):
    """NCCL all-to-all exchange and the peer-memory pull kernel give the oracle's ket."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    # world_size 2 is the configuration verified on 2 x B200 in round 1 (both exchange modes, incl. the
    # sharded MeasureFock).  A 4-rank run of this worker (5 modes, g = 2) exposed an exchange livelock in
    # the scheduler (too few evictable axes); it is fixed and covered on gloo by the (4, 5, 6) case above,
    # but the round's GPU budget ended before a 4-GPU re-run.  bench.py at 4 / 8 ranks was not affected.
    lines = _run("gpu", 2, 5, 6, exchange, timeout=150)
    assert all(l["p2p"] == (exchange == "p2p") for l in lines)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange,lazy,flags", [("push", False, ()), ("p2p", True, ()), ("p2p", False, ("loss",)),
                                                 ("p2p", False, ("fock",)), ("p2p", False, ("addmode",)),
                                                 ("p2p", False, ("ckpt",))])
def test_sharded_paths_written_after_round1_on_gpus(exchange, lazy, flags):
    """The push form of the peer-memory exchange, the sharded lazy vacuum, sharded density matrices, Fock inputs,
    add_mode and per-rank checkpoints on GPUs (green on 2 x B200 in round 2, profiles/r02_pytest_sharding_gpu_*.log)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    lines = _run("gpu", 2, 4 if flags else 5, 6, exchange, timeout=150, lazy=lazy, flags=flags)
    assert all(l.get("p2p", True) for l in lines)   # the add_mode / checkpoint workers report "ok" only
