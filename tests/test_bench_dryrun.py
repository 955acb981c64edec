"""bench.py's control flow and JSON contract, dry-run on the CPU: the CUDA kernels are replaced by
the numpy double of the C ABI and the CUDA events by wall-clock stand-ins, so the NUMBERS mean
nothing -- what is checked is that every workload / option combination runs to the end and prints
one JSON line with the keys the driver reads (a Python error here would cost a round's benchmark)."""
import argparse
import json
import time

import pytest
import torch

from fake_lib import FakeLib


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


@pytest.fixture
def bench(monkeypatch):
    from strawberryfields_b200 import circuit, lib
    import bench as B

    monkeypatch.setattr(lib, "_lib", FakeLib())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    monkeypatch.setattr(B, "DEVICE", "cpu")
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    return B


def _args(**kw):
    base = dict(gpus=1, steps=2, warmup=1, impl="b200", modes=4, cutoff=3, no_cpu_baseline=True, workload="c2",
                batch=2, exchange="auto", fuse="fold", from_vacuum=False, no_parity=False, no_ten_mode=False,
                parity_probs=12, exchange_overlap=8)
    base.update(kw)
    return argparse.Namespace(**base)


CONTRACT = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"]


@pytest.mark.parametrize("kw", [
    {}, {"from_vacuum": True}, {"fuse": "tile"}, {"fuse": "off"},
    {"workload": "c3", "modes": 3}, {"workload": "c3", "modes": 3, "from_vacuum": True},
    {"workload": "c4", "modes": 3}, {"workload": "c4", "modes": 3, "from_vacuum": True},
], ids=lambda kw: "-".join("%s=%s" % i for i in kw.items()) or "default")
def test_b200_arm_prints_the_contract(kw, bench, capsys):
    bench.b200_arm(_args(**kw))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in CONTRACT:
        assert key in line, key
    assert line["metric"] == "fock_amp_gate_updates_per_s" and line["unit"] == "updates/s"
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert "workload" in line["config"]


def test_c1_arm(bench, capsys):
    bench.c1_arm(_args(workload="c1", steps=1, warmup=1))
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][0])
    assert line["golden_probabilities_ok"] and line["gpu_launches"] > 0


@pytest.mark.parametrize("exchange,mode", [("auto", ""), ("push", ""), ("p2p", "from_vacuum")])
def test_sharded_arm(exchange, mode):
    """the torchrun arm (BASELINE config 5 shape) on 2 gloo ranks: rank 0 prints the line, with the
    exchange section"""
    import os
    import socket
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(root, "tests", "bench_dryrun_worker.py"), "5", "4", exchange] + ([mode] if mode else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in CONTRACT:
        assert key in line, key
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    # the sharded arm proves its own correctness in the line: single-photon transfer, comparison with an
    # unsharded run of the same circuit on rank 0, seeded MeasureFock
    assert line["parity"]["max_abs_err"] <= 1e-12 and line["parity"]["measure_equal"]
    assert line["parity"]["vs_unsharded_run_on_rank0"]["fock_probs_compared"] == 12
    assert line["single_gpu_same_workload_ms"] > 0 and line["strong_efficiency"] > 0
    if not mode:  # from vacuum the replicated prefix may leave nothing to exchange at this size
        assert line["exchange"]["all_to_all_per_step"] >= 1
        assert ("p2p_" in " ".join(line["exchange"])) == (exchange != "auto")
    else:
        assert line["config"]["state_at_step_start"].startswith("vacuum")
