#!/usr/bin/env python
"""Dry run of bench.py's GPU arm on a machine without a GPU.

    python tools/bench_dryrun.py [bench.py flags]

Executes bench.run_ours() from the first line to the printed JSON line with the DEVICE replaced by stand-ins:
torch.cuda (streams, events, pinned memory) and the library handle classes are mocks that compute nothing, every
"device" tensor is a CPU tensor, every event pair reports 15 ms.  What it checks is the host-side flow of the
benchmark -- input generation, sharding, the order of the measurement sections, the CPU checker leg and the
assembly of the JSON line with all the keys the driver reads -- so that an edit to bench.py cannot break the
round-end run for a reason unrelated to the device.  The numbers it prints are meaningless and labelled so
("data": "DRY RUN"); nothing here is a measurement and bench.py itself never imports this file.
"""
import contextlib
import importlib
import importlib.util
import io
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "jackal-navigation_b200"
EVENT_MS = 15.0


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return EVENT_MS


class _Handle:
    """Stands in for jn.Elas / jn.ObstacleScan / jn.Calibration: accepts every call bench.py makes."""
    _h = None
    calls = []

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def f(*a, **k):
            _Handle.calls.append(name)
            return 0
        return f


class _Lib:
    def __getattr__(self, name):
        if name == "jn_elas_profile_read":
            def read(h, buf):
                for i in range(7):
                    buf[i] = 1.0 + i
                return 0
            return _Fn(read)
        return _Fn(lambda *a: 0)


class _Fn:
    def __init__(self, f):
        self.f = f
        self.argtypes = None

    def __call__(self, *a):
        return self.f(*a)


def install_mocks():
    import torch
    real_device = torch.device
    cpu = real_device("cpu")
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.device = lambda *a, **k: cpu
    # one process stands in for a rank of an N > 1 launch (RANK / WORLD_SIZE / LOCAL_RANK in the environment):
    # the collectives become no-ops, so the max over ranks is this rank's own time
    import torch.distributed as dist
    dist.init_process_group = lambda *a, **k: None
    dist.barrier = lambda *a, **k: None
    dist.all_reduce = lambda *a, **k: None
    dist.destroy_process_group = lambda *a, **k: None
    synth = importlib.import_module(PKG + ".synth")
    sharding = importlib.import_module(PKG + ".sharding")
    fake = types.ModuleType(PKG)
    fake.ROBOTICS, fake.MIDDLEBURY = 0, 1
    fake.parameters = lambda *a, **k: dict(k)
    fake.Elas = _Handle
    fake.ObstacleScan = _Handle
    fake.Calibration = _Handle
    fake.lib = lambda: _Lib()
    count = [0]

    def launch_count():
        count[0] += 500
        return count[0]
    fake.launch_count = launch_count
    sys.modules[PKG] = fake
    sys.modules[PKG + ".synth"] = synth
    sys.modules[PKG + ".sharding"] = sharding


def load_bench():
    spec = importlib.util.spec_from_file_location("jn_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["jn_bench"] = m          # the worker pools pickle bench.py's functions by module name
    spec.loader.exec_module(m)
    return m


def dry_run(argv):
    """Returns the JSON line bench.py's GPU arm prints for `argv` (a list of bench.py flags)."""
    install_mocks()
    b = load_bench()
    old = sys.argv
    sys.argv = ["bench.py"] + list(argv)
    try:
        a = b.parse()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            b.run_ours(a)
    finally:
        sys.argv = old
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    if int(os.environ.get("RANK", "0")) != 0:
        assert not lines, "only rank 0 prints"
        return None
    assert len(lines) == 1, "bench.py must print exactly one JSON line, got %d" % len(lines)
    line = json.loads(lines[0])
    line["data"] = "DRY RUN"
    return line


if __name__ == "__main__":
    _line = dry_run(sys.argv[1:] or ["--width", "320", "--height", "240", "--disp-max", "64", "--batch", "4",
                                              "--steps", "2", "--warmup", "1"])
    if _line is not None:
        print(json.dumps(_line))
