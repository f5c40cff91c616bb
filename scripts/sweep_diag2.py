"""The configs[1] sweep of bench.py with the host-side phases of every call timed (upload enqueue, each C-ABI call,
download): which phase carries the late calls."""
import collections
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import runia_core_b200 as R  # noqa: E402
from runia_core_b200 import _device, _lib  # noqa: E402
from runia_core_b200.inference import postprocessors as PP  # noqa: E402

T = collections.defaultdict(list)


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[name].append((time.perf_counter() - t0) * 1e3)
        return r
    return w


_device.to_device = timed("to_device", _device.to_device)
PP.to_host = timed("to_host", PP.to_host)
_call = _lib.call
_lib.call = lambda name, *a: timed("call:" + name, _call)(name, *a)


def cpu_times():
    with open("/proc/stat") as f:
        v = [int(x) for x in f.readline().split()[1:]]
    return v  # user nice system idle iowait irq softirq steal


def ctxt():
    out = {}
    with open("/proc/self/status") as f:
        for line in f:
            if "ctxt_switches" in line:
                k, v = line.split(":")
                out[k] = int(v)
    return out


def brief(s):
    return {k: [round(v["ms"], 3), round(v["ms_mean"], 3), round(v["ms_max"], 2)] for k, v in s.items() if isinstance(v, dict)}


c0, x0 = cpu_times(), ctxt()
t0 = time.time()
out = {"sweep": brief(bench._extra_sweep_config2(R))}
c1, x1 = cpu_times(), ctxt()
dt = [b - a for a, b in zip(c0, c1)]
out["host"] = {"wall_s": round(time.time() - t0, 2), "jiffies": dict(zip("user nice system idle iowait irq softirq steal".split(), dt[:8])),
               "ctxt": {k: x1[k] - x0[k] for k in x0}}
out["phases_ms"] = {k: {"n": len(v), "median": round(float(np.median(v)), 3), "p90": round(float(np.percentile(v, 90)), 3),
                        "max": round(float(np.max(v)), 2)} for k, v in T.items() if len(v) >= 20}
print(json.dumps(out))
