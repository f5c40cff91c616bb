"""Host -> device copy rates of one GPU box: pageable `tensor.to(device)` (driver staging), the native staging
engine (`runia_stage_h2d`, csrc/stage.cu) and pinned memory, at the sizes of a 10k x 512 call and a 1M x 256 call.
RUNIA_B200_STAGE_THREADS selects the engine's thread count (read once per process)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _device  # noqa: E402


def rate(fn, nbytes, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"ms": round(dt * 1e3, 3), "GBps": round(nbytes / dt / 1e9, 1)}


def main():
    dev = torch.device("cuda", 0)
    out = {"cpu_count": os.cpu_count(), "stage_threads_env": os.environ.get("RUNIA_B200_STAGE_THREADS")}
    for name, shape, reps in (("20MB", (10_000, 512), 100), ("1GB", (1 << 20, 256), 5)):
        a = np.random.rand(*shape).astype(np.float32)
        pinned = torch.from_numpy(a).pin_memory()
        dst = torch.empty(shape, dtype=torch.float32, device=dev)
        nb = a.nbytes
        out[name] = {
            "pageable_to": rate(lambda: torch.from_numpy(a).to(dev), nb, reps),
            "engine": rate(lambda: _device.HostPipe.get(dev).upload(a, dst), nb, reps),
            "to_device": rate(lambda: _device.to_device(a), nb, reps),
            "pinned_copy": rate(lambda: dst.copy_(pinned, non_blocking=True), nb, reps),
        }
        assert torch.equal(_device.to_device(a).cpu(), torch.from_numpy(a))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
