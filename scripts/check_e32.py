"""GPU check of the four-warps-per-item n_mc = 32 entropy kernel against the one-warp-per-item kernel it replaces
(RUNIA_B200_E32_OFF=1 routes to the old one), plus timings of both and of the n_mc = 16 kernel."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
g = torch.Generator(device="cuda").manual_seed(5)
for n_items, D in [(3000, 512), (700, 100), (1201, 1024), (5, 4), (297, 132)]:
    z = torch.randn(n_items, 1, D, generator=g, device="cuda") + 0.1 * torch.randn(n_items, 32, D, generator=g, device="cuda")
    z[torch.rand(n_items, 32, D, generator=g, device="cuda") < 0.3] = 0.0
    z = z.reshape(n_items * 32, D).contiguous()
    os.environ.pop("RUNIA_B200_E32_OFF", None)
    hm1, hz1 = _ops.mcd_entropy(z, 32)
    hm1, hz1 = hm1.clone(), hz1.clone()
    os.environ["RUNIA_B200_E32_OFF"] = "1"
    hm0, hz0 = _ops.mcd_entropy(z, 32)
    os.environ.pop("RUNIA_B200_E32_OFF", None)
    torch.cuda.synchronize()
    dz = (hz1 - hz0).abs().max().item()
    dm = ((hm1 - hm0).abs() / hm0.abs().clamp_min(1e-9)).max().item()
    out[f"diff_{n_items}x{D}"] = {"hz_max_abs": dz, "hmvn_max_rel": dm}
    print(n_items, D, dz, dm, flush=True)

n_items, D = 30000, 512
z = torch.randn(n_items * 32, D, device="cuda")
alg = n_items * (32 * D * 4 + D * 8 + 8)
ms = timed(lambda: _ops.mcd_entropy(z, 32))
os.environ["RUNIA_B200_E32_OFF"] = "1"
ms_old = timed(lambda: _ops.mcd_entropy(z, 32))
os.environ.pop("RUNIA_B200_E32_OFF", None)
out["n32"] = {"ms": ms, "hbm_frac": alg / ms / 1e6 / 6553.0, "ms_one_warp_per_item": ms_old}
del z
n_items = 40000
z = torch.randn(n_items * 24, D, device="cuda")
ms = timed(lambda: _ops.mcd_entropy(z, 24))
os.environ["RUNIA_B200_E32_OFF"] = "1"
ms_old = timed(lambda: _ops.mcd_entropy(z, 24))
os.environ.pop("RUNIA_B200_E32_OFF", None)
out["n24"] = {"ms": ms, "ms_one_warp_per_item": ms_old}
del z
n_items = 60000
z = torch.randn(n_items * 16, D, device="cuda")
alg = n_items * (16 * D * 4 + D * 8 + 8)
ms = timed(lambda: _ops.mcd_entropy(z, 16))
out["n16"] = {"ms": ms, "hbm_frac": alg / ms / 1e6 / 6553.0}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check_e32.json", "w"), indent=1)
