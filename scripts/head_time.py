"""ReAct / DICE head (2M x 512, C = 10): the narrow tcgen05 kernel against the FP32 SIMT kernel at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
X = torch.relu(torch.randn(2_000_000, 512, generator=g, device=dev))
W = 0.05 * torch.randn(10, 512, generator=g, device=dev)
b = torch.randn(10, generator=g, device=dev)
planes = _ops.linear_planes(W)


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


t_tc = timed(lambda: _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes))
o_tc = _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes).clone()
keep = _ops.LINEAR_TC_MIN_ROWS
_ops.LINEAR_TC_MIN_ROWS = 1 << 62
t_simt = timed(lambda: _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes))
o_simt = _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes).clone()
_ops.LINEAR_TC_MIN_ROWS = keep
ref = torch.logsumexp(torch.clamp(X[:4096].double(), max=1.0) @ W.double().T + b.double(), dim=1)
print({"tc_ms": t_tc, "simt_ms": t_simt, "tc_err": (o_tc[:4096].double() - ref).abs().max().item(),
       "simt_err": (o_simt[:4096].double() - ref).abs().max().item(), "hbm_frac_tc": 4.104e9 / t_tc / 1e6 / 6553,
       "hbm_frac_simt": 4.104e9 / t_simt / 1e6 / 6553})
