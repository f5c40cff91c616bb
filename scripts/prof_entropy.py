"""One entropy16 launch at the configs[0] shape for ncu --set full --import-source on."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
n_items, n_mc, D = 60_000, int(os.environ.get("N_MC", "16")), 512
z = torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev)
z = z.reshape(n_items * n_mc, D).contiguous()
for _ in range(3):
    _ops.mcd_entropy(z, n_mc)
torch.cuda.synchronize()
