"""One launch each of entropy16_kernel (60k x 16 x 512) and entropy32_kernel (30k x 32 x 512) after a warm-up pair:
the shape ncu captures for profiles/ (ncu -k regex:entropy --launch-skip 2 -c 2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

z16 = torch.randn(60000 * 16, 512, device="cuda")
z32 = torch.randn(30000 * 32, 512, device="cuda")
for _ in range(2):
    _ops.mcd_entropy(z16, 16)
    _ops.mcd_entropy(z32, 32)
    torch.cuda.synchronize()
