"""entropy kernel timing: n_mc = 16 (configs[0] shape, 60k items x 512; configs[2] width 1024) and n_mc = 32."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    out = {}
    for n_items, n_mc, D in ((60_000, 16, 512), (30_000, 16, 1024), (30_000, 32, 512), (60_000, 8, 512)):
        z = torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev)
        z = z.reshape(n_items * n_mc, D).contiguous()
        for _ in range(3):
            _ops.mcd_entropy(z, n_mc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _ops.mcd_entropy(z, n_mc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        alg = n_items * (n_mc * D * 4 + D * 8 + 8)
        out[f"n{n_mc}_D{D}"] = {"ms": round(ms, 4), "items_per_s": round(n_items / ms * 1e3), "GBps": round(alg / ms / 1e6, 1),
                                "frac_6553": round(alg / ms / 1e6 / 6553, 4)}
        del z
    print(json.dumps(out))


if __name__ == "__main__":
    main()
