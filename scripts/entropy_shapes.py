import sys, torch, time
sys.path.insert(0, '/root/repo')
from runia_core_b200 import _ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
for n_mc, D, items in ((16, 512, 60000), (32, 512, 30000), (8, 512, 60000), (10, 512, 60000), (16, 510, 60000), (32, 1024, 10000)):
    z = (torch.randn(items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(items, n_mc, D, generator=g, device=dev)).reshape(-1, D).contiguous()
    for _ in range(2):
        _ops.mcd_entropy(z, n_mc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        _ops.mcd_entropy(z, n_mc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"n_mc={n_mc} D={D} items={items}: {ms:.3f} ms  {items/ms*1e3:.3e} items/s  {z.numel()*4/ms/1e6:.0f} GB/s")
