"""cProfile of the setup() of the class-conditional baselines (second instance: library start-up excluded)."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402

rng = np.random.RandomState(11)
C, d, ntr, nte = 10, 512, 50_000, 10_000
means = rng.randn(C, d).astype(np.float32)
ytr = rng.randint(0, C, ntr)
train = (means[ytr] + rng.randn(ntr, d)).astype(np.float32)
valid = (means[rng.randint(0, C, nte)] + rng.randn(nte, d)).astype(np.float32)
W = (0.05 * rng.randn(C, d)).astype(np.float32)
b = rng.randn(C).astype(np.float32)
lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
kw = dict(valid_feats=valid, train_labels=ytr, train_logits=lg(train), valid_logits=lg(valid),
          final_linear_layer_params={"weight": W, "bias": b})
I = R.inference
for name, ctor in {"mahalanobis": lambda: I.Mahalanobis(flip_sign=False, num_classes=C), "vim": lambda: I.ViM(flip_sign=False),
                   "ddu": lambda: I.DDU(flip_sign=False, num_classes=C), "md": lambda: I.MDLatentSpace()}.items():
    ctor().setup(train, **kw)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    ctor().setup(train, **kw)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print("=====", name)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:32]))
