"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference hot-path modules from
/root/reference so that golden vectors can be generated in the build container.

Nothing in the product (`runia_core_b200/`), in `-m gpu` tests, in `smoke()` or in
`bench.py` may import this file: `/root/reference` does not exist on the GPU box.
It is used by `oracle/gen_golden.py` (the fixtures under `tests/golden/` are what the tests consume)
(which skips itself when `/root/reference` is absent).

How it works: the reference's package `__init__`
files pull mlflow / pacmap / dropblock / lightning, none of which are installed and
none of which are on the scoring path.  We register empty package modules whose
`__path__` points into the reference tree, so `importlib` loads the real
`postprocessors.py`, `funcs.py`, `abstract_classes.py`, `entropy.py`, `baselines.py`
byte-for-byte, and provide tiny stand-ins for the four absent third-party modules:

* `omegaconf.DictConfig`   - attribute-access dict (type hints + cfg.k_neighbors only)
* `faiss.IndexFlatL2`      - exact squared-L2 brute force, float32, FLT_MAX / -1 padding
* `dropblock.DropBlock2D`  - identity (only needed for an import chain)
* `entropy_estimators.continuous.get_h` - Kozachenko-Leonenko estimator
  (entropy-estimators==0.0.1, requirements.txt:3): published algorithm restated.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "runia_core"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class _IndexFlatL2:
    """faiss.IndexFlatL2 semantics: exact squared L2 in float32, ascending, pads with
    FLT_MAX / -1 when k > ntotal."""

    def __init__(self, d):
        self.d = d
        self.xb = np.zeros((0, d), np.float32)
        self.ntotal = 0

    def add(self, x):
        x = np.ascontiguousarray(x, np.float32)
        self.xb = np.vstack([self.xb, x])
        self.ntotal = len(self.xb)

    def search(self, q, k):
        q = np.ascontiguousarray(q, np.float32)
        D = np.empty((len(q), k), np.float32)
        I = np.empty((len(q), k), np.int64)
        for r in range(len(q)):
            diff = self.xb - q[r][None, :]
            d2 = np.einsum("ij,ij->i", diff, diff).astype(np.float32)
            order = np.argsort(d2, kind="stable")[:k]
            n = len(order)
            D[r, :n] = d2[order]
            I[r, :n] = order
            D[r, n:] = np.finfo(np.float32).max
            I[r, n:] = -1
        return D, I


def _get_h(x, k=1, norm="max", min_dist=0.0, workers=1):
    from scipy.spatial import cKDTree
    from scipy.special import digamma

    x = np.asarray(x, np.float64)
    if x.ndim == 1:
        x = x[:, None]
    n, d = x.shape
    assert norm == "max"
    r = cKDTree(x).query(x, k + 1, eps=0, p=np.inf)[0][:, -1]
    r[r < min_dist] = min_dist
    return -digamma(k) + digamma(n) + (d / float(n)) * np.sum(np.log(2 * r))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference modules: .pp (postprocessors), .funcs,
    .abstract, .entropy, .baselines.  Occupies `runia_core*` in sys.modules, so call it
    only from a process that does not also need another `runia_core`."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    import torch

    _mod("runia_core").__path__ = [REFERENCE_ROOT + "/runia_core"]
    for sub in ("inference", "evaluation", "feature_extraction", "llm_uncertainty"):
        _mod("runia_core." + sub).__path__ = [REFERENCE_ROOT + "/runia_core/" + sub]
    if "omegaconf" not in sys.modules:
        _mod("omegaconf", DictConfig=_DictConfig,
             OmegaConf=types.SimpleNamespace(create=_DictConfig))
    if "faiss" not in sys.modules:
        _mod("faiss", IndexFlatL2=_IndexFlatL2)

    class DropBlock2D(torch.nn.Module):
        def __init__(self, drop_prob=0.0, block_size=1):
            super().__init__()

        def forward(self, x):
            return x

    if "dropblock" not in sys.modules:
        _mod("dropblock", DropBlock2D=DropBlock2D)
    if "entropy_estimators" not in sys.modules:
        cont = _mod("entropy_estimators.continuous", get_h=_get_h)
        _mod("entropy_estimators", continuous=cont)
    try:
        import tqdm.contrib.concurrent  # noqa: F401
    except Exception:  # pragma: no cover
        def process_map(fn, *iterables, **kw):
            return [fn(*a) for a in zip(*iterables)]
        _mod("tqdm.contrib.concurrent", process_map=process_map)

    _loaded["pp"] = importlib.import_module("runia_core.inference.postprocessors")
    _loaded["funcs"] = importlib.import_module("runia_core.inference.funcs")
    _loaded["abstract"] = importlib.import_module("runia_core.inference.abstract_classes")
    _loaded["entropy"] = importlib.import_module("runia_core.evaluation.entropy")
    try:
        _loaded["baselines"] = importlib.import_module("runia_core.evaluation.baselines")
    except Exception as e:  # baselines imports tqdm etc.; not required for goldens
        _loaded["baselines"] = None
        _loaded["baselines_error"] = repr(e)
    # dimensionality_reduction.py imports pacmap / matplotlib at module level (plots only)
    for name in ("pacmap", "matplotlib", "matplotlib.pyplot"):
        try:
            importlib.import_module(name)
        except Exception:
            # names that appear in the reference's type annotations
            _mod(name, scatter=None, Figure=None, Axes=None, PaCMAP=None)
    _loaded["dimred"] = importlib.import_module("runia_core.dimensionality_reduction")
    _loaded["DictConfig"] = _DictConfig
    return types.SimpleNamespace(**_loaded)
