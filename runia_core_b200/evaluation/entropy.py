"""`get_dl_h_z` / `single_image_entropy_calculation` with the reference's signatures
(`runia_core/evaluation/entropy.py:20-93`), computed by one bandwidth-bound CUDA kernel
(`runia_mcd_entropy_f32`) instead of N*D Python-level estimator calls."""
from typing import Tuple, Union

import numpy as np
import torch
from torch import Tensor

from .. import _ops
from .._device import to_device, to_host

__all__ = ["get_dl_h_z", "single_image_entropy_calculation"]

MAX_MC_SAMPLES = 128  # entropy.cu kEntropyMaxN; the tuned kernels cover 2..32, the generic ones the rest


def single_image_entropy_calculation(sample: np.ndarray, neighbors: int) -> np.ndarray:
    """Per-dimension entropy of ONE item: sample [n_mc, D] -> [D] float64 (entropy.py:20-38)."""
    z = to_device(sample, torch.float32)
    assert z.dim() == 2
    _, h_z = _ops.mcd_entropy(z.contiguous(), z.shape[0], k=int(neighbors), want_joint=False)
    return to_host(h_z)[0]


def get_dl_h_z(dl_z_samples: Union[Tensor, np.ndarray], mcd_samples_nro: int = 32,
               parallel_run: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Joint and per-dimension Kozachenko-Leonenko entropies of MC-dropout latent samples.

    dl_z_samples: [N * mcd_samples_nro, D], item-major.  Returns (h_mvn [N, 1], h_z [N, D]),
    float64.  `parallel_run` (a multiprocessing pool upstream, entropy.py:85-91) is accepted and
    ignored: all items are reduced in one launch.  Like upstream, a Tensor whose row count is not
    a multiple of mcd_samples_nro yields a last item from the short chunk (entropy.py:56-58), an
    ndarray raises (np.split, entropy.py:60-62)."""
    n_mc = int(mcd_samples_nro)
    if n_mc > MAX_MC_SAMPLES or n_mc < 2:
        raise NotImplementedError(f"mcd_samples_nro={n_mc}: the CUDA estimator handles 2..{MAX_MC_SAMPLES} samples")
    is_tensor = isinstance(dl_z_samples, Tensor)
    rows = dl_z_samples.shape[0]
    if not is_tensor and rows % n_mc != 0:
        raise ValueError("array split does not result in an equal division")
    z = dl_z_samples  # host arrays are staged by the operator itself (pinned ring, copy / kernel overlap)
    if is_tensor and z.is_cuda:
        z = to_device(z, torch.float32)
    assert z.ndim == 2
    k = _ops.entropy_k(n_mc)
    n_full = rows // n_mc
    rem = rows - n_full * n_mc
    h_mvn, h_z = _ops.mcd_entropy(z[: n_full * n_mc], n_mc, k=k)
    h_mvn, h_z = to_host(h_mvn), to_host(h_z)
    if rem:
        if rem <= k:
            # fewer samples than neighbours: the estimator's k-th distance is infinite upstream
            t_mvn = np.full((1,), np.inf)
            t_z = np.full((1, z.shape[1]), np.inf)
        else:
            tm, tz = _ops.mcd_entropy(z[n_full * n_mc:], rem, k=k)
            t_mvn, t_z = to_host(tm), to_host(tz)
        h_mvn = np.concatenate([h_mvn, t_mvn])
        h_z = np.concatenate([h_z, t_z])
    return np.expand_dims(h_mvn, axis=1), h_z
