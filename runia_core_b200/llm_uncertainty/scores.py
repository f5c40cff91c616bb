"""`eigen_score` with the reference's signature (`runia_core/llm_uncertainty/scores.py:49-66`,
`utils.py:102-117`), computed by one CUDA kernel from the n x n Gram matrix of the centred samples
instead of a float64 SVD of the d x d covariance.  The other scores of that file (entropies,
perplexity, RAUQ) are out of scope."""
from typing import Tuple

import torch

from .. import _lib
from .._device import device, stream_ptr, to_device

__all__ = ["eigen_score"]


def _construct_embedding_matrix(hidden_states: Tuple[torch.Tensor, ...], token_index: int = -1,
                                layer_index: int = 15) -> torch.Tensor:
    """utils.py:102-117: embeddings of the chosen token / layer, one row per sampled generation."""
    return hidden_states[token_index][layer_index].squeeze()


def eigen_score(hidden_states: Tuple[torch.Tensor, ...], alpha: float = 1e-3) -> float:
    """Mean log singular value of cov(E^T) + alpha I (Chen et al. 2024)."""
    e = to_device(_construct_embedding_matrix(hidden_states), torch.float32)
    assert e.dim() == 2, "embedding matrix must be (num_samples, hidden_size)"
    n, d = e.shape
    out = torch.empty((1,), dtype=torch.float64, device=device())
    _lib.call("runia_eigen_score_f32", e.data_ptr(), n, d, float(alpha), out.data_ptr(), stream_ptr())
    return float(out.item())
