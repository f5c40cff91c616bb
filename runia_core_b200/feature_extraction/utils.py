"""`get_mean_or_fullmean_ls_sample` with the reference's signature
(`runia_core/feature_extraction/utils.py:70-92`): spatial mean of a convolutional activation map in one
coalesced pass on the device."""
import torch
from torch import Tensor

from .. import _lib
from .._device import stream_ptr, to_device

__all__ = ["get_mean_or_fullmean_ls_sample"]


def get_mean_or_fullmean_ls_sample(latent_sample: Tensor, method: str = "fullmean") -> Tensor:
    """[B, C, H, W] -> [B, C, H] ("mean": over W) or [B, C] ("fullmean": over W then H), squeezed like
    upstream (`torch.squeeze` drops every size-1 dimension)."""
    assert method in ("mean", "fullmean")
    assert latent_sample.dim() == 4, "activation map must be (batch, channels, height, width)"
    x = to_device(latent_sample, torch.float32)
    B, C, H, W = x.shape
    full = method == "fullmean"
    out = torch.empty((B, C) if full else (B, C, H), dtype=torch.float32, device=x.device)
    _lib.call("runia_spatial_mean_f32", x.data_ptr(), B * C, H, W, 1 if full else 0, out.data_ptr(), stream_ptr())
    out = torch.squeeze(out)
    return out if latent_sample.is_cuda else out.cpu()
