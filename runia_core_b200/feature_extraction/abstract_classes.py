"""`MCSamplerModule` of the reference's `runia_core/feature_extraction/abstract_classes.py:32-101`: n_mc
DropBlock2D layers applied to one hooked activation map, each reduced to its H x W mean -- the producer of the
`[n_mc, C]` sample rows that `get_dl_h_z` consumes.

The reference instantiates `dropblock.DropBlock2D` (third party, dropblock==0.3.0) n_mc times and runs n_mc
full passes over the map.  Here the Bernoulli seeds are drawn exactly as those layers draw them (one
`torch.rand(B, H, W)` per layer, in layer order, from torch's default CPU generator, compared with
`drop_prob / block_size**2`), so the random stream is the reference's, and everything after the draw -- block
dilation, masking, renormalisation, spatial mean, for all n_mc samples -- is one pass over the map on the GPU
(`runia_mc_dropblock_mean_f32`).
"""
import torch

from .. import _ops
from .utils import get_mean_or_fullmean_ls_sample

__all__ = ["MCSamplerModule"]


class MCSamplerModule(torch.nn.Module):
    def __init__(self, mc_samples: int, block_size: int, drop_prob: float, layer_type: str = "Conv"):
        super().__init__()
        assert layer_type in ("Conv", "FC", "RPN")
        self.layer_type = layer_type
        self.mc_samples = mc_samples
        self.block_size = block_size
        self.drop_prob = drop_prob

    def draw_seeds(self, latent_rep: torch.Tensor) -> torch.Tensor:
        """[n_mc, B, H, W] uint8: the `(torch.rand(B, H, W) < gamma)` draws of the n_mc DropBlock2D layers."""
        gamma = self.drop_prob / (self.block_size**2)
        shape = (latent_rep.shape[0], *latent_rep.shape[2:])
        # ONE draw of [n_mc, B, H, W]: torch's CPU uniform_ fills serially from the generator, so this is exactly the
        # sequence of the n_mc per-layer `torch.rand(B, H, W)` calls (tests/test_cabi_and_host.py checks it), at a tenth
        # of their cost
        return (torch.rand(self.mc_samples, *shape) < gamma).to(torch.uint8)

    def sample_batch(self, latent_rep: torch.Tensor, seeds: torch.Tensor = None) -> torch.Tensor:
        """[B, C, H, W] -> device rows [B * n_mc, C], item-major (rows b * n_mc .. b * n_mc + n_mc - 1 are image b):
        what the reference obtains by calling `forward` image by image and concatenating."""
        assert latent_rep.dim() == 4, f"Expected input with 4 dimensions (bsize, channels, height, width), got {latent_rep.dim()}"
        if not self.training or self.drop_prob == 0.0:  # DropBlock2D is the identity in eval mode
            mean = get_mean_or_fullmean_ls_sample(latent_rep, method="fullmean").reshape(latent_rep.shape[0], 1, -1)
            return mean.expand(-1, self.mc_samples, -1).reshape(-1, latent_rep.shape[1]).contiguous()
        if seeds is None:
            seeds = self.draw_seeds(latent_rep)
        return _ops.mc_dropblock_mean(latent_rep, seeds, self.block_size)

    def forward(self, latent_rep):
        """Reference semantics (one image per call): "Conv": [1, C, H, W] -> [n_mc, C] (fused kernel);
        "FC" / "RPN": the unreduced masked maps [n_mc, numel]."""
        if self.layer_type != "Conv":
            # "FC" / "RPN": the masked maps themselves, flattened ([n_mc, B*C*H*W]); no reduction follows
            if not self.training or self.drop_prob == 0.0:
                return latent_rep.reshape(1, -1).repeat(self.mc_samples, 1)
            out = _ops.mc_dropblock_apply(latent_rep, self.draw_seeds(latent_rep), self.block_size)
            return out if latent_rep.is_cuda else out.cpu()
        rows = self.sample_batch(latent_rep)
        if latent_rep.shape[0] == 1:
            return rows
        # the reference's reshape(1, -1) per sample for a batch: [n_mc, B * C]
        B, C = latent_rep.shape[0], latent_rep.shape[1]
        return rows.reshape(B, self.mc_samples, C).permute(1, 0, 2).reshape(self.mc_samples, B * C).contiguous()
