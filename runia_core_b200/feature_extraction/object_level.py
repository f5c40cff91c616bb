"""The two reducers of the reference's object-level path (`runia_core/feature_extraction/object_level.py:254-366`)
that turn hooked detector feature maps + predicted boxes into the rows the scoring hot path consumes:

* `_reduce_features_to_rois`: RoIAlign every hooked map to the boxes, mean (optionally std) of every RoI over its
  P x P bins -> one row per detected object (the LaRD / LaREM / LaRED input of BASELINE configs[2]);
* `_dropblock_rois_get_entropy`: RoIAlign, MC-DropBlock sampling of every RoI map, per-dimension entropy.

The reference calls torchvision's `roi_align`, then `torch.mean` per object and per map in a Python loop; here the
mean never materialises the RoI maps (`runia_roi_align_mean_f32`) and the DropBlock path chains three kernels on the
device (`runia_roi_align_f32` -> `runia_mc_dropblock_mean_f32` -> `runia_mcd_entropy_f32`).  The detector wrappers
around them (`BoxFeaturesExtractor*`, hooks, datasets) are out of scope."""
from typing import List, Tuple

import torch
from torch import Tensor

from .. import _ops

__all__ = ["_reduce_features_to_rois", "_dropblock_rois_get_entropy"]


def _scale(latent: Tensor, img_shape: Tuple[int, ...]) -> float:
    return latent.shape[3] / img_shape[1]  # like upstream: map width over img_shape[1]


def _reduce_features_to_rois(latent_mcd_sample: List[Tensor], output_sizes: Tuple[int], boxes: Tensor,
                             img_shape: Tuple[int, ...], sampling_ratio: int, n_hooked_reps: int,
                             n_detected_objects: int, return_stds: bool = False) -> Tuple[List[Tensor], List[Tensor]]:
    """([1, sum_j C_j] per object, same for the stds or []) -- object_level.py:254-309."""
    means, stds = [], []
    for j in range(n_hooked_reps):
        m, s = _ops.roi_align_mean(latent_mcd_sample[j], boxes, output_sizes[j], _scale(latent_mcd_sample[j], img_shape),
                                   sampling_ratio, aligned=True, want_std=return_stds)
        means.append(m)
        stds.append(s)
    all_means = torch.cat(means, dim=1)[:n_detected_objects]
    n_objects_means = [all_means[i].reshape(1, -1) for i in range(all_means.shape[0])]
    n_objects_stds = []
    if return_stds:
        all_stds = torch.cat(stds, dim=1)[:n_detected_objects]
        n_objects_stds = [all_stds[i].reshape(1, -1) for i in range(all_stds.shape[0])]
    return n_objects_means, n_objects_stds


def _dropblock_rois_get_entropy(latent_mcd_sample: List[Tensor], output_sizes: Tuple[int], boxes: Tensor,
                                img_shape: Tuple[int, ...], sampling_ratio: int, n_hooked_reps: int, n_mcd_steps: int,
                                mc_sampler) -> Tensor:
    """Per-dimension entropies [n_objects, sum_j C_j] of the MC-DropBlock means of every RoI -- object_level.py:312-366.
    Like upstream, the hooked maps must share their pooled size when there are several (they are concatenated along
    the channels), every detection is sampled on its own (one DropBlock normalisation per RoI) and `get_dl_h_z` runs
    with mcd_samples_nro = n_mcd_steps."""
    rois = [_ops.roi_align(latent_mcd_sample[j], boxes, output_sizes[j], _scale(latent_mcd_sample[j], img_shape),
                           sampling_ratio, aligned=True) for j in range(n_hooked_reps)]
    rois = torch.cat(rois, dim=1) if len(rois) > 1 else rois[0]
    if hasattr(mc_sampler, "sample_batch") and getattr(mc_sampler, "layer_type", "Conv") == "Conv" and \
            mc_sampler.training and mc_sampler.drop_prob > 0.0:
        # upstream samples detection after detection (n_mc draws of [1, P, P] each): draw in that order, then hand
        # the sampler all detections at once -- rows come out item-major [K * n_mc, C], what the loop concatenates
        K, _, P_h, P_w = rois.shape
        gamma = mc_sampler.drop_prob / (mc_sampler.block_size**2)
        seeds = (torch.rand(K, mc_sampler.mc_samples, P_h, P_w) < gamma).permute(1, 0, 2, 3).to(torch.uint8).contiguous()
        rows = mc_sampler.sample_batch(rois, seeds=seeds)
    else:
        rows = torch.cat([mc_sampler(det.unsqueeze(0)) for det in rois], dim=0)
    _, h_z = _ops.mcd_entropy(rows if rows.is_cuda else rows.contiguous(), n_mcd_steps)
    return h_z.to(torch.float32).cpu()  # upstream: Tensor(entropies) -- float32 on the host
