"""`LaRExInference` of the reference's `runia_core/inference/image_level.py:31-120`: online LaREx scoring of
one image -- model forward, MC-DropBlock samples of the hooked latent map, per-dimension entropy, optional PCA,
LaRED / LaREM score.

The reference hops host <-> device between every stage (`get_dl_h_z` copies the samples to the CPU and runs
`n_mc * D` KD-tree queries, sklearn's PCA and the postprocessor are NumPy).  Here the chain after the model
forward stays on the GPU: `runia_mc_dropblock_mean_f32` -> `runia_mcd_entropy_f32` -> `runia_pca_transform_*`
-> `runia_rownorm_score_*`, four launches on one stream and a single copy of the score back to the host.
"""
import numpy as np
import torch

from .. import _ops
from .._device import to_device, to_host
from .abstract_classes import ProbabilisticInferenceModule

__all__ = ["LaRExInference", "FoldedLaREM"]


class FoldedLaREM:
    """PCA transform + LaREM (`apply_pca_transform` -> `MDLatentSpace.postprocess`,
    dimensionality_reduction.py:86 + postprocessors.py:241-243) as one contraction over the raw latents:
    the fitted PCA and the factored precision are multiplied together once on the host, and the row scorer
    reads each raw latent once -- no intermediate [N, d] array.  `postprocess` has the postprocessor's signature."""

    def __init__(self, pca_transform, md_postprocessor):
        self._state = _ops.md_fold_pca(pca_transform.mean_, pca_transform.components_, pca_transform.explained_variance_,
                                       pca_transform.whiten, md_postprocessor.feats_mean, md_postprocessor.precision)
        if self._state is None:
            raise ValueError("PCA and precision cannot be folded (rank-deficient product)")

    def postprocess_device(self, test_data) -> torch.Tensor:
        return _ops.md_score(test_data, self._state, torch.float64)

    def postprocess(self, test_data, **kwargs) -> np.ndarray:
        assert test_data.ndim == 2, "test_feats must be 2 dimensional"
        return to_host(self.postprocess_device(test_data))

    __call__ = postprocess


class LaRExInference(ProbabilisticInferenceModule):
    def __init__(self, model, postprocessor, drop_block_prob: float, drop_block_size: int, mcd_samples_nro: int,
                 mcd_sampler, pca_transform=None, layer_type="Conv"):
        super().__init__(model=model, postprocessor=postprocessor, drop_block_prob=drop_block_prob,
                         drop_block_size=drop_block_size, mcd_samples_nro=mcd_samples_nro)
        self.layer_type = layer_type
        self.pca_transform = pca_transform
        self.mc_sampler = mcd_sampler(mc_samples=self.mcd_samples_nro, layer_type=layer_type,
                                      drop_prob=self.drop_block_prob, block_size=self.drop_block_size)
        self.mc_sampler.to(self.device)
        self.mc_sampler.train()
        self._folded = None
        if pca_transform is not None and hasattr(postprocessor, "precision") and hasattr(postprocessor, "feats_mean") \
                and hasattr(pca_transform, "components_") and getattr(postprocessor, "precision", None) is not None:
            try:
                self._folded = FoldedLaREM(pca_transform, postprocessor)  # PCA + LaREM as one launch
            except ValueError:
                self._folded = None

    def score_samples(self, mc_samples_t) -> np.ndarray:
        """[N * n_mc, D] MC samples (any device) -> LaREx scores [N]: entropy -> PCA -> postprocessor, on the GPU
        when the PCA / postprocessor expose their device entry points, through their public methods otherwise."""
        z = to_device(mc_samples_t, torch.float32)
        n_mc = self.mcd_samples_nro
        if z.shape[0] % n_mc != 0:
            raise ValueError(f"{z.shape[0]} sample rows are not a multiple of mcd_samples_nro={n_mc}")
        _, h_z = _ops.mcd_entropy(z, n_mc, k=_ops.entropy_k(n_mc), want_joint=False)
        if self._folded is not None:
            return to_host(self._folded.postprocess_device(h_z))
        if self.pca_transform:
            h_z = self.pca_transform.transform_device(h_z) if hasattr(self.pca_transform, "transform_device") \
                else self.pca_transform.transform(to_host(h_z))
        if hasattr(self.postprocessor, "postprocess_device") and isinstance(h_z, torch.Tensor):
            return to_host(self.postprocessor.postprocess_device(h_z))
        return self.postprocessor.postprocess(to_host(h_z) if isinstance(h_z, torch.Tensor) else h_z)

    def get_score(self, input_image, layer_hook):
        """(model output, LaREx score [1]) for one image (image_level.py:95-120)."""
        with torch.no_grad():
            try:
                input_image = input_image.to(self.device)
            except AttributeError:  # pragma: no cover
                pass
            output = self.model(input_image)
            latent_rep = layer_hook.output  # latent representation sample
        mc_samples_t = self.mc_sampler(latent_rep)
        return output, self.score_samples(mc_samples_t)

    def get_score_full_inference(self, input_image, layer_hook):
        raise NotImplementedError
