"""`LaRExInference` of the reference's `runia_core/inference/image_level.py:31-120`: online LaREx scoring of
one image -- model forward, MC-DropBlock samples of the hooked latent map, per-dimension entropy, optional PCA,
LaRED / LaREM score.

The reference hops host <-> device between every stage (`get_dl_h_z` copies the samples to the CPU and runs
`n_mc * D` KD-tree queries, sklearn's PCA and the postprocessor are NumPy).  Here the chain after the model
forward stays on the GPU: `runia_mc_dropblock_mean_f32` -> `runia_mcd_entropy_f32` -> `runia_pca_transform_*`
-> `runia_rownorm_score_*`, four launches on one stream and a single copy of the score back to the host.

`get_score` replays that chain as ONE CUDA graph per latent-map shape (static device buffers for the map, the seeds,
the sample rows and the score; the seeds reach the device through a pinned buffer) -- the launches, their argument
marshalling and the allocator calls leave the per-image path.
"""
import numpy as np
import torch

from .. import _ops
from .._device import device, to_device, to_host
from ..feature_extraction.utils import get_mean_or_fullmean_ls_sample
from .abstract_classes import InferenceModule, ProbabilisticInferenceModule, record_time

__all__ = ["LaRExInference", "LaRDInference", "FoldedLaREM"]


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
        self._graphs = {}
        self.use_cuda_graph = True
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

    # ---- CUDA-graph fast path: sampler -> entropy -> folded PCA + LaREM on static buffers -------------------------
    def _graph_for(self, latent_rep):
        """The captured chain for this latent-map shape, or None when the configuration has no single-launch-chain
        form (FC / RPN layers, eval-mode sampler, no folded PCA + LaREM)."""
        smp = self.mc_sampler
        if self._folded is None or getattr(smp, "layer_type", None) != "Conv" or not smp.training or \
                smp.drop_prob == 0.0 or not hasattr(smp, "draw_seeds") or latent_rep.dim() != 4:
            return None
        key = tuple(latent_rep.shape)
        g = self._graphs.get(key)
        if g is not None:
            return g
        dev = device()
        B, C, H, W = key
        n_mc = self.mcd_samples_nro
        st = {"x": torch.empty(key, dtype=torch.float32, device=dev),
              "seed_host": torch.empty((n_mc, B, H, W), dtype=torch.uint8).pin_memory(),
              "seed": torch.empty((n_mc, B, H, W), dtype=torch.uint8, device=dev),
              "score_host": torch.empty((B,), dtype=torch.float64).pin_memory()}

        def chain():
            rows = _ops.mc_dropblock_mean(st["x"], st["seed"], smp.block_size)
            _, h_z = _ops.mcd_entropy(rows, n_mc, k=_ops.entropy_k(n_mc), want_joint=False)
            return self._folded.postprocess_device(h_z)

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture (attribute calls, lazy initialisation)
            st["x"].zero_()
            st["seed"].zero_()
            chain()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["out"] = chain()
        st["graph"] = graph
        self._graphs[key] = st
        return st

    def _score_graph(self, st, latent_rep):
        st["seed_host"].copy_(self.mc_sampler.draw_seeds(latent_rep))
        st["x"].copy_(latent_rep, non_blocking=True)
        st["seed"].copy_(st["seed_host"], non_blocking=True)
        st["graph"].replay()
        st["score_host"].copy_(st["out"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return st["score_host"].numpy().copy()

    def get_score(self, input_image, layer_hook):
        """(model output, LaREx score [1]) for one image (image_level.py:95-120)."""
        with torch.no_grad():
            try:
                input_image = input_image.to(self.device)
            except AttributeError:  # pragma: no cover
                pass
            output = self.model(input_image)
            latent_rep = layer_hook.output  # latent representation sample
        return output, self.score_latent(latent_rep)

    def score_latent(self, latent_rep) -> np.ndarray:
        """LaREx score(s) of one hooked latent map [B, C, H, W] (the part of get_score after the model forward)."""
        st = self._graph_for(latent_rep) if self.use_cuda_graph and isinstance(latent_rep, torch.Tensor) and \
            latent_rep.is_cuda and latent_rep.dtype == torch.float32 else None
        if st is not None:
            return self._score_graph(st, latent_rep)
        if getattr(self.mc_sampler, "layer_type", None) == "Conv" and hasattr(self.mc_sampler, "sample_batch") and \
                isinstance(latent_rep, torch.Tensor) and latent_rep.dim() == 4:
            return self.score_samples(self.mc_sampler.sample_batch(latent_rep))  # item-major rows: one score per map
        return self.score_samples(self.mc_sampler(latent_rep))

    @record_time
    def test_time_inference(self, input_image, layer_hook):
        """(get_score result, seconds) -- image_level.py:122-134."""
        return self.get_score(input_image, layer_hook)

    @record_time
    def get_layer_mc_samples(self, input_image, layer_hook):
        """Model forward + MC sampling only, timed (image_level.py:136-155)."""
        with torch.no_grad():
            input_image = input_image.to(self.device)
            _ = self.model(input_image)
            latent_rep = layer_hook.output
        return self.mc_sampler(latent_rep)

    @record_time
    def get_mc_samples_full_inference(self, input_image, layer_hook):
        """mcd_samples_nro complete forward passes, hooked maps concatenated (image_level.py:157-183)."""
        mc_samples = []
        with torch.no_grad():
            for _ in range(self.mcd_samples_nro):
                try:
                    input_image = input_image.to(self.device)
                except AttributeError:
                    pass
                _ = self.model(input_image)
                mc_samples.append(layer_hook.output)
            return torch.cat(mc_samples).cpu().numpy()

    @record_time
    def get_score_full_inference(self, input_image, layer_hook):
        """Abstract upstream as well (image_level.py:185-198: `raise NotImplementedError`, "should be implemented
        in child class")."""
        raise NotImplementedError


class LaRDInference(InferenceModule):
    """LaRD (image_level.py:201-315): representation reduction + density score, no MC sampling and no entropy.
    model forward -> H x W mean per channel ("Conv") or column mean ("FC") -> optional PCA -> KDE / MD
    postprocessor.  The reduction, the projection and the score stay on the GPU when the PCA / postprocessor expose
    their device entry points (B200PCA.transform_device, MDLatentSpace.postprocess_device)."""

    def __init__(self, model, postprocessor, pca_transform=None, layer_type="Conv") -> None:
        super().__init__(model, postprocessor)
        self.layer_type = layer_type
        if self.layer_type == "Conv":
            self._reducer = self._reduce_conv_representation
        elif self.layer_type == "FC":
            self._reducer = self._reduce_fc_representation
        else:
            pass  # "RPN" lives in its own subclass upstream
        self.pca_transform = pca_transform

    def score_latent(self, latent_rep) -> np.ndarray:
        if self.layer_type == "Conv" and isinstance(latent_rep, torch.Tensor) and latent_rep.dim() == 4 and \
                hasattr(self.postprocessor, "postprocess_device") and \
                (self.pca_transform is None or hasattr(self.pca_transform, "transform_device")):
            z = get_mean_or_fullmean_ls_sample(to_device(latent_rep, torch.float32), "fullmean").reshape(1, -1)
            if self.pca_transform is not None:
                z = self.pca_transform.transform_device(z)
            return to_host(self.postprocessor.postprocess_device(z))
        z = self._reducer(latent_rep)
        if self.pca_transform:
            from ..dimensionality_reduction import apply_pca_transform

            z = apply_pca_transform(z, self.pca_transform)
        return self.postprocessor.postprocess(z)

    def get_score(self, input_image, layer_hook):
        """(model output, LaRD score [1]) for one image (image_level.py:244-266)."""
        with torch.no_grad():
            try:
                input_image = input_image.to(self.device)
            except AttributeError:
                pass
            output = self.model(input_image)
            latent_rep = layer_hook.output
        return output, self.score_latent(latent_rep)

    @record_time
    def test_time_inference(self, input_image, layer_hook):
        return self.get_score(input_image, layer_hook)

    @staticmethod
    def _reduce_conv_representation(representation) -> np.ndarray:
        return get_mean_or_fullmean_ls_sample(representation, "fullmean").cpu().numpy().reshape(1, -1)

    @staticmethod
    def _reduce_fc_representation(representation) -> np.ndarray:
        if representation.ndim > 1:
            return torch.mean(representation, dim=1).cpu().numpy().reshape(1, -1)
        return representation.reshape(1, -1).cpu().numpy()
