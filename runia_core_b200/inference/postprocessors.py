"""Registry and the 16 scoring operators of the reference's
`runia_core/inference/postprocessors.py`, with the same registry keys, class names, constructor
signatures, `setup()` / `postprocess()` keyword arguments, state attributes, warnings and
assertion messages -- and every `postprocess()` body replaced by calls into the CUDA library.

`setup()` keeps the reference's one-off host fits (sklearn / NumPy / torch.distributions, cited
per class) and then uploads the fitted state; scoring InD validation data for thresholds already
runs on the GPU.  Inputs may be NumPy arrays (as upstream) or torch tensors, including CUDA
tensors (no host round trip); outputs are NumPy arrays of the dtype the reference returns.
"""
import warnings
from typing import Dict, List, Union

import numpy as np
import torch
from scipy.linalg import pinvh
from sklearn.covariance import EmpiricalCovariance
from torch import Tensor

from .. import _ops
from .._device import to_device, to_host
from .abstract_classes import OodPostprocessor, Postprocessor
from .funcs import RouteDICE, gmm_fit, mahalanobis_preprocess

__all__ = [
    "postprocessors_dict",
    "postprocessor_input_dict",
    "register_postprocessor",
    "DetectorKDE",
    "FlatL2Index",
    "LaREMPostprocessor",
    "LaREDPostprocessor",
]

_VALID_INPUT_TYPES = ("latent_space_means", "features", "logits")
postprocessors_dict: Dict[str, type] = {}
postprocessor_input_dict: Dict[str, List[str]] = {}


def register_postprocessor(postprocessor_name: str, postprocessor_input: List[str]):
    """Class decorator filling the two registries (postprocessors.py:50-75)."""

    def decorator(cls):
        for input_type in postprocessor_input:
            assert (
                input_type in _VALID_INPUT_TYPES
            ), f"Invalid input type {input_type}. Specify at least one of {_VALID_INPUT_TYPES}."
        postprocessors_dict[postprocessor_name] = cls
        postprocessor_input_dict[postprocessor_name] = postprocessor_input
        __all__.append(cls.__name__)
        return cls

    return decorator


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, Tensor) else np.asarray(x)


def _linear_params(kwargs):
    w, b = kwargs["final_linear_layer_params"]["weight"], kwargs["final_linear_layer_params"]["bias"]
    return _np(w), _np(b)


# ------------------------------------------------------------------------------------------------
# LaRED: Gaussian KDE over the latent bank            reference: postprocessors.py:78-178
# ------------------------------------------------------------------------------------------------
def _head_planes(W):
    """TF32 planes of a head's weight matrix for the tensor-core linear head, built once at setup
    (None when the head does not qualify: a width that is not a multiple of 4 or above 4096)."""
    if W.shape[1] % 4 == 0 and W.shape[1] <= _ops.TC_MAX_K:
        return _ops.linear_planes(W)
    return None


def _device_fit(feats, labels, num_classes):
    """float32 features -> (class means [C, d] float32, counts [C], shared precision [d, d] float64) from
    device statistics: NumPy-ordered means (bit-identical to `feats[labels == c].mean(0)`), the float64
    covariance of the class-centred rows (np.cov(..., bias=1), what EmpiricalCovariance.fit holds) and
    scipy's pinvh like sklearn's `_set_covariance`."""
    means, counts, xf, lab = _ops.class_means(feats, labels, num_classes)
    n_used = int(counts.sum())
    if n_used == 0:
        raise ValueError(f"Found array with 0 sample(s) (shape=(0, {xf.shape[1]})) while a minimum of 1 is required.")
    cov = _ops.centered_covariance(xf, lab, means, n_used)
    if not np.isfinite(cov).all():  # sklearn's validate_data refuses such rows
        raise ValueError("Input X contains NaN or infinity.")
    return to_host(means), counts, (_ops.pinvh(cov) if cov.shape[0] >= 64 else pinvh(cov, check_finite=False))


class DetectorKDE:
    """Gaussian kernel density estimate of the training embeddings (postprocessors.py:78-128).
    The reference fits sklearn's KernelDensity (a KD-tree) and queries it exactly (atol=rtol=0);
    here `density` is the device-resident bank and scoring is the closed form
    logsumexp_i(-|q-x_i|^2 / 2h^2) - log N - d/2 log(2 pi h^2) in one fused kernel."""

    def __init__(self, train_embeddings, save_path=None, kernel="gaussian", bandwidth=1.0, bank_group=None,
                 bank_rows_are_local=False) -> None:
        """bank_group: torch.distributed process group over which the bank is sharded (every rank keeps a contiguous
        slice; scoring all-reduces the running (max, sum-exp) pair: sharding.kde_score_sharded)."""
        if kernel != "gaussian":
            raise NotImplementedError("only the Gaussian kernel (the reference's default) is implemented")
        self.kernel = kernel
        self.bandwidth = bandwidth
        self.train_embeddings = train_embeddings
        self.save_path = save_path
        self._group = bank_group
        self._local = bank_rows_are_local
        self.density = self.density_fit()

    def density_fit(self):
        emb = _np(self.train_embeddings)
        if self._group is None:
            center = np.asarray(emb, np.float64).mean(0)
            return _ops.kde_bank(self.train_embeddings, self.bandwidth, center=center)
        import torch.distributed as dist

        from .. import sharding

        lo, hi, total = sharding.shard_of_bank(int(emb.shape[0]), self._group, self._local)
        mine = emb if self._local else emb[lo:hi]
        # any common centre is valid (it only keeps |q - b| small in float32): the mean of the whole bank
        acc = to_device(np.asarray(mine, np.float64).sum(0), torch.float64)
        if dist.get_world_size(self._group) > 1:
            dist.all_reduce(acc, group=self._group)
        return _ops.kde_bank(mine, self.bandwidth, center=acc / total, n_total=total)

    def get_density_scores(self, test_embeddings):
        if self._group is not None:
            from .. import sharding

            return to_host(sharding.kde_score_sharded(test_embeddings, self.density, group=self._group))
        return to_host(_ops.kde_score(test_embeddings, self.density))


@register_postprocessor("KDE", postprocessor_input=["latent_space_means"])
class KDELatentSpace(Postprocessor):
    """LaRED score (postprocessors.py:131-178): log-density, float64."""

    def __init__(self, cfg=None):
        super().__init__(cfg)
        self.detector = None

    def setup(self, ind_train_data: np.ndarray, **kwargs) -> None:
        assert ind_train_data.ndim == 2, "ind_feats must be 2 dimensional"
        if not self._setup_flag:
            self.detector = DetectorKDE(train_embeddings=ind_train_data, bank_group=kwargs.get("bank_group"),
                                        bank_rows_are_local=bool(kwargs.get("bank_rows_are_local", False)))
            self._setup_flag = True
        else:
            warnings.warn("KDEPostprocessor already trained")

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert test_data.ndim == 2, "ood_feats must be 2 dimensional"
        return self.detector.get_density_scores(test_data)


# ------------------------------------------------------------------------------------------------
# LaREM: Mahalanobis distance to the training distribution     reference: postprocessors.py:181-244
# ------------------------------------------------------------------------------------------------
@register_postprocessor("MD", postprocessor_input=["latent_space_means"])
class MDLatentSpace(Postprocessor):
    """-(x-mu)^T P (x-mu), float64.  setup: mean, EmpiricalCovariance(pinvh) on the host
    (postprocessors.py:212-220); postprocess: runia_rownorm_score_f32 against the factored P."""

    def __init__(self, cfg=None):
        super().__init__(cfg)
        self.feats_mean = None
        self.precision = None
        self.centered_data = None
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs) -> None:
        assert ind_train_data.ndim == 2, "ind_feats must be 2 dimensional"
        if not self._setup_flag:
            ind_train_data = _np(ind_train_data)
            if kwargs.get("bank_group") is not None:
                # training rows sharded over the ranks (bank_rows_are_local) or replicated (each rank fits its slice):
                # per-rank sufficient statistics + one all-reduce of d*d + d doubles (sharding.fit_mean_precision_sharded)
                from .. import sharding

                group, local = kwargs["bank_group"], bool(kwargs.get("bank_rows_are_local", False))
                lo, hi, _ = sharding.shard_of_bank(int(ind_train_data.shape[0]), group, local)
                mine = np.ascontiguousarray(ind_train_data if local else ind_train_data[lo:hi], np.float32)
                self.feats_mean, _, self.precision = sharding.fit_mean_precision_sharded(mine, None, 1, group=group)
                self.centered_data = ind_train_data - self.feats_mean
            elif ind_train_data.dtype == np.float32 and ind_train_data.shape[0] > 0:
                self.feats_mean, _, self.precision = _device_fit(ind_train_data, None, 1)
                self.centered_data = ind_train_data - self.feats_mean
            else:  # other dtypes: the reference's own host fit
                self.feats_mean = np.mean(ind_train_data, 0, keepdims=True)
                self.centered_data = ind_train_data - self.feats_mean
                ec = EmpiricalCovariance(assume_centered=False)
                ec.fit(self.centered_data)
                self.precision = ec.precision_
            self._state = _ops.md_prepare(self.feats_mean, self.precision)
            self._setup_flag = True
        else:
            warnings.warn("MDPostprocessor already trained")

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert test_data.ndim == 2, "test_feats must be 2 dimensional"
        return to_host(self.postprocess_device(test_data))

    def postprocess_device(self, test_data) -> Tensor:
        """Scores as a float64 CUDA tensor (no host copy): the form the online LaREx chain consumes."""
        if self._state is None:  # attributes assigned by hand (checkpoint restore)
            self._state = _ops.md_prepare(self.feats_mean, self.precision)
        return _ops.md_score(test_data, self._state, torch.float64)


# ------------------------------------------------------------------------------------------------
# cMD: class-conditional LaREM (float32 torch in the reference)       postprocessors.py:247-357
# ------------------------------------------------------------------------------------------------
@register_postprocessor("cMD", postprocessor_input=["latent_space_means"])
class cMDLatentSpace(Postprocessor):
    def __init__(self, cfg=None):
        super().__init__(cfg)
        try:
            self.num_classes = cfg.num_classes
        except AttributeError:
            self.num_classes = 10
        self.feats_mean = None
        self.precision = None
        self.class_mean = None
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs) -> None:
        try:
            ind_train_labels = kwargs["ind_train_labels"]
        except KeyError:
            raise ValueError("id_labels not provided. Pass ID train labels as 'ind_train_labels' argument.")
        ind_train_labels = _np(ind_train_labels)
        feats = _np(ind_train_data).astype(np.float32)
        assert feats.ndim == 2, "ind_feats must be 2 dimensional"
        if not self._setup_flag:
            cm, counts, precision = _device_fit(feats, ind_train_labels, self.num_classes)
            for c in np.flatnonzero(counts == 0):
                warnings.warn(f"No examples for class {c} to build class-wise Mahalanobis Distance score")
            self.class_mean = torch.from_numpy(cm)  # [#classes, d], like the reference
            self.precision = torch.from_numpy(precision).float()
            self._state = _ops.classcond_prepare(cm, precision)
            self._setup_flag = True
        else:
            warnings.warn("cMDPostprocessor already trained")

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        if "pred_labels" not in kwargs:
            raise ValueError("pred_logits not provided")
        assert test_data.ndim == 2, "test_feats must be 2 dimensional"
        return to_host(_ops.classcond_score(test_data, self._state, torch.float32))


# ------------------------------------------------------------------------------------------------
# kNN over L2-normalised latents                       postprocessors.py:360-423 / 789-883
# ------------------------------------------------------------------------------------------------
class FlatL2Index:
    """Stand-in for `faiss.IndexFlatL2` (the only faiss class the reference uses,
    postprocessors.py:396-397, 850-851): exact squared-L2 search over a device-resident bank."""

    def __init__(self, d: int, bank_group=None, idx_offset: int = 0, ntotal: int = None):
        """bank_group: a torch.distributed process group whose ranks each hold a contiguous slice of the bank
        (idx_offset = global index of this rank's first row, ntotal = rows of the whole bank); searches then merge
        the per-rank top-k over NCCL and return the same result on every rank (sharding.knn_search_sharded)."""
        self.d = d
        self.ntotal = 0
        self._bank = None
        self._group = bank_group
        self._offset = int(idx_offset)
        self._ntotal_global = ntotal

    def add(self, x):
        t = to_device(x, torch.float32)
        assert t.dim() == 2 and t.shape[1] == self.d
        full = t if self._bank is None else torch.cat([self._bank.bank, t])
        self._bank = _ops.knn_bank(full.contiguous(), idx_offset=self._offset)
        self.ntotal = int(full.shape[0]) if self._ntotal_global is None else int(self._ntotal_global)

    def _search(self, q, k):
        if self._group is None:
            return None
        from .. import sharding

        return sharding.knn_search_sharded(q, self._bank, k, group=self._group)

    def search(self, x, k: int):
        q = to_device(x, torch.float32)
        merged = self._search(q, k)
        if merged is not None:
            return to_host(merged[0]), to_host(merged[1])
        res = _ops.knn_search(q, self._bank, k, check_status=False)
        return to_host(res["dist"]), to_host(res["idx"])

    def kth_distance(self, q: torch.Tensor, k: int) -> torch.Tensor:
        merged = self._search(q, k)
        if merged is not None:
            return merged[2]
        return _ops.knn_search(q, self._bank, k, want_idx=False, want_dist=False, check_status=False)["kth"]


def _sharded_index(bank_normed: torch.Tensor, d: int, kwargs) -> FlatL2Index:
    """FlatL2Index over the (already normalised) training rows; with `bank_group=` in the setup kwargs every rank keeps
    only its slice of the bank (`bank_rows_are_local=True`: the rows handed in ARE this rank's slice)."""
    group = kwargs.get("bank_group")
    if group is None:
        index = FlatL2Index(d)
        index.add(bank_normed)
        return index
    from .. import sharding

    local = bool(kwargs.get("bank_rows_are_local", False))
    lo, hi, total = sharding.shard_of_bank(int(bank_normed.shape[0]), group, local)
    index = FlatL2Index(d, bank_group=group, idx_offset=lo, ntotal=total)
    index.add(bank_normed if local else bank_normed[lo:hi].contiguous())
    return index


def _knn_scores(index: FlatL2Index, test_data, k: int) -> np.ndarray:
    qn = _ops.normalize_rows(test_data)
    return -to_host(index.kth_distance(qn, k))


@register_postprocessor("KNN", postprocessor_input=["latent_space_means"])
class KNNLatentSpace(Postprocessor):
    """minus the squared distance to the K-th neighbour among normalised training latents, float32
    (postprocessors.py:385-423); K from cfg.k_neighbors, default 50."""

    def __init__(self, cfg=None):
        super().__init__(cfg)
        try:
            self.K = cfg.k_neighbors
        except AttributeError:
            self.K = 50
        self.activation_log = None
        self.index = None

    def setup(self, ind_train_data: np.ndarray, **kwargs) -> None:
        assert ind_train_data.ndim == 2, "ind_train_feats must be 2 dimensional"
        if not self._setup_flag:
            bank = _ops.normalize_rows(ind_train_data)
            self.activation_log = to_host(bank)
            self.index = _sharded_index(bank, ind_train_data.shape[1], kwargs)
            self._setup_flag = True
        else:
            warnings.warn("KNNPostprocessor already trained")

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert test_data.ndim == 2, "test_feats must be 2 dimensional"
        return _knn_scores(self.index, test_data, self.K)


# ------------------------------------------------------------------------------------------------
# GMM (LaREG) and DDU: class-wise Gaussian mixture log-density   postprocessors.py:426-492, 694-786
# ------------------------------------------------------------------------------------------------
def _gmm_state(gmm):
    return _ops.gmm_prepare(gmm.loc, gmm.scale_tril)


@register_postprocessor("GMM", postprocessor_input=["latent_space_means"])
class GMMLatentSpace(Postprocessor):
    def __init__(self, cfg=None):
        super().__init__(cfg)
        try:
            self.num_classes = cfg.num_classes
        except AttributeError:
            self.num_classes = 10
        self.gmm = None
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs) -> None:
        assert ind_train_data.ndim == 2, "ind_train_feats must be 2 dimensional"
        if not self._setup_flag:
            try:
                labels = kwargs["ind_train_labels"]
            except KeyError:
                raise ValueError("id_labels not provided")
            self.gmm, _ = gmm_fit(embeddings=Tensor(_np(ind_train_data)), labels=Tensor(_np(labels)),
                                  num_classes=self.num_classes)
            self._state = _gmm_state(self.gmm)
            self._setup_flag = True
        else:
            warnings.warn("GMMPostprocessor already trained")

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert test_data.ndim == 2, "test_feats must be 2 dimensional"
        return to_host(_ops.gmm_lse(test_data, self._state))


# ------------------------------------------------------------------------------------------------
# logit-space scores: one fused pass                              postprocessors.py:495-691
# ------------------------------------------------------------------------------------------------
def _logit_score(test_data, which, gamma=0.1, M=None):
    e, m, g, in_dtype = _ops.logit_scores(test_data, gamma=gamma, M=M, energy=which == "energy",
                                          msp=which == "msp", gen=which == "gen")
    out = to_host({"energy": e, "msp": m, "gen": g}[which])
    return out.astype(np.float64) if in_dtype == torch.float64 else out


@register_postprocessor("energy", postprocessor_input=["logits"])
class Energy(OodPostprocessor):
    def setup(self, ind_train_data: np.ndarray, **kwargs):
        ind_scores = self.flip_sign_fn(_logit_score(ind_train_data, "energy"))
        self.set_threshold(ind_scores)

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(_logit_score(test_data, "energy"))


@register_postprocessor("msp", postprocessor_input=["logits"])
class MSP(OodPostprocessor):
    def setup(self, ind_train_data: np.ndarray, **kwargs):
        ind_scores = self.flip_sign_fn(_logit_score(ind_train_data, "msp"))
        self.set_threshold(ind_scores)

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(_logit_score(test_data, "msp"))


@register_postprocessor("gen", postprocessor_input=["logits"])
class GEN(OodPostprocessor):
    def __init__(self, flip_sign: bool, gamma: float, num_classes: int, cfg=None):
        super().__init__(flip_sign, cfg)
        self.gamma = gamma
        self.num_classes = num_classes

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        ind_scores = self.flip_sign_fn(_logit_score(ind_train_data, "gen", self.gamma, self.num_classes))
        self.set_threshold(ind_scores)

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(_logit_score(test_data, "gen", self.gamma, self.num_classes))


@register_postprocessor("ddu", postprocessor_input=["features"])
class DDU(OodPostprocessor):
    def __init__(self, flip_sign: bool, num_classes: int, cfg=None):
        super().__init__(flip_sign, cfg)
        self.num_classes = num_classes
        self.gmm = None
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "valid_feats" in kwargs, "valid_feats must be provided for DDU"
        assert "train_labels" in kwargs, "train_labels must be provided for DDU"
        # like the reference (:751-755) the fit runs where self.device points: torch on the GPU when there is one
        self.gmm, _ = gmm_fit(embeddings=Tensor(_np(ind_train_data)).to(self.device),
                              labels=Tensor(_np(kwargs["train_labels"])).to(self.device), num_classes=self.num_classes)
        self._state = _gmm_state(self.gmm)
        ind_scores = self.flip_sign_fn(to_host(_ops.gmm_lse(kwargs["valid_feats"], self._state)))
        self.set_threshold(ind_scores)

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(to_host(_ops.gmm_lse(test_data, self._state)))


@register_postprocessor("knn", postprocessor_input=["features"])
class KNN(OodPostprocessor):
    def __init__(self, flip_sign: bool, k_neighbors: int, cfg=None):
        super().__init__(flip_sign, cfg)
        self.k_neighbors = k_neighbors
        self.gmm = None
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.index = None

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "valid_feats" in kwargs, "valid_feats must be provided for KNN setup"
        bank = _ops.normalize_rows(ind_train_data)
        self.index = _sharded_index(bank, ind_train_data.shape[1], kwargs)
        # like the reference (postprocessors.py:852-854): postprocess() already applies flip_sign_fn and setup flips
        # the result once more, so with flip_sign=True the threshold comes from the UN-flipped validation scores
        ind_scores = self.flip_sign_fn(self.postprocess(kwargs["valid_feats"]))
        self.set_threshold(ind_scores)

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        return self.flip_sign_fn(_knn_scores(self.index, test_data, self.k_neighbors))


@register_postprocessor("mahalanobis", postprocessor_input=["features"])
class Mahalanobis(OodPostprocessor):
    def __init__(self, flip_sign: bool, num_classes: int, cfg=None):
        super().__init__(flip_sign, cfg)
        self.num_classes = num_classes
        self.class_mean = None
        self.precision = None
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "train_labels" in kwargs, "train_labels must be provided for Mahalanobis"
        assert "valid_feats" in kwargs, "valid_feats must be provided for Mahalanobis"
        ind = {"train features": _np(ind_train_data), "train labels": _np(kwargs["train_labels"])}
        self.class_mean, self.precision = mahalanobis_preprocess(ind_data=ind, num_classes=self.num_classes)
        self._state = _ops.classcond_prepare(self.class_mean, self.precision)
        ind_scores = to_host(_ops.classcond_score(kwargs["valid_feats"], self._state, torch.float64))
        self.set_threshold(self.flip_sign_fn(ind_scores))

    def postprocess(self, test_data: Union[np.ndarray, Tensor], **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(to_host(_ops.classcond_score(test_data, self._state, torch.float64)))


@register_postprocessor("vim", postprocessor_input=["features", "logits"])
class ViM(OodPostprocessor):
    """-alpha * ||(x-u) NS|| + logsumexp(logits) (postprocessors.py:983-1112).  setup: pinv of the head on the host
    ([C, d]); covariance of the shifted rows and its eigendecomposition on the device for float32 features of width
    >= 64 (otherwise the reference's host expressions: EmpiricalCovariance(assume_centered=True), np.linalg.eig)."""

    def __init__(self, flip_sign: bool, cfg=None):
        super().__init__(flip_sign, cfg)
        self.u = None
        self.DIM = None
        self.NS = None
        self.alpha = None
        self._state = None

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "final_linear_layer_params" in kwargs, "final_linear_layer_params must be provided for ViM"
        assert "train_logits" in kwargs, "train_logits must be provided for ViM"
        assert "valid_feats" in kwargs, "valid_feats must be provided for ViM"
        assert "valid_logits" in kwargs, "valid_logits must be provided for ViM"
        w, b = _linear_params(kwargs)
        train = _np(ind_train_data)
        with _ops.host_blas_single_thread():
            self.u = -np.matmul(np.linalg.pinv(w), b)
        if train.shape[-1] >= 2048:
            self.DIM = 1000
        elif train.shape[-1] >= 768:
            self.DIM = 512
        else:
            self.DIM = train.shape[-1] // 2
        if train.dtype == np.float32 and train.ndim == 2 and train.shape[1] >= 64 and train.shape[0] > 0 \
                and np.asarray(self.u).dtype in (np.float32, np.float64):
            # device fit: float64 Gram matrix of the shifted rows -- r = f32(x - u) for a float32 u (a float32 head:
            # NumPy subtracts in float32; `runia_centered_gram_f64`), r = f64(x) - u for a float64 one
            # (`runia_shifted_gram_f64`) -- and the Jacobi eigensolver.  The residual norm only depends on the SPAN of
            # the discarded eigenvectors, which a symmetric solver and the reference's np.linalg.eig agree on.
            train = to_device(train)
            if np.asarray(self.u).dtype == np.float64:
                cov = _ops.shifted_covariance(train, self.u)
            else:
                G, _ = _ops.centered_gram(train, None, to_device(np.ascontiguousarray(self.u, np.float32).reshape(1, -1)))
                cov = (G / train.shape[0]).cpu().numpy()
            eig_vals, eigen_vectors = _ops.eigh(cov)
        else:
            ec = EmpiricalCovariance(assume_centered=True)
            ec.fit(train - self.u)
            eig_vals, eigen_vectors = np.linalg.eig(ec.covariance_)
        self.NS = np.ascontiguousarray((eigen_vectors.T[np.argsort(eig_vals * -1)[self.DIM:]]).T)
        st = _ops.vim_prepare(self.u, self.NS, 1.0)
        vlogit_train = to_host(_ops.residual_norm(train, st))
        self.alpha = _np(kwargs["train_logits"]).max(axis=-1).mean() / vlogit_train.mean()
        self._state = _ops.vim_prepare(self.u, self.NS, self.alpha)
        ind_scores = to_host(_ops.vim_score(kwargs["valid_feats"], kwargs["valid_logits"], self._state))
        self.set_threshold(self.flip_sign_fn(ind_scores))

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return to_host(_ops.vim_score(test_data, kwargs["logits"], self._state))  # no sign flip upstream


# ------------------------------------------------------------------------------------------------
# feature-shaping baselines: clip / mask / prune -> linear -> log-sum-exp   postprocessors.py:1115-1621
# ------------------------------------------------------------------------------------------------
@register_postprocessor("ash", postprocessor_input=["features"])
class ASH(OodPostprocessor):
    def __init__(self, flip_sign: bool, ash_percentile: int = 85, cfg=None):
        super().__init__(flip_sign, cfg)
        self.ash_percentile = ash_percentile
        self.w = None
        self.b = None

    def _score(self, x):
        n = x.shape[1]
        k = n - int(np.round(n * self.ash_percentile / 100.0))
        return to_host(_ops.ash_linear_lse(x, self._w, self._b, k, planes=self._planes))

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "final_linear_layer_params" in kwargs, "final_linear_layer_params must be provided for ASH"
        assert "valid_feats" in kwargs, "valid_feats must be provided for ASH"
        self.w, self.b = _linear_params(kwargs)
        self._w, self._b = to_device(self.w, torch.float32), to_device(self.b, torch.float32)
        self._planes = None if _ops.head_fits_smem(*self._w.shape) else _head_planes(self._w)
        # the reference thresholds on the TRAIN features here (postprocessors.py:1185)
        self.set_threshold(self.flip_sign_fn(self._score(ind_train_data)))

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(self._score(test_data))


def _train_percentile(ind_train_data, q):
    """np.percentile(train.flatten(), q) (postprocessors.py:1433, 1576); float32 banks are sorted on the GPU."""
    a = _np(ind_train_data)
    if a.dtype == np.float32 and a.size >= 1 << 16:
        return _ops.percentile_f32(a, q)
    return np.percentile(a.flatten(), q)


def _make_dice_layer(ind_train_data, kwargs, num_classes, percentile):
    w, b = kwargs["final_linear_layer_params"]["weight"], kwargs["final_linear_layer_params"]["bias"]
    params = {"weight": Tensor(w) if isinstance(w, np.ndarray) else w,
              "bias": Tensor(b) if isinstance(b, np.ndarray) else b}
    info = Tensor(_np(ind_train_data)).mean(0).cpu().numpy()
    layer = RouteDICE(in_features=ind_train_data.shape[1], out_features=num_classes, bias=True,
                      p=percentile, info=info)
    layer.load_state_dict(params)
    layer.eval()
    layer.calculate_mask_weight()
    return layer


@register_postprocessor("dice", postprocessor_input=["features"])
class DICE(OodPostprocessor):
    def __init__(self, flip_sign: bool, dice_percentile: int = 90, num_classes: int = 10, cfg=None):
        super().__init__(flip_sign, cfg)
        self.dice_percentile = dice_percentile
        self.num_classes = num_classes
        self.dice_layer = None
        self.device = "cuda" if torch.cuda.is_available() else "cpu"

    def _score(self, x):
        return to_host(_ops.clip_linear_lse(x, self.dice_layer.masked_w, self._b, planes=getattr(self, "_planes", None)))

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "final_linear_layer_params" in kwargs, "final_linear_layer_params must be provided for DICE"
        assert "valid_feats" in kwargs, "valid_feats must be provided for DICE"
        self.dice_layer = _make_dice_layer(ind_train_data, kwargs, self.num_classes, self.dice_percentile)
        self._b = to_device(self.dice_layer.bias.detach(), torch.float32)
        self._planes = _head_planes(self.dice_layer.masked_w)
        self.set_threshold(self.flip_sign_fn(self._score(kwargs["valid_feats"])))

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(self._score(test_data))


@register_postprocessor("react", postprocessor_input=["features"])
class ReAct(OodPostprocessor):
    def __init__(self, flip_sign: bool, react_percentile: int = 90, cfg=None):
        super().__init__(flip_sign, cfg)
        self.react_percentile = react_percentile
        self.activation_threshold = None
        self.w = None
        self.b = None

    def _score(self, x):
        return to_host(_ops.clip_linear_lse(x, self._w, self._b, clip=float(self.activation_threshold), planes=getattr(self, "_planes", None)))

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "final_linear_layer_params" in kwargs, "final_linear_layer_params must be provided for ReAct"
        assert "valid_feats" in kwargs, "valid_feats must be provided for ReAct"
        self.w, self.b = _linear_params(kwargs)
        self._w, self._b = to_device(self.w, torch.float32), to_device(self.b, torch.float32)
        self._planes = _head_planes(self._w)
        self.activation_threshold = _train_percentile(ind_train_data, self.react_percentile)
        self.set_threshold(self.flip_sign_fn(self._score(kwargs["valid_feats"])))

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(self._score(test_data))


@register_postprocessor("dice_react", postprocessor_input=["features"])
class DICEReAct(OodPostprocessor):
    def __init__(self, flip_sign: bool, dice_percentile: int = 90, react_percentile: int = 90,
                 num_classes: int = 10, cfg=None):
        super().__init__(flip_sign, cfg)
        self.dice_percentile = dice_percentile
        self.react_percentile = react_percentile
        self.num_classes = num_classes
        self.dice_layer = None
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.react_activation_threshold = None

    def _score(self, x):
        return to_host(_ops.clip_linear_lse(x, self.dice_layer.masked_w, self._b,
                                            clip=float(self.react_activation_threshold), planes=getattr(self, "_planes", None)))

    def setup(self, ind_train_data: np.ndarray, **kwargs):
        assert "final_linear_layer_params" in kwargs, "final_linear_layer_params must be provided for DICE"
        assert "valid_feats" in kwargs, "valid_feats must be provided for DICE"
        self.dice_layer = _make_dice_layer(ind_train_data, kwargs, self.num_classes, self.dice_percentile)
        self._b = to_device(self.dice_layer.bias.detach(), torch.float32)
        self._planes = _head_planes(self.dice_layer.masked_w)
        self.react_activation_threshold = _train_percentile(ind_train_data, self.react_percentile)
        self.set_threshold(self.flip_sign_fn(self._score(kwargs["valid_feats"])))

    def postprocess(self, test_data: np.ndarray, **kwargs) -> np.ndarray:
        assert self._setup_flag, "setup() must be called before postprocess()"
        return self.flip_sign_fn(self._score(test_data))


# Names used by the reference's README / the north star for the two headline scorers
LaREMPostprocessor = MDLatentSpace
LaREDPostprocessor = KDELatentSpace
