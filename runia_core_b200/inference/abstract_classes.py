"""Boundary types of the scoring hot path: same names, constructor signatures, flags and error
behaviour as the reference's `runia_core/inference/abstract_classes.py:35-211, 373-424`
(of the model-running classes only `InferenceModule` / `ProbabilisticInferenceModule`, :217-320, exist:
the bases of the online LaREx chain in `image_level.py`; the object-detection ones are out of scope)."""
from abc import ABC, abstractmethod
from time import monotonic
from typing import Dict, List, Union

import numpy as np
from numpy import ndarray

__all__ = [
    "record_time",
    "Postprocessor",
    "OodPostprocessor",
    "InferenceModule",
    "ProbabilisticInferenceModule",
    "get_baselines_thresholds",
    "get_method_threshold",
]


def record_time(function):
    """Returns (result, seconds) -- abstract_classes.py:35-52."""

    def wrapper(*args, **kwargs):
        t0 = monotonic()
        result = function(*args, **kwargs)
        return result, monotonic() - t0

    return wrapper


class Postprocessor(ABC):
    """Scoring operator interface (abstract_classes.py:58-130): `setup(ind_train_data, **kw)`
    fits on in-distribution data, `postprocess(test_data, **kw)` returns one score per row,
    calling the object is `postprocess`.  `cfg` is accepted and ignored here, as upstream."""

    def __init__(self, cfg=None):
        self._setup_flag = False

    @abstractmethod
    def setup(self, ind_train_data: ndarray, **kwargs) -> None:
        raise NotImplementedError

    @abstractmethod
    def postprocess(self, test_data: ndarray, **kwargs) -> ndarray:
        raise NotImplementedError

    def __call__(self, test_data: ndarray, **kwargs) -> ndarray:
        return self.postprocess(test_data, **kwargs)


class OodPostprocessor(Postprocessor):
    """Adds the sign convention and the detection threshold (abstract_classes.py:133-211)."""

    def __init__(self, flip_sign: bool, cfg=None):
        super().__init__(cfg)
        self.flip_sign = flip_sign
        self.threshold: Union[float, None] = None

    def flip_sign_fn(self, scores: Union[Dict[str, ndarray], ndarray]):
        if self.flip_sign:
            if isinstance(scores, dict):
                for method, values in scores.items():
                    scores[method] = values * -1
            elif isinstance(scores, ndarray):
                scores = scores * -1
            else:
                raise ValueError("scores must be a dict or ndarray")
        return scores

    def set_threshold(self, ind_test_scores: ndarray, z_score_percentile: float = 1.645) -> None:
        self.threshold = get_method_threshold(scores=ind_test_scores, z_score_percentile=z_score_percentile)
        self._setup_flag = True

    def setup(self, ind_train_data: ndarray, **kwargs) -> None:
        raise NotImplementedError

    def postprocess(self, test_data: ndarray, **kwargs) -> ndarray:
        raise NotImplementedError


class InferenceModule:
    """Model + postprocessor holder; moves the model to CUDA when present (abstract_classes.py:217-279)."""

    def __init__(self, model, postprocessor):
        import torch

        self.model = model
        self.postprocessor = postprocessor
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        try:
            self.model.to(self.device)
        except AttributeError:
            pass

    def get_score(self, input_image, *args, **kwargs):
        raise NotImplementedError


class ProbabilisticInferenceModule(InferenceModule):
    """Adds the MC-DropBlock parameters (abstract_classes.py:282-320)."""

    def __init__(self, model, postprocessor, drop_block_prob: float, drop_block_size: int, mcd_samples_nro: int):
        super().__init__(model, postprocessor)
        self.drop_block_prob = drop_block_prob
        self.drop_block_size = drop_block_size
        self.mcd_samples_nro = mcd_samples_nro


def get_method_threshold(scores: np.ndarray, z_score_percentile: float):
    """mean - z * std: 95 % of InD scores lie above it (abstract_classes.py:408-424)."""
    return float(np.mean(scores)) - z_score_percentile * float(np.std(scores))


def get_baselines_thresholds(baselines_names: List[str], baselines_scores_dict: Dict[str, np.ndarray],
                             z_score_percentile: float = 1.645) -> Dict[str, float]:
    """abstract_classes.py:373-405; the pseudo-method "raw" gets threshold 0."""
    out = {}
    for name in baselines_names:
        out[name] = 0.0 if name == "raw" else get_method_threshold(
            scores=baselines_scores_dict[name], z_score_percentile=z_score_percentile)
    return out
