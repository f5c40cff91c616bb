"""`calculate_all_baselines` and its helpers with the reference's names, signatures, dictionary keys and quirks
(`runia_core/evaluation/baselines.py:37-854`): the driver that runs every OoD baseline over pre-extracted features and
logits.  Upstream writes one near-identical function per baseline; here one table says, per baseline, which
postprocessor to build, what `setup()` gets and what `postprocess()` scores, and the public per-baseline functions
are thin wrappers over it -- every score comes from the CUDA postprocessors of `inference/postprocessors.py`."""
from typing import Dict, List, Tuple, Union

import numpy as np

from ..inference.postprocessors import ASH, DDU, DICE, GEN, KNN, MSP, DICEReAct, Energy, Mahalanobis, ReAct, ViM
from .. import _ops
from .._device import to_host

__all__ = [
    "remove_latent_features",
    "calculate_all_baselines",
    "get_labels_from_logits",
]  # `baseline_name_dict` (plot titles for the reference's plotting helpers) is presentation metadata: out of scope

Arr = Dict[str, np.ndarray]


def _run(postp, key: str, on: str, ind: Arr, ood: Arr, ood_names: List[str], scores: Arr, setup_kw: dict,
         with_logits: bool = False) -> Tuple[Arr, Arr]:
    """setup on the InD train split, then score the InD valid split (-> ind[key]) and every OoD set
    (-> scores["<ood> <key>"]); `on` is "features" or "logits"."""
    postp.setup(ind_train_data=ind[f"train {on}"], **setup_kw)
    extra = {"logits": ind["valid logits"]} if with_logits else {}
    ind[key] = postp.postprocess(test_data=ind[f"valid {on}"], **extra)
    for name in ood_names:
        extra = {"logits": ood[f"{name} logits"]} if with_logits else {}
        scores[f"{name} {key}"] = postp.postprocess(test_data=ood[f"{name} {on}"], **extra)
    return ind, scores


def get_dice_score_from_features(fc_params, ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, percentile):
    print("Calculating DICE score")
    postp = DICE(flip_sign=False, dice_percentile=percentile, num_classes=ind_data_dict["train logits"].shape[1])
    return _run(postp, "dice", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(valid_feats=ind_data_dict["valid features"], final_linear_layer_params=fc_params))


def get_react_score_from_features(fc_params, ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, percentile):
    print("Calculating ReAct score")
    postp = ReAct(flip_sign=False, react_percentile=percentile)
    return _run(postp, "react", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(valid_feats=ind_data_dict["valid features"], final_linear_layer_params=fc_params))


def get_dice_react_score_from_features(fc_params, ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                                       dice_percentile, react_percentile):
    print("Calculating DICE + ReAct score")
    postp = DICEReAct(flip_sign=False, dice_percentile=dice_percentile, react_percentile=react_percentile,
                      num_classes=ind_data_dict["train logits"].shape[1])
    return _run(postp, "dice_react", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(valid_feats=ind_data_dict["valid features"], final_linear_layer_params=fc_params))


def get_ash_score_from_features(fc_params, ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, ash_percentile):
    print("Calculating ASH score")
    postp = ASH(flip_sign=False, ash_percentile=ash_percentile)
    return _run(postp, "ash", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(valid_feats=ind_data_dict["valid features"], final_linear_layer_params=fc_params))


def get_gen_score_from_logits(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, gamma, gen_m):
    print("Calculating GEN score")
    postp = GEN(flip_sign=False, gamma=gamma, num_classes=gen_m)
    return _run(postp, "gen", "logits", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, {})


def calculate_vim_score(fc_params, ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict):
    print("Calculating ViM score")
    postp = ViM(flip_sign=False)
    return _run(postp, "vim", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(train_logits=ind_data_dict["train logits"], valid_feats=ind_data_dict["valid features"],
                     valid_logits=ind_data_dict["valid logits"], final_linear_layer_params=fc_params), with_logits=True)


def get_msp_score_from_logits(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict):
    print("Calculating MSP score")
    return _run(MSP(flip_sign=False), "msp", "logits", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, {})


def get_raw_score_from_logits(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict):
    """"raw" = the maximum softmax probability without a postprocessor object (baselines.py:395-425)."""
    msp = lambda lg: to_host(_ops.logit_scores(lg, energy=False, msp=True, gen=False)[1])  # noqa: E731
    ind_data_dict["raw"] = msp(ind_data_dict["valid logits"])
    for name in ood_names:
        ood_baselines_dict[f"{name} raw"] = msp(ood_data_dict[f"{name} logits"])
    return ind_data_dict, ood_baselines_dict


def get_energy_score_from_logits(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict):
    print("Calculating Energy score")
    return _run(Energy(flip_sign=False), "energy", "logits", ind_data_dict, ood_data_dict, ood_names,
                ood_baselines_dict, {})


def get_mahalanobis_score_from_features(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, num_classes):
    print("Calculating Mahalanobis distance score")
    postp = Mahalanobis(flip_sign=False, num_classes=num_classes)
    return _run(postp, "mdist", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(train_labels=ind_data_dict["train labels"], valid_feats=ind_data_dict["valid features"]))


def get_knn_score_from_features(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, k_neighbors=10):
    print("Calculating kNN score")
    postp = KNN(flip_sign=False, k_neighbors=k_neighbors)
    return _run(postp, "knn", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(valid_feats=ind_data_dict["valid features"]))


def get_ddu_score_from_features(ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict, num_classes):
    print("Calculating DDU score")
    postp = DDU(flip_sign=False, num_classes=num_classes)
    return _run(postp, "ddu", "features", ind_data_dict, ood_data_dict, ood_names, ood_baselines_dict,
                dict(train_labels=ind_data_dict["train labels"], valid_feats=ind_data_dict["valid features"]))


def _labels(logits):
    """argmax labels; a trailing background column (11 or 21 logits) is dropped first (baselines.py:645-676)."""
    if logits.shape[1] in (21, 11):
        logits = logits[:, :-1]
    return np.argmax(logits, axis=-1)


def get_labels_from_logits(id_data: Arr, ood_data: Arr, ood_names: List[str]):
    """Replaces the logits of the InD / OoD dictionaries by predicted labels (baselines.py:614-683): "train logits"
    / "valid logits" / "<ood> logits" are popped, "train labels" / "valid labels" / "<ood> labels" are written;
    empty lists mean "no logits were extracted" and give empty label arrays."""
    tr = id_data.pop("train logits", None) if "train logits" in id_data else None
    va = id_data.pop("valid logits", None) if "valid logits" in id_data else None
    if isinstance(tr, np.ndarray) or isinstance(va, np.ndarray):
        id_data["train labels"] = _labels(tr) if tr is not None else np.asarray([], dtype=int)
        id_data["valid labels"] = _labels(va) if va is not None else np.asarray([], dtype=int)
    elif isinstance(tr, list) and len(tr) == 0 and isinstance(va, list) and len(va) == 0:
        id_data["train labels"] = np.asarray([], dtype=int)
        id_data["valid labels"] = np.asarray([], dtype=int)
    else:
        raise NotImplementedError
    for name in ood_names:
        lg = ood_data.pop(f"{name} logits", None)
        if isinstance(lg, np.ndarray):
            ood_data[f"{name} labels"] = _labels(lg)
        elif isinstance(lg, list) and len(lg) == 0:
            ood_data[f"{name} labels"] = np.asarray([], dtype=int)
        else:
            raise NotImplementedError
    return id_data, ood_data


def remove_latent_features(id_data: Arr, ood_data: Arr, ood_names: List[str]):
    """Drops the feature arrays (baselines.py:686-710); missing keys are ignored."""
    id_data.pop("train features", None)
    id_data.pop("valid features", None)
    for name in ood_names:
        ood_data.pop(f"{name} features", None)
    return id_data, ood_data


def calculate_all_baselines(baselines_names: List[str], ind_data_dict: Arr, ood_data_dict: Arr,
                            fc_params: Union[Dict[str, np.ndarray], None], cfg, num_classes: int):
    """Runs the selected baselines in the reference's order (baselines.py:713-854): the logit / feature baselines
    first, then `get_labels_from_logits` (which REPLACES the logits by labels), then the two label-based ones
    ("mdist", "ddu").  Returns (ind_data_dict, ood_data_dict, {"<ood> <baseline>": scores})."""
    if num_classes > 21 and "gen" in baselines_names:
        raise ValueError(
            "Implementation of gen baseline does not yet support num_classes greater than 21. "
            "Otherwise implement M parameter specification"
        )
    names, scores = cfg.ood_datasets, {}
    common = dict(ind_data_dict=ind_data_dict, ood_data_dict=ood_data_dict, ood_names=names, ood_baselines_dict=scores)
    steps = [
        ("vim", lambda: calculate_vim_score(fc_params=fc_params, **common)),
        ("msp", lambda: get_msp_score_from_logits(**common)),
        ("raw", lambda: get_raw_score_from_logits(**common)),
        ("knn", lambda: get_knn_score_from_features(k_neighbors=cfg.k_neighbors, **common)),
        ("energy", lambda: get_energy_score_from_logits(**common)),
        ("ash", lambda: get_ash_score_from_features(fc_params=fc_params, ash_percentile=cfg.ash_percentile, **common)),
        ("gen", lambda: get_gen_score_from_logits(gamma=cfg.gen_gamma, gen_m=num_classes, **common)),
        ("react", lambda: get_react_score_from_features(fc_params=fc_params, percentile=cfg.react_percentile, **common)),
        ("dice", lambda: get_dice_score_from_features(fc_params=fc_params, percentile=cfg.dice_percentile, **common)),
        ("dice_react", lambda: get_dice_react_score_from_features(
            fc_params=fc_params, dice_percentile=cfg.dice_percentile, react_percentile=cfg.react_percentile, **common)),
    ]
    for name, fn in steps:
        if name in baselines_names:
            fn()
    ind_data_dict, ood_data_dict = get_labels_from_logits(id_data=ind_data_dict, ood_data=ood_data_dict, ood_names=names)
    common["ind_data_dict"], common["ood_data_dict"] = ind_data_dict, ood_data_dict
    if "mdist" in baselines_names:
        get_mahalanobis_score_from_features(num_classes=num_classes, **common)
    if "ddu" in baselines_names:
        get_ddu_score_from_features(num_classes=num_classes, **common)
    return ind_data_dict, ood_data_dict, scores
