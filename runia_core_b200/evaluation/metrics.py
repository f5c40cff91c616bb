"""`get_auroc_results` with the reference's signature and result table
(`runia_core/evaluation/metrics.py:37-100`), computed on the GPU: radix sort + fused ROC / PR scan
(`runia_ood_metrics`) instead of torchmetrics' CPU sort.  Plotting, MLflow logging and the
experiment drivers of that file are out of scope."""
from typing import Tuple, Union

import numpy as np
import pandas as pd

from .. import _ops

__all__ = ["get_auroc_results"]


def get_auroc_results(detect_exp_name: str, ind_samples_scores: np.ndarray, ood_samples_scores: np.ndarray,
                      return_results_for_mlflow: bool = False) -> Union[pd.DataFrame, Tuple[pd.DataFrame, dict]]:
    """AUROC, FPR@95, AUPR and the ROC curve of InD (positive class) vs OoD scores; same columns, row
    name and optional MLflow dictionary as upstream (metrics.py:82-100)."""
    m = _ops.ood_metrics(ind_samples_scores, ood_samples_scores, want_curve=True)
    # upstream reports float32 numbers (torchmetrics tensors): round the same way
    auroc = float(np.float32(m["auroc"]))
    fpr95 = float(np.float32(m["fpr95"]))
    aupr = float(np.float32(m["aupr"]))
    results_table = pd.DataFrame.from_dict(
        {detect_exp_name: [auroc, fpr95, aupr, m["fpr"].cpu().tolist(), m["tpr"].cpu().tolist()]},
        orient="index", columns=["auroc", "fpr@95", "aupr", "fpr", "tpr"])
    if not return_results_for_mlflow:
        return results_table
    results_for_mlflow = results_table.loc[detect_exp_name, ["auroc", "fpr@95", "aupr"]].to_dict()
    results_for_mlflow["fpr_95"] = results_for_mlflow.pop("fpr@95")  # MLflow does not accept '@'
    return results_table, results_for_mlflow
