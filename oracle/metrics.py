"""
oracle/metrics.py -- TEST INFRASTRUCTURE ONLY (not a product path).

CPU restatement of the reference's per-batch metrics, /root/reference/problem.py:44-64 (`ProblemMetrics`): sklearn micro /
macro F1 on argmax predictions (`classification`), on `preds > 0` (`multilabel_classification`), and the mean absolute error
(`regression_mae`).  sklearn is the third-party dependency the reference itself calls (unpinned there; scikit-learn 1.9 here),
so the restatement IS those calls; `f1_counts` restates sklearn's published definition in numpy and is pinned against it
in tests/test_oracle_metrics.py.
"""
import warnings

import numpy as np


def classification(y_true, y_pred):
    from sklearn import metrics
    y_pred = np.argmax(y_pred, axis=1)
    return {"micro": float(metrics.f1_score(y_true, y_pred, average="micro")),
            "macro": float(metrics.f1_score(y_true, y_pred, average="macro"))}


def multilabel_classification(y_true, y_pred):
    from sklearn import metrics
    y_pred = (y_pred > 0).astype(int)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')                 # UndefinedMetricWarning for empty labels (F1 := 0)
        return {"micro": float(metrics.f1_score(y_true, y_pred, average="micro")),
                "macro": float(metrics.f1_score(y_true, y_pred, average="macro"))}


def regression_mae(y_true, y_pred):
    return float(np.abs(y_true - y_pred).mean())


def f1_counts(tp, fp, fn, present_only):
    """micro / macro F1 from per-label counts -- the definition the device kernel implements."""
    tp, fp, fn = [np.asarray(a, dtype=np.float64) for a in (tp, fp, fn)]
    den = 2 * tp + fp + fn
    f1 = np.where(den > 0, 2 * tp / np.maximum(den, 1), 0.0)
    labels = den > 0 if present_only else np.ones_like(den, dtype=bool)
    micro = 2 * tp.sum() / max(2 * tp.sum() + fp.sum() + fn.sum(), 1e-300)
    return {"micro": float(micro), "macro": float(f1[labels].mean()) if labels.any() else 0.0}
