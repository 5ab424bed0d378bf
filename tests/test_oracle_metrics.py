"""CPU: the metric restatement (oracle/metrics.py) against sklearn itself -- the dependency the reference calls at
problem.py:44-58 -- including the label-set conventions the device kernel has to reproduce."""
import numpy as np

from oracle import metrics


def _counts(y_true, y_pred, C):
    tp = np.array([np.sum((y_true == c) & (y_pred == c)) for c in range(C)])
    fp = np.array([np.sum((y_true != c) & (y_pred == c)) for c in range(C)])
    fn = np.array([np.sum((y_true == c) & (y_pred != c)) for c in range(C)])
    return tp, fp, fn


def test_classification_counts_definition_matches_sklearn():
    rs = np.random.RandomState(0)
    for C, n in ((5, 64), (41, 512), (7, 3)):
        logits = rs.randn(n, C).astype(np.float32)
        y = rs.randint(0, C - 1, size=(n, 1))                       # class C-1 never true: present only if predicted
        want = metrics.classification(y, logits)
        got = metrics.f1_counts(*_counts(y.reshape(-1), logits.argmax(1), C), present_only=True)
        assert abs(got['micro'] - want['micro']) < 1e-12 and abs(got['macro'] - want['macro']) < 1e-12


def test_multilabel_counts_definition_matches_sklearn():
    rs = np.random.RandomState(1)
    n, L = 200, 12
    logits = rs.randn(n, L).astype(np.float32)
    logits[:, 3] = -1.0                                               # a label that is never predicted ...
    y = (rs.rand(n, L) < 0.3).astype(np.float32)
    y[:, 3] = 0                                                       # ... nor true: F1 0, still counted in macro
    want = metrics.multilabel_classification(y, logits)
    p = logits > 0
    tp, fp, fn = (p & (y > 0)).sum(0), (p & (y == 0)).sum(0), (~p & (y > 0)).sum(0)
    got = metrics.f1_counts(tp, fp, fn, present_only=False)
    assert abs(got['micro'] - want['micro']) < 1e-12 and abs(got['macro'] - want['macro']) < 1e-12
