"""Fold ensembling tail (BASELINE config 5): example_scripts/combine_preds.py restated without pandas/sklearn.

  average_probability   :29-31   concat folds -> group by id -> mean prob (ids returned sorted, as groupby does)
  threshold_optimization:34-63   100-point grid on [0, 1], maximise *binary* F1 of (prob > t), first arg-max wins
  majority_voting       :21-26
"""
from __future__ import annotations

import numpy as np


def average_probability(fold_ids, fold_probs):
    """fold_ids / fold_probs: one sequence per fold. Returns (sorted unique ids, float64 mean prob per id)."""
    acc: dict[str, list] = {}
    for ids, probs in zip(fold_ids, fold_probs):
        for i, p in zip(ids, probs):
            e = acc.setdefault(i, [0.0, 0])
            e[0] += float(p)
            e[1] += 1
    ids = sorted(acc)
    return ids, np.array([acc[i][0] / acc[i][1] for i in ids], dtype=np.float64)


def binary_f1(y_true, y_pred) -> float:
    """sklearn.metrics.f1_score(y_true, y_pred) for binary labels (positive class = 1; 0.0 when undefined)."""
    y_true = np.asarray(y_true).astype(bool)
    y_pred = np.asarray(y_pred).astype(bool)
    tp = int(np.sum(y_true & y_pred))
    fp = int(np.sum(~y_true & y_pred))
    fn = int(np.sum(y_true & ~y_pred))
    denom = 2 * tp + fp + fn
    return 0.0 if denom == 0 else 2.0 * tp / denom


def macro_f1(y_true, y_pred) -> float:
    """Official task metric (scorer/task2.py:109): unweighted mean of the per-class F1 over both classes."""
    y_true = np.asarray(y_true).astype(int)
    y_pred = np.asarray(y_pred).astype(int)
    return 0.5 * (binary_f1(y_true == 1, y_pred == 1) + binary_f1(y_true == 0, y_pred == 0))


def find_optimal_threshold(y_true, y_prob):
    """combine_preds.py:35-47 -> (threshold, best binary F1)."""
    thresholds = np.linspace(0, 1, 100)
    y_prob = np.asarray(y_prob, dtype=np.float64)
    scores = [binary_f1(y_true, y_prob > t) for t in thresholds]
    k = int(np.argmax(scores))
    return float(thresholds[k]), float(scores[k])


def threshold_optimization(ids, probs, gold: dict):
    """gold: id -> 'propaganda' | 'not_propaganda'. Returns (threshold, f1, labels per id)."""
    y_true = np.array([1 if gold[i] == "propaganda" else 0 for i in ids])
    t, f1 = find_optimal_threshold(y_true, probs)
    labels = ["propaganda" if p > t else "not_propaganda" for p in probs]
    return t, f1, labels


def majority_voting(fold_probs):
    votes = np.stack([np.asarray(p) > 0.5 for p in fold_probs]).sum(0)
    # pandas .mode(axis=1)[0] picks the smallest label on ties ('not_propaganda' < 'propaganda')
    return ["propaganda" if 2 * v > len(fold_probs) else "not_propaganda" for v in votes]


def roc_optimal_threshold(y_true, y_prob):
    """Multimodal_example_task2C.py:819-822: threshold = thresholds[argmax(tpr - fpr)] of sklearn's roc_curve
    (drop_intermediate=True changes only which collinear points are kept, not the maximiser's value)."""
    y_true = np.asarray(y_true).astype(int)
    y_prob = np.asarray(y_prob, dtype=np.float64)
    order = np.argsort(-y_prob, kind="mergesort")
    ys, ps = y_true[order], y_prob[order]
    distinct = np.where(np.diff(ps))[0]
    idx = np.r_[distinct, ys.size - 1]
    tps = np.cumsum(ys)[idx]
    fps = 1 + idx - tps
    P, N = max(int(ys.sum()), 1), max(int((1 - ys).sum()), 1)
    j = tps / P - fps / N
    k = int(np.argmax(j))
    if j[k] <= 0:            # roc_curve prepends (0,0) at threshold inf; it wins only if nothing beats chance
        return float("inf")
    return float(ps[idx][k])
