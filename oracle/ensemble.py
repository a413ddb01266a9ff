"""TEST INFRASTRUCTURE ONLY -- pandas/sklearn restatement of the reference's fold-ensembling script,
example_scripts/combine_preds.py (functions at :7-9, :21-26, :29-31, :34-63), kept as close to the original
expression-by-expression as possible so that it can be pinned against the script's own printed output
(tests/golden/combine_preds_golden.json, produced by running the unmodified script; see tests/golden/make_golden.py).
"""
import numpy as np
import pandas as pd
from sklearn.metrics import f1_score


def majority_voting(dfs):                                            # combine_preds.py:21-26
    binary_predictions = [df['prob'].apply(lambda x: 'propaganda' if x > 0.5 else 'not_propaganda') for df in dfs]
    majority_vote = pd.concat(binary_predictions, axis=1).mode(axis=1)[0]
    result = dfs[0][['id']].copy()
    result['label'] = majority_vote
    return result


def average_probability(dfs):                                        # combine_preds.py:29-31
    return pd.concat([df[['id', 'prob']] for df in dfs]).groupby('id').mean().reset_index()


def find_optimal_threshold(y_true, y_prob):                          # combine_preds.py:35-47
    thresholds = np.linspace(0, 1, 100)
    f1_scores = [f1_score(y_true, y_prob > t) for t in thresholds]
    return thresholds[np.argmax(f1_scores)], f1_scores[np.argmax(f1_scores)]


def threshold_optimization(df, labels_df):                           # combine_preds.py:34-63
    merged_df = pd.merge(df, labels_df, on='id', how='left')
    y_true = merged_df['class_label'].apply(lambda x: 1 if x == 'propaganda' else 0).values
    y_prob = merged_df['prob'].values
    optimal_threshold, best = find_optimal_threshold(y_true, y_prob)
    merged_df['label'] = merged_df['prob'].apply(lambda x: 'propaganda' if x > optimal_threshold else 'not_propaganda')
    return merged_df[['id', 'prob', 'label']], optimal_threshold, best
