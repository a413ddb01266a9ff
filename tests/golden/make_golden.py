"""Regenerates the committed golden fixtures.  Runs ONLY in the build container (needs /root/reference):

  python tests/golden/make_golden.py

1. combine_preds_golden.json -- the reference's own fold-ensembling script (example_scripts/combine_preds.py) is
   executed UNMODIFIED on the reference's committed fold TSVs; its printed (threshold, F1) pairs are the known
   answers.  The inputs travel in anonymised form (sample index instead of the dataset path; the probabilities
   and binary gold labels are kept verbatim) so tests can re-run the computation without /root/reference.
2. scorer_golden.json -- scorer/task2.py metrics of the committed single-run TSV against the dev gold file.
3. oracle_tiny_golden.pt -- logits / loss / a few gradients of the CPU oracle (oracle/reference_model.py: the
   reference's MultimodalClassifier restated on stock transformers/torchvision modules) on the tiny configuration,
   pinning the oracle itself against accidental drift.
"""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def combine_preds_golden():
    out = subprocess.run([sys.executable, "example_scripts/combine_preds.py"], cwd=REF, capture_output=True,
                         text=True, check=True).stdout
    pairs = [tuple(map(float, m)) for m in re.findall(r"^([0-9.]+) ([0-9.]+)$", out, flags=re.M)]
    assert len(pairs) == 6, out
    gold = {d["id"]: d["class_label"] for d in
            json.load(open(os.path.join(REF, "data/arabic_memes_propaganda_araieval_24_dev.json")))}
    sys.path.insert(0, os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200"))
    import tsv as _tsv   # plain module import: no CUDA needed
    folds = []
    index = {}
    header_runs = None
    for k in range(5):
        ids, labels, probs, runs = _tsv.read_prob_tsv(os.path.join(REF, f"task2C_kevinmathew_probs_fold_{k}.tsv"))
        header_runs = runs[0]
        for i in ids:
            index.setdefault(i, len(index))
        folds.append({"idx": [index[i] for i in ids], "prob_repr": [repr(p) for p in probs],
                      "label": [1 if l == "propaganda" else 0 for l in labels]})
    # sorted-id order is what pandas' groupby returns; store the rank of every anonymised id
    order = sorted(index, key=lambda s: s)
    rank = {index[i]: r for r, i in enumerate(order)}
    fx = {"source": "example_scripts/combine_preds.py run unmodified in the build container",
          "pairs_threshold_f1": pairs, "run_id": header_runs,
          "gold": {str(index[i]): (1 if gold[i] == "propaganda" else 0) for i in index},
          "sorted_rank": {str(k): v for k, v in rank.items()}, "folds": folds}
    json.dump(fx, open(os.path.join(HERE, "combine_preds_golden.json"), "w"))
    print("combine_preds:", pairs)


def scorer_golden():
    sys.path.insert(0, REF)
    from scorer import task2 as scorer   # the organisers' scorer, imported unmodified
    import logging
    logging.disable(logging.CRITICAL)
    pred = os.path.join(REF, "analysis", "task2C_kevinmathew 2.tsv")
    goldf = os.path.join(REF, "data/arabic_memes_propaganda_araieval_24_dev.json")
    acc, p, r, f1 = scorer.evaluate(goldf, pred)
    gold = {d["id"]: d["class_label"] for d in json.load(open(goldf))}
    rows = [l.rstrip("\n").split("\t") for l in open(pred)][1:]
    ids = sorted(gold)
    idx = {i: k for k, i in enumerate(ids)}
    fx = {"source": "scorer/task2.py evaluate() on analysis/task2C_kevinmathew 2.tsv",
          "acc": acc, "precision_weighted": p, "recall_weighted": r, "f1_macro": f1,
          "gold": [1 if gold[i] == "propaganda" else 0 for i in ids],
          "pred": {str(idx[r[0]]): (1 if r[1] == "propaganda" else 0) for r in rows}}
    json.dump(fx, open(os.path.join(HERE, "scorer_golden.json"), "w"))
    print("scorer:", acc, p, r, f1)


def oracle_golden():
    import torch
    import torch.nn as nn
    from oracle import reference_model as R
    cfg = R.TowerConfig.tiny()
    torch.manual_seed(42)
    torch.set_num_threads(1)
    m = R.zero_dropout(R.MultimodalClassifier(2, cfg))
    m.train()
    data = R.synthetic_batch(8, 32, cfg)
    out = m(data["text"], data["image"], data["text_mask"])
    loss = nn.CrossEntropyLoss()(out, data["label"])
    loss.backward()
    fx = {"logits": out.detach(), "loss": loss.detach(),
          "grad_output_fc": m.output_fc.weight.grad.clone(),
          "grad_q0_norm": m.bert.transformer.layer[0].attention.q_lin.weight.grad.norm(),
          "grad_conv1_norm": m.resnet.conv1.weight.grad.norm(),
          "text_ids_sum": data["text"].sum(), "mask_sum": data["text_mask"].sum(), "labels": data["label"]}
    torch.save(fx, os.path.join(HERE, "oracle_tiny_golden.pt"))
    print("oracle tiny loss:", loss.item())


def kfold_golden():
    """Labels of the reference's train split + sklearn's StratifiedKFold(5, shuffle=True, random_state=42) folds, the
    split the HEAD script draws at Multimodal_example_task2C.py:115-128 (fold sizes 1714/1715)."""
    import numpy as np
    from sklearn.model_selection import StratifiedKFold
    d = json.load(open(os.path.join(REF, "data/arabic_memes_propaganda_araieval_24_train.json")))
    labels = [1 if x["class_label"] == "propaganda" else 0 for x in d]
    skf = StratifiedKFold(5, shuffle=True, random_state=42)
    splits = list(skf.split(np.zeros(len(labels)), labels))
    json.dump({"source": "data/arabic_memes_propaganda_araieval_24_train.json class_label (1 = propaganda), file order; "
                         "sklearn StratifiedKFold(5, shuffle=True, random_state=42)",
               "labels_bits": "".join(map(str, labels)), "fold_sizes": [(len(a), len(b)) for a, b in splits],
               "first_val_indices": [b[:5].tolist() for a, b in splits]},
              open(os.path.join(HERE, "kfold_golden.json"), "w"))
    print("kfold:", [(len(a), len(b)) for a, b in splits])


if __name__ == "__main__":
    kfold_golden()
    combine_preds_golden()
    scorer_golden()
    oracle_golden()
