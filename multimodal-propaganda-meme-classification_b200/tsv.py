"""Prediction-file writers and readers with the reference's exact schemas.

  * 3-column label file  ``id\\tlabel\\trun_id``       example_scripts/Multimodal_example_task2C.txt:271-280,
                                                      Multimodal_example_task2C.py:857-867; accepted by
                                                      format_checker/task2.py:20,31 and scorer/task2.py:50
  * 4-column prob file   ``id\\tlabel\\tprob\\trun_id`` Multimodal_example_task2C.py:869-879 (prob printed with
                                                      Python's float repr of the fp32 value widened to double,
                                                      e.g. ``0.37347766757011414``)
"""
from __future__ import annotations

import re

LABELS = ("not_propaganda", "propaganda")
_LINE_RE = re.compile(r"^([\w:]+\/.*?\.[\w:]+)\t(propaganda|not_propaganda)\t[\w-]+")  # format_checker/task2.py:20


def write_label_tsv(path, ids, labels, run_id):
    with open(path, "w") as f:
        f.write("id\tlabel\trun_id\n")
        for i, l in zip(ids, labels):
            f.write(f"{i}\t{l}\t{run_id}\n")


def write_prob_tsv(path, ids, labels, probs, run_id):
    with open(path, "w") as f:
        f.write("id\tlabel\tprob\trun_id\n")
        for i, l, p in zip(ids, labels, probs):
            f.write(f"{i}\t{l}\t{float(p)!r}\t{run_id}\n")


def read_prob_tsv(path):
    """-> (ids, labels, probs as float64 list, run_ids)"""
    ids, labels, probs, runs = [], [], [], []
    with open(path) as f:
        header = f.readline().rstrip("\n").split("\t")
        if header != ["id", "label", "prob", "run_id"]:
            raise ValueError(f"{path}: unexpected header {header}")
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            i, l, p, r = line.split("\t")
            ids.append(i)
            labels.append(l)
            probs.append(float(p))
            runs.append(r)
    return ids, labels, probs, runs


def check_label_tsv(path) -> bool:
    """Same acceptance rule as the organisers' format checker (format_checker/task2.py:25-39)."""
    with open(path) as f:
        lines = f.read().split("\n")
    for line in lines[1:]:
        if not line:
            continue
        if not _LINE_RE.match(line) or len(line.split("\t")) != 3:
            return False
    return True
