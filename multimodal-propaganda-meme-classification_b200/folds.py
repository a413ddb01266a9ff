"""The participant script's fold driver on the engine: ``setup(k)`` and ``for k in [0..4]: setup(k)``
(example_scripts/Multimodal_example_task2C.py:50-192, :882-885), followed by the ensembling tail
(example_scripts/combine_preds.py).

What ``setup(k)`` does in the reference, in order, and where it is here:

    seed_everything()                                                    :57         loop_head.seed_everything
    read_data(train_file) / read_data(test_file)                         :94-112,141 data.read_data (or records passed in)
    StratifiedKFold(5, shuffle=True, random_state=42).split -> fold k    :117-131    loop_head.stratified_kfold
    label -> id map, compute_class_weight('balanced') (printed only)     :114,133-140 balanced_class_weights
    three datasets, three DataLoader(batch 16, shuffle=True)             :142-164    ``make_dataset`` + DataLoader
    model, sigmoid focal loss, Adam(get_params(lr)), 10 % linear warm-up :166-176    model_factory, SigmoidFocalLoss,
                                                                                     FusedAdam, optim.get_linear_...
    8 epochs of train() [mid-epoch + end-of-epoch test/val, best-F1 TSVs] then test/val + the two "ALL |" lines
                                                                         :179-192    loop_head.train / test

The script keeps everything in module globals; here the same values travel in a ``FoldRun`` that ``setup`` returns.
The model / dataset constructors are arguments because the script's own (AraBERT + RoBERTa + ResNet-18 pretrained
checkpoints, BLIP captions) need the network; ``b200mm.MultimodalClassifierHEAD`` / ``b200mm.data.MemeDataset`` are
the engine's equivalents.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Callable, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import ensemble, loop_head
from .data import L2ID
from .loop import SigmoidFocalLoss
from .optim import FusedAdam, get_linear_schedule_with_warmup
from .tsv import read_prob_tsv, write_label_tsv

N_SPLITS = 5          # :116
SPLIT_SEED = 42       # :117
BATCH_SIZE = 16       # :73
LEARNING_RATE = 1e-5  # :68
NUM_EPOCHS = 8        # :171
WARMUP_RATIO = 0.1    # :173


def balanced_class_weights(labels) -> np.ndarray:
    """sklearn.utils.class_weight.compute_class_weight('balanced', classes=np.unique(y), y=y) (:137-138):
    n_samples / (n_classes * bincount(y))."""
    y = np.asarray(labels)
    classes, counts = np.unique(y, return_counts=True)
    return len(y) / (len(classes) * counts.astype(np.float64))


def _subset(records: dict, idx) -> dict:
    return {k: [v[i] for i in idx] for k, v in records.items()}


def _ids(records: dict) -> dict:
    out = dict(records)
    if "label" in out:
        out["label"] = [L2ID[l] if isinstance(l, str) else int(l) for l in out["label"]]
    return out


@dataclasses.dataclass
class FoldRun:
    """The globals ``setup(k)`` leaves behind in the reference (:51-55), as one object."""
    fold: int
    model: object
    criterion: object
    optimizer: object
    scheduler: object
    train_loader: DataLoader
    val_loader: DataLoader
    test_loader: DataLoader
    class_weights: np.ndarray
    total_steps: int
    warmup_steps: int
    best_macro_f1: float = 0.0
    history: list = dataclasses.field(default_factory=list)
    prob_tsv: str | None = None
    label_tsv: str | None = None


def setup(k: int, *, train_records: dict, test_records: dict, model_factory: Callable[[], object],
          make_dataset: Callable[[dict], object], device, batch_size: int = BATCH_SIZE,
          learning_rate: float = LEARNING_RATE, num_epochs: int = NUM_EPOCHS, n_splits: int = N_SPLITS,
          out_dir: str = ".", team_name: str = "kevinmathew", run_id: str | None = None, num_workers: int = 0,
          collate_fn=None, image_transform=None, log=print) -> FoldRun:
    """One fold of the reference's ``setup(k)`` (Multimodal_example_task2C.py:50-192).

    train_records / test_records: ``data.read_data`` dicts (``id``, ``text``, ``image``, ``label`` lists).
    model_factory():              a fresh model for this fold (the script builds ``MultimodalClassifier(fusion_method)``).
    make_dataset(records):        a Dataset yielding the reference's batch-dict keys for those records.
    image_transform:              data.GpuImageTransform for loaders that ship uint8 pixels (data.collate_packed);
                                  ``GpuImageTransform('square', train=True, augment=True)`` is the script's transform
                                  (:222-235), which its train, validation and test datasets all share."""
    loop_head.seed_everything()
    labels = train_records["label"]
    splits = list(loop_head.stratified_kfold(labels, n_splits, SPLIT_SEED))
    train_idx, val_idx = splits[k]
    train_r = _ids(_subset(train_records, train_idx))
    val_r = _ids(_subset(train_records, val_idx))
    test_r = _ids(test_records)
    class_weights = balanced_class_weights(train_r["label"])
    log(f"class weights: {class_weights}")
    train_ds, val_ds, test_ds = make_dataset(train_r), make_dataset(val_r), make_dataset(test_r)
    log(f"train_df len: {len(train_ds)}")
    log(f"val_df len: {len(val_ds)}")
    log(f"test_df len: {len(test_ds)}")
    kw = dict(batch_size=batch_size, shuffle=True, drop_last=False, num_workers=num_workers, collate_fn=collate_fn)
    train_loader, val_loader, test_loader = DataLoader(train_ds, **kw), DataLoader(val_ds, **kw), DataLoader(test_ds, **kw)

    model = model_factory()
    criterion = SigmoidFocalLoss(alpha=0.25, gamma=2.0)
    optimizer = FusedAdam(loop_head.get_params(model, learning_rate), lr=learning_rate,
                          max_grad_norm=loop_head.CLIP_NORM)
    total_steps = len(train_loader) * num_epochs
    warmup_steps = int(WARMUP_RATIO * total_steps)
    scheduler = get_linear_schedule_with_warmup(optimizer, num_warmup_steps=warmup_steps,
                                                num_training_steps=total_steps)
    run = FoldRun(k, model, criterion, optimizer, scheduler, train_loader, val_loader, test_loader, class_weights,
                  total_steps, warmup_steps)
    state = {}
    ev = {"fold": k, "out_dir": out_dir, "team_name": team_name, "run_id": run_id}
    for epoch in range(num_epochs):
        train_loss, acc = loop_head.train(model, train_loader, criterion, optimizer, scheduler, device, epoch,
                                          test_loader=test_loader, val_loader=val_loader, state=state,
                                          evaluate_kwargs=ev, log=log, image_transform=image_transform)
        t_loss, t_acc, t_f1, t_thr = loop_head.test(model, test_loader, criterion, device, epoch, log,
                                                    image_transform=image_transform)
        v_loss, v_acc, v_f1, v_thr = loop_head.test(model, val_loader, criterion, device, epoch, log,
                                                    image_transform=image_transform)
        log("  ALL | Epoch {}/{}: Train Loss = {:.4f}, Test Loss = {:.4f}, Train Accuracy = {:.4f}, Test Accuracy = "
            "{:.4f}, F1 = {:.4f}".format(epoch + 1, num_epochs, train_loss, t_loss, acc, t_acc, t_f1))
        log("  ALL | Epoch {}/{}: Train Loss = {:.4f}, Val Loss = {:.4f}, Train Accuracy = {:.4f}, Val Accuracy = "
            "{:.4f}, F1 = {:.4f}".format(epoch + 1, num_epochs, train_loss, v_loss, acc, v_acc, v_f1))
        run.history.append({"epoch": epoch, "train_loss": train_loss, "train_acc": acc, "test_loss": t_loss,
                            "test_acc": t_acc, "test_f1": t_f1, "test_threshold": t_thr, "val_loss": v_loss,
                            "val_acc": v_acc, "val_f1": v_f1})
    run.best_macro_f1 = state.get("best_macro_f1", 0.0)
    label_tsv = os.path.join(out_dir, f"task2C_{team_name}.tsv")
    prob_tsv = os.path.join(out_dir, f"task2C_{team_name}_probs_fold_{k}.tsv")
    run.label_tsv = label_tsv if os.path.exists(label_tsv) else None
    run.prob_tsv = prob_tsv if os.path.exists(prob_tsv) else None
    return run


def run_folds(folds: Sequence[int] = (0, 1, 2, 3, 4), *, log=print, **setup_kwargs):
    """``for k in [0, 1, 2, 3, 4]: setup(k=k)`` (:882-885).  Returns the FoldRuns (models are dropped between folds
    unless ``keep_models=True`` is passed: five resident models are only needed for ensemble inference)."""
    keep = setup_kwargs.pop("keep_models", False)
    runs = []
    for k in folds:
        log(f"training for fold: {k}")
        run = setup(k, log=log, **setup_kwargs)
        if not keep:
            run.model = run.optimizer = run.scheduler = None
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        runs.append(run)
    return runs


def combine_folds(prob_tsvs: Sequence[str], gold: dict | None = None, *, out_path: str | None = None,
                  run_id: str = "ensemble", log=print):
    """The tail after the fold loop (combine_preds.py:9-31, 66-88): read the per-fold probability TSVs, average per id,
    and -- when gold labels are given -- pick the F1-optimal threshold on the 100-point grid; otherwise threshold at
    0.5.  Writes the 3-column submission TSV when ``out_path`` is set.  Returns (ids, mean_prob, labels, threshold, f1)."""
    fold_ids, fold_probs = [], []
    for p in prob_tsvs:
        ids, _, probs, _ = read_prob_tsv(p)
        fold_ids.append(ids)
        fold_probs.append(probs)
    ids, mean_prob = ensemble.average_probability(fold_ids, fold_probs)
    if gold is not None:
        thr, f1, labels = ensemble.threshold_optimization(ids, mean_prob, gold)
        log(f"Optimal Threshold: {thr}")
        log(f"Optimal F1: {f1}")
    else:
        thr, f1 = 0.5, None
        labels = ["propaganda" if p > thr else "not_propaganda" for p in mean_prob]
    if out_path is not None:
        write_label_tsv(out_path, ids, labels, run_id)
    return ids, mean_prob, labels, thr, f1
