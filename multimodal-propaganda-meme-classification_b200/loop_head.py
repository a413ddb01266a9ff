"""The participant's ("HEAD") training-loop API of example_scripts/Multimodal_example_task2C.py, on the engine:

    seed_everything(seed)                                                   :42-48
    stratified_kfold(labels, 5, 42)      == StratifiedKFold(5, shuffle=True, random_state=42).split   :115-128
    get_params(model, lr)                param groups: head @ lr, text tower @ 0.8 lr, image tower @ 0.8 lr  :645-664
    get_linear_schedule_with_warmup      (b200mm.optim)                     :169-174
    train(model, train_loader, criterion, optimizer, scheduler, device, epoch, scaler=None)   :689-776
    test(model, test_loader, criterion, device, epoch) -> (loss, acc, macro_f1, roc_threshold) :779-834
    evaluate(model, test_loader, threshold, device)  -> task2C_<team>.tsv + task2C_<team>_probs_fold_<k>.tsv  :837-879

with a single-logit head + sigmoid focal loss (``MultimodalClassifier(num_classes=1, squeeze_output=True)`` and
``SigmoidFocalLoss``; :167, :641-643) on the two-tower model, or with the script's own three-tower model
(``b200mm.MultimodalClassifierHEAD``: caption tower, BatchNorm1d projection heads, ConcatAttention3; batches then
also carry ``caption_text`` / ``caption_text_mask``, :293-303).

What is deliberately NOT reproduced: clipping before un-scaling under AMP (:712-717) --
the engine is bf16 with fp32 master weights and needs no loss scaling, so ``scaler`` is accepted and ignored; the
fp32 branch's ``clip_grad_norm_(model.parameters(), 10.0)`` (:728-730) is ENFORCED by ``train`` on every step: a
``FusedAdam`` without a clip threshold gets ``max_grad_norm = 10.0`` (the clip is folded into the Adam kernel), any
other optimizer is preceded by ``torch.nn.utils.clip_grad_norm_``; and the script's accidental switch to eval mode
after its first mid-epoch check (``train(reference_eval_mode_quirk=True)`` reproduces it).
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import ensemble
from .loop import ID2L, SigmoidFocalLoss, _extra_inputs, _fused, _to_device
from .tsv import write_label_tsv, write_prob_tsv


CLIP_NORM = 10.0   # clip_grad_norm_(model.parameters(), 10.0), Multimodal_example_task2C.py:713-715, 728-730


def seed_everything(seed: int = 42) -> None:
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def stratified_kfold(labels, n_splits: int = 5, seed: int = 42):
    """sklearn.model_selection.StratifiedKFold(n_splits, shuffle=True, random_state=seed).split(X, labels), restated
    (sklearn/model_selection/_split.py, _make_test_folds): classes are encoded in order of first appearance, every
    class's samples are dealt round-robin to the folds and the per-class fold assignment is shuffled with ONE shared
    RandomState.  Yields (train_idx, val_idx) per fold."""
    y = np.asarray(labels)
    _, y_idx, y_inv = np.unique(y, return_index=True, return_inverse=True)
    _, class_perm = np.unique(y_idx, return_inverse=True)
    y_enc = class_perm[y_inv]
    n_classes = len(y_idx)
    y_order = np.sort(y_enc)
    allocation = np.asarray([np.bincount(y_order[i::n_splits], minlength=n_classes) for i in range(n_splits)])
    rng = np.random.RandomState(seed)
    test_folds = np.empty(len(y), dtype="i")
    for k in range(n_classes):
        folds_for_class = np.arange(n_splits).repeat(allocation[:, k])
        rng.shuffle(folds_for_class)
        test_folds[y_enc == k] = folds_for_class
    idx = np.arange(len(y))
    for f in range(n_splits):
        yield idx[test_folds != f], idx[test_folds == f]


def param_group_index(name: str) -> int:
    """The script's substring rules (:650-658), in its order: 'fusion_layer' -> group 0 (lr); 'text_model' -> group 1
    (0.8 lr) -- which also captures ``caption_text_model.*``; 'image_model' -> group 2 (0.8 lr); anything else -> 0."""
    return 0 if "fusion_layer" in name else 1 if "text_model" in name else 2 if "image_model" in name else 0


def get_params(model, lr: float):
    """Three param groups as in the HEAD script (:645-664): everything else @ lr, text tower(s) @ 0.8 lr, image tower
    @ 0.8 lr.  A model with the script's attribute names (``text_model`` / ``image_model`` / ``fusion_layer``) is grouped
    by the script's own substring rules; the two-tower engine model by its tower prefixes (``bert.*`` / ``resnet.*``)."""
    if hasattr(model, "get_params"):      # the three-tower HEAD model carries the reference's own method
        return model.get_params(lr)
    named = list(model.named_parameters())
    groups = ([], [], [])
    if any("text_model" in n or "image_model" in n or "fusion_layer" in n for n, _ in named):
        for name, p in named:
            groups[param_group_index(name)].append(p)
    else:
        for name, p in named:
            groups[1 if name.startswith("bert.") else 2 if name.startswith("resnet.") else 0].append(p)
    return [{"params": groups[0], "lr": lr}, {"params": groups[1], "lr": lr * 0.8},
            {"params": groups[2], "lr": lr * 0.8}]


def _probs(output):
    return torch.sigmoid(output.float().reshape(-1))


def test(model, test_loader, criterion, device, epoch=0, log=print, *, image_transform=None):
    """``image_transform``: the data.GpuImageTransform applied to uint8 image batches (the script's Dataset uses ONE
    transform object -- augmentations included -- for its train, validation and test datasets alike, :222-235)."""
    model.eval()
    test_loss, n = 0.0, 0
    true_labels, predicted_probs = [], []
    fused = _fused(criterion) and hasattr(model, "eval_step_fused")
    with torch.no_grad():
        for batch_idx, data in enumerate(test_loader, 1):
            text, image, mask, labels = _to_device(data, device, image_transform)
            extra = _extra_inputs(data, device)
            if fused:
                output, loss, _ = model.eval_step_fused(text, image, mask, *extra, labels, loss_kind=criterion.loss_kind,
                                                        alpha=criterion.alpha, gamma=criterion.gamma)
                loss_v = loss.item()
            else:
                output = model(text, image, mask, *extra)
                loss_v = criterion(output, labels.float()).item()
            test_loss += loss_v * labels.size(0)
            n += labels.size(0)
            predicted_probs.extend(_probs(output).cpu().numpy())
            true_labels.extend(labels.float().cpu().numpy())
            if batch_idx % 10 == 0:
                log(f" TEST | Epoch [{epoch}] | Batch [{batch_idx}/{len(test_loader)}] | Loss: {loss_v:.4f} |")
    y = np.asarray(true_labels)
    p = np.asarray(predicted_probs, dtype=np.float64)
    optimal_threshold = ensemble.roc_optimal_threshold(y, p)           # roc_curve + argmax(tpr - fpr), :819-822
    log(f"Optimal Threshold: {optimal_threshold}")
    predicted = (p > optimal_threshold).astype(float)
    denom = len(test_loader.dataset) if hasattr(test_loader, "dataset") else n
    accuracy = float((predicted == y).sum()) / denom
    macro_f1 = ensemble.macro_f1(y, predicted)
    test_loss /= denom
    log(f" TEST | Epoch [{epoch}] | Testing Loss: {test_loss:.4f} | Accuracy: {accuracy:.4f} | "
        f"Macro F1: {macro_f1:.4f} | optim t: {optimal_threshold} |")
    return test_loss, accuracy, macro_f1, optimal_threshold


def evaluate(model, test_loader, t_optimal_threshold, device, *, fold=0, team_name="kevinmathew",
             run_id=None, out_dir=".", image_transform=None):
    """Writes ``task2C_<team>.tsv`` (id, label, run_id) and ``task2C_<team>_probs_fold_<fold>.tsv``
    (id, label, prob, run_id) -- the schemas of the reference's committed prediction files."""
    model.eval()
    ids, probs = [], []
    with torch.no_grad():
        for data in test_loader:
            text, image, mask, _ = _to_device(data, device, image_transform)
            probs.append(_probs(model(text, image, mask, *_extra_inputs(data, device))).cpu())
            ids.extend(list(data["id"]))
    probs = torch.cat(probs).numpy() if probs else np.zeros(0, dtype=np.float32)
    labels = [ID2L[int(p > t_optimal_threshold)] for p in probs]
    run_id = run_id or f"{team_name}_resnet50_distilbert-base-multilingual-cased_concatenation.tsv"
    f1 = os.path.join(out_dir, f"task2C_{team_name}.tsv")
    f2 = os.path.join(out_dir, f"task2C_{team_name}_probs_fold_{fold}.tsv")
    write_label_tsv(f1, ids, labels, run_id)
    write_prob_tsv(f2, ids, labels, probs, run_id)
    return f1, f2


def train(model, train_loader, criterion, optimizer, scheduler, device, epoch=0, scaler=None, *, test_loader=None,
          val_loader=None, state=None, evaluate_kwargs=None, log=print, image_transform=None,
          reference_eval_mode_quirk: bool = False):
    """One epoch. ``state`` (dict) carries ``best_macro_f1`` across epochs like the script's global (:766-769).
    ``image_transform``: data.GpuImageTransform for uint8 image batches -- ``GpuImageTransform('square', train=True,
    augment=True)`` is the script's transform (:222-235: Resize((224, 224)), flip, ColorJitter, RandomRotation,
    Normalize), run on the device; it is also handed to the mid-epoch ``test`` / ``evaluate`` calls, as the script's
    datasets all share that transform.

    ``reference_eval_mode_quirk``: the script's ``train`` never calls ``model.train()`` again after its mid-epoch
    ``test`` / ``evaluate`` calls (:752-769 -> ``model.eval()`` at :780 / :838), so from the first check of every epoch
    (batch ``total // 2``) to the epoch's end it keeps optimising in EVAL mode -- dropout off, BatchNorm normalising
    with its running statistics.  By default this loop restores training mode after a check; ``True`` reproduces the
    script to the letter (tests/test_cpu.py replays the script's own functions against it, tests/golden/
    make_reference_golden.py), for a model whose backward supports eval-mode BatchNorm (any torch module; not the
    fused engine step)."""
    model.train()
    state = state if state is not None else {}
    train_loss, correct, n = 0.0, 0, 0
    total_batches = len(train_loader)
    check_interval = max(total_batches // 2, 1)
    batch_losses = []
    fused = _fused(criterion) and hasattr(model, "train_step_fused")
    fused_optim = hasattr(optimizer, "max_grad_norm")
    if fused and reference_eval_mode_quirk:
        raise ValueError("reference_eval_mode_quirk needs the generic (torch criterion) route: the fused engine step "
                         "differentiates training-mode BatchNorm only")
    if fused_optim and optimizer.max_grad_norm is None:
        optimizer.max_grad_norm = CLIP_NORM              # reference: clip_grad_norm_(..., 10.0) on every step
    for batch_idx, data in enumerate(train_loader, 1):
        optimizer.zero_grad()
        text, image, mask, labels = _to_device(data, device, image_transform)
        extra = _extra_inputs(data, device)
        if fused:
            output, loss, ok = model.train_step_fused(text, image, mask, *extra, labels, loss_kind=criterion.loss_kind,
                                                      alpha=criterion.alpha, gamma=criterion.gamma)
            if not fused_optim:
                grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), CLIP_NORM)
            optimizer.step()
            loss_v, ok_v = loss.item(), ok.item()
        else:
            output = model(text, image, mask, *extra)
            loss = criterion(output, labels.float() if output.dim() == 1 else labels)
            loss.backward()
            if not fused_optim:
                grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), CLIP_NORM)     # :728-730
            optimizer.step()
            loss_v = loss.item()
            pred = (_probs(output) > 0.5).float() if output.dim() == 1 or output.shape[-1] == 1 else output.argmax(1)
            ok_v = (pred == labels).sum().item()
        scheduler.step()
        bs = labels.size(0)
        train_loss += loss_v * bs
        batch_losses.append(loss_v)
        correct += ok_v
        n += bs
        if batch_idx % 10 == 0:
            gn = getattr(optimizer, "last_grad_norm", None)
            gn = float(gn.sqrt().item()) if gn is not None else (float(grad_norm) if not fused_optim else float("nan"))
            log(f"TRAIN | Epoch [{epoch}] | Batch [{batch_idx}/{total_batches}] | "
                f"Loss: {sum(batch_losses) / len(batch_losses):.4f} | LR: {scheduler.get_last_lr()[0]} | "
                f"Grad Norm: {gn:.4f} |")
            batch_losses = []
        if test_loader is not None and (batch_idx % check_interval == 0 or batch_idx == total_batches):
            t_loss, t_acc, t_f1, t_thr = test(model, test_loader, criterion, device, epoch, log,
                                              image_transform=image_transform)
            if val_loader is not None:
                v_loss, v_acc, v_f1, v_thr = test(model, val_loader, criterion, device, epoch, log,
                                                  image_transform=image_transform)
                log(f" VAL | Epoch [{epoch}] | Batch [{batch_idx}/{total_batches}] | Test Loss: {v_loss:.4f} | "
                    f"Acc: {v_acc:.4f} | F1: {v_f1:.4f} | thresh: {v_thr}")
            log(f" TEST | Epoch [{epoch}] | Batch [{batch_idx}/{total_batches}] | Test Loss: {t_loss:.4f} | "
                f"Acc: {t_acc:.4f} | F1: {t_f1:.4f} | thresh: {t_thr}")
            if t_f1 > state.get("best_macro_f1", 0.0):
                state["best_macro_f1"] = t_f1
                evaluate(model, test_loader, t_thr, device, image_transform=image_transform, **(evaluate_kwargs or {}))
            if not reference_eval_mode_quirk:
                model.train()
    denom = len(train_loader.dataset) if hasattr(train_loader, "dataset") else n
    train_loss /= denom
    accuracy = correct / denom
    log(f"TRAIN | Epoch [{epoch}] | Training Loss: {train_loss:.4f} | Accuracy: {accuracy:.4f} |")
    return train_loss, accuracy
