"""Training / evaluation loops with the reference's signatures and return values.

    train(model, train_loader, criterion, optimizer, device) -> (mean loss, accuracy)   .txt:200-223
    test(model, test_loader, criterion, device)              -> (mean loss, accuracy)   .txt:225-242
    evaluate(model, test_loader, device)                     -> writes task2C_<team>.tsv .txt:259-280

(example_scripts/Multimodal_example_task2C.txt).  Batches are the reference's dicts
``{"id", "text", "text_mask", "image", "label"}`` (.txt:61-69); they are moved to the device here exactly where the
reference does (``data[...].to(device)``, .txt:206-211), from pinned memory when the loader provides it.

When ``criterion`` is one of this package's loss objects the step runs fully fused on the device
(output layer + loss + backward in one kernel, see ``MultimodalClassifier.train_step_fused``); with any other
criterion the generic ``loss = criterion(model(...), labels); loss.backward()`` path is used, as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .tsv import write_label_tsv

ID2L = {0: "not_propaganda", 1: "propaganda"}


class CrossEntropyLoss:
    """nn.CrossEntropyLoss() stand-in (mean reduction) that lets the loops take the fused head+loss kernel."""
    loss_kind = ops.LOSS_CE
    alpha, gamma = 0.25, 2.0

    def __call__(self, output, labels):
        return F.cross_entropy(output, labels)


class SigmoidFocalLoss:
    """torchvision.ops.sigmoid_focal_loss(alpha=.25, gamma=2, reduction='mean') stand-in
    (Multimodal_example_task2C.py:167); for single-logit heads."""
    loss_kind = ops.LOSS_FOCAL

    def __init__(self, alpha=0.25, gamma=2.0):
        self.alpha, self.gamma = alpha, gamma

    def __call__(self, output, labels):
        from torchvision.ops import sigmoid_focal_loss
        return sigmoid_focal_loss(output.squeeze(-1), labels.float(), alpha=self.alpha, gamma=self.gamma,
                                  reduction="mean")


def _to_device(data, device):
    text = data["text"].to(device, non_blocking=True)
    image = data["image"].to(device, non_blocking=True)
    mask = data["text_mask"].to(device, non_blocking=True)
    labels = data["label"].to(device, non_blocking=True) if "label" in data else None
    return text, image, mask, labels


def _fused(criterion):
    return isinstance(criterion, (CrossEntropyLoss, SigmoidFocalLoss))


def train(model, train_loader, criterion, optimizer, device, scheduler=None, on_step=None):
    model.train()
    train_loss = 0.0
    correct = 0
    n = 0
    fused = _fused(criterion) and hasattr(model, "train_step_fused")
    for data in train_loader:
        optimizer.zero_grad()
        text, image, mask, labels = _to_device(data, device)
        if fused:
            _, loss, ok = model.train_step_fused(text, image, mask, labels, loss_kind=criterion.loss_kind,
                                                 alpha=criterion.alpha, gamma=criterion.gamma)
            optimizer.step()
            loss_v, ok_v = loss.item(), ok.item()     # the reference's two per-step D2H syncs (.txt:218-220)
        else:
            output = model(text, image, mask)
            loss = criterion(output, labels)
            loss.backward()
            optimizer.step()
            loss_v = loss.item()
            _, predicted = torch.max(output, 1)
            ok_v = (predicted == labels).sum().item()
        if scheduler is not None:
            scheduler.step()
        bs = labels.size(0)
        train_loss += loss_v * bs
        correct += ok_v
        n += bs
        if on_step is not None:
            on_step(loss_v, bs)
    denom = len(train_loader.dataset) if hasattr(train_loader, "dataset") else n
    return train_loss / denom, correct / denom


def test(model, test_loader, criterion, device):
    model.eval()
    test_loss = 0.0
    correct = 0
    n = 0
    fused = _fused(criterion) and hasattr(model, "eval_step_fused")
    with torch.no_grad():
        for data in test_loader:
            text, image, mask, labels = _to_device(data, device)
            if fused:
                _, loss, ok = model.eval_step_fused(text, image, mask, labels, loss_kind=criterion.loss_kind,
                                                    alpha=criterion.alpha, gamma=criterion.gamma)
                loss_v, ok_v = loss.item(), ok.item()
            else:
                output = model(text, image, mask)
                loss_v = criterion(output, labels).item()
                _, predicted = torch.max(output, 1)
                ok_v = (predicted == labels).sum().item()
            bs = labels.size(0)
            test_loss += loss_v * bs
            correct += ok_v
            n += bs
    denom = len(test_loader.dataset) if hasattr(test_loader, "dataset") else n
    return test_loss / denom, correct / denom


def predict(model, test_loader, device):
    """Eval-mode forward over a loader. Returns (ids, logits fp32 [N, C] on the host)."""
    model.eval()
    ids, outs = [], []
    with torch.no_grad():
        for data in test_loader:
            text, image, mask, _ = _to_device(data, device)
            outs.append(model(text, image, mask).float().cpu())
            ids.extend(list(data["id"]))
    return ids, torch.cat(outs) if outs else torch.empty(0)


def evaluate(model, test_loader, device, out_path="task2C_TeamName.tsv", run_id="DistilBERT+ResNet"):
    """Reference: .txt:259-280 -- argmax label per id, 3-column TSV ``id\\tlabel\\trun_id``."""
    ids, logits = predict(model, test_loader, device)
    labels = [ID2L[int(i)] for i in logits.argmax(1).tolist()]
    write_label_tsv(out_path, ids, labels, run_id)
    return list(zip(ids, labels))
