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


_EVAL_TRANSFORM = None


def _image_to_device(data, device, image_transform=None):
    """fp32 pixels are copied as they are; uint8 pixels (fixed-size or packed, data.collate_packed) take the GPU
    transform: ``image_transform`` (a data.GpuImageTransform) or, by default, the organiser script's deterministic
    evaluation form (Resize(256) / CenterCrop(224) / Normalize)."""
    global _EVAL_TRANSFORM
    if "jpeg_table" in data:           # entropy-decoded JPEG batch (jpeg.collate_jpeg): finish the decode on the device
        from .jpeg import reconstruct_batch
        data = dict(data)
        data["image_packed"], data["image_table"] = reconstruct_batch(data, device)
    if "image_packed" in data or data["image"].dtype == torch.uint8:
        tr = image_transform
        if tr is None:
            if _EVAL_TRANSFORM is None:
                from .data import GpuImageTransform
                _EVAL_TRANSFORM = GpuImageTransform("center_crop")
            tr = _EVAL_TRANSFORM
        if "image_packed" in data:
            return tr.packed(data["image_packed"].to(device, non_blocking=True),
                             data["image_table"].to(device, non_blocking=True))
        return tr.fixed(data["image"].to(device, non_blocking=True))
    return data["image"].to(device, non_blocking=True)


def _to_device(data, device, image_transform=None):
    text = data["text"].to(device, non_blocking=True)
    image = _image_to_device(data, device, image_transform)
    mask = data["text_mask"].to(device, non_blocking=True)
    labels = data["label"].to(device, non_blocking=True) if "label" in data else None
    return text, image, mask, labels


def _extra_inputs(data, device):
    """The HEAD script's batches also carry the BLIP caption tokens (Multimodal_example_task2C.py:293-303, 706-708):
    returned as extra positional inputs of the model when present."""
    if "caption_text" not in data:
        return ()
    return (data["caption_text"].to(device, non_blocking=True), data["caption_text_mask"].to(device, non_blocking=True))


class DevicePrefetcher:
    """Iterates a loader of the reference's batch dicts and yields (text, image, mask, labels, batch) already on the
    device, copying batch i+1 on a side stream while step i computes (the reference does the copy synchronously at
    the top of every step, .txt:206-211; its loop is input-bound).  The overlap needs pinned host
    tensors (``DataLoader(pin_memory=True)`` / ``data.collate_packed``); pageable ones are copied the way the
    reference copies them.

    ``image`` may arrive in four forms (the fourth: ``jpeg_coefs`` / ``jpeg_qtabs`` / ``jpeg_table`` from
    jpeg.collate_jpeg -- files entropy-decoded in the loader worker, reconstructed to pixels here on the device): fp32 [B, 3, H, W] (what the reference's CPU transform produces -- copied as
    is), uint8 [B, H, W, 3] at network resolution, or a packed variable-size batch (``image_packed`` +
    ``image_table``, data.collate_packed).  The uint8 forms cross PCIe at a quarter of the bytes and are turned into
    the normalised fp32 tensor by ONE kernel on the copy stream (``image_transform``, a data.GpuImageTransform;
    default: the organiser script's Resize(256) / CenterCrop(224) / Normalize)."""

    KEYS = ("text", "image", "text_mask", "label", "caption_text", "caption_text_mask", "image_packed", "image_table")

    def __init__(self, loader, device, pin: bool = False, image_transform=None):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.stream = torch.cuda.Stream(device=self.device)
        self.image_transform = image_transform

    def _transform(self):
        if self.image_transform is None:
            from .data import GpuImageTransform
            self.image_transform = GpuImageTransform("center_crop")
        return self.image_transform

    def _stage(self, data):
        out = {}
        with torch.cuda.stream(self.stream):
            for k in self.KEYS:
                if k in data:
                    t = data[k]
                    if self.pin and not t.is_cuda and not t.is_pinned():
                        t = t.pin_memory()
                    out[k] = t.to(self.device, non_blocking=True)
            if "jpeg_table" in data:   # coefficients cross PCIe; IDCT / up-sampling / colour conversion run here
                from .jpeg import reconstruct_batch
                out["image_packed"], out["image_table"] = reconstruct_batch(data, self.device)
            if "image_packed" in out:
                out["image"] = self._transform().packed(out.pop("image_packed"), out.pop("image_table"))
            elif "image" in out and out["image"].dtype == torch.uint8:
                out["image"] = self._transform().fixed(out["image"])
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev, data

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev, raw = nxt
            try:
                nxt = self._stage(next(it))      # issue the next copy before this step's kernels are queued
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in cur.values():
                t.record_stream(torch.cuda.current_stream(self.device))
            if "caption_text" in cur:            # HEAD-script batches: hand the device copies to _extra_inputs
                raw = dict(raw, caption_text=cur["caption_text"], caption_text_mask=cur["caption_text_mask"])
            yield cur["text"], cur["image"], cur["text_mask"], cur.get("label"), raw

    def __len__(self):
        return len(self.loader)


class StepReadback:
    """Per-step loss / correct-count readback without draining the launch queue.  The reference reads ``loss.item()``
    right after ``optimizer.step()`` (.txt:218-220), which blocks the host until the whole step has run and leaves the
    GPU idle while the next step's first kernels are being launched (about 1 ms of a 33 ms step here).  ``push`` queues
    an asynchronous copy of this step's two scalars into pinned host memory and hands back the PREVIOUS step's values
    (already complete by then); ``flush`` returns the last ones.  Every step's result still crosses to the host."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.host = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.dev = [torch.zeros(2, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.pending = None
        self.i = 0

    def _take(self):
        if self.pending is None:
            return None
        slot = self.pending
        self.events[slot].synchronize()
        self.pending = None
        return float(self.host[slot][0]), int(round(float(self.host[slot][1])))

    def push(self, loss, ok):
        prev = self._take()
        slot = self.i & 1
        self.i += 1
        self.dev[slot][0].copy_(loss.reshape(()), non_blocking=True)
        self.dev[slot][1].copy_(ok.reshape(()), non_blocking=True)
        self.host[slot].copy_(self.dev[slot], non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        self.pending = slot
        return prev

    def flush(self):
        return self._take()

    bytes_per_step = 8


def _fused(criterion):
    return isinstance(criterion, (CrossEntropyLoss, SigmoidFocalLoss))


def train(model, train_loader, criterion, optimizer, device, scheduler=None, on_step=None, image_transform=None,
          graph=None):
    """One epoch (.txt:200-223).  ``graph``: a ``b200mm.GraphedTrainStep`` built over (model, optimizer, criterion) --
    the loop body then runs as one CUDA-graph replay per step (the small-batch regime, where launching the step's
    ~500 kernels is what bounds it); batches whose shape differs from the captured one take the eager step."""
    model.train()
    train_loss = 0.0
    correct = 0
    n = 0
    fused = _fused(criterion) and hasattr(model, "train_step_fused")
    readback = StepReadback(device) if fused else None
    prev_bs = 0
    if graph is not None and not fused:
        raise ValueError("graph= needs the fused criterion (b200mm.CrossEntropyLoss / SigmoidFocalLoss)")

    def account(done, bs):
        nonlocal train_loss, correct
        train_loss += done[0] * bs
        correct += done[1]
        if on_step is not None:
            on_step(done[0], bs)

    for text, image, mask, labels, data in DevicePrefetcher(train_loader, device, image_transform=image_transform):
        if graph is not None:
            _, loss, ok = graph(text, image, mask, labels)
            if scheduler is not None:
                scheduler.step()
            done = readback.push(loss, ok)
            if done is not None:
                account(done, prev_bs)
            prev_bs = labels.size(0)
            n += prev_bs
            continue
        optimizer.zero_grad()
        if fused:
            _, loss, ok = model.train_step_fused(text, image, mask, labels, loss_kind=criterion.loss_kind,
                                                 alpha=criterion.alpha, gamma=criterion.gamma)
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            # the reference's per-step loss / accuracy reads (.txt:218-220), one step late so the host never waits
            done = readback.push(loss, ok)
            if done is not None:
                account(done, prev_bs)
            prev_bs = labels.size(0)
            n += prev_bs
            continue
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
    if readback is not None:
        done = readback.flush()
        if done is not None:
            account(done, prev_bs)
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


def get_features(model, loader, device):
    """Feature extraction for the SVM baseline (baselines/extract_feat.py:52-67 ``get_features``):
    eval-mode towers -> one image feature vector and one text feature vector per id, returned as the two
    ``{id: [float, ...]}`` dicts that script dumps to JSON for baselines/subtask_2c.py:74-95.  The towers are the
    engine's own (image: the tower's output vector -- 1000-d ResNet logits, or the ViT CLS feature; text: the pooled
    token the classifier head reads), not the script's ConvNeXt / AraBERT pooler."""
    model.eval()
    img_feats, text_feats = {}, {}
    with torch.no_grad():
        for text, image, mask, _, data in DevicePrefetcher(loader, device):
            B, S = text.shape
            model._prep() if hasattr(model, "_prep") else model.store.refresh_shadow()
            h = model.text.forward(text, mask, training=False)
            off = S - 1 if getattr(model, "pooling", "cls") == "last" else 0
            t = ops.gather_rows(h, B, S, off).float().cpu()
            v = model.img.forward(image, training=False).float().cpu()
            for k, i in enumerate(data["id"]):
                img_feats[i] = v[k].tolist()
                text_feats[i] = t[k].tolist()
    return img_feats, text_feats
