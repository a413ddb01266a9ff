"""Golden vectors from the REFERENCE'S OWN CODE, run in the build container (needs /root/reference):

    python tests/golden/make_reference_golden.py      ->  tests/golden/reference_run_golden.pt

What runs is the reference's source text, executed verbatim:

  organiser script  example_scripts/Multimodal_example_task2C.txt
      class MultimodalDataset (:28-72), class MultimodalClassifier (:152-197), def train (:200-223), def test (:225-242),
      def evaluate (:259-280)
  participant script example_scripts/Multimodal_example_task2C.py
      class MultimodalDataset (:206-304), LLMWithClassificationHead (:307-392), ConcatAttention3 (:476-499), CustomDenseNet161 (:562-585),
      MultimodalClassifier incl. get_params (:587-685), def train (:689-776), test (:779-834), evaluate (:837-879)

  SVM baseline       baselines/subtask_2c.py
      def run_imgbert_baseline (:74-95), fed with feature files written by this repo's write_features_json

with the names they look up at run time bound to the stock libraries -- except the three network-bound constructors,
which return from-config modules (tests/golden/refpin.py).  Weights are set by name (refpin.reseed_by_name), inputs are
refpin.batches: both are reproducible without the reference, so tests/test_cpu.py can demand that the oracle
(oracle/reference_model.py) and the repo's host-side loops reproduce these vectors.  This pins the oracle to the
reference's module code on the library versions of this image (transformers 5.5 / torchvision 0.26 / torch 2.11; the
reference's lock file names 4.39.2 / 0.17.2 / 2.2.2).
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/example_scripts"
sys.path.insert(0, HERE)
import refpin  # noqa: E402


def _top_level_block(lines, start_prefix):
    """Source lines of the top-level class / def starting with ``start_prefix`` (the organiser file is a notebook dump
    that does not parse as a whole)."""
    i0 = next(i for i, l in enumerate(lines) if l.startswith(start_prefix))
    i1 = i0 + 1
    while i1 < len(lines) and (lines[i1].strip() == "" or lines[i1][0] in " \t"):
        i1 += 1
    return "".join(lines[i0:i1]), (i0 + 1, i1)


def organiser_namespace():
    lines = open(os.path.join(REF, "Multimodal_example_task2C.txt"), encoding="utf-8").readlines()
    ns = {"torch": torch, "nn": nn, "optim": optim, "tqdm": lambda it: it, "text_model_name": "stub",
          "AutoModel": types.SimpleNamespace(from_pretrained=lambda name: refpin.distilbert()),
          "models": types.SimpleNamespace(resnet50=lambda pretrained=True: refpin.resnet50_small())}
    spans = {}
    for prefix in ("class MultimodalClassifier", "def train(", "def test(", "def evaluate("):
        src, span = _top_level_block(lines, prefix)
        exec(compile(src, f"Multimodal_example_task2C.txt:{span[0]}", "exec"), ns)
        spans[prefix] = span
    return ns, spans


def participant_namespace(workdir):
    from sklearn.metrics import f1_score, roc_curve
    from torchvision.ops import sigmoid_focal_loss
    path = os.path.join(REF, "Multimodal_example_task2C.py")
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    archs = {"stub-text": "bert", "stub-caption": "roberta"}
    ns = {"torch": torch, "nn": nn, "np": np, "tqdm": lambda it: it, "roc_curve": roc_curve, "f1_score": f1_score,
          "AutoModel": types.SimpleNamespace(from_pretrained=lambda name: refpin.bert(archs[name])),
          "timm": types.SimpleNamespace(create_model=lambda name, pretrained=True: refpin.timm_resnet18()),
          "text_model": "stub-text", "english_text_model": "stub-caption", "image_model": "resnet18",
          "fusion_method": "concatenation", "USE_FP16": False, "fold": 3, "best_macro_f1": 0.0,
          "sigmoid_focal_loss": sigmoid_focal_loss}
    want = {"LLMWithClassificationHead", "ConcatAttention3", "CustomDenseNet161", "MultimodalClassifier", "train", "test",
            "evaluate"}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want:
            exec(compile(ast.get_source_segment(src, node), f"Multimodal_example_task2C.py:{node.lineno}", "exec"), ns)
    return ns


def run_organiser():
    ns, spans = organiser_namespace()
    torch.manual_seed(0)
    model = refpin.reseed_by_name(ns["MultimodalClassifier"](2), seed=1)
    keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    data = refpin.batches(4, 4, seed=11, captions=False)
    # (A) one dropout-free training-mode forward / backward (the regime of every engine parity test)
    from oracle.reference_model import zero_dropout
    zero_dropout(model).train()
    b = data[0]
    out = model(b["text"], b["image"], b["text_mask"])
    loss = nn.CrossEntropyLoss()(out, b["label"])
    loss.backward()
    fx = {"spans": spans, "state_keys": keys, "logits": out.detach().clone(), "loss": loss.detach().clone(),
          "grad_norms": refpin.param_norms(model, grads=True),
          "grad_output_fc": model.output_fc.weight.grad.clone(), "grad_fusion_bias": model.fusion_fc.bias.grad.clone()}
    # (B) the reference's own train() / test() (.txt:200-242) with its dropout (0.3 head, 0.1 towers) drawing from torch's
    #     generator, Adam(lr=2e-5) and CrossEntropyLoss as at .txt:248-249
    model = refpin.reseed_by_name(ns["MultimodalClassifier"](2), seed=1)
    opt = optim.Adam(model.parameters(), lr=2e-5)
    torch.manual_seed(123)
    tr = ns["train"](model, refpin.ListLoader(data[:3]), nn.CrossEntropyLoss(), opt, torch.device("cpu"))
    te = ns["test"](model, refpin.ListLoader(data[3:]), nn.CrossEntropyLoss(), torch.device("cpu"))
    fx.update(train_return=tr, test_return=te, post_train_norms=refpin.param_norms(model))
    # the script's evaluate() (.txt:259-280) writes task2C_TeamName.tsv into the working directory
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ns["evaluate"](model, refpin.ListLoader(data[2:]), torch.device("cpu"))
            fx["tsv_label"] = open("task2C_TeamName.tsv").read()
        finally:
            os.chdir(cwd)
    return fx


def run_dataset():
    """The organiser script's own ``MultimodalDataset`` (.txt:28-72), executed verbatim on three JPEG files: what
    ``__getitem__`` hands the loop -- token ids / mask at the script's max length 512, label, and the image after the
    script's Resize(256) / CenterCrop(224) / ToTensor / Normalize on the PIL image.  The files themselves travel in the
    fixture (their encoder's output is not guaranteed to be reproducible elsewhere); the image tensors are stored as the
    uint8 pixel values they were computed from ((x * std + mean) * 255 is integral up to rounding noise)."""
    from PIL import Image
    from torch.utils.data import DataLoader, Dataset
    from torchvision import transforms
    lines = open(os.path.join(REF, "Multimodal_example_task2C.txt"), encoding="utf-8").readlines()
    with tempfile.TemporaryDirectory() as tmp:
        tok = refpin.EncodePlusTokenizer(tmp)
        ns = {"torch": torch, "Dataset": Dataset, "Image": Image, "transforms": transforms, "train_max_seq_len": 512,
              "AutoTokenizer": types.SimpleNamespace(from_pretrained=lambda name: tok)}
        src, span = _top_level_block(lines, "class MultimodalDataset")
        exec(compile(src, f"Multimodal_example_task2C.txt:{span[0]}", "exec"), ns)
        files = refpin.dataset_jpegs()
        paths = []
        for i, f in enumerate(files):
            paths.append(os.path.join(tmp, f"img_{i}.jpg"))
            open(paths[-1], "wb").write(f)
        ids = [f"data/x/img_{i}.jpg" for i in range(len(files))]
        ds = ns["MultimodalDataset"](ids, refpin.DATASET_TEXTS, paths, refpin.DATASET_LABELS)
        batch = next(iter(DataLoader(ds, batch_size=len(files), shuffle=False)))
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 3, 1, 1)
    px = (batch["image"] * std + mean) * 255.0
    assert (px - px.round()).abs().max() < 1e-3
    return {"span": span, "files": files, "id": list(batch["id"]), "text": batch["text"], "text_mask": batch["text_mask"],
            "label": batch["label"], "image_u8": px.round().to(torch.uint8), "keys": sorted(batch.keys())}


def run_participant_dataset():
    """The participant script's ``MultimodalDataset`` (.py:206-304), executed verbatim: its Compose -- Resize((224, 224)),
    RandomHorizontalFlip, ColorJitter(.1, .1, .1, .1), RandomRotation(15), ToTensor, Normalize -- on the PIL image, with
    torch's generator seeded before every ``__getitem__`` so that the draws can be replayed without the reference
    (torch.rand(1) for the flip, ColorJitter.get_params, RandomRotation.get_params, in that order); BLIP captioning
    (:195-204, CUDA + a download) is a stub that returns fixed strings."""
    from PIL import Image
    from torch.utils.data import Dataset
    from torchvision import transforms
    path = os.path.join(REF, "Multimodal_example_task2C.py")
    src = open(path, encoding="utf-8").read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "MultimodalDataset")

    class ImageCaptioning:
        def generate_caption(self, images, texts):
            return [f"{t} the propaganda" for t in texts]

    with tempfile.TemporaryDirectory() as tmp:
        tok = refpin.EncodePlusTokenizer(tmp)
        ns = {"torch": torch, "Dataset": Dataset, "Image": Image, "transforms": transforms, "tqdm": lambda it: it,
              "train_max_seq_len": 512, "text_model": "stub-text", "english_text_model": "stub-caption",
              "AutoTokenizer": types.SimpleNamespace(from_pretrained=lambda name: tok),
              "ImageCaptioning": ImageCaptioning}
        exec(compile(ast.get_source_segment(src, node), f"Multimodal_example_task2C.py:{node.lineno}", "exec"), ns)
        files = refpin.dataset_jpegs()[:2]
        paths = []
        for i, f in enumerate(files):
            paths.append(os.path.join(tmp, f"img_{i}.jpg"))
            open(paths[-1], "wb").write(f)
        ids = [f"data/x/img_{i}.jpg" for i in range(len(files))]
        ds = ns["MultimodalDataset"](ids, refpin.DATASET_TEXTS[:2], paths, refpin.DATASET_LABELS[:2])
        items = []
        for i in range(len(files)):
            torch.manual_seed(refpin.DATASET_AUG_SEED + i)
            items.append(ds[i])
    mean = torch.tensor((0.485, 0.456, 0.406)).view(3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(3, 1, 1)
    px = torch.stack([(it["image"] * std + mean) * 255.0 for it in items])
    assert (px - px.round()).abs().max() < 1e-3
    return {"line": node.lineno, "keys": sorted(items[0].keys()), "captions": list(ds.precalculated_captions),
            "text": torch.stack([it["text"] for it in items]), "text_mask": torch.stack([it["text_mask"] for it in items]),
            "caption_text": torch.stack([it["caption_text"] for it in items]),
            "caption_text_mask": torch.stack([it["caption_text_mask"] for it in items]),
            "label": torch.stack([it["label"] for it in items]), "image_u8": px.round().to(torch.uint8)}


def run_svm_consumer():
    """The consumer of the feature-extraction contract, ``run_imgbert_baseline`` (baselines/subtask_2c.py:74-95), executed
    verbatim on feature files written by THIS repo's ``write_features_json``: the results TSV it produces."""
    import json
    from os.path import join
    from sklearn.svm import SVC
    sys.path.insert(0, ROOT)
    from b200mm.features import write_features_json
    path = "/root/reference/baselines/subtask_2c.py"
    src = open(path, encoding="utf-8").read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "run_imgbert_baseline")
    ns = {"json": json, "join": join, "np": np, "SVC": SVC}
    exec(compile(ast.get_source_segment(src, node), f"subtask_2c.py:{node.lineno}", "exec"), ns)
    corpus = refpin.svm_feature_corpus()
    with tempfile.TemporaryDirectory() as tmp:
        for split in ("train", "dev"):
            c = corpus[split]
            write_features_json(os.path.join(tmp, "features", f"{split}_feats.json"), c["imgfeats"], c["textfeats"])
            json.dump([{"id": i, "class_label": l} for i, l in zip(c["id"], c["class_label"])],
                      open(os.path.join(tmp, f"{split}.json"), "w"))
        out = os.path.join(tmp, "results.tsv")
        ns["run_imgbert_baseline"](tmp, "dev", "train.json", "dev.json", out)
        return {"line": node.lineno, "results_tsv": open(out).read()}


def run_participant():
    from transformers import get_linear_schedule_with_warmup
    with tempfile.TemporaryDirectory() as tmp:
        ns = participant_namespace(tmp)
        torch.manual_seed(0)
        model = refpin.reseed_by_name(ns["MultimodalClassifier"]("concatenation"), seed=2)
        keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        groups = [sorted(n for n, p in model.named_parameters() if any(p is q for q in g["params"]))
                  for g in model.get_params(1e-5)]
        lrs = [g["lr"] for g in model.get_params(1e-5)]
        data = refpin.batches(6, 6, seed=21, captions=True)
        # (A) dropout-free training-mode forward / backward with the script's loss (:167, :711)
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
        model.train()
        b = data[0]
        out = model(b["text"], b["image"], b["text_mask"], b["caption_text"], b["caption_text_mask"])
        loss = ns["sigmoid_focal_loss"](out, b["label"].float(), alpha=0.25, gamma=2.0, reduction="mean")
        loss.backward()
        fx = {"state_keys": keys, "param_groups": groups, "group_lrs": lrs, "logits": out.detach().clone(),
              "loss": loss.detach().clone(), "grad_norms": refpin.param_norms(model, grads=True)}
        # (B) the script's own train() (:689-776) -- which calls its test() and evaluate() mid-epoch on the globals
        #     test_df / val_df -- over 4 batches, then test(); dropout active, Adam over get_params, linear warm-up
        model = refpin.reseed_by_name(ns["MultimodalClassifier"]("concatenation"), seed=2)
        opt = optim.Adam(model.get_params(1e-4))
        sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=1, num_training_steps=8)
        ns["test_df"], ns["val_df"] = refpin.ListLoader(data[4:5]), refpin.ListLoader(data[5:6])
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            torch.manual_seed(321)
            tr = ns["train"](model, refpin.ListLoader(data[:4]), ns["sigmoid_focal_loss"], opt, sched,
                             torch.device("cpu"), 0, None)
            te = ns["test"](model, ns["test_df"], ns["sigmoid_focal_loss"], torch.device("cpu"), 0)
            tsv_label = open("task2C_kevinmathew.tsv").read()
            tsv_prob = open("task2C_kevinmathew_probs_fold_3.tsv").read()
        finally:
            os.chdir(cwd)
        fx.update(train_return=tuple(float(v) for v in tr), test_return=tuple(float(v) for v in te),
                  best_macro_f1=float(ns["best_macro_f1"]), tsv_label=tsv_label, tsv_prob=tsv_prob,
                  post_train_norms=refpin.param_norms(model), training_flag_after_train=bool(model.training))
        return fx


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    fx = {"organiser": run_organiser(), "participant": run_participant(), "dataset": run_dataset(),
          "participant_dataset": run_participant_dataset(), "svm_consumer": run_svm_consumer(),
          "versions": {"torch": torch.__version__, "transformers": __import__("transformers").__version__,
                       "torchvision": __import__("torchvision").__version__}}
    out = os.path.join(HERE, "reference_run_golden.pt")
    torch.save(fx, out)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("organiser train/test:", fx["organiser"]["train_return"], fx["organiser"]["test_return"])
    print("participant train/test:", fx["participant"]["train_return"], fx["participant"]["test_return"],
          "model.training after train():", fx["participant"]["training_flag_after_train"])
