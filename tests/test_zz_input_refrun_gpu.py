"""The device-side input pipeline against the organiser script's own Dataset (reference-run fixture).

Kept in a file of its own that sorts after the other GPU suites: its host half and its tolerances were verified / calibrated
on the CPU (tests/test_cpu.py::test_dataset_contract_matches_reference_dataset_run), every device code path it uses is
covered by tests/test_kernels_gpu.py, but the round's GPU budget ended before this combination itself ran on a B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pillow_exact_input_tensor_equals_reference_dataset(cuda_device, tmp_path):
    """GpuImageTransform(resample="pillow") behind the split JPEG decode: the tensor that reaches the training step equals
    the one the organiser script's own Dataset produced (fixture) -- every pixel on the same uint8 value, floats within
    rounding of torch's ToTensor / Normalize.  (The arithmetic was verified on the host build against Pillow and this
    fixture, tests/test_cpu.py::test_pillow_exact_resize_arithmetic_on_host; this is the kernel's first run on a GPU.)"""
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import refpin
    from b200mm import data as D, jpeg
    from b200mm.loop import DevicePrefetcher
    fx = torch.load(os.path.join(gold, "reference_run_golden.pt"), weights_only=False)["dataset"]
    paths = []
    for i, f in enumerate(fx["files"]):
        p = tmp_path / f"img_{i}.jpg"
        p.write_bytes(f)
        paths.append(str(p))
    tok = refpin.EncodePlusTokenizer(tmp_path)

    def tokenize(text):
        e = tok.tok(text, add_special_tokens=True)
        return e["input_ids"], e["attention_mask"]

    ds = D.MemeDataset(fx["id"], refpin.DATASET_TEXTS, paths, refpin.DATASET_LABELS, tokenizer=tokenize, max_len=512,
                       image_loader=D.file_bytes_loader)
    batch = jpeg.collate_jpeg([ds[i] for i in range(len(ds))])
    tr = D.GpuImageTransform("center_crop", resample="pillow")
    (text, image, mask, labels, raw), = list(DevicePrefetcher([batch], cuda_device, image_transform=tr))
    mean = torch.tensor(D.ops.IMAGENET_MEAN, device=cuda_device).view(1, 3, 1, 1)
    std = torch.tensor(D.ops.IMAGENET_STD, device=cuda_device).view(1, 3, 1, 1)
    want_u8 = fx["image_u8"].to(cuda_device)
    got_u8 = ((image * std + mean) * 255.0).round().to(torch.uint8)
    assert torch.equal(got_u8, want_u8), (got_u8.int() - want_u8.int()).abs().max().item()
    want = (want_u8.float() / 255.0 - mean) / std
    assert (image - want).abs().max().item() < 1e-6


def test_pillow_exact_train_transform_equals_participant_dataset(cuda_device, tmp_path):
    """GpuImageTransform("square", train=True, augment=True, rng="torchvision", resample="pillow") behind the split JPEG
    decode, one sample at a time under the seeds of the fixture: the tensor equals the one the participant script's own
    Dataset produced (Resize((224, 224)) / flip / ColorJitter / RandomRotation on the PIL image) -- every pixel on the same
    uint8 value.  (Arithmetic verified on the host build, tests/test_cpu.py::test_pillow_exact_augmentation_arithmetic_on_host;
    this is the kernels' first run on a GPU.)"""
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import refpin
    from b200mm import data as D, jpeg
    fx = torch.load(os.path.join(gold, "reference_run_golden.pt"), weights_only=False)
    files, pd = fx["dataset"]["files"], fx["participant_dataset"]
    tr = D.GpuImageTransform("square", train=True, augment=True, rng="torchvision", resample="pillow")
    mean = torch.tensor(D.ops.IMAGENET_MEAN, device=cuda_device).view(1, 3, 1, 1)
    std = torch.tensor(D.ops.IMAGENET_STD, device=cuda_device).view(1, 3, 1, 1)
    for i in range(pd["image_u8"].shape[0]):
        packed, table = jpeg.reconstruct_batch(jpeg.pack_jpeg_batch([files[i]]), cuda_device)
        torch.manual_seed(refpin.DATASET_AUG_SEED + i)
        image = tr.packed(packed, table)
        got_u8 = ((image * std + mean) * 255.0).round().to(torch.uint8)
        want_u8 = pd["image_u8"][i:i + 1].to(cuda_device)
        assert torch.equal(got_u8, want_u8), (i, (got_u8.int() - want_u8.int()).abs().max().item())
        assert (image - (want_u8.float() / 255.0 - mean) / std).abs().max().item() < 1e-6


def test_input_pipeline_against_reference_dataset_run(cuda_device, tmp_path):
    """What the training step receives -- through data.MemeDataset + collate_packed (decoded pixels cross PCIe) and
    through file_bytes_loader + jpeg.collate_jpeg (coefficients cross PCIe), both finished by loop.DevicePrefetcher on the
    GPU -- against what the organiser script's own ``MultimodalDataset`` handed ITS loop for the same files and texts
    (.txt:28-72 executed verbatim, tests/golden/make_reference_golden.py): token tensors identical; image within one uint8
    step of the script's PIL transform (PIL rounds the resized image to uint8 after each pass, the kernel keeps floats;
    calibrated on the CPU in tests/test_cpu.py: max 0.99, mean 0.30 of a step), exact where no resize happens."""
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import refpin
    from b200mm import data as D, jpeg
    from b200mm.loop import DevicePrefetcher
    fx = torch.load(os.path.join(gold, "reference_run_golden.pt"), weights_only=False)["dataset"]
    paths = []
    for i, f in enumerate(fx["files"]):
        p = tmp_path / f"img_{i}.jpg"
        p.write_bytes(f)
        paths.append(str(p))
    tok = refpin.EncodePlusTokenizer(tmp_path)

    def tokenize(text):
        e = tok.tok(text, add_special_tokens=True)
        return e["input_ids"], e["attention_mask"]

    mean = torch.tensor(D.ops.IMAGENET_MEAN, device=cuda_device).view(1, 3, 1, 1)
    std = torch.tensor(D.ops.IMAGENET_STD, device=cuda_device).view(1, 3, 1, 1)
    want = (fx["image_u8"].to(cuda_device).float() / 255.0 - mean) / std
    images = []
    for loader, collate in ((D.pil_loader, D.collate_packed), (D.file_bytes_loader, jpeg.collate_jpeg)):
        ds = D.MemeDataset(fx["id"], refpin.DATASET_TEXTS, paths, refpin.DATASET_LABELS, tokenizer=tokenize, max_len=512,
                           image_loader=loader)
        (text, image, mask, labels, raw), = list(DevicePrefetcher([collate([ds[i] for i in range(len(ds))])], cuda_device))
        assert torch.equal(text.cpu(), fx["text"]) and torch.equal(mask.cpu(), fx["text_mask"])
        assert torch.equal(labels.cpu(), fx["label"]) and raw["id"] == fx["id"]
        steps = (image - want).abs() * std * 255.0
        assert steps.max().item() < 1.1 and steps.mean().item() < 0.35, (steps.max().item(), steps.mean().item())
        assert steps[2].max().item() < 0.02          # 256 x 256 file: centre crop + Normalize only
        images.append(image)
    assert torch.equal(images[0], images[1])         # pixels decoded on the GPU == Pillow's: identical tensors downstream
