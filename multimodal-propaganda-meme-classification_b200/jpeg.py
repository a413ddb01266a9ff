"""JPEG decode split between loader workers and the GPU.

The reference decodes every image on the host, per sample and per epoch: ``Image.open(path).convert("RGB")``
(example_scripts/Multimodal_example_task2C.txt:50; Multimodal_example_task2C.py:270).  Here

  * the host (a DataLoader worker; the C calls release the GIL) parses the file and Huffman-decodes it into quantised
    DCT coefficients -- ``entropy_decode`` / ``collate_jpeg`` (csrc/jpeg_decode.cu, host half);
  * the device turns a whole batch of coefficient sets into packed uint8 RGB images in two launches --
    ``reconstruct_batch`` (dequantisation + inverse DCT, chroma up-sampling + colour conversion) -- which
    ``ops.preprocess_u8_packed`` resizes / crops / normalises.  Decoded pixels never exist on the host.

The pixels are bit-identical to Pillow's (libjpeg-turbo's default path is restated in integer arithmetic,
csrc/jpeg_math.cuh; tests/test_cpu.py and tests/test_kernels_gpu.py compare with ``Image.open``).  Files outside the
supported set (CMYK, arithmetic coding, 12-bit, RGB-coded, 4:4:0 / exotic sampling) raise ``UnsupportedJpeg`` from
``entropy_decode``.  By default a batch that contains such a file fails loudly; ``unsupported="pil"`` (opt-in) hands
those files -- and non-JPEG files such as PNGs -- to the loader the reference itself uses (Pillow) and ships their
pixels beside the coefficients, so that a real dataset with a stray CMYK or PNG meme still trains.
"""
from __future__ import annotations

import io
import threading

import numpy as np
import torch
import torch.utils.data

from . import _lib

INFO_INTS = 32
TABLE_COLS = 32
JPEG_UNSUPPORTED, JPEG_CORRUPT = -10, -11


class UnsupportedJpeg(ValueError):
    """A valid file the split decoder does not handle (decode it with the loader the reference uses)."""


class CorruptJpeg(OSError):
    """Not a JPEG, damaged or truncated (Pillow raises OSError for those as well)."""


def _as_bytes(data) -> np.ndarray:
    if isinstance(data, torch.Tensor):
        data = data.numpy()
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _check(rc: int, what: str):
    if rc == JPEG_UNSUPPORTED:
        raise UnsupportedJpeg(f"{what}: coding not handled by the split decoder")
    if rc == JPEG_CORRUPT:
        raise CorruptJpeg(f"{what}: not a JPEG, or damaged / truncated")
    if rc != 0:
        raise _lib.B200MMError(f"{what}: error {rc}")


def parse(data) -> np.ndarray:
    """Frame header of a JPEG file -> int32 [32] (include/b200mm.h: width, height, components, progressive, hs, vs,
    blocks per row [3], block rows [3], component width [3], height [3], first coefficient [3], coefficients, restart)."""
    a = _as_bytes(data)
    info = np.zeros(INFO_INTS, dtype=np.int32)
    _check(_lib.load().b200mm_jpeg_parse(a.ctypes.data, a.size, info.ctypes.data), "b200mm_jpeg_parse")
    return info


def entropy_decode(data, out: np.ndarray | None = None):
    """Huffman-decodes every scan of the file on the host.  Returns (coefs int16 [info[21]], qtabs uint16 [3, 64], info).
    ``out``: an int16 array to decode into (at least info[21] long), e.g. a slice of a pinned batch buffer."""
    a = _as_bytes(data)
    info = parse(a)
    n = int(info[21])
    coefs = np.empty(n, dtype=np.int16) if out is None else out[:n]
    if coefs.size != n or coefs.dtype != np.int16 or not coefs.flags.c_contiguous:
        raise ValueError("out must be a contiguous int16 array of at least info[21] elements")
    qtabs = np.zeros((3, 64), dtype=np.uint16)
    _check(_lib.load().b200mm_jpeg_entropy_decode(a.ctypes.data, a.size, coefs.ctypes.data, qtabs.ctypes.data,
                                                  info.ctypes.data), "b200mm_jpeg_entropy_decode")
    return coefs, qtabs, info


class _Scratch:
    """Per-thread reusable work space of ``entropy_decode_sparse`` (dense coefficients + worst-case entry arrays)."""

    def __init__(self):
        self.dense = np.empty(0, dtype=np.int16)
        self.idx = np.empty(0, dtype=np.uint8)
        self.val = np.empty(0, dtype=np.int16)

    def fit(self, n):
        if self.dense.size < n:
            self.dense, self.idx, self.val = (np.empty(n, dtype=np.int16), np.empty(n, dtype=np.uint8),
                                              np.empty(n, dtype=np.int16))


_TLS = threading.local()


def entropy_decode_sparse(data):
    """``entropy_decode`` with the result compacted to its non-zero coefficients (a typical file keeps 5-15 % of them).
    Returns (block_off int32 [blocks + 1], idx uint8 [nnz], val int16 [nnz], qtabs uint16 [3, 64], info): entries
    [block_off[b], block_off[b + 1]) belong to block b (blocks in the dense layout's order), ``idx`` = position inside
    the block in natural order."""
    a = _as_bytes(data)
    info = parse(a)
    n = int(info[21])
    sc = getattr(_TLS, "scratch", None)
    if sc is None:
        sc = _TLS.scratch = _Scratch()
    sc.fit(n)
    block_off = np.empty(n // 64 + 1, dtype=np.int32)
    qtabs = np.zeros((3, 64), dtype=np.uint16)
    nnz = np.zeros(1, dtype=np.int64)
    _check(_lib.load().b200mm_jpeg_entropy_decode_sparse(a.ctypes.data, a.size, sc.dense.ctypes.data,
                                                         block_off.ctypes.data, sc.idx.ctypes.data, sc.val.ctypes.data, n,
                                                         qtabs.ctypes.data, info.ctypes.data, nnz.ctypes.data),
           "b200mm_jpeg_entropy_decode_sparse")
    k = int(nnz[0])
    return block_off, sc.idx[:k].copy(), sc.val[:k].copy(), qtabs, info


def _pil_pixels(data) -> torch.Tensor:
    from PIL import Image
    with Image.open(io.BytesIO(_as_bytes(data).tobytes())) as im:
        return torch.from_numpy(np.asarray(im.convert("RGB"), dtype=np.uint8).copy())


class CoefRing:
    """Recycled (pinned) coefficient buffers for a collate that runs in the training process: a fresh 236 MB pinned
    allocation per 256-image batch costs ~130 ms, first-touch page faults of a fresh pageable one ~0.5 ms per image.
    ``slots`` buffers are handed out in turn, so a buffer is overwritten ``slots`` batches later -- keep fewer batches
    than that in flight (``DevicePrefetcher`` holds two).  Buffers grow to the largest batch seen."""

    def __init__(self, slots: int = 3, pin: bool = True):
        self.slots, self.pin, self._next = slots, pin, 0
        self._bufs = [None] * slots

    def take(self, n_coefs: int) -> torch.Tensor:
        i = self._next
        self._next = (i + 1) % self.slots
        if self._bufs[i] is None or self._bufs[i].numel() < n_coefs:
            self._bufs[i] = torch.empty(max(n_coefs, 8), dtype=torch.int16, pin_memory=self.pin)
        return self._bufs[i][:max(n_coefs, 8)]


def pack_jpeg_batch(files, pin: bool = True, unsupported: str = "raise", threads: int = 1,
                    ring: CoefRing | None = None, sparse: bool = False) -> dict:
    """files: the raw bytes of the batch's image files.  ``unsupported``: 'raise' (default) or 'pil' -- what to do with a
    file outside the split decoder's set (module docstring).  ``threads`` > 1 Huffman-decodes the files of the batch
    concurrently (the C call releases the GIL; every file writes its own slice of the batch buffer).  ``pin``: allocate
    the buffers in pinned memory here -- for a collate that runs in the training process; inside DataLoader WORKER
    processes pass ``pin=False`` and let ``DataLoader(pin_memory=True)`` pin (a worker must not create a CUDA context).
    ``ring``: a ``CoefRing`` whose recycled buffer receives the coefficients instead of a fresh allocation.
    ``sparse=True`` ships only the non-zero coefficients: ``jpeg_coefs`` is replaced by ``jpeg_sp_off`` int32
    [blocks of the batch + 1], ``jpeg_sp_idx`` uint8 [nnz], ``jpeg_sp_val`` int16 [nnz], and table columns 17..19 hold the
    first BLOCK of each component in the batch's block numbering (3 B per non-zero + 4 B per block cross PCIe instead of
    128 B per block).
    Returns the batch in the form that crosses PCIe:

      jpeg_coefs  int16 [total]       coefficients of every split-decoded image (pinned)
      jpeg_qtabs  int16 [n, 3, 64]    quantisation tables (uint16 bit patterns)
      jpeg_table  int64 [n, 32]       per-image geometry and offsets (include/b200mm.h: b200mm_jpeg_reconstruct)
      jpeg_meta   int64 [5]           max blocks / width / height over the batch, bytes of plane scratch, bytes of output
      jpeg_raw    list of (index, uint8 [H, W, 3])   with unsupported='pil': the images Pillow decoded
    """
    if unsupported not in ("raise", "pil"):
        raise ValueError("unsupported must be 'raise' or 'pil'")
    if sparse and ring is not None:
        raise ValueError("ring= recycles the DENSE coefficient buffer; the sparse form allocates exact-size buffers")
    n = len(files)
    infos, raw = [], []
    for i, f in enumerate(files):
        try:
            infos.append(parse(f))
        except (UnsupportedJpeg, CorruptJpeg):
            if unsupported == "raise":
                raise
            px = _pil_pixels(f)                          # raises for files Pillow cannot read either, as the reference does
            raw.append((i, px))
            info = np.zeros(INFO_INTS, dtype=np.int32)
            info[0], info[1] = px.shape[1], px.shape[0]
            infos.append(info)
    table = torch.zeros(n, TABLE_COLS, dtype=torch.int64)
    coef_off = plane_off = out_off = 0
    max_blocks = max_w = max_h = 1
    for i, info in enumerate(infos):
        w, h, nc = int(info[0]), int(info[1]), int(info[2])
        row = table[i]
        row[0], row[1], row[2], row[3], row[4] = w, h, nc, int(info[4]), int(info[5])
        blocks = 0
        for c in range(nc):
            wb, hb = int(info[6 + c]), int(info[9 + c])
            row[5 + c], row[8 + c], row[11 + c], row[14 + c] = wb, hb, int(info[12 + c]), int(info[15 + c])
            row[17 + c] = (coef_off + int(info[18 + c])) // 64 if sparse else coef_off + int(info[18 + c])
            row[20 + c] = plane_off + blocks * 64
            blocks += wb * hb
        row[23], row[24] = out_off, blocks
        coef_off += int(info[21])
        plane_off += blocks * 64
        out_off += (w * h * 3 + 15) // 16 * 16
        max_blocks, max_w, max_h = max(max_blocks, blocks), max(max_w, w), max(max_h, h)
    qtabs = torch.zeros(n, 3, 64, dtype=torch.int16, pin_memory=pin)
    qnp = qtabs.numpy().view(np.uint16)
    raw_idx = {i for i, _ in raw}
    coef_start = [0] * n
    acc = 0
    for i, info in enumerate(infos):
        coef_start[i] = acc
        acc += int(info[21])
    sparse_parts = [None] * n
    if not sparse:
        coefs = ring.take(coef_off) if ring is not None else torch.empty(max(coef_off, 8), dtype=torch.int16,
                                                                         pin_memory=pin)
        cnp = coefs.numpy()

    def decode_one(i):
        try:
            if sparse:
                off, idx, val, q, _ = entropy_decode_sparse(files[i])
                sparse_parts[i] = (off, idx, val)
            else:
                _, q, _ = entropy_decode(files[i], out=cnp[coef_start[i]:coef_start[i] + int(infos[i][21])])
        except UnsupportedJpeg:                          # found past the frame header (e.g. RGB-coded components)
            if unsupported == "raise":
                raise
            return i, None
        return i, q

    todo = [i for i in range(n) if i not in raw_idx]
    if threads > 1 and len(todo) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as pool:
            results = list(pool.map(decode_one, todo))
    else:
        results = [decode_one(i) for i in todo]
    for i, q in results:
        if q is None:
            table[i, 2] = table[i, 24] = 0               # the kernels skip this image; its pixels are copied in
            raw.append((i, _pil_pixels(files[i])))
        else:
            qnp[i] = q
    meta = torch.tensor([max_blocks, max_w, max_h, max(plane_off, 8), max(out_off, 16)], dtype=torch.int64)
    out = {"jpeg_qtabs": qtabs, "jpeg_table": table.pin_memory() if pin else table, "jpeg_meta": meta, "jpeg_raw": raw}
    if not sparse:
        out["jpeg_coefs"] = coefs
        return out
    # concatenate the files' entry lists; block offsets become batch-wide (a skipped / Pillow-decoded image keeps its
    # block range with zero entries, so the block numbering stays the dense one)
    total_blocks = coef_off // 64
    nnz = sum(p[1].size for p in sparse_parts if p is not None)
    sp_off = torch.empty(total_blocks + 1, dtype=torch.int32, pin_memory=pin)
    sp_idx = torch.empty(max(nnz, 16), dtype=torch.uint8, pin_memory=pin)
    sp_val = torch.empty(max(nnz, 16), dtype=torch.int16, pin_memory=pin)
    onp, inp, vnp = sp_off.numpy(), sp_idx.numpy(), sp_val.numpy()
    e = 0
    for i in range(n):
        b0, nb = coef_start[i] // 64, int(infos[i][21]) // 64
        part = sparse_parts[i]
        if part is None:
            onp[b0:b0 + nb] = e
            continue
        off, idx, val = part
        onp[b0:b0 + nb] = off[:nb] + e
        inp[e:e + idx.size] = idx
        vnp[e:e + val.size] = val
        e += idx.size
    onp[total_blocks] = e
    out.update(jpeg_sp_off=sp_off, jpeg_sp_idx=sp_idx, jpeg_sp_val=sp_val)
    return out


def reconstruct_batch(batch: dict, device=None):
    """Device half: ``pack_jpeg_batch``'s tensors (host or already on the device) -> (packed uint8 RGB CUDA buffer,
    int64 CUDA table [3, n] = byte offset | height | width) -- what ``ops.preprocess_u8_packed`` /
    ``data.GpuImageTransform.packed`` take.  Runs on the current stream."""
    sparse = "jpeg_sp_off" in batch
    dev = torch.device(device) if device is not None else batch["jpeg_qtabs"].device
    if dev.type != "cuda":
        raise _lib.B200MMError("reconstruct_batch needs a CUDA device (b200mm has no CPU path)")
    if sparse:
        sp_off, sp_idx, sp_val = (batch[k].to(dev, non_blocking=True) for k in ("jpeg_sp_off", "jpeg_sp_idx", "jpeg_sp_val"))
    else:
        coefs = batch["jpeg_coefs"].to(dev, non_blocking=True)
    qtabs = batch["jpeg_qtabs"].to(dev, non_blocking=True)
    table = batch["jpeg_table"].to(dev, non_blocking=True)
    max_blocks, max_w, max_h, plane_bytes, out_bytes = (int(v) for v in batch["jpeg_meta"])
    n = table.shape[0]
    planes = torch.empty(plane_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    if sparse:
        _lib.call("b200mm_jpeg_reconstruct_sparse", sp_off.data_ptr(), sp_idx.data_ptr(), sp_val.data_ptr(),
                  qtabs.data_ptr(), table.data_ptr(), n, max_blocks, max_w, max_h, planes.data_ptr(), out.data_ptr(), stream)
    else:
        _lib.call("b200mm_jpeg_reconstruct", coefs.data_ptr(), qtabs.data_ptr(), table.data_ptr(), n, max_blocks, max_w,
                  max_h, planes.data_ptr(), out.data_ptr(), stream)
    host_table = batch["jpeg_table"]
    for i, px in batch.get("jpeg_raw", ()):              # images the reference's loader decoded: their pixels are copied in
        o = int(host_table[i, 23])
        out[o:o + px.numel()].copy_(px.reshape(-1), non_blocking=True)
    image_table = torch.stack([table[:, 23], table[:, 1], table[:, 0]])
    return out, image_table


def decode_jpeg(files, device=None, unsupported: str = "raise", sparse: bool = False):
    """Convenience: list of file bytes -> list of uint8 CUDA tensors [H, W, 3] (views of one packed buffer)."""
    batch = pack_jpeg_batch(files, unsupported=unsupported, sparse=sparse)
    out, _ = reconstruct_batch(batch, device or "cuda")
    t = batch["jpeg_table"]
    return [out[int(t[i, 23]):int(t[i, 23]) + int(t[i, 0]) * int(t[i, 1]) * 3].view(int(t[i, 1]), int(t[i, 0]), 3)
            for i in range(t.shape[0])]


def collate_jpeg(samples, pin: bool | None = None, unsupported: str = "raise", threads: int = 1,
                 ring: CoefRing | None = None, sparse: bool = False):
    """DataLoader ``collate_fn`` for datasets whose ``image`` is the FILE CONTENT (bytes / uint8 1-D tensor, see
    ``data.file_bytes_loader``): stacks the token tensors / labels and entropy-decodes the batch's files in the worker
    (``pack_jpeg_batch``; ``functools.partial(collate_jpeg, unsupported="pil")`` for datasets with stray non-JPEG / CMYK
    files, ``threads=k`` to decode a batch's files concurrently).  ``pin=None`` pins the buffers when the collate runs in
    the training process and leaves pinning to ``DataLoader(pin_memory=True)`` inside worker processes.  ``loop.DevicePrefetcher`` finishes the decode on the device and runs the image transform."""
    if pin is None:     # pin here only in the training process: a DataLoader worker must not create a CUDA context
        pin = torch.utils.data.get_worker_info() is None and torch.cuda.is_available()
    out = {"id": [s["id"] for s in samples]}
    for k in ("text", "text_mask", "caption_text", "caption_text_mask", "label"):
        if k in samples[0]:
            t = torch.stack([s[k] for s in samples])
            out[k] = t.pin_memory() if pin else t
    out.update(pack_jpeg_batch([s["image"] for s in samples], pin=pin, unsupported=unsupported, threads=threads,
                               ring=ring, sparse=sparse))
    return out
