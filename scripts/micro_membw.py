"""HBM bandwidth by traffic mix on this GPU (CUDA-graph replays of 20 back-to-back torch kernels): fill, copy, sum.
CAUTION when reading the numbers: torch's bf16 fill kernel stores 8 bytes per thread and reaches 3.9 TB/s, an fp32
zero_() (16-byte stores) reaches 6.9 TB/s -- the low figure is that kernel's, not a write limit of the HBM (it was
briefly mistaken for one; profiles/membw_r02.log)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")


def graph_b2b(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for mb in (103, 411, 1024):
    n = mb * 1024 * 1024 // 2
    a = torch.empty(n, device=dev, dtype=torch.bfloat16)
    b = torch.empty(n, device=dev, dtype=torch.bfloat16)
    f32 = torch.empty(n // 2, device=dev, dtype=torch.float32)
    t = graph_b2b(lambda: f32.zero_()); print(f"{mb:5d} MB fp32 zero_ (write only) {t:7.1f} us  {mb * 1.048576 / t * 1e3:6.0f} GB/s")
    t = graph_b2b(lambda: a.fill_(1.0)); print(f"{mb:5d} MB fill (write only)      {t:7.1f} us  {mb * 1.048576 / t * 1e3:6.0f} GB/s")
    t = graph_b2b(lambda: a.zero_()); print(f"{mb:5d} MB zero_ (write only)     {t:7.1f} us  {mb * 1.048576 / t * 1e3:6.0f} GB/s")
    t = graph_b2b(lambda: b.copy_(a)); print(f"{mb:5d} MB copy (1 read : 1 write) {t:7.1f} us  {2 * mb * 1.048576 / t * 1e3:6.0f} GB/s")
    t = graph_b2b(lambda: torch.sum(a.view(torch.int16))); print(f"{mb:5d} MB sum (read only)        {t:7.1f} us  {mb * 1.048576 / t * 1e3:6.0f} GB/s")
