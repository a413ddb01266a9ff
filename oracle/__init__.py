"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's task-2C hot path (fp32, stock PyTorch / transformers / torchvision
modules, exactly the libraries the reference calls).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product path
(b200mm) never does and has no CPU fallback.
"""
