#!/usr/bin/env python
"""Diagnostic: time layer ops with experiment builds of the library (no epilogue stores / no A gathers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
variant = sys.argv[1]
if variant != "base":
    capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libxtb200_%s.so" % variant)
from xingtian_b200.engine import Net, _ptr, stream_ptr
from xingtian_b200.model import archs
lib = capi.lib()
B = 320
arch = archs.ppo_cnn((84, 84, 4), 4, [256], "relu", True)
net = Net(arch, max_batch=B)
net.params.normal_(0, 0.05); net.params_changed()
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device="cuda")
net.forward(obs, B)
for n, _, _, _ in arch["layers"]:
    net.tensor_grad(n)[:B].normal_()
res = []
for li, which in [(0, 0), (0, 1), (1, 0), (1, 2), (2, 2), (3, 0)]:
    f = lambda: capi.check(lib.xtb_net_bench_layer(net.handle, li, which, _ptr(obs), None, B, stream_ptr()))
    # 20 back-to-back launches inside one event pair: launch gaps overlap, per-launch GPU time dominates
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    b.record(); torch.cuda.synchronize()
    res.append("L%d/%d %.1f" % (li, which, a.elapsed_time(b) * 1e3 / 20))
print(variant, " us per launch:", "  ".join(res))
