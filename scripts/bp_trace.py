#!/usr/bin/env python
"""Timeline of one tensor-core kernel launch (needs lib/libxtb200_trace.so: scripts/exp_variants.sh).
usage: XTB_LIB_PATH=xingtian_b200/lib/libxtb200_trace.so python scripts/bp_trace.py <layer> <which 0|2> [batch]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
from xingtian_b200.engine import Net, _ptr, stream_ptr
from xingtian_b200.model import archs
layer, which = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 320
lib = capi.lib()
h = C.CDLL(capi.LIB_PATH)
h.xtb_bp_trace_set.argtypes = [C.c_void_p]
arch = archs.ppo_cnn((84, 84, 4), 4, [256], "relu", True)
net = Net(arch, max_batch=B)
net.params.normal_(0, 0.05); net.params_changed()
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device="cuda")
net.forward(obs, B)
for n, _, _, _ in arch["layers"]:
    net.tensor_grad(n)[:B].normal_()
run = lambda: capi.check(lib.xtb_net_bench_layer(net.handle, layer, which, _ptr(obs), None, B, stream_ptr()))
for _ in range(3):
    run()
torch.cuda.synchronize()
NTR = 22
buf = torch.zeros(4 * 3 * NTR, dtype=torch.int64, device="cuda")
h.xtb_bp_trace_set(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
h.xtb_bp_trace_set(None)
d = buf.cpu().numpy().reshape(4, 3, NTR)
names = {1: "P wait-empty", 2: "P got-empty", 3: "P issued", 10: "M wait-tmem", 11: "M got-tmem", 12: "M full", 13: "M issued", 20: "E wait-acc", 21: "E got-acc", 22: "E stored"}
for cta in range(2):
    ev = []
    for role in range(3):
        n = int(d[cta, role, 0])
        for v in d[cta, role, 1:min(n, NTR - 1) + 1]:
            ev.append((int(v) >> 8, int(v) & 0xff))
    ev.sort()
    if not ev:
        continue
    t0 = ev[0][0]
    print("---- CTA %d: %d events, span %.2f us" % (cta, len(ev), (ev[-1][0] - t0) / 1e3))
    for t, c in ev[:160]:
        print("%8.2f us  %s" % ((t - t0) / 1e3, names.get(c, str(c))))
