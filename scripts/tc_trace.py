#!/usr/bin/env python
"""Timeline of one tensor-core kernel launch (needs lib/libxtb200_trace.so built with -DXTB_TC_TRACE).
usage: python scripts/tc_trace.py <layer> <which 0|1|2> [batch]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libxtb200_trace.so")
from xingtian_b200.engine import Net, _ptr, stream_ptr
from xingtian_b200.model import archs
layer, which = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 320
lib = capi.lib()
lib.xtb_tc_trace_set.argtypes = [C.c_void_p]
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
SL = 256
buf = torch.zeros(300 * SL, dtype=torch.int64, device="cuda")
lib.xtb_tc_trace_set(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
lib.xtb_tc_trace_set(None)
t = buf.cpu().numpy().reshape(300, SL).view(np.uint64)
ctas = [i for i in range(300) if t[i, 0] > 0]
ev = {}
t0 = None
rows = []
for c in ctas:
    n = int(min(t[c, 0], SL - 1))
    e = [(int(x >> np.uint64(8)), int(x & np.uint64(255))) for x in t[c, 1:n + 1]]
    rows.append(e)
    mn = min(x[0] for x in e)
    t0 = mn if t0 is None else min(t0, mn)
print("CTAs traced:", len(ctas), " events in CTA0:", len(rows[0]))
def stat(code, first=True):
    v = []
    for e in rows:
        xs = [x[0] - t0 for x in e if x[1] == code]
        if xs: v.append(xs[0] if first else xs[-1])
    return (np.min(v) / 1e3, np.median(v) / 1e3, np.max(v) / 1e3) if v else None
for code, name in [(1, "prologue done"), (2, "dependency ok"), (10, "first stage issued"), (20, "first stage landed"), (30, "first accumulator ready"), (31, "first tile stored")]:
    print("%-26s min/med/max us: %s" % (name, stat(code)))
for code, name in [(10, "last stage issued"), (20, "last stage landed"), (30, "last accumulator ready"), (31, "last tile stored"), (3, "dealloc")]:
    print("%-26s min/med/max us: %s" % (name, stat(code, False)))
e = sorted(rows[0])
print("CTA0 timeline (us, code):", [(round((x[0] - t0) / 1e3, 2), x[1]) for x in e if x[1] not in (1, 2)][:90])
