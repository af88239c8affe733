#!/usr/bin/env python
"""Per-role cycle accounting of the tensor-core kernels (needs lib/libxtb200_prof.so built with -DXTB_TC_PROF).
usage: python scripts/tc_prof.py [batch]   -> one line per (layer, op): average us per CTA spent in each phase"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libxtb200_prof.so")
from xingtian_b200.engine import Net, _ptr, stream_ptr
from xingtian_b200.model import archs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
lib = capi.lib()
lib.xtb_tc_prof_set.argtypes = [C.c_void_p]
arch = archs.ppo_cnn((84, 84, 4), 4, [256], "relu", True)
net = Net(arch, max_batch=B)
net.params.normal_(0, 0.05); net.params_changed()
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device="cuda")
net.forward(obs, B)
for n, _, _, _ in arch["layers"]:
    net.tensor_grad(n)[:B].normal_()
ghz = 1.965
names = {0: ["tile setup", "wait empty", "issue A", "issue B", "loop top", "arrive"],
         1: ["wait tmem-empty", "wait full", "mma issue", "tile top"],
         2: ["wait acc", "tmem ld", "store16", "tile top/arrive", "finish"]}
for li, which in [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (2, 2), (3, 0), (3, 1), (3, 2)]:
    run = lambda: capi.check(lib.xtb_net_bench_layer(net.handle, li, which, _ptr(obs), None, B, stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(300 * 32, dtype=torch.int64, device="cuda")
    lib.xtb_tc_prof_set(C.c_void_p(buf.data_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run(); a.record(); run(); b.record()
    torch.cuda.synchronize()
    lib.xtb_tc_prof_set(None)
    raw = buf.cpu().numpy().reshape(300, 32)
    live = raw[:, 24] > 0
    st = raw[live][:, 24:29].astype(np.float64)
    t0 = st[:, 0].min()
    st = (st - t0) / 1e3
    print("   CTA entry %.1f..%.1f us | prologue done +%.2f | dependency wait +%.2f | roles done (tid0) at %.1f..%.1f | all roles done at %.1f..%.1f (median %.1f)" % (
        st[:, 0].min(), st[:, 0].max(), (st[:, 1] - st[:, 0]).mean(), (st[:, 2] - st[:, 1]).mean(), st[:, 3].min(), st[:, 3].max(),
        st[:, 4].min(), st[:, 4].max(), np.median(st[:, 4])))
    t = raw[:, :24].astype(np.float64) / (ghz * 1e3)
    act = t[live]
    print("L%d/%d  launch %.1f us, %d CTAs" % (li, which, a.elapsed_time(b) * 1e3, len(act)))
    for role, rn in [(0, "producer"), (1, "mma     "), (2, "epilogue")]:
        seg = act[:, role * 8:role * 8 + 8]
        print("   %s total %5.1f (max %5.1f) | " % (rn, seg.sum(1).mean(), seg.sum(1).max()) + "  ".join("%s %.2f" % (names[role][i], seg[:, i].mean()) for i in range(len(names[role]))))
