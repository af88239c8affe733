#!/usr/bin/env python
"""Per-kernel device times (CUDA events, warm L2) of every layer op of PpoCnn at a given batch:
forward / weight-gradient / data-gradient launched alone through xtb_net_bench_layer, plus the
optimizer and loss kernels.  Usage: python scripts/layer_times.py [batch] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
from xingtian_b200.engine import Net, Adam, _ptr, stream_ptr
from xingtian_b200.model import archs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
arch_name = sys.argv[3] if len(sys.argv) > 3 else "ppo"
lib = capi.lib()
arch = {"ppo": lambda: archs.ppo_cnn((84, 84, 4), 4, [256], "relu", True), "impala": lambda: archs.impala_cnn((84, 84, 4), 4),
        "dqn": lambda: archs.dqn_cnn((84, 84, 4), 4)}[arch_name]()
net = Net(arch, max_batch=B)
net.params.normal_(0, 0.05); net.params_changed()
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device="cuda")
net.forward(obs, B)
for n, _, _, _ in arch["layers"]:
    net.tensor_grad(n)[:B].normal_()
torch.cuda.synchronize()


def timeit(fn):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3


rows = []
tot = 0.0
for i, (name, kind, src, sp) in enumerate(arch["layers"]):
    for which, tag in ((0, "fwd"), (1, "wgrad"), (2, "dgrad")):
        if which == 2 and src == "obs":
            continue
        t = timeit(lambda: capi.check(lib.xtb_net_bench_layer(net.handle, i, which, _ptr(obs), None, B, stream_ptr())))
        rows.append((name, tag, t)); tot += t
opt = Adam(net, 2.5e-4)
t = timeit(lambda: opt.step()); rows.append(("adam(3 kernels)", "opt", t)); tot += t
for r in rows:
    print("%-28s %-6s %8.1f us" % r)
print("sum %.1f us  (mode tc=%d, B=%d)" % (tot, lib.xtb_get_tc_mode(), B))
