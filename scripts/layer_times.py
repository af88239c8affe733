#!/usr/bin/env python
"""Per-kernel device times of every layer op of a network at a given batch: forward / weight-gradient / data-gradient
launched through xtb_net_bench_layer, `reps` back-to-back launches inside one CUDA-event pair (warm L2), so the figure
is the launch-to-launch period of the kernel (its duration plus the inter-kernel gap).
Usage: python scripts/layer_times.py [batch] [reps] [ppo|impala|dqn]      (XTB_LIB_PATH selects an experiment build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xingtian_b200 import capi
from xingtian_b200.engine import Net, Adam, _ptr, stream_ptr
from xingtian_b200.model import archs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
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
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


rows = []
tot = 0.0
t = timeit(lambda: capi.check(lib.xtb_net_bench_layer(net.handle, 0, 3, _ptr(obs), None, B, stream_ptr())))
rows.append(("frame decode", "", t)); tot += t
for i, (name, kind, src, sp) in enumerate(arch["layers"]):
    for which, tag in ((0, "fwd"), (1, "wgrad"), (2, "dgrad")):
        if which == 2 and src == "obs":
            continue
        t = timeit(lambda: capi.check(lib.xtb_net_bench_layer(net.handle, i, which, _ptr(obs), None, B, stream_ptr())))
        rows.append((name, tag, t)); tot += t
opt = Adam(net, 2.5e-4)
t = timeit(lambda: opt.step()); rows.append(("adam + weight blobs", "opt", t)); tot += t
for r in rows:
    print("%-28s %-6s %8.1f us" % r)
print("sum %.1f us  (mode tc=%d, B=%d, lib=%s)" % (tot, lib.xtb_get_tc_mode(), B, os.path.basename(capi.LIB_PATH)))
