#!/usr/bin/env python
"""Host-side cost of the IMPALA reference-facing calls at C3 size (E=64): predict (both paths), prepare_data, train."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xingtian_b200 as xb
from xingtian_b200 import synth

E, T, B = 64, 128, 512
info = {"actor": {"model_name": "ImpalaCnnOpt", "state_dim": [84, 84, 4], "action_dim": 4, "input_dtype": "uint8", "state_mean": 0.0,
                  "state_std": 255.0, "max_batch": B, "model_config": {"LR": 0.0005, "sample_batch_step": T, "grad_norm_clip": 40.0, "init_seed": 0}}}
alg = xb.alg_builder("IMPALAOpt", info, {"instance_num": E, "agent_num": 1, "BATCH_SIZE": B})
m = alg.actor
ro = synth.ppo_rollout(0, E, T)
frames = [np.ascontiguousarray(ro["obs"][np.arange(E) * T + t]) for t in range(8)]
u = np.random.default_rng(0).random((E, 4)).astype(np.float32) * 0.99 + 0.005


def bench(fn, n):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("predict, one native call (graph)      %8.1f us" % bench(lambda i: m.predict(frames[i & 7]), 200))
print("predict, eager path (uniforms given)  %8.1f us" % bench(lambda i: m.predict(frames[i & 7], uniforms=u), 200))
traj = [dict(cur_state=ro["obs"][e * T:(e + 1) * T], logit=ro["logits"][e * T:(e + 1) * T], action=ro["action"][e * T:(e + 1) * T],
             reward=ro["reward"][e * T:(e + 1) * T].astype(np.float32), done=ro["done"][e * T:(e + 1) * T]) for e in range(E)]


def it(i):
    t0 = time.perf_counter()
    for e in range(E):
        alg.prepare_data(traj[e])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    alg.train()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    it.acc[0] += t1 - t0; it.acc[1] += t2 - t1


it.acc = [0.0, 0.0]
for i in range(2):
    it(i)
it.acc = [0.0, 0.0]
for i in range(4):
    it(i)
print("prepare_data x %d                      %8.2f ms" % (E, it.acc[0] / 4 * 1e3))
print("train() (%d steps)                     %8.2f ms" % (E * T // B, it.acc[1] / 4 * 1e3))
