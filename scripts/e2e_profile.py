#!/usr/bin/env python
"""Host-side time of the plugin entry points (predict / prepare_data / train) at C2 shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xingtian_b200 as xb
E, T = 32, 128
info = {"actor": {"model_name": "PpoCnn", "state_dim": [84, 84, 4], "action_dim": 4, "input_dtype": "uint8",
                  "model_config": {"BATCH_SIZE": 320, "ENTROPY_LOSS": 0.003, "LOSS_CLIPPING": 0.1, "LR": 0.00025, "NUM_SGD_ITER": 4,
                                   "hidden_sizes": [256], "action_type": "Categorical", "init_seed": 0}, "max_predict_batch": 320}}
alg = xb.alg_builder("PPO", info, {"instance_num": E, "agent_num": 1})
rng = np.random.default_rng(0)      # synthetic rollout of the C2 shapes (the numbers do not matter for host timing)
ro = dict(obs=rng.integers(0, 256, (E * T, 84, 84, 4), dtype=np.uint8), action=rng.integers(0, 4, E * T).astype(np.int32),
          logp=np.full((E * T, 1), -1.386, np.float32), value=rng.standard_normal((E, T + 1, 1)).astype(np.float32),
          reward=rng.integers(-1, 2, (E * T, 1)).astype(np.float32), done=rng.random((E * T, 1)) < 0.01)
host_obs = [np.ascontiguousarray(ro["obs"][np.arange(E) * T + t]) for t in range(T)]
alg.actor.keep_predict_obs(E, T)      # frames reach the device once, inside predict()
traj = [dict(ring_rows=(e, 0, T), action=ro["action"][e * T:(e + 1) * T], logp=ro["logp"][e * T:(e + 1) * T],
             value=ro["value"][e], reward=ro["reward"][e * T:(e + 1) * T], done=ro["done"][e * T:(e + 1) * T]) for e in range(E)]
for rep in range(4):
    alg.actor._obs_ring["t"] = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(T):
        alg.actor.predict(host_obs[t])
    t1 = time.perf_counter()
    for e in range(E):
        alg.prepare_data(traj[e])
    t2a = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    alg._device_gae()
    t3a = time.perf_counter()
    perm = alg.actor.make_perm(alg._count)
    t3b = time.perf_counter()
    loss = alg.actor.train_device(alg._count, perm)
    t3 = time.perf_counter()
    alg._init_train_list()
    print("predict x128 %.2f ms | prepare_data x32 %.2f ms host + %.2f ms drain | device GAE call %.2f | make_perm %.2f | train_device (launch + sync + loss) %.2f ms"
          % ((t1 - t0) * 1e3, (t2a - t1) * 1e3, (t2 - t2a) * 1e3, (t3a - t2) * 1e3, (t3b - t3a) * 1e3, (t3 - t3b) * 1e3))

# ---- split of one predict() call: staged H2D alone, device inference alone ------------------------------
from xingtian_b200.engine import _ptr, stream_ptr
from xingtian_b200.capi import check
m = alg.actor
io = m._predict_io(E)
lib = m.net.lib
for rep in range(2):
    t0 = time.perf_counter()
    for t in range(T):
        check(lib.xtb_copy_h2d_staged(io["obs_ptr"], host_obs[t].ctypes.data, host_obs[t].nbytes, stream_ptr()))
        torch.cuda.current_stream().synchronize()
    t1 = time.perf_counter()
    for t in range(T):
        m.rollout_infer_device(io["obs"], None, E, 1, io["act"], io["logp"], io["val"])
        io["pin_out"].copy_(io["out_dev"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    pin = torch.empty(host_obs[0].shape, dtype=torch.uint8).pin_memory()
    for t in range(T):
        io["obs"].copy_(pin, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t3 = time.perf_counter()
    for t in range(T):
        np.copyto(pin.numpy(), host_obs[t])
    t4 = time.perf_counter()
    print("per call: staged h2d+sync %.1f us | infer+d2h+sync %.1f us | pinned h2d+sync %.1f us | np.copyto %.1f us" %
          ((t1 - t0) / T * 1e6, (t2 - t1) / T * 1e6, (t3 - t2) / T * 1e6, (t4 - t3) / T * 1e6))
