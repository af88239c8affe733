#!/usr/bin/env python
"""bench.py -- learner env-steps/sec on the Atari 84x84x4 PPO hot path (BASELINE.json metric).

One "step" = one learner iteration of examples/breakout_ppo.yaml at E=32 envs (BASELINE config 2,
C2): T=128 batched policy-inference calls over the E stacked observations, GAE over [E,T], then
PPO.train = NUM_SGD_ITER(4) epochs x ceil(4096/320)=13 shuffled minibatches of fused
forward/loss/backward/clip/Adam.  N>1: one process per GPU, E=32 envs PER rank (weak scaling),
gradients all-reduced over NCCL every SGD step.

  value : N_samples / device time, rollout already resident in HBM (synthetic, seeded)
  e2e   : same iteration through the reference-facing plugin API (Algorithm.predict /
          prepare_data / train) with HOST numpy buffers; H2D/D2H inside the timed region
  --impl reference : the CPU restatement of the reference learner (oracle/) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E, T, A = 32, 128, 4
BATCH, EPOCHS = 320, 4
STATE = (84, 84, 4)
FLOP_PER_ENV_STEP = 164e6          # 12.62 MFLOP inference + 4 x 37.9 MFLOP train passes
MODEL_CFG = {"BATCH_SIZE": BATCH, "CRITIC_LOSS_COEF": 1.0, "ENTROPY_LOSS": 0.003, "LOSS_CLIPPING": 0.1,
             "LR": 0.00025, "MAX_GRAD_NORM": 5.0, "NUM_SGD_ITER": EPOCHS, "SUMMARY": False,
             "VF_SHARE_LAYERS": True, "activation": "relu", "hidden_sizes": [256],
             "action_type": "Categorical", "init_seed": 0}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sus=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of the reference learner on host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_iteration(state, sgd_steps=4, infer_steps=8):
    """One bounded sample of the C2 iteration on the CPU: `infer_steps` of the T=128 inference calls
    (batch E), the full GAE, `sgd_steps` of the 52 SGD minibatches (B=320); returns the time of a
    full iteration extrapolated from the three measured parts."""
    import torch
    from oracle import xt_oracle as orc
    ro, learner, arch = state["ro"], state["learner"], state["arch"]
    w = dict(zip(learner.names, learner.params))
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in range(infer_steps):
            rows = np.arange(E) * T + t
            logits, v = orc.forward(arch, w, ro["obs"][rows])
            orc.gumbel_argmax(logits.numpy(), state["u"][t])
    t_inf = (time.perf_counter() - t0) / infer_steps
    t0 = time.perf_counter()
    advs = [orc.gae(ro["value"][e], ro["reward"][e * T:(e + 1) * T], ro["done"][e * T:(e + 1) * T]) for e in range(E)]
    t_gae = time.perf_counter() - t0
    adv = np.concatenate([a[0] for a in advs]).astype(np.float32)
    ov = np.concatenate([a[1] for a in advs]); tv = np.concatenate([a[2] for a in advs]).astype(np.float32)
    inds = np.random.permutation(E * T)
    t0 = time.perf_counter()
    for s in range(sgd_steps):
        mb = inds[s * BATCH:(s + 1) * BATCH]
        learner.sgd_step(ro["obs"][mb], ro["action"][mb], ro["logp"][mb], adv[mb], ov[mb], tv[mb])
    t_sgd = (time.perf_counter() - t0) / sgd_steps
    n_sgd = EPOCHS * ((E * T + BATCH - 1) // BATCH)
    return t_inf * T + t_gae + t_sgd * n_sgd


def pick_threads(state):
    """The reference arm may use every host core, but oversubscribing small convolutions is slower than
    using fewer threads: time one SGD minibatch at a few thread counts and keep the fastest."""
    import torch
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    ro, learner = state["ro"], state["learner"]
    mb = np.arange(BATCH)
    z = np.zeros(BATCH, np.float32)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        learner.loss_and_grads(ro["obs"][mb], ro["action"][mb], ro["logp"][mb], z, z, z)
        t0 = time.perf_counter()
        learner.loss_and_grads(ro["obs"][mb], ro["action"][mb], ro["logp"][mb], z, z, z)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    state["threads"] = best
    return best


def make_cpu_state():
    import torch
    from oracle import xt_oracle as orc
    arch = orc.ppo_cnn_arch()
    w = orc.init_weights(arch, seed=0)
    ro = orc.synth_ppo_rollout(0, E, T)
    learner = orc.PpoLearner(arch, w, lr=0.00025, batch_size=BATCH, ent_coef=0.003, clip_ratio=0.1, num_sgd_iter=EPOCHS)
    u = np.random.default_rng(0).random((T, E, A)).astype(np.float32) * 0.998 + 0.001
    st = dict(arch=arch, ro=ro, learner=learner, u=u)
    pick_threads(st)
    return st


SAMPLE_DESC = ("8 of 128 inference calls (batch 32) + full GAE + 4 of 52 SGD minibatches (B=320) of the C2 iteration, "
               "extrapolated to the full iteration; torch-CPU fp32 restatement of xt.model.ppo/xt.algorithm.ppo")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    st = make_cpu_state()
    for _ in range(max(1, args.warmup)):
        cpu_reference_iteration(st)
    times = [cpu_reference_iteration(st) for _ in range(args.steps)]
    t = float(np.mean(times))
    val = E * T / t
    out = {"impl": "reference", "metric": "learner env-steps/sec (Atari 84x84x4 PPO)", "value": val, "unit": "env-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "breakout_ppo C2: PpoCnn E=32 T=128 N=4096 B=320 x4 epochs (52 SGD steps) + 128 batched inference calls + GAE"},
           "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": st["threads"], "host_cpus": os.cpu_count(), "kind": "port", "sample": SAMPLE_DESC},
           "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import xingtian_b200 as xb
    from xingtian_b200 import capi, engine
    from xingtian_b200.engine import _ptr, stream_ptr
    from oracle import xt_oracle as orc   # synthetic rollout generator + cpu baseline only
    lib = capi.lib()
    dev = torch.device("cuda", local)
    info = {"actor": {"model_name": "PpoCnn", "state_dim": list(STATE), "action_dim": A, "input_dtype": "uint8",
                      "model_config": dict(MODEL_CFG), "device": "cuda:%d" % local, "max_predict_batch": BATCH}}
    alg = xb.alg_builder("PPO", info, {"instance_num": E, "agent_num": 1})
    model = alg.actor
    dp = engine.GradAllReduce(model.net) if world > 1 else None
    ro = orc.synth_ppo_rollout(rank, E, T)       # every rank: its own envs
    n = E * T
    # ---- device-resident rollout -----------------------------------------------------------------
    store = model.rollout
    store.reserve(n)
    store.obs[:n].copy_(torch.from_numpy(ro["obs"]))
    reward_d = torch.from_numpy(ro["reward"].astype(np.float32)).to(dev)
    done_d = torch.from_numpy(ro["done"].view(np.uint8)).to(dev)
    value_d = torch.zeros(E, T + 1, device=dev)
    step_idx = (torch.arange(E, dtype=torch.int32, device=dev)[None, :] * T + torch.arange(T, dtype=torch.int32, device=dev)[:, None]).contiguous()
    act_t = torch.empty(T, E, dtype=torch.int32, device=dev); logp_t = torch.empty(T, E, device=dev); val_t = torch.empty(T + 1, E, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    seg = {"infer": [], "gae": [], "train": []}

    def device_iteration(record=False):
        # (1) batched inference over the E stacked observations, once per env step
        if record:
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
        model.rollout_infer_device(store.obs, step_idx, E, T, act_t, logp_t, val_t)
        val_t[T].copy_(val_t[T - 1])            # bootstrap value (synthetic rollout: no next observation)
        store.action[:n].copy_(act_t.t().reshape(-1)); store.old_logp[:n].copy_(logp_t.t().reshape(-1))
        value_d.copy_(val_t.t())
        if record:
            e1.record()
        # (2) GAE on [E,T]
        capi.check(lib.xtb_gae(_ptr(value_d), _ptr(reward_d), _ptr(done_d), E, T, 0.99, 0.95, 1,
                               _ptr(store.adv), _ptr(store.old_v), _ptr(store.target_v), stream_ptr()))
        if record:
            e2.record()
        # (3) minibatch SGD
        loss = model.train_device(n)
        if record:
            e3.record(); torch.cuda.synchronize()
            seg["infer"].append(e0.elapsed_time(e1)); seg["gae"].append(e1.elapsed_time(e2)); seg["train"].append(e2.elapsed_time(e3))
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    np.random.seed(1234 + rank)
    for _ in range(args.warmup):
        flush.fill_(1)
        device_iteration()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.xtb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i)                           # L2 flush between timed iterations (outside the events)
        ev[i][0].record()
        device_iteration()
        ev[i][1].record()
    barrier()
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    launches = lib.xtb_launch_count() - launches0
    for _ in range(2):
        device_iteration(record=True)
    breakdown = {k: float(np.mean(v)) for k, v in seg.items()}
    # ---- e2e through the plugin API with host buffers ----------------------------------------------
    host_obs = [np.ascontiguousarray(ro["obs"][np.arange(E) * T + t]) for t in range(T)]
    traj = []
    for e in range(E):
        sl = slice(e * T, (e + 1) * T)
        traj.append(dict(cur_state=ro["obs"][sl], action=ro["action"][sl], logp=ro["logp"][sl],
                         value=ro["value"][e], reward=ro["reward"][sl], done=ro["done"][sl]))
    alg.sign_clip_reward = True

    def e2e_iteration():
        for t in range(T):
            model.predict(host_obs[t])           # H2D obs, D2H (action, logp, value)
        for e in range(E):
            alg.prepare_data(traj[e])            # H2D trajectory (pinned bounce), device GAE
        return alg.train()                       # D2H loss trace

    e2e_steps = max(1, min(args.steps, 5))
    e2e_iteration()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_iteration()
    barrier()
    e2e_s_local = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    # ---- max over ranks --------------------------------------------------------------------------
    tt = torch.tensor([ms_local, e2e_s_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: every layer op of one SGD minibatch (B=320) is launched alone and
    #      timed with CUDA events on the launching stream (L2 flushed before each launch); the one with the largest
    #      time is reported against the measured bf16 tensor peak with its ALGORITHMIC flops (2*M*N*K).
    pk = peaks()
    net = model.net
    idx0 = step_idx.reshape(-1)[:BATCH].contiguous()
    net.forward(store.obs, BATCH, idx=idx0)
    for name, _, _, _ in model.arch["layers"]:
        net.tensor_grad(name)[:BATCH].normal_()
    shapes = {"obs": STATE}
    ops = []
    for li, (name, kind, src, sp) in enumerate(model.arch["layers"]):
        ish = shapes[src]
        if kind == "conv":
            oh, ow = (ish[0] - sp["k"]) // sp["s"] + 1, (ish[1] - sp["k"]) // sp["s"] + 1
            shapes[name] = (oh, ow, sp["cout"])
            gm, gn, gk = BATCH * oh * ow, sp["cout"], sp["k"] * sp["k"] * ish[2]
        else:
            shapes[name] = (sp["n"],)
            gm, gn, gk = BATCH, sp["n"], int(np.prod(ish))
        for which, tag in ((0, "forward"), (1, "weight-gradient"), (2, "data-gradient")):
            if which == 2 and src == "obs":
                continue
            ops.append((li, which, "%s %s (M=%d N=%d K=%d)" % (name, tag, gm, gn, gk), 2.0 * gm * gn * gk))
    reps = 20
    best = None
    for li, which, label, flop in ops:
        run = lambda: capi.check(lib.xtb_net_bench_layer(net.handle, li, which, _ptr(store.obs), _ptr(idx0), BATCH, stream_ptr()))
        for _ in range(3):
            run()
        # average launch duration: `reps` launches of the kernel back to back inside one event pair on the launching
        # stream (operands L2-resident, as they are inside the SGD step where the previous layer has just produced them)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if best is None or ms > best[0]:
            best = (ms, label, flop, li, which)
    k_ms, k_label, k_flop, k_li, k_which = best
    # the same kernel alone with a cold L2 (256 MiB flush before the launch), single launch between two events
    run = lambda: capi.check(lib.xtb_net_bench_layer(net.handle, k_li, k_which, _ptr(store.obs), _ptr(idx0), BATCH, stream_ptr()))
    cold = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        cold.append(a.elapsed_time(b))
    achieved = k_flop / (k_ms * 1e-3) / 1e12
    traffic = None
    try:   # dram bytes of this kernel from the committed ncu --set full capture (profiles/), if present
        with open(os.path.join(ROOT, "profiles", "r1_dominant_kernel.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch", {}).get("L%d/%d" % (k_li, k_which))
    except Exception:
        pass
    roofline = {"kernel": "tc_gemm_kernel: " + k_label, "bound": "tensor", "achieved": achieved, "peak": pk["tf"],
                "unit": "TFLOP/s", "frac": achieved / pk["tf"], "peak_source": pk["src"] + " bf16 burst (cuBLAS)",
                "traffic": traffic, "ms_per_launch": k_ms, "ms_single_launch_cold_l2": float(np.median(cold)),
                "note": "ms_per_launch = mean of 20 back-to-back launches (CUDA events on the launching stream); flops are "
                        "algorithmic 2MNK, bf16x3 issues 2-3 tensor-core MACs per algorithmic MAC; traffic = dram bytes "
                        "of the cold-cache ncu capture in profiles/"}
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)
    whole = {"achieved_tflops": value * FLOP_PER_ENV_STEP / 1e12 / world, "frac_of_sustained_bf16": value * FLOP_PER_ENV_STEP / 1e12 / world / pk["tf_sus"]}
    # ---- cpu baseline (rank 0, N=1) --------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        st = make_cpu_state()
        cpu_reference_iteration(st, sgd_steps=2, infer_steps=4)
        tc = float(np.mean([cpu_reference_iteration(st) for _ in range(3)]))
        cpu = {"value": n / tc, "unit": "env-steps/s", "cores": st["threads"], "host_cpus": os.cpu_count(), "kind": "port", "sample": SAMPLE_DESC}
    h2d = 2 * n * int(np.prod(STATE)) + n * (4 + 4 + 4 + 1) + E * 4 + EPOCHS * n * 4
    d2h = T * E * 12 + EPOCHS * ((n + BATCH - 1) // BATCH) * 4
    out = {"metric": "learner env-steps/sec (Atari 84x84x4 PPO)", "value": value, "unit": "env-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32 (bf16x3 split on tcgen05, fp32 accumulate in TMEM)", "data": "synthetic",
           "config": {"workload": "breakout_ppo C2: PpoCnn E=32/GPU T=128 N=4096 B=320 x4 epochs (52 SGD steps) + 128 batched inference calls + GAE",
                      "timing": "CUDA events per iteration, max over ranks; 256 MiB L2 flush between timed iterations",
                      "parallelism": "dp%d" % world},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": world * n / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
           "roofline": roofline, "whole_step": whole, "breakdown_ms": breakdown, "cpu_baseline": cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
