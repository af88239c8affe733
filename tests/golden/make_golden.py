#!/usr/bin/env python
"""Generate golden fixtures by EXECUTING the reference's own source files from /root/reference.

The reference package cannot be imported as a whole here (python 3.12: no `imp`; tensorflow, gym, lz4
absent), but its non-TensorFlow arithmetic is plain numpy/python.  This script loads the individual
reference files with stub modules standing in for the unavailable imports (tf sessions, registries,
base classes) and records what the reference code itself computes:

  gae.npz            PPO.data_proc                (xt/agent/ppo/ppo.py:77-106)
  ppo_minibatch.npz  PPO.train minibatch slicing  (xt/model/ppo/ppo.py:111-132)   [sess.run stubbed]
  dqn_target.npz     DQN.train TD-target loop     (xt/algorithm/dqn/dqn.py:61-103) [actor.predict stubbed]
  impala_proc.npz    IMPALAOpt._data_proc/train slicing (xt/algorithm/impala/impala_opt.py:73-147)

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures are committed; nothing on the GPU box reads /root/reference."""
import importlib.util
import os
import random
import sys
import types
from collections import defaultdict

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _Reg(object):
    def __call__(self, cls):
        return cls


def install_stubs():
    if not hasattr(np, "bool"):
        np.bool = bool   # impala_opt.py:143 uses the alias removed in numpy>=1.24
    regs = type("Registers", (), {"agent": _Reg(), "model": _Reg(), "algorithm": _Reg(), "env": _Reg(), "comm": _Reg()})
    for pkg in ("xt", "xt.agent", "xt.agent.ppo", "xt.model", "xt.model.ppo", "xt.algorithm", "xt.algorithm.dqn",
                "xt.algorithm.impala", "zeus", "zeus.common", "zeus.common.util", "zeus.common.ipc", "absl"):
        _mod(pkg).__path__ = []
    _mod("zeus.common.util.register", Registers=regs)
    _load("zeus.common.util.common", "zeus/common/util/common.py")      # the real import_config
    _mod("absl.logging", info=lambda *a, **k: None, debug=lambda *a, **k: None, warning=lambda *a, **k: None)
    sys.modules["absl"].logging = sys.modules["absl.logging"]


def golden_gae():
    class Agent(object):
        def get_trajectory(self, last_pred=None):
            return self.trajectory
    sys.modules["xt.agent"].Agent = Agent
    _load("xt.agent.ppo.default_config", "xt/agent/ppo/default_config.py")
    ref = _load("xt.agent.ppo.ppo", "xt/agent/ppo/ppo.py")
    rng = np.random.default_rng(2024)
    out = {}
    for case, (T, pdone) in enumerate([(128, 0.01), (128, 0.2), (200, 0.0), (7, 0.3), (1, 0.0), (33, 1.0)]):
        ag = object.__new__(ref.PPO)
        ag.trajectory = defaultdict(list)
        value = [rng.standard_normal(1).astype(np.float32) for _ in range(T + 1)]     # model output v[0] per step (:49)
        reward = [float(x) for x in rng.choice([-1.0, 0.0, 1.0], size=T, p=[0.05, 0.9, 0.05])]
        done = [bool(x) for x in (rng.random(T) < pdone)]
        ag.trajectory["cur_state"] = [rng.standard_normal(4).astype(np.float32) for _ in range(T)]
        ag.trajectory["action"] = [int(x) for x in rng.integers(0, 2, T)]
        ag.trajectory["logp"] = [rng.standard_normal(1).astype(np.float32) for _ in range(T)]
        ag.trajectory["value"] = list(value)
        ag.trajectory["reward"] = list(reward)
        ag.trajectory["done"] = list(done)
        ag.data_proc()
        out["c%d_value" % case] = np.asarray(value)
        out["c%d_reward" % case] = np.asarray(reward)
        out["c%d_done" % case] = np.asarray(done)
        out["c%d_adv" % case] = np.asarray(ag.trajectory["adv"])
        out["c%d_old_value" % case] = np.asarray(ag.trajectory["old_value"])
        out["c%d_target_value" % case] = np.asarray(ag.trajectory["target_value"])
    np.savez(os.path.join(OUT, "gae.npz"), **out)
    print("gae.npz", {k: v.dtype for k, v in out.items() if k.startswith("c0")})


def golden_ppo_minibatch():
    class XTModel(object):
        pass

    class _Ctx(object):
        def __enter__(self): return self
        def __exit__(self, *a): return False

    sys.modules["xt.model"].XTModel = XTModel
    sys.modules["xt.model.ppo"].actor_loss_with_entropy = None
    sys.modules["xt.model.ppo"].critic_loss = None
    _load("xt.model.ppo.default_config", "xt/model/ppo/default_config.py")
    _mod("xt.model.tf_compat", tf=types.SimpleNamespace())
    _mod("xt.model.tf_dist", make_dist=lambda *a: None)
    _mod("xt.model.tf_utils", TFVariables=object)
    ref = _load("xt.model.ppo.ppo", "xt/model/ppo/ppo.py")
    out = {}
    for case, (n, bs, iters, seed) in enumerate([(4096, 320, 4, 0), (64, 24, 2, 123), (200, 200, 8, 9), (7, 3, 1, 5)]):
        m = object.__new__(ref.PPO)
        m.graph = types.SimpleNamespace(as_default=lambda: _Ctx())
        m.num_sgd_iter, m._batch_size = iters, bs
        for ph in ("state_ph", "behavior_action_ph", "old_logp_ph", "adv_ph", "old_v_ph", "target_v_ph", "train_op", "loss"):
            setattr(m, ph, ph)
        seen = []

        class Sess(object):
            def run(self, fetches, feed_dict):
                seen.append(np.asarray(feed_dict["adv_ph"]).reshape(-1).astype(np.int64).copy())
                assert np.array_equal(feed_dict["state_ph"].reshape(-1), feed_dict["adv_ph"].reshape(-1))
                return [None, float(len(seen))]
        m.sess = Sess()
        ids = np.arange(n, dtype=np.float64)
        np.random.seed(seed)
        mean_loss = m.train([ids.copy()], [ids.copy(), ids.copy(), ids.copy(), ids.copy(), ids.copy()])
        out["c%d_cfg" % case] = np.asarray([n, bs, iters, seed])
        out["c%d_order" % case] = np.concatenate(seen)
        out["c%d_sizes" % case] = np.asarray([len(s) for s in seen])
        out["c%d_mean_loss" % case] = np.asarray(mean_loss)
    np.savez(os.path.join(OUT, "ppo_minibatch.npz"), **out)
    print("ppo_minibatch.npz", out["c0_sizes"][:14], out["c1_sizes"])


def golden_dqn():
    class Algorithm(object):
        pass
    sys.modules["xt.algorithm"].Algorithm = Algorithm
    _load("xt.algorithm.dqn.default_config", "xt/algorithm/dqn/default_config.py")
    rb = _load("xt.algorithm.replay_buffer", "xt/algorithm/replay_buffer.py")
    sys.modules["xt.model"].model_builder = lambda info: None
    ref = _load("xt.algorithm.dqn.dqn", "xt/algorithm/dqn/dqn.py")
    out = {}
    rng = np.random.default_rng(77)
    for case, (A, B, n, double) in enumerate([(4, 32, 100, False), (6, 16, 40, True), (4, 512, 700, False)]):
        ref.BATCH_SIZE = B
        ref.TARGET_UPDATE_FREQ = 10 ** 9
        alg = object.__new__(ref.DQN)
        alg.double_dqn = double
        alg.train_count = 0
        alg.buff = rb.ReplayBuffer(1000)
        # states are scalars (ids); the stub networks are fixed tables indexed by id
        qtab = rng.standard_normal((n, A)).astype(np.float32)        # online Q(s)
        qn_online = rng.standard_normal((n, A)).astype(np.float32)   # online Q(s')
        qn_target = rng.standard_normal((n, A)).astype(np.float32)   # target Q(s')
        act = rng.integers(0, A, n); rew = np.sign(rng.standard_normal(n)); done = rng.random(n) < 0.15
        for i in range(n):
            alg.buff.add((np.float32(i), int(act[i]), float(rew[i]), np.float32(i + 100000), bool(done[i])))
        captured = {}

        class Actor(object):
            def predict(self, x):
                x = np.asarray(x).astype(np.int64)
                return (qtab[x] if x.max() < 100000 else qn_online[x - 100000]).copy()

            def train(self, states, y):
                captured["states"] = np.asarray(states).astype(np.int64).copy(); captured["y"] = np.asarray(y).copy()
                return 0.5

        class Target(object):
            def predict(self, x):
                return qn_target[np.asarray(x).astype(np.int64) - 100000].copy()
        alg.actor, alg.target_actor = Actor(), Target()
        random.seed(31 + case)
        alg.train()
        pre = "c%d_" % case
        out.update({pre + "q": qtab, pre + "qn_online": qn_online, pre + "qn_target": qn_target, pre + "action": act,
                    pre + "reward": rew, pre + "done": done, pre + "batch_ids": captured["states"], pre + "y": captured["y"],
                    pre + "cfg": np.asarray([A, B, n, int(double), 31 + case])})
    np.savez(os.path.join(OUT, "dqn_target.npz"), **out)
    print("dqn_target.npz", out["c0_y"].shape, out["c0_y"].dtype)


def golden_impala():
    class Algorithm(object):
        pass
    sys.modules["xt.algorithm"].Algorithm = Algorithm
    _load("xt.algorithm.impala.default_config", "xt/algorithm/impala/default_config.py")
    _mod("zeus.common.ipc.uni_comm", UniComm=object)
    _mod("xt.model.tf_compat", loss_to_val=lambda x: x)
    _mod("xt.algorithm.alg_utils", DivideDistPolicy=object, FIFODistPolicy=object, EqualDistPolicy=object)
    ref = _load("xt.algorithm.impala.impala_opt", "xt/algorithm/impala/impala_opt.py")
    ref.BATCH_SIZE = 48
    alg = object.__new__(ref.IMPALAOpt)
    alg.states, alg.behavior_logits, alg.actions, alg.dones, alg.rewards = [], [], [], [], []
    slices = []

    class Actor(object):
        def train(self, state, label):
            slices.append((np.asarray(state).copy(), [np.asarray(l).copy() for l in label]))
            return float(len(slices))
    alg.actor = Actor()
    rng = np.random.default_rng(5)
    total = 0
    for msg in range(5):
        n = 32   # vector_env_size(2) x 16 steps
        ids = np.arange(total, total + n)
        total += n
        alg.prepare_data(dict(cur_state=ids.astype(np.float32), logit=np.stack([ids, ids], 1).astype(np.float32), action=ids.astype(np.int32),
                              done=[bool(x) for x in (ids % 7 == 0)], reward=[float(x) for x in ids * 0.5]))
    mean_loss = alg.train()
    out = {"batch_size": np.asarray(48), "n_slices": np.asarray(len(slices)), "mean_loss": np.asarray(mean_loss),
           "slice_sizes": np.asarray([len(s[0]) for s in slices]), "state_order": np.concatenate([s[0] for s in slices]),
           "done_dtype_is_bool": np.asarray(slices[0][1][2].dtype == np.bool_), "done": np.concatenate([s[1][2] for s in slices]),
           "reward": np.concatenate([s[1][3] for s in slices])}
    np.savez(os.path.join(OUT, "impala_proc.npz"), **out)
    print("impala_proc.npz", out["slice_sizes"])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    install_stubs()
    golden_gae()
    golden_ppo_minibatch()
    golden_dqn()
    golden_impala()
