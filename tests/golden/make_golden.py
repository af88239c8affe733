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

and, with a small numpy stand-in for the dozen TensorFlow ops they call (class TFShim below: eager float32, the documented
TF-1.15 semantics of exp/log/min/max/clip/reduce_*/one_hot/softmax-xent/scan(reverse)), the reference's loss code itself:

  tf_losses.npz      CategoricalDist.neglog_prob/entropy      (xt/model/tf_dist.py:89-113)
                     actor_loss_with_entropy, critic_loss     (xt/model/ppo/__init__.py:4-25)
                     vtrace.from_logic_outputs                (xt/model/impala/vtrace.py:39-115)
                     vtrace_loss and its three terms          (xt/model/impala/impala_cnn_opt.py:299-351)
  This pins the structure of those functions (indexing, clipping, bootstrap handling, sums vs means); the TF kernels
  behind the individual ops stay unpinned.

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


# ------------------------------------------------------------------------------------------------------------------
# numpy stand-in for the TensorFlow ops used by the reference's loss / distribution / V-trace code
# ------------------------------------------------------------------------------------------------------------------
class _NN(object):
    @staticmethod
    def log_softmax(x, axis=-1):
        x = np.asarray(x, np.float32)
        m = x.max(axis=axis, keepdims=True)
        return (x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True), dtype=np.float32)).astype(np.float32)

    @staticmethod
    def softmax(x, axis=-1):
        return np.exp(_NN.log_softmax(x, axis)).astype(np.float32)

    @staticmethod
    def softmax_cross_entropy_with_logits_v2(labels, logits):
        return (-(np.asarray(labels, np.float32) * _NN.log_softmax(logits)).sum(-1)).astype(np.float32)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels, logits):
        ls = _NN.log_softmax(logits)
        return (-np.take_along_axis(ls, np.asarray(labels, np.int64)[..., None], -1)[..., 0]).astype(np.float32)


class _Shape(tuple):
    def assert_has_rank(self, rank):
        assert len(self) == rank, (tuple(self), rank)

    def as_list(self):
        return list(self)


class _Tensor(np.ndarray):
    """ndarray whose .shape answers TensorShape.assert_has_rank (vtrace.py:67-69)."""

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))


class TFShim(object):
    float32, int32 = np.float32, np.int32
    nn = _NN

    class _Ctx(object):
        def __enter__(self): return self
        def __exit__(self, *a): return False

    @staticmethod
    def device(name): return TFShim._Ctx()
    @staticmethod
    def convert_to_tensor(x, dtype=None): return np.asarray(x, dtype).view(_Tensor)
    @staticmethod
    def exp(x): return np.exp(np.asarray(x, np.float32)).astype(np.float32)
    @staticmethod
    def log(x): return np.log(np.asarray(x, np.float32)).astype(np.float32)
    @staticmethod
    def square(x): return np.square(np.asarray(x, np.float32))
    @staticmethod
    def add(a, b): return np.asarray(a, np.float32) + np.asarray(b, np.float32)
    @staticmethod
    def minimum(a, b): return np.minimum(np.float32(a) if np.isscalar(a) else a, np.float32(b) if np.isscalar(b) else b)
    @staticmethod
    def maximum(a, b): return np.maximum(np.float32(a) if np.isscalar(a) else a, np.float32(b) if np.isscalar(b) else b)
    @staticmethod
    def clip_by_value(x, lo, hi): return np.clip(np.asarray(x, np.float32), np.float32(lo), np.float32(hi))
    @staticmethod
    def reduce_mean(x, axis=None, keepdims=False): return np.mean(np.asarray(x, np.float32), axis=axis, keepdims=keepdims, dtype=np.float32)
    @staticmethod
    def reduce_sum(x, axis=None, keepdims=False): return np.sum(np.asarray(x, np.float32), axis=axis, keepdims=keepdims, dtype=np.float32)
    @staticmethod
    def reduce_max(x, axis=None, keepdims=False): return np.max(np.asarray(x, np.float32), axis=axis, keepdims=keepdims)
    @staticmethod
    def one_hot(x, depth): return np.eye(depth, dtype=np.float32)[np.asarray(x, np.int64)]
    @staticmethod
    def expand_dims(x, axis): return np.expand_dims(x, axis)
    @staticmethod
    def concat(xs, axis=0): return np.concatenate(xs, axis=axis)
    @staticmethod
    def zeros_like(x): return np.zeros_like(x)
    @staticmethod
    def stop_gradient(x): return x

    @staticmethod
    def scan(fn, elems, initializer, parallel_iterations=None, back_prop=None, reverse=False):
        """tf.scan over the leading axis of a tuple of tensors; reverse=True walks from the last element to the first
        and places each result at the index of the element it consumed."""
        n = elems[0].shape[0]
        out = [None] * n
        acc = initializer
        for i in (range(n - 1, -1, -1) if reverse else range(n)):
            acc = fn(acc, tuple(e[i] for e in elems))
            out[i] = acc
        return np.stack(out, 0).astype(np.float32)


class _AnyName(types.ModuleType):
    """xt.model.tf_compat stand-in: `tf` is the shim, every other imported name (Conv2D, Saver, ...) is an inert placeholder."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def golden_tf_losses():
    compat = _AnyName("xt.model.tf_compat")
    compat.tf = TFShim
    compat.DTYPE_MAP = {"float32": np.float32}
    sys.modules["xt.model.tf_compat"] = compat
    dist_mod = _load("xt.model.tf_dist", "xt/model/tf_dist.py")
    ppo_mod = _load("xt.model.ppo", "xt/model/ppo/__init__.py")
    _mod("xt.model.impala").__path__ = []
    vtrace = _load("xt.model.impala.vtrace", "xt/model/impala/vtrace.py")
    sys.modules["xt.model.impala"].vtrace = vtrace
    # imports of impala_cnn_opt.py that are irrelevant to its loss functions
    for name in ("tensorflow", "tensorflow.python", "tensorflow.python.util"):
        _mod(name).__path__ = []
    _mod("tensorflow.python.util.deprecation", _PRINT_DEPRECATION_WARNINGS=False)
    sys.modules["tensorflow.python.util"].deprecation = sys.modules["tensorflow.python.util.deprecation"]
    sys.modules["xt.model"].XTModel = type("XTModel", (), {})
    _load("xt.model.impala.default_config", "xt/model/impala/default_config.py")
    _mod("xt.model.atari_model", get_atari_filter=None)
    _mod("xt.model.tf_utils", TFVariables=None, restore_tf_variable=None)
    _mod("xt.model.model_utils", state_transform=None, custom_norm_initializer=None)
    imp = _load("xt.model.impala.impala_cnn_opt", "xt/model/impala/impala_cnn_opt.py")

    rng = np.random.default_rng(77)
    out = {}
    # --- categorical distribution + PPO losses (shapes of xt/model/ppo/ppo.py:62-92: actions [B], the rest [B,1])
    for case, (B, A, scale) in enumerate([(5, 4, 1.0), (320, 4, 3.0), (257, 6, 30.0), (1, 2, 0.1)]):
        logits = (rng.standard_normal((B, A)) * scale).astype(np.float32)
        act = rng.integers(0, A, B).astype(np.int32)
        adv = rng.standard_normal((B, 1)).astype(np.float32)
        old_logp = (-np.abs(rng.standard_normal((B, 1))) - 0.3).astype(np.float32)
        old_v, target_v = rng.standard_normal((B, 1)).astype(np.float32), rng.standard_normal((B, 1)).astype(np.float32)
        out_v = (old_v + 4.0 * rng.standard_normal((B, 1))).astype(np.float32)
        dist = dist_mod.CategoricalDist(A)
        dist.init_by_param(logits)
        pre = "ppo%d_" % case
        out[pre + "logits"], out[pre + "action"], out[pre + "adv"], out[pre + "old_logp"] = logits, act, adv, old_logp
        out[pre + "old_v"], out[pre + "target_v"], out[pre + "out_v"] = old_v, target_v, out_v
        out[pre + "neglogp"] = dist.neglog_prob(act)
        out[pre + "entropy"] = dist.entropy()
        for clip, ent in ((0.1, 0.003), (0.2, 0.01)):
            out[pre + "actor_loss_%g_%g" % (clip, ent)] = np.float32(ppo_mod.actor_loss_with_entropy(dist, adv, old_logp, act, clip, ent))
        for vclip in (5.0, 0.5):
            out[pre + "critic_loss_%g" % vclip] = np.float32(ppo_mod.critic_loss(target_v, out_v, old_v, vclip))
    # --- V-trace (time-major [T', B, ...]) and the IMPALA loss terms
    for case, (T, B, A) in enumerate([(7, 1, 4), (127, 4, 4), (19, 3, 6)]):
        bp = rng.standard_normal((T, B, A)).astype(np.float32)
        tp = (bp + 0.5 * rng.standard_normal((T, B, A))).astype(np.float32)
        act = rng.integers(0, A, (T, B)).astype(np.int32)
        disc = ((rng.random((T, B)) > 0.05) * 0.99).astype(np.float32)
        rew = np.clip(rng.standard_normal((T, B)), -1, 1).astype(np.float32)
        val = rng.standard_normal((T, B)).astype(np.float32)
        boot = rng.standard_normal(B).astype(np.float32)
        vs, pg = vtrace.from_logic_outputs(bp, tp, act, disc, rew, val, boot)
        pre = "vt%d_" % case
        for k, v in dict(bp=bp, tp=tp, action=act, discount=disc, reward=rew, value=val, bootstrap=boot, vs=vs, pg_adv=pg).items():
            out[pre + k] = v
        out[pre + "pi_loss"] = np.float32(imp.calc_pi_loss(tp, act, pg))
        out[pre + "baseline_loss"] = np.float32(imp.calc_baseline_loss(vs - val))
        out[pre + "entropy_loss"] = np.float32(imp.calc_entropy_loss(tp))
        out[pre + "total_loss"] = np.float32(imp.vtrace_loss(bp, tp, act, disc, rew, val, boot))
    np.savez(os.path.join(OUT, "tf_losses.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("tf_losses.npz", len(out), "arrays; vt1 total", out["vt1_total_loss"], "ppo1 actor", out["ppo1_actor_loss_0.1_0.003"])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    install_stubs()
    golden_gae()
    golden_ppo_minibatch()
    golden_dqn()
    golden_impala()
    golden_tf_losses()
