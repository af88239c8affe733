"""ImpalaCnnOpt on the CUDA engine (xt/model/impala/impala_cnn_opt.py:64-297)."""
import numpy as np
import torch

from .. import capi
from ..capi import check
from ..engine import Adam, Net, _ptr, stream_ptr
from ..registry import Registers, import_config
from . import archs
from .base import XTModel, glorot_uniform_

# xt/model/impala/default_config.py
LR = 0.0003
ENTROPY_LOSS = 0.01
GAMMA = 0.99


@Registers.model
class ImpalaCnnOpt(XTModel):
    """IMPALA conv net with the V-trace loss evaluated inside the train step."""

    def __init__(self, model_info):
        model_config = model_info.get("model_config", dict())
        import_config(globals(), model_config)
        self.input_dtype = model_info.get("input_dtype", "float32")
        self.sta_mean = model_info.get("state_mean", 0.)
        self.sta_std = model_info.get("state_std", 255.)
        self.state_dim = model_info["state_dim"]
        self.action_dim = model_info["action_dim"]
        self.lr_schedule = model_config.get("lr_schedule", None)
        self.opt_type = model_config.get("opt_type", "adam")
        self.lr = LR
        self.grad_norm_clip = model_config.get("grad_norm_clip", 40.0)
        self.sample_batch_steps = model_config.get("sample_batch_step", 50)
        self._init_seed = model_config.get("init_seed")
        self.use_graph = bool(model_config.get("use_cuda_graph", True))
        if self.opt_type not in ("adam", "rmsprop"):
            raise KeyError("invalid opt_type: {}".format(self.opt_type))          # impala_cnn_opt.py:207-208
        if self.lr_schedule and len(self.lr_schedule) != 2:
            raise ValueError("lr_schedule invalid: {}".format(self.lr_schedule))  # impala_cnn_opt.py:237-240 (logging.fatal)
        if self.input_dtype not in ("uint8",) or abs(self.sta_mean) >= 1e-4:
            # state_transform (model_utils.py:192-201): uint8 with mean~0 => x/std
            raise NotImplementedError("ImpalaCnnOpt: only uint8 observations with state_mean 0 are supported")
        super().__init__(model_info)

    def create_model(self, model_info):
        arch = archs.impala_cnn(self.state_dim, self.action_dim)
        arch["scale"] = 1.0 / float(self.sta_std)
        self.arch = arch
        self.net = Net(arch, max_batch=int(model_info.get("max_batch", 512)), device=self.device)
        rng = np.random.default_rng(self._init_seed) if self._init_seed is not None else np.random.default_rng()
        glorot_uniform_(self.net, rng)
        # baseline head: custom_norm_initializer(0.01) (model_utils.py:204-211, impala_cnn_opt.py:146)
        name = "explore_agent/dense/kernel"
        shape = self.net.ptable[name][1]
        o = rng.standard_normal(shape).astype(np.float32)
        o *= 0.01 / np.sqrt(np.square(o).sum(axis=0, keepdims=True))
        self.net.view(name).copy_(torch.from_numpy(o))
        self.net.params_changed()
        self.opt = Adam(self.net, self.lr, eps=1e-8, clip_mode=capi.CLIP_GLOBAL_NORM, clip=self.grad_norm_clip)
        if self.opt_type == "rmsprop":
            self.opt.use_rmsprop(decay=0.99, epsilon=0.1)          # impala_cnn_opt.py:205-206 (the schedule is Adam-only there)
        self._global_step = 0
        self._bufs = {}
        self._sample_seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._sample_offset = 0
        self.logit_name, self.base_name = arch["outputs"]
        return self.net

    def _buffers(self, n):
        b = self._bufs.get(n)
        if b is None:
            dev = self.device
            b = dict(obs=torch.empty((n,) + tuple(self.state_dim), dtype=torch.uint8, device=dev),
                     bp=torch.empty(n, self.action_dim, dtype=torch.float32, device=dev),
                     action=torch.empty(n, dtype=torch.int32, device=dev),
                     done=torch.empty(n, dtype=torch.uint8, device=dev),
                     reward=torch.empty(n, dtype=torch.float32, device=dev),
                     logp=torch.empty(n, dtype=torch.float32, device=dev),
                     loss=torch.zeros(1, dtype=torch.float32, device=dev))
            self._bufs[n] = b
        return b

    def train_device(self, obs, bp_logits, actions, dones, rewards, n, loss_buf):
        """One V-trace SGD step on device-resident tensors ([n] env-major, n = k*sample_batch_step)."""
        net = self.net
        s = int(self.sample_batch_steps)
        if n % s != 0 or n < s:
            raise ValueError("batch of %d samples is not a whole number of %d-step trajectories" % (n, s))
        net.ensure_batch(n)
        loss_buf.zero_()
        if self.lr_schedule and self.opt_type == "adam":
            self.opt.set_lr(self.scheduled_lr(self._global_step))
        self._global_step += 1
        # forward, in-graph V-trace + losses, backward, clip + Adam: one native call, replayed as a CUDA graph
        check(net.lib.xtb_impala_train(net.handle, self.opt.handle, _ptr(obs), None, _ptr(bp_logits), _ptr(actions), _ptr(dones),
                                       _ptr(rewards), int(n), s, float(GAMMA), net.tid[self.logit_name], net.tid[self.base_name],
                                       _ptr(loss_buf), 1 if self.use_graph else 0, stream_ptr()))
        return loss_buf

    def scheduled_lr(self, global_step, decay_step=20000.0):
        """impala_cnn_opt.py:234-249: tf.train.linear_cosine_decay(lr_schedule[0][1], global_step, 20000,
        beta=lr_schedule[1][1] / 20000) with the TF defaults num_periods=0.5, alpha=0:
        lr * ((alpha + (D - s) / D) * 0.5 * (1 + cos(pi * 2 * num_periods * s / D)) + beta), s = min(global_step, D)."""
        import math
        base, beta = float(self.lr_schedule[0][1]), float(self.lr_schedule[1][1]) / float(decay_step)
        s = min(float(global_step), float(decay_step))
        linear = (decay_step - s) / decay_step
        cosine = 0.5 * (1.0 + math.cos(math.pi * 2.0 * 0.5 * s / decay_step))
        return base * (linear * cosine + beta)

    def train(self, state, label):
        """impala_cnn_opt.py:251-265: train(state, [bp_logic_outs, actions, dones, rewards]) -> loss."""
        bp_logic_outs, actions, dones, rewards = label
        n = len(state)
        b = self._buffers(n)
        b["obs"].copy_(torch.from_numpy(np.ascontiguousarray(state, np.uint8)), non_blocking=True)
        b["bp"].copy_(torch.from_numpy(np.ascontiguousarray(bp_logic_outs, np.float32)))
        b["action"].copy_(torch.from_numpy(np.ascontiguousarray(actions, np.int32).reshape(-1)))
        b["done"].copy_(torch.from_numpy(np.ascontiguousarray(dones, np.bool_).reshape(-1).view(np.uint8)))
        b["reward"].copy_(torch.from_numpy(np.ascontiguousarray(rewards, np.float32).reshape(-1)))
        loss = self.train_device(b["obs"], b["bp"], b["action"], b["done"], b["reward"], n, b["loss"])
        return float(loss.cpu()[0])

    def _predict_io(self, n):
        """Persistent staging of the host-facing predict(): device frames, packed device / pinned result blocks."""
        io = self._bufs.get(("io", n))
        if io is None:
            dev = self.device
            out_dev = torch.empty(3, n, dtype=torch.float32, device=dev)
            io = dict(obs=torch.empty((n,) + tuple(self.state_dim), dtype=torch.uint8, device=dev), out_dev=out_dev,
                      pin_out=torch.empty(3, n, dtype=torch.float32).pin_memory(),
                      pin_logits=torch.empty(n, self.action_dim, dtype=torch.float32).pin_memory())
            io["out_np"], io["logits_np"] = io["pin_out"].numpy(), io["pin_logits"].numpy()
            self._bufs[("io", n)] = io
        return io

    def rollout_infer_device(self, obs_dev, step_idx, n_env, n_step, action, logp, value):
        """n_step batched policy evaluations on device-resident frames as ONE CUDA graph (the learner-side replacement of
        the explorers' per-step predict calls): time-major action / logp / baseline [n_step, n_env]; the logits of the last
        step stay in net.tensor(logit_name)."""
        import ctypes as C
        if getattr(self, "_offset_dev", None) is None:
            self._offset_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.net.ensure_batch(n_env)
        check(self.net.lib.xtb_ppo_rollout_infer(self.net.handle, _ptr(obs_dev), _ptr(step_idx), int(n_env), int(n_step),
                                                 self.net.tid[self.logit_name], self.net.tid[self.base_name],
                                                 C.c_uint64(self._sample_seed), _ptr(self._offset_dev), _ptr(action), _ptr(logp),
                                                 _ptr(value), 1 if self.use_graph else 0, stream_ptr()))

    def predict(self, state, uniforms=None):
        """impala_cnn_opt.py:267-277: [logits [B,A], baseline [B], action [B]]."""
        state = np.ascontiguousarray(state, np.uint8)
        n = state.shape[0]
        net = self.net
        if uniforms is None and n <= net.max_batch:
            # staged H2D -> graphed forward + fused heads + Philox sampling -> packed D2H (+ logits) -> sync: one native call
            import ctypes as C
            io = self._predict_io(n)
            if getattr(self, "_offset_dev", None) is None:
                self._offset_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            net.ensure_batch(n)
            check(net.lib.xtb_actor_predict_host(net.handle, state.ctypes.data, state.nbytes, _ptr(io["obs"]), n,
                                                 net.tid[self.logit_name], net.tid[self.base_name], C.c_uint64(self._sample_seed),
                                                 _ptr(self._offset_dev), _ptr(io["out_dev"]), _ptr(io["pin_out"]),
                                                 _ptr(io["pin_logits"]), 1 if self.use_graph else 0, stream_ptr()))
            out = io["out_np"]
            return [io["logits_np"].copy(), out[2].copy(), out[0].view(np.int32).copy()]
        b = self._buffers(n)
        net.ensure_batch(n)
        b["obs"].copy_(torch.from_numpy(state), non_blocking=True)
        net.forward(b["obs"], n)
        u = None
        if uniforms is not None:
            u = torch.from_numpy(np.ascontiguousarray(uniforms, np.float32)).to(self.device)
        import ctypes as C
        check(net.lib.xtb_categorical_sample(_ptr(net.tensor(self.logit_name)), n, self.action_dim, _ptr(u),
                                             C.c_uint64(self._sample_seed), C.c_uint64(self._sample_offset),
                                             _ptr(b["action"]), _ptr(b["logp"]), stream_ptr()))
        self._sample_offset += 1
        logits = net.tensor(self.logit_name)[:n].cpu().numpy()
        base = net.tensor(self.base_name)[:n, 0].cpu().numpy()
        return [logits, base, b["action"].cpu().numpy()]
