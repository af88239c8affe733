"""Algorithm base class and weight-distribution policies: the surface TrainWorker drives
(xt/algorithm/algorithm.py:34-237, xt/algorithm/alg_utils.py:25-117).

When the reference package is importable (the drop-in case: this package loaded under xt_main), `Algorithm` SUBCLASSES the
reference's own `xt.algorithm.algorithm.Algorithm` and only replaces construction (the model comes from this package's
registry, not from xt.model), so every method TrainWorker calls is the reference's code.  Stand-alone (tests, bench, the
GPU box: no `xt` on the path) the small `_StandaloneSurface` below provides the same methods.

_StandaloneSurface restates an interface of XingTian (MIT License):
    Copyright (C) 2020. Huawei Technologies Co., Ltd. All rights reserved.
    Permission is hereby granted, free of charge, to any person obtaining a copy of this software and associated
    documentation files (the "Software"), to deal in the Software without restriction, including without limitation the
    rights to use, copy, modify, merge, publish, distribute, sublicense, and/or sell copies of the Software, and to permit
    persons to whom the Software is furnished to do so, subject to the following conditions: The above copyright notice and
    this permission notice shall be included in all copies or substantial portions of the Software.
    THE SOFTWARE IS PROVIDED "AS IS", WITHOUT WARRANTY OF ANY KIND, EXPRESS OR IMPLIED, INCLUDING BUT NOT LIMITED TO THE
    WARRANTIES OF MERCHANTABILITY, FITNESS FOR A PARTICULAR PURPOSE AND NONINFRINGEMENT. IN NO EVENT SHALL THE AUTHORS OR
    COPYRIGHT HOLDERS BE LIABLE FOR ANY CLAIM, DAMAGES OR OTHER LIABILITY, WHETHER IN AN ACTION OF CONTRACT, TORT OR
    OTHERWISE, ARISING FROM, OUT OF OR IN CONNECTION WITH THE SOFTWARE OR THE USE OR OTHER DEALINGS IN THE SOFTWARE.
"""
import importlib.util
import logging
import os
from collections import defaultdict, deque

import numpy as np

from ..registry import model_builder

MODEL_PREFIX = "actor"
ZFILL_LENGTH = 5


def _clip_explorer_id(raw_dist_info, clip_set):
    if not clip_set:
        return raw_dist_info
    if isinstance(raw_dist_info, dict):
        if raw_dist_info["explorer_id"] == -1:
            raw_dist_info["explorer_id"] = clip_set
        else:
            raw_dist_info["explorer_id"] = [i for i in raw_dist_info["explorer_id"] if i in clip_set]
    return raw_dist_info


class DefaultAlgDistPolicy(object):
    """alg_utils.py:37-47: broadcast to every explorer."""

    def __init__(self, actor_num, **kwargs):
        self.actor_num = actor_num
        self.default_policy = {"broker_id": -1, "explorer_id": -1}

    def get_dist_info(self, model_index, explorer_set=None):
        return _clip_explorer_id(self.default_policy, explorer_set)

    def add_processed_ctr_info(self, ctr_info):
        pass


class FIFODistPolicy(DefaultAlgDistPolicy):
    """alg_utils.py:70-94: send weights to the explorers whose data was just consumed."""

    def __init__(self, actor_num, prepare_times, **kwargs):
        super().__init__(actor_num, **kwargs)
        self._processed_agent = deque()
        self.prepare_data_times = prepare_times

    def add_processed_ctr_info(self, ctr_info):
        self._processed_agent.append(ctr_info)

    def get_dist_info(self, model_index, explorer_set=None):
        if model_index < 0:
            return self.default_policy
        relation = defaultdict(set)
        while self._processed_agent:
            info = self._processed_agent.popleft()
            relation[info[0]].update((info[1],))
        out = [{"broker_id": b, "explorer_id": list(e)} for b, e in relation.items()]
        return _clip_explorer_id(out, explorer_set)


def _reference_algorithm():
    """The reference's Algorithm class if its package is on the path (never imported speculatively: find_spec first)."""
    if os.environ.get("XTB_STANDALONE_BASE") == "1":
        return None
    try:
        if importlib.util.find_spec("xt") is None or importlib.util.find_spec("xt.algorithm") is None:
            return None
        from xt.algorithm.algorithm import Algorithm as ref_cls
        return ref_cls
    except Exception as err:       # half-installed reference (missing TF, absl ...): stand alone
        logging.debug("reference Algorithm not importable: %s", err)
        return None


# alg_config key -> (attribute, default); algorithm.py:63-81
_CONFIG_ATTRS = (("learning_starts", "learning_starts", 0), ("train_per_checkpoint", "_train_per_checkpoint", 1),
                 ("save_model", "if_save_model", False), ("save_interval", "save_interval", 500))


class _StandaloneSurface(object):
    """The methods TrainWorker / the agents call on an algorithm (algorithm.py:83-237), for use without the reference."""

    buff = None
    actor = None

    # -- checkpoint cadence ------------------------------------------------------------------------------------------
    def if_save(self, train_count):
        if not self.if_save_model:
            return False
        return True if train_count % self.save_interval == 0 else None      # None, as algorithm.py:83-88 falls through

    def checkpoint_ready(self, train_count, **kwargs):
        self._train_ready = False
        return train_count % self.train_per_checkpoint == 0

    train_per_checkpoint = property(lambda self: self._train_per_checkpoint,
                                    lambda self, interval: setattr(self, "_train_per_checkpoint", interval))
    prepare_data_times = property(lambda self: self._prepare_times_per_train)
    weights_map = property(lambda self: self._weights_map, lambda self, info: setattr(self, "_weights_map", info))

    @staticmethod
    def update_weights_map(agent_in_group="agent_0", agent_in_env="agent_0"):
        return {}

    # -- data / training hooks ---------------------------------------------------------------------------------------
    def prepare_data(self, train_data, **kwargs):
        raise NotImplementedError

    def train(self, **kwargs):
        raise NotImplementedError

    def train_ready(self, elapsed_episode, **kwargs):
        short = bool(getattr(self, "buff")) and self.learning_starts > 0 and self.buff.size() < self.learning_starts
        self._train_ready = not short
        return self._train_ready

    def predict(self, state):
        """Greedy action of one state (algorithm.py:124-135)."""
        return np.argmax(self.actor.predict(state.reshape((1,) + state.shape)))

    # -- weights / checkpoints ---------------------------------------------------------------------------------------
    def save(self, model_path, model_index):
        stem = "{}_{}".format(MODEL_PREFIX, str(model_index).zfill(ZFILL_LENGTH))
        return [self.actor.save_model(os.path.join(model_path, stem))]

    def restore(self, model_name=None, model_weights=None):
        if model_weights is None:
            logging.debug("%s load model: %s", self.alg_name, model_name)
            self.actor.load_model(model_name)
        else:
            self.actor.set_weights(model_weights)

    def get_weights(self):
        return self.actor.get_weights()

    def set_weights(self, weights):
        return self.actor.set_weights(weights)

    def shutdown(self):
        pass


_REF = _reference_algorithm()


class Algorithm(_REF if _REF is not None else _StandaloneSurface):
    """xt/algorithm/algorithm.py:34-237 with the model built by this package's registry (CUDA models)."""

    inherits_reference = _REF is not None

    def __init__(self, alg_name, model_info, alg_config=None, **kwargs):
        # construction is always ours: the reference's __init__ (algorithm.py:44-81) would call xt.model.model_builder
        alg_config = alg_config if alg_config is not None else {}
        self.actor = model_builder(model_info)
        self.state_dim, self.action_dim = model_info.get("state_dim"), model_info.get("action_dim")
        self.alg_name, self.alg_config, self.model_info = alg_name, alg_config, model_info
        self.train_count = 0
        self.async_flag = True
        self._train_ready = True
        self._weights_map = self.update_weights_map()
        per_train = alg_config["instance_num"] * alg_config["agent_num"]
        self._prepare_times_per_train = alg_config.get("prepare_times_per_train", per_train)
        self.dist_model_policy = DefaultAlgDistPolicy(alg_config["instance_num"], prepare_times=self._prepare_times_per_train)
        for key, attr, default in _CONFIG_ATTRS:
            setattr(self, attr, alg_config.get(key, default))
