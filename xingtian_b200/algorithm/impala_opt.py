"""IMPALAOpt algorithm (xt/algorithm/impala/impala_opt.py:37-147)."""
import os

import numpy as np
import torch

from ..registry import Registers, import_config
from .base import Algorithm, FIFODistPolicy

BATCH_SIZE = 200   # xt/algorithm/impala/default_config.py


@Registers.algorithm
class IMPALAOpt(Algorithm):
    """Buffers (state, behaviour logits, action, done, reward) trajectories on the device and trains
    in BATCH_SIZE slices; V-trace runs inside the train step."""

    def __init__(self, model_info, alg_config, **kwargs):
        import_config(globals(), alg_config)
        super().__init__(alg_name="impala", model_info=model_info["actor"], alg_config=alg_config)
        self._chunks = []
        self.async_flag = False
        self.dist_model_policy = FIFODistPolicy(alg_config["instance_num"], prepare_times=self._prepare_times_per_train)

    @staticmethod
    def _data_proc(episode_data):
        """impala_opt.py:125-147."""
        states = episode_data["cur_state"]
        behavior_logits = episode_data["logit"]
        actions = episode_data["action"]
        dones = np.asarray(episode_data["done"], dtype=np.bool_)
        rewards = np.asarray(episode_data["reward"])
        return states, behavior_logits, actions, dones, rewards

    def prepare_data(self, train_data, **kwargs):
        """impala_opt.py:110-117 -- staged to the device immediately."""
        state, logit, action, done, reward = self._data_proc(train_data)
        dev = self.actor.device
        st = np.ascontiguousarray(state, np.uint8)
        t = torch.from_numpy(st)
        self._chunks.append(dict(
            obs=(t.pin_memory() if st.nbytes >= (1 << 20) else t).to(dev, non_blocking=True),
            bp=torch.from_numpy(np.ascontiguousarray(logit, np.float32)).to(dev),
            action=torch.from_numpy(np.ascontiguousarray(action, np.int32).reshape(-1)).to(dev),
            done=torch.from_numpy(np.ascontiguousarray(done, np.bool_).reshape(-1).view(np.uint8)).to(dev),
            reward=torch.from_numpy(np.ascontiguousarray(reward, np.float32).reshape(-1)).to(dev)))

    def train(self, **kwargs):
        """impala_opt.py:73-106: concatenate, slice by BATCH_SIZE, one SGD step per slice, mean loss."""
        cat = {k: torch.cat([c[k] for c in self._chunks]) if len(self._chunks) > 1 else self._chunks[0][k]
               for k in ("obs", "bp", "action", "done", "reward")}
        nbatch = cat["obs"].shape[0]
        count = (nbatch + BATCH_SIZE - 1) // BATCH_SIZE
        losses = torch.zeros(count, dtype=torch.float32, device=self.actor.device)
        tmp = torch.zeros(1, dtype=torch.float32, device=self.actor.device)
        for i in range(count):
            s, e = i * BATCH_SIZE, min(nbatch, (i + 1) * BATCH_SIZE)
            self.actor.train_device(cat["obs"][s:e], cat["bp"][s:e], cat["action"][s:e], cat["done"][s:e],
                                    cat["reward"][s:e], e - s, tmp)
            losses[i:i + 1].copy_(tmp)
        self._chunks.clear()
        return float(losses.mean().cpu())

    def save(self, model_path, model_index):
        """impala_opt.py:103-108."""
        actor_name = "actor" + str(model_index).zfill(5)
        actor_name = self.actor.save_model(os.path.join(model_path, actor_name))
        return [actor_name.split("/")[-1]]

    def predict(self, state):
        """impala_opt.py:119-123."""
        return self.actor.predict(state)
